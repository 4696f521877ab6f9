/*
 * oracle/ozz_archive.c -- TEST INFRASTRUCTURE (see dmk_oracle.h).
 *
 * Restates ozz::io::IArchive >> ozz::animation::Skeleton / Animation as darmok drives them in
 * OzzUtils::loadFromData (src/skeleton_ozz.cpp:27-44): TestTag<T>() then `archive >> obj`.
 * Byte layout per SURVEY.md Appendix A.1-A.3 (Skeleton v2; Animation v6 = ozz-animation 0.14 and v7 = ozz-animation >= 0.15).
 * ozz is not vendored in /root/reference -> parity unpinned (header of dmk_oracle.h).
 */
#include "dmk_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    const uint8_t* p;
    size_t n, pos;
    int swap; /* archive endianness differs from host */
    int fail;
} rd_t;

static void rd_bytes(rd_t* r, void* out, size_t n) {
    /* DataOzzStream::Read clamps to the bytes available (src/skeleton_ozz.cpp:147-160); a short
     * read leaves the rest of the destination untouched and we flag it. */
    if (r->pos + n > r->n) {
        r->fail = 1;
        n = r->n - r->pos;
    }
    memcpy(out, r->p + r->pos, n);
    r->pos += n;
}
static void swap_n(void* v, size_t n) {
    uint8_t* b = (uint8_t*)v;
    for (size_t i = 0; i < n / 2; ++i) {
        uint8_t t = b[i];
        b[i] = b[n - 1 - i];
        b[n - 1 - i] = t;
    }
}
#define RD_SCALAR(name, type)                 \
    static type name(rd_t* r) {               \
        type v = 0;                           \
        rd_bytes(r, &v, sizeof v);            \
        if (r->swap) swap_n(&v, sizeof v);    \
        return v;                             \
    }
RD_SCALAR(rd_u8, uint8_t)
RD_SCALAR(rd_u16, uint16_t)
RD_SCALAR(rd_i16, int16_t)
RD_SCALAR(rd_u32, uint32_t)
RD_SCALAR(rd_i32, int32_t)
RD_SCALAR(rd_f32, float)

static void set_err(char* err, size_t errn, const char* msg) {
    if (err && errn) snprintf(err, errn, "%s", msg);
}

/* IArchive ctor reads the endianness byte; TestTag<T> compares the NUL-terminated tag and rewinds;
 * operator>> re-reads the tag and then the uint32 version. */
static int open_object(rd_t* r, const uint8_t* bytes, size_t n, const char* tag, uint32_t* version) {
    r->p = bytes;
    r->n = n;
    r->pos = 0;
    r->fail = 0;
    if (n < 1) return 0;
    uint8_t endian = rd_u8(r); /* ozz::Endianness: 0 = big, 1 = little */
    const uint16_t probe = 1;
    uint8_t host_little = *(const uint8_t*)&probe;
    r->swap = (endian != host_little);
    size_t tl = strlen(tag) + 1;
    if (r->pos + tl > n || memcmp(bytes + r->pos, tag, tl) != 0) return 0;
    r->pos += tl;
    *version = rd_u32(r);
    return !r->fail;
}

orc_skeleton* orc_skeleton_load(const uint8_t* bytes, size_t n, char* err, size_t errn) {
    rd_t r;
    uint32_t version = 0;
    if (!open_object(&r, bytes, n, "ozz-skeleton", &version)) {
        set_err(err, errn, "archive does not contain the type"); /* src/skeleton_ozz.cpp:39 */
        return NULL;
    }
    if (version != 2) {
        set_err(err, errn, "unsupported ozz skeleton version");
        return NULL;
    }
    orc_skeleton* s = (orc_skeleton*)calloc(1, sizeof *s);
    int32_t num_joints = rd_i32(&r);
    if (num_joints < 0 || num_joints > ORC_MAX_JOINTS || r.fail) {
        set_err(err, errn, "corrupt ozz skeleton");
        free(s);
        return NULL;
    }
    s->num_joints = num_joints;
    s->num_soa = (num_joints + 3) / 4;
    if (num_joints == 0) return s;
    int32_t chars = rd_i32(&r);
    if (chars < num_joints || (size_t)chars > n) {
        set_err(err, errn, "corrupt ozz skeleton");
        free(s);
        return NULL;
    }
    s->names_len = chars;
    s->names_buf = (char*)calloc((size_t)chars + 1, 1);
    rd_bytes(&r, s->names_buf, (size_t)chars);
    s->names = (char**)calloc((size_t)num_joints, sizeof(char*));
    {
        char* c = s->names_buf;
        for (int i = 0; i < num_joints; ++i) {
            s->names[i] = c;
            c += strlen(c) + 1;
            if (c > s->names_buf + chars) c = s->names_buf + chars;
        }
    }
    s->parents = (int16_t*)calloc((size_t)num_joints, sizeof(int16_t));
    for (int i = 0; i < num_joints; ++i) s->parents[i] = rd_i16(&r);
    s->rest = (float*)calloc((size_t)s->num_soa * ORC_SOA_FLOATS, sizeof(float));
    for (int i = 0; i < s->num_soa * ORC_SOA_FLOATS; ++i) s->rest[i] = rd_f32(&r);
    if (r.fail) {
        set_err(err, errn, "corrupt ozz skeleton");
        orc_skeleton_free(s);
        return NULL;
    }
    for (int i = 0; i < num_joints; ++i) {
        /* depth-first order invariant that LocalToModelJob relies on (parent index < child index) */
        if (s->parents[i] >= i || s->parents[i] < ORC_NO_PARENT) {
            set_err(err, errn, "corrupt ozz skeleton");
            orc_skeleton_free(s);
            return NULL;
        }
    }
    return s;
}

void orc_skeleton_free(orc_skeleton* s) {
    if (!s) return;
    free(s->names_buf);
    free(s->names);
    free(s->parents);
    free(s->rest);
    free(s);
}
int orc_skeleton_num_joints(const orc_skeleton* s) { return s->num_joints; }
int orc_skeleton_num_soa(const orc_skeleton* s) { return s->num_soa; }
const char* orc_skeleton_joint_name(const orc_skeleton* s, int i) { return s->names[i]; }
void orc_skeleton_parents(const orc_skeleton* s, int16_t* out) {
    memcpy(out, s->parents, (size_t)s->num_joints * sizeof(int16_t));
}
void orc_skeleton_rest_pose(const orc_skeleton* s, float* out) {
    memcpy(out, s->rest, (size_t)s->num_soa * ORC_SOA_FLOATS * sizeof(float));
}
int orc_skeleton_find_joint(const orc_skeleton* s, const char* name) {
    /* SkeletalAnimatorImpl::getJointModelMatrix: first match of a linear scan (src/skeleton_ozz.cpp:968-975) */
    for (int i = 0; i < s->num_joints; ++i)
        if (strcmp(s->names[i], name) == 0) return i;
    return -1;
}

/* ---- Animation archive v6 (ozz-animation <= 0.14): ratio and track in every key ---- */
static int load_animation_v6(rd_t* r, size_t n, orc_animation* a) {
    a->duration = rd_f32(r);
    a->num_tracks = rd_i32(r);
    int32_t name_len = rd_i32(r);
    a->n_t = rd_i32(r);
    a->n_r = rd_i32(r);
    a->n_s = rd_i32(r);
    size_t need = (size_t)(name_len < 0 ? 0 : name_len) + 12u * (size_t)(a->n_t < 0 ? 0 : a->n_t) +
                  14u * (size_t)(a->n_r < 0 ? 0 : a->n_r) + 12u * (size_t)(a->n_s < 0 ? 0 : a->n_s);
    if (r->fail || a->num_tracks < 0 || a->num_tracks > ORC_MAX_JOINTS || name_len < 0 || a->n_t < 0 ||
        a->n_r < 0 || a->n_s < 0 || r->pos + need > n)
        return 0;
    a->num_soa = (a->num_tracks + 3) / 4;
    a->name = (char*)calloc((size_t)name_len + 1, 1);
    rd_bytes(r, a->name, (size_t)name_len);
    a->t = (orc_f3key*)calloc((size_t)a->n_t + 1, sizeof(orc_f3key));
    a->r = (orc_qkey*)calloc((size_t)a->n_r + 1, sizeof(orc_qkey));
    a->s = (orc_f3key*)calloc((size_t)a->n_s + 1, sizeof(orc_f3key));
    for (int i = 0; i < a->n_t; ++i) {
        a->t[i].ratio = rd_f32(r);
        a->t[i].track = rd_u16(r);
        for (int k = 0; k < 3; ++k) a->t[i].value[k] = rd_u16(r);
    }
    for (int i = 0; i < a->n_r; ++i) {
        a->r[i].ratio = rd_f32(r);
        a->r[i].track = rd_u16(r);
        a->r[i].largest = rd_u8(r);
        a->r[i].sign = rd_u8(r);
        for (int k = 0; k < 3; ++k) a->r[i].value[k] = rd_i16(r);
    }
    for (int i = 0; i < a->n_s; ++i) {
        a->s[i].ratio = rd_f32(r);
        a->s[i].track = rd_u16(r);
        for (int k = 0; k < 3; ++k) a->s[i].value[k] = rd_u16(r);
    }
    return !r->fail;
}

/* ---- Animation archive v7 (ozz-animation >= 0.15, SURVEY.md Appendix A.3').  WRITTEN FROM MEMORY of upstream and never
 * checked against a file written by ozz (none exists in the build container): keys no longer carry ratio and track; a shared
 * timepoints[] array, and per component ratios (u8 / u16 index into timepoints), previouses (u16 distance back to the
 * previous key of the same track), iframes (cursor snapshots for seeking: read and ignored) and 3 x u16 values per key.
 * Decoded into the same key lists as v6, so the cursor and the interpolation below do not know the difference, except for
 * the quantisation of a rotation key (orc_qkey.format). ---- */
typedef struct { uint32_t keys, iframe_entries, iframe_desc; } ctrl_counts;

static int read_ctrl_v7(rd_t* r, size_t n, const ctrl_counts* c, const float* timepoints, uint32_t timepoints_count, int padded,
                        float* ratio, uint16_t* track, uint16_t (*value)[3]) {
    const int wide = timepoints_count > 256;
    size_t need = (size_t)c->keys * ((wide ? 2u : 1u) + 2u + 6u) + c->iframe_entries + 4u * (size_t)c->iframe_desc + 4u;
    if (r->pos + need > n) return 0;
    for (uint32_t i = 0; i < c->keys; ++i) {
        uint32_t index = wide ? rd_u16(r) : rd_u8(r);
        if (index >= timepoints_count) return 0;
        ratio[i] = timepoints[index];
    }
    for (uint32_t i = 0; i < c->keys; ++i) {
        uint16_t back = rd_u16(r);
        if (i < (uint32_t)padded) {
            track[i] = (uint16_t)i;
        } else {
            if (back < 1 || back > i) return 0;
            track[i] = track[i - back];
        }
    }
    for (uint32_t i = 0; i < c->iframe_entries; ++i) (void)rd_u8(r);
    for (uint32_t i = 0; i < c->iframe_desc; ++i) (void)rd_u32(r);
    (void)rd_f32(r); /* iframe_interval */
    for (uint32_t i = 0; i < c->keys; ++i)
        for (int k = 0; k < 3; ++k) value[i][k] = rd_u16(r);
    return !r->fail;
}

static int load_animation_v7(rd_t* r, size_t n, orc_animation* a) {
    a->duration = rd_f32(r);
    uint32_t num_tracks = rd_u32(r), name_len = rd_u32(r), timepoints_count = rd_u32(r);
    ctrl_counts t, q, s;
    t.keys = rd_u32(r); q.keys = rd_u32(r); s.keys = rd_u32(r);
    t.iframe_entries = rd_u32(r); t.iframe_desc = rd_u32(r);
    q.iframe_entries = rd_u32(r); q.iframe_desc = rd_u32(r);
    s.iframe_entries = rd_u32(r); s.iframe_desc = rd_u32(r);
    if (r->fail || num_tracks > ORC_MAX_JOINTS || timepoints_count > 65536u || t.keys > n || q.keys > n || s.keys > n ||
        r->pos + (size_t)name_len + 4u * (size_t)timepoints_count > n)
        return 0;
    a->num_tracks = (int)num_tracks;
    a->num_soa = (a->num_tracks + 3) / 4;
    a->n_t = (int)t.keys; a->n_r = (int)q.keys; a->n_s = (int)s.keys;
    a->name = (char*)calloc((size_t)name_len + 1, 1);
    rd_bytes(r, a->name, (size_t)name_len);
    float* timepoints = (float*)calloc((size_t)timepoints_count + 1, sizeof(float));
    for (uint32_t i = 0; i < timepoints_count; ++i) timepoints[i] = rd_f32(r);
    a->t = (orc_f3key*)calloc((size_t)a->n_t + 1, sizeof(orc_f3key));
    a->r = (orc_qkey*)calloc((size_t)a->n_r + 1, sizeof(orc_qkey));
    a->s = (orc_f3key*)calloc((size_t)a->n_s + 1, sizeof(orc_f3key));
    const int padded = a->num_soa * 4;
    size_t most = t.keys > q.keys ? t.keys : q.keys;
    if (s.keys > most) most = s.keys;
    float* ratio = (float*)calloc(most + 1, sizeof(float));
    uint16_t* track = (uint16_t*)calloc(most + 1, sizeof(uint16_t));
    uint16_t (*value)[3] = (uint16_t (*)[3])calloc(most + 1, sizeof(uint16_t[3]));
    int ok = !r->fail;
    if (ok) ok = read_ctrl_v7(r, n, &t, timepoints, timepoints_count, padded, ratio, track, value);
    for (int i = 0; ok && i < a->n_t; ++i) {
        a->t[i].ratio = ratio[i]; a->t[i].track = track[i];
        for (int k = 0; k < 3; ++k) a->t[i].value[k] = value[i][k];
    }
    if (ok) ok = read_ctrl_v7(r, n, &q, timepoints, timepoints_count, padded, ratio, track, value);
    for (int i = 0; ok && i < a->n_r; ++i) {
        const uint64_t packed = (uint64_t)value[i][0] | ((uint64_t)value[i][1] << 16) | ((uint64_t)value[i][2] << 32);
        a->r[i].ratio = ratio[i]; a->r[i].track = track[i];
        a->r[i].largest = (uint8_t)(packed & 3u);
        a->r[i].sign = (uint8_t)((packed >> 2) & 1u);
        a->r[i].value[0] = (int16_t)((packed >> 3) & 0x7fffu);
        a->r[i].value[1] = (int16_t)((packed >> 18) & 0x7fffu);
        a->r[i].value[2] = (int16_t)((packed >> 33) & 0x7fffu);
        a->r[i].format = 1;
    }
    if (ok) ok = read_ctrl_v7(r, n, &s, timepoints, timepoints_count, padded, ratio, track, value);
    for (int i = 0; ok && i < a->n_s; ++i) {
        a->s[i].ratio = ratio[i]; a->s[i].track = track[i];
        for (int k = 0; k < 3; ++k) a->s[i].value[k] = value[i][k];
    }
    free(timepoints); free(ratio); free(track); free(value);
    return ok;
}

orc_animation* orc_animation_load(const uint8_t* bytes, size_t n, char* err, size_t errn) {
    rd_t r;
    uint32_t version = 0;
    if (!open_object(&r, bytes, n, "ozz-animation", &version)) {
        set_err(err, errn, "archive does not contain the type");
        return NULL;
    }
    if (version != 6 && version != 7) {
        set_err(err, errn, "unsupported ozz animation version");
        return NULL;
    }
    orc_animation* a = (orc_animation*)calloc(1, sizeof *a);
    int ok = version == 6 ? load_animation_v6(&r, n, a) : load_animation_v7(&r, n, a);
    int padded = a->num_soa * 4;
    /* the sampling cursor needs 2 keys per padded track and in-range track ids */
    if (ok && a->num_tracks > 0 && (a->n_t < 2 * padded || a->n_r < 2 * padded || a->n_s < 2 * padded)) ok = 0;
    for (int i = 0; ok && i < a->n_t; ++i) ok = a->t[i].track < padded;
    for (int i = 0; ok && i < a->n_r; ++i) ok = a->r[i].track < padded && a->r[i].largest < 4 && a->r[i].sign < 2;
    for (int i = 0; ok && i < a->n_s; ++i) ok = a->s[i].track < padded;
    if (!ok) {
        set_err(err, errn, "corrupt ozz animation");
        orc_animation_free(a);
        return NULL;
    }
    return a;
}

void orc_animation_free(orc_animation* a) {
    if (!a) return;
    free(a->name);
    free(a->t);
    free(a->r);
    free(a->s);
    free(a);
}
float orc_animation_duration(const orc_animation* a) { return a->duration; }
const char* orc_animation_name(const orc_animation* a) { return a->name; }
int orc_animation_num_tracks(const orc_animation* a) { return a->num_tracks; }
int orc_animation_num_soa(const orc_animation* a) { return a->num_soa; }
void orc_animation_key_counts(const orc_animation* a, int* nt, int* nr, int* ns) {
    *nt = a->n_t;
    *nr = a->n_r;
    *ns = a->n_s;
}
