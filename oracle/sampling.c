/*
 * oracle/sampling.c -- TEST INFRASTRUCTURE (see dmk_oracle.h).
 *
 * Restates ozz::animation::SamplingJob::Run + SamplingJob::Context (ozz-animation 0.14,
 * src/animation/runtime/sampling_job.cc) as driven by
 * OzzSkeletalAnimatorAnimationState::update (src/skeleton_ozz.cpp:411-440):
 *     sampling.animation/context/ratio/output ... sampling.Run()
 * SURVEY.md Appendix A.3-A.5.  Scalar simd_ref semantics: RcpEst(x) = 1/x, RSqrtEst(x) = 1/sqrt(x).
 * ozz is not vendored in /root/reference -> parity unpinned.
 */
#include "dmk_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* one SoA group of decompressed bracketing keys (ozz internal::InterpSoaFloat3 / InterpSoaQuaternion) */
typedef struct {
    float ratio[2][4];
    float value[2][3][4];
} interp_f3;
typedef struct {
    float ratio[2][4];
    float value[2][4][4];
} interp_q;

struct orc_sampling_context {
    int max_soa_tracks;
    const orc_animation* animation;
    float ratio;
    int t_cursor, r_cursor, s_cursor;
    int *t_keys, *r_keys, *s_keys;          /* 2 entries (left,right) per padded track */
    uint8_t *t_outdated, *r_outdated, *s_outdated;
    interp_f3* soa_t;
    interp_q* soa_r;
    interp_f3* soa_s;
};

orc_sampling_context* orc_sampling_context_create(int max_tracks) {
    /* SamplingJob::Context::Resize(max_tracks): max_soa_tracks = (max_tracks + 3) / 4 */
    orc_sampling_context* c = (orc_sampling_context*)calloc(1, sizeof *c);
    int soa = (max_tracks + 3) / 4;
    c->max_soa_tracks = soa;
    size_t tracks = (size_t)soa * 4;
    c->t_keys = (int*)calloc(tracks * 2 + 1, sizeof(int));
    c->r_keys = (int*)calloc(tracks * 2 + 1, sizeof(int));
    c->s_keys = (int*)calloc(tracks * 2 + 1, sizeof(int));
    size_t flags = (size_t)(soa + 7) / 8 + 1;
    c->t_outdated = (uint8_t*)calloc(flags, 1);
    c->r_outdated = (uint8_t*)calloc(flags, 1);
    c->s_outdated = (uint8_t*)calloc(flags, 1);
    c->soa_t = (interp_f3*)calloc((size_t)soa + 1, sizeof(interp_f3));
    c->soa_r = (interp_q*)calloc((size_t)soa + 1, sizeof(interp_q));
    c->soa_s = (interp_f3*)calloc((size_t)soa + 1, sizeof(interp_f3));
    orc_sampling_context_invalidate(c);
    return c;
}

void orc_sampling_context_free(orc_sampling_context* c) {
    if (!c) return;
    free(c->t_keys);
    free(c->r_keys);
    free(c->s_keys);
    free(c->t_outdated);
    free(c->r_outdated);
    free(c->s_outdated);
    free(c->soa_t);
    free(c->soa_r);
    free(c->soa_s);
    free(c);
}

void orc_sampling_context_invalidate(orc_sampling_context* c) {
    /* Context::Invalidate */
    c->animation = NULL;
    c->ratio = 0.f;
    c->t_cursor = c->r_cursor = c->s_cursor = 0;
}

/* Context::Step: restart the cursors when the animation changes or time goes backwards */
static void context_step(orc_sampling_context* c, const orc_animation* a, float ratio) {
    if (c->animation != a || ratio < c->ratio) {
        c->animation = a;
        c->t_cursor = c->r_cursor = c->s_cursor = 0;
    }
    c->ratio = ratio;
}

/* exact IEEE binary16 -> binary32 (ozz math::HalfToFloat) */
static float half_to_float(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1fu;
    uint32_t man = h & 0x3ffu;
    uint32_t bits;
    if (exp == 0) {
        if (man == 0) {
            bits = sign;
        } else { /* subnormal half: value = man * 2^-24, exact in binary32 */
            float f = (float)man * 5.9604644775390625e-08f;
            memcpy(&bits, &f, 4);
            bits |= sign;
        }
    } else if (exp == 31) {
        bits = sign | 0x7f800000u | (man << 13);
    } else {
        bits = sign | ((exp + 112u) << 23) | (man << 13);
    }
    float out;
    memcpy(&out, &bits, 4);
    return out;
}

/* UpdateCacheCursor<_Key> (sampling_job.cc): the key stream is sorted so that a key appears when the
 * previous key of its track becomes the left bound; the loop stops at the first key not yet needed. */
#define DEFINE_UPDATE_CURSOR(fn, key_t)                                                            \
    static void fn(float ratio, int num_soa_tracks, const key_t* keys, int num_keys, int* cursor_io, \
                   int* cache, uint8_t* outdated) {                                                \
        const int num_tracks = num_soa_tracks * 4;                                                 \
        int cursor;                                                                                \
        if (!*cursor_io) {                                                                         \
            for (int i = 0; i < num_soa_tracks; ++i) {                                             \
                const int in0 = i * 4, in1 = in0 + num_tracks, out = i * 4 * 2;                    \
                for (int l = 0; l < 4; ++l) {                                                      \
                    cache[out + 2 * l + 0] = in0 + l;                                              \
                    cache[out + 2 * l + 1] = in1 + l;                                              \
                }                                                                                  \
            }                                                                                      \
            cursor = num_tracks * 2;                                                               \
            const int num_flags = (num_soa_tracks + 7) / 8;                                        \
            for (int i = 0; i < num_flags - 1; ++i) outdated[i] = 0xff;                            \
            outdated[num_flags - 1] = (uint8_t)(0xff >> (num_flags * 8 - num_soa_tracks));         \
        } else {                                                                                   \
            cursor = *cursor_io;                                                                   \
        }                                                                                          \
        while (cursor < num_keys && keys[cache[keys[cursor].track * 2 + 1]].ratio <= ratio) {      \
            const int track = keys[cursor].track;                                                  \
            outdated[track / 32] |= (uint8_t)(1 << ((track & 0x1f) / 4));                          \
            const int base = track * 2;                                                            \
            cache[base] = cache[base + 1];                                                         \
            cache[base + 1] = cursor;                                                              \
            ++cursor;                                                                              \
        }                                                                                          \
        *cursor_io = cursor;                                                                       \
    }
DEFINE_UPDATE_CURSOR(update_cursor_f3, orc_f3key)
DEFINE_UPDATE_CURSOR(update_cursor_q, orc_qkey)

/* DecompressFloat3: 4 keys -> SoaFloat3 via exact half->float */
static void decompress_f3(const orc_f3key* k[4], float out[3][4]) {
    for (int l = 0; l < 4; ++l)
        for (int c = 0; c < 3; ++c) out[c][l] = half_to_float(k[l]->value[c]);
}

/* DecompressQuaternion (Appendix A.4): smallest-three, 16 bit, largest component rebuilt */
static void decompress_q(const orc_qkey* k[4], float out[4][4]) {
    static const int kCpntMapping[4][4] = {{0, 0, 1, 2}, {0, 0, 1, 2}, {0, 1, 0, 2}, {0, 1, 2, 0}};
    const float kInt2Float = 1.f / (32767.f * 1.4142135623730950488016887242097f);
    for (int l = 0; l < 4; ++l) {
        const int* m = kCpntMapping[k[l]->largest];
        int cmp[4];
        for (int c = 0; c < 4; ++c) cmp[c] = k[l]->value[m[c]];
        cmp[k[l]->largest] = 0;
        float cpnt[4];
        if (k[l]->format == 1) {
            /* archive v7 (ozz >= 0.15, from memory -- see ozz_archive.c): magnitudes mapped from [-1/sqrt 2, 1/sqrt 2] */
            const float kScale = 1.4142135623730950488016887242097f / 32767.f, kOffset = -0.70710678118654752440084436210485f;
            for (int c = 0; c < 4; ++c) cpnt[c] = c == k[l]->largest ? 0.f : (float)cmp[c] * kScale + kOffset;
        } else {
            for (int c = 0; c < 4; ++c) cpnt[c] = kInt2Float * (float)cmp[c];
        }
        const float dot = cpnt[0] * cpnt[0] + cpnt[1] * cpnt[1] + cpnt[2] * cpnt[2] + cpnt[3] * cpnt[3];
        float ww0 = 1.f - dot;
        if (!(ww0 > 1e-16f)) ww0 = 1e-16f;   /* math::Max(eps, one - dot) */
        float w0 = ww0 * (1.f / sqrtf(ww0)); /* x * RSqrtEst(x), simd_ref: RSqrtEst = 1/sqrt */
        if (k[l]->sign) w0 = -w0;            /* Or(w0, sign << 31), w0 >= 0 */
        cpnt[k[l]->largest] = w0;            /* Or into the zeroed lane */
        for (int c = 0; c < 4; ++c) out[c][l] = cpnt[c];
    }
}

/* UpdateInterpKeyframes: refresh the decompressed bracketing keys of every outdated SoA group */
static void update_interp_f3(int num_soa_tracks, const orc_f3key* keys, const int* interp, uint8_t* outdated,
                             interp_f3* ik) {
    const int num_flags = (num_soa_tracks + 7) / 8;
    for (int j = 0; j < num_flags; ++j) {
        uint8_t od = outdated[j];
        outdated[j] = 0;
        for (int i = j * 8; od; ++i, od >>= 1) {
            if (!(od & 1)) continue;
            const int base = i * 4 * 2;
            for (int side = 0; side < 2; ++side) {
                const orc_f3key* k[4];
                for (int l = 0; l < 4; ++l) {
                    k[l] = &keys[interp[base + 2 * l + side]];
                    ik[i].ratio[side][l] = k[l]->ratio;
                }
                decompress_f3(k, ik[i].value[side]);
            }
        }
    }
}
static void update_interp_q(int num_soa_tracks, const orc_qkey* keys, const int* interp, uint8_t* outdated,
                            interp_q* ik) {
    const int num_flags = (num_soa_tracks + 7) / 8;
    for (int j = 0; j < num_flags; ++j) {
        uint8_t od = outdated[j];
        outdated[j] = 0;
        for (int i = j * 8; od; ++i, od >>= 1) {
            if (!(od & 1)) continue;
            const int base = i * 4 * 2;
            for (int side = 0; side < 2; ++side) {
                const orc_qkey* k[4];
                for (int l = 0; l < 4; ++l) {
                    k[l] = &keys[interp[base + 2 * l + side]];
                    ik[i].ratio[side][l] = k[l]->ratio;
                }
                decompress_q(k, ik[i].value[side]);
            }
        }
    }
}

/* Interpolates (Appendix A.5): alpha = (t - r0) * RcpEst(r1 - r0); Lerp T,S; NLerpEst R */
static void interpolates(float anim_ratio, int num_soa, const interp_f3* t, const interp_q* r, const interp_f3* s,
                         float* out_soa) {
    for (int i = 0; i < num_soa; ++i) {
        float* o = out_soa + (size_t)i * ORC_SOA_FLOATS;
        for (int l = 0; l < 4; ++l) {
            const float at = (anim_ratio - t[i].ratio[0][l]) * (1.f / (t[i].ratio[1][l] - t[i].ratio[0][l]));
            const float ar = (anim_ratio - r[i].ratio[0][l]) * (1.f / (r[i].ratio[1][l] - r[i].ratio[0][l]));
            const float as = (anim_ratio - s[i].ratio[0][l]) * (1.f / (s[i].ratio[1][l] - s[i].ratio[0][l]));
            for (int c = 0; c < 3; ++c) /* Lerp(a, b, f) = (b - a) * f + a */
                o[(0 + c) * 4 + l] = (t[i].value[1][c][l] - t[i].value[0][c][l]) * at + t[i].value[0][c][l];
            float q[4];
            for (int c = 0; c < 4; ++c) q[c] = (r[i].value[1][c][l] - r[i].value[0][c][l]) * ar + r[i].value[0][c][l];
            const float len2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
            const float inv_len = 1.f / sqrtf(len2);
            for (int c = 0; c < 4; ++c) o[(3 + c) * 4 + l] = q[c] * inv_len;
            for (int c = 0; c < 3; ++c)
                o[(7 + c) * 4 + l] = (s[i].value[1][c][l] - s[i].value[0][c][l]) * as + s[i].value[0][c][l];
        }
    }
}

static void export_cache(const orc_sampling_context* c, int num_soa, int32_t* out_cache) {
    if (!out_cache) return;
    const int n = num_soa * 4 * 2;
    for (int i = 0; i < n; ++i) {
        out_cache[i] = c->t_keys[i];
        out_cache[n + i] = c->r_keys[i];
        out_cache[2 * n + i] = c->s_keys[i];
    }
}

int orc_sample(const orc_animation* a, orc_sampling_context* c, float ratio, float* out_soa, int out_soa_count,
               int32_t* out_cache) {
    /* SamplingJob::Validate: animation && context, context->max_soa_tracks() >= animation->num_soa_tracks() */
    if (!a || !c) return 0;
    const int num_soa_tracks = a->num_soa;
    if (c->max_soa_tracks < num_soa_tracks) return 0;
    if (num_soa_tracks == 0) return 1;
    float anim_ratio = ratio; /* math::Clamp(0, ratio, 1) = Max(0, Min(ratio, 1)) */
    if (!(anim_ratio < 1.f)) anim_ratio = 1.f;
    if (!(anim_ratio > 0.f)) anim_ratio = 0.f;
    context_step(c, a, anim_ratio);
    update_cursor_f3(anim_ratio, num_soa_tracks, a->t, a->n_t, &c->t_cursor, c->t_keys, c->t_outdated);
    update_interp_f3(num_soa_tracks, a->t, c->t_keys, c->t_outdated, c->soa_t);
    update_cursor_q(anim_ratio, num_soa_tracks, a->r, a->n_r, &c->r_cursor, c->r_keys, c->r_outdated);
    update_interp_q(num_soa_tracks, a->r, c->r_keys, c->r_outdated, c->soa_r);
    update_cursor_f3(anim_ratio, num_soa_tracks, a->s, a->n_s, &c->s_cursor, c->s_keys, c->s_outdated);
    update_interp_f3(num_soa_tracks, a->s, c->s_keys, c->s_outdated, c->soa_s);
    const int n = out_soa_count < num_soa_tracks ? out_soa_count : num_soa_tracks;
    interpolates(anim_ratio, n, c->soa_t, c->soa_r, c->soa_s, out_soa);
    export_cache(c, num_soa_tracks, out_cache);
    return 1;
}

/* Stateless equivalent (Appendix A.3): per track, right = first key of the track (in stream order) whose
 * ratio is > t, or the track's last key when none is; left = the previous key of the same track.  This is
 * what the per-(character, joint) GPU search computes; tests assert it equals the cursor form. */
#define DEFINE_STATELESS(fn, key_t)                                                       \
    static void fn(float ratio, int num_tracks, const key_t* keys, int num_keys, int* cache) { \
        for (int tr = 0; tr < num_tracks; ++tr) cache[2 * tr] = cache[2 * tr + 1] = -1;   \
        /* walk the stream once, tracking for each track its last two keys seen so far; */  \
        /* a track freezes when its right key has ratio > t                           */  \
        for (int i = 0; i < num_keys; ++i) {                                              \
            const int tr = keys[i].track;                                                 \
            int* c = cache + 2 * tr;                                                      \
            if (c[1] >= 0 && c[0] >= 0 && keys[c[1]].ratio > ratio) continue;             \
            c[0] = c[1];                                                                  \
            c[1] = i;                                                                     \
        }                                                                                 \
    }
DEFINE_STATELESS(stateless_f3, orc_f3key)
DEFINE_STATELESS(stateless_q, orc_qkey)

int orc_sample_stateless(const orc_animation* a, float ratio, float* out_soa, int out_soa_count,
                         int32_t* out_cache) {
    if (!a) return 0;
    const int num_soa = a->num_soa;
    if (num_soa == 0) return 1;
    float anim_ratio = ratio;
    if (!(anim_ratio < 1.f)) anim_ratio = 1.f;
    if (!(anim_ratio > 0.f)) anim_ratio = 0.f;
    const int tracks = num_soa * 4;
    int* ct = (int*)malloc(sizeof(int) * (size_t)tracks * 2 * 3);
    int *cr = ct + tracks * 2, *cs = cr + tracks * 2;
    stateless_f3(anim_ratio, tracks, a->t, a->n_t, ct);
    stateless_q(anim_ratio, tracks, a->r, a->n_r, cr);
    stateless_f3(anim_ratio, tracks, a->s, a->n_s, cs);
    interp_f3* st = (interp_f3*)calloc((size_t)num_soa, sizeof(interp_f3));
    interp_q* sr = (interp_q*)calloc((size_t)num_soa, sizeof(interp_q));
    interp_f3* ss = (interp_f3*)calloc((size_t)num_soa, sizeof(interp_f3));
    uint8_t* all = (uint8_t*)malloc((size_t)(num_soa + 7) / 8);
    for (int pass = 0; pass < 3; ++pass) {
        const int num_flags = (num_soa + 7) / 8;
        for (int i = 0; i < num_flags - 1; ++i) all[i] = 0xff;
        all[num_flags - 1] = (uint8_t)(0xff >> (num_flags * 8 - num_soa));
        if (pass == 0) update_interp_f3(num_soa, a->t, ct, all, st);
        if (pass == 1) update_interp_q(num_soa, a->r, cr, all, sr);
        if (pass == 2) update_interp_f3(num_soa, a->s, cs, all, ss);
    }
    const int n = out_soa_count < num_soa ? out_soa_count : num_soa;
    interpolates(anim_ratio, n, st, sr, ss, out_soa);
    if (out_cache) {
        const int m = tracks * 2;
        for (int i = 0; i < m; ++i) {
            out_cache[i] = ct[i];
            out_cache[m + i] = cr[i];
            out_cache[2 * m + i] = cs[i];
        }
    }
    free(all);
    free(st);
    free(sr);
    free(ss);
    free(ct);
    return 1;
}

int orc_clip_clock(float* normalized_time, int* looped, float dt, float clip_duration, float speed, int loop) {
    /* OzzSkeletalAnimatorAnimationState::update (src/skeleton_ozz.cpp:411-428):
     * hasFinished() = _looped && !loop, tested before and after the _looped update;
     * getDuration() = anim.duration / (speed == 0 ? 1 : speed) (:398-404);
     * setNormalizedTime = fmodf(x, 1) (:392-396). */
    if (*looped && !loop) return 0;
    float sp = speed == 0.f ? 1.f : speed;
    float duration = clip_duration / sp;
    float norm_delta = dt / duration;
    float norm_time = *normalized_time + norm_delta;
    *looped = norm_time > 1.f;
    if (*looped && !loop) return 0;
    *normalized_time = fmodf(norm_time, 1.f);
    return 1;
}
