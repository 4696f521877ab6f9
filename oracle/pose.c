/*
 * oracle/pose.c -- TEST INFRASTRUCTURE (see dmk_oracle.h).
 *
 * BlendingJob, LocalToModelJob (ozz-animation 0.14, restated: SURVEY.md Appendix A.6-A.7; ozz is absent
 * from /root/reference -> parity unpinned), darmok's handedness flip, skin palette and the LBS shader.
 */
#include "dmk_oracle.h"

#include <math.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------
 * ozz BlendingJob::Run, non-additive layers without joint masks -- the only form darmok builds:
 *   state blend tree   src/skeleton_ozz.cpp:556-608  (threshold = definition, rest pose = clip at (0,0))
 *   transition         src/skeleton_ozz.cpp:709-724  (2 layers (1-v, v), threshold = bx::kFloatSmallest)
 * ---------------------------------------------------------------------------------------------- */
static void blend_1st_pass(const float* in, float w, float* out) {
    for (int i = 0; i < ORC_SOA_FLOATS; ++i) out[i] = in[i] * w; /* T, R and S all scaled */
}
static void blend_n_pass(const float* in, float w, float* out) {
    for (int l = 0; l < 4; ++l) {
        for (int c = 0; c < 3; ++c) out[c * 4 + l] = out[c * 4 + l] + in[c * 4 + l] * w;
        /* negate opposed quaternions: sign = signbit(Dot(out.R, in.R)), in.R xor sign */
        const float dot = out[12 + l] * in[12 + l] + out[16 + l] * in[16 + l] + out[20 + l] * in[20 + l] +
                          out[24 + l] * in[24 + l];
        const int neg = signbit(dot);
        for (int c = 3; c < 7; ++c) {
            const float q = neg ? -in[c * 4 + l] : in[c * 4 + l];
            out[c * 4 + l] = out[c * 4 + l] + q * w;
        }
        for (int c = 7; c < 10; ++c) out[c * 4 + l] = out[c * 4 + l] + in[c * 4 + l] * w;
    }
}

int orc_blend(int num_layers, const float* const* layer_soa, const float* weights, const float* rest_soa,
              float threshold, int num_soa, float* out_soa) {
    /* BlendingJob::Validate: threshold > 0, rest pose not empty (a layer without transform also fails) */
    if (!(threshold > 0.f) || !rest_soa || num_soa <= 0 || !out_soa) return 0;
    for (int i = 0; i < num_layers; ++i)
        if (!layer_soa[i]) return 0;
    float accumulated = 0.f;
    int num_passes = 0;
    /* BlendLayers */
    for (int i = 0; i < num_layers; ++i) {
        const float w = weights[i];
        if (w <= 0.f) continue; /* also skips NaN?  no: NaN <= 0 is false, ozz lets a NaN weight through */
        accumulated += w;
        for (int g = 0; g < num_soa; ++g) {
            const float* in = layer_soa[i] + (size_t)g * ORC_SOA_FLOATS;
            float* out = out_soa + (size_t)g * ORC_SOA_FLOATS;
            if (num_passes == 0)
                blend_1st_pass(in, w, out);
            else
                blend_n_pass(in, w, out);
        }
        ++num_passes;
    }
    /* BlendRestPose (no partial pass) */
    const float bp_weight = threshold - accumulated;
    if (bp_weight > 0.f) {
        if (num_passes == 0) {
            accumulated = 1.f;
            memcpy(out_soa, rest_soa, sizeof(float) * ORC_SOA_FLOATS * (size_t)num_soa);
        } else {
            accumulated = threshold;
            for (int g = 0; g < num_soa; ++g)
                blend_n_pass(rest_soa + (size_t)g * ORC_SOA_FLOATS, bp_weight, out_soa + (size_t)g * ORC_SOA_FLOATS);
        }
    }
    /* Normalize: rotation = NormalizeEst(rotation); translation, scale *= 1 / accumulated */
    const float ratio = 1.f / accumulated;
    for (int g = 0; g < num_soa; ++g) {
        float* o = out_soa + (size_t)g * ORC_SOA_FLOATS;
        for (int l = 0; l < 4; ++l) {
            const float len2 = o[12 + l] * o[12 + l] + o[16 + l] * o[16 + l] + o[20 + l] * o[20 + l] + o[24 + l] * o[24 + l];
            const float inv_len = 1.f / sqrtf(len2);
            for (int c = 3; c < 7; ++c) o[c * 4 + l] = o[c * 4 + l] * inv_len;
            for (int c = 0; c < 3; ++c) o[c * 4 + l] = o[c * 4 + l] * ratio;
            for (int c = 7; c < 10; ++c) o[c * 4 + l] = o[c * 4 + l] * ratio;
        }
    }
    return 1;
}

/* ------------------------------------------------------------------------------------------------
 * ozz LocalToModelJob::Run with root = identity, from = kNoParent, to = kMaxJoints (darmok sets only
 * input/output/skeleton: src/skeleton_ozz.cpp:911-915 and :1009-1039).
 * ---------------------------------------------------------------------------------------------- */
/* ozz Float4x4 operator*: ret.col[j] = (a.col0 * b[j].x + a.col1 * b[j].y) + (a.col2 * b[j].z + a.col3 * b[j].w) */
static void ozz_mat_mul(const float* a, const float* b, float* out) {
    float r[16];
    for (int j = 0; j < 4; ++j)
        for (int i = 0; i < 4; ++i)
            r[j * 4 + i] = (a[0 + i] * b[j * 4 + 0] + a[4 + i] * b[j * 4 + 1]) +
                           (a[8 + i] * b[j * 4 + 2] + a[12 + i] * b[j * 4 + 3]);
    memcpy(out, r, sizeof r);
}

/* SoaFloat4x4::FromAffine for lane l of one SoaTransform, written as an AoS column-major matrix */
static void from_affine(const float* soa, int l, float* m) {
    const float tx = soa[0 + l], ty = soa[4 + l], tz = soa[8 + l];
    const float x = soa[12 + l], y = soa[16 + l], z = soa[20 + l], w = soa[24 + l];
    const float sx = soa[28 + l], sy = soa[32 + l], sz = soa[36 + l];
    const float xx = x * x, xy = x * y, xz = x * z, xw = x * w;
    const float yy = y * y, yz = y * z, yw = y * w;
    const float zz = z * z, zw = z * w;
    const float one = 1.f, two = one + one;
    m[0] = sx * (one - two * (yy + zz));
    m[1] = sx * two * (xy + zw);
    m[2] = sx * two * (xz - yw);
    m[3] = 0.f;
    m[4] = sy * two * (xy - zw);
    m[5] = sy * (one - two * (xx + zz));
    m[6] = sy * two * (yz + xw);
    m[7] = 0.f;
    m[8] = sz * two * (xz + yw);
    m[9] = sz * two * (yz - xw);
    m[10] = sz * (one - two * (xx + yy));
    m[11] = 0.f;
    m[12] = tx;
    m[13] = ty;
    m[14] = tz;
    m[15] = one;
}

int orc_local_to_model(const orc_skeleton* s, const float* locals_soa, float* models) {
    /* LocalToModelJob::Validate: skeleton set, input/output large enough */
    if (!s || !locals_soa || !models) return 0;
    static const float identity[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    for (int i = 0; i < s->num_joints; ++i) {
        float local[16];
        from_affine(locals_soa + (size_t)(i / 4) * ORC_SOA_FLOATS, i & 3, local);
        const int parent = s->parents[i];
        const float* pm = parent == ORC_NO_PARENT ? identity : models + (size_t)parent * 16;
        ozz_mat_mul(pm, local, models + (size_t)i * 16);
    }
    return 1;
}

/* ------------------------------------------------------------------------------------------------
 * glm::mat4 helpers in glm's scalar operation order
 * ---------------------------------------------------------------------------------------------- */
void orc_mat4_mul(const float* a, const float* b, float* out) {
    /* Result[j] = ((A0 * B[j].x + A1 * B[j].y) + A2 * B[j].z) + A3 * B[j].w */
    float r[16];
    for (int j = 0; j < 4; ++j)
        for (int i = 0; i < 4; ++i)
            r[j * 4 + i] = ((a[0 + i] * b[j * 4 + 0] + a[4 + i] * b[j * 4 + 1]) + a[8 + i] * b[j * 4 + 2]) +
                           a[12 + i] * b[j * 4 + 3];
    memcpy(out, r, sizeof r);
}

void orc_mat4_inverse(const float* in, float* out) {
    /* glm::detail::compute_inverse<4,4>: cofactor expansion; m[c][r] = in[c*4 + r] */
#define M(c, r) in[(c) * 4 + (r)]
    const float c00 = M(2, 2) * M(3, 3) - M(3, 2) * M(2, 3);
    const float c02 = M(1, 2) * M(3, 3) - M(3, 2) * M(1, 3);
    const float c03 = M(1, 2) * M(2, 3) - M(2, 2) * M(1, 3);
    const float c04 = M(2, 1) * M(3, 3) - M(3, 1) * M(2, 3);
    const float c06 = M(1, 1) * M(3, 3) - M(3, 1) * M(1, 3);
    const float c07 = M(1, 1) * M(2, 3) - M(2, 1) * M(1, 3);
    const float c08 = M(2, 1) * M(3, 2) - M(3, 1) * M(2, 2);
    const float c10 = M(1, 1) * M(3, 2) - M(3, 1) * M(1, 2);
    const float c11 = M(1, 1) * M(2, 2) - M(2, 1) * M(1, 2);
    const float c12 = M(2, 0) * M(3, 3) - M(3, 0) * M(2, 3);
    const float c14 = M(1, 0) * M(3, 3) - M(3, 0) * M(1, 3);
    const float c15 = M(1, 0) * M(2, 3) - M(2, 0) * M(1, 3);
    const float c16 = M(2, 0) * M(3, 2) - M(3, 0) * M(2, 2);
    const float c18 = M(1, 0) * M(3, 2) - M(3, 0) * M(1, 2);
    const float c19 = M(1, 0) * M(2, 2) - M(2, 0) * M(1, 2);
    const float c20 = M(2, 0) * M(3, 1) - M(3, 0) * M(2, 1);
    const float c22 = M(1, 0) * M(3, 1) - M(3, 0) * M(1, 1);
    const float c23 = M(1, 0) * M(2, 1) - M(2, 0) * M(1, 1);
    const float f0[4] = {c00, c00, c02, c03}, f1[4] = {c04, c04, c06, c07}, f2[4] = {c08, c08, c10, c11};
    const float f3[4] = {c12, c12, c14, c15}, f4[4] = {c16, c16, c18, c19}, f5[4] = {c20, c20, c22, c23};
    const float v0[4] = {M(1, 0), M(0, 0), M(0, 0), M(0, 0)}, v1[4] = {M(1, 1), M(0, 1), M(0, 1), M(0, 1)};
    const float v2[4] = {M(1, 2), M(0, 2), M(0, 2), M(0, 2)}, v3[4] = {M(1, 3), M(0, 3), M(0, 3), M(0, 3)};
    static const float sa[4] = {+1, -1, +1, -1}, sb[4] = {-1, +1, -1, +1};
    float inv[16];
    for (int i = 0; i < 4; ++i) {
        inv[0 + i] = ((v1[i] * f0[i] - v2[i] * f1[i]) + v3[i] * f2[i]) * sa[i];
        inv[4 + i] = ((v0[i] * f0[i] - v2[i] * f3[i]) + v3[i] * f4[i]) * sb[i];
        inv[8 + i] = ((v0[i] * f1[i] - v1[i] * f3[i]) + v3[i] * f5[i]) * sa[i];
        inv[12 + i] = ((v0[i] * f2[i] - v1[i] * f4[i]) + v2[i] * f5[i]) * sb[i];
    }
    const float d0 = M(0, 0) * inv[0], d1 = M(0, 1) * inv[4], d2 = M(0, 2) * inv[8], d3 = M(0, 3) * inv[12];
    const float det = (d0 + d1) + (d2 + d3);
    const float ood = 1.f / det;
    for (int i = 0; i < 16; ++i) out[i] = inv[i] * ood;
#undef M
}

void orc_flip_handedness(const float* m, float* out) {
    /* Math::flipHandedness (src/math.cpp:21-25): zNegate * mat * zNegate, zNegate = scale(1, 1, -1) */
    static const float zneg[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, -1, 0, 0, 0, 0, 1};
    float t[16];
    orc_mat4_mul(zneg, m, t);
    orc_mat4_mul(t, zneg, out);
}

void orc_palette(const float* models, int num_joints, const int32_t* remap, const float* inv_bind, int k,
                 float* out) {
    /* SkeletalAnimationRenderComponent::beforeRenderEntity (src/skeleton.cpp:236-249):
     *   _skinning = [mat4(1)] + [animator->getJointModelMatrix(joint.name) * joint.inverseBindPose ...]
     * getJointModelMatrix (src/skeleton_ozz.cpp:962-977): OzzUtils::convert(_models[i]) or mat4(1). */
    static const float identity[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    memcpy(out, identity, sizeof identity);
    for (int j = 0; j < k; ++j) {
        float model[16];
        if (remap[j] >= 0 && remap[j] < num_joints)
            orc_flip_handedness(models + (size_t)remap[j] * 16, model);
        else
            memcpy(model, identity, sizeof identity);
        orc_mat4_mul(model, inv_bind + (size_t)j * 16, out + (size_t)(j + 1) * 16);
    }
}

void orc_skin(const float* palette, int palette_size, const float* pos, const float* nrm, const float* tan,
              const float* indices, const float* weights, int num_vertices, float* out) {
    /* darmok_skinning.inc.slang:5-28 + forward.slang:33-38: the same blended matrix is applied to
     * float4(pos, 1), float4(normal, 0), float4(tangent, 0); no renormalisation at this step. */
    (void)palette_size;
    for (int v = 0; v < num_vertices; ++v) {
        const float* idx = indices + (size_t)v * 4;
        const float* w = weights + (size_t)v * 4;
        const float* in[3] = {pos + (size_t)v * 3, nrm + (size_t)v * 3, tan + (size_t)v * 3};
        float* o = out + (size_t)v * 9;
        if (idx[0] >= 0.f) {
            float m[16];
            const float* p0 = palette + (size_t)(int)(idx[0] + 1.f) * 16;
            const float* p1 = palette + (size_t)(int)(idx[1] + 1.f) * 16;
            const float* p2 = palette + (size_t)(int)(idx[2] + 1.f) * 16;
            const float* p3 = palette + (size_t)(int)(idx[3] + 1.f) * 16;
            for (int e = 0; e < 16; ++e) m[e] = ((w[0] * p0[e] + w[1] * p1[e]) + w[2] * p2[e]) + w[3] * p3[e];
            for (int a = 0; a < 3; ++a) {
                const float vw = a == 0 ? 1.f : 0.f;
                for (int r = 0; r < 3; ++r) /* mul(skin, v): row r of skin dot v */
                    o[a * 3 + r] = ((m[0 + r] * in[a][0] + m[4 + r] * in[a][1]) + m[8 + r] * in[a][2]) + m[12 + r] * vw;
            }
        } else {
            for (int a = 0; a < 3; ++a)
                for (int r = 0; r < 3; ++r) o[a * 3 + r] = in[a][r];
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * SkeletalAnimatorImpl::getJointModelMatrixes(dir) (src/skeleton_ozz.cpp:979-1004), the debug-bone matrices
 * RenderableSkeleton draws (src/skeleton.cpp:62-113): for every joint i with a parent p, keyed by the PARENT's name
 * (unordered_map::emplace: the first child of a parent wins):
 *   parentPos = convert(models[p])[3], childPos = convert(models[i])[3], diff = childPos - parentPos
 *   mat = Math::transform(parentPos, glm::rotation(dir, normalize(diff)), vec3(length(diff)))
 * glm::rotation (glm/gtx/quaternion.inl, restated from the published source; glm is not in the container):
 * identity when cos >= 1 - eps, a half turn about a perpendicular axis when cos < -1 + eps, else Stan Melax's form.
 * out: 16 floats per joint (entry of joint p = the bone from p to its first child), has[p] = 0 for joints without children.
 * ---------------------------------------------------------------------------------------------- */
static void glm_rotation(const float* orig, const float* dest, float* q /* x y z w */) {
    const float eps = 1.1920929e-07f; /* glm::epsilon<float>() */
    const float cos_theta = (orig[0] * dest[0] + orig[1] * dest[1]) + orig[2] * dest[2];
    float axis[3];
    if (cos_theta >= 1.f - eps) {
        q[0] = q[1] = q[2] = 0.f;
        q[3] = 1.f;
        return;
    }
    if (cos_theta < -1.f + eps) {
        /* cross((0,0,1), orig), or cross((1,0,0), orig) when that is (nearly) zero */
        axis[0] = 0.f * orig[2] - orig[1] * 1.f;
        axis[1] = 1.f * orig[0] - orig[2] * 0.f;
        axis[2] = 0.f * orig[1] - orig[0] * 0.f;
        if ((axis[0] * axis[0] + axis[1] * axis[1]) + axis[2] * axis[2] < eps) {
            axis[0] = 0.f * orig[2] - orig[1] * 0.f;
            axis[1] = 0.f * orig[0] - orig[2] * 1.f;
            axis[2] = 1.f * orig[1] - orig[0] * 0.f;
        }
        const float inv = 1.f / sqrtf((axis[0] * axis[0] + axis[1] * axis[1]) + axis[2] * axis[2]);
        /* angleAxis(pi, axis): s = sin(pi / 2), w = cos(pi / 2) in float */
        const float half = 3.14159265358979323846264338327950288f * 0.5f;
        const float s = sinf(half);
        q[0] = axis[0] * inv * s;
        q[1] = axis[1] * inv * s;
        q[2] = axis[2] * inv * s;
        q[3] = cosf(half);
        return;
    }
    axis[0] = orig[1] * dest[2] - dest[1] * orig[2];
    axis[1] = orig[2] * dest[0] - dest[2] * orig[0];
    axis[2] = orig[0] * dest[1] - dest[0] * orig[1];
    const float s = sqrtf((1.f + cos_theta) * 2.f);
    const float invs = 1.f / s;
    q[0] = axis[0] * invs;
    q[1] = axis[1] * invs;
    q[2] = axis[2] * invs;
    q[3] = s * 0.5f;
}

void orc_bone_matrices(const orc_skeleton* s, const float* models, const float* dir, float* out, uint8_t* has) {
    const int n = s->num_joints;
    memset(out, 0, (size_t)n * 16 * sizeof(float));
    memset(has, 0, (size_t)n);
    for (int i = 0; i < n; ++i) {
        const int p = s->parents[i];
        if (p == ORC_NO_PARENT || has[p]) continue; /* emplace keeps the first child's entry */
        float pm[16], cm[16];
        orc_flip_handedness(models + (size_t)p * 16, pm);
        orc_flip_handedness(models + (size_t)i * 16, cm);
        const float diff[3] = {cm[12] - pm[12], cm[13] - pm[13], cm[14] - pm[14]};
        const float len2 = (diff[0] * diff[0] + diff[1] * diff[1]) + diff[2] * diff[2];
        const float inv = 1.f / sqrtf(len2); /* glm::normalize = v * inversesqrt(dot(v, v)) */
        const float nd[3] = {diff[0] * inv, diff[1] * inv, diff[2] * inv};
        float q[4];
        glm_rotation(dir, nd, q);
        const float len = sqrtf(len2); /* glm::length */
        const float scale[3] = {len, len, len};
        orc_transform_local_matrix(pm + 12, q, scale, out + (size_t)p * 16);
        has[p] = 1;
    }
}
