// pose_kernel.cu -- K1..K3: clip clock + keyframe sampling + layer blend + local-to-model + skin palette.
//
// One warp per character.  Replaces, per character and per tick (paths relative to the darmok repository):
//   OzzSkeletalAnimatorAnimationState::update  src/skeleton_ozz.cpp:411-440   (clip clock, ozz SamplingJob)
//   BlendingJob in state/transition update     src/skeleton_ozz.cpp:556-608, :709-724
//   LocalToModelJob in SkeletalAnimatorImpl    src/skeleton_ozz.cpp:1034-1039
//   getJointModelMatrix + palette build        src/skeleton_ozz.cpp:962-977, src/skeleton.cpp:236-249
//
// Arithmetic follows the ozz-animation 0.14 scalar ("simd_ref") semantics operation by operation so that key
// indices are identical and matrices agree with the CPU restatement: this TU is compiled with -fmad=false
// (no contraction), IEEE division and square root (-prec-div/-prec-sqrt defaults), no flush-to-zero.
//
// Data layout: key tracks are SoA per (clip, component, track) and shared by all characters (L1/L2 resident);
// the per-skeleton tables (parents, level-order schedule, remap, inverse binds) are staged in shared memory once
// per persistent block; hierarchy concatenation runs level by level, one (joint, column) task per lane.
#include "device_intrinsics.cuh"
#include "device_types.cuh"
#include "kernels.cuh"

#include <cuda_fp16.h>

namespace dmk {

namespace {

constexpr int kWarpsPerBlock = 4;   // maximum; skeletons with many joints run fewer warps per block (shared memory)
constexpr int kMS = 13;  // shared-memory stride of a 3x4 matrix in floats: odd, so lanes on consecutive joints hit distinct banks

// 1/x and 1/sqrt(x) of the scalar reference (simd_ref RcpEst / RSqrtEst): correctly rounded reciprocal == 1.f / x
__device__ __forceinline__ float rcp_exact(float x) { return __frcp_rn(x); }
__device__ __forceinline__ float rsqrt_exact(float x) { return __frcp_rn(__fsqrt_rn(x)); }

// Stateless equivalent of ozz's sampling cursor: right = first key (index >= 1) of the
// track with ratio > t, else the last key; left = right - 1.
__device__ __forceinline__ void find_keys(const float* __restrict__ ratio, uint2 d, float t, int& left, int& right) {
    const float* base = ratio + d.x;
    int a = 1, b = (int)d.y - 1;
    while (a < b) {
        const int m = (a + b) >> 1;
        if (__ldg(base + m) > t) b = m; else a = m + 1;
    }
    right = (int)d.x + a;
    left = right - 1;
}

// The two bracketing keys of one track at ratio t: ratios, packed values and (optionally) the ozz stream indices.
struct Bracket {
    float r0, r1;
    uint2 v0, v1;
};

// F: compile-time shape of the sampling path, so that the common case carries no dead branches:
//   kBuckets  the clip set has a bucket table (else per-track binary search)
//   kKeys     the ozz key-stream indices are written out (parity tests only)
enum : int { kBuckets = 1, kKeys = 2 };

template <int F>
__device__ __forceinline__ Bracket fetch_bracket(const ClipSetDev& cs, int comp, int clip, int track, float t,
                                                 int32_t* key_out) {
    Bracket k;
    const int P = cs.padded_tracks;
    if (F & kBuckets) {
        // bucket table: one coalesced 40-byte entry per (bucket, track)
        const uint2 bd = __ldg(cs.bucket_dir + clip);
        const int G = (int)bd.y;
        int b = (int)(t * (float)G);  // G is a power of two: exact
        b = b < G - 1 ? b : G - 1;
        const uint32_t e = bd.x + (uint32_t)((comp * G + b) * P + track);  // tables hold < 2^32 entries
        const float4 a = __ldg(cs.bucket_a + e);
        const uint4 vb = __ldg(cs.bucket_b + e);
        const uint2 vc = __ldg(cs.bucket_c + e);
        const bool lower = (a.y > t) || (a.z < 0.f);
        k.r0 = lower ? a.x : a.y;
        k.r1 = lower ? a.y : a.z;
        k.v0 = lower ? make_uint2(vb.x, vb.y) : make_uint2(vb.z, vb.w);
        k.v1 = lower ? make_uint2(vb.z, vb.w) : vc;
        if (F & kKeys) {
            const uint2 d = __ldg(cs.dir + (size_t)(clip * 3 + comp) * P + track);
            const int right = (int)d.x + __float_as_int(a.w) + (lower ? 0 : 1);
            key_out[0] = __ldg(cs.stream[comp] + right - 1);
            key_out[1] = __ldg(cs.stream[comp] + right);
        }
        return k;
    }
    const uint2 d = __ldg(cs.dir + (size_t)(clip * 3 + comp) * P + track);
    int l, r;
    find_keys(cs.ratio[comp], d, t, l, r);
    k.r0 = __ldg(cs.ratio[comp] + l);
    k.r1 = __ldg(cs.ratio[comp] + r);
    k.v0 = __ldg(cs.value[comp] + l);
    k.v1 = __ldg(cs.value[comp] + r);
    if (F & kKeys) {
        key_out[0] = __ldg(cs.stream[comp] + l);
        key_out[1] = __ldg(cs.stream[comp] + r);
    }
    return k;
}

__device__ __forceinline__ float half_bits_to_float(unsigned h) {
    return __half2float(__ushort_as_half((unsigned short)h));
}

// Float3 track (translation or scale): DecompressFloat3 + Lerp
template <int F>
__device__ __forceinline__ void sample_f3(const ClipSetDev& cs, int comp, int clip, int track, float t, float out[3],
                                          int32_t* key_out) {
    const Bracket k = fetch_bracket<F>(cs, comp, clip, track, t, key_out);
    const float r0 = k.r0, r1 = k.r1;
    const uint2 v0 = k.v0, v1 = k.v1;
    const float alpha = (t - r0) * rcp_exact(r1 - r0);
    const float a[3] = {half_bits_to_float(v0.x & 0xffffu), half_bits_to_float(v0.x >> 16), half_bits_to_float(v0.y & 0xffffu)};
    const float b[3] = {half_bits_to_float(v1.x & 0xffffu), half_bits_to_float(v1.x >> 16), half_bits_to_float(v1.y & 0xffffu)};
#pragma unroll
    for (int c = 0; c < 3; ++c) out[c] = (b[c] - a[c]) * alpha + a[c];
}

// DecompressQuaternion (Appendix A.4) for one key
__device__ __forceinline__ void decode_quat(uint2 v, float q[4]) {
    float f0, f1, f2;
    if ((v.y >> 19) & 1u) {   // archive v7 (ozz >= 0.15): 15-bit magnitudes mapped from [-1/sqrt 2, 1/sqrt 2] (clip-uniform branch)
        const float kScale = 1.4142135623730950488016887242097f / 32767.f, kOffset = -0.70710678118654752440084436210485f;
        f0 = (float)(v.x & 0xffffu) * kScale + kOffset;
        f1 = (float)(v.x >> 16) * kScale + kOffset;
        f2 = (float)(v.y & 0xffffu) * kScale + kOffset;
    } else {                  // archive v6: signed 16 bit
        const float kInt2Float = 1.f / (32767.f * 1.4142135623730950488016887242097f);
        f0 = kInt2Float * (float)(short)(v.x & 0xffffu);
        f1 = kInt2Float * (float)(short)(v.x >> 16);
        f2 = kInt2Float * (float)(short)(v.y & 0xffffu);
    }
    const unsigned largest = (v.y >> 16) & 3u;
    // the rebuilt lane is zero in ozz's dot product, so the sum is ((f0^2 + f1^2) + f2^2) for every mapping
    const float dot = (f0 * f0 + f1 * f1) + f2 * f2;
    float ww0 = 1.f - dot;
    ww0 = (ww0 > 1e-16f) ? ww0 : 1e-16f;
    float w0 = ww0 * rsqrt_exact(ww0);
    if ((v.y >> 18) & 1u) w0 = -w0;
    q[0] = largest == 0 ? w0 : f0;
    q[1] = largest == 0 ? f0 : (largest == 1 ? w0 : f1);
    q[2] = largest <= 1 ? f1 : (largest == 2 ? w0 : f2);
    q[3] = largest == 3 ? w0 : f2;
}

// rotation track: decode both keys, NLerpEst
template <int F>
__device__ __forceinline__ void sample_quat(const ClipSetDev& cs, int clip, int track, float t, float out[4],
                                            int32_t* key_out) {
    const Bracket k = fetch_bracket<F>(cs, 1, clip, track, t, key_out);
    const float r0 = k.r0, r1 = k.r1;
    const uint2 v0 = k.v0, v1 = k.v1;
    const float alpha = (t - r0) * rcp_exact(r1 - r0);
    float a[4], b[4], q[4];
    decode_quat(v0, a);
    decode_quat(v1, b);
#pragma unroll
    for (int c = 0; c < 4; ++c) q[c] = (b[c] - a[c]) * alpha + a[c];
    const float len2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    const float inv_len = rsqrt_exact(len2);
#pragma unroll
    for (int c = 0; c < 4; ++c) out[c] = q[c] * inv_len;
}

struct Xform {  // one joint's local transform
    float t[3], r[4], s[3];
};

template <int F>
__device__ __forceinline__ void sample_joint(const ClipSetDev& cs, int clip, int track, float ratio, Xform& x,
                                             int32_t* key_out /* [3][P][2] base for this layer (kKeys) */) {
    const int P = cs.padded_tracks;
    sample_f3<F>(cs, 0, clip, track, ratio, x.t, (F & kKeys) ? key_out + ((size_t)0 * P + track) * 2 : nullptr);
    sample_quat<F>(cs, clip, track, ratio, x.r, (F & kKeys) ? key_out + ((size_t)1 * P + track) * 2 : nullptr);
    sample_f3<F>(cs, 2, clip, track, ratio, x.s, (F & kKeys) ? key_out + ((size_t)2 * P + track) * 2 : nullptr);
}

// OZZ_BLEND_N_PASS for one joint
__device__ __forceinline__ void blend_n_pass(Xform& acc, const Xform& in, float w) {
#pragma unroll
    for (int c = 0; c < 3; ++c) acc.t[c] = acc.t[c] + in.t[c] * w;
    const float dot = acc.r[0] * in.r[0] + acc.r[1] * in.r[1] + acc.r[2] * in.r[2] + acc.r[3] * in.r[3];
    const bool neg = (__float_as_uint(dot) >> 31) != 0u;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float q = neg ? -in.r[c] : in.r[c];
        acc.r[c] = acc.r[c] + q * w;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) acc.s[c] = acc.s[c] + in.s[c] * w;
}

// Normalize step of BlendingJob for one joint
__device__ __forceinline__ void blend_normalize(Xform& acc, float inv_weight) {
    const float len2 = acc.r[0] * acc.r[0] + acc.r[1] * acc.r[1] + acc.r[2] * acc.r[2] + acc.r[3] * acc.r[3];
    const float inv_len = rsqrt_exact(len2);
#pragma unroll
    for (int k = 0; k < 4; ++k) acc.r[k] = acc.r[k] * inv_len;
#pragma unroll
    for (int k = 0; k < 3; ++k) { acc.t[k] = acc.t[k] * inv_weight; acc.s[k] = acc.s[k] * inv_weight; }
}

__device__ __forceinline__ void blend_first_pass(Xform& acc, const Xform& in, float w) {  // OZZ_BLEND_1ST_PASS
#pragma unroll
    for (int k = 0; k < 3; ++k) acc.t[k] = in.t[k] * w;
#pragma unroll
    for (int k = 0; k < 4; ++k) acc.r[k] = in.r[k] * w;
#pragma unroll
    for (int k = 0; k < 3; ++k) acc.s[k] = in.s[k] * w;
}

// One group of layers [begin, begin + count) of a character, for joint jt.  `layer(l, clip, ratio, weight)` yields the
// resolved state of layer l (clip index validated, ratio clamped to [0, 1] as SamplingJob::Run does).
template <int F, typename LayerFn>
__device__ __forceinline__ Xform eval_group(const ClipSetDev& cs, const LayerFn& layer, int begin, int count, int mode, int rest_idx,
                                            float threshold, int jt, int32_t* key_base, int P) {
    Xform acc;
    if (mode == kGroupZero) {   // the never-sampled clip of a state: darmok's zero-initialised locals are the pose
#pragma unroll
        for (int k = 0; k < 3; ++k) acc.t[k] = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) acc.r[k] = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) acc.s[k] = 0.f;
        return acc;
    }
    if (mode == kGroupPassthrough) {
        // every clip of a state is sampled each tick (:538-545); only one of them is the pose
        for (int l = begin; l < begin + count; ++l) {
            const bool chosen = l == begin + rest_idx;
            if (!chosen && !(F & kKeys)) continue;
            int cl;
            float tl, wl;
            layer(l, cl, tl, wl);
            Xform in;
            sample_joint<F>(cs, cl, jt, tl, in, (F & kKeys) ? key_base + (size_t)l * 3 * P * 2 : nullptr);
            if (chosen) acc = in;
        }
        return acc;
    }
    // BlendingJob bookkeeping is independent of the joint: accumulated weight, passes, rest-pose weight
    float accumulated = 0.f;
    int num_passes = 0;
    for (int l = begin; l < begin + count; ++l) {
        int cl;
        float tl, w;
        layer(l, cl, tl, w);
        if (!(w <= 0.f)) { accumulated += w; ++num_passes; }
    }
    const float bp_weight = threshold - accumulated;
    const bool need_rest = bp_weight > 0.f;
    float norm_weight = accumulated;
    if (need_rest) norm_weight = (num_passes == 0) ? 1.f : threshold;
    const float inv_weight = rcp_exact(norm_weight);
    Xform rest;
    int pass = 0;
#pragma unroll 1
    for (int l = begin; l < begin + count; ++l) {
        int cl;
        float tl, w;
        layer(l, cl, tl, w);
        const bool contributes = !(w <= 0.f);
        const bool is_rest = need_rest && l == begin + rest_idx;
        if (!contributes && !is_rest && !(F & kKeys)) continue;
        Xform in;
        sample_joint<F>(cs, cl, jt, tl, in, (F & kKeys) ? key_base + (size_t)l * 3 * P * 2 : nullptr);
        if (is_rest) rest = in;
        if (!contributes) continue;
        if (pass == 0) blend_first_pass(acc, in, w);
        else blend_n_pass(acc, in, w);
        ++pass;
    }
    if (need_rest) {  // BlendRestPose
        if (pass == 0) acc = rest;
        else blend_n_pass(acc, rest, bp_weight);
    }
    blend_normalize(acc, inv_weight);
    return acc;
}

// The transition's BlendingJob: layers (w0, a), (w1, b), rest pose = a (src/skeleton_ozz.cpp:709-724)
__device__ __forceinline__ Xform blend_two(const Xform& a, const Xform& b, float w0, float w1, float threshold) {
    Xform acc;
    float accumulated = 0.f;
    int pass = 0;
    if (!(w0 <= 0.f)) { accumulated += w0; blend_first_pass(acc, a, w0); ++pass; }
    if (!(w1 <= 0.f)) {
        accumulated += w1;
        if (pass == 0) blend_first_pass(acc, b, w1);
        else blend_n_pass(acc, b, w1);
        ++pass;
    }
    const float bp_weight = threshold - accumulated;
    float norm_weight = accumulated;
    if (bp_weight > 0.f) {
        if (pass == 0) { acc = a; norm_weight = 1.f; }
        else { blend_n_pass(acc, a, bp_weight); norm_weight = threshold; }
    }
    blend_normalize(acc, rcp_exact(norm_weight));
    return acc;
}

// SoaFloat4x4::FromAffine for one joint -> 3x4 affine, columns of 3 (m[c*3 + r])
__device__ __forceinline__ void from_affine(const Xform& x, float* m) {
    const float qx = x.r[0], qy = x.r[1], qz = x.r[2], qw = x.r[3];
    const float xx = qx * qx, xy = qx * qy, xz = qx * qz, xw = qx * qw;
    const float yy = qy * qy, yz = qy * qz, yw = qy * qw;
    const float zz = qz * qz, zw = qz * qw;
    const float one = 1.f, two = 2.f;
    m[0] = x.s[0] * (one - two * (yy + zz));
    m[1] = x.s[0] * two * (xy + zw);
    m[2] = x.s[0] * two * (xz - yw);
    m[3] = x.s[1] * two * (xy - zw);
    m[4] = x.s[1] * (one - two * (xx + zz));
    m[5] = x.s[1] * two * (yz + xw);
    m[6] = x.s[2] * two * (xz + yw);
    m[7] = x.s[2] * two * (yz - xw);
    m[8] = x.s[2] * (one - two * (xx + yy));
    m[9] = x.t[0];
    m[10] = x.t[1];
    m[11] = x.t[2];
}

__device__ __forceinline__ PoseProgram load_program(const PoseParams& p, int64_t c) {
    PoseProgram prog;
    if (p.program) {
        prog = p.program[c];
    } else {  // flat mode: one group; a state with a single clip never runs BlendingJob (:552-555, :617-624)
        const int L = p.num_layers;
        prog.n0 = (uint8_t)L; prog.n1 = 0; prog.rest0 = 0; prog.rest1 = 0;
        prog.mode0 = L == 1 ? kGroupPassthrough : kGroupBlend; prog.mode1 = kGroupPassthrough;
        prog.top = kTopGroup0; prog.reserved = 0;
        prog.threshold0 = p.threshold; prog.threshold1 = p.threshold; prog.w0 = 1.f; prog.w1 = 0.f;
    }
    return prog;
}

// ---- K0: clip clocks, one thread per (layer, character) -----------------------------------------------------------
// OzzSkeletalAnimatorAnimationState::update (src/skeleton_ozz.cpp:411-428)
__global__ void __launch_bounds__(256) advance_kernel(const PoseParams p) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p.n * p.num_layers) return;
    const int ci = p.clip[idx];
    if (ci < 0 || ci >= p.clips.num_clips) return;  // reported by the sampling kernel
    const bool loop = p.loop ? p.loop[idx] != 0 : true;
    bool looped = p.looped[idx] != 0;
    if (looped && !loop) return;
    float t = p.ratio[idx];
    float sp = p.speed[idx];
    sp = sp == 0.f ? 1.f : sp;
    const float duration = __fdiv_rn(__ldg(p.clips.duration + ci), sp);
    const float norm_time = t + __fdiv_rn(p.dt, duration);
    looped = norm_time > 1.f;
    if (!(looped && !loop)) t = fmodf(norm_time, 1.f);
    p.looped[idx] = looped ? 1 : 0;
    p.ratio[idx] = t;
}

// ---- K1: sample + blend, one thread per (character, joint) ----------------------------------------------------------
// Flat mapping: every lane of every warp carries a joint (a 67-joint skeleton would waste a third of a
// warp-per-character layout), small register footprint, many warps per SM to cover the L2 latency of the table reads.
// Output: the joint's local 3x4 matrix (SoaFloat4x4::FromAffine), 16 floats per joint so that a thread writes two whole
// sectors (256-bit stores); the model kernel reads them back coalesced.
#ifndef DMK_SAMPLE_MIN_BLOCKS
#define DMK_SAMPLE_MIN_BLOCKS 8   // 64 registers: 32 resident warps per SM hide the L2 latency of the table reads (measured best)
#endif
#ifndef DMK_SAMPLE_BLOCK
#define DMK_SAMPLE_BLOCK 128
#endif
// One (character, joint): the layers of the character's pose program sampled and blended, FromAffine -> m[12].  Returns false for
// a character without a state (kTopSkip: the animator keeps its previous matrices, src/skeleton_ozz.cpp:1029-1032).
template <int F>
__device__ __forceinline__ bool sample_local_matrix(const PoseParams& p, int64_t c, int j, float m[12]) {
    const int P = p.skel.padded_joints;
    const PoseProgram prog = load_program(p, c);
    if (prog.top == kTopSkip) return false;
    const int64_t n = p.n;
    const int L = p.num_layers;
    bool bad = false;
    auto layer = [&](int l, int& clip, float& ratio, float& weight) {
        const int64_t idx = (int64_t)l * n + c;
        int ci = __ldg(p.clip + idx);
        if (ci < 0 || ci >= p.clips.num_clips) { bad = true; ci = 0; }
        float t = p.ratio[idx];
        t = (t < 1.f) ? t : 1.f;   // SamplingJob::Run clamps the ratio to [0, 1]
        t = (t > 0.f) ? t : 0.f;
        clip = ci;
        ratio = t;
        weight = __ldg(p.weight + idx);
    };
    int32_t* kb = (F & kKeys) ? p.key_indices + (size_t)c * L * 3 * P * 2 : nullptr;
    Xform acc;
    if (prog.top == kTopRestPose) {  // load-time pose: LocalToModelJob over joint_rest_poses (:901-916)
        const float* r = p.skel.rest + (size_t)j * 10;
        acc.t[0] = r[0]; acc.t[1] = r[1]; acc.t[2] = r[2];
        acc.r[0] = r[3]; acc.r[1] = r[4]; acc.r[2] = r[5]; acc.r[3] = r[6];
        acc.s[0] = r[7]; acc.s[1] = r[8]; acc.s[2] = r[9];
    } else {
        acc = eval_group<F>(p.clips, layer, 0, prog.n0, prog.mode0, prog.rest0, prog.threshold0, j, kb, P);
        if (prog.top == kTopTransition || ((F & kKeys) && prog.n1)) {
            const Xform cur = eval_group<F>(p.clips, layer, prog.n0, prog.n1, prog.mode1, prog.rest1, prog.threshold1, j, kb, P);
            if (prog.top == kTopTransition)  // 2-layer BlendingJob, threshold bx::kFloatSmallest, rest = previous (:709-719)
                acc = blend_two(acc, cur, prog.w0, prog.w1, 1.17549435e-38f);
        }
    }
    if (bad && j == 0) atomicCAS(p.error_flag, 0, (int32_t)min(c + 1, (int64_t)0x7fffffff));
    if (p.locals) {
        float* o = p.locals + ((size_t)c * p.skel.num_soa + (j >> 2)) * kSoaFloats + (j & 3);
        o[0] = acc.t[0]; o[4] = acc.t[1]; o[8] = acc.t[2];
        o[12] = acc.r[0]; o[16] = acc.r[1]; o[20] = acc.r[2]; o[24] = acc.r[3];
        o[28] = acc.s[0]; o[32] = acc.s[1]; o[36] = acc.s[2];
    }
    from_affine(acc, m);
    return true;
}

template <int F>
__global__ void __launch_bounds__(DMK_SAMPLE_BLOCK, DMK_SAMPLE_MIN_BLOCKS) sample_kernel(const PoseParams p) {
    const int P = p.skel.padded_joints;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= p.n * P) return;
    const int64_t c = tid / P;
    const int j = (int)(tid - c * P);
    float m[12];
    if (!sample_local_matrix<F>(p, c, j, m)) return;
    float* out = p.scratch + (size_t)tid * 16;
    st_v8(out, m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7]);
    st_v8(out + 8, m[8], m[9], m[10], m[11], 0.f, 0.f, 0.f, 0.f);
}

// LocalToModelJob for one character by one warp: level order, one (joint, column) task per lane per round; s_loc / s_mod are the
// character's local and model matrices in shared memory ([P][kMS] each).
__device__ __forceinline__ void local_to_model_rounds(int rounds, const uint16_t* s_sched, const int16_t* s_parents, const float* s_identity,
                                                      const float* s_loc, float* s_mod, int lane) {
    for (int rd = 0; rd < rounds; ++rd) {
        const unsigned task = s_sched[rd * 32 + lane];
        if (task != kIdleTask) {
            const int j = task & 0xfff, col = task >> 12;
            const int parent = s_parents[j];
            const float* pm = parent < 0 ? s_identity : s_mod + (size_t)parent * kMS;
            const float* lc = s_loc + (size_t)j * kMS + col * 3;
            const float l0 = lc[0], l1 = lc[1], l2 = lc[2], l3 = col == 3 ? 1.f : 0.f;
            float* o = s_mod + (size_t)j * kMS + col * 3;
#pragma unroll
            for (int r = 0; r < 3; ++r)  // ozz Float4x4 product: (a0*x + a1*y) + (a2*z + a3*w)
                o[r] = (pm[r] * l0 + pm[3 + r] * l1) + (pm[6 + r] * l2 + pm[9 + r] * l3);
        }
        __syncwarp();
    }
}

// Palette slot k + 1 of one character: flipHandedness(model[jm]) * inverse bind (src/skeleton_ozz.cpp:962-977, src/skeleton.cpp:236-249),
// written as the column-major mat4 (pal) and / or the row-major 3x4 the skinning kernel reads (p34): 256-bit full-sector stores.
__device__ __forceinline__ void palette_slot(int jm, const float* s_mod, const float* b, float* pal, float* p34, int k) {
    float m[12];
    if (jm >= 0) {
        const float* src = s_mod + (size_t)jm * kMS;  // Math::flipHandedness: Z * M * Z
        m[0] = src[0]; m[1] = src[1]; m[2] = -src[2];
        m[3] = src[3]; m[4] = src[4]; m[5] = -src[5];
        m[6] = -src[6]; m[7] = -src[7]; m[8] = src[8];
        m[9] = src[9]; m[10] = src[10]; m[11] = -src[11];
    } else {  // unknown joint name: getJointModelMatrix returns mat4(1)
#pragma unroll
        for (int e = 0; e < 12; ++e) m[e] = (e == 0 || e == 4 || e == 8) ? 1.f : 0.f;
    }
    float v[4][3];  // [column][row]
#pragma unroll
    for (int col = 0; col < 4; ++col) {
        const float b0 = b[col * 3 + 0], b1 = b[col * 3 + 1], b2 = b[col * 3 + 2];
        const float b3 = col == 3 ? 1.f : 0.f;
#pragma unroll
        for (int r = 0; r < 3; ++r)  // glm mat4 * mat4: ((A0*b0 + A1*b1) + A2*b2) + A3*b3
            v[col][r] = ((m[r] * b0 + m[3 + r] * b1) + m[6 + r] * b2) + m[9 + r] * b3;
    }
    if (pal) {
        float* o = pal + (size_t)(k + 1) * 16;
        st_v8(o, v[0][0], v[0][1], v[0][2], 0.f, v[1][0], v[1][1], v[1][2], 0.f);
        st_v8(o + 8, v[2][0], v[2][1], v[2][2], 0.f, v[3][0], v[3][1], v[3][2], 1.f);
    }
    if (p34) {
        float* o = p34 + (size_t)(k + 1) * 16;
        st_v8(o, v[0][0], v[1][0], v[2][0], v[3][0], v[0][1], v[1][1], v[2][1], v[3][1]);
        st_v8(o + 8, v[0][2], v[1][2], v[2][2], v[3][2], 0.f, 0.f, 0.f, 0.f);
    }
}

// ---- K2 + K3: local-to-model and skin palette, one warp per character ------------------------------------------------
__global__ void __launch_bounds__(kWarpsPerBlock * 32) model_kernel(const PoseParams p) {
    DMK_DYNAMIC_SMEM(smem_raw, 16);
    const int J = p.skel.num_joints, P = p.skel.padded_joints, K = p.arm.num_joints, PS = K + 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // ---- shared tables, staged once per persistent block ----
    float* s_inv_bind = reinterpret_cast<float*>(smem_raw);                   // [K][kMS]
    float* s_identity = s_inv_bind + (size_t)K * kMS;                         // [12] (+4 pad)
    int32_t* s_remap = reinterpret_cast<int32_t*>(s_identity + 16);           // [K]
    int16_t* s_parents = reinterpret_cast<int16_t*>(s_remap + ((K + 3) & ~3)); // [P]
    uint16_t* s_sched = reinterpret_cast<uint16_t*>(s_parents + ((P + 7) & ~7)); // [rounds][32]
    const int per_warp = (2 * P * kMS + 3) & ~3;                              // local + model matrices of one character
    float* s_warp0 = reinterpret_cast<float*>(
        (reinterpret_cast<uintptr_t>(s_sched + (size_t)p.skel.rounds * 32) + 15) & ~uintptr_t(15));
    float* s_loc = s_warp0 + (size_t)warp * per_warp;  // [P][kMS] local matrices
    float* s_mod = s_loc + (size_t)P * kMS;             // [P][kMS] model matrices

    for (int i = threadIdx.x; i < K * 12; i += blockDim.x) s_inv_bind[(i / 12) * kMS + i % 12] = p.arm.inv_bind[i];
    for (int i = threadIdx.x; i < K; i += blockDim.x) s_remap[i] = p.arm.remap[i];
    for (int i = threadIdx.x; i < P; i += blockDim.x) s_parents[i] = i < J ? p.skel.parents[i] : (int16_t)-1;
    for (int i = threadIdx.x; i < p.skel.rounds * 32; i += blockDim.x) s_sched[i] = p.skel.schedule[i];
    if (threadIdx.x < 12) s_identity[threadIdx.x] = (threadIdx.x == 0 || threadIdx.x == 4 || threadIdx.x == 8) ? 1.f : 0.f;
    __syncthreads();

    const int64_t n = p.n;
    const int warps_per_block = blockDim.x >> 5;
    const int64_t warps_total = (int64_t)gridDim.x * warps_per_block;

    for (int64_t c = (int64_t)blockIdx.x * warps_per_block + warp; c < n; c += warps_total) {
        if (p.program && p.program[c].top == kTopSkip) continue;

        // ---- local matrices written by the sampling kernel: coalesced 256-bit loads -> shared memory ----
        for (int j = lane; j < J; j += 32) {
            float v[16];
            const float* src = p.scratch + ((size_t)c * P + j) * 16;
            ld_v8(src, v);
            ld_v8(src + 8, v + 8);
            float* d = s_loc + (size_t)j * kMS;
#pragma unroll
            for (int e = 0; e < 12; ++e) d[e] = v[e];
        }
        __syncwarp();

        // ---- K2: local-to-model, level order; one (joint, column) per lane per round ----
        local_to_model_rounds(p.skel.rounds, s_sched, s_parents, s_identity, s_loc, s_mod, lane);

        // ---- K3: flip handedness + palette = model * inverse bind ----
        if (p.models) {
            for (int j = lane; j < J; j += 32) {
                const float* m = s_mod + (size_t)j * kMS;
                float* o = p.models + ((size_t)c * J + j) * 16;
                o[0] = m[0]; o[1] = m[1]; o[2] = -m[2]; o[3] = 0.f;
                o[4] = m[3]; o[5] = m[4]; o[6] = -m[5]; o[7] = 0.f;
                o[8] = -m[6]; o[9] = -m[7]; o[10] = m[8]; o[11] = 0.f;
                o[12] = m[9]; o[13] = m[10]; o[14] = -m[11]; o[15] = 1.f;
            }
        }
        if (p.palette || p.palette34) {
            // Each lane owns whole palette slots and writes them as 256-bit full-sector stores (a mat4 = 2 sectors; the
            // 3x4 row-major copy for the skinning kernel is padded to 4 rows = 2 sectors as well): no staging pass.
            float* pal = p.palette ? p.palette + (size_t)c * PS * 16 : nullptr;
            float* p34 = p.palette34 ? p.palette34 + (size_t)c * PS * 16 : nullptr;
            if (lane == 0) {  // slot 0 = mat4(1) (src/skeleton.cpp:238)
                if (pal) { st_v8(pal, 1, 0, 0, 0, 0, 1, 0, 0); st_v8(pal + 8, 0, 0, 1, 0, 0, 0, 0, 1); }
                if (p34) { st_v8(p34, 1, 0, 0, 0, 0, 1, 0, 0); st_v8(p34 + 8, 0, 0, 1, 0, 0, 0, 0, 0); }
            }
            // armature joints in the ORDER the host chose (ArmatureDev): the 32 lanes of a pass read model matrices of joints that differ
            // mod 32, i.e. from 32 different banks (stride kMS is odd); the entry carries the joint and the palette slot it fills
            for (int i = lane; i < K; i += 32) {
                const int32_t e = s_remap[i];
                palette_slot((int)(int16_t)(e & 0xffff), s_mod, s_inv_bind + (size_t)i * kMS, pal, p34, e >> 16);
            }
        }
        __syncwarp();  // the matrix arrays are reused by the next character
    }
}

// ---- K1 + K2 + K3 in one kernel (round 2): no round trip of the local matrices through HBM ---------------------------------
// The two-kernel path above writes 64 B of local matrix per joint and reads them back (profiles/r01_sample_kernel.txt,
// r01_model_kernel.txt: 1.65 GB of DRAM traffic per 100k characters against 0.44 GB of algorithmic bytes).  Here a block owns
// `cpb` characters at a time:
//   phase 1  one thread per (character, joint): sampling + blend + FromAffine, as sample_kernel, into shared memory;
//   phase 2  one warp per character: LocalToModelJob, level order (local_to_model_rounds);
//   phase 3  one thread per (character, palette slot): flip + inverse bind, 256-bit stores of the palette(s).
// Same device functions, same operation order: results are bit-identical to the two-kernel path.  MEASURED (B200, 100k characters,
// profiles/r02_pose_fused_kernel.txt): DRAM traffic 0.40 GB instead of 1.65 GB, but 742 us instead of 613 us -- the path is
// latency bound (barrier 2.8 and long-scoreboard 3.6 stalled warps per issued instruction at 27 resident warps), not bandwidth
// bound, so the two kernels stay the default and this one is opt-in (DMK_OPT_POSE_FUSED).
constexpr int kFusedMaxThreads = 288;   // 4 characters x 68 padded joints = 272 threads: 9 warps, 16 idle lanes
#ifndef DMK_FUSED_MIN_BLOCKS
#define DMK_FUSED_MIN_BLOCKS 3
#endif
template <int F>
__global__ void __launch_bounds__(kFusedMaxThreads, DMK_FUSED_MIN_BLOCKS) pose_fused_kernel(const PoseParams p, int cpb) {
    DMK_DYNAMIC_SMEM(smem_raw, 16);
    const int J = p.skel.num_joints, P = p.skel.padded_joints, K = p.arm.num_joints, PS = K + 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    float* s_inv_bind = reinterpret_cast<float*>(smem_raw);                   // [K][kMS]
    float* s_identity = s_inv_bind + (size_t)K * kMS;                         // [12] (+4 pad)
    int32_t* s_remap = reinterpret_cast<int32_t*>(s_identity + 16);           // [K]
    int16_t* s_parents = reinterpret_cast<int16_t*>(s_remap + ((K + 3) & ~3)); // [P]
    uint16_t* s_sched = reinterpret_cast<uint16_t*>(s_parents + ((P + 7) & ~7)); // [rounds][32]
    const int per_char = (2 * P * kMS + 3) & ~3;                              // local + model matrices of one character
    float* s_char0 = reinterpret_cast<float*>(
        (reinterpret_cast<uintptr_t>(s_sched + (size_t)p.skel.rounds * 32) + 15) & ~uintptr_t(15));
    uint8_t* s_live = reinterpret_cast<uint8_t*>(s_char0 + (size_t)cpb * per_char);   // [cpb] character has a state this frame

    for (int i = threadIdx.x; i < K * 12; i += blockDim.x) s_inv_bind[(i / 12) * kMS + i % 12] = p.arm.inv_bind[i];
    for (int i = threadIdx.x; i < K; i += blockDim.x) s_remap[i] = p.arm.remap[i];
    for (int i = threadIdx.x; i < P; i += blockDim.x) s_parents[i] = i < J ? p.skel.parents[i] : (int16_t)-1;
    for (int i = threadIdx.x; i < p.skel.rounds * 32; i += blockDim.x) s_sched[i] = p.skel.schedule[i];
    if (threadIdx.x < 12) s_identity[threadIdx.x] = (threadIdx.x == 0 || threadIdx.x == 4 || threadIdx.x == 8) ? 1.f : 0.f;
    __syncthreads();

    const int64_t n = p.n;
    const int slot = threadIdx.x / P, j = threadIdx.x - slot * P;   // phase 1: this thread's (character of the group, joint)
    const int64_t groups = (n + cpb - 1) / cpb;
    for (int64_t g = blockIdx.x; g < groups; g += gridDim.x) {
        const int64_t c0 = g * cpb;
        // ---- phase 1: sample + blend + FromAffine -> shared memory ----
        if (slot < cpb && c0 + slot < n) {
            float m[12];
            const bool live = sample_local_matrix<F>(p, c0 + slot, j, m);
            if (j == 0) s_live[slot] = live ? 1 : 0;
            if (live) {
                float* d = s_char0 + (size_t)slot * per_char + (size_t)j * kMS;
#pragma unroll
                for (int e = 0; e < 12; ++e) d[e] = m[e];
            }
        }
        __syncthreads();
        // ---- phase 2: local-to-model, one warp per character ----
        for (int s2 = warp; s2 < cpb; s2 += (blockDim.x >> 5)) {
            if (c0 + s2 >= n || !s_live[s2]) continue;
            float* s_loc = s_char0 + (size_t)s2 * per_char;
            local_to_model_rounds(p.skel.rounds, s_sched, s_parents, s_identity, s_loc, s_loc + (size_t)P * kMS, lane);
        }
        __syncthreads();
        // ---- phase 3: models (optional) and palette slots, one thread per (character, joint / slot) ----
        if (p.models) {
            for (int idx = threadIdx.x; idx < cpb * J; idx += blockDim.x) {
                const int s3 = idx / J, jj = idx - s3 * J;
                if (c0 + s3 >= n || !s_live[s3]) continue;
                const float* m = s_char0 + (size_t)s3 * per_char + (size_t)P * kMS + (size_t)jj * kMS;
                float* o = p.models + ((size_t)(c0 + s3) * J + jj) * 16;
                o[0] = m[0]; o[1] = m[1]; o[2] = -m[2]; o[3] = 0.f;
                o[4] = m[3]; o[5] = m[4]; o[6] = -m[5]; o[7] = 0.f;
                o[8] = -m[6]; o[9] = -m[7]; o[10] = m[8]; o[11] = 0.f;
                o[12] = m[9]; o[13] = m[10]; o[14] = -m[11]; o[15] = 1.f;
            }
        }
        if (p.palette || p.palette34) {
            for (int idx = threadIdx.x; idx < cpb * PS; idx += blockDim.x) {
                const int s3 = idx / PS, k = idx - s3 * PS - 1;   // k = -1: slot 0 = mat4(1) (src/skeleton.cpp:238)
                if (c0 + s3 >= n || !s_live[s3]) continue;
                float* pal = p.palette ? p.palette + (size_t)(c0 + s3) * PS * 16 : nullptr;
                float* p34 = p.palette34 ? p.palette34 + (size_t)(c0 + s3) * PS * 16 : nullptr;
                if (k < 0) {
                    if (pal) { st_v8(pal, 1, 0, 0, 0, 0, 1, 0, 0); st_v8(pal + 8, 0, 0, 1, 0, 0, 0, 0, 1); }
                    if (p34) { st_v8(p34, 1, 0, 0, 0, 0, 1, 0, 0); st_v8(p34 + 8, 0, 0, 1, 0, 0, 0, 0, 0); }
                } else {
                    const int32_t e = s_remap[k];   // (entry k of the host's processing order: joint | palette slot << 16)
                    palette_slot((int)(int16_t)(e & 0xffff), s_char0 + (size_t)s3 * per_char + (size_t)P * kMS, s_inv_bind + (size_t)k * kMS, pal, p34, e >> 16);
                }
            }
        }
        __syncthreads();  // the matrix arrays are reused by the next group
    }
}

}  // namespace

static size_t fused_smem_bytes_for(const PoseParams& p, int cpb) {
    const int P = p.skel.padded_joints, K = p.arm.num_joints;
    size_t b = 0;
    b += (size_t)K * kMS * 4 + 16 * 4;
    b += (size_t)((K + 3) & ~3) * 4;
    b += (size_t)((P + 7) & ~7) * 2;
    b += (size_t)p.skel.rounds * 32 * 2;
    b += 16;
    b += (size_t)((2 * P * kMS + 3) & ~3) * 4 * cpb;
    b += (size_t)((cpb + 15) & ~15);
    return b;
}

static size_t pose_smem_bytes_for(const PoseParams& p, int warps) {
    const int P = p.skel.padded_joints, K = p.arm.num_joints;
    size_t b = 0;
    b += (size_t)K * kMS * 4 + 16 * 4;
    b += (size_t)((K + 3) & ~3) * 4;
    b += (size_t)((P + 7) & ~7) * 2;
    b += (size_t)p.skel.rounds * 32 * 2;
    b += 16;
    b += (size_t)((2 * P * kMS + 3) & ~3) * 4 * warps;
    return b;
}

size_t pose_smem_bytes(const PoseParams& p) { return pose_smem_bytes_for(p, kWarpsPerBlock); }

cudaError_t launch_pose(const PoseParams& p, int num_sms, cudaStream_t stream, int* launches) {
    if (p.n <= 0) return cudaSuccess;
    if (p.advance) {
        const int64_t total = p.n * p.num_layers;
        DMK_LAUNCH(advance_kernel, (unsigned)((total + 255) / 256), 256, 0, stream)(p);
        if (launches) ++*launches;
    }
    // fused sampling + local-to-model + palette where a group of characters fits a block (skeletons up to 288 padded joints)
    if (p.fused && p.skel.padded_joints <= kFusedMaxThreads) {
        const int P = p.skel.padded_joints;
        int cpb = kFusedMaxThreads / P;
        if (cpb > 8) cpb = 8;
        const int threads = (cpb * P + 31) / 32 * 32;
        const size_t smem = fused_smem_bytes_for(p, cpb);
        const int shape = (p.clips.bucket_dir ? kBuckets : 0) | (p.key_indices ? kKeys : 0);
        auto go = [&](auto kernel) {
            cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            int per_sm = 0;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem);
            if (e != cudaSuccess) return e;
            if (per_sm < 1) return cudaErrorInvalidConfiguration;
            int64_t blocks = (p.n + cpb - 1) / cpb;
            const int64_t cap = (int64_t)num_sms * per_sm;
            if (blocks > cap) blocks = cap;
            DMK_LAUNCH(kernel, (unsigned)blocks, threads, smem, stream)(p, cpb);
            if (launches) ++*launches;
            return cudaGetLastError();
        };
        switch (shape) {
        case 0: return go(pose_fused_kernel<0>);
        case kBuckets: return go(pose_fused_kernel<kBuckets>);
        case kKeys: return go(pose_fused_kernel<kKeys>);
        default: return go(pose_fused_kernel<kBuckets | kKeys>);
        }
    }
    {
        const int64_t total = p.n * p.skel.padded_joints;
        constexpr int kBlock = DMK_SAMPLE_BLOCK;
        const unsigned grid = (unsigned)((total + kBlock - 1) / kBlock);
        const int shape = (p.clips.bucket_dir ? kBuckets : 0) | (p.key_indices ? kKeys : 0);
        switch (shape) {
        case 0: DMK_LAUNCH(sample_kernel<0>, grid, kBlock, 0, stream)(p); break;
        case kBuckets: DMK_LAUNCH(sample_kernel<kBuckets>, grid, kBlock, 0, stream)(p); break;
        case kKeys: DMK_LAUNCH(sample_kernel<kKeys>, grid, kBlock, 0, stream)(p); break;
        default: DMK_LAUNCH(sample_kernel<kBuckets | kKeys>, grid, kBlock, 0, stream)(p); break;
        }
        if (launches) ++*launches;
    }
    return launch_model(p, num_sms, stream, launches);
}

// local-to-model + flip + palette from the local matrices a sampling launch left in p.scratch: the second half of the pose
// path, also run on its own for every further armature of a crowd (one Armature per Skinnable mesh, src/skeleton.cpp:223-251)
cudaError_t launch_model(const PoseParams& p, int num_sms, cudaStream_t stream, int* launches) {
    if (p.n <= 0) return cudaSuccess;
    int warps = kWarpsPerBlock;
    while (warps > 1 && pose_smem_bytes_for(p, warps) > 200 * 1024) warps >>= 1;
    const size_t smem = pose_smem_bytes_for(p, warps);
    cudaError_t e = cudaFuncSetAttribute(model_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, model_kernel, warps * 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorInvalidConfiguration;
    int64_t blocks = (p.n + warps - 1) / warps;
    const int64_t cap = (int64_t)num_sms * per_sm;
    if (blocks > cap) blocks = cap;
    DMK_LAUNCH(model_kernel, (unsigned)blocks, warps * 32, smem, stream)(p);
    if (launches) ++*launches;
    return cudaGetLastError();
}

}  // namespace dmk
