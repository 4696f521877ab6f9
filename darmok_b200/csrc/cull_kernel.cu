// cull_kernel.cu -- K4: FrustumCuller::updateCulled for a whole crowd, one thread per instance.
//
// Replaces (paths relative to the darmok repository):
//   FrustumCuller::updateCulled       src/culling.cpp:264-285
//   Frustum::operator*=               src/shape.cpp:1317-1324   (camera frustum corners -> entity local space)
//   Frustum::getPlanes / getPlane     src/shape.cpp:1235-1272   (6 planes from corner triples)
//   Plane(Triangle), signedDistanceTo src/shape.cpp:423-427, :449-452
//   Plane::isInFront(BoundingBox)     src/shape.cpp:459-469,  BoundingBox::getCorners :1057-1069
//   Frustum::canSee                   src/shape.cpp:1213-1223
//
// Visibility must be bit-exact with the CPU: this TU is compiled with -fmad=false, uses IEEE sqrt/div and follows
// glm's scalar operation order.  The only algebraic shortcut is exact: "all 8 box corners
// have signed distance > 0" is evaluated on the corner that minimises every rounded product -- IEEE rounding is
// monotonic, so fl(fl(px+py)+pz)-d of that corner is the minimum over the 8 corners (NaN corners never fail the
// reference's `<= 0` test, and fminf drops them the same way).
#include "device_intrinsics.cuh"
#include "device_types.cuh"
#include "kernels.cuh"

namespace dmk {

namespace {

struct V3 { float x, y, z; };

__device__ __forceinline__ V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ float dot3(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ V3 cross3(V3 x, V3 y) {
    return {x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y};
}
__device__ __forceinline__ V3 normalize3(V3 v) {
    const float inv = __frcp_rn(__fsqrt_rn(dot3(v, v)));   // correctly rounded 1 / sqrt: glm::inversesqrt = 1 / sqrt(x)
    return {v.x * inv, v.y * inv, v.z * inv};
}

// true when plane(a, b, c) has the whole box strictly in front (Plane::isInFront(BoundingBox))
__device__ __forceinline__ bool box_in_front(V3 a, V3 b, V3 c, V3 mn, V3 mx) {
    const V3 n = normalize3(cross3(sub(b, a), sub(c, a)));   // Plane(Triangle): normal
    const float d = dot3(n, a);                              //                  distance
    const V3 nn = normalize3(n);                             // signedDistanceTo normalises again
    const float px = fminf(nn.x * mn.x, nn.x * mx.x);
    const float py = fminf(nn.y * mn.y, nn.y * mx.y);
    const float pz = fminf(nn.z * mn.z, nn.z * mx.z);
    const float sd_min = ((px + py) + pz) - d;
    return !(sd_min <= 0.f);
}

// Transform::update (src/transform.cpp:286-309): _localMatrix = glm::translate(p) * glm::mat4_cast(q) * glm::scale(s),
// _localInverse = glm::inverse(_localMatrix), all in glm's scalar operation order.  m[column][row].
// The two products only add exact zeros to R[c][r] * s[c] and to p, so the local matrix is [R * diag(s) | p].
__device__ __forceinline__ void local_from_values(float px, float py, float pz, float4 q, float sx, float sy, float sz, float m[4][4]) {
    const float qxx = q.x * q.x, qyy = q.y * q.y, qzz = q.z * q.z, qxz = q.x * q.z, qxy = q.x * q.y, qyz = q.y * q.z;
    const float qwx = q.w * q.x, qwy = q.w * q.y, qwz = q.w * q.z;
    m[0][0] = (1.f - 2.f * (qyy + qzz)) * sx; m[0][1] = (2.f * (qxy + qwz)) * sx;       m[0][2] = (2.f * (qxz - qwy)) * sx;       m[0][3] = 0.f;
    m[1][0] = (2.f * (qxy - qwz)) * sy;       m[1][1] = (1.f - 2.f * (qxx + qzz)) * sy; m[1][2] = (2.f * (qyz + qwx)) * sy;       m[1][3] = 0.f;
    m[2][0] = (2.f * (qxz + qwy)) * sz;       m[2][1] = (2.f * (qyz - qwx)) * sz;       m[2][2] = (1.f - 2.f * (qxx + qyy)) * sz; m[2][3] = 0.f;
    m[3][0] = px; m[3][1] = py; m[3][2] = pz; m[3][3] = 1.f;
}
__device__ __forceinline__ void local_from_trs(const float* __restrict__ pos, const float* __restrict__ rot, const float* __restrict__ scl, int64_t i,
                                               float m[4][4]) {
    local_from_values(__ldg(pos + i * 3), __ldg(pos + i * 3 + 1), __ldg(pos + i * 3 + 2), __ldg(reinterpret_cast<const float4*>(rot) + i),
                      __ldg(scl + i * 3), __ldg(scl + i * 3 + 1), __ldg(scl + i * 3 + 2), m);
}

// glm::detail::compute_inverse<4, 4>
__device__ __forceinline__ void inverse_glm(const float m[4][4], float o[4][4]) {
    const float c00 = m[2][2] * m[3][3] - m[3][2] * m[2][3], c02 = m[1][2] * m[3][3] - m[3][2] * m[1][3];
    const float c03 = m[1][2] * m[2][3] - m[2][2] * m[1][3], c04 = m[2][1] * m[3][3] - m[3][1] * m[2][3];
    const float c06 = m[1][1] * m[3][3] - m[3][1] * m[1][3], c07 = m[1][1] * m[2][3] - m[2][1] * m[1][3];
    const float c08 = m[2][1] * m[3][2] - m[3][1] * m[2][2], c10 = m[1][1] * m[3][2] - m[3][1] * m[1][2];
    const float c11 = m[1][1] * m[2][2] - m[2][1] * m[1][2], c12 = m[2][0] * m[3][3] - m[3][0] * m[2][3];
    const float c14 = m[1][0] * m[3][3] - m[3][0] * m[1][3], c15 = m[1][0] * m[2][3] - m[2][0] * m[1][3];
    const float c16 = m[2][0] * m[3][2] - m[3][0] * m[2][2], c18 = m[1][0] * m[3][2] - m[3][0] * m[1][2];
    const float c19 = m[1][0] * m[2][2] - m[2][0] * m[1][2], c20 = m[2][0] * m[3][1] - m[3][0] * m[2][1];
    const float c22 = m[1][0] * m[3][1] - m[3][0] * m[1][1], c23 = m[1][0] * m[2][1] - m[2][0] * m[1][1];
    const float f0[4] = {c00, c00, c02, c03}, f1[4] = {c04, c04, c06, c07}, f2[4] = {c08, c08, c10, c11};
    const float f3[4] = {c12, c12, c14, c15}, f4[4] = {c16, c16, c18, c19}, f5[4] = {c20, c20, c22, c23};
    const float v0[4] = {m[1][0], m[0][0], m[0][0], m[0][0]}, v1[4] = {m[1][1], m[0][1], m[0][1], m[0][1]};
    const float v2[4] = {m[1][2], m[0][2], m[0][2], m[0][2]}, v3[4] = {m[1][3], m[0][3], m[0][3], m[0][3]};
    float inv[4][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float sa = (k & 1) ? -1.f : 1.f, sb = -sa;
        inv[0][k] = ((v1[k] * f0[k] - v2[k] * f1[k]) + v3[k] * f2[k]) * sa;
        inv[1][k] = ((v0[k] * f0[k] - v2[k] * f3[k]) + v3[k] * f4[k]) * sb;
        inv[2][k] = ((v0[k] * f1[k] - v1[k] * f3[k]) + v3[k] * f5[k]) * sa;
        inv[3][k] = ((v0[k] * f2[k] - v1[k] * f4[k]) + v2[k] * f5[k]) * sb;
    }
    const float d0 = m[0][0] * inv[0][0], d1 = m[0][1] * inv[1][0], d2 = m[0][2] * inv[2][0], d3 = m[0][3] * inv[3][0];
    const float ood = __frcp_rn((d0 + d1) + (d2 + d3));
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int r = 0; r < 4; ++r) o[c][r] = inv[c][r] * ood;
}

// glm::mat4 * glm::mat4: Result[j] = ((A[0] * B[j][0] + A[1] * B[j][1]) + A[2] * B[j][2]) + A[3] * B[j][3]; o may alias a or b
__device__ __forceinline__ void mul_glm(const float a[4][4], const float b[4][4], float o[4][4]) {
    float r[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) r[j][i] = ((a[0][i] * b[j][0] + a[1][i] * b[j][1]) + a[2][i] * b[j][2]) + a[3][i] * b[j][3];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) o[j][i] = r[j][i];
}

// Transform::update of the group nodes the crowd's entities hang under (src/transform.cpp:300-305):
//   world = parent.world * local,   worldInverse = localInverse * parent.worldInverse,   the parent first.
// One thread per node walks its ancestor chain and multiplies from the root down: the same products in the same order as
// the reference's recursion, recomputed per node instead of shared (a few hundred nodes against 1e5..1e7 entities).
constexpr int kMaxTransformDepth = 32;

__global__ void __launch_bounds__(128) transform_nodes_kernel(const float* __restrict__ pos, const float* __restrict__ rot,
                                                              const float* __restrict__ scl, const int32_t* __restrict__ parent, int num_nodes,
                                                              float* __restrict__ world, float* __restrict__ world_inverse) {
    const int node = blockIdx.x * blockDim.x + threadIdx.x;
    if (node >= num_nodes) return;
    int chain[kMaxTransformDepth];
    int depth = 0;
    for (int k = node; k >= 0 && depth < kMaxTransformDepth; k = parent[k]) chain[depth++] = k;
    float w[4][4], wi[4][4], local[4][4], local_inverse[4][4];
    local_from_trs(pos, rot, scl, chain[depth - 1], w);
    inverse_glm(w, wi);
    for (int d = depth - 2; d >= 0; --d) {
        local_from_trs(pos, rot, scl, chain[d], local);
        inverse_glm(local, local_inverse);
        mul_glm(w, local, w);
        mul_glm(local_inverse, wi, wi);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        reinterpret_cast<float4*>(world + (size_t)node * 16)[c] = make_float4(w[c][0], w[c][1], w[c][2], w[c][3]);
        reinterpret_cast<float4*>(world_inverse + (size_t)node * 16)[c] = make_float4(wi[c][0], wi[c][1], wi[c][2], wi[c][3]);
    }
}

__device__ __forceinline__ void world_inverse_from_trs(const CullParams& p, int64_t i, float4& o0, float4& o1, float4& o2, float4& o3) {
    float m[4][4], inv[4][4];
    local_from_trs(p.position, p.rotation, p.scale, i, m);
    inverse_glm(m, inv);
    if (p.entity_parent) {   // uniform branch: the crowd either hangs under group nodes or not
        const int node = __ldg(p.entity_parent + i);
        if (node >= 0) {
            float pw[4][4];
            const float4* src = reinterpret_cast<const float4*>(p.node_world_inverse + (size_t)node * 16);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float4 v = __ldg(src + c);
                pw[c][0] = v.x; pw[c][1] = v.y; pw[c][2] = v.z; pw[c][3] = v.w;
            }
            mul_glm(inv, pw, inv);
        }
    }
    o0 = make_float4(inv[0][0], inv[0][1], inv[0][2], inv[0][3]);
    o1 = make_float4(inv[1][0], inv[1][1], inv[1][2], inv[1][3]);
    o2 = make_float4(inv[2][0], inv[2][1], inv[2][2], inv[2][3]);
    o3 = make_float4(inv[3][0], inv[3][1], inv[3][2], inv[3][3]);
}

template <bool FROM_TRS>
__global__ void __launch_bounds__(256) cull_kernel(const CullParams p) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool visible = false;
    if (i < p.n) {
        float4 c0, c1, c2, c3;
        if (FROM_TRS) {
            world_inverse_from_trs(p, i, c0, c1, c2, c3);
            if (p.world_inverse_out) {
                float4* o = reinterpret_cast<float4*>(p.world_inverse_out + i * 16);
                o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
            }
        } else {
            const float4* m4 = reinterpret_cast<const float4*>(p.world_inverse + i * 16);
            c0 = __ldg(m4 + 0); c1 = __ldg(m4 + 1); c2 = __ldg(m4 + 2); c3 = __ldg(m4 + 3);
        }
        const float2* bb = reinterpret_cast<const float2*>(p.bbox + i * 6);
        const float2 b0 = __ldg(bb + 0), b1 = __ldg(bb + 1), b2 = __ldg(bb + 2);
        const V3 mn = {b0.x, b0.y, b1.x}, mx = {b1.y, b2.x, b2.y};
        V3 c[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            // trans * vec4(corner, 1): (m0*x + m1*y) + (m2*z + m3*w), truncated to vec3
            const float x = p.corners[k * 3 + 0], y = p.corners[k * 3 + 1], z = p.corners[k * 3 + 2];
            c[k].x = (c0.x * x + c1.x * y) + (c2.x * z + c3.x * 1.f);
            c[k].y = (c0.y * x + c1.y * y) + (c2.y * z + c3.y * 1.f);
            c[k].z = (c0.z * x + c1.z * y) + (c2.z * z + c3.z * 1.f);
        }
        // Frustum::CornerType: NBL 0, NBR 1, NTL 2, NTR 3, FBL 4, FBR 5, FTL 6, FTR 7
        bool culled = box_in_front(c[3], c[1], c[0], mn, mx);        // Near:   NTR NBR NBL
        culled = culled || box_in_front(c[6], c[4], c[5], mn, mx);   // Far:    FTL FBL FBR
        culled = culled || box_in_front(c[4], c[0], c[1], mn, mx);   // Bottom: FBL NBL NBR
        culled = culled || box_in_front(c[7], c[3], c[2], mn, mx);   // Top:    FTR NTR NTL
        culled = culled || box_in_front(c[2], c[0], c[4], mn, mx);   // Left:   NTL NBL FBL
        culled = culled || box_in_front(c[7], c[5], c[1], mn, mx);   // Right:  FTR FBR NBR
        visible = !culled;
    }
    const unsigned bits = __ballot_sync(0xffffffffu, visible);
    if ((threadIdx.x & 31) == 0 && i < p.n) p.vis_bits[i >> 5] = bits;
}

// ---- ordered stream compaction of the visible ids from the bitmask ------------------------------------
constexpr int kScanBlock = 1024;  // words per block

__global__ void __launch_bounds__(kScanBlock) compact_count_kernel(const uint32_t* __restrict__ bits, int64_t words,
                                                                   int32_t* __restrict__ block_sums) {
    __shared__ int warp_sums[32];
    const int64_t w = (int64_t)blockIdx.x * kScanBlock + threadIdx.x;
    int v = w < words ? __popc(bits[w]) : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        int s = warp_sums[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) block_sums[blockIdx.x] = s;
    }
}

// single block: exclusive scan of the block sums in place; total -> count_out
__global__ void __launch_bounds__(1024) compact_scan_kernel(int32_t* block_sums, int num_blocks, int32_t* count_out) {
    __shared__ int warp_tot[32];
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < num_blocks; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < num_blocks ? block_sums[i] : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if ((threadIdx.x & 31) >= o) inc += t;
        }
        if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = inc;
        __syncthreads();
        if (threadIdx.x < 32) {
            int s = warp_tot[threadIdx.x];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, s, o);
                if (threadIdx.x >= o) s += t;
            }
            warp_tot[threadIdx.x] = s;  // inclusive totals of the warps
        }
        __syncthreads();
        const int warp_off = (threadIdx.x >> 5) ? warp_tot[(threadIdx.x >> 5) - 1] : 0;
        const int carry = carry_s;
        if (i < num_blocks) block_sums[i] = carry + warp_off + inc - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + warp_off + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) count_out[0] = carry_s;
}

__global__ void __launch_bounds__(kScanBlock) compact_scatter_kernel(const uint32_t* __restrict__ bits, int64_t words,
                                                                     const int32_t* __restrict__ block_offsets,
                                                                     int32_t* __restrict__ ids_out) {
    __shared__ int warp_tot[32];
    const int64_t w = (int64_t)blockIdx.x * kScanBlock + threadIdx.x;
    uint32_t word = w < words ? bits[w] : 0u;
    const int v = __popc(word);
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if ((threadIdx.x & 31) >= o) inc += t;
    }
    if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = inc;
    __syncthreads();
    if (threadIdx.x < 32) {
        int s = warp_tot[threadIdx.x];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, s, o);
            if (threadIdx.x >= o) s += t;
        }
        warp_tot[threadIdx.x] = s;
    }
    __syncthreads();
    int pos = block_offsets[blockIdx.x] + ((threadIdx.x >> 5) ? warp_tot[(threadIdx.x >> 5) - 1] : 0) + inc - v;
    while (word) {
        const int b = __ffs(word) - 1;
        word &= word - 1;
        ids_out[pos++] = (int32_t)(w * 32 + b);
    }
}

// `from34`: `palette` is the 3x4 row-major palette (same 16 floats per slot); the output is column-major mat4 either way
__global__ void pack_palettes_kernel(const float4* __restrict__ palette, int palette_vec4, const int32_t* __restrict__ ids,
                                     const int32_t* __restrict__ count, int64_t capacity, float4* __restrict__ out, int from34) {
    int64_t total = count[0];
    if (total > capacity) total = capacity;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    for (int64_t i = warp; i < total; i += warps) {
        if (from34) warp_copy_palette34_as_mat4(reinterpret_cast<const float*>(palette + (size_t)ids[i] * palette_vec4), out + (size_t)i * palette_vec4, palette_vec4, lane);
        else warp_copy_palette(palette + (size_t)ids[i] * palette_vec4, out + (size_t)i * palette_vec4, palette_vec4, lane);
    }
}

}  // namespace

cudaError_t launch_cull(const CullParams& p, cudaStream_t stream) {
    if (p.n <= 0) return cudaSuccess;
    const int64_t blocks = (p.n + 255) / 256;
    if (p.position) DMK_LAUNCH(cull_kernel<true>, (unsigned)blocks, 256, 0, stream)(p);
    else DMK_LAUNCH(cull_kernel<false>, (unsigned)blocks, 256, 0, stream)(p);
    return cudaGetLastError();
}

// SkeletalAnimatorImpl::getJointModelMatrixes(dir) (src/skeleton_ozz.cpp:979-1004): the debug-bone matrix from joint p to its
// first child, Math::transform(parentPos, glm::rotation(dir, normalize(diff)), vec3(length(diff))).  models are the
// converted (left-handed) joint matrices the crowd keeps for getJointModelMatrix.  One thread per (character, joint).
__global__ void __launch_bounds__(128) bone_matrices_kernel(const float* __restrict__ models, const int16_t* __restrict__ first_child, int num_joints,
                                                            int64_t count, float dx, float dy, float dz, float half_turn_sin, float half_turn_cos,
                                                            float* __restrict__ out) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= count * num_joints) return;
    const int64_t c = tid / num_joints;
    const int p = (int)(tid - c * num_joints);
    float4* o = reinterpret_cast<float4*>(out + tid * 16);
    const int child = first_child[p];
    if (child < 0) {
        o[0] = o[1] = o[2] = o[3] = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }
    const float4 pp = __ldg(reinterpret_cast<const float4*>(models + (c * num_joints + p) * 16) + 3);
    const float4 cp = __ldg(reinterpret_cast<const float4*>(models + (c * num_joints + child) * 16) + 3);
    const float fx = cp.x - pp.x, fy = cp.y - pp.y, fz = cp.z - pp.z;
    const float len2 = (fx * fx + fy * fy) + fz * fz;
    const float inv = __frcp_rn(__fsqrt_rn(len2));   // glm::normalize = v * inversesqrt(dot(v, v))
    const float nx = fx * inv, ny = fy * inv, nz = fz * inv;
    // glm::rotation(orig = dir, dest = n) (glm/gtx/quaternion.inl)
    const float eps = 1.1920929e-07f;
    const float cos_theta = (dx * nx + dy * ny) + dz * nz;
    float4 q;
    if (cos_theta >= 1.f - eps) {
        q = make_float4(0.f, 0.f, 0.f, 1.f);
    } else if (cos_theta < -1.f + eps) {
        float ax = 0.f * dz - dy * 1.f, ay = 1.f * dx - dz * 0.f, az = 0.f * dy - dx * 0.f;   // cross((0,0,1), dir)
        if ((ax * ax + ay * ay) + az * az < eps) {                                             // cross((1,0,0), dir)
            ax = 0.f * dz - dy * 0.f; ay = 0.f * dx - dz * 1.f; az = 1.f * dy - dx * 0.f;
        }
        const float ainv = __frcp_rn(__fsqrt_rn((ax * ax + ay * ay) + az * az));
        q = make_float4(ax * ainv * half_turn_sin, ay * ainv * half_turn_sin, az * ainv * half_turn_sin, half_turn_cos);
    } else {
        const float ax = dy * nz - ny * dz, ay = dz * nx - nz * dx, az = dx * ny - nx * dy;
        const float s = __fsqrt_rn((1.f + cos_theta) * 2.f);
        const float invs = __frcp_rn(s);
        q = make_float4(ax * invs, ay * invs, az * invs, s * 0.5f);
    }
    const float len = __fsqrt_rn(len2);
    float m[4][4];
    local_from_values(pp.x, pp.y, pp.z, q, len, len, len, m);
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = make_float4(m[k][0], m[k][1], m[k][2], m[k][3]);
}

cudaError_t launch_bone_matrices(const float* models, const int16_t* first_child, int num_joints, int64_t count, const float* dir, float half_turn_sin,
                                 float half_turn_cos, float* out, cudaStream_t stream) {
    const int64_t threads = count * num_joints;
    if (threads <= 0) return cudaSuccess;
    DMK_LAUNCH(bone_matrices_kernel, (unsigned)((threads + 127) / 128), 128, 0, stream)(models, first_child, num_joints, count, dir[0], dir[1], dir[2],
                                                                                half_turn_sin, half_turn_cos, out);
    return cudaGetLastError();
}

cudaError_t launch_transform_nodes(const float* pos, const float* rot, const float* scl, const int32_t* parent, int num_nodes, float* world,
                                   float* world_inverse, cudaStream_t stream) {
    if (num_nodes <= 0) return cudaSuccess;
    DMK_LAUNCH(transform_nodes_kernel, (num_nodes + 127) / 128, 128, 0, stream)(pos, rot, scl, parent, num_nodes, world, world_inverse);
    return cudaGetLastError();
}

size_t compact_scratch_ints(int64_t n) {
    const int64_t words = (n + 31) / 32;
    return (size_t)((words + kScanBlock - 1) / kScanBlock + 1);
}

cudaError_t launch_compact(const uint32_t* vis_bits, int64_t n, int32_t* ids_out, int32_t* count_out, int32_t* scratch,
                           cudaStream_t stream, int* launches) {
    if (n <= 0) return cudaMemsetAsync(count_out, 0, sizeof(int32_t), stream);
    const int64_t words = (n + 31) / 32;
    const int blocks = (int)((words + kScanBlock - 1) / kScanBlock);
    DMK_LAUNCH(compact_count_kernel, blocks, kScanBlock, 0, stream)(vis_bits, words, scratch);
    DMK_LAUNCH(compact_scan_kernel, 1, 1024, 0, stream)(scratch, blocks, count_out);
    DMK_LAUNCH(compact_scatter_kernel, blocks, kScanBlock, 0, stream)(vis_bits, words, scratch, ids_out);
    if (launches) *launches += 3;
    return cudaGetLastError();
}

cudaError_t launch_pack_palettes(const float* palette, int palette_floats, const int32_t* ids, const int32_t* count,
                                 int64_t capacity, float* out, int num_sms, cudaStream_t stream, bool from34) {
    DMK_LAUNCH(pack_palettes_kernel, num_sms * 4, 256, 0, stream)(reinterpret_cast<const float4*>(palette), palette_floats / 4, ids,
                                                          count, capacity, reinterpret_cast<float4*>(out), from34 ? 1 : 0);
    return cudaGetLastError();
}

}  // namespace dmk
