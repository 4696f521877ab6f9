// skin_kernel.cu -- K5: 4-influence linear-blend skinning of positions, normals and tangents for a whole crowd.
//
// Replaces the LBS vertex stage of the reference (paths relative to the darmok repository):
//   applySkinning / getSkinningMatrix   assets/shaders/core/darmok_skinning.inc.slang:5-28
//   pos(w=1) / normal(w=0) / tangent(w=0) all through the same blended matrix   assets/shaders/core/forward.slang:33-38
//   influence packing (indices -1 -> palette slot 0 = identity, weight 0)        src/mesh_core.cpp:336-356
//
// Shape of the kernel (see DESIGN.md "K5"):
//   * HBM traffic is a pure write stream (36 B/vertex out, 0.65 B/vertex in); the on-chip limit is the gather of
//     4 palette matrices per vertex from shared memory (192 B/vertex against 36 B/vertex of HBM writes).
//   * One persistent CTA per SM owns a TILE of the (shared) mesh: each thread keeps 8 consecutive vertices --
//     position, normal, tangent, weights, packed palette slots -- in registers for the whole launch and loops over
//     characters, so bind-pose data is read once per CTA, not once per character.
//   * Per character the 3x4 row-major palette (written by the pose kernel) arrives by cp.async.bulk (TMA bulk
//     copy) into a double-buffered staging area, signalled by an mbarrier, one character ahead.
//   * The palette is then replicated 8x into bank-group-private copies: lane L reads copy (L & 7), whose 16-byte
//     chunks all live in bank group (L & 7), so the random per-vertex gathers (LDS.128) are conflict free
//     whatever the bone indices are.
//   * 8 vertices x 36 B = 288 B = 9 full 32-byte sectors per thread: results leave as 256-bit stores
//     (st.global.v8.f32 -> STG.E.256), every store writing whole sectors.
#include "device_types.cuh"
#include "kernels.cuh"

namespace dmk {

namespace {

constexpr int kVPT = 8;           // vertices per thread: lcm(36 B, 32 B) / 36 B
constexpr int kMaxThreads = 320;  // 65536 registers / 204
constexpr int kRepl = 8;          // palette replicas = bank groups of a 16-byte access

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA bulk copy global -> shared, completion counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void st_v8(float* p, const float* v) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
                 "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}

__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

template <int REPL>
__global__ void __launch_bounds__(kMaxThreads, 1) skin_kernel(const SkinParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int PS = p.palette_size;
    const uint32_t pal_bytes = (uint32_t)PS * 48u;             // 3x4 floats per slot
    const uint32_t repl_bytes = pal_bytes * REPL;
    unsigned char* s_repl = smem;                               // [2][PS*3][REPL] float4
    unsigned char* s_stage = smem + 2 * (size_t)repl_bytes;     // [2][PS*3] float4 (bulk copy destination)
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_stage + 2 * (size_t)pal_bytes);

    const int tid = threadIdx.x, lane = tid & 31;
    const int tiles_per_pass = min(p.num_tiles, (int)gridDim.x);
    const int ngroups = gridDim.x / tiles_per_pass;
    const int tslot = blockIdx.x % tiles_per_pass, group = blockIdx.x / tiles_per_pass;
    if (group >= ngroups) return;

    int64_t nchars = p.n;
    if (p.char_ids) nchars = p.char_count[0];
    const int64_t my_count = nchars > group ? (nchars - group + ngroups - 1) / ngroups : 0;
    if (my_count == 0) return;

    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    uint32_t issued = 0;  // bulk copies issued so far by thread 0 across tile passes (buffer = issued & 1)
    uint32_t consumed = 0;

    for (int tile = tslot; tile < p.num_tiles; tile += tiles_per_pass) {
        // ---- this thread's 8 vertices: loaded once, kept in registers for every character ----
        const int v0 = tile * p.tile_vertices + tid * kVPT;
        const bool active = tid * kVPT < p.tile_vertices && v0 < p.vpad;
        float px[kVPT], py[kVPT], pz[kVPT], nx[kVPT], ny[kVPT], nz[kVPT], tx[kVPT], ty[kVPT], tz[kVPT];
        float w0[kVPT], w1[kVPT], w2[kVPT], w3[kVPT];
        uint32_t slots[kVPT];
#pragma unroll
        for (int v = 0; v < kVPT; ++v) {
            const int vi = active ? v0 + v : 0;
            px[v] = __ldg(p.pos + (size_t)vi * 3 + 0); py[v] = __ldg(p.pos + (size_t)vi * 3 + 1); pz[v] = __ldg(p.pos + (size_t)vi * 3 + 2);
            nx[v] = __ldg(p.nrm + (size_t)vi * 3 + 0); ny[v] = __ldg(p.nrm + (size_t)vi * 3 + 1); nz[v] = __ldg(p.nrm + (size_t)vi * 3 + 2);
            tx[v] = __ldg(p.tan + (size_t)vi * 3 + 0); ty[v] = __ldg(p.tan + (size_t)vi * 3 + 1); tz[v] = __ldg(p.tan + (size_t)vi * 3 + 2);
            const float4 w = __ldg(reinterpret_cast<const float4*>(p.weights) + vi);
            w0[v] = w.x; w1[v] = w.y; w2[v] = w.z; w3[v] = w.w;
            slots[v] = __ldg(p.slots + vi);
        }

        auto char_of = [&](int64_t i) -> int64_t {
            const int64_t k = group + i * ngroups;
            return p.char_ids ? (int64_t)__ldg(p.char_ids + k) : k;
        };
        auto issue = [&](int64_t i) {  // thread 0 only
            if (i < my_count) {
                const uint32_t b = issued & 1u;
                mbar_expect_tx(&s_bar[b], pal_bytes);
                bulk_g2s(s_stage + (size_t)b * pal_bytes, p.palette34 + (size_t)char_of(i) * PS * 12, pal_bytes, &s_bar[b]);
                ++issued;
            }
        };
        // staging -> 8 bank-group-private replicas; lane L writes replica (L + k) & 7 at step k: conflict free
        auto replicate = [&](uint32_t b) {
            const float4* src = reinterpret_cast<const float4*>(s_stage + (size_t)b * pal_bytes);
            float4* dst = reinterpret_cast<float4*>(s_repl + (size_t)b * repl_bytes);
            for (int t = tid; t < PS * 3; t += blockDim.x) {
                const float4 val = src[t];
#pragma unroll
                for (int k = 0; k < REPL; ++k) dst[t * REPL + ((lane + k) & (REPL - 1))] = val;
            }
        };

        if (tid == 0) { issue(0); issue(1); }
        mbar_wait(&s_bar[consumed & 1u], (consumed >> 1) & 1u);
        replicate(consumed & 1u);
        ++consumed;
        __syncthreads();
        if (tid == 0) issue(2);

        const uint32_t lane_off = (uint32_t)(lane & (REPL - 1)) * 16u;
        for (int64_t i = 0; i < my_count; ++i) {
            // buffer of character i within this tile pass: the one replicated last
            const uint32_t cur = (consumed - 1u) & 1u;
            if (active) {
                const uint32_t base = smem_u32(s_repl + (size_t)cur * repl_bytes) + lane_off;
                float* out = p.out + ((size_t)char_of(i) * p.vpad + v0) * 9;
                float res[kVPT * 9];
#pragma unroll
                for (int v = 0; v < kVPT; ++v) {
                    const uint32_t s = slots[v];
                    const uint32_t a0 = base + (s & 0xffu) * (48u * REPL);
                    const uint32_t a1 = base + ((s >> 8) & 0xffu) * (48u * REPL);
                    const uint32_t a2 = base + ((s >> 16) & 0xffu) * (48u * REPL);
                    const uint32_t a3 = base + (s >> 24) * (48u * REPL);
                    float m[3][4];
#pragma unroll
                    for (int r = 0; r < 3; ++r) {  // row r of M = sum_i w_i * P[slot_i]
                        const float4 q0 = lds128(a0 + r * 16u * REPL), q1 = lds128(a1 + r * 16u * REPL);
                        const float4 q2 = lds128(a2 + r * 16u * REPL), q3 = lds128(a3 + r * 16u * REPL);
                        m[r][0] = fmaf(w3[v], q3.x, fmaf(w2[v], q2.x, fmaf(w1[v], q1.x, w0[v] * q0.x)));
                        m[r][1] = fmaf(w3[v], q3.y, fmaf(w2[v], q2.y, fmaf(w1[v], q1.y, w0[v] * q0.y)));
                        m[r][2] = fmaf(w3[v], q3.z, fmaf(w2[v], q2.z, fmaf(w1[v], q1.z, w0[v] * q0.z)));
                        m[r][3] = fmaf(w3[v], q3.w, fmaf(w2[v], q2.w, fmaf(w1[v], q1.w, w0[v] * q0.w)));
                    }
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        res[v * 9 + r] = fmaf(m[r][2], pz[v], fmaf(m[r][1], py[v], fmaf(m[r][0], px[v], m[r][3])));
                        res[v * 9 + 3 + r] = fmaf(m[r][2], nz[v], fmaf(m[r][1], ny[v], m[r][0] * nx[v]));
                        res[v * 9 + 6 + r] = fmaf(m[r][2], tz[v], fmaf(m[r][1], ty[v], m[r][0] * tx[v]));
                    }
                    // flush every 32-byte sector that just became complete
#pragma unroll
                    for (int sct = (v * 9) / 8; sct < ((v + 1) * 9) / 8; ++sct) st_v8(out + sct * 8, res + sct * 8);
                }
            }
            if (i + 1 < my_count) {
                mbar_wait(&s_bar[consumed & 1u], (consumed >> 1) & 1u);
                replicate(consumed & 1u);
                ++consumed;
            }
            __syncthreads();
            if (tid == 0) issue(i + 3);
        }
        // tile pass boundary: every issued copy has been consumed (issue() stops at my_count)
    }
}

}  // namespace

SkinPlan plan_skin(int vpad, int palette_size, int num_sms) {
    SkinPlan plan{};
    const int max_tile = kMaxThreads * kVPT;
    plan.num_tiles = (vpad + max_tile - 1) / max_tile;
    int tv = (vpad + plan.num_tiles - 1) / plan.num_tiles;
    tv = (tv + kVPT - 1) / kVPT * kVPT;
    plan.tile_vertices = tv;
    plan.threads = ((tv / kVPT) + 31) / 32 * 32;
    plan.replicas = kRepl;
    plan.smem = (size_t)2 * palette_size * 48 * (kRepl + 1) + 16;
    int occ = 1;
    cudaFuncSetAttribute(skin_kernel<kRepl>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, skin_kernel<kRepl>, plan.threads, plan.smem) != cudaSuccess || occ < 1)
        occ = 1;
    int grid = num_sms * occ;
    const int tpp = plan.num_tiles < grid ? plan.num_tiles : grid;
    grid = grid / tpp * tpp;
    plan.grid = grid;
    return plan;
}

cudaError_t launch_skin(const SkinParams& p, const SkinPlan& plan, cudaStream_t stream) {
    if (p.n <= 0 || p.vpad <= 0) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(skin_kernel<kRepl>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem);
    if (e != cudaSuccess) return e;
    skin_kernel<kRepl><<<plan.grid, plan.threads, plan.smem, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace dmk
