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
//     copy) into a double-buffered staging area, signalled by an mbarrier, two characters ahead.
//   * There is no block-wide barrier in the character loop: warps hand replica buffers over through mbarriers
//     (3-deep ring), so a warp only ever waits for the others to have STARTED the previous character.
//   * The palette is then replicated 8x into bank-group-private copies: lane L reads copy (L & 7), whose 16-byte
//     chunks all live in bank group (L & 7), so the random per-vertex gathers (LDS.128) are conflict free
//     whatever the bone indices are.
//   * 8 vertices x 36 B = 288 B = 9 full 32-byte sectors per thread: results leave as 256-bit stores
//     (st.global.v8.f32 -> STG.E.256), every store writing whole sectors.
#include "device_intrinsics.cuh"
#include "device_types.cuh"
#include "kernels.cuh"

#include <cstdlib>
#include <type_traits>

namespace dmk {

namespace {

constexpr int kVPT = 8;           // vertices per thread: lcm(36 B, 32 B) / 36 B
constexpr int kMaxThreads = 256;  // 8 warps = 2 per SM sub-partition (16K registers each): up to 255 registers per thread;
                                  // a 3rd warp per sub-partition (any block of 9..12 warps) would cap the kernel at 168
constexpr int kRepl = 8;          // palette replicas = bank groups of a 16-byte access

constexpr int kNumRepl = 3;   // replica ring: fill(i+1) may overlap compute(i) and the slowest warp's compute(i-1)
constexpr int kNumStage = 2;  // bulk-copy staging ring

// One vertex: gather 4 palette matrices row by row (row r+1 is in flight while row r is consumed), blend, apply.
struct Rows4 { float4 q0, q1, q2, q3; };

template <int REPL, bool SKIP>
__device__ __forceinline__ Rows4 load_row(uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, int r, uint32_t used) {
    Rows4 x;   // `used`: bit k set when influence k has a non-zero weight (influence 0 always is fetched)
    x.q0 = lds128(a0 + r * 16u * REPL);
    if (SKIP) {
        x.q1 = lds128_if(a1 + r * 16u * REPL, used & 2u);
        x.q2 = lds128_if(a2 + r * 16u * REPL, used & 4u);
        x.q3 = lds128_if(a3 + r * 16u * REPL, used & 8u);
    } else {
        x.q1 = lds128(a1 + r * 16u * REPL);
        x.q2 = lds128(a2 + r * 16u * REPL);
        x.q3 = lds128(a3 + r * 16u * REPL);
    }
    return x;
}

__device__ __forceinline__ float4 blend_row(const Rows4& x, float w0, float w1, float w2, float w3) {
    float4 m;
    m.x = fmaf(w3, x.q3.x, fmaf(w2, x.q2.x, fmaf(w1, x.q1.x, w0 * x.q0.x)));
    m.y = fmaf(w3, x.q3.y, fmaf(w2, x.q2.y, fmaf(w1, x.q1.y, w0 * x.q0.y)));
    m.z = fmaf(w3, x.q3.z, fmaf(w2, x.q2.z, fmaf(w1, x.q1.z, w0 * x.q0.z)));
    m.w = fmaf(w3, x.q3.w, fmaf(w2, x.q2.w, fmaf(w1, x.q1.w, w0 * x.q0.w)));
    return m;
}

// VPT / PLANAR / MAXT: the shape of the output side.
//   VPT = 8, PLANAR = false (default): [V][9] interleaved vertices, 8 per thread, 256-bit full-sector stores 288 B apart.
//   VPT = 4, PLANAR = true: nine component planes [9][V] per character (dmk.h DMK_OPT_SKIN_PLANAR_OUTPUT); a thread's
//     4 consecutive vertices are 16 contiguous bytes of every plane, so each warp store covers 512 contiguous bytes --
//     whole 128-byte lines instead of 32 scattered sectors -- and half the vertices per thread leave room for 10 (MAXT =
//     320, up to 204 registers) or 14 warps (MAXT = 448, up to 144 registers) per SM instead of 7.
//   VPT = 8, PLANAR = true: the same nine planes from the default thread shape -- a thread's 8 vertices are 32 contiguous bytes of
//     every plane (one st.v8 per plane, held back until the 8th vertex is done): the store layout alone, occupancy as the default.
//   VPT = 8, STREAMS = true: three float3 streams per character, [N][3][V][3] (positions, normals, tangents: the usual
//     non-interleaved vertex-buffer layout); a thread's 8 vertices are 96 contiguous bytes = 3 whole sectors of each stream,
//     flushed as they complete like the interleaved sectors; a warp store then touches 24 lines instead of 32.
template <int REPL, bool SKIP, int VPT = 8, bool PLANAR = false, int MAXT = kMaxThreads, bool STREAMS = false>
__global__ void __launch_bounds__(MAXT, 1) skin_kernel(const SkinParams p) {
    static_assert(VPT == 8 || (PLANAR && VPT == 4), "interleaved / stream output needs 8 vertices per thread (whole sectors); planar takes 4 (st.v4 per plane) or 8 (st.v8)");
    static_assert(!(PLANAR && STREAMS), "one output layout at a time");
    constexpr int kVPT = VPT;   // shadows the file-level constant inside this kernel
    DMK_DYNAMIC_SMEM(smem, 128);
    const int PS = p.palette_size;
    const uint32_t pal_bytes = (uint32_t)PS * 64u;             // staged palette: 3 rows of 4 floats + 1 pad row per slot
    const uint32_t repl_bytes = (uint32_t)PS * 48u * REPL;     // replicated: the 3 rows only
    unsigned char* s_repl = smem;                                        // [kNumRepl][PS*3][REPL] float4
    unsigned char* s_stage = smem + kNumRepl * (size_t)repl_bytes;       // [kNumStage][PS*3] float4 (bulk copy dst)
    uint64_t* s_tma_bar = reinterpret_cast<uint64_t*>(s_stage + kNumStage * (size_t)pal_bytes);  // [kNumStage]
    uint64_t* s_full_bar = s_tma_bar + kNumStage;                        // [kNumRepl]

    const int tid = threadIdx.x, lane = tid & 31;
    const int nwarps = blockDim.x >> 5;
    const int tiles_per_pass = min(p.num_tiles, (int)gridDim.x);
    const int ngroups = gridDim.x / tiles_per_pass;
    const int tslot = blockIdx.x % tiles_per_pass, group = blockIdx.x / tiles_per_pass;
    if (group >= ngroups) return;

    int64_t nchars = p.n;
    if (p.char_ids) nchars = p.char_count[0];
    const int64_t my_count = nchars > group ? (nchars - group + ngroups - 1) / ngroups : 0;
    if (my_count == 0) return;

    if (tid == 0) {
        for (int k = 0; k < kNumStage; ++k) mbar_init(&s_tma_bar[k], 1);
        for (int k = 0; k < kNumRepl; ++k) mbar_init(&s_full_bar[k], nwarps);
        mbar_init_fence();
    }
    __syncthreads();

    // running counts over the whole launch: item q (a character of some tile pass) uses stage q % kNumStage (its
    // (q / kNumStage)-th use) and replica buffer q % kNumRepl (its (q / kNumRepl)-th use)
    uint32_t q_issue = 0;  // thread 0: items whose bulk copy has been issued
    uint32_t q_fill = 0;   // items this thread has replicated
    uint32_t q_comp = 0;   // items this thread has computed

    const uint32_t lane_off = (uint32_t)(lane & (REPL - 1)) * 16u;
    const uint32_t repl_base = smem_u32(s_repl) + lane_off;

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
        uint32_t used[kVPT];
#pragma unroll
        for (int v = 0; v < kVPT; ++v) used[v] = 1u | (w1[v] != 0.f ? 2u : 0u) | (w2[v] != 0.f ? 4u : 0u) | (w3[v] != 0.f ? 8u : 0u);

        auto char_of = [&](int64_t i) -> int64_t {
            const int64_t k = group + i * ngroups;
            return p.char_ids ? (int64_t)__ldg(p.char_ids + k) : k;
        };
        auto issue = [&](int64_t i) {  // thread 0 only: bulk copy of character i's palette into the staging ring
            if (i < my_count) {
                const uint32_t b = q_issue % kNumStage;
                mbar_expect_tx(&s_tma_bar[b], pal_bytes);
                bulk_g2s(s_stage + (size_t)b * pal_bytes, p.palette34 + (size_t)char_of(i) * PS * 16, pal_bytes, &s_tma_bar[b]);
                ++q_issue;
            }
        };
        // staging -> 8 bank-group-private replicas; lane L writes replica (L + k) & 7 at step k: conflict free.
        // Ends with one mbarrier arrival per warp on the replica buffer's "full" barrier.
        auto fill = [&]() {
            const uint32_t sb = q_fill % kNumStage, rb = q_fill % kNumRepl;
            mbar_wait(&s_tma_bar[sb], (q_fill / kNumStage) & 1u);
            const float4* src = reinterpret_cast<const float4*>(s_stage + (size_t)sb * pal_bytes);
            float4* dst = reinterpret_cast<float4*>(s_repl + (size_t)rb * repl_bytes);
            for (int t = tid; t < PS * 3; t += blockDim.x) {
                const float4 val = src[(t / 3) * 4 + t % 3];
#pragma unroll
                for (int k = 0; k < REPL; ++k) dst[t * REPL + ((lane + k) & (REPL - 1))] = val;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_full_bar[rb]);
            ++q_fill;
        };

        // prologue of the tile pass: stage items 0 and 1, replicate item 0.  Every warp has finished its previous
        // pass's last compute only when it gets here, but another warp may still be computing it: the block-wide
        // barrier below (once per tile pass, not per character) closes that window.
        if (tile != tslot) __syncthreads();
        if (tid == 0) { issue(0); issue(1); }
        fill();

        int64_t cid = char_of(0);
        for (int64_t i = 0; i < my_count; ++i) {
            // (1) every warp has replicated item i  => also: every warp is past compute(i - 2)
            mbar_wait(&s_full_bar[q_comp % kNumRepl], (q_comp / kNumRepl) & 1u);
            // (2) so the staging slot of item i is drained and the replica buffer of item i - 2 is free:
            //     stage item i + 2, replicate item i + 1 (its bulk copy was issued one iteration ago)
            if (tid == 0) issue(i + 2);
            if (i + 1 < my_count) fill();
            const int64_t cid_next = (i + 1 < my_count) ? char_of(i + 1) : 0;
            // (3) skin this thread's 8 vertices of character i
            if (active) {
                const uint32_t base = repl_base + (q_comp % kNumRepl) * repl_bytes;
                float* out = p.out + ((size_t)cid * p.vpad + v0) * 9;
                float res[kVPT * 9];
                uint32_t a0, a1, a2, a3;
                auto addr = [&](uint32_t s) {
                    // opaque to the optimiser: otherwise the 32 (vertex, slot) offsets are hoisted out of the
                    // character loop and cost 24 more live registers than the 8 packed slot words
                    asm volatile("" : "+r"(s));
                    a0 = base + (s & 0xffu) * (48u * REPL);
                    a1 = base + ((s >> 8) & 0xffu) * (48u * REPL);
                    a2 = base + ((s >> 16) & 0xffu) * (48u * REPL);
                    a3 = base + (s >> 24) * (48u * REPL);
                };
                addr(slots[0]);
                Rows4 cur = load_row<REPL, SKIP>(a0, a1, a2, a3, 0, used[0]);
#pragma unroll
                for (int v = 0; v < kVPT; ++v) {
                    float4 m[3];
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        Rows4 nxt;  // next row (or row 0 of the next vertex) is requested before this one is consumed
                        if (r < 2) {
                            nxt = load_row<REPL, SKIP>(a0, a1, a2, a3, r + 1, used[v]);
                        } else if (v + 1 < kVPT) {
                            addr(slots[v + 1]);
                            nxt = load_row<REPL, SKIP>(a0, a1, a2, a3, 0, used[v + 1]);
                        }
                        m[r] = blend_row(cur, w0[v], w1[v], w2[v], w3[v]);
                        cur = nxt;
                    }
                    const float m0[3] = {m[0].x, m[1].x, m[2].x}, m1[3] = {m[0].y, m[1].y, m[2].y};
                    const float m2[3] = {m[0].z, m[1].z, m[2].z}, m3[3] = {m[0].w, m[1].w, m[2].w};
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        // interleaved / planar: vertex-major (v * 9 + attribute * 3 + r); streams: attribute-major (a * 24 + v * 3 + r)
                        res[STREAMS ? v * 3 + r : v * 9 + r] = fmaf(m2[r], pz[v], fmaf(m1[r], py[v], fmaf(m0[r], px[v], m3[r])));
                        res[STREAMS ? kVPT * 3 + v * 3 + r : v * 9 + 3 + r] = fmaf(m2[r], nz[v], fmaf(m1[r], ny[v], m0[r] * nx[v]));
                        res[STREAMS ? kVPT * 6 + v * 3 + r : v * 9 + 6 + r] = fmaf(m2[r], tz[v], fmaf(m1[r], ty[v], m0[r] * tx[v]));
                    }
                    // flush every 32-byte sector that just became complete
                    if (!PLANAR && !STREAMS) {
#pragma unroll
                        for (int sct = (v * 9) / 8; sct < ((v + 1) * 9) / 8; ++sct) st_v8(out + sct * 8, res + sct * 8);
                    }
                    if (STREAMS) {
#pragma unroll
                        for (int a = 0; a < 3; ++a) {
                            float* stream = p.out + (((size_t)cid * 3 + a) * p.vpad + v0) * 3;
#pragma unroll
                            for (int sct = (v * 3) / 8; sct < ((v + 1) * 3) / 8; ++sct) st_v8(stream + sct * 8, res + a * kVPT * 3 + sct * 8);
                        }
                    }
                }
                if (PLANAR) {
                    // component planes: plane c holds component c (position.xyz, normal.xyz, tangent.xyz) of every vertex of
                    // the character; this thread's 4 vertices are 16 contiguous bytes of each
                    float* plane = p.out + (size_t)cid * 9 * p.vpad + v0;
#pragma unroll
                    for (int c = 0; c < 9; ++c) {
                        if (VPT == 4) {
                            st_v4(plane + (size_t)c * p.vpad, res[c], res[9 + c], res[18 + c], res[27 + c]);
                        } else {   // 8 vertices: 32 contiguous bytes of the plane, one whole sector per lane
                            const float v8[8] = {res[c], res[9 + c], res[18 + c], res[27 + c], res[36 + c], res[45 + c], res[54 + c], res[63 + c]};
                            st_v8(plane + (size_t)c * p.vpad, v8);
                        }
                    }
                }
            }
            ++q_comp;
            cid = cid_next;
        }
    }
}


// ------------------------------------------------------------------------------------------------------------------------
// Warp-specialised shape of the same kernel (round 2).  What the captures said about the kernel above
// (profiles/r02_skin_kernel_layout_default.txt): the L1 data pipe is 91 % busy -- 828 M gather wavefronts, 64 M replica
// stores, 235 M global-store wavefronts (stores cost 64 B per cycle whatever their layout) -- and 18 % of the stall samples
// are warps parked on an mbarrier: every warp replicates a slice of the next palette BEFORE it skins the current one, so a
// bulk copy that arrives late (issued only one character ahead, behind 4 TB/s of writes) stops all 7 warps.  Here
//   * the palette path runs without the LSU: ONE thread (of a producer warp) issues, per character, 8 bulk copies of the
//     palette from global memory into a ring of up to 4 replica buffers, several characters ahead; the copies' bytes
//     complete the buffer's "full" mbarrier; the compute warps return the buffer through its "empty" mbarrier;
//   * the COMPUTE warps (up to 7) do nothing but wait for a full buffer, skin their 8 vertices and hand the buffer back.
//     No warp ever waits for another compute warp;
//   * SHARE (meshes stored with DMK_MESH_CLUSTER_BONES): entry 0 of the thread's first vertex names the block's bone; its 3
//     rows are read once per character into registers and every vertex whose entry 0 is that bone skips its gather.  With
//     SKIP (absent influences are predicated off, whole quarter-warps at a time) the gathers drop from 96 to ~51 LDS.128
//     per thread and character (dmk_mesh_cluster_stats), which takes the kernel off the L1 data pipe.
// LAYOUT: 0 interleaved [V][9], 3 three float3 streams [3][V][3], 4 nine planes [9][V]  (DMK_OPT_SKIN_PLANAR_OUTPUT values).
#ifndef DMK_SKIN_OUT_BUFS
#define DMK_SKIN_OUT_BUFS 8
#endif
constexpr int kOutBufs = DMK_SKIN_OUT_BUFS;   // staging buffers per compute warp of the bulk-store output path (a divisor of 8).  With 2 the warp
// waits for the copy of two rounds ago to have READ its buffer, which takes longer than two rounds (5.95 ms per launch, measured)
constexpr int kMaxRing = 4;   // replica ring of the warp-specialised kernel (SkinParams::num_stages: 4, fewer where shared memory is short).
// MEASURED (profiles/r02_skin_ring_sleep_experiment.jsonl): rings of 5 and 6 buffers, and a producer that sleeps between the polls of
// the "empty" barrier (its spin is 13 % of the warp instructions the kernel executes, one lane each) change neither the burst
// nor the power-capped time -- the palettes are not what the compute warps wait for.

// Replica layout of the warp-specialised kernel: REPL plain copies of the character's palette34 ([PS][16] floats, rows 0..2 of
// every slot used), each written by ONE bulk copy straight from global memory -- no thread ever touches the palette on its
// way in.  Replica r starts at r * RS with RS = 64 PS + 16 pad and pad chosen so that RS / 16 = 1 (mod 8): row `row` of slot
// s in replica r then sits in bank group (r + 4 s + row) mod 8, and lane L reads replica (L + 4 (s & 1)) mod 8, i.e. bank
// group (L + row) mod 8 whatever the slot -- the 8 lanes of a quarter-warp never collide, as with the interleaved replicas
// of the round-1 kernel, for 2 more integer instructions per gather.
__host__ __device__ inline uint32_t replica_stride(int palette_size) { return (uint32_t)palette_size * 64u + ((palette_size & 1) ? 80u : 16u); }

//
// VPT = 4 (round 2, DMK_OPT_SKIN_VERTICES_PER_THREAD): an 8-vertex block -- still the unit of the output's whole sectors and
// of the clustered mesh order -- is shared by TWO lanes of a warp, lane L (first 4 vertices) and lane L ^ 8 (last 4), so the 8
// lanes of a quarter-warp still hold the same position of 8 consecutive blocks (what the predicated gathers need to cost no
// wavefront).  4 x 36 B = 144 B is 4.5 sectors: the second lane passes its first 16 bytes to the first one by 4 shuffles, which
// then writes 5 whole sectors, the second lane 4.  Half the registers per thread (13 x 4 bind-pose values instead of 13 x 8),
// so 10 compute warps per SM instead of 7 (MAXT = 352).  Measured slower than 8 vertices per thread (see kMaxThreads4).
//
// STRIDED (round 2, the bulk-store output path; meshes stored by cluster_mesh_strided): thread (warp w, lane l) owns the stored vertices
// tile_begin + 256 w + k nl + l, k = 0..7 (nl = lanes of the warp that have vertices), so round k of a warp is nl consecutive
// vertices = nl x 36 contiguous bytes of the output.  The lanes write them to a per-warp staging buffer in shared memory (9 x STS.32,
// stride 9 words: conflict free) and ONE cp.async.bulk shared -> global per round sends them (two buffers per warp: the copy of round
// k - 2 must have finished READING its buffer before round k overwrites it).  No st.global, no register staging of whole sectors, and
// the output leaves as full 128-byte lines instead of 32 scattered sectors per store instruction.  tools/ubench_out_path.cu: with the
// clustered mesh's ~6 gathers per vertex the memory side alone runs 3.21 ms with st.global.v8 and 2.99 ms this way (with the 12
// gathers of the caller's order it is the other way round: the bulk copies' reads compete with the gathers, 4.17 against 3.40 ms).
// In the kernel it LOSES (5.95 against 3.55 ms): see strided_enabled() below.  Opt-in, DMK_SKIN_BULK_STORE=1.
template <int REPL, bool SKIP, bool SHARE, int LAYOUT, int VPT = 8, int MAXT = kMaxThreads, bool STRIDED = false>
__global__ void __launch_bounds__(MAXT, 1) skin_ws_kernel(const SkinParams p) {
    static_assert(LAYOUT == 0 || LAYOUT == 3 || LAYOUT == 4, "interleaved, three streams or nine planes");
    static_assert(VPT == 8 || (VPT == 4 && LAYOUT == 0), "4 vertices per thread: interleaved output only");
    static_assert(!STRIDED || (VPT == 8 && LAYOUT == 0), "bulk-store output path: interleaved output, 8 vertices per thread");
    constexpr int kVPT = VPT;
    constexpr bool STREAMS = LAYOUT == 3, PLANAR = LAYOUT == 4;
    DMK_DYNAMIC_SMEM(smem, 128);
    const int PS = p.palette_size;
    const uint32_t pal_bytes = (uint32_t)PS * 64u;             // one palette34: 3 rows of 4 floats + 1 pad row per slot
    const uint32_t RS = replica_stride(PS);
    const uint32_t buf_bytes = RS * REPL;                      // one ring buffer: REPL replicas
    const int ring = p.num_stages;
    unsigned char* s_repl = smem;                                        // [ring][REPL] palettes, RS apart
    uint64_t* s_full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)ring * buf_bytes);   // [kMaxRing] bulk copies -> compute warps
    uint64_t* s_empty_bar = s_full_bar + kMaxRing;                       // [kMaxRing] compute warps -> producer
    uint64_t* s_stage_bar = s_empty_bar + kMaxRing;                      // [kMaxRing] shared-memory replication: replica 0 has arrived
    volatile int32_t* s_cid = reinterpret_cast<volatile int32_t*>(s_stage_bar + kMaxRing);   // [kMaxRing] character of the item in a ring buffer, -1 = end of the tile pass
    // (STRIDED) [compute warps][kOutBufs][32 x 36 B], on a 128-byte boundary of the block's shared memory
    unsigned char* s_out = smem + (((size_t)ring * buf_bytes + 8 * 3 * kMaxRing + 4 * kMaxRing + 127) & ~(size_t)127);
    const bool s2s = p.replicate_in_smem != 0;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ncompute = (blockDim.x >> 5) - 1;                          // the last warp is the producer
    // Which tiles and characters a block works on.  Dynamic (default, p.next_item): every block walks ALL tiles, starting at its own
    // (blockIdx.x mod num_tiles), and takes the next character of the tile from a counter until none is left -- so blocks that are
    // done with their tile help with the others, no SM has to stay idle to make the grid a multiple of the tile count, and tiles of
    // different weight even out.  Static (p.next_item == nullptr, A/B): block (tile slot, group) takes characters group, group +
    // ngroups, ... of the tiles tslot, tslot + tiles_per_pass, ...
    const bool dynamic = p.next_item != nullptr;
    const int tiles_per_pass = min(p.num_tiles, (int)gridDim.x);
    const int ngroups = gridDim.x / tiles_per_pass;
    const int tslot = blockIdx.x % tiles_per_pass, group = blockIdx.x / tiles_per_pass;
    if (!(dynamic && p.num_tiles <= (int)gridDim.x) && group >= ngroups) return;
    // (a mesh with more tiles than blocks -- or the static split -- : tiles tslot, tslot + tiles_per_pass, ...; helping is limited to the
    // next 7 tiles so that a huge mesh with few characters does not pay one empty hand-shake per tile and block)
    const bool walk = dynamic && p.num_tiles <= (int)gridDim.x;
    const int tile_iters = walk ? min(p.num_tiles, 8) : (p.num_tiles - tslot + tiles_per_pass - 1) / tiles_per_pass;
    auto tile_at = [&](int j) { return walk ? (int)((blockIdx.x + (unsigned)j) % (unsigned)p.num_tiles) : tslot + j * tiles_per_pass; };

    int64_t nchars = p.n;
    if (p.char_ids) nchars = p.char_count[0];
    const int64_t my_count = nchars > group ? (nchars - group + ngroups - 1) / ngroups : 0;   // (static split)

    if (tid == 0) {
        for (int k = 0; k < ring; ++k) {
            mbar_init(&s_full_bar[k], 1);
            mbar_init(&s_empty_bar[k], ncompute);
            mbar_init(&s_stage_bar[k], 1);
        }
        mbar_init_fence();
    }
    __syncthreads();

    if (warp == ncompute) {
        // ---------------- producer: one thread.  Per (tile pass, character): wait until the ring buffer is free, then fill its REPL
        // replicas with bulk copies; their bytes complete the buffer's "full" barrier, nothing else is needed.  Two ways:
        //   * REPL copies from global memory (L2): simplest, but 8 x 4352 B per character and block is 10 GB of L2 reads per
        //     100k x 5k launch next to 18 GB of writes (profiles/r02_skin_ws_clustered_v4.txt: lts throughput 74 %);
        //   * (replicate_in_smem) ONE copy from global memory into replica 0, then REPL - 1 shared-memory -> shared-memory bulk
        //     copies (cp.async.bulk.shared::cluster.shared::cta) one item behind, so the wait for replica 0 is never exposed.
        //     MEASURED: slower (4.16 ms against 3.52 ms; the copies run at 15 B/cycle/SM, tools/ubench_s2s.cu, and take
        //     shared-memory bandwidth from the gathers), so the first way is the default and this one a diagnostic option.
        //
        // WHICH character: by default the blocks of a tile take the next one nobody has taken from a counter in global memory
        // (p.next_item[tile], zeroed by launch_skin).  MEASURED (profiles/r02_skin_cta_finish_times_*.txt): with the static
        // split (every block the same number of characters) most blocks of the 100k x 5k launch are done after 3.17-3.25 ms but
        // eight of them -- the same SMs in every launch -- need 3.6 ms, and the kernel lasts as long as they do.
        // The character of a ring buffer travels in s_cid (written before the arrival on the buffer's barrier); -1 ends a tile pass.
        if (lane != 0) return;
        uint32_t rb = 0, rphase = 0, sb = 0, sphase = 0;
        bool wrapped = false;
        auto advance_rb = [&] { if (++rb == (uint32_t)ring) { rb = 0; rphase ^= 1u; wrapped = true; } };
        auto advance_sb = [&] { if (++sb == (uint32_t)ring) { sb = 0; sphase ^= 1u; } };
        for (int j = 0; j < tile_iters; ++j) {
            const int tile = tile_at(j);
            int64_t issued = 0, staged = 0, i_static = 0;   // items of this tile pass: requested / (s2s) replicated inside shared memory
            bool exhausted = false;
            for (;;) {
                if (!exhausted) {
                    int64_t k;
                    if (p.next_item) k = atomicAdd(p.next_item + tile, 1);
                    else k = i_static < my_count ? group + (i_static++) * ngroups : nchars;
                    if (k >= nchars) {
                        exhausted = true;
                    } else {
                        if (wrapped) mbar_wait(&s_empty_bar[rb], rphase ^ 1u);   // every compute warp is done with the item `ring` items back
                        const int32_t cid = p.char_ids ? __ldg(p.char_ids + k) : (int32_t)k;
                        s_cid[rb] = cid;
                        const float* src = p.palette34 + (size_t)cid * PS * 16;
                        unsigned char* dst = s_repl + (size_t)rb * buf_bytes;
                        if (s2s) {
                            mbar_expect_tx(&s_stage_bar[rb], pal_bytes);
                            bulk_g2s(dst, src, pal_bytes, &s_stage_bar[rb]);
                        } else {
                            mbar_expect_tx(&s_full_bar[rb], pal_bytes * REPL);
#pragma unroll
                            for (int r = 0; r < REPL; ++r) bulk_g2s(dst + (size_t)r * RS, src, pal_bytes, &s_full_bar[rb]);
                        }
                        ++issued;
                        advance_rb();
                    }
                }
                // (s2s) the item issued one iteration ago -- its replica 0 has had a whole item's time to arrive -- or, at the end, the last one
                if (s2s && staged < issued && (exhausted || staged + 1 < issued)) {
                    mbar_wait(&s_stage_bar[sb], sphase);
                    unsigned char* buf = s_repl + (size_t)sb * buf_bytes;
                    mbar_expect_tx(&s_full_bar[sb], pal_bytes * (REPL - 1));
#pragma unroll
                    for (int r = 1; r < REPL; ++r) bulk_s2s(buf + (size_t)r * RS, buf, pal_bytes, &s_full_bar[sb]);
                    ++staged;
                    advance_sb();
                }
                if (exhausted && (!s2s || staged == issued)) break;
            }
            // the end of this tile pass takes a ring slot of its own
            if (wrapped) mbar_wait(&s_empty_bar[rb], rphase ^ 1u);
            s_cid[rb] = -1;
            if (s2s) {
                mbar_arrive(&s_stage_bar[rb]);
                advance_sb();
            }
            mbar_arrive(&s_full_bar[rb]);
            advance_rb();
        }
        return;
    }

    // ---------------- compute warps ----------------
    // this lane's replica for even slots, and how far away its replica for odd slots is (REPL < 8: the same one)
    const uint32_t lane_repl = (uint32_t)lane & (REPL - 1);
    const uint32_t repl_base = smem_u32(s_repl) + lane_repl * RS;
    const uint32_t odd_delta = (((lane_repl ^ 4u) & (REPL - 1)) - lane_repl) * RS;   // (mod 2^32)
    uint32_t rb = 0, rphase = 0;
    for (int j = 0; j < tile_iters; ++j) {
        const int tile = tile_at(j);
        // the first item of this tile pass -- or already its end, when the tile's characters were all taken before this block got here:
        // then there is nothing to load
        if (s2s) mbar_wait(&s_stage_bar[rb], rphase);
        mbar_wait(&s_full_bar[rb], rphase);
        if (s_cid[rb] < 0) {
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty_bar[rb]);
            if (++rb == (uint32_t)ring) { rb = 0; rphase ^= 1u; }
            continue;
        }
        // ---- this thread's 8 (4) vertices: loaded once, kept in registers for every character ----
        // the 8-vertex block of the tile this lane works on, and which half of it (VPT = 4: lanes L and L ^ 8 share a block)
        const int blk = VPT == 8 ? tid : warp * 16 + ((lane >> 4) << 3) + (lane & 7);
        const int half = VPT == 8 ? 0 : (lane >> 3) & 1;
        // STRIDED: nl = lanes of this warp that have vertices (a multiple of 8: tiles and the vertex pitch are multiples of 64)
        const int tile_verts = min(p.tile_vertices, p.vpad - tile * p.tile_vertices);
        const int nl = STRIDED ? max(0, min(32, tile_verts / 8 - 32 * warp)) : 0;
        const int v0 = STRIDED ? tile * p.tile_vertices + warp * 256 + lane : tile * p.tile_vertices + blk * 8 + half * 4;
        const int vstep = STRIDED ? nl : 1;
        const bool active = STRIDED ? lane < nl : (blk * 8 < p.tile_vertices && v0 < p.vpad);
        float px[kVPT], py[kVPT], pz[kVPT], nx[kVPT], ny[kVPT], nz[kVPT], tx[kVPT], ty[kVPT], tz[kVPT];
        float w0[kVPT], w1[kVPT], w2[kVPT], w3[kVPT];
        uint32_t slots[kVPT];
#pragma unroll
        for (int v = 0; v < kVPT; ++v) {
            const int vi = active ? v0 + v * vstep : 0;
            px[v] = __ldg(p.pos + (size_t)vi * 3 + 0); py[v] = __ldg(p.pos + (size_t)vi * 3 + 1); pz[v] = __ldg(p.pos + (size_t)vi * 3 + 2);
            nx[v] = __ldg(p.nrm + (size_t)vi * 3 + 0); ny[v] = __ldg(p.nrm + (size_t)vi * 3 + 1); nz[v] = __ldg(p.nrm + (size_t)vi * 3 + 2);
            tx[v] = __ldg(p.tan + (size_t)vi * 3 + 0); ty[v] = __ldg(p.tan + (size_t)vi * 3 + 1); tz[v] = __ldg(p.tan + (size_t)vi * 3 + 2);
            const float4 w = __ldg(reinterpret_cast<const float4*>(p.weights) + vi);
            w0[v] = w.x; w1[v] = w.y; w2[v] = w.z; w3[v] = w.w;
            slots[v] = __ldg(p.slots + vi);
        }
        // bit 4 v + k: influence k of vertex v is fetched from shared memory (k = 0: unless it is the block's bone)
        uint32_t used = 0;
        // the block's bone: entry 0 of the block's FIRST vertex (the other half's vertex 0 for the second lane of a pair)
        const uint32_t block_slot = (VPT == 4 && half && active ? __ldg(p.slots + v0 - 4) : slots[0]) & 0xffu;
        // VPT = 4: the whole warp runs the item when any lane has vertices (the pair's shuffles need both lanes), stores are predicated
        const bool warp_active = (VPT == 8 && !STRIDED) ? active : __ballot_sync(0xffffffffu, active) != 0u;
        const uint32_t stage_base = STRIDED ? smem_u32(s_out) + (uint32_t)warp * (kOutBufs * 1152u) + (uint32_t)lane * 36u : 0u;
        const uint32_t st_first = active && half == 0, st_second = active && half == 1;
#pragma unroll
        for (int v = 0; v < kVPT; ++v) {
            uint32_t u = (SHARE && (slots[v] & 0xffu) == block_slot) ? 0u : 1u;
            if (SKIP) u |= (w1[v] != 0.f ? 2u : 0u) | (w2[v] != 0.f ? 4u : 0u) | (w3[v] != 0.f ? 8u : 0u);
            else u |= 14u;
            used |= u << (4 * v);
        }
        const uint32_t block_off = block_slot * 64u + ((block_slot & 1u) ? odd_delta : 0u);
        // bit v: some lane of this warp still gathers entry 0 of its vertex v (mixed blocks only)
        uint32_t rare = 0;
        if (SHARE) {
#pragma unroll
            for (int v = 0; v < kVPT; ++v) rare |= __ballot_sync(0xffffffffu, (used >> (4 * v) & 1u) != 0u) != 0u ? 1u << v : 0u;
        }

        for (;;) {
            if (s2s) mbar_wait(&s_stage_bar[rb], rphase);   // replica 0 (the async copy that the other replicas were made from)
            mbar_wait(&s_full_bar[rb], rphase);
            const int64_t cid = s_cid[rb];                  // written by the producer before its arrival on the barrier; -1: the pass is over
            // One character for this thread's 8 vertices.  RARE: some lane of the warp still gathers entry 0 of some vertex (a mixed
            // block of a clustered mesh).  The common instantiation has no branch in it: one basic block of ~1000 instructions, so the
            // scheduler can put the gathers of vertex v + 1 under the arithmetic of vertex v.
            auto item = [&](auto rare_tag) {
                constexpr bool RARE = decltype(rare_tag)::value;
                const uint32_t base = repl_base + rb * buf_bytes;
                float4 b0, b1, b2;   // the block's bone, once per character
                if (SHARE) {
                    b0 = lds128(base + block_off);
                    b1 = lds128(base + block_off + 16u);
                    b2 = lds128(base + block_off + 32u);
                }
                float res[kVPT * 9];
                float recv[4];   // (VPT = 4) the first 16 bytes of the block's second half
                uint32_t a0, a1, a2, a3;
                auto slot_addr = [&](uint32_t slot) { return base + slot * 64u + ((slot & 1u) ? odd_delta : 0u); };
                auto addr = [&](uint32_t s) {
                    asm volatile("" : "+r"(s));   // keeps the 32 (vertex, slot) offsets from being hoisted into 24 more registers
                    a0 = slot_addr(s & 0xffu);
                    a1 = slot_addr((s >> 8) & 0xffu);
                    a2 = slot_addr((s >> 16) & 0xffu);
                    a3 = slot_addr(s >> 24);
                };
                auto load = [&](int v, int r) {
                    const uint32_t u = used >> (4 * v);
                    Rows4 x;
                    x.q0 = lds128(a0 + r * 16u);
                    if (SKIP) {
                        x.q1 = lds128_if(a1 + r * 16u, u & 2u);
                        x.q2 = lds128_if(a2 + r * 16u, u & 4u);
                        x.q3 = lds128_if(a3 + r * 16u, u & 8u);
                    } else {
                        x.q1 = lds128(a1 + r * 16u);
                        x.q2 = lds128(a2 + r * 16u);
                        x.q3 = lds128(a3 + r * 16u);
                    }
                    return x;
                };
                // SHARE: rows of influences 1..3; a skipped gather keeps what is there and meets a zero weight, so they start at
                // zero and only ever hold rows of this character's palette
                float4 T[3][3];
                if (SHARE) {
#pragma unroll
                    for (int k = 0; k < 3; ++k)
#pragma unroll
                        for (int r = 0; r < 3; ++r) T[k][r] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                Rows4 cur;
                if (!SHARE) {
                    addr(slots[0]);
                    cur = load(0, 0);
                }
#pragma unroll
                for (int v = 0; v < kVPT; ++v) {
                    float4 m[3];
                    if (SHARE) {
                        // entry 0 is the block's bone (already in registers) for every lane of the warp, except in the few
                        // mixed blocks: a warp-uniform branch, so the common path carries neither the gather nor a select
                        addr(slots[v]);
                        float4 e0 = b0, e1 = b1, e2 = b2;
                        if (RARE && (rare >> v & 1u)) {
                            const uint32_t take = used >> (4 * v) & 1u;
                            e0 = lds128_or(a0, take, b0);
                            e1 = lds128_or(a0 + 16u, take, b1);
                            e2 = lds128_or(a0 + 32u, take, b2);
                        }
                        m[0] = make_float4(w0[v] * e0.x, w0[v] * e0.y, w0[v] * e0.z, w0[v] * e0.w);
                        m[1] = make_float4(w0[v] * e1.x, w0[v] * e1.y, w0[v] * e1.z, w0[v] * e1.w);
                        m[2] = make_float4(w0[v] * e2.x, w0[v] * e2.y, w0[v] * e2.z, w0[v] * e2.w);
#pragma unroll
                        for (int r = 0; r < 3; ++r) {
                            lds128_keep(T[0][r], a1 + r * 16u, w1[v] != 0.f);
                            lds128_keep(T[1][r], a2 + r * 16u, w2[v] != 0.f);
                            lds128_keep(T[2][r], a3 + r * 16u, w3[v] != 0.f);
                        }
#pragma unroll
                        for (int r = 0; r < 3; ++r) {
                            m[r].x = fmaf(w3[v], T[2][r].x, fmaf(w2[v], T[1][r].x, fmaf(w1[v], T[0][r].x, m[r].x)));
                            m[r].y = fmaf(w3[v], T[2][r].y, fmaf(w2[v], T[1][r].y, fmaf(w1[v], T[0][r].y, m[r].y)));
                            m[r].z = fmaf(w3[v], T[2][r].z, fmaf(w2[v], T[1][r].z, fmaf(w1[v], T[0][r].z, m[r].z)));
                            m[r].w = fmaf(w3[v], T[2][r].w, fmaf(w2[v], T[1][r].w, fmaf(w1[v], T[0][r].w, m[r].w)));
                        }
                    } else {
#pragma unroll
                        for (int r = 0; r < 3; ++r) {
                            Rows4 nxt;  // next row (or row 0 of the next vertex) is requested before this one is consumed
                            if (r < 2) {
                                nxt = load(v, r + 1);
                            } else if (v + 1 < kVPT) {
                                addr(slots[v + 1]);
                                nxt = load(v + 1, 0);
                            }
                            m[r] = blend_row(cur, w0[v], w1[v], w2[v], w3[v]);
                            cur = nxt;
                        }
                    }
                    const float m0[3] = {m[0].x, m[1].x, m[2].x}, m1[3] = {m[0].y, m[1].y, m[2].y};
                    const float m2[3] = {m[0].z, m[1].z, m[2].z}, m3[3] = {m[0].w, m[1].w, m[2].w};
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        // interleaved / planar: vertex-major (v * 9 + attribute * 3 + r); streams: attribute-major (a * 24 + v * 3 + r)
                        res[STREAMS ? v * 3 + r : v * 9 + r] = fmaf(m2[r], pz[v], fmaf(m1[r], py[v], fmaf(m0[r], px[v], m3[r])));
                        res[STREAMS ? kVPT * 3 + v * 3 + r : v * 9 + 3 + r] = fmaf(m2[r], nz[v], fmaf(m1[r], ny[v], m0[r] * nx[v]));
                        res[STREAMS ? kVPT * 6 + v * 3 + r : v * 9 + 6 + r] = fmaf(m2[r], tz[v], fmaf(m1[r], ty[v], m0[r] * tx[v]));
                    }
                    if (STRIDED) {
                        // round v of this warp: nl consecutive vertices.  Buffer v % kOutBufs was last read by the bulk copy issued kOutBufs
                        // rounds ago (with a buffer per round: by the same round of the previous character)
                        const uint32_t buf = stage_base + (uint32_t)(v % kOutBufs) * 1152u;
                        if (lane == 0) bulk_wait_read<kOutBufs - 1>();
                        __syncwarp();
                        if (active) {
#pragma unroll
                            for (int c = 0; c < 9; ++c) sts32(buf + 4u * c, res[v * 9 + c]);
                        }
                        fence_proxy_async_smem();   // the lanes' writes, before the async proxy reads them
                        __syncwarp();
                        if (lane == 0) {
                            bulk_s2g(p.out + ((size_t)cid * p.vpad + (size_t)(tile * p.tile_vertices + warp * 256 + v * nl)) * 9, buf, (uint32_t)nl * 36u);
                            bulk_commit();
                        }
                    }
                    if (LAYOUT == 0 && VPT == 8 && !STRIDED) {   // flush every 32-byte sector that just became complete
                        float* out = p.out + ((size_t)cid * p.vpad + v0) * 9;
#pragma unroll
                        for (int sct = (v * 9) / 8; sct < ((v + 1) * 9) / 8; ++sct) st_v8(out + sct * 8, res + sct * 8);
                    }
                    if (LAYOUT == 0 && VPT == 4) {
                        // this lane's 36 floats start on a sector boundary (first half of the block) or 16 bytes behind one
                        // (second half).  First half: sectors [8 s, 8 s + 8), s = 0..3, and a fifth made of its last 4 floats and
                        // the second half's first 4 (shuffled over as soon as they exist); second half: sectors [4 + 8 s, 12 + 8 s).
                        float* out = p.out + ((size_t)cid * p.vpad + v0) * 9;
                        if (v == 0) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) recv[k] = __shfl_xor_sync(0xffffffffu, res[k], 8);
                        }
#pragma unroll
                        for (int sct = 0; sct < 4; ++sct) {
                            if (8 * (sct + 1) > 9 * v && 8 * (sct + 1) <= 9 * (v + 1)) st_v8_if(out + 8 * sct, res + 8 * sct, st_first);
                            if (4 + 8 * (sct + 1) > 9 * v && 4 + 8 * (sct + 1) <= 9 * (v + 1)) st_v8_if(out + 4 + 8 * sct, res + 4 + 8 * sct, st_second);
                        }
                        if (v == 3) {
                            const float last[8] = {res[32], res[33], res[34], res[35], recv[0], recv[1], recv[2], recv[3]};
                            st_v8_if(out + 32, last, st_first);
                        }
                    }
                    if (STREAMS) {
#pragma unroll
                        for (int a = 0; a < 3; ++a) {
                            float* stream = p.out + (((size_t)cid * 3 + a) * p.vpad + v0) * 3;
#pragma unroll
                            for (int sct = (v * 3) / 8; sct < ((v + 1) * 3) / 8; ++sct) st_v8(stream + sct * 8, res + a * kVPT * 3 + sct * 8);
                        }
                    }
                }
                if (PLANAR) {   // one whole sector of every component plane
                    float* plane = p.out + (size_t)cid * 9 * p.vpad + v0;
#pragma unroll
                    for (int c = 0; c < 9; ++c) {
                        const float v8[8] = {res[c], res[9 + c], res[18 + c], res[27 + c], res[36 + c], res[45 + c], res[54 + c], res[63 + c]};
                        st_v8(plane + (size_t)c * p.vpad, v8);
                    }
                }
                        };
            if (warp_active && cid >= 0) {
                if (SHARE && rare != 0u) item(std::true_type{});
                else item(std::false_type{});
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty_bar[rb]);
            if (++rb == (uint32_t)ring) { rb = 0; rphase ^= 1u; }
            if (cid < 0) break;
        }
    }
    if (STRIDED && lane == 0) bulk_wait_all();   // the staging buffers must outlive the copies that read them
}

}  // namespace

namespace {
constexpr size_t kMaxSmem = 227 * 1024;   // per-CTA shared memory limit on sm_100
constexpr int kComputeThreads = kMaxThreads - 32;   // warp-specialised shape: the 8th warp is the producer
// 4 vertices per thread: 10 compute warps + the producer (up to 184 registers; ptxas takes 153-168).  MEASURED (r02 call 11,
// profiles/r02_skin_vpt4_experiment.jsonl): 3.73-3.79 ms against 3.54-3.59 ms for 8 vertices per thread on the clustered mesh,
// 4.04 against 3.89 on the caller's order; 14 + 1 warps at 128 registers (spills) 5.10 ms.  More resident warps do not help:
// the kernel is not short of warps to issue from, it is short of shared-memory and store bandwidth per SM.  Kept selectable
// (DMK_SKIN_VPT=4 in the environment) as a diagnostic.
constexpr int kMaxThreads4 = 352;
// legacy: the round-1 kernel (every warp replicates and skins); the VPT = 4 layouts (1, 2) only exist in that shape
bool is_legacy(int variant, bool legacy) { return legacy || variant == 1 || variant == 2; }
constexpr size_t kOutStageBytes = (size_t)(kComputeThreads / 32) * kOutBufs * 1152;   // bulk-store output path: kOutBufs rounds of 32 vertices per compute warp
size_t skin_smem_bytes(int palette_size, int repl, bool legacy, int ring = kMaxRing, bool strided = false) {
    if (legacy) return (size_t)palette_size * (48 * kNumRepl * repl + 64 * kNumStage) + 8 * (kNumStage + kNumRepl) + 16;
    return (size_t)replica_stride(palette_size) * repl * ring + 8 * 3 * kMaxRing + 4 * kMaxRing + 16 + (strided ? kOutStageBytes + 128 : 0);
}
template <typename F>
auto with_legacy_kernel(int repl, bool skip, int variant, F&& f) {
    if (variant == 4) return skip ? f(skin_kernel<8, true, 8, true>) : f(skin_kernel<8, false, 8, true>);   // planar output, the default kernel's thread shape
    if (variant == 3) return skip ? f(skin_kernel<8, true, 8, false, kMaxThreads, true>) : f(skin_kernel<8, false, 8, false, kMaxThreads, true>);   // three float3 streams
    if (variant == 1) return skip ? f(skin_kernel<8, true, 4, true, 320>) : f(skin_kernel<8, false, 4, true, 320>);   // planar output, up to 10 warps
    if (variant == 2) return skip ? f(skin_kernel<8, true, 4, true, 448>) : f(skin_kernel<8, false, 4, true, 448>);   // planar output, up to 14 warps
    if (skip && repl == 8) return f(skin_kernel<8, true>);   // meshes grouped by influence count
    switch (repl) {
    case 8: return f(skin_kernel<8, false>);
    case 4: return f(skin_kernel<4, false>);
    case 2: return f(skin_kernel<2, false>);
    default: return f(skin_kernel<1, false>);
    }
}
// mode: 0 caller's vertex order, 1 influence-sorted (predicated gathers), 2 clustered by bone (block bone in registers + predicated gathers)
template <int LAYOUT, typename F>
auto with_ws_layout(int repl, int mode, int vpt, bool strided, F&& f) {
    if constexpr (LAYOUT == 0) {
        if (repl == 8 && vpt == 8 && strided && mode == 3) return f(skin_ws_kernel<8, true, true, 0, 8, kMaxThreads, true>);
        if (repl == 8 && vpt == 4) {
            if (mode == 2) return f(skin_ws_kernel<8, true, true, 0, 4, kMaxThreads4>);
            if (mode == 1) return f(skin_ws_kernel<8, true, false, 0, 4, kMaxThreads4>);
            return f(skin_ws_kernel<8, false, false, 0, 4, kMaxThreads4>);
        }
    }
    if (mode == 3) mode = 2;   // a mesh stored for the bulk-store path under another kernel: clustered semantics hold on any order
    if (repl == 8) {
        if (mode == 2) return f(skin_ws_kernel<8, true, true, LAYOUT>);
        if (mode == 1) return f(skin_ws_kernel<8, true, false, LAYOUT>);
        return f(skin_ws_kernel<8, false, false, LAYOUT>);
    }
    if (LAYOUT != 0) return f(skin_ws_kernel<8, false, false, LAYOUT>);   // (refused earlier: these layouts need 8 replicas)
    // palettes too large for 8 replicas: fewer, with bank conflicts; sorted meshes take the clustered kernel's predicates
    // (SHARE only ever compares entry 0 with the block's first vertex, so it is correct on any order)
    switch (repl) {
    case 4: return mode ? f(skin_ws_kernel<4, true, true, 0>) : f(skin_ws_kernel<4, false, false, 0>);
    case 2: return mode ? f(skin_ws_kernel<2, true, true, 0>) : f(skin_ws_kernel<2, false, false, 0>);
    default: return mode ? f(skin_ws_kernel<1, true, true, 0>) : f(skin_ws_kernel<1, false, false, 0>);
    }
}
template <typename F>
auto with_kernel(int repl, int mode, int variant, bool legacy, int vpt, bool strided, F&& f) {
    if (is_legacy(variant, legacy)) return with_legacy_kernel(repl, mode != 0, variant, f);
    if (variant == 3) return with_ws_layout<3>(repl, mode, vpt, strided, f);
    if (variant == 4) return with_ws_layout<4>(repl, mode, vpt, strided, f);
    return with_ws_layout<0>(repl, mode, vpt, strided, f);
}
// The bulk-store output path is OFF unless DMK_SKIN_BULK_STORE=1 is in the environment.  MEASURED (profiles/r02_skin_bulk_store_experiment.txt):
// correct on hardware, but 5.95 ms per launch against 3.55 ms, whatever the staging depth (2, 4 or 8 buffers per warp).  The bulk-copy
// engine of an SM moves about 23 bytes per clock in total (tools/ubench_out_path.cu mode 2: 64.5 KB of stores per 2770 clocks; the
// shared->shared copies of tools/ubench_s2s.cu: 15 B/clock): the 8 palette replicas (35 KB per character and block) already take half
// of that, and 60 KB of output on top need 4300 clocks per character where the whole item has 3400.  The microbenchmark that promised
// 2.99 against 3.21 ms had no palette copies in it.  Kept as a diagnostic (and as the record of why the output leaves through st.global).
bool strided_enabled() {
    static const bool v = [] {
        const char* e = std::getenv("DMK_SKIN_BULK_STORE");
        return e && std::atoi(e) == 1;
    }();
    return v;
}
int default_vpt() {
    static const int v = [] {
        const char* e = std::getenv("DMK_SKIN_VPT");
        return e && std::atoi(e) == 4 ? 4 : 8;
    }();
    return v;
}
}  // namespace

SkinPlan plan_skin(int vpad, int palette_size, int num_sms, int variant, bool legacy, int ws_vpt) {
    SkinPlan plan{};
    plan.variant = variant;
    plan.legacy = is_legacy(variant, legacy);
    // 8 replicas make the gathers conflict free; palettes too large for that (> 226 slots; > 181 in the round-1 shape) fall back to fewer replicas
    // (and give up ring depth, down to 2 buffers, before giving up replicas)
    plan.replicas = kRepl;
    plan.stages = plan.legacy ? kNumStage : kMaxRing;
    auto fits = [&] { return skin_smem_bytes(palette_size, plan.replicas, plan.legacy, plan.stages) <= kMaxSmem; };
    while (!plan.legacy && plan.stages > 2 && !fits()) --plan.stages;
    while (plan.replicas > 1 && !fits()) plan.replicas /= 2;
    plan.smem = skin_smem_bytes(palette_size, plan.replicas, plan.legacy, plan.stages);
    // warp-specialised kernel, interleaved output, 8 replicas: 4 vertices per thread (two lanes per 8-vertex block) on request
    if (ws_vpt == 0) ws_vpt = default_vpt();
    plan.ws_vpt = (!plan.legacy && variant == 0 && plan.replicas == 8 && ws_vpt == 4) ? 4 : 8;
    // bulk-store output path (meshes stored by cluster_mesh_strided): interleaved output, 8 replicas, pitch a multiple of 64 (a warp's
    // lanes in use are then a multiple of 8, so every round -- lanes x 36 bytes -- starts and ends on a 32-byte sector)
    plan.strided_store = !plan.legacy && variant == 0 && plan.replicas == 8 && plan.ws_vpt == 8 && vpad % 64 == 0 && strided_enabled() &&
                         skin_smem_bytes(palette_size, plan.replicas, false, plan.stages, true) <= kMaxSmem;
    if (plan.strided_store) plan.smem = skin_smem_bytes(palette_size, plan.replicas, false, plan.stages, true);
    const int vpt = (variant == 1 || variant == 2) ? 4 : plan.ws_vpt;
    const int max_threads = variant == 1 ? 320 : variant == 2 ? 448 : plan.legacy ? kMaxThreads : plan.ws_vpt == 4 ? kMaxThreads4 - 32 : kComputeThreads;
    const int max_tile = max_threads * vpt;
    plan.num_tiles = (vpad + max_tile - 1) / max_tile;
    int tv = (vpad + plan.num_tiles - 1) / plan.num_tiles;
    tv = (tv + kVPT - 1) / kVPT * kVPT;      // multiple of 8 for either shape
    if (plan.strided_store) tv = (tv + 63) / 64 * 64;   // every warp's lanes in use are a multiple of 8: rounds are whole sectors
    plan.tile_vertices = tv;
    plan.threads = ((tv / vpt) + 31) / 32 * 32 + (plan.legacy ? 0 : 32);   // + the producer warp
    int occ = 1;
    with_kernel(plan.replicas, plan.strided_store ? 3 : 0, variant, legacy, plan.ws_vpt, plan.strided_store, [&](auto kernel) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, plan.threads, plan.smem) != cudaSuccess || occ < 1) occ = 1;
        return 0;
    });
    int grid = num_sms * occ;
    if (plan.legacy) {   // the round-1 kernel's static split wants a multiple of the tile count; the dynamic assignment uses every SM
        const int tpp = plan.num_tiles < grid ? plan.num_tiles : grid;
        grid = grid / tpp * tpp;
    }
    plan.grid = grid;
    return plan;
}

cudaError_t launch_skin(const SkinParams& p, const SkinPlan& plan, cudaStream_t stream) {
    if (p.n <= 0 || p.vpad <= 0) return cudaSuccess;
    SkinParams q = p;
    q.num_stages = plan.stages;
    if (q.next_item && !plan.legacy) {   // dynamic character assignment: one counter per tile, zero at every launch
        const cudaError_t e = cudaMemsetAsync(q.next_item, 0, sizeof(int32_t) * (size_t)plan.num_tiles, stream);
        if (e != cudaSuccess) return e;
    }
    q.replicate_in_smem = plan.replicate_in_smem ? 1 : 0;
    return with_kernel(plan.replicas, p.vertex_order_mode, plan.variant, plan.legacy, plan.ws_vpt, plan.strided_store, [&, p = q](auto kernel) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem);
        if (e != cudaSuccess) return e;
        if (p.replicate_in_smem) DMK_LAUNCH_CLUSTER_OF_ONE(kernel, plan.grid, plan.threads, plan.smem, stream)(p);
        else DMK_LAUNCH(kernel, plan.grid, plan.threads, plan.smem, stream)(p);
        return cudaGetLastError();
    });
}

}  // namespace dmk
