// device_intrinsics.cuh -- everything in the kernels that is not C++: the PTX wrappers (mbarrier, TMA bulk copy,
// 128/256-bit shared and global accesses, system-scope acquire/release, %globaltimer), the dynamic shared-memory
// declaration and the <<< >>> launch syntax.  The kernels only use these through the names below.
//
// tests/devsim/ builds the same kernel sources with g++ and runs them on the CPU in lockstep (one fiber per CUDA thread)
// to check the kernels' indexing and hand-off protocols without a GPU; it defines DMK_DEVSIM and supplies these names
// itself.  That build is test infrastructure under tests/: the product library is always the nvcc build of this branch.
#pragma once

#ifdef DMK_DEVSIM
#include <dmk_devsim_intrinsics.hpp>
#else

#include <cstdint>
#include <cuda_runtime.h>

// kernel<<<grid, block, dynamic shared bytes, stream>>>(arguments...)  ==  DMK_LAUNCH(kernel, grid, block, bytes, stream)(arguments...)
#define DMK_LAUNCH(kernel, grid, block, smem_bytes, stream) kernel<<<(grid), (block), (smem_bytes), (stream)>>>
// the same as a thread-block cluster of ONE block: what cp.async.bulk shared::cta -> shared::cluster needs even when source and
// destination are the same block's shared memory (a plain launch traps with "illegal instruction": tools/ubench_s2s.cu)
namespace dmk {
template <class... P>
struct ClusterOfOneLaunch {
    void (*kernel)(P...);
    unsigned grid, block;
    size_t smem;
    cudaStream_t stream;
    template <class... A>
    void operator()(A&&... args) const {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(block);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 1;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, kernel, P(args)...);
    }
};
template <class... P>
ClusterOfOneLaunch<P...> cluster_of_one(void (*kernel)(P...), unsigned grid, unsigned block, size_t smem, cudaStream_t stream) {
    return {kernel, grid, block, smem, stream};
}
}  // namespace dmk
#define DMK_LAUNCH_CLUSTER_OF_ONE(kernel, grid, block, smem_bytes, stream) ::dmk::cluster_of_one(kernel, (grid), (block), (smem_bytes), (stream))
// the block's dynamic shared memory as `unsigned char name[]`
#define DMK_DYNAMIC_SMEM(name, alignment) extern __shared__ __align__(alignment) unsigned char name[]

namespace dmk {
namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier -------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    // try_wait suspends the thread in hardware until the phase completes or a time limit passes.  A protocol bug becomes a
    // trap instead of a hung GPU; the bound is in TIME (%globaltimer, 10 s), not in polls, so that a slow phase -- a profiler
    // replay, a preempted context, a side-stream kernel holding the SM -- is never mistaken for one.
    uint32_t done = 0;
    uint32_t spins = 0;
    uint64_t t0 = 0;
    const uint32_t addr = smem_u32(bar);
    do {
        asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
            "selp.u32 %0, 1, 0, P1;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (!done && (++spins & 0xfffu) == 0) {   // every 4096 polls: look at the clock
            uint64_t now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > 10000000000ull) __trap();
        }
    } while (!done);
}
// mbarrier arrive (release at CTA scope): one arrival per warp after its lanes' shared-memory writes
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// TMA bulk copy global -> shared, completion counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// bulk copy shared -> shared inside the CTA (a CTA is a cluster of one: its shared::cta addresses are valid shared::cluster ones)
__device__ __forceinline__ void bulk_s2s(void* dst_smem, const void* src_smem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "r"(smem_u32(src_smem)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// bulk copy shared -> global (the skinning kernel's output path): issued by one thread, tracked in that thread's bulk groups
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, uint32_t src_smem_addr, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(src_smem_addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all but the N most recent groups of this thread have finished READING their shared-memory source
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy writes to shared memory (st.shared) made visible to the async proxy (bulk copies) of this CTA
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void sts32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }

// ---- wide global / shared accesses --------------------------------------------------------------------------------
#ifndef DMK_ST_V8_OP
#define DMK_ST_V8_OP "st.global.v8.f32"   // (experiments: -DDMK_ST_V8_OP='"st.global.cs.v8.f32"')
#endif
__device__ __forceinline__ void st_v8(float* p, const float* v) {
    asm volatile(DMK_ST_V8_OP " [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
                 "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}
__device__ __forceinline__ void st_v8(float* p, float a, float b, float c, float d, float e, float f, float g, float h) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d), "f"(e), "f"(f), "f"(g), "f"(h)
                 : "memory");
}
// the same where `take` is non-zero (a warp whose lanes store different register sets: one predicated store per set)
__device__ __forceinline__ void st_v8_if(float* p, const float* v, uint32_t take) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "setp.ne.u32 P1, %9, 0;\n"
        "@P1 st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n"
        "}\n" ::"l"(p),
        "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "r"(take)
        : "memory");
}
__device__ __forceinline__ void ld_v8(const float* p, float* v) {
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p));
}
__device__ __forceinline__ void st_v4(float* p, float a, float b, float c, float d) {
    asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
// Predicated gather: an influence with zero weight contributes w * P = 0, so its rows are not fetched.  A quarter-warp
// whose 8 lanes all skip costs no shared-memory wavefront (meshes whose vertices are grouped by influence count --
// DMK_MESH_SORT_BY_INFLUENCES, or any mesh with spatially coherent skinning -- gather proportionally fewer bytes).
__device__ __forceinline__ float4 lds128_if(uint32_t addr, uint32_t take) {
    float4 v;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "setp.ne.u32 P1, %5, 0;\n"
        "mov.f32 %0, 0f00000000;\n"
        "mov.f32 %1, 0f00000000;\n"
        "mov.f32 %2, 0f00000000;\n"
        "mov.f32 %3, 0f00000000;\n"
        "@P1 ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];\n"
        "}\n"
        : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
        : "r"(addr), "r"(take));
    return v;
}

// Gather unless the row is already in registers: `fallback` (the block's bone, read once per character) where `take` is 0.
__device__ __forceinline__ float4 lds128_or(uint32_t addr, uint32_t take, const float4& fallback) {
    float4 v;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "setp.ne.u32 P1, %5, 0;\n"
        "mov.f32 %0, %6;\n"
        "mov.f32 %1, %7;\n"
        "mov.f32 %2, %8;\n"
        "mov.f32 %3, %9;\n"
        "@P1 ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];\n"
        "}\n"
        : "=&f"(v.x), "=&f"(v.y), "=&f"(v.z), "=&f"(v.w)
        : "r"(addr), "r"(take), "f"(fallback.x), "f"(fallback.y), "f"(fallback.z), "f"(fallback.w));
    return v;
}

// Predicated gather that leaves the registers as they are where `take` is 0 (no zero fill): the caller multiplies by a
// zero weight then, so the registers must hold something finite -- zero, or a row fetched earlier.
__device__ __forceinline__ void lds128_keep(float4& v, uint32_t addr, uint32_t take) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "setp.ne.u32 P1, %5, 0;\n"
        "@P1 ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];\n"
        "}\n"
        : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w)
        : "r"(addr), "r"(take));
}

// One warp copies one palette (palette_vec4 float4): 8 loads in flight per lane before the first store, so that a handful of
// warps keeps enough bytes on the wire -- with one load per store a whole SM moved 14 GB/s over NVLink (r02, 2 GPUs), bound by
// the latency of its own loads.
__device__ __forceinline__ void warp_copy_palette(const float4* __restrict__ src, float4* __restrict__ dst, int palette_vec4, int lane) {
    constexpr int kInFlight = 8;
    for (int base = 0; base < palette_vec4; base += 32 * kInFlight) {
        float4 v[kInFlight];
#pragma unroll
        for (int u = 0; u < kInFlight; ++u) {
            const int k = base + u * 32 + lane;
            if (k < palette_vec4) v[u] = __ldg(src + k);
        }
#pragma unroll
        for (int u = 0; u < kInFlight; ++u) {
            const int k = base + u * 32 + lane;
            if (k < palette_vec4) dst[k] = v[u];
        }
    }
}

// The same from the 3x4 row-major palette the skinning kernel reads ([slot][3 rows + 1 unused row][4]) into column-major mat4:
// float4 k of the destination is column k % 4 of slot k / 4, bottom row (0, 0, 0, 1).
__device__ __forceinline__ void warp_copy_palette34_as_mat4(const float* __restrict__ src34, float4* __restrict__ dst, int palette_vec4, int lane) {
    constexpr int kInFlight = 8;
    for (int base = 0; base < palette_vec4; base += 32 * kInFlight) {
        float4 v[kInFlight];
#pragma unroll
        for (int u = 0; u < kInFlight; ++u) {
            const int k = base + u * 32 + lane;
            if (k < palette_vec4) {
                const float* s = src34 + (size_t)(k >> 2) * 16 + (k & 3);
                v[u] = make_float4(__ldg(s), __ldg(s + 4), __ldg(s + 8), (k & 3) == 3 ? 1.f : 0.f);
            }
        }
#pragma unroll
        for (int u = 0; u < kInFlight; ++u) {
            const int k = base + u * 32 + lane;
            if (k < palette_vec4) dst[k] = v[u];
        }
    }
}

// ---- system scope (peer memory over NVLink) -----------------------------------------------------------------------
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint64_t global_timer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

}  // namespace
}  // namespace dmk

#endif  // DMK_DEVSIM
