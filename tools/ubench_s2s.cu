// ubench_s2s.cu -- does cp.async.bulk shared::cta -> shared::cluster work inside ONE CTA (plain launch / cluster of 1), and how fast is it?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_s2s.bin tools/ubench_s2s.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{ .reg .pred P1; mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2; selp.u32 %0, 1, 0, P1; }" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return done != 0;
}
template <bool MAPA>
__device__ __forceinline__ void bulk_s2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    uint32_t d = smem_u32(dst), b = smem_u32(bar);
    if (MAPA) {
        uint32_t rank;
        asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
        asm volatile("mapa.shared::cluster.u32 %0, %0, %1;" : "+r"(d) : "r"(rank));
        asm volatile("mapa.shared::cluster.u32 %0, %0, %1;" : "+r"(b) : "r"(rank));
    }
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "r"(smem_u32(src)), "r"(bytes), "r"(b) : "memory");
}

template <bool MAPA>
__global__ void k(int iters, int copies, uint32_t bytes, uint32_t stride, long long* cycles, int* ok) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
    float* a = reinterpret_cast<float*>(smem + 128);
    for (uint32_t i = threadIdx.x; i < bytes / 4; i += blockDim.x) a[i] = (float)i + blockIdx.x;
    if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes of `a` visible to the async proxy
    __syncthreads();
    long long t0 = clock64();
    if (threadIdx.x == 0) {
        uint32_t phase = 0;
        for (int it = 0; it < iters; ++it) {
            mbar_expect_tx(bar, bytes * copies);
            for (int r = 1; r <= copies; ++r) bulk_s2s<MAPA>(smem + 128 + (size_t)r * stride, a, bytes, bar);
            long long spin = 0;
            while (!mbar_try(bar, phase)) if (++spin > (1ll << 24)) { ok[blockIdx.x] = -1; return; }
            phase ^= 1;
        }
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    int good = 1;
    for (int r = 1; r <= copies; ++r) {
        const float* b = reinterpret_cast<const float*>(smem + 128 + (size_t)r * stride);
        for (uint32_t i = threadIdx.x; i < bytes / 4; i += blockDim.x) good &= (b[i] == (float)i + blockIdx.x);
    }
    if (!good) ok[blockIdx.x] = 0;
}

template <bool MAPA>
int run(const char* name, bool cluster_launch) {
    const int iters = 2000, copies = 7, blocks = 148;
    const uint32_t bytes = 4352, stride = 4352 + 16;
    const size_t smem = 128 + (size_t)(copies + 1) * stride;
    long long* cyc; int* ok;
    cudaMalloc(&cyc, blocks * sizeof(long long)); cudaMalloc(&ok, blocks * sizeof(int));
    int ones[148]; for (int i = 0; i < 148; ++i) ones[i] = 1;
    cudaMemcpy(ok, ones, sizeof ones, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k<MAPA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaError_t e;
    if (cluster_launch) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, k<MAPA>, iters, copies, bytes, stride, cyc, ok);
    } else {
        k<MAPA><<<blocks, 128, smem>>>(iters, copies, bytes, stride, cyc, ok);
        e = cudaGetLastError();
    }
    cudaError_t e2 = cudaDeviceSynchronize();
    if (e != cudaSuccess || e2 != cudaSuccess) { printf("%s: %s / %s\n", name, cudaGetErrorString(e), cudaGetErrorString(e2)); return 1; }
    long long h[148]; int hok[148];
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost); cudaMemcpy(hok, ok, sizeof hok, cudaMemcpyDeviceToHost);
    int good = 1; long long mx = 0;
    for (int i = 0; i < blocks; ++i) { good &= hok[i] == 1; if (h[i] > mx) mx = h[i]; }
    printf("%s: data %s, %.1f cycles per %d x %u B (%.1f B/cycle/SM)\n", name, good ? "ok" : "WRONG", (double)mx / iters, copies, bytes, (double)copies * bytes * iters / mx);
    return 0;
}

int main(int argc, char** argv) {   // one variant per process: an illegal instruction kills the context
    const int which = argc > 1 ? atoi(argv[1]) : 0;
    if (which == 0) return run<false>("plain launch, shared::cta addresses", false);
    if (which == 1) return run<true>("plain launch, mapa addresses", false);
    if (which == 2) return run<false>("cluster(1) launch, shared::cta addresses", true);
    return run<true>("cluster(1) launch, mapa addresses", true);
}
