// ubench_tma_store.cu -- do cp.async.bulk shared->global stores steal shared-memory bandwidth from LDS?
// 7 warps stream LDS.128 (conflict free); warp 7 optionally pushes smem->global bulk stores at full tilt.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>  // 0: LDS only, 1: LDS + TMA stores, 2: TMA stores only, 3: LDS + STG.256 strided by warp 7
__global__ void __launch_bounds__(256) k(float* out, float* sink, int iters, int stores) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float4* s4 = reinterpret_cast<float4*>(smem);
    for (int i = tid; i < 8192; i += 256) s4[i] = make_float4(i, 1, 2, 3);   // 128 KB
    __syncthreads();
    if (warp < 7) {
        if (MODE == 2) return;
        float acc = 0.f;
        uint32_t a = smem_u32(smem) + tid * 16;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                float4 v;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a + ((u * 4096 + it * 64) & 0xfff0)));
                acc += v.x + v.w;
            }
        }
        if (acc == 12345.f) sink[0] = acc;
    } else {
        if (MODE == 0) return;
        if (MODE == 3) {
            float* base = out + (size_t)blockIdx.x * stores * 2304;
            for (int s = 0; s < stores; ++s)
#pragma unroll
                for (int q = 0; q < 9; ++q) {
                    float* p = base + (size_t)s * 2304 + lane * 72 + q * 8;
                    asm volatile("st.global.v8.f32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p), "f"(1.f) : "memory");
                }
            return;
        }
        if (lane == 0) {
            char* base = reinterpret_cast<char*>(out) + (size_t)blockIdx.x * stores * 9216;
            for (int s = 0; s < stores; ++s) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(base + (size_t)s * 9216),
                             "r"(smem_u32(smem) + (s & 7) * 9216), "r"(9216) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    }
}

int main() {
    const int iters = 20000, stores = 6000;
    float *d, *sink;
    cudaMalloc(&d, (size_t)148 * stores * 9216);
    cudaMalloc(&sink, 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const size_t smem = 128 * 1024;
    cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int mode = 0; mode < 4; ++mode) {
        float best = 1e9;
        for (int it = 0; it < 3; ++it) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148, 256, smem>>>(d, sink, iters, stores);
            if (mode == 1) k<1><<<148, 256, smem>>>(d, sink, iters, stores);
            if (mode == 2) k<2><<<148, 256, smem>>>(d, sink, iters, stores);
            if (mode == 3) k<3><<<148, 256, smem>>>(d, sink, iters, stores);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        const double lds_bytes = 148.0 * 7 * 32 * 16 * 16 * iters, st_bytes = 148.0 * stores * 9216;
        printf("mode %d: %.3f ms | LDS %.1f GB/s (if active) | store %.1f GB/s (if active)\n", mode, best, lds_bytes / best / 1e6, st_bytes / best / 1e6);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
