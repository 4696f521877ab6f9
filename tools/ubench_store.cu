// ubench_store.cu -- microbenchmark: cost of global store patterns on sm_100a (time + l1tex wavefronts under ncu).
// Patterns write the same N bytes:  0: st.v8 per lane, lanes 288 B apart (the skin kernel's pattern)
//                                   1: st.v4 fully coalesced (lane i -> base + 16 i)
//                                   2: st.v8 coalesced (lane i -> base + 32 i)
//                                   3: st.v4 x2 per 32 B, lanes 288 B apart
//                                   4: st.v8 per lane, lanes 96 B apart (three attribute planes of float3, 8 vertices per thread)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_store tools/ubench_store.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

__device__ __forceinline__ void st8(float* p, float v) {
    asm volatile("st.global.v8.f32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void st4(float* p, float v) {
    asm volatile("st.global.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(p), "f"(v) : "memory");
}

// g_next != nullptr: the blocks take chunks from a counter (dynamic) instead of each writing its own chunks_per_block chunks -- the
// static split lasts as long as the slowest SM (the skinning kernel's finding, profiles/r02_skin_cta_finish_times_static.txt)
__device__ unsigned long long g_counter;
template <int PATTERN>
__global__ void __launch_bounds__(256) k(float* out, size_t chunks_per_block, float v, int dynamic = 0) {
    // each block writes chunks of 256 threads * 288 B = 73728 B
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ unsigned long long s_chunk;
    const size_t total = chunks_per_block * gridDim.x;
    for (size_t c = 0;; ++c) {
        size_t chunk;
        if (dynamic) {
            __syncthreads();
            if (tid == 0) s_chunk = atomicAdd(&g_counter, 1ull);
            __syncthreads();
            chunk = s_chunk;
            if (chunk >= total) break;
        } else {
            if (c >= chunks_per_block) break;
            chunk = (size_t)blockIdx.x * chunks_per_block + c;
        }
        float* base = out + chunk * (256 * 72);
        float* wbase = base + warp * (32 * 72);
        if (PATTERN == 0) {
#pragma unroll
            for (int s = 0; s < 9; ++s) st8(wbase + lane * 72 + s * 8, v);
        } else if (PATTERN == 1) {
#pragma unroll
            for (int s = 0; s < 18; ++s) st4(wbase + s * 128 + lane * 4, v);
        } else if (PATTERN == 2) {
#pragma unroll
            for (int s = 0; s < 9; ++s) st8(wbase + s * 256 + lane * 8, v);
        } else if (PATTERN == 3) {
#pragma unroll
            for (int s = 0; s < 9; ++s) { st4(wbase + lane * 72 + s * 8, v); st4(wbase + lane * 72 + s * 8 + 4, v); }
        } else {
            // plane a of the chunk = 256 threads * 24 floats; a warp's part of it is 32 * 24 floats
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int s = 0; s < 3; ++s) st8(base + a * (256 * 24) + warp * (32 * 24) + lane * 24 + s * 8, v);
        }
    }
}

int main(int argc, char** argv) {
    const size_t chunks_per_block = 1600, blocks = 148;
    const size_t bytes = blocks * chunks_per_block * 256 * 288;
    float* d;
    cudaMalloc(&d, bytes);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    if (argc >= 2 && !strcmp(argv[1], "dynamic")) {
        // ubench_store dynamic: every pattern with the static split and with chunks taken from a counter
        for (int pat = 0; pat < 5; ++pat) {
            if (pat == 3) continue;
            for (int dyn = 0; dyn < 2; ++dyn) {
                float best = 1e9;
                for (int it = 0; it < 4; ++it) {
                    unsigned long long zero = 0;
                    cudaMemcpyToSymbol(g_counter, &zero, sizeof(zero));
                    cudaEventRecord(e0);
                    if (pat == 0) k<0><<<blocks, 256>>>(d, chunks_per_block, 1.f, dyn);
                    if (pat == 1) k<1><<<blocks, 256>>>(d, chunks_per_block, 1.f, dyn);
                    if (pat == 2) k<2><<<blocks, 256>>>(d, chunks_per_block, 1.f, dyn);
                    if (pat == 4) k<4><<<blocks, 256>>>(d, chunks_per_block, 1.f, dyn);
                    cudaEventRecord(e1);
                    cudaEventSynchronize(e1);
                    float ms; cudaEventElapsedTime(&ms, e0, e1);
                    if (ms < best) best = ms;
                }
                printf("pattern %d %s: %.3f ms  %.1f GB/s\n", pat, dyn ? "dynamic" : "static ", best, bytes / best / 1e6);
            }
        }
        printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
        return 0;
    }
    if (argc >= 3 && !strcmp(argv[1], "sustained")) {
        // ubench_store sustained <pattern 0 | 1 | 4> [launches]: the same launch back to back (what does the power cap leave of it?)
        const int pat = atoi(argv[2]), launches = argc >= 4 ? atoi(argv[3]) : 600;
        for (int a = 0; a < launches; a += 50) {
            cudaEventRecord(e0);
            for (int it = 0; it < 50; ++it) {
                if (pat == 0) k<0><<<blocks, 256>>>(d, chunks_per_block, 1.f);
                else if (pat == 1) k<1><<<blocks, 256>>>(d, chunks_per_block, 1.f);
                else k<4><<<blocks, 256>>>(d, chunks_per_block, 1.f);
            }
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("pattern %d launches %d-%d: %.3f ms each  %.1f GB/s\n", pat, a, a + 50, ms / 50, bytes / (ms / 50) / 1e6);
        }
        return 0;
    }
    for (int pat = 0; pat < 5; ++pat) {
        float best = 1e9;
        for (int it = 0; it < 4; ++it) {
            cudaEventRecord(e0);
            if (pat == 0) k<0><<<blocks, 256>>>(d, chunks_per_block, 1.f);
            if (pat == 1) k<1><<<blocks, 256>>>(d, chunks_per_block, 1.f);
            if (pat == 2) k<2><<<blocks, 256>>>(d, chunks_per_block, 1.f);
            if (pat == 3) k<3><<<blocks, 256>>>(d, chunks_per_block, 1.f);
            if (pat == 4) k<4><<<blocks, 256>>>(d, chunks_per_block, 1.f);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        printf("pattern %d: %.3f ms  %.1f GB/s  (%zu bytes)\n", pat, best, bytes / best / 1e6, bytes);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
