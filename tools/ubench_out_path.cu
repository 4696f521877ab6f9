// ubench_out_path.cu -- the skinning kernel's output side in isolation: per "item" a warp produces 256 vertices x 36 B next to G
// conflict-free LDS.128 gathers per lane and vertex.  How should the 9216 bytes leave the SM?
//   mode 0: as the kernel does today: a lane owns 8 consecutive vertices (288 B = 9 sectors), st.global.v8.f32 as sectors complete
//   mode 1: a lane owns vertices lane, lane + 32, ...: a round of 32 vertices is 1152 contiguous bytes, written to shared memory
//           (9 x STS.32 per lane, stride 9 words: conflict free) and sent by ONE cp.async.bulk shared -> global per round (double buffered)
//   mode 2: mode 1 without the gathers (the store path alone);  mode 3: mode 0 without the gathers
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_out_path.bin tools/ubench_out_path.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE, int G>
__global__ void __launch_bounds__(256, 1) k(float* out, float* sink, int items) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float4* s4 = reinterpret_cast<float4*>(smem);
    for (int i = tid; i < 8192; i += 256) s4[i] = make_float4(i * 1e-9f, 1e-9f, 2e-9f, 3e-9f);   // 128 KB of "palette replicas"
    __syncthreads();
    if (warp == 7) return;
    const bool gather = MODE < 2;
    const bool tma = (MODE == 1 || MODE == 2);
    unsigned char* stage = smem + 128 * 1024 + warp * 2 * 1152;     // [2][1152] per warp
    const uint32_t stage_a = smem_u32(stage);
    const uint32_t lds_base = smem_u32(smem) + lane * 16;
    float* wout = out + ((size_t)blockIdx.x * 7 + warp) * (size_t)items * 2304;   // 2304 floats per item and warp
    float acc = 0.f;
    for (int it = 0; it < items; ++it) {
        float res[72];
#pragma unroll
        for (int v = 0; v < 8; ++v) {
            float m[9];
#pragma unroll
            for (int c = 0; c < 9; ++c) m[c] = acc + (float)c;
            if (gather) {
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    float4 x;
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w)
                                 : "r"(lds_base + (uint32_t)(((it * 8 + v) * G + g) * 512 & 0x1fe00)));
                    m[(g * 4 + 0) % 9] = fmaf(x.x, 1.0001f, m[(g * 4 + 0) % 9]);
                    m[(g * 4 + 1) % 9] = fmaf(x.y, 1.0001f, m[(g * 4 + 1) % 9]);
                    m[(g * 4 + 2) % 9] = fmaf(x.z, 1.0001f, m[(g * 4 + 2) % 9]);
                    m[(g * 4 + 3) % 9] = fmaf(x.w, 1.0001f, m[(g * 4 + 3) % 9]);
                }
            }
            acc = m[0] * 1e-9f;
            if (!tma) {
#pragma unroll
                for (int c = 0; c < 9; ++c) res[v * 9 + c] = m[c];
                float* o = wout + (size_t)it * 2304 + lane * 72;
#pragma unroll
                for (int sct = (v * 9) / 8; sct < ((v + 1) * 9) / 8; ++sct)
                    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(o + sct * 8), "f"(res[sct * 8]), "f"(res[sct * 8 + 1]),
                                 "f"(res[sct * 8 + 2]), "f"(res[sct * 8 + 3]), "f"(res[sct * 8 + 4]), "f"(res[sct * 8 + 5]), "f"(res[sct * 8 + 6]),
                                 "f"(res[sct * 8 + 7]) : "memory");
            } else {
                const uint32_t buf = stage_a + (v & 1) * 1152;
                // the bulk store that read this buffer two rounds ago must have finished READING it
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                __syncwarp();
#pragma unroll
                for (int c = 0; c < 9; ++c) asm volatile("st.shared.f32 [%0], %1;" ::"r"(buf + (lane * 9 + c) * 4), "f"(m[c]) : "memory");
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(wout + (size_t)it * 2304 + v * 288), "r"(buf), "r"(1152)
                                 : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
        }
    }
    if (tma && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (acc == 12345.f) sink[0] = acc;
}

template <int MODE, int G>
float run(float* d, float* sink, int items) {
    const size_t smem = 128 * 1024 + 7 * 2 * 1152;
    cudaFuncSetAttribute(k<MODE, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k<MODE, G><<<148, 256, smem>>>(d, sink, items);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    const int items = 1900;   // 148 x 7 x 1900 x 9216 B = 18.1 GB, the skinning kernel's output per launch
    const size_t bytes = (size_t)148 * 7 * items * 9216;
    float *d, *sink;
    cudaMalloc(&d, bytes);
    cudaMalloc(&sink, 4);
    const char* names[4] = {"LDS + st.v8 strided (today)", "LDS + STS + bulk store per 32 vertices", "STS + bulk store alone", "st.v8 strided alone"};
    float t[4] = {run<0, 6>(d, sink, items), run<1, 6>(d, sink, items), run<2, 6>(d, sink, items), run<3, 6>(d, sink, items)};
    for (int m = 0; m < 4; ++m) printf("mode %d  %-42s %.3f ms  %.0f GB/s\n", m, names[m], t[m], bytes / t[m] / 1e6);
    float t12[2] = {run<0, 12>(d, sink, items), run<1, 12>(d, sink, items)};
    printf("12 gathers per vertex (the caller-order kernel): st.v8 %.3f ms, bulk store %.3f ms\n", t12[0], t12[1]);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
