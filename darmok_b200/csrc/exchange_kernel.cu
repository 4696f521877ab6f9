// exchange_kernel.cu -- K6': visible-palette exchange over peer memory (NVLink 5 / NVSwitch), SURVEY.md section 8(e).
//
// The reference is one process on one GPU; its palette goes from SkeletalAnimationRenderComponent::beforeRenderEntity
// (src/skeleton.cpp:223-251) straight into a bgfx uniform.  With the crowd sharded over G GPUs the palettes of the
// characters FrustumCuller left visible (src/culling.cpp:264-285) have to reach the GPU that renders: that is the one
// exchange of the path.  Here it is ONE kernel per rank that packs and transfers: every warp copies the palette of one
// visible character from the rank's palette buffer into slab[rank] in the RENDER GPU's memory (mapped into this process
// with CUDA IPC) with 128-bit stores over NVLink, and the last block to finish publishes {count, frame} behind a
// system-scope release.  No rank waits for another rank's KERNEL to be resident (NCCL's send/recv pair does, which is
// what stalls next to a persistent grid): a producer only ever waits for a word the render GPU wrote two frames ago.
//
// Protocol (all words live in the header at the start of the render GPU's allocation; frames count 1, 2, 3, ...):
//   published[r]   last frame whose palettes rank r has completely stored            written by rank r, read by the owner
//   count[b][r]    visible characters of rank r in the frame using buffer b          written by rank r, read by the owner
//                  (NOT clamped: a count above the capacity says that only the first `capacity` palettes were stored;
//                  the owner's wait records the frame in overflow_frame and dmk_exchange_status reports it)
//   consumed       last frame the owner has finished reading                         written by the owner, read by all
// The slab is double buffered (frame & 1).  Rank r may store frame f once consumed >= f - 2 (buffer f & 1 was last used
// by frame f - 2); the owner reads frame f once published[r] >= f for every r.  Dependencies only point to earlier
// frames, so the waits cannot form a cycle.  Every wait is bounded (globaltimer): on expiry the kernel raises the
// error word and gives up instead of hanging the GPU.
#include "device_intrinsics.cuh"
#include "kernels.cuh"

namespace dmk {
namespace {

// waits until *word has reached `need` (wrap-safe); false when `timeout_ns` passed first
__device__ bool wait_reached(const uint32_t* word, uint32_t need, uint64_t timeout_ns) {
    if (static_cast<int32_t>(ld_acquire_sys(word) - need) >= 0) return true;
    const uint64_t t0 = global_timer_ns();
    for (;;) {
        __nanosleep(256);
        if (static_cast<int32_t>(ld_acquire_sys(word) - need) >= 0) return true;
        if (global_timer_ns() - t0 > timeout_ns) return false;
    }
}

// One warp per visible character.  palette: this rank's [N][palette_vec4] float4; slab / hdr: peer (or local) memory.
__global__ void __launch_bounds__(256) push_palettes_kernel(const float4* __restrict__ palette, int palette_vec4, const int32_t* __restrict__ ids,
                                                            const int32_t* __restrict__ count, int64_t capacity, float4* __restrict__ slab,
                                                            ExchangeHeader* hdr, int rank, uint32_t frame, uint32_t* done, uint32_t* local_error,
                                                            uint64_t timeout_ns, int from34) {
    __shared__ int go;
    if (threadIdx.x == 0) {
        // flow control: the buffer of this frame was last read for frame - 2
        go = frame <= 2 || wait_reached(&hdr->consumed, frame - 2, timeout_ns);
        if (!go) atomicExch(local_error, kExchangeErrTimeoutProducer);
    }
    __syncthreads();
    const int64_t visible = count[0];
    const int64_t total = visible > capacity ? capacity : visible;
    if (go) {
        const int warps = (gridDim.x * blockDim.x) >> 5;
        const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
        for (int64_t i = warp; i < total; i += warps) {
            if (from34) warp_copy_palette34_as_mat4(reinterpret_cast<const float*>(palette + static_cast<size_t>(ids[i]) * palette_vec4), slab + static_cast<size_t>(i) * palette_vec4, palette_vec4, lane);
            else warp_copy_palette(palette + static_cast<size_t>(ids[i]) * palette_vec4, slab + static_cast<size_t>(i) * palette_vec4, palette_vec4, lane);
        }
    }
    // last block out publishes: fence (release of this thread's stores at system scope), block barrier, ticket
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t ticket = atomicAdd(done, 1u);
        if (ticket == gridDim.x - 1) {
            *done = 0;               // next launch on this stream starts from zero
            __threadfence_system();  // acquire side of the tickets + release side of the words below
            if (*reinterpret_cast<volatile uint32_t*>(local_error) == 0) {   // a block that timed out stored nothing: do not publish
                hdr->count[frame & 1][rank] = static_cast<int32_t>(visible);   // unclamped: the owner sees an overflow
                st_release_sys(&hdr->published[rank], frame);
            }
        }
    }
}

// Copy-engine transport (dmk_crowd_push_visible_palettes with DMK_PUSH_COPY_ENGINE): the palettes are packed into a local
// buffer by the pack kernel, moved by ONE cudaMemcpyAsync into the render GPU's slab (a DMA engine, no SM), and these two
// single-thread kernels do the flow control in front of the copy and the publication behind it.
__global__ void wait_consumed_kernel(ExchangeHeader* hdr, uint32_t frame, uint32_t* local_error, uint64_t timeout_ns) {
    if (frame > 2 && !wait_reached(&hdr->consumed, frame - 2, timeout_ns)) atomicExch(local_error, kExchangeErrTimeoutProducer);
}
__global__ void publish_kernel(ExchangeHeader* hdr, int rank, uint32_t frame, const int32_t* __restrict__ count, const uint32_t* local_error) {
    if (*reinterpret_cast<const volatile uint32_t*>(local_error) != 0) return;   // the flow-control wait gave up: nothing valid was copied
    __threadfence_system();
    hdr->count[frame & 1][rank] = count[0];
    st_release_sys(&hdr->published[rank], frame);
}

// Owner: one thread per rank waits for that rank's frame, then copies its count.
__global__ void wait_published_kernel(ExchangeHeader* hdr, int world, uint32_t frame, int32_t* __restrict__ counts_out, int64_t capacity,
                                      uint64_t timeout_ns) {
    const int r = threadIdx.x;
    if (r >= world) return;
    if (wait_reached(&hdr->published[r], frame, timeout_ns)) {
        const int32_t visible = *reinterpret_cast<volatile int32_t*>(&hdr->count[frame & 1][r]);
        counts_out[r] = visible;
        if (visible > capacity) *reinterpret_cast<volatile uint32_t*>(&hdr->overflow_frame) = frame;   // (frames are waited for in order)
    } else {
        counts_out[r] = 0;
        atomicExch(&hdr->error, kExchangeErrTimeoutOwner);
    }
}

__global__ void release_frame_kernel(ExchangeHeader* hdr, uint32_t frame) {
    __threadfence_system();
    st_release_sys(&hdr->consumed, frame);
}

}  // namespace

cudaError_t launch_push_palettes(const float* palette, int palette_floats, const int32_t* ids, const int32_t* count, int64_t capacity, float* slab,
                                 ExchangeHeader* hdr, int rank, uint32_t frame, uint32_t* done, uint32_t* local_error, uint64_t timeout_ns, int num_ctas,
                                 cudaStream_t stream, bool from34) {
    DMK_LAUNCH(push_palettes_kernel, num_ctas, 256, 0, stream)(reinterpret_cast<const float4*>(palette), palette_floats / 4, ids, count, capacity,
                                                       reinterpret_cast<float4*>(slab), hdr, rank, frame, done, local_error, timeout_ns, from34 ? 1 : 0);
    return cudaGetLastError();
}
cudaError_t launch_wait_published(ExchangeHeader* hdr, int world, uint32_t frame, int32_t* counts_out, int64_t capacity, uint64_t timeout_ns,
                                  cudaStream_t stream) {
    DMK_LAUNCH(wait_published_kernel, 1, kExchangeMaxRanks, 0, stream)(hdr, world, frame, counts_out, capacity, timeout_ns);
    return cudaGetLastError();
}
cudaError_t launch_wait_consumed(ExchangeHeader* hdr, uint32_t frame, uint32_t* local_error, uint64_t timeout_ns, cudaStream_t stream) {
    DMK_LAUNCH(wait_consumed_kernel, 1, 1, 0, stream)(hdr, frame, local_error, timeout_ns);
    return cudaGetLastError();
}
cudaError_t launch_publish(ExchangeHeader* hdr, int rank, uint32_t frame, const int32_t* count, const uint32_t* local_error, cudaStream_t stream) {
    DMK_LAUNCH(publish_kernel, 1, 1, 0, stream)(hdr, rank, frame, count, local_error);
    return cudaGetLastError();
}
cudaError_t launch_release_frame(ExchangeHeader* hdr, uint32_t frame, cudaStream_t stream) {
    DMK_LAUNCH(release_frame_kernel, 1, 1, 0, stream)(hdr, frame);
    return cudaGetLastError();
}

}  // namespace dmk
