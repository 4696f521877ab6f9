// kernels.cuh -- launchers of the sm_100a kernels (one per .cu translation unit).
#pragma once

#include "device_types.cuh"

namespace dmk {

// K1..K3 (pose_kernel.cu, built with -fmad=false)
size_t pose_smem_bytes(const PoseParams& p);
cudaError_t launch_pose(const PoseParams& p, int num_sms, cudaStream_t stream, int* launches);
// K2 + K3 only, from the local matrices of the last sampling launch (p.scratch): further armatures of the same crowd
cudaError_t launch_model(const PoseParams& p, int num_sms, cudaStream_t stream, int* launches);

// K4 (cull_kernel.cu, built with -fmad=false): visibility bitmask; then ordered compaction of visible ids
cudaError_t launch_cull(const CullParams& p, cudaStream_t stream);
// getJointModelMatrixes(dir): debug-bone matrices [count][J][16] from the converted joint matrices [count][J][16]
cudaError_t launch_bone_matrices(const float* models, const int16_t* first_child, int num_joints, int64_t count, const float* dir, float half_turn_sin,
                                 float half_turn_cos, float* out, cudaStream_t stream);
// Transform::update over the group nodes (parents first), any node order, chains up to 32 deep
cudaError_t launch_transform_nodes(const float* pos, const float* rot, const float* scl, const int32_t* parent, int num_nodes, float* world,
                                   float* world_inverse, cudaStream_t stream);
// ids_out[i] = i-th visible instance (ascending); count_out[0] = number of visible; scratch >= ceil(words/1024)+1 ints
cudaError_t launch_compact(const uint32_t* vis_bits, int64_t n, int32_t* ids_out, int32_t* count_out,
                           int32_t* scratch, cudaStream_t stream, int* launches);
size_t compact_scratch_ints(int64_t n);

// K5 (skin_kernel.cu)
struct SkinPlan {
    int threads;        // threads per block (multiple of 32)
    int num_tiles;      // vertex tiles per character
    int tile_vertices;  // vertices per tile (multiple of 8)
    int grid;           // persistent blocks
    int replicas;       // palette replicas in shared memory (bank-conflict-free gathers)
    size_t smem;
    bool replicate_in_smem;   // one copy from global memory per character + REPL - 1 shared-memory copies (else REPL copies from global memory)
    int stages;         // bulk-copy staging depth
    bool legacy;        // the round-1 kernel shape (every warp replicates and skins) instead of producer warp + compute warps
    int variant;        // 0: interleaved [V][9] output, 8 vertices per thread; 1 / 2: planar [9][V] output, 4 per thread, <= 320 / 448 threads;
                        // 3: three float3 streams [3][V][3], 8 per thread; 4: planar [9][V], 8 per thread
    bool strided_store; // the bulk-store output path is available for this shape (used when the mesh is stored for it: vertex_order_mode 3)
    int ws_vpt;         // vertices per thread of the warp-specialised kernel: 8, or 4 (two lanes per 8-vertex block; variant 0 with 8 replicas only)
};
// ws_vpt: 8 or 4, 0 = the default (8, or what the environment variable DMK_SKIN_VPT says)
SkinPlan plan_skin(int vpad, int palette_size, int num_sms, int variant = 0, bool legacy = false, int ws_vpt = 0);
cudaError_t launch_skin(const SkinParams& p, const SkinPlan& plan, cudaStream_t stream);

// K6 payload: gather palettes of the listed characters into a dense buffer
cudaError_t launch_pack_palettes(const float* palette, int palette_floats, const int32_t* ids, const int32_t* count,
                                 int64_t capacity, float* out, int num_sms, cudaStream_t stream, bool from34 = false);

// K6' (exchange_kernel.cu): pack + transfer of the visible palettes into the render GPU's slab over peer memory; the
// owner-side wait for a frame of every rank and the release of its buffer
cudaError_t launch_push_palettes(const float* palette, int palette_floats, const int32_t* ids, const int32_t* count, int64_t capacity, float* slab,
                                 ExchangeHeader* hdr, int rank, uint32_t frame, uint32_t* done, uint32_t* local_error, uint64_t timeout_ns, int num_ctas,
                                 cudaStream_t stream, bool from34 = false);
cudaError_t launch_wait_published(ExchangeHeader* hdr, int world, uint32_t frame, int32_t* counts_out, int64_t capacity, uint64_t timeout_ns,
                                  cudaStream_t stream);
// copy-engine transport: flow control in front of the DMA copy, publication behind it (one thread each)
cudaError_t launch_wait_consumed(ExchangeHeader* hdr, uint32_t frame, uint32_t* local_error, uint64_t timeout_ns, cudaStream_t stream);
cudaError_t launch_publish(ExchangeHeader* hdr, int rank, uint32_t frame, const int32_t* count, const uint32_t* local_error, cudaStream_t stream);
cudaError_t launch_release_frame(ExchangeHeader* hdr, uint32_t frame, cudaStream_t stream);

}  // namespace dmk

namespace dmk {
// device-side animator state machine (animator_kernel.cu, built with -fmad=false)
cudaError_t launch_animator_init(const AnimatorParams& p, float* cmd_speed, float* in_blend, float* in_speed, cudaStream_t stream);
cudaError_t launch_animator_query(const AnimatorParams& p, AnimStateInfo* out, cudaStream_t stream);
cudaError_t launch_animator(const AnimatorParams& p, cudaStream_t stream);
}  // namespace dmk
