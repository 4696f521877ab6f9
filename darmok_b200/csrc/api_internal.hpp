// api_internal.hpp -- what the translation units of the C ABI share: the objects behind the opaque handles of include/dmk.h, error
// reporting, device allocation helpers and the kernel timer.  (api_context.cu, api_assets.cu, api_crowd.cu, api_results.cu,
// api_exchange.cu: host side only; the kernels are in *_kernel.cu.)
//
// No CPU fallback lives here: every compute entry point launches the sm_100a kernels or fails with DMK_ERR_CUDA.
#pragma once

#include "../../include/dmk.h"

#include "animator_flatten.hpp"
#include "device_types.cuh"
#include "host_math.hpp"
#include "kernels.cuh"
#include "mesh_cluster.hpp"
#include "ozz_io.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <new>
#include <numeric>
#include <string>
#include <vector>

using namespace dmk;


// ------------------------------------------------------------------------------------------------------------
// objects behind the opaque handles
// ------------------------------------------------------------------------------------------------------------
struct dmk_context_s {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool owns_stream = false;
    int num_sms = 0;
    std::string err;
    uint64_t launches = 0;
    // grow-only scratch of the host-buffer cull entry point
    void* cull_scratch = nullptr;
    size_t cull_scratch_bytes = 0;
    // optional per-kernel timing
    bool timing = false;
    struct Timed { int kernel; cudaEvent_t start, stop; };
    std::vector<Timed> pending;
    std::vector<cudaEvent_t> event_pool;
    double kernel_ms[DMK_KERNEL_COUNT] = {};
    uint64_t kernel_launches[DMK_KERNEL_COUNT] = {};
};

struct dmk_skeleton_s {
    dmk_context_t ctx;
    SkeletonData data;
    int16_t* d_parents = nullptr;
    uint16_t* d_schedule = nullptr;
    float* d_rest = nullptr;  // [P][10]
    int rounds = 0;
};

struct dmk_clip_s {
    dmk_context_t ctx;
    AnimationData data;
};

struct dmk_clipset_s {
    dmk_context_t ctx;
    int num_clips = 0;
    int padded_tracks = 0;
    std::vector<float> durations;
    uint2* d_dir = nullptr;
    float* d_ratio[3] = {nullptr, nullptr, nullptr};
    uint2* d_value[3] = {nullptr, nullptr, nullptr};
    int32_t* d_stream[3] = {nullptr, nullptr, nullptr};
    float* d_duration = nullptr;
    // bucket table (see ClipSetDev); absent when a clip does not admit one
    uint2* d_bucket_dir = nullptr;
    float4* d_bucket_a = nullptr;
    uint4* d_bucket_b = nullptr;
    uint2* d_bucket_c = nullptr;
    bool has_buckets = false;
};

struct dmk_armature_s {
    dmk_context_t ctx;
    int num_joints = 0;
    std::vector<int32_t> remap;
    int32_t* d_remap = nullptr;
    float* d_inv_bind = nullptr;  // [K][12]
};

struct dmk_mesh_s {
    dmk_context_t ctx;
    int num_vertices = 0;
    int vpad = 0;
    int palette_size = 0;
    float *d_pos = nullptr, *d_nrm = nullptr, *d_tan = nullptr, *d_weights = nullptr;
    uint32_t* d_slots = nullptr;
    std::vector<int32_t> order;  // order[i] = caller's index of the vertex stored at position i
    bool sorted = false;
    int sort_block = 0;          // vertices per thread the influence grouping was dealt for (8 or 4)
    bool clustered = false;      // DMK_MESH_CLUSTER_BONES: every 8-vertex block names its common bone in entry 0 of its first vertex
    bool strided = false;        // ... stored in the strided layout of the bulk-store output path (cluster_mesh_strided; pitch a multiple of 64)
    int64_t cluster_gather_wavefronts = 0;
    int cluster_shared_blocks = 0;
};

struct dmk_crowd_s {
    dmk_context_t ctx;
    dmk_skeleton_t skel;
    dmk_clipset_t clips;
    dmk_armature_t arm;
    dmk_mesh_t mesh;
    int64_t n = 0;
    int max_layers = 0;
    int num_layers = 0;
    float threshold = 0.f;
    bool layers_set = false, has_loop = false, instances_set = false, camera_set = false;
    // layer state [max_layers][n]
    int32_t* d_clip = nullptr;
    float *d_ratio = nullptr, *d_weight = nullptr, *d_speed = nullptr;
    uint8_t *d_loop = nullptr, *d_looped = nullptr;
    PoseProgram* d_program = nullptr;  // [n], allocated by dmk_crowd_set_programs
    bool use_program = false;
    // pinned staging for the per-frame host state machine
    // two pinned staging buffers used alternately: the host may fill one while the copies out of the other are in flight
    void* h_stage[2] = {nullptr, nullptr};
    size_t h_stage_bytes[2] = {0, 0};
    cudaEvent_t stage_done[2] = {nullptr, nullptr};
    int stage_next = 0;
    // results
    float *d_palette = nullptr, *d_palette34 = nullptr, *d_skinned = nullptr;
    float *d_models = nullptr, *d_locals = nullptr;
    float* d_local_matrices = nullptr;  // [n][P][16] sampling kernel -> model kernel
    int32_t* d_key_indices = nullptr;
    int32_t* d_error = nullptr;
    int32_t* h_error = nullptr;  // pinned
    // cull
    float *d_world_inverse = nullptr, *d_bbox = nullptr;
    float *d_position = nullptr, *d_rotation = nullptr, *d_scale = nullptr;
    bool use_trs = false;
    // group nodes the entities' Transforms hang under (dmk_crowd_set_transform_nodes)
    int num_nodes = 0;
    float *d_node_position = nullptr, *d_node_rotation = nullptr, *d_node_scale = nullptr, *d_node_world = nullptr, *d_node_world_inverse = nullptr;
    int32_t *d_node_parent = nullptr, *d_entity_parent = nullptr;
    bool keep_world_inverse = false;   // dmk_crowd_get_world_inverses was asked for: the cull also stores what it derives
    uint32_t* d_vis_bits = nullptr;
    int32_t *d_visible_ids = nullptr, *d_visible_count = nullptr, *d_scratch = nullptr;
    // pinned ring for DMK_STEP_READBACK_VISIBILITY: [words] bits + 1 count, two frames deep
    uint32_t* h_vis[2] = {nullptr, nullptr};
    cudaEvent_t vis_ready[2] = {nullptr, nullptr};
    int vis_head = 0, vis_pending = 0;
    float corners[24] = {0};
    bool culled_once = false;
    bool force_search = false;  // tests: exercise the per-track binary search even when a bucket table exists
    SkinPlan plan{};
    // further (armature, mesh) pairs skinned from the same animator (dmk_crowd_add_skin); skin 0 is the pair above
    struct ExtraSkin {
        dmk_armature_t arm = nullptr;
        dmk_mesh_t mesh = nullptr;
        float *d_palette = nullptr, *d_palette34 = nullptr, *d_skinned = nullptr;
        SkinPlan plan{};
    };
    std::vector<ExtraSkin> extra_skins;
    int reserved_sms = 0;
    // dmk_crowd_update_layers (the per-frame upload) copies on its own stream, behind the last step that read the layer arrays
    // (layers_idle, recorded after the pose part of a step) and in front of the next one (upload_done): the copy runs under the
    // skinning kernel of the frame before instead of in front of the pose kernels
    cudaStream_t upload_stream = nullptr;
    cudaEvent_t layers_idle = nullptr, upload_done = nullptr;
    bool layers_idle_recorded = false, upload_pending = false;
    int32_t* d_skin_next = nullptr;   // [skin_next_capacity] per-tile counters of the skinning kernel's dynamic character assignment
    int skin_next_capacity = 0;
    int skin_variant = 0;   // DMK_OPT_SKIN_PLANAR_OUTPUT
    bool skin_legacy = false;   // DMK_OPT_SKIN_ROUND1_KERNEL
    bool pose_fused = false;   // DMK_OPT_POSE_FUSED
    bool skin_s2s = false;     // DMK_OPT_SKIN_REPLICATE_IN_SMEM
    bool palette_3x4_only = false;   // DMK_OPT_PALETTE_3X4_ONLY: d_palette is not written; consumers expand d_palette34
    // device-side animators (dmk_crowd_set_animator)
    AnimStateDef* d_anim_states = nullptr;
    AnimClipDef* d_anim_clips = nullptr;
    AnimTransitionDef* d_anim_transitions = nullptr;
    float* d_anim_tables = nullptr;
    uint32_t* d_anim_slots = nullptr;
    int32_t* d_anim_transition = nullptr;
    float* d_anim_transition_time = nullptr;
    AnimEvents* d_anim_events = nullptr;
    AnimStateInfo* d_anim_info = nullptr;
    int32_t *d_anim_cmd = nullptr, *d_anim_status = nullptr;
    float *d_anim_cmd_speed = nullptr, *d_anim_blend = nullptr, *d_anim_speed = nullptr;
    int anim_num_states = 0, anim_num_transitions = 0, anim_layers = 0;
    bool animator_set = false, animator_ticked = false;
};

// visible-palette exchange over peer memory (exchange_kernel.cu)
struct dmk_exchange_s {
    enum Kind { Owner, Opened, Attached };
    dmk_context_t ctx = nullptr;
    Kind kind = Owner;
    int world = 0, rank = 0, palette_size = 0;
    int64_t capacity = 0;
    void* base = nullptr;             // ExchangeHeader, then float[2][world][capacity][palette_size][16], in the render GPU's memory
    uint32_t* d_local = nullptr;      // this rank's own words: {tickets of the push grid, error}
    int32_t* d_counts = nullptr;      // owner: [2][kExchangeMaxRanks] counts of the waited frame, per buffer
    bool staged = false;              // dmk_crowd_stage_visible_palettes ran, dmk_exchange_send_staged has not yet
    float* d_stage = nullptr;         // copy-engine transport: this rank's packed palettes [capacity][palette_size][16] (allocated at first use)
    uint32_t last_pushed = 0, last_waited = 0, last_released = 0;
    uint64_t timeout_ns = 30ull * 1000 * 1000 * 1000;
    size_t frame_floats() const { return static_cast<size_t>(world) * static_cast<size_t>(capacity) * static_cast<size_t>(palette_size) * 16; }
    size_t bytes() const { return kExchangeHeaderBytes + 2 * frame_floats() * sizeof(float); }
    ExchangeHeader* header() const { return static_cast<ExchangeHeader*>(base); }
    float* buffer(uint32_t frame) const {
        return reinterpret_cast<float*>(static_cast<char*>(base) + kExchangeHeaderBytes) + (frame & 1u) * frame_floats();
    }
};

// ------------------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------------------
extern thread_local std::string g_create_error;   // creation errors before a context exists (api_context.cu)

static dmk_status fail(dmk_context_t ctx, dmk_status code, const std::string& msg) {
    if (ctx) ctx->err = msg; else g_create_error = msg;
    return code;
}
static dmk_status cuda_fail(dmk_context_t ctx, cudaError_t e, const char* what) {
    return fail(ctx, DMK_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define DMK_CUDA(ctx, expr)                                        \
    do {                                                           \
        cudaError_t e__ = (expr);                                  \
        if (e__ != cudaSuccess) return cuda_fail(ctx, e__, #expr); \
    } while (0)

// Owner of a handle under construction: a failing create call releases the device memory it has allocated so far through
// the handle's own destroy function.
template <typename T, void (*Destroy)(T*)>
struct HandleDeleter {
    void operator()(T* p) const { Destroy(p); }
};
template <typename T, void (*Destroy)(T*)>
using Building = std::unique_ptr<T, HandleDeleter<T, Destroy>>;

template <typename T>
static cudaError_t dev_alloc(T** p, size_t count) {
    *p = nullptr;
    if (count == 0) return cudaSuccess;
    return cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T));
}
template <typename T>
static cudaError_t upload(dmk_context_t ctx, T** p, const std::vector<T>& v) {
    cudaError_t e = dev_alloc(p, v.size());
    if (e != cudaSuccess || v.empty()) return e;
    e = cudaMemcpyAsync(*p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(ctx->stream);  // v may be a temporary
}
static void dev_free(void* p) {
    if (p) cudaFree(p);
}
static dmk_status bind(dmk_context_t ctx) {
    DMK_CUDA(ctx, cudaSetDevice(ctx->device));
    return DMK_OK;
}

namespace {
// RAII scope that brackets a launch with CUDA events when kernel timing is on
struct KernelTimer {
    dmk_context_t ctx;
    dmk_context_s::Timed t{};
    bool on;
    cudaStream_t stream;
    KernelTimer(dmk_context_t c, int kernel, cudaStream_t on_stream = nullptr) : ctx(c), on(c->timing), stream(on_stream ? on_stream : c->stream) {
        if (!on) return;
        t.kernel = kernel;
        for (cudaEvent_t* e : {&t.start, &t.stop}) {
            if (!ctx->event_pool.empty()) {
                *e = ctx->event_pool.back();
                ctx->event_pool.pop_back();
            } else if (cudaEventCreate(e) != cudaSuccess) {
                on = false;
                return;
            }
        }
        cudaEventRecord(t.start, stream);
    }
    ~KernelTimer() {
        if (!on) return;
        cudaEventRecord(t.stop, stream);
        ctx->pending.push_back(t);
    }
};
void drain_timers(dmk_context_t ctx) {
    for (auto& t : ctx->pending) {
        float ms = 0.f;
        if (cudaEventSynchronize(t.stop) == cudaSuccess && cudaEventElapsedTime(&ms, t.start, t.stop) == cudaSuccess) {
            ctx->kernel_ms[t.kernel] += ms;
            ++ctx->kernel_launches[t.kernel];
        }
        ctx->event_pool.push_back(t.start);
        ctx->event_pool.push_back(t.stop);
    }
    ctx->pending.clear();
}

}  // namespace

// The context's stream waits for a pending per-frame upload (dmk_crowd_update_layers) before anything touches the layer arrays.
static inline dmk_status join_layer_upload(dmk_crowd_t c) {
    if (c->upload_pending) {
        DMK_CUDA(c->ctx, cudaStreamWaitEvent(c->ctx->stream, c->upload_done, 0));
        c->upload_pending = false;
    }
    return DMK_OK;
}
// ... and tells the upload stream when the layer arrays are free again (after the kernels of a step that read or write them)
static inline dmk_status mark_layers_idle(dmk_crowd_t c) {
    if (!c->layers_idle) DMK_CUDA(c->ctx, cudaEventCreateWithFlags(&c->layers_idle, cudaEventDisableTiming));
    DMK_CUDA(c->ctx, cudaEventRecord(c->layers_idle, c->ctx->stream));
    c->layers_idle_recorded = true;
    return DMK_OK;
}
