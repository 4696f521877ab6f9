// api.cu -- implementation of the C ABI declared in include/dmk.h (host side; the kernels are in *_kernel.cu).
//
// No CPU fallback lives here: every compute entry point launches the sm_100a kernels or fails with DMK_ERR_CUDA.
#include "../../include/dmk.h"

#include "animator_flatten.hpp"
#include "device_types.cuh"
#include "host_math.hpp"
#include "kernels.cuh"
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
    int skin_variant = 0;   // DMK_OPT_SKIN_PLANAR_OUTPUT
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
static thread_local std::string g_create_error;

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

struct PackedComponent {
    std::vector<float> ratio;
    std::vector<uint2> value;
    std::vector<int32_t> stream;
};
uint2 pack_key(const Float3Key& k) {
    return make_uint2(static_cast<uint32_t>(k.value[0]) | (static_cast<uint32_t>(k.value[1]) << 16), static_cast<uint32_t>(k.value[2]));
}
uint2 pack_key(const QuaternionKey& k) {
    return make_uint2(static_cast<uint32_t>(static_cast<uint16_t>(k.value[0])) | (static_cast<uint32_t>(static_cast<uint16_t>(k.value[1])) << 16),
                      static_cast<uint32_t>(static_cast<uint16_t>(k.value[2])) | (static_cast<uint32_t>(k.largest) << 16) |
                          (static_cast<uint32_t>(k.sign) << 18));
}
// keys of the ozz stream regrouped per track (stream order within a track = time order)
template <typename Key>
bool regroup(const std::vector<Key>& keys, int padded, PackedComponent& pc, uint2* dir, std::string& err) {
    std::vector<uint32_t> count(padded, 0);
    for (const auto& k : keys) ++count[k.track];
    uint32_t first = static_cast<uint32_t>(pc.ratio.size());
    std::vector<uint32_t> cursor(padded);
    for (int t = 0; t < padded; ++t) {
        if (count[t] < 2) {
            err = "corrupt ozz animation archive (a track has fewer than 2 keys)";
            return false;
        }
        dir[t] = make_uint2(first, count[t]);
        cursor[t] = first;
        first += count[t];
    }
    pc.ratio.resize(first);
    pc.value.resize(first);
    pc.stream.resize(first);
    for (size_t i = 0; i < keys.size(); ++i) {
        const uint32_t at = cursor[keys[i].track]++;
        pc.ratio[at] = keys[i].ratio;
        pc.value[at] = pack_key(keys[i]);
        pc.stream[at] = static_cast<int32_t>(i);
    }
    for (int t = 0; t < padded; ++t)
        for (uint32_t i = dir[t].x + 1; i < dir[t].x + dir[t].y; ++i)
            if (pc.ratio[i] < pc.ratio[i - 1]) {
                err = "corrupt ozz animation archive (track keys are not sorted by ratio)";
                return false;
            }
    return true;
}

// Bucket table of one clip (see ClipSetDev).  Returns false when no G <= kMaxBuckets separates the keys.
constexpr int kMinBuckets = 16, kMaxBuckets = 1024;
struct BucketTable {
    std::vector<uint2> dir;
    std::vector<float4> a;
    std::vector<uint4> b;
    std::vector<uint2> c;
};
bool build_buckets(const PackedComponent* pc, const uint2* dir, int padded, BucketTable& tbl) {
    for (int G = kMinBuckets; G <= kMaxBuckets; G *= 2) {
        bool ok = true;
        const size_t first = tbl.a.size();
        tbl.a.resize(first + (size_t)3 * G * padded);
        tbl.b.resize(tbl.a.size());
        tbl.c.resize(tbl.a.size());
        for (int comp = 0; comp < 3 && ok; ++comp) {
            for (int t = 0; t < padded && ok; ++t) {
                const uint2 d = dir[comp * padded + t];
                const float* r = pc[comp].ratio.data() + d.x;
                const uint2* v = pc[comp].value.data() + d.x;
                const int n = (int)d.y;
                int base = 1;
                for (int b = 0; b < G; ++b) {
                    const float tb = (float)b / (float)G, te = (float)(b + 1) / (float)G;
                    while (base < n - 1 && !(r[base] > tb)) ++base;  // first key (>= 1) with ratio > tb, else the last
                    if (!(base >= n - 2 || r[base + 1] >= te)) { ok = false; break; }
                    const size_t e = first + ((size_t)comp * G + b) * padded + t;
                    const bool last = base == n - 1;
                    tbl.a[e] = make_float4(r[base - 1], r[base], last ? -1.f : r[base + 1], 0.f);
                    std::memcpy(&tbl.a[e].w, &base, sizeof(int));
                    tbl.b[e] = make_uint4(v[base - 1].x, v[base - 1].y, v[base].x, v[base].y);
                    tbl.c[e] = last ? v[base] : v[base + 1];
                }
            }
        }
        if (ok) {
            tbl.dir.push_back(make_uint2((uint32_t)first, (uint32_t)G));
            return true;
        }
        tbl.a.resize(first);
        tbl.b.resize(first);
        tbl.c.resize(first);
    }
    return false;
}
}  // namespace

// ------------------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------------------
extern "C" {

const char* dmk_version(void) { return "darmok_b200 0.1 (sm_100a)"; }

dmk_status dmk_context_create(int device, void* cuda_stream, dmk_context_t* out) {
    if (!out) return fail(nullptr, DMK_ERR_INVALID, "out is null");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, DMK_ERR_CUDA, std::string("no CUDA device (this library has no CPU fallback): ") +
                                               (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    if (device < 0 || device >= count) return fail(nullptr, DMK_ERR_INVALID, "device index out of range");
    if ((e = cudaSetDevice(device)) != cudaSuccess) return cuda_fail(nullptr, e, "cudaSetDevice");
    auto ctx = std::make_unique<dmk_context_s>();
    ctx->device = device;
    cudaDeviceProp prop{};
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return cuda_fail(nullptr, e, "cudaGetDeviceProperties");
    if (prop.major < 10)
        return fail(nullptr, DMK_ERR_CUDA, std::string("device is not sm_100a class (found sm_") + std::to_string(prop.major) +
                                               std::to_string(prop.minor) + "); kernels are built for B200 only");
    ctx->num_sms = prop.multiProcessorCount;
    if (cuda_stream) {
        ctx->stream = static_cast<cudaStream_t>(cuda_stream);
    } else {
        if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess)
            return cuda_fail(nullptr, e, "cudaStreamCreate");
        ctx->owns_stream = true;
    }
    *out = ctx.release();
    return DMK_OK;
}

void dmk_context_destroy(dmk_context_t ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    dev_free(ctx->cull_scratch);
    drain_timers(ctx);
    for (cudaEvent_t e : ctx->event_pool) cudaEventDestroy(e);
    if (ctx->owns_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* dmk_last_error(dmk_context_t ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }
void* dmk_context_stream(dmk_context_t ctx) { return ctx ? ctx->stream : nullptr; }
uint64_t dmk_context_launch_count(dmk_context_t ctx) { return ctx ? ctx->launches : 0; }

dmk_status dmk_context_enable_kernel_timing(dmk_context_t ctx, int on) {
    if (!ctx) return DMK_ERR_INVALID;
    ctx->timing = on != 0;
    return DMK_OK;
}
dmk_status dmk_context_kernel_time(dmk_context_t ctx, int kernel, double* total_ms, uint64_t* launches) {
    if (!ctx || kernel < 0 || kernel >= DMK_KERNEL_COUNT) return DMK_ERR_INVALID;
    if (bind(ctx)) return DMK_ERR_CUDA;
    drain_timers(ctx);
    if (total_ms) *total_ms = ctx->kernel_ms[kernel];
    if (launches) *launches = ctx->kernel_launches[kernel];
    return DMK_OK;
}
dmk_status dmk_context_reset_kernel_timing(dmk_context_t ctx) {
    if (!ctx) return DMK_ERR_INVALID;
    if (bind(ctx)) return DMK_ERR_CUDA;
    drain_timers(ctx);
    for (int k = 0; k < DMK_KERNEL_COUNT; ++k) {
        ctx->kernel_ms[k] = 0;
        ctx->kernel_launches[k] = 0;
    }
    return DMK_OK;
}

dmk_status dmk_sync(dmk_context_t ctx) {
    if (!ctx) return DMK_ERR_INVALID;
    if (bind(ctx)) return DMK_ERR_CUDA;
    DMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DMK_OK;
}

// ------------------------------------------------------------------------------------------------------------
// skeleton
// ------------------------------------------------------------------------------------------------------------
dmk_status dmk_skeleton_load_ozz(dmk_context_t ctx, const void* bytes, size_t size, dmk_skeleton_t* out) {
    if (!ctx || !out) return fail(ctx, DMK_ERR_INVALID, "null argument");
    *out = nullptr;
    Building<dmk_skeleton_s, dmk_skeleton_destroy> skel(new dmk_skeleton_s());
    skel->ctx = ctx;
    std::string err;
    if (!parse_skeleton(bytes, size, skel->data, err)) return fail(ctx, DMK_ERR_FORMAT, err);
    if (bind(ctx)) return DMK_ERR_CUDA;
    const int J = skel->data.num_joints;
    // level-order schedule of LocalToModelJob: a joint's model matrix needs its parent's, so joints are processed by
    // depth; within a depth every (joint, column) pair is an independent task, 32 tasks per round (one per lane)
    std::vector<int> depth(J, 0);
    int max_depth = 0;
    for (int j = 0; j < J; ++j) {
        depth[j] = skel->data.parents[j] < 0 ? 0 : depth[skel->data.parents[j]] + 1;
        max_depth = std::max(max_depth, depth[j]);
    }
    std::vector<uint16_t> sched;
    for (int d = 0; d <= max_depth && J > 0; ++d) {
        size_t in_round = 0;
        for (int j = 0; j < J; ++j) {
            if (depth[j] != d) continue;
            for (int c = 0; c < 4; ++c) {
                sched.push_back(static_cast<uint16_t>(j | (c << 12)));
                ++in_round;
            }
        }
        while (in_round % 32) {
            sched.push_back(kIdleTask);
            ++in_round;
        }
    }
    skel->rounds = static_cast<int>(sched.size() / 32);
    DMK_CUDA(ctx, upload(ctx, &skel->d_parents, skel->data.parents));
    DMK_CUDA(ctx, upload(ctx, &skel->d_schedule, sched));
    {
        const int P = skel->data.num_soa() * 4;
        std::vector<float> rest(static_cast<size_t>(P) * 10);
        for (int j = 0; j < P; ++j)
            for (int k = 0; k < 10; ++k) rest[static_cast<size_t>(j) * 10 + k] = skel->data.rest_soa[static_cast<size_t>(j / 4) * 40 + k * 4 + (j & 3)];
        DMK_CUDA(ctx, upload(ctx, &skel->d_rest, rest));
    }
    *out = skel.release();
    return DMK_OK;
}

void dmk_skeleton_destroy(dmk_skeleton_t skel) {
    if (!skel) return;
    cudaSetDevice(skel->ctx->device);
    dev_free(skel->d_parents);
    dev_free(skel->d_schedule);
    dev_free(skel->d_rest);
    delete skel;
}
int dmk_skeleton_num_joints(dmk_skeleton_t skel) { return skel ? skel->data.num_joints : 0; }
const char* dmk_skeleton_joint_name(dmk_skeleton_t skel, int i) {
    return (skel && i >= 0 && i < skel->data.num_joints) ? skel->data.names[i].c_str() : nullptr;
}
int dmk_skeleton_joint_parent(dmk_skeleton_t skel, int i) {
    return (skel && i >= 0 && i < skel->data.num_joints) ? skel->data.parents[i] : -1;
}
int dmk_skeleton_find_joint(dmk_skeleton_t skel, const char* name) {
    if (!skel || !name) return -1;
    for (int i = 0; i < skel->data.num_joints; ++i)
        if (skel->data.names[i] == name) return i;
    return -1;
}

// ------------------------------------------------------------------------------------------------------------
// clips
// ------------------------------------------------------------------------------------------------------------
dmk_status dmk_clip_load_ozz(dmk_context_t ctx, const void* bytes, size_t size, dmk_clip_t* out) {
    if (!ctx || !out) return fail(ctx, DMK_ERR_INVALID, "null argument");
    *out = nullptr;
    auto clip = std::make_unique<dmk_clip_s>();
    clip->ctx = ctx;
    std::string err;
    if (!parse_animation(bytes, size, clip->data, err)) return fail(ctx, DMK_ERR_FORMAT, err);
    *out = clip.release();
    return DMK_OK;
}
void dmk_clip_destroy(dmk_clip_t clip) { delete clip; }
float dmk_clip_duration(dmk_clip_t clip) { return clip ? clip->data.duration : 0.f; }
const char* dmk_clip_name(dmk_clip_t clip) { return clip ? clip->data.name.c_str() : nullptr; }
int dmk_clip_num_tracks(dmk_clip_t clip) { return clip ? clip->data.num_tracks : 0; }


dmk_status dmk_clipset_create(dmk_context_t ctx, dmk_skeleton_t skel, const dmk_clip_t* clips, int num_clips, dmk_clipset_t* out) {
    if (!ctx || !skel || !clips || !out || num_clips <= 0) return fail(ctx, DMK_ERR_INVALID, "null argument or empty clip list");
    *out = nullptr;
    if (bind(ctx)) return DMK_ERR_CUDA;
    const int padded = skel->data.num_soa() * 4;
    Building<dmk_clipset_s, dmk_clipset_destroy> set(new dmk_clipset_s());
    set->ctx = ctx;
    set->num_clips = num_clips;
    set->padded_tracks = padded;
    std::vector<uint2> dir(static_cast<size_t>(num_clips) * 3 * padded);
    PackedComponent pc[3];
    for (int c = 0; c < num_clips; ++c) {
        if (!clips[c]) return fail(ctx, DMK_ERR_INVALID, "null clip");
        const AnimationData& a = clips[c]->data;
        if (a.num_soa() != skel->data.num_soa())
            return fail(ctx, DMK_ERR_UNSUPPORTED,
                        "animation '" + a.name + "' has " + std::to_string(a.num_tracks) + " tracks, skeleton has " +
                            std::to_string(skel->data.num_joints) + " joints (SoA track counts must match)");
        std::string err;
        uint2* d = dir.data() + static_cast<size_t>(c) * 3 * padded;
        if (!regroup(a.translations, padded, pc[0], d, err) || !regroup(a.rotations, padded, pc[1], d + padded, err) ||
            !regroup(a.scales, padded, pc[2], d + 2 * padded, err))
            return fail(ctx, DMK_ERR_FORMAT, err);
        set->durations.push_back(a.duration);
    }
    // bucket tables need every clip's keys regrouped first: rebuild the per-clip views over the packed arrays
    BucketTable tbl;
    bool buckets = true;
    for (int c = 0; c < num_clips && buckets; ++c)
        buckets = build_buckets(pc, dir.data() + static_cast<size_t>(c) * 3 * padded, padded, tbl);
    if (buckets) {
        DMK_CUDA(ctx, upload(ctx, &set->d_bucket_dir, tbl.dir));
        DMK_CUDA(ctx, upload(ctx, &set->d_bucket_a, tbl.a));
        DMK_CUDA(ctx, upload(ctx, &set->d_bucket_b, tbl.b));
        DMK_CUDA(ctx, upload(ctx, &set->d_bucket_c, tbl.c));
        set->has_buckets = true;
    }
    DMK_CUDA(ctx, upload(ctx, &set->d_dir, dir));
    for (int k = 0; k < 3; ++k) {
        DMK_CUDA(ctx, upload(ctx, &set->d_ratio[k], pc[k].ratio));
        DMK_CUDA(ctx, upload(ctx, &set->d_value[k], pc[k].value));
        DMK_CUDA(ctx, upload(ctx, &set->d_stream[k], pc[k].stream));
    }
    DMK_CUDA(ctx, upload(ctx, &set->d_duration, set->durations));
    *out = set.release();
    return DMK_OK;
}

void dmk_clipset_destroy(dmk_clipset_t set) {
    if (!set) return;
    cudaSetDevice(set->ctx->device);
    dev_free(set->d_dir);
    for (int k = 0; k < 3; ++k) {
        dev_free(set->d_ratio[k]);
        dev_free(set->d_value[k]);
        dev_free(set->d_stream[k]);
    }
    dev_free(set->d_duration);
    dev_free(set->d_bucket_dir);
    dev_free(set->d_bucket_a);
    dev_free(set->d_bucket_b);
    dev_free(set->d_bucket_c);
    delete set;
}

// ------------------------------------------------------------------------------------------------------------
// armature + mesh
// ------------------------------------------------------------------------------------------------------------
dmk_status dmk_armature_create(dmk_context_t ctx, dmk_skeleton_t skel, const char* const* joint_names, const float* inv_bind,
                               int num_joints, dmk_armature_t* out) {
    if (!ctx || !skel || !out || num_joints < 0 || (num_joints > 0 && (!joint_names || !inv_bind)))
        return fail(ctx, DMK_ERR_INVALID, "null argument");
    *out = nullptr;
    if (num_joints + 1 > kMaxPalette)
        return fail(ctx, DMK_ERR_UNSUPPORTED, "armature has more than 255 joints (palette slots are bytes; the reference caps at 64)");
    if (bind(ctx)) return DMK_ERR_CUDA;
    Building<dmk_armature_s, dmk_armature_destroy> arm(new dmk_armature_s());
    arm->ctx = ctx;
    arm->num_joints = num_joints;
    std::vector<float> ib(static_cast<size_t>(num_joints) * 12);
    for (int k = 0; k < num_joints; ++k) {
        const float* m = inv_bind + static_cast<size_t>(k) * 16;
        if (m[3] != 0.f || m[7] != 0.f || m[11] != 0.f || m[15] != 1.f)
            return fail(ctx, DMK_ERR_UNSUPPORTED, "inverse bind pose of joint " + std::to_string(k) + " is not affine (bottom row must be 0,0,0,1)");
        for (int c = 0; c < 4; ++c)
            for (int r = 0; r < 3; ++r) ib[static_cast<size_t>(k) * 12 + c * 3 + r] = m[c * 4 + r];
        arm->remap.push_back(joint_names[k] ? dmk_skeleton_find_joint(skel, joint_names[k]) : -1);
    }
    DMK_CUDA(ctx, upload(ctx, &arm->d_remap, arm->remap));
    DMK_CUDA(ctx, upload(ctx, &arm->d_inv_bind, ib));
    *out = arm.release();
    return DMK_OK;
}
void dmk_armature_destroy(dmk_armature_t arm) {
    if (!arm) return;
    cudaSetDevice(arm->ctx->device);
    dev_free(arm->d_remap);
    dev_free(arm->d_inv_bind);
    delete arm;
}
int dmk_armature_palette_size(dmk_armature_t arm) { return arm ? arm->num_joints + 1 : 0; }

dmk_status dmk_mesh_create(dmk_context_t ctx, const float* position, const float* normal, const float* tangent, const float* indices,
                           const float* weights, int num_vertices, int palette_size, dmk_mesh_t* out) {
    return dmk_mesh_create_ex(ctx, position, normal, tangent, indices, weights, num_vertices, palette_size, 0, out);
}

dmk_status dmk_mesh_vertex_order(dmk_mesh_t mesh, int32_t* out_order) {
    if (!mesh || !out_order) return DMK_ERR_INVALID;
    std::copy(mesh->order.begin(), mesh->order.end(), out_order);
    return DMK_OK;
}

dmk_status dmk_mesh_create_ex(dmk_context_t ctx, const float* position, const float* normal, const float* tangent, const float* indices,
                              const float* weights, int num_vertices, int palette_size, uint32_t flags, dmk_mesh_t* out) {
    if (!ctx || !out || num_vertices < 0 || (num_vertices > 0 && (!position || !normal || !tangent || !indices || !weights)))
        return fail(ctx, DMK_ERR_INVALID, "null argument");
    *out = nullptr;
    if (palette_size < 1 || palette_size > kMaxPalette) return fail(ctx, DMK_ERR_INVALID, "palette_size must be in [1, 256]");
    if (bind(ctx)) return DMK_ERR_CUDA;
    Building<dmk_mesh_s, dmk_mesh_destroy> mesh(new dmk_mesh_s());
    mesh->ctx = ctx;
    mesh->num_vertices = num_vertices;
    mesh->palette_size = palette_size;
    const int vpad = (num_vertices + 7) / 8 * 8;
    mesh->vpad = vpad;
    std::vector<float> pos(static_cast<size_t>(vpad) * 3, 0.f), nrm(pos.size(), 0.f), tan(pos.size(), 0.f);
    std::vector<float> w(static_cast<size_t>(vpad) * 4, 0.f);
    std::vector<uint32_t> slots(vpad, 0u);
    // storage order: the caller's, or grouped by the number of non-zero weights (stable, most influences first)
    mesh->order.resize(static_cast<size_t>(num_vertices));
    for (int v = 0; v < num_vertices; ++v) mesh->order[static_cast<size_t>(v)] = v;
    if (flags & DMK_MESH_SORT_BY_INFLUENCES) {
        auto count = [&](int v) {
            const float* idx = indices + static_cast<size_t>(v) * 4;
            if (!(idx[0] >= 0.f)) return 1;
            int n = 1;
            for (int c = 1; c < 4; ++c) n += weights[static_cast<size_t>(v) * 4 + c] != 0.f ? 1 : 0;
            return n;
        };
        std::stable_sort(mesh->order.begin(), mesh->order.end(), [&](int a, int b) { return count(a) > count(b); });
        // Every thread of the skinning kernel owns 8 consecutive stored vertices and a warp executes the k-th vertex of its
        // 32 threads together.  Fill the storage "column major" over those 8-vertex blocks: sorted vertex i goes to block
        // i % B, slot i / B.  Then slot k of (almost) every block holds vertices with the same influence count -- the
        // gathers of an absent influence are skipped by whole warps -- and every thread, warp and tile gets the same mix.
        // (4-vertex blocks for the 4-vertices-per-thread shapes, DMK_MESH_SORT_BLOCKS_OF_4.)
        const size_t per = (flags & DMK_MESH_SORT_BLOCKS_OF_4) ? 4 : 8;
        const size_t V = mesh->order.size(), B = static_cast<size_t>(vpad) / per;
        std::vector<int32_t> dealt(V);
        size_t next = 0;
        for (size_t slot = 0; slot < per; ++slot)
            for (size_t blk = 0; blk < B; ++blk) {
                const size_t pos = blk * per + slot;
                if (pos < V) dealt[pos] = mesh->order[next++];
            }
        mesh->order = dealt;
        mesh->sorted = true;
        mesh->sort_block = static_cast<int>(per);
    }
    for (int s = 0; s < vpad; ++s) {
        float* wv = &w[static_cast<size_t>(s) * 4];
        wv[0] = 1.f;  // padding and unskinned vertices: 1 * identity
        if (s >= num_vertices) continue;
        const int v = mesh->order[static_cast<size_t>(s)];
        for (int k = 0; k < 3; ++k) {
            pos[static_cast<size_t>(s) * 3 + k] = position[static_cast<size_t>(v) * 3 + k];
            nrm[static_cast<size_t>(s) * 3 + k] = normal[static_cast<size_t>(v) * 3 + k];
            tan[static_cast<size_t>(s) * 3 + k] = tangent[static_cast<size_t>(v) * 3 + k];
        }
        const float* idx = indices + static_cast<size_t>(v) * 4;
        if (!(idx[0] >= 0.f)) continue;  // applySkinning: `if (indices.x >= 0)` else the vertex passes through
        uint32_t packed = 0;
        for (int c = 0; c < 4; ++c) {
            const int slot = static_cast<int>(idx[c] + 1.f);  // getSkinningMatrixPart: u_skinning[int(v + 1)]
            if (!(idx[c] >= -1.f) || slot < 0 || slot >= palette_size)
                return fail(ctx, DMK_ERR_INVALID, "vertex " + std::to_string(v) + " references bone " + std::to_string(idx[c]) +
                                                      " outside the palette (size " + std::to_string(palette_size) + ")");
            packed |= static_cast<uint32_t>(slot) << (8 * c);
            wv[c] = weights[static_cast<size_t>(v) * 4 + c];
        }
        slots[s] = packed;
    }
    DMK_CUDA(ctx, upload(ctx, &mesh->d_pos, pos));
    DMK_CUDA(ctx, upload(ctx, &mesh->d_nrm, nrm));
    DMK_CUDA(ctx, upload(ctx, &mesh->d_tan, tan));
    DMK_CUDA(ctx, upload(ctx, &mesh->d_weights, w));
    DMK_CUDA(ctx, upload(ctx, &mesh->d_slots, slots));
    *out = mesh.release();
    return DMK_OK;
}
void dmk_mesh_destroy(dmk_mesh_t mesh) {
    if (!mesh) return;
    cudaSetDevice(mesh->ctx->device);
    dev_free(mesh->d_pos);
    dev_free(mesh->d_nrm);
    dev_free(mesh->d_tan);
    dev_free(mesh->d_weights);
    dev_free(mesh->d_slots);
    delete mesh;
}
int dmk_mesh_num_vertices(dmk_mesh_t mesh) { return mesh ? mesh->num_vertices : 0; }
int dmk_mesh_vertex_pitch(dmk_mesh_t mesh) { return mesh ? mesh->vpad : 0; }

// ------------------------------------------------------------------------------------------------------------
// crowd
// ------------------------------------------------------------------------------------------------------------
dmk_status dmk_crowd_create(dmk_context_t ctx, dmk_skeleton_t skel, dmk_clipset_t clips, dmk_armature_t arm, dmk_mesh_t mesh,
                            int64_t n, int max_layers, dmk_crowd_t* out) {
    if (!ctx || !skel || !clips || !out) return fail(ctx, DMK_ERR_INVALID, "null argument");
    *out = nullptr;
    if (n < 0 || n > 0x7fffffff) return fail(ctx, DMK_ERR_INVALID, "num_characters out of range");
    if (max_layers < 1 || max_layers > kMaxLayers) return fail(ctx, DMK_ERR_INVALID, "max_layers must be in [1, 32]");
    if (mesh && !arm) return fail(ctx, DMK_ERR_INVALID, "a mesh needs an armature");
    if (mesh && mesh->palette_size != arm->num_joints + 1)
        return fail(ctx, DMK_ERR_INVALID, "mesh palette_size does not match the armature");
    if (skel->data.num_joints == 0) return fail(ctx, DMK_ERR_INVALID, "skeleton has no joints");
    if (bind(ctx)) return DMK_ERR_CUDA;
    Building<dmk_crowd_s, dmk_crowd_destroy> crowd(new dmk_crowd_s());
    dmk_crowd_s* c = crowd.get();
    c->ctx = ctx;
    c->skel = skel;
    c->clips = clips;
    c->arm = arm;
    c->mesh = mesh;
    c->n = n;
    c->max_layers = max_layers;
    const size_t ln = static_cast<size_t>(max_layers) * static_cast<size_t>(n);
    DMK_CUDA(ctx, dev_alloc(&c->d_clip, ln));
    DMK_CUDA(ctx, dev_alloc(&c->d_ratio, ln));
    DMK_CUDA(ctx, dev_alloc(&c->d_weight, ln));
    DMK_CUDA(ctx, dev_alloc(&c->d_speed, ln));
    DMK_CUDA(ctx, dev_alloc(&c->d_loop, ln));
    DMK_CUDA(ctx, dev_alloc(&c->d_looped, ln));
    DMK_CUDA(ctx, dev_alloc(&c->d_local_matrices, static_cast<size_t>(n) * skel->data.num_soa() * 4 * 16));
    DMK_CUDA(ctx, dev_alloc(&c->d_error, 1));
    DMK_CUDA(ctx, cudaMemsetAsync(c->d_error, 0, sizeof(int32_t), ctx->stream));
    DMK_CUDA(ctx, cudaHostAlloc(reinterpret_cast<void**>(&c->h_error), sizeof(int32_t), cudaHostAllocDefault));
    *c->h_error = 0;
    if (arm) {
        const size_t ps = static_cast<size_t>(arm->num_joints + 1);
        DMK_CUDA(ctx, dev_alloc(&c->d_palette, static_cast<size_t>(n) * ps * 16));
        DMK_CUDA(ctx, dev_alloc(&c->d_palette34, static_cast<size_t>(n) * ps * 16));
    }
    if (mesh) {
        DMK_CUDA(ctx, dev_alloc(&c->d_skinned, static_cast<size_t>(n) * static_cast<size_t>(mesh->vpad) * 9));
        c->plan = plan_skin(mesh->vpad, mesh->palette_size, ctx->num_sms);
    }
    const size_t words = static_cast<size_t>((n + 31) / 32);
    DMK_CUDA(ctx, dev_alloc(&c->d_vis_bits, words + 1));
    DMK_CUDA(ctx, dev_alloc(&c->d_visible_ids, static_cast<size_t>(n) + 1));
    DMK_CUDA(ctx, dev_alloc(&c->d_visible_count, 1));
    DMK_CUDA(ctx, dev_alloc(&c->d_scratch, compact_scratch_ints(n)));
    DMK_CUDA(ctx, cudaMemsetAsync(c->d_visible_count, 0, sizeof(int32_t), ctx->stream));
    *out = crowd.release();
    return DMK_OK;
}

void dmk_crowd_destroy(dmk_crowd_t c) {
    if (!c) return;
    cudaSetDevice(c->ctx->device);
    cudaStreamSynchronize(c->ctx->stream);
    for (void* p : {(void*)c->d_clip, (void*)c->d_ratio, (void*)c->d_weight, (void*)c->d_speed, (void*)c->d_loop, (void*)c->d_looped,
                    (void*)c->d_program, (void*)c->d_local_matrices, (void*)c->d_palette, (void*)c->d_palette34, (void*)c->d_skinned, (void*)c->d_models, (void*)c->d_locals,
                    (void*)c->d_key_indices, (void*)c->d_error, (void*)c->d_world_inverse, (void*)c->d_bbox, (void*)c->d_position, (void*)c->d_rotation, (void*)c->d_scale, (void*)c->d_vis_bits,
                    (void*)c->d_visible_ids, (void*)c->d_visible_count, (void*)c->d_scratch, (void*)c->d_anim_states, (void*)c->d_anim_clips,
                    (void*)c->d_anim_transitions, (void*)c->d_anim_tables, (void*)c->d_anim_slots, (void*)c->d_anim_transition,
                    (void*)c->d_anim_transition_time, (void*)c->d_anim_events, (void*)c->d_anim_info, (void*)c->d_node_position, (void*)c->d_node_rotation,
                    (void*)c->d_node_scale, (void*)c->d_node_world, (void*)c->d_node_world_inverse, (void*)c->d_node_parent, (void*)c->d_entity_parent, (void*)c->d_anim_cmd, (void*)c->d_anim_status,
                    (void*)c->d_anim_cmd_speed, (void*)c->d_anim_blend, (void*)c->d_anim_speed})
        dev_free(p);
    for (auto& x : c->extra_skins) {
        dev_free(x.d_palette);
        dev_free(x.d_palette34);
        dev_free(x.d_skinned);
    }
    for (int b = 0; b < 2; ++b) {
        if (c->h_vis[b]) cudaFreeHost(c->h_vis[b]);
        if (c->vis_ready[b]) cudaEventDestroy(c->vis_ready[b]);
        if (c->h_stage[b]) cudaFreeHost(c->h_stage[b]);
        if (c->stage_done[b]) cudaEventDestroy(c->stage_done[b]);
    }
    if (c->h_error) cudaFreeHost(c->h_error);
    delete c;
}
int64_t dmk_crowd_size(dmk_crowd_t c) { return c ? c->n : 0; }
dmk_status dmk_crowd_set_option(dmk_crowd_t c, int option, int value) {
    if (!c) return DMK_ERR_INVALID;
    if (option == DMK_OPT_FORCE_KEY_SEARCH) {
        c->force_search = value != 0;
        return DMK_OK;
    }
    if (option == DMK_OPT_SKIN_RESERVED_SMS) {
        if (value < 0 || value >= c->ctx->num_sms) return fail(c->ctx, DMK_ERR_INVALID, "reserved SM count out of range");
        c->reserved_sms = value;
        if (c->mesh) c->plan = plan_skin(c->mesh->vpad, c->mesh->palette_size, c->ctx->num_sms - value, c->skin_variant);
        for (auto& x : c->extra_skins)
            if (x.mesh) x.plan = plan_skin(x.mesh->vpad, x.mesh->palette_size, c->ctx->num_sms - value, c->skin_variant);
        return DMK_OK;
    }
    if (option == DMK_OPT_SKIN_PLANAR_OUTPUT) {
        if (value < 0 || value > 3) return fail(c->ctx, DMK_ERR_INVALID, "DMK_OPT_SKIN_PLANAR_OUTPUT takes 0, 1, 2 or 3");
        if (value) {
            auto supported = [&](dmk_mesh_t m) { return !m || plan_skin(m->vpad, m->palette_size, c->ctx->num_sms, value).replicas == 8; };
            bool ok = supported(c->mesh);
            for (auto& x : c->extra_skins) ok = ok && supported(x.mesh);
            if (!ok) return fail(c->ctx, DMK_ERR_UNSUPPORTED, "planar skinned output needs a palette of at most 181 slots");
        }
        c->skin_variant = value;
        if (c->mesh) c->plan = plan_skin(c->mesh->vpad, c->mesh->palette_size, c->ctx->num_sms - c->reserved_sms, value);
        for (auto& x : c->extra_skins)
            if (x.mesh) x.plan = plan_skin(x.mesh->vpad, x.mesh->palette_size, c->ctx->num_sms - c->reserved_sms, value);
        return DMK_OK;
    }
    if (option == DMK_OPT_KEEP_WORLD_INVERSE) {
        if (bind(c->ctx)) return DMK_ERR_CUDA;
        if (value && !c->d_world_inverse) DMK_CUDA(c->ctx, dev_alloc(&c->d_world_inverse, static_cast<size_t>(c->n) * 16));
        c->keep_world_inverse = value != 0;
        return DMK_OK;
    }
    return fail(c->ctx, DMK_ERR_INVALID, "unknown option");
}

dmk_status dmk_skin_plan(int vertex_pitch, int palette_size, int num_sms, int layout, int64_t* out7) {
    if (!out7 || vertex_pitch < 8 || vertex_pitch % 8 || palette_size < 1 || palette_size > kMaxPalette || num_sms < 1 || layout < 0 || layout > 3)
        return fail(nullptr, DMK_ERR_INVALID, "dmk_skin_plan: argument out of range");
    const SkinPlan plan = plan_skin(vertex_pitch, palette_size, num_sms, layout);
    cudaGetLastError();   // without a device the occupancy query fails and the plan assumes one block per SM
    out7[0] = plan.threads;
    out7[1] = plan.num_tiles;
    out7[2] = plan.tile_vertices;
    out7[3] = plan.grid;
    out7[4] = plan.replicas;
    out7[5] = static_cast<int64_t>(plan.smem);
    out7[6] = (layout == 1 || layout == 2) ? 4 : 8;
    return DMK_OK;
}

dmk_status dmk_crowd_enable_outputs(dmk_crowd_t c, uint32_t outputs) {
    if (!c) return DMK_ERR_INVALID;
    dmk_context_t ctx = c->ctx;
    if (bind(ctx)) return DMK_ERR_CUDA;
    const size_t n = static_cast<size_t>(c->n);
    if ((outputs & DMK_OUT_MODELS) && !c->d_models) DMK_CUDA(ctx, dev_alloc(&c->d_models, n * c->skel->data.num_joints * 16));
    if ((outputs & DMK_OUT_LOCALS) && !c->d_locals) DMK_CUDA(ctx, dev_alloc(&c->d_locals, n * c->skel->data.num_soa() * kSoaFloats));
    if ((outputs & DMK_OUT_KEY_INDICES) && !c->d_key_indices)
        DMK_CUDA(ctx, dev_alloc(&c->d_key_indices, n * c->max_layers * 3 * c->skel->data.num_soa() * 4 * 2));
    return DMK_OK;
}

// Returns a pinned staging buffer of at least `bytes` that no in-flight copy is reading; stage_release() must be
// called after the copies out of it have been enqueued.
static dmk_status stage_acquire(dmk_crowd_t c, size_t bytes, char** out) {
    const int b = c->stage_next;
    if (c->stage_done[b]) DMK_CUDA(c->ctx, cudaEventSynchronize(c->stage_done[b]));
    else DMK_CUDA(c->ctx, cudaEventCreateWithFlags(&c->stage_done[b], cudaEventDisableTiming));
    if (c->h_stage_bytes[b] < bytes) {
        if (c->h_stage[b]) cudaFreeHost(c->h_stage[b]);
        c->h_stage[b] = nullptr;
        c->h_stage_bytes[b] = 0;
        DMK_CUDA(c->ctx, cudaHostAlloc(&c->h_stage[b], bytes, cudaHostAllocDefault));
        c->h_stage_bytes[b] = bytes;
    }
    *out = static_cast<char*>(c->h_stage[b]);
    return DMK_OK;
}
static dmk_status stage_release(dmk_crowd_t c) {
    DMK_CUDA(c->ctx, cudaEventRecord(c->stage_done[c->stage_next], c->ctx->stream));
    c->stage_next ^= 1;
    return DMK_OK;
}

dmk_status dmk_crowd_set_layers(dmk_crowd_t c, int num_layers, const int32_t* clip, const float* ratio, const float* weight,
                                const float* speed, const uint8_t* loop, float threshold) {
    if (!c) return DMK_ERR_INVALID;
    dmk_context_t ctx = c->ctx;
    if (num_layers < 1 || num_layers > c->max_layers) return fail(ctx, DMK_ERR_INVALID, "num_layers out of range");
    if (!clip || !ratio || !weight) return fail(ctx, DMK_ERR_INVALID, "null layer array");
    if (bind(ctx)) return DMK_ERR_CUDA;
    const size_t ln = static_cast<size_t>(num_layers) * static_cast<size_t>(c->n);
    for (size_t i = 0; i < ln; ++i)
        if (clip[i] < 0 || clip[i] >= c->clips->num_clips)
            return fail(ctx, DMK_ERR_INVALID, "layer references clip " + std::to_string(clip[i]) + " outside the clip set");
    cudaStream_t s = ctx->stream;
    DMK_CUDA(ctx, cudaMemcpyAsync(c->d_clip, clip, ln * 4, cudaMemcpyHostToDevice, s));
    DMK_CUDA(ctx, cudaMemcpyAsync(c->d_ratio, ratio, ln * 4, cudaMemcpyHostToDevice, s));
    DMK_CUDA(ctx, cudaMemcpyAsync(c->d_weight, weight, ln * 4, cudaMemcpyHostToDevice, s));
    if (speed) {
        DMK_CUDA(ctx, cudaMemcpyAsync(c->d_speed, speed, ln * 4, cudaMemcpyHostToDevice, s));
    } else {
        std::vector<float> ones(ln, 1.f);
        DMK_CUDA(ctx, cudaMemcpyAsync(c->d_speed, ones.data(), ln * 4, cudaMemcpyHostToDevice, s));
        DMK_CUDA(ctx, cudaStreamSynchronize(s));
    }
    c->has_loop = loop != nullptr;
    if (loop) DMK_CUDA(ctx, cudaMemcpyAsync(c->d_loop, loop, ln, cudaMemcpyHostToDevice, s));
    DMK_CUDA(ctx, cudaMemsetAsync(c->d_looped, 0, ln, s));
    DMK_CUDA(ctx, cudaStreamSynchronize(s));  // host arrays are the caller's
    c->num_layers = num_layers;
    c->threshold = threshold;
    c->layers_set = true;
    c->use_program = false;
    return DMK_OK;
}

dmk_status dmk_crowd_set_programs(dmk_crowd_t c, int num_layers, const int32_t* clip, const float* ratio, const float* weight,
                                  const dmk_pose_program* programs) {
    if (!c) return DMK_ERR_INVALID;
    dmk_context_t ctx = c->ctx;
    static_assert(sizeof(dmk_pose_program) == sizeof(PoseProgram), "C ABI struct and device struct must match");
    if (num_layers < 1 || num_layers > c->max_layers) return fail(ctx, DMK_ERR_INVALID, "num_layers out of range");
    if (!clip || !ratio || !weight || !programs) return fail(ctx, DMK_ERR_INVALID, "null argument");
    if (bind(ctx)) return DMK_ERR_CUDA;
    const size_t n = static_cast<size_t>(c->n);
    for (size_t i = 0; i < n; ++i) {
        const dmk_pose_program& g = programs[i];
        if (g.top > DMK_TOP_SKIP || g.mode0 > DMK_GROUP_ZERO || g.mode1 > DMK_GROUP_ZERO || g.num_layers0 + g.num_layers1 > num_layers)
            return fail(ctx, DMK_ERR_INVALID, "malformed pose program for character " + std::to_string(i));
        if (g.top == DMK_TOP_SKIP || g.top == DMK_TOP_REST_POSE) continue;
        const int groups = g.top == DMK_TOP_TRANSITION ? 2 : 1;
        for (int k = 0; k < groups; ++k) {
            const int cnt = k ? g.num_layers1 : g.num_layers0, rest = k ? g.rest1 : g.rest0, mode = k ? g.mode1 : g.mode0;
            const float th = k ? g.threshold1 : g.threshold0;
            if (cnt < 1 || rest >= cnt) return fail(ctx, DMK_ERR_INVALID, "malformed pose program for character " + std::to_string(i));
            // BlendingJob::Validate: threshold > 0 (darmok reports "error in the blending job", src/skeleton_ozz.cpp:605-608)
            if (mode == DMK_GROUP_BLEND && !(th > 0.f)) return fail(ctx, DMK_ERR_JOB, "error in the blending job");
        }
        for (int l = 0; l < g.num_layers0 + g.num_layers1; ++l)
            if (clip[static_cast<size_t>(l) * n + i] < 0 || clip[static_cast<size_t>(l) * n + i] >= c->clips->num_clips)
                return fail(ctx, DMK_ERR_INVALID, "layer references a clip outside the clip set");
    }
    if (!c->d_program) DMK_CUDA(ctx, dev_alloc(&c->d_program, n));
    const size_t lbytes = static_cast<size_t>(num_layers) * n * 4, pbytes = n * sizeof(PoseProgram);
    char* h = nullptr;
    if (dmk_status st = stage_acquire(c, 3 * lbytes + pbytes, &h)) return st;
    std::memcpy(h, clip, lbytes);
    std::memcpy(h + lbytes, ratio, lbytes);
    std::memcpy(h + 2 * lbytes, weight, lbytes);
    std::memcpy(h + 3 * lbytes, programs, pbytes);
    cudaStream_t s = ctx->stream;
    DMK_CUDA(ctx, cudaMemcpyAsync(c->d_clip, h, lbytes, cudaMemcpyHostToDevice, s));
    DMK_CUDA(ctx, cudaMemcpyAsync(c->d_ratio, h + lbytes, lbytes, cudaMemcpyHostToDevice, s));
    DMK_CUDA(ctx, cudaMemcpyAsync(c->d_weight, h + 2 * lbytes, lbytes, cudaMemcpyHostToDevice, s));
    DMK_CUDA(ctx, cudaMemcpyAsync(c->d_program, h + 3 * lbytes, pbytes, cudaMemcpyHostToDevice, s));
    if (dmk_status st = stage_release(c)) return st;
    c->num_layers = num_layers;
    c->layers_set = true;
    c->use_program = true;
    c->has_loop = false;
    return DMK_OK;
}

dmk_status dmk_crowd_update_layers(dmk_crowd_t c, int num_layers, const int32_t* clip, const float* ratio, const float* weight) {
    if (!c) return DMK_ERR_INVALID;
    dmk_context_t ctx = c->ctx;
    if (!c->layers_set || num_layers != c->num_layers) return fail(ctx, DMK_ERR_INVALID, "call dmk_crowd_set_layers first (same num_layers)");
    if (!clip || !ratio || !weight) return fail(ctx, DMK_ERR_INVALID, "null layer array");
    if (bind(ctx)) return DMK_ERR_CUDA;
    const size_t bytes = static_cast<size_t>(num_layers) * static_cast<size_t>(c->n) * 4;
    char* h = nullptr;
    if (dmk_status st = stage_acquire(c, 3 * bytes, &h)) return st;
    std::memcpy(h, clip, bytes);
    std::memcpy(h + bytes, ratio, bytes);
    std::memcpy(h + 2 * bytes, weight, bytes);
    DMK_CUDA(ctx, cudaMemcpyAsync(c->d_clip, h, bytes, cudaMemcpyHostToDevice, ctx->stream));
    DMK_CUDA(ctx, cudaMemcpyAsync(c->d_ratio, h + bytes, bytes, cudaMemcpyHostToDevice, ctx->stream));
    DMK_CUDA(ctx, cudaMemcpyAsync(c->d_weight, h + 2 * bytes, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return stage_release(c);
}

// ---- device-side animators ------------------------------------------------------------
// (the host-side flattening of the definition is in animator_flatten.hpp)
static AnimatorParams animator_params(dmk_crowd_t c, float dt) {
    AnimatorParams a{};
    a.states = c->d_anim_states;
    a.clips = c->d_anim_clips;
    a.transitions = c->d_anim_transitions;
    a.num_states = c->anim_num_states;
    a.num_transitions = c->anim_num_transitions;
    a.tables = c->d_anim_tables;
    a.slots = c->d_anim_slots;
    a.transition_def = c->d_anim_transition;
    a.transition_time = c->d_anim_transition_time;
    a.cmd_state = c->d_anim_cmd;
    a.cmd_speed = c->d_anim_cmd_speed;
    a.in_blend = c->d_anim_blend;
    a.in_speed = c->d_anim_speed;
    a.events = c->d_anim_events;
    a.status = c->d_anim_status;
    a.n = c->n;
    a.num_layers = c->anim_layers;
    a.dt = dt;
    a.clip = c->d_clip;
    a.ratio = c->d_ratio;
    a.weight = c->d_weight;
    a.program = c->d_program;
    return a;
}

dmk_status dmk_crowd_set_animator(dmk_crowd_t c, const dmk_anim_state_def* states, int num_states, const dmk_anim_clip_def* clips,
                                  int num_clips, const dmk_anim_transition_def* transitions, int num_transitions) {
    if (!c) return DMK_ERR_INVALID;
    dmk_context_t ctx = c->ctx;
    if (!states || !clips || num_states < 1 || num_states > 32767 || num_clips < 1 || num_transitions < 0 || (num_transitions && !transitions))
        return fail(ctx, DMK_ERR_INVALID, "malformed animator definition");
    if (bind(ctx)) return DMK_ERR_CUDA;
    FlatAnimator flat;
    if (const std::string msg = flatten_animator(states, num_states, clips, num_clips, transitions, num_transitions, c->clips->durations, flat); !msg.empty())
        return fail(ctx, DMK_ERR_INVALID, msg);
    const int max_clips = flat.max_clips;
    if (2 * max_clips > c->max_layers)
        return fail(ctx, DMK_ERR_INVALID, "the crowd needs max_layers >= " + std::to_string(2 * max_clips) + " for this animator (two states blend in a transition)");
    const std::vector<AnimStateDef>& hs = flat.states;
    const std::vector<AnimClipDef>& hc = flat.clips;
    const std::vector<AnimTransitionDef>& ht = flat.transitions;
    const std::vector<float>& tables = flat.tables;
    for (void* p : {(void*)c->d_anim_states, (void*)c->d_anim_clips, (void*)c->d_anim_transitions, (void*)c->d_anim_tables}) dev_free(p);
    c->d_anim_states = nullptr; c->d_anim_clips = nullptr; c->d_anim_transitions = nullptr; c->d_anim_tables = nullptr;
    DMK_CUDA(ctx, upload(ctx, &c->d_anim_tables, tables));
    DMK_CUDA(ctx, upload(ctx, &c->d_anim_states, hs));
    DMK_CUDA(ctx, upload(ctx, &c->d_anim_clips, hc));
    DMK_CUDA(ctx, upload(ctx, &c->d_anim_transitions, ht));
    const size_t n = static_cast<size_t>(c->n);
    if (!c->d_anim_slots) {
        DMK_CUDA(ctx, dev_alloc(&c->d_anim_slots, static_cast<size_t>(kNumSlots) * kSlotWords * n));
        DMK_CUDA(ctx, dev_alloc(&c->d_anim_transition, n));
        DMK_CUDA(ctx, dev_alloc(&c->d_anim_transition_time, n));
        DMK_CUDA(ctx, dev_alloc(&c->d_anim_events, n));
        DMK_CUDA(ctx, dev_alloc(&c->d_anim_cmd, n));
        DMK_CUDA(ctx, dev_alloc(&c->d_anim_status, n));
        DMK_CUDA(ctx, dev_alloc(&c->d_anim_cmd_speed, n));
        DMK_CUDA(ctx, dev_alloc(&c->d_anim_blend, 2 * n));
        DMK_CUDA(ctx, dev_alloc(&c->d_anim_speed, n));
    }
    if (!c->d_program) DMK_CUDA(ctx, dev_alloc(&c->d_program, n));
    DMK_CUDA(ctx, launch_animator_init(animator_params(c, 0.f), c->d_anim_cmd_speed, c->d_anim_blend, c->d_anim_speed, ctx->stream));
    DMK_CUDA(ctx, cudaMemsetAsync(c->d_anim_status, 0, n * sizeof(int32_t), ctx->stream));
    DMK_CUDA(ctx, cudaMemsetAsync(c->d_anim_events, 0xff, n * sizeof(AnimEvents), ctx->stream));
    DMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // the upload sources are locals
    ++ctx->launches;
    c->anim_num_states = num_states;
    c->anim_num_transitions = num_transitions;
    c->anim_layers = 2 * max_clips;
    c->animator_set = true;
    c->animator_ticked = false;
    return DMK_OK;
}

dmk_status dmk_crowd_animator_commands(dmk_crowd_t c, const int32_t* state, const float* speed) {
    if (!c) return DMK_ERR_INVALID;
    dmk_context_t ctx = c->ctx;
    if (!c->animator_set) return fail(ctx, DMK_ERR_INVALID, "dmk_crowd_set_animator has not been called");
    if (!state) return fail(ctx, DMK_ERR_INVALID, "null command array");
    if (bind(ctx)) return DMK_ERR_CUDA;
    const size_t n = static_cast<size_t>(c->n);
    // an unknown state name is "nothing happens" in the reference (play() returns false): ids outside the table do the same
    char* h = nullptr;
    if (dmk_status st = stage_acquire(c, 2 * n * 4, &h)) return st;
    std::memcpy(h, state, n * 4);
    float* hs = reinterpret_cast<float*>(h + n * 4);
    if (speed) std::memcpy(hs, speed, n * 4);
    else std::fill(hs, hs + n, 1.f);
    DMK_CUDA(ctx, cudaMemcpyAsync(c->d_anim_cmd, h, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    DMK_CUDA(ctx, cudaMemcpyAsync(c->d_anim_cmd_speed, hs, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    return stage_release(c);
}

dmk_status dmk_crowd_animator_set_inputs(dmk_crowd_t c, const float* blend_xy, const float* speed) {
    if (!c) return DMK_ERR_INVALID;
    dmk_context_t ctx = c->ctx;
    if (!c->animator_set) return fail(ctx, DMK_ERR_INVALID, "dmk_crowd_set_animator has not been called");
    if (!blend_xy && !speed) return DMK_OK;
    if (bind(ctx)) return DMK_ERR_CUDA;
    const size_t n = static_cast<size_t>(c->n);
    char* h = nullptr;
    if (dmk_status st = stage_acquire(c, 3 * n * 4, &h)) return st;
    if (blend_xy) {
        std::memcpy(h, blend_xy, 2 * n * 4);
        DMK_CUDA(ctx, cudaMemcpyAsync(c->d_anim_blend, h, 2 * n * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    if (speed) {
        std::memcpy(h + 2 * n * 4, speed, n * 4);
        DMK_CUDA(ctx, cudaMemcpyAsync(c->d_anim_speed, h + 2 * n * 4, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    return stage_release(c);
}

dmk_status dmk_crowd_get_animator_events(dmk_crowd_t c, dmk_anim_events* out_events, int32_t* out_status) {
    if (!c) return DMK_ERR_INVALID;
    dmk_context_t ctx = c->ctx;
    static_assert(sizeof(dmk_anim_events) == sizeof(AnimEvents), "C ABI struct and device struct must match");
    if (!c->animator_set) return fail(ctx, DMK_ERR_INVALID, "dmk_crowd_set_animator has not been called");
    if (bind(ctx)) return DMK_ERR_CUDA;
    const size_t n = static_cast<size_t>(c->n);
    if (out_events) DMK_CUDA(ctx, cudaMemcpyAsync(out_events, c->d_anim_events, n * sizeof(AnimEvents), cudaMemcpyDeviceToHost, ctx->stream));
    if (out_status) DMK_CUDA(ctx, cudaMemcpyAsync(out_status, c->d_anim_status, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    DMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DMK_OK;
}

dmk_status dmk_crowd_get_animator_state(dmk_crowd_t c, dmk_anim_state* out_states) {
    if (!c) return DMK_ERR_INVALID;
    dmk_context_t ctx = c->ctx;
    static_assert(sizeof(dmk_anim_state) == sizeof(AnimStateInfo), "C ABI struct and device struct must match");
    if (!c->animator_set) return fail(ctx, DMK_ERR_INVALID, "dmk_crowd_set_animator has not been called");
    if (!out_states) return fail(ctx, DMK_ERR_INVALID, "null output");
    if (bind(ctx)) return DMK_ERR_CUDA;
    const size_t n = static_cast<size_t>(c->n);
    if (!c->d_anim_info) DMK_CUDA(ctx, dev_alloc(&c->d_anim_info, n));
    DMK_CUDA(ctx, launch_animator_query(animator_params(c, 0.f), c->d_anim_info, ctx->stream));
    ++ctx->launches;
    DMK_CUDA(ctx, cudaMemcpyAsync(out_states, c->d_anim_info, n * sizeof(AnimStateInfo), cudaMemcpyDeviceToHost, ctx->stream));
    DMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DMK_OK;
}

dmk_status dmk_crowd_get_programs(dmk_crowd_t c, int num_layers, int32_t* out_clip, float* out_ratio, float* out_weight,
                                  dmk_pose_program* out_programs) {
    if (!c) return DMK_ERR_INVALID;
    dmk_context_t ctx = c->ctx;
    if (num_layers < 0 || num_layers > c->max_layers) return fail(ctx, DMK_ERR_INVALID, "num_layers out of range");
    if (out_programs && !c->d_program) return fail(ctx, DMK_ERR_INVALID, "the crowd has no pose programs");
    if (bind(ctx)) return DMK_ERR_CUDA;
    const size_t n = static_cast<size_t>(c->n), ln = static_cast<size_t>(num_layers) * n;
    if (out_clip) DMK_CUDA(ctx, cudaMemcpyAsync(out_clip, c->d_clip, ln * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_ratio) DMK_CUDA(ctx, cudaMemcpyAsync(out_ratio, c->d_ratio, ln * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_weight) DMK_CUDA(ctx, cudaMemcpyAsync(out_weight, c->d_weight, ln * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_programs) DMK_CUDA(ctx, cudaMemcpyAsync(out_programs, c->d_program, n * sizeof(PoseProgram), cudaMemcpyDeviceToHost, ctx->stream));
    DMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DMK_OK;
}

dmk_status dmk_crowd_set_instances(dmk_crowd_t c, const float* world_inverse, const float* bbox) {
    if (!c) return DMK_ERR_INVALID;
    dmk_context_t ctx = c->ctx;
    if (!world_inverse || !bbox) return fail(ctx, DMK_ERR_INVALID, "null instance array");
    if (bind(ctx)) return DMK_ERR_CUDA;
    const size_t n = static_cast<size_t>(c->n);
    if (!c->d_world_inverse) DMK_CUDA(ctx, dev_alloc(&c->d_world_inverse, n * 16));
    if (!c->d_bbox) DMK_CUDA(ctx, dev_alloc(&c->d_bbox, n * 6));
    DMK_CUDA(ctx, cudaMemcpyAsync(c->d_world_inverse, world_inverse, n * 64, cudaMemcpyHostToDevice, ctx->stream));
    DMK_CUDA(ctx, cudaMemcpyAsync(c->d_bbox, bbox, n * 24, cudaMemcpyHostToDevice, ctx->stream));
    DMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    c->instances_set = true;
    c->use_trs = false;
    return DMK_OK;
}

dmk_status dmk_crowd_set_transforms(dmk_crowd_t c, const float* position, const float* rotation, const float* scale, const float* bbox) {
    if (!c) return DMK_ERR_INVALID;
    dmk_context_t ctx = c->ctx;
    if (!position || !rotation || !scale || !bbox) return fail(ctx, DMK_ERR_INVALID, "null transform array");
    if (bind(ctx)) return DMK_ERR_CUDA;
    const size_t n = static_cast<size_t>(c->n);
    if (!c->d_position) DMK_CUDA(ctx, dev_alloc(&c->d_position, n * 3));
    if (!c->d_rotation) DMK_CUDA(ctx, dev_alloc(&c->d_rotation, n * 4));
    if (!c->d_scale) DMK_CUDA(ctx, dev_alloc(&c->d_scale, n * 3));
    if (!c->d_bbox) DMK_CUDA(ctx, dev_alloc(&c->d_bbox, n * 6));
    DMK_CUDA(ctx, cudaMemcpyAsync(c->d_position, position, n * 12, cudaMemcpyHostToDevice, ctx->stream));
    DMK_CUDA(ctx, cudaMemcpyAsync(c->d_rotation, rotation, n * 16, cudaMemcpyHostToDevice, ctx->stream));
    DMK_CUDA(ctx, cudaMemcpyAsync(c->d_scale, scale, n * 12, cudaMemcpyHostToDevice, ctx->stream));
    DMK_CUDA(ctx, cudaMemcpyAsync(c->d_bbox, bbox, n * 24, cudaMemcpyHostToDevice, ctx->stream));
    DMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    c->instances_set = true;
    c->use_trs = true;
    return DMK_OK;
}

dmk_status dmk_crowd_set_transform_nodes(dmk_crowd_t c, int num_nodes, const float* position, const float* rotation, const float* scale,
                                         const int32_t* node_parent, const int32_t* entity_parent) {
    if (!c) return DMK_ERR_INVALID;
    dmk_context_t ctx = c->ctx;
    if (num_nodes < 0) return fail(ctx, DMK_ERR_INVALID, "negative node count");
    if (bind(ctx)) return DMK_ERR_CUDA;
    DMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // a running cull may still read the old arrays
    for (void* p : {(void*)c->d_node_position, (void*)c->d_node_rotation, (void*)c->d_node_scale, (void*)c->d_node_world,
                    (void*)c->d_node_world_inverse, (void*)c->d_node_parent, (void*)c->d_entity_parent})
        dev_free(p);
    c->d_node_position = c->d_node_rotation = c->d_node_scale = c->d_node_world = c->d_node_world_inverse = nullptr;
    c->d_node_parent = c->d_entity_parent = nullptr;
    c->num_nodes = 0;
    if (num_nodes == 0) return DMK_OK;
    if (!position || !rotation || !scale || !node_parent || !entity_parent) return fail(ctx, DMK_ERR_INVALID, "null node array");
    const size_t m = static_cast<size_t>(num_nodes), n = static_cast<size_t>(c->n);
    for (size_t i = 0; i < m; ++i) {   // a forest: parents in range, no cycles, bounded depth
        int depth = 0;
        for (int k = static_cast<int>(i); k >= 0; k = node_parent[k]) {
            if (k >= num_nodes) return fail(ctx, DMK_ERR_INVALID, "node " + std::to_string(i) + ": parent index out of range");
            if (++depth > 32) return fail(ctx, DMK_ERR_INVALID, "node " + std::to_string(i) + ": parent chain deeper than 32 (or cyclic)");
        }
    }
    for (size_t i = 0; i < n; ++i)
        if (entity_parent[i] < -1 || entity_parent[i] >= num_nodes)
            return fail(ctx, DMK_ERR_INVALID, "entity " + std::to_string(i) + ": parent node out of range");
    cudaStream_t s = ctx->stream;
    DMK_CUDA(ctx, dev_alloc(&c->d_node_position, m * 3));
    DMK_CUDA(ctx, dev_alloc(&c->d_node_rotation, m * 4));
    DMK_CUDA(ctx, dev_alloc(&c->d_node_scale, m * 3));
    DMK_CUDA(ctx, dev_alloc(&c->d_node_world, m * 16));
    DMK_CUDA(ctx, dev_alloc(&c->d_node_world_inverse, m * 16));
    DMK_CUDA(ctx, dev_alloc(&c->d_node_parent, m));
    DMK_CUDA(ctx, dev_alloc(&c->d_entity_parent, n));
    DMK_CUDA(ctx, cudaMemcpyAsync(c->d_node_position, position, m * 12, cudaMemcpyHostToDevice, s));
    DMK_CUDA(ctx, cudaMemcpyAsync(c->d_node_rotation, rotation, m * 16, cudaMemcpyHostToDevice, s));
    DMK_CUDA(ctx, cudaMemcpyAsync(c->d_node_scale, scale, m * 12, cudaMemcpyHostToDevice, s));
    DMK_CUDA(ctx, cudaMemcpyAsync(c->d_node_parent, node_parent, m * 4, cudaMemcpyHostToDevice, s));
    DMK_CUDA(ctx, cudaMemcpyAsync(c->d_entity_parent, entity_parent, n * 4, cudaMemcpyHostToDevice, s));
    DMK_CUDA(ctx, cudaStreamSynchronize(s));
    c->num_nodes = num_nodes;
    return DMK_OK;
}

dmk_status dmk_crowd_get_world_inverses(dmk_crowd_t c, int64_t first, int64_t count, float* out) {
    if (!c) return DMK_ERR_INVALID;
    dmk_context_t ctx = c->ctx;
    if (!out || first < 0 || count < 0 || first + count > c->n) return fail(ctx, DMK_ERR_INVALID, "character range out of bounds");
    if (!c->d_world_inverse || !c->culled_once || (c->use_trs && !c->keep_world_inverse))
        return fail(ctx, DMK_ERR_INVALID, "needs DMK_OPT_KEEP_WORLD_INVERSE (or dmk_crowd_set_instances) and a cull step");
    if (bind(ctx)) return DMK_ERR_CUDA;
    DMK_CUDA(ctx, cudaMemcpyAsync(out, c->d_world_inverse + first * 16, static_cast<size_t>(count) * 64, cudaMemcpyDeviceToHost, ctx->stream));
    DMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DMK_OK;
}
const float* dmk_crowd_world_inverse_dev(dmk_crowd_t c) { return c ? c->d_world_inverse : nullptr; }

dmk_status dmk_crowd_get_transform_nodes(dmk_crowd_t c, float* out_world, float* out_world_inverse) {
    if (!c) return DMK_ERR_INVALID;
    dmk_context_t ctx = c->ctx;
    if (!c->num_nodes || !c->culled_once) return fail(ctx, DMK_ERR_INVALID, "needs dmk_crowd_set_transform_nodes and a cull step");
    if (bind(ctx)) return DMK_ERR_CUDA;
    const size_t bytes = static_cast<size_t>(c->num_nodes) * 64;
    if (out_world) DMK_CUDA(ctx, cudaMemcpyAsync(out_world, c->d_node_world, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_world_inverse) DMK_CUDA(ctx, cudaMemcpyAsync(out_world_inverse, c->d_node_world_inverse, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    DMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DMK_OK;
}

dmk_status dmk_frustum_corners(const float* view_proj, float near_depth, float* out_corners) {
    if (!view_proj || !out_corners) return DMK_ERR_INVALID;
    frustum_corners_glm(view_proj, near_depth, out_corners);
    return DMK_OK;
}

dmk_status dmk_crowd_set_camera(dmk_crowd_t c, const float* view_proj, float near_depth) {
    if (!c) return DMK_ERR_INVALID;
    if (!view_proj) return fail(c->ctx, DMK_ERR_INVALID, "null view_proj");
    frustum_corners_glm(view_proj, near_depth, c->corners);
    c->camera_set = true;
    return DMK_OK;
}
dmk_status dmk_crowd_set_frustum_corners(dmk_crowd_t c, const float* corners) {
    if (!c) return DMK_ERR_INVALID;
    if (!corners) return fail(c->ctx, DMK_ERR_INVALID, "null corners");
    std::memcpy(c->corners, corners, sizeof c->corners);
    c->camera_set = true;
    return DMK_OK;
}

dmk_status dmk_crowd_step(dmk_crowd_t c, float dt, uint32_t flags) {
    if (!c) return DMK_ERR_INVALID;
    dmk_context_t ctx = c->ctx;
    if (bind(ctx)) return DMK_ERR_CUDA;
    cudaStream_t s = ctx->stream;
    if ((flags & DMK_STEP_ADVANCE) && !(flags & DMK_STEP_POSE))
        return fail(ctx, DMK_ERR_UNSUPPORTED, "DMK_STEP_ADVANCE runs inside the pose kernel: combine it with DMK_STEP_POSE");
    if (flags & DMK_STEP_ANIMATOR) {
        if (!c->animator_set) return fail(ctx, DMK_ERR_INVALID, "dmk_crowd_set_animator has not been called");
        if (flags & DMK_STEP_ADVANCE) return fail(ctx, DMK_ERR_UNSUPPORTED, "the device animators own the clip clocks: do not combine them with DMK_STEP_ADVANCE");
        const AnimatorParams a = animator_params(c, dt);
        {
            KernelTimer timer(ctx, DMK_KERNEL_ANIMATOR);
            DMK_CUDA(ctx, launch_animator(a, s));
        }
        ++ctx->launches;
        c->num_layers = c->anim_layers;
        c->layers_set = true;
        c->use_program = true;
        c->has_loop = false;
        c->animator_ticked = true;
    }
    if (flags & DMK_STEP_POSE) {
        if (!c->layers_set) return fail(ctx, DMK_ERR_INVALID, "dmk_crowd_set_layers has not been called");
        // BlendingJob::Validate: threshold > 0 (darmok: "error in the blending job", src/skeleton_ozz.cpp:605-608)
        if (!c->use_program && c->num_layers > 1 && !(c->threshold > 0.f)) return fail(ctx, DMK_ERR_JOB, "error in the blending job");
        if (c->use_program && (flags & DMK_STEP_ADVANCE))
            return fail(ctx, DMK_ERR_UNSUPPORTED, "pose programs carry host-resolved ratios: do not combine them with DMK_STEP_ADVANCE");
        PoseParams p{};
        p.clips.dir = c->clips->d_dir;
        for (int k = 0; k < 3; ++k) {
            p.clips.ratio[k] = c->clips->d_ratio[k];
            p.clips.value[k] = c->clips->d_value[k];
            p.clips.stream[k] = c->clips->d_stream[k];
        }
        p.clips.duration = c->clips->d_duration;
        if (c->clips->has_buckets && !c->force_search) {
            p.clips.bucket_dir = c->clips->d_bucket_dir;
            p.clips.bucket_a = c->clips->d_bucket_a;
            p.clips.bucket_b = c->clips->d_bucket_b;
            p.clips.bucket_c = c->clips->d_bucket_c;
        }
        p.clips.num_clips = c->clips->num_clips;
        p.clips.padded_tracks = c->clips->padded_tracks;
        p.skel.parents = c->skel->d_parents;
        p.skel.schedule = c->skel->d_schedule;
        p.skel.rest = c->skel->d_rest;
        p.skel.num_joints = c->skel->data.num_joints;
        p.skel.padded_joints = c->skel->data.num_soa() * 4;
        p.skel.num_soa = c->skel->data.num_soa();
        p.skel.rounds = c->skel->rounds;
        if (c->arm) {
            p.arm.remap = c->arm->d_remap;
            p.arm.inv_bind = c->arm->d_inv_bind;
            p.arm.num_joints = c->arm->num_joints;
        }
        p.n = c->n;
        p.num_layers = c->num_layers;
        p.clip = c->d_clip;
        p.ratio = c->d_ratio;
        p.weight = c->d_weight;
        p.speed = c->d_speed;
        p.loop = c->has_loop ? c->d_loop : nullptr;
        p.looped = c->d_looped;
        p.threshold = c->threshold;
        p.program = c->use_program ? c->d_program : nullptr;
        p.dt = dt;
        p.advance = (flags & DMK_STEP_ADVANCE) ? 1 : 0;
        p.palette = c->d_palette;
        p.palette34 = c->d_palette34;
        p.models = c->d_models;
        p.locals = c->d_locals;
        p.key_indices = c->d_key_indices;
        p.error_flag = c->d_error;
        p.scratch = c->d_local_matrices;
        int pose_launches = 0;
        {
            KernelTimer timer(ctx, DMK_KERNEL_POSE);
            DMK_CUDA(ctx, launch_pose(p, ctx->num_sms, s, &pose_launches));
            // every further armature: its own palette from the same local matrices (beforeRenderEntity runs once per Skinnable)
            for (const auto& x : c->extra_skins) {
                PoseParams q = p;
                q.arm.remap = x.arm->d_remap;
                q.arm.inv_bind = x.arm->d_inv_bind;
                q.arm.num_joints = x.arm->num_joints;
                q.palette = x.d_palette;
                q.palette34 = x.d_palette34;
                q.models = nullptr;
                q.locals = nullptr;
                q.key_indices = nullptr;
                DMK_CUDA(ctx, launch_model(q, ctx->num_sms, s, &pose_launches));
            }
        }
        ctx->launches += pose_launches;
    }
    if (flags & DMK_STEP_CULL) {
        if (!c->instances_set || !c->camera_set)
            return fail(ctx, DMK_ERR_INVALID, "cull needs dmk_crowd_set_instances and dmk_crowd_set_camera");
        CullParams p{};
        int node_launches = 0;
        if (c->use_trs) {
            p.position = c->d_position;
            p.rotation = c->d_rotation;
            p.scale = c->d_scale;
            if (c->num_nodes) {
                DMK_CUDA(ctx, launch_transform_nodes(c->d_node_position, c->d_node_rotation, c->d_node_scale, c->d_node_parent, c->num_nodes,
                                                     c->d_node_world, c->d_node_world_inverse, s));
                node_launches = 1;
                p.entity_parent = c->d_entity_parent;
                p.node_world_inverse = c->d_node_world_inverse;
            }
            if (c->keep_world_inverse) p.world_inverse_out = c->d_world_inverse;
        } else {
            p.world_inverse = c->d_world_inverse;
        }
        p.bbox = c->d_bbox;
        p.n = c->n;
        p.vis_bits = c->d_vis_bits;
        std::memcpy(p.corners, c->corners, sizeof p.corners);
        int extra = 0;
        {
            KernelTimer timer(ctx, DMK_KERNEL_CULL);
            DMK_CUDA(ctx, launch_cull(p, s));
            DMK_CUDA(ctx, launch_compact(c->d_vis_bits, c->n, c->d_visible_ids, c->d_visible_count, c->d_scratch, s, &extra));
        }
        ctx->launches += 1 + extra + node_launches;
        c->culled_once = true;
        if (flags & DMK_STEP_READBACK_VISIBILITY) {
            if (c->vis_pending == 2) return fail(ctx, DMK_ERR_INVALID, "two visibility read-backs are already outstanding: collect one first");
            const int b = (c->vis_head + c->vis_pending) & 1;
            const size_t words = static_cast<size_t>((c->n + 31) / 32);
            for (int k = 0; k < 2 && !c->h_vis[1]; ++k) {   // both ring slots at first use: pinned allocation synchronises
                DMK_CUDA(ctx, cudaHostAlloc(reinterpret_cast<void**>(&c->h_vis[k]), (words + 1) * 4, cudaHostAllocDefault));
                DMK_CUDA(ctx, cudaEventCreateWithFlags(&c->vis_ready[k], cudaEventDisableTiming));
            }
            DMK_CUDA(ctx, cudaMemcpyAsync(c->h_vis[b], c->d_vis_bits, words * 4, cudaMemcpyDeviceToHost, s));
            DMK_CUDA(ctx, cudaMemcpyAsync(c->h_vis[b] + words, c->d_visible_count, 4, cudaMemcpyDeviceToHost, s));
            DMK_CUDA(ctx, cudaEventRecord(c->vis_ready[b], s));
            ++c->vis_pending;
        }
    }
    if (flags & DMK_STEP_SKIN) {
        bool any_mesh = c->mesh != nullptr;
        for (const auto& x : c->extra_skins) any_mesh = any_mesh || x.mesh != nullptr;
        if (!any_mesh || !c->arm) return fail(ctx, DMK_ERR_INVALID, "skinning needs a mesh and an armature");
        const bool visible_only = (flags & DMK_STEP_SKIN_VISIBLE_ONLY) != 0;
        if (visible_only && !c->culled_once) return fail(ctx, DMK_ERR_INVALID, "DMK_STEP_SKIN_VISIBLE_ONLY needs a cull step");
        SkinParams p{};
        p.char_ids = visible_only ? c->d_visible_ids : nullptr;
        p.char_count = visible_only ? c->d_visible_count : nullptr;
        p.n = c->n;
        if (c->mesh) {
            p.palette34 = c->d_palette34;
            p.pos = c->mesh->d_pos;
            p.nrm = c->mesh->d_nrm;
            p.tan = c->mesh->d_tan;
            p.weights = c->mesh->d_weights;
            p.slots = c->mesh->d_slots;
            p.out = c->d_skinned;
            p.vpad = c->mesh->vpad;
            p.palette_size = c->mesh->palette_size;
            p.num_tiles = c->plan.num_tiles;
            p.tile_vertices = c->plan.tile_vertices;
            p.skip_zero_weights = c->mesh->sorted ? 1 : 0;
            {
                KernelTimer timer(ctx, DMK_KERNEL_SKIN);
                DMK_CUDA(ctx, launch_skin(p, c->plan, s));
            }
            ++ctx->launches;
        }
        for (const auto& x : c->extra_skins) {
            if (!x.mesh) continue;
            SkinParams q = p;
            q.palette34 = x.d_palette34;
            q.pos = x.mesh->d_pos;
            q.nrm = x.mesh->d_nrm;
            q.tan = x.mesh->d_tan;
            q.weights = x.mesh->d_weights;
            q.slots = x.mesh->d_slots;
            q.out = x.d_skinned;
            q.vpad = x.mesh->vpad;
            q.palette_size = x.mesh->palette_size;
            q.num_tiles = x.plan.num_tiles;
            q.tile_vertices = x.plan.tile_vertices;
            q.skip_zero_weights = x.mesh->sorted ? 1 : 0;
            {
                KernelTimer timer(ctx, DMK_KERNEL_SKIN);
                DMK_CUDA(ctx, launch_skin(q, x.plan, s));
            }
            ++ctx->launches;
        }
    }
    if (flags & DMK_STEP_POSE)
        DMK_CUDA(ctx, cudaMemcpyAsync(c->h_error, c->d_error, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    return DMK_OK;
}

// ---- results -----------------------------------------------------------------------------------------------
const float* dmk_crowd_palette_dev(dmk_crowd_t c) { return c ? c->d_palette : nullptr; }
const float* dmk_crowd_skinned_dev(dmk_crowd_t c) { return c ? c->d_skinned : nullptr; }
const uint32_t* dmk_crowd_visibility_dev(dmk_crowd_t c) { return c ? c->d_vis_bits : nullptr; }
const int32_t* dmk_crowd_visible_ids_dev(dmk_crowd_t c) { return c ? c->d_visible_ids : nullptr; }
const int32_t* dmk_crowd_visible_count_dev(dmk_crowd_t c) { return c ? c->d_visible_count : nullptr; }

static dmk_status finish_and_check(dmk_crowd_t c) {
    dmk_context_t ctx = c->ctx;
    DMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (*c->h_error != 0)
        return fail(ctx, DMK_ERR_INVALID, "character " + std::to_string(*c->h_error - 1) + " references a clip outside the clip set");
    return DMK_OK;
}
static dmk_status check_range(dmk_crowd_t c, int64_t first, int64_t count, const void* out) {
    if (!c) return DMK_ERR_INVALID;
    if (!out || first < 0 || count < 0 || first + count > c->n) return fail(c->ctx, DMK_ERR_INVALID, "character range out of bounds");
    if (bind(c->ctx)) return DMK_ERR_CUDA;
    return DMK_OK;
}

dmk_status dmk_crowd_get_palette(dmk_crowd_t c, int64_t first, int64_t count, float* out) {
    if (dmk_status st = check_range(c, first, count, out)) return st;
    if (!c->arm) return fail(c->ctx, DMK_ERR_INVALID, "crowd has no armature");
    if (count == 0) return DMK_OK;
    const size_t per = static_cast<size_t>(c->arm->num_joints + 1) * 16;
    DMK_CUDA(c->ctx, cudaMemcpyAsync(out, c->d_palette + first * per, count * per * 4, cudaMemcpyDeviceToHost, c->ctx->stream));
    return finish_and_check(c);
}
dmk_status dmk_crowd_get_models(dmk_crowd_t c, int64_t first, int64_t count, float* out) {
    if (dmk_status st = check_range(c, first, count, out)) return st;
    if (!c->d_models) return fail(c->ctx, DMK_ERR_INVALID, "enable DMK_OUT_MODELS with dmk_crowd_enable_outputs before the step");
    const size_t per = static_cast<size_t>(c->skel->data.num_joints) * 16;
    DMK_CUDA(c->ctx, cudaMemcpyAsync(out, c->d_models + first * per, count * per * 4, cudaMemcpyDeviceToHost, c->ctx->stream));
    return finish_and_check(c);
}
dmk_status dmk_crowd_get_locals(dmk_crowd_t c, int64_t first, int64_t count, float* out) {
    if (dmk_status st = check_range(c, first, count, out)) return st;
    if (!c->d_locals) return fail(c->ctx, DMK_ERR_INVALID, "enable DMK_OUT_LOCALS with dmk_crowd_enable_outputs before the step");
    const size_t per = static_cast<size_t>(c->skel->data.num_soa()) * kSoaFloats;
    DMK_CUDA(c->ctx, cudaMemcpyAsync(out, c->d_locals + first * per, count * per * 4, cudaMemcpyDeviceToHost, c->ctx->stream));
    return finish_and_check(c);
}
dmk_status dmk_crowd_get_key_indices(dmk_crowd_t c, int64_t first, int64_t count, int32_t* out) {
    if (dmk_status st = check_range(c, first, count, out)) return st;
    if (!c->d_key_indices) return fail(c->ctx, DMK_ERR_INVALID, "enable DMK_OUT_KEY_INDICES with dmk_crowd_enable_outputs before the step");
    // device layout uses the current num_layers as the layer stride
    const size_t per = static_cast<size_t>(c->num_layers) * 3 * c->skel->data.num_soa() * 4 * 2;
    DMK_CUDA(c->ctx, cudaMemcpyAsync(out, c->d_key_indices + first * per, count * per * 4, cudaMemcpyDeviceToHost, c->ctx->stream));
    return finish_and_check(c);
}
dmk_status dmk_crowd_get_bone_matrices(dmk_crowd_t c, int64_t first, int64_t count, const float* dir, float* out, uint8_t* out_has_bone) {
    if (dmk_status st = check_range(c, first, count, out)) return st;
    dmk_context_t ctx = c->ctx;
    if (!dir) return fail(ctx, DMK_ERR_INVALID, "null direction");
    if (!c->d_models) return fail(ctx, DMK_ERR_INVALID, "enable DMK_OUT_MODELS with dmk_crowd_enable_outputs before the step");
    const int J = c->skel->data.num_joints;
    std::vector<int16_t> first_child(static_cast<size_t>(J), -1);
    for (int i = 0; i < J; ++i) {
        const int p = c->skel->data.parents[static_cast<size_t>(i)];
        if (p >= 0 && first_child[static_cast<size_t>(p)] < 0) first_child[static_cast<size_t>(p)] = static_cast<int16_t>(i);
    }
    if (out_has_bone)
        for (int i = 0; i < J; ++i) out_has_bone[i] = first_child[static_cast<size_t>(i)] >= 0;
    if (count == 0) return DMK_OK;
    int16_t* d_first = nullptr;
    float* d_out = nullptr;
    DMK_CUDA(ctx, upload(ctx, &d_first, first_child));
    cudaError_t e = dev_alloc(&d_out, static_cast<size_t>(count) * J * 16);
    // angleAxis(pi, axis) of glm::rotation's opposite-direction case: sin / cos of pi / 2 in float, from the host's libm
    const float half = 3.14159265358979323846264338327950288f * 0.5f;
    if (e == cudaSuccess)
        e = launch_bone_matrices(c->d_models + first * J * 16, d_first, J, count, dir, std::sin(half), std::cos(half), d_out, ctx->stream);
    ++ctx->launches;
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, static_cast<size_t>(count) * J * 64, cudaMemcpyDeviceToHost, ctx->stream);
    const cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
    dev_free(d_first);
    dev_free(d_out);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "bone matrices");
    if (e2 != cudaSuccess) return cuda_fail(ctx, e2, "bone matrices");
    if (*c->h_error != 0)
        return fail(ctx, DMK_ERR_INVALID, "character " + std::to_string(*c->h_error - 1) + " references a clip outside the clip set");
    return DMK_OK;
}
// [count][V][9] on the host from either device layout
static dmk_status read_skinned(dmk_crowd_t c, dmk_mesh_t mesh, const float* d_skinned, int64_t first, int64_t count, float* out) {
    const size_t vpad = static_cast<size_t>(mesh->vpad), nv = static_cast<size_t>(mesh->num_vertices);
    const size_t pitch = vpad * 9, row = nv * 9;
    if (!c->skin_variant) {
        DMK_CUDA(c->ctx, cudaMemcpy2DAsync(out, row * 4, d_skinned + first * pitch, pitch * 4, row * 4, count, cudaMemcpyDeviceToHost, c->ctx->stream));
        return finish_and_check(c);
    }
    // on the device: nine component planes [9][vpad] (1, 2) or three float3 streams [3][vpad][3] (3) per character
    std::vector<float> planes;
    try {
        planes.resize(static_cast<size_t>(count) * pitch);
    } catch (const std::bad_alloc&) {   // no exception crosses the C ABI
        return fail(c->ctx, DMK_ERR_INVALID, "not enough host memory to re-interleave that many characters: read a smaller range");
    }
    DMK_CUDA(c->ctx, cudaMemcpyAsync(planes.data(), d_skinned + first * pitch, planes.size() * 4, cudaMemcpyDeviceToHost, c->ctx->stream));
    if (dmk_status st = finish_and_check(c)) return st;
    for (size_t i = 0; i < static_cast<size_t>(count); ++i) {
        if (c->skin_variant == 3) {
            for (size_t a = 0; a < 3; ++a) {
                const float* src = planes.data() + i * pitch + a * vpad * 3;
                float* dst = out + i * row + a * 3;
                for (size_t v = 0; v < nv; ++v)
                    for (size_t r = 0; r < 3; ++r) dst[v * 9 + r] = src[v * 3 + r];
            }
            continue;
        }
        for (size_t k = 0; k < 9; ++k) {
            const float* src = planes.data() + i * pitch + k * vpad;
            float* dst = out + i * row + k;
            for (size_t v = 0; v < nv; ++v) dst[v * 9] = src[v];
        }
    }
    return DMK_OK;
}
dmk_status dmk_crowd_get_skinned(dmk_crowd_t c, int64_t first, int64_t count, float* out) {
    if (dmk_status st = check_range(c, first, count, out)) return st;
    if (!c->mesh) return fail(c->ctx, DMK_ERR_INVALID, "crowd has no mesh");
    if (count == 0) return DMK_OK;
    return read_skinned(c, c->mesh, c->d_skinned, first, count, out);
}
// ---- several skinned meshes per character -----------------------------------------------------------------------
dmk_status dmk_crowd_add_skin(dmk_crowd_t c, dmk_armature_t arm, dmk_mesh_t mesh, int* out_skin) {
    if (!c) return DMK_ERR_INVALID;
    dmk_context_t ctx = c->ctx;
    if (!arm) return fail(ctx, DMK_ERR_INVALID, "a skin needs an armature");
    if (arm->ctx != ctx || (mesh && mesh->ctx != ctx)) return fail(ctx, DMK_ERR_INVALID, "armature / mesh belong to another context");
    if (mesh && mesh->palette_size != arm->num_joints + 1) return fail(ctx, DMK_ERR_INVALID, "mesh palette_size does not match the armature");
    if (!c->arm) return fail(ctx, DMK_ERR_INVALID, "the crowd was created without an armature: pass the first one to dmk_crowd_create");
    if (bind(ctx)) return DMK_ERR_CUDA;
    dmk_crowd_s::ExtraSkin x;
    x.arm = arm;
    x.mesh = mesh;
    const size_t ps = static_cast<size_t>(arm->num_joints + 1), n = static_cast<size_t>(c->n);
    cudaError_t e = dev_alloc(&x.d_palette, n * ps * 16);
    if (e == cudaSuccess) e = dev_alloc(&x.d_palette34, n * ps * 16);
    if (e == cudaSuccess && mesh) e = dev_alloc(&x.d_skinned, n * static_cast<size_t>(mesh->vpad) * 9);
    if (e != cudaSuccess) {
        dev_free(x.d_palette);
        dev_free(x.d_palette34);
        dev_free(x.d_skinned);
        return cuda_fail(ctx, e, "dmk_crowd_add_skin allocation");
    }
    if (mesh && c->skin_variant && plan_skin(mesh->vpad, mesh->palette_size, ctx->num_sms, c->skin_variant).replicas != 8) {
        dev_free(x.d_palette);
        dev_free(x.d_palette34);
        dev_free(x.d_skinned);
        return fail(ctx, DMK_ERR_UNSUPPORTED, "planar skinned output needs a palette of at most 181 slots");
    }
    if (mesh) x.plan = plan_skin(mesh->vpad, mesh->palette_size, ctx->num_sms - c->reserved_sms, c->skin_variant);
    c->extra_skins.push_back(x);
    if (out_skin) *out_skin = static_cast<int>(c->extra_skins.size());
    return DMK_OK;
}
int dmk_crowd_num_skins(dmk_crowd_t c) { return c ? (c->arm ? 1 : 0) + static_cast<int>(c->extra_skins.size()) : 0; }
namespace {
struct SkinView {
    dmk_armature_t arm = nullptr;
    dmk_mesh_t mesh = nullptr;
    float *palette = nullptr, *skinned = nullptr;
};
bool skin_view(dmk_crowd_t c, int skin, SkinView* v) {
    if (skin == 0) {
        *v = {c->arm, c->mesh, c->d_palette, c->d_skinned};
        return c->arm != nullptr;
    }
    if (skin < 0 || static_cast<size_t>(skin) > c->extra_skins.size()) return false;
    const auto& x = c->extra_skins[static_cast<size_t>(skin) - 1];
    *v = {x.arm, x.mesh, x.d_palette, x.d_skinned};
    return true;
}
}  // namespace
const float* dmk_crowd_skin_palette_dev(dmk_crowd_t c, int skin) {
    SkinView v;
    return c && skin_view(c, skin, &v) ? v.palette : nullptr;
}
const float* dmk_crowd_skin_skinned_dev(dmk_crowd_t c, int skin) {
    SkinView v;
    return c && skin_view(c, skin, &v) ? v.skinned : nullptr;
}
dmk_status dmk_crowd_get_skin_palette(dmk_crowd_t c, int skin, int64_t first, int64_t count, float* out) {
    if (dmk_status st = check_range(c, first, count, out)) return st;
    SkinView v;
    if (!skin_view(c, skin, &v)) return fail(c->ctx, DMK_ERR_INVALID, "no such skin");
    if (count == 0) return DMK_OK;
    const size_t per = static_cast<size_t>(v.arm->num_joints + 1) * 16;
    DMK_CUDA(c->ctx, cudaMemcpyAsync(out, v.palette + first * per, count * per * 4, cudaMemcpyDeviceToHost, c->ctx->stream));
    return finish_and_check(c);
}
dmk_status dmk_crowd_get_skin_skinned(dmk_crowd_t c, int skin, int64_t first, int64_t count, float* out) {
    if (dmk_status st = check_range(c, first, count, out)) return st;
    SkinView v;
    if (!skin_view(c, skin, &v)) return fail(c->ctx, DMK_ERR_INVALID, "no such skin");
    if (!v.mesh) return fail(c->ctx, DMK_ERR_INVALID, "the skin has no mesh");
    if (count == 0) return DMK_OK;
    return read_skinned(c, v.mesh, v.skinned, first, count, out);
}

dmk_status dmk_crowd_get_visibility(dmk_crowd_t c, uint32_t* out_bits, int64_t* out_visible_count) {
    if (!c) return DMK_ERR_INVALID;
    if (!c->culled_once) return fail(c->ctx, DMK_ERR_INVALID, "no cull step has run");
    if (bind(c->ctx)) return DMK_ERR_CUDA;
    const size_t words = static_cast<size_t>((c->n + 31) / 32);
    if (out_bits) DMK_CUDA(c->ctx, cudaMemcpyAsync(out_bits, c->d_vis_bits, words * 4, cudaMemcpyDeviceToHost, c->ctx->stream));
    int32_t count = 0;
    DMK_CUDA(c->ctx, cudaMemcpyAsync(&count, c->d_visible_count, 4, cudaMemcpyDeviceToHost, c->ctx->stream));
    if (dmk_status st = finish_and_check(c)) return st;
    if (out_visible_count) *out_visible_count = count;
    return DMK_OK;
}
dmk_status dmk_crowd_collect_visibility(dmk_crowd_t c, uint32_t* out_bits, int64_t* out_visible_count) {
    if (!c) return DMK_ERR_INVALID;
    if (c->vis_pending == 0) return fail(c->ctx, DMK_ERR_INVALID, "no step with DMK_STEP_READBACK_VISIBILITY is outstanding");
    if (bind(c->ctx)) return DMK_ERR_CUDA;
    const int b = c->vis_head;
    DMK_CUDA(c->ctx, cudaEventSynchronize(c->vis_ready[b]));
    const size_t words = static_cast<size_t>((c->n + 31) / 32);
    if (out_bits) std::memcpy(out_bits, c->h_vis[b], words * 4);
    if (out_visible_count) *out_visible_count = static_cast<int32_t>(c->h_vis[b][words]);
    c->vis_head ^= 1;
    --c->vis_pending;
    return DMK_OK;
}
dmk_status dmk_crowd_get_ratios(dmk_crowd_t c, float* out_ratio, uint8_t* out_looped) {
    if (!c) return DMK_ERR_INVALID;
    if (!c->layers_set) return fail(c->ctx, DMK_ERR_INVALID, "dmk_crowd_set_layers has not been called");
    if (bind(c->ctx)) return DMK_ERR_CUDA;
    const size_t ln = static_cast<size_t>(c->num_layers) * static_cast<size_t>(c->n);
    if (out_ratio) DMK_CUDA(c->ctx, cudaMemcpyAsync(out_ratio, c->d_ratio, ln * 4, cudaMemcpyDeviceToHost, c->ctx->stream));
    if (out_looped) DMK_CUDA(c->ctx, cudaMemcpyAsync(out_looped, c->d_looped, ln, cudaMemcpyDeviceToHost, c->ctx->stream));
    return finish_and_check(c);
}

dmk_status dmk_crowd_pack_visible_palettes(dmk_crowd_t c, float* out_dev, int64_t capacity) {
    if (!c) return DMK_ERR_INVALID;
    if (!out_dev || capacity < 0) return fail(c->ctx, DMK_ERR_INVALID, "null output");
    if (!c->arm || !c->culled_once) return fail(c->ctx, DMK_ERR_INVALID, "needs an armature and a cull step");
    if (c->n == 0) return DMK_OK;
    if (bind(c->ctx)) return DMK_ERR_CUDA;
    {
        KernelTimer timer(c->ctx, DMK_KERNEL_PACK);
        DMK_CUDA(c->ctx, launch_pack_palettes(c->d_palette, (c->arm->num_joints + 1) * 16, c->d_visible_ids, c->d_visible_count, capacity,
                                              out_dev, c->ctx->num_sms, c->ctx->stream));
    }
    ++c->ctx->launches;
    return DMK_OK;
}

// ---- visible-palette exchange over peer memory ---------------------------------------------------------------
static dmk_status exchange_check_shape(dmk_context_t ctx, int world, int rank, int64_t capacity, int palette_size) {
    if (world < 1 || world > kExchangeMaxRanks) return fail(ctx, DMK_ERR_INVALID, "world must be in [1, DMK_EXCHANGE_MAX_RANKS]");
    if (rank < 0 || rank >= world) return fail(ctx, DMK_ERR_INVALID, "rank out of range");
    if (capacity < 0 || capacity > 0x7fffffff) return fail(ctx, DMK_ERR_INVALID, "capacity out of range");
    if (palette_size < 1 || palette_size > kMaxPalette) return fail(ctx, DMK_ERR_INVALID, "palette_size out of range");
    return DMK_OK;
}
static dmk_status exchange_alloc_local(dmk_exchange_t x) {
    DMK_CUDA(x->ctx, dev_alloc(&x->d_local, 2));
    DMK_CUDA(x->ctx, cudaMemsetAsync(x->d_local, 0, 2 * sizeof(uint32_t), x->ctx->stream));
    DMK_CUDA(x->ctx, cudaStreamSynchronize(x->ctx->stream));
    return DMK_OK;
}
void dmk_exchange_destroy(dmk_exchange_t x) {
    if (!x) return;
    cudaSetDevice(x->ctx->device);
    cudaDeviceSynchronize();   // the exchange may have run on side streams
    if (x->base) {
        if (x->kind == dmk_exchange_s::Owner) cudaFree(x->base);
        else if (x->kind == dmk_exchange_s::Opened) cudaIpcCloseMemHandle(x->base);
    }
    dev_free(x->d_local);
    dev_free(x->d_counts);
    delete x;
}
dmk_status dmk_exchange_create(dmk_context_t ctx, int world, int rank, int64_t capacity, int palette_size, uint8_t* handle_out,
                               dmk_exchange_t* out) {
    if (!ctx || !out) return fail(ctx, DMK_ERR_INVALID, "null argument");
    *out = nullptr;
    static_assert(sizeof(cudaIpcMemHandle_t) == DMK_EXCHANGE_HANDLE_BYTES, "DMK_EXCHANGE_HANDLE_BYTES");
    static_assert(kExchangeMaxRanks == DMK_EXCHANGE_MAX_RANKS, "DMK_EXCHANGE_MAX_RANKS");
    if (dmk_status st = exchange_check_shape(ctx, world, rank, capacity, palette_size)) return st;
    if (bind(ctx)) return DMK_ERR_CUDA;
    Building<dmk_exchange_s, dmk_exchange_destroy> x(new dmk_exchange_s());
    x->ctx = ctx;
    x->kind = dmk_exchange_s::Owner;
    x->world = world;
    x->rank = rank;
    x->capacity = capacity;
    x->palette_size = palette_size;
    DMK_CUDA(ctx, cudaMalloc(&x->base, x->bytes()));
    DMK_CUDA(ctx, cudaMemsetAsync(x->base, 0, kExchangeHeaderBytes, ctx->stream));
    DMK_CUDA(ctx, dev_alloc(&x->d_counts, 2 * kExchangeMaxRanks));
    DMK_CUDA(ctx, cudaMemsetAsync(x->d_counts, 0, 2 * kExchangeMaxRanks * sizeof(int32_t), ctx->stream));
    if (dmk_status st = exchange_alloc_local(x.get())) return st;
    if (handle_out) {
        cudaIpcMemHandle_t h;
        DMK_CUDA(ctx, cudaIpcGetMemHandle(&h, x->base));
        std::memcpy(handle_out, &h, sizeof h);
    }
    *out = x.release();
    return DMK_OK;
}
dmk_status dmk_exchange_open(dmk_context_t ctx, int world, int rank, int64_t capacity, int palette_size, const uint8_t* handle,
                             dmk_exchange_t* out) {
    if (!ctx || !out || !handle) return fail(ctx, DMK_ERR_INVALID, "null argument");
    *out = nullptr;
    if (dmk_status st = exchange_check_shape(ctx, world, rank, capacity, palette_size)) return st;
    if (bind(ctx)) return DMK_ERR_CUDA;
    Building<dmk_exchange_s, dmk_exchange_destroy> x(new dmk_exchange_s());
    x->ctx = ctx;
    x->kind = dmk_exchange_s::Opened;
    x->world = world;
    x->rank = rank;
    x->capacity = capacity;
    x->palette_size = palette_size;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof h);
    DMK_CUDA(ctx, cudaIpcOpenMemHandle(&x->base, h, cudaIpcMemLazyEnablePeerAccess));
    if (dmk_status st = exchange_alloc_local(x.get())) return st;
    *out = x.release();
    return DMK_OK;
}
dmk_status dmk_exchange_attach(dmk_exchange_t owner, dmk_context_t ctx, int rank, dmk_exchange_t* out) {
    if (!owner || !ctx || !out) return fail(ctx, DMK_ERR_INVALID, "null argument");
    *out = nullptr;
    if (owner->kind != dmk_exchange_s::Owner) return fail(ctx, DMK_ERR_INVALID, "attach needs the owning exchange");
    if (rank < 0 || rank >= owner->world) return fail(ctx, DMK_ERR_INVALID, "rank out of range");
    if (bind(ctx)) return DMK_ERR_CUDA;
    if (ctx->device != owner->ctx->device) {
        int can = 0;
        DMK_CUDA(ctx, cudaDeviceCanAccessPeer(&can, ctx->device, owner->ctx->device));
        if (!can) return fail(ctx, DMK_ERR_UNSUPPORTED, "no peer access to the render GPU");
        const cudaError_t e = cudaDeviceEnablePeerAccess(owner->ctx->device, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        else if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaDeviceEnablePeerAccess");
    }
    Building<dmk_exchange_s, dmk_exchange_destroy> x(new dmk_exchange_s());
    x->ctx = ctx;
    x->kind = dmk_exchange_s::Attached;
    x->world = owner->world;
    x->rank = rank;
    x->capacity = owner->capacity;
    x->palette_size = owner->palette_size;
    x->base = owner->base;
    x->timeout_ns = owner->timeout_ns;
    if (dmk_status st = exchange_alloc_local(x.get())) return st;
    *out = x.release();
    return DMK_OK;
}
dmk_status dmk_exchange_set_timeout_ms(dmk_exchange_t x, int64_t ms) {
    if (!x) return DMK_ERR_INVALID;
    if (ms < 1) return fail(x->ctx, DMK_ERR_INVALID, "timeout must be at least 1 ms");
    x->timeout_ns = static_cast<uint64_t>(ms) * 1000000ull;
    return DMK_OK;
}
dmk_status dmk_crowd_push_visible_palettes(dmk_crowd_t c, dmk_exchange_t x, uint32_t frame, int num_ctas, void* cuda_stream) {
    if (!c || !x) return DMK_ERR_INVALID;
    dmk_context_t ctx = c->ctx;
    if (x->ctx != ctx) return fail(ctx, DMK_ERR_INVALID, "the exchange belongs to another context");
    if (!c->arm) return fail(ctx, DMK_ERR_INVALID, "needs an armature");
    if (c->n > 0 && !c->culled_once) return fail(ctx, DMK_ERR_INVALID, "needs a cull step");
    if (c->arm->num_joints + 1 != x->palette_size) return fail(ctx, DMK_ERR_INVALID, "palette size differs from the exchange's");
    if (frame != x->last_pushed + 1) return fail(ctx, DMK_ERR_INVALID, "frames are pushed in order: expected the previous frame + 1");
    if (num_ctas < 0) return fail(ctx, DMK_ERR_INVALID, "num_ctas is negative");
    if (bind(ctx)) return DMK_ERR_CUDA;
    cudaStream_t s = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->stream;
    if (num_ctas == 0) num_ctas = ctx->num_sms * 4;
    float* slab = x->buffer(frame) + static_cast<size_t>(x->rank) * static_cast<size_t>(x->capacity) * static_cast<size_t>(x->palette_size) * 16;
    {
        KernelTimer timer(ctx, DMK_KERNEL_PACK, s);
        DMK_CUDA(ctx, launch_push_palettes(c->d_palette, x->palette_size * 16, c->d_visible_ids, c->d_visible_count, x->capacity, slab, x->header(),
                                           x->rank, frame, x->d_local, x->d_local + 1, x->timeout_ns, num_ctas, s));
    }
    ++ctx->launches;
    x->last_pushed = frame;
    return DMK_OK;
}
dmk_status dmk_exchange_wait(dmk_exchange_t x, uint32_t frame, void* cuda_stream) {
    if (!x) return DMK_ERR_INVALID;
    dmk_context_t ctx = x->ctx;
    if (x->kind != dmk_exchange_s::Owner) return fail(ctx, DMK_ERR_INVALID, "only the render rank waits");
    if (frame != x->last_waited + 1) return fail(ctx, DMK_ERR_INVALID, "frames are waited for in order: expected the previous frame + 1");
    if (bind(ctx)) return DMK_ERR_CUDA;
    cudaStream_t s = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->stream;
    DMK_CUDA(ctx, launch_wait_published(x->header(), x->world, frame, x->d_counts + (frame & 1u) * kExchangeMaxRanks, x->timeout_ns, s));
    ++ctx->launches;
    x->last_waited = frame;
    return DMK_OK;
}
dmk_status dmk_exchange_frame_dev(dmk_exchange_t x, uint32_t frame, float** out_slab_dev, int32_t** out_counts_dev) {
    if (!x) return DMK_ERR_INVALID;
    if (x->kind != dmk_exchange_s::Owner) return fail(x->ctx, DMK_ERR_INVALID, "only the render rank reads the slab");
    if (out_slab_dev) *out_slab_dev = x->buffer(frame);
    if (out_counts_dev) *out_counts_dev = x->d_counts + (frame & 1u) * kExchangeMaxRanks;
    return DMK_OK;
}
dmk_status dmk_exchange_release(dmk_exchange_t x, uint32_t frame, void* cuda_stream) {
    if (!x) return DMK_ERR_INVALID;
    dmk_context_t ctx = x->ctx;
    if (x->kind != dmk_exchange_s::Owner) return fail(ctx, DMK_ERR_INVALID, "only the render rank releases");
    if (frame != x->last_released + 1 || frame > x->last_waited)
        return fail(ctx, DMK_ERR_INVALID, "frames are released in order, after their wait");
    if (bind(ctx)) return DMK_ERR_CUDA;
    cudaStream_t s = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->stream;
    DMK_CUDA(ctx, launch_release_frame(x->header(), frame, s));
    ++ctx->launches;
    x->last_released = frame;
    return DMK_OK;
}
dmk_status dmk_exchange_status(dmk_exchange_t x, void* cuda_stream, int* out_error) {
    if (!x) return DMK_ERR_INVALID;
    dmk_context_t ctx = x->ctx;
    if (bind(ctx)) return DMK_ERR_CUDA;
    cudaStream_t s = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->stream;
    uint32_t local = 0, owner = 0;
    DMK_CUDA(ctx, cudaMemcpyAsync(&local, x->d_local + 1, sizeof local, cudaMemcpyDeviceToHost, s));
    if (x->kind == dmk_exchange_s::Owner)
        DMK_CUDA(ctx, cudaMemcpyAsync(&owner, &x->header()->error, sizeof owner, cudaMemcpyDeviceToHost, s));
    DMK_CUDA(ctx, cudaStreamSynchronize(s));
    const int err = local ? static_cast<int>(local) : static_cast<int>(owner);
    if (out_error) *out_error = err;
    if (err == static_cast<int>(kExchangeErrTimeoutProducer))
        return fail(ctx, DMK_ERR_CUDA, "palette exchange: this rank's push gave up waiting for the render rank to release a buffer");
    if (err) return fail(ctx, DMK_ERR_CUDA, "palette exchange: the render rank gave up waiting for a rank's frame");
    return DMK_OK;
}

// ---- stand-alone cull ----------------------------------------------------------------------------------------
dmk_status dmk_cull_dev(dmk_context_t ctx, const float* corners_host, const float* world_inverse_dev, const float* bbox_dev, int64_t n,
                        uint32_t* vis_bits_dev) {
    if (!ctx) return DMK_ERR_INVALID;
    if (!corners_host || n < 0 || (n > 0 && (!world_inverse_dev || !bbox_dev || !vis_bits_dev))) return fail(ctx, DMK_ERR_INVALID, "null argument");
    if (bind(ctx)) return DMK_ERR_CUDA;
    CullParams p{};
    p.world_inverse = world_inverse_dev;
    p.bbox = bbox_dev;
    p.n = n;
    p.vis_bits = vis_bits_dev;
    std::memcpy(p.corners, corners_host, sizeof p.corners);
    DMK_CUDA(ctx, launch_cull(p, ctx->stream));
    ++ctx->launches;
    return DMK_OK;
}

dmk_status dmk_cull_trs_dev(dmk_context_t ctx, const float* corners_host, const float* position_dev, const float* rotation_dev,
                            const float* scale_dev, const float* bbox_dev, int64_t n, uint32_t* vis_bits_dev, float* world_inverse_out_dev) {
    if (!ctx) return DMK_ERR_INVALID;
    if (!corners_host || n < 0 || (n > 0 && (!position_dev || !rotation_dev || !scale_dev || !bbox_dev || !vis_bits_dev)))
        return fail(ctx, DMK_ERR_INVALID, "null argument");
    if (bind(ctx)) return DMK_ERR_CUDA;
    CullParams p{};
    p.position = position_dev;
    p.rotation = rotation_dev;
    p.scale = scale_dev;
    p.world_inverse_out = world_inverse_out_dev;
    p.bbox = bbox_dev;
    p.n = n;
    p.vis_bits = vis_bits_dev;
    std::memcpy(p.corners, corners_host, sizeof p.corners);
    DMK_CUDA(ctx, launch_cull(p, ctx->stream));
    ++ctx->launches;
    return DMK_OK;
}

dmk_status dmk_cull_host(dmk_context_t ctx, const float* view_proj, float near_depth, const float* world_inverse, const float* bbox,
                         int64_t n, uint32_t* vis_bits) {
    if (!ctx) return DMK_ERR_INVALID;
    if (!view_proj || n < 0 || (n > 0 && (!world_inverse || !bbox || !vis_bits))) return fail(ctx, DMK_ERR_INVALID, "null argument");
    if (n == 0) return DMK_OK;
    if (bind(ctx)) return DMK_ERR_CUDA;
    const size_t words = static_cast<size_t>((n + 31) / 32);
    const size_t need = static_cast<size_t>(n) * 88 + words * 4 + 256;
    if (ctx->cull_scratch_bytes < need) {
        dev_free(ctx->cull_scratch);
        ctx->cull_scratch = nullptr;
        ctx->cull_scratch_bytes = 0;
        DMK_CUDA(ctx, cudaMalloc(&ctx->cull_scratch, need));
        ctx->cull_scratch_bytes = need;
    }
    float* d_wi = static_cast<float*>(ctx->cull_scratch);
    float* d_bb = d_wi + static_cast<size_t>(n) * 16;
    uint32_t* d_bits = reinterpret_cast<uint32_t*>(d_bb + static_cast<size_t>(n) * 6 + ((n * 6) % 2 ? 1 : 0));
    float corners[24];
    frustum_corners_glm(view_proj, near_depth, corners);
    DMK_CUDA(ctx, cudaMemcpyAsync(d_wi, world_inverse, static_cast<size_t>(n) * 64, cudaMemcpyHostToDevice, ctx->stream));
    DMK_CUDA(ctx, cudaMemcpyAsync(d_bb, bbox, static_cast<size_t>(n) * 24, cudaMemcpyHostToDevice, ctx->stream));
    if (dmk_status st = dmk_cull_dev(ctx, corners, d_wi, d_bb, n, d_bits)) return st;
    DMK_CUDA(ctx, cudaMemcpyAsync(vis_bits, d_bits, words * 4, cudaMemcpyDeviceToHost, ctx->stream));
    DMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DMK_OK;
}

}  // extern "C"
