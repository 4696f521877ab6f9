// device_types.cuh -- plain structs shared by the host side of the C ABI and the sm_100a kernels.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace dmk {

constexpr int kSoaFloats = 40;      // ozz SoaTransform: T.xyz R.xyzw S.xyz, 4 lanes each
constexpr int kMaxLayers = 32;      // layer state lives in the lanes of the character's warp
constexpr int kMaxPalette = 256;    // palette slots are packed as bytes in the skinning kernel
constexpr uint16_t kIdleTask = 0xffffu;

// Clip set re-laid out for per-track search: keys of one (clip, component, track) are contiguous and sorted.
struct ClipSetDev {
    const uint2* dir;        // [num_clips][3][P] : {first key, key count}
    const float* ratio[3];   // key ratios per component (T, R, S)
    const uint2* value[3];   // packed key values: T/S = 3 halves; R = 3 int16 + largest (2 bit) + sign (1 bit)
    const int32_t* stream[3];// index of the key in the ozz key stream (what ozz's sampling cache stores)
    const float* duration;   // [num_clips]
    int num_clips;
    int padded_tracks;       // P = 4 * num_soa
    // Bucket table (null when a clip has tracks too dense for it: the kernel then searches the tracks above).
    // Time is cut into G = 2^k buckets per clip such that a bucket holds at most one key boundary of any track, so
    // for t in bucket b the bracketing keys are (base-1, base) or (base, base+1) with base = first key > b/G.
    // Entries are laid out [clip][component][bucket][track]: the 32 lanes of a warp (consecutive tracks) read
    // consecutive entries -- one coalesced request per field instead of a per-lane binary search.
    const uint2* bucket_dir; // [num_clips] : {first entry, G}
    const float4* bucket_a;  // {ratio[base-1], ratio[base], ratio[base+1] or -1 when base is the last key, base as int bits}
    const uint4* bucket_b;   // {value[base-1], value[base]}
    const uint2* bucket_c;   // value[base+1]
};

// Per-character pose program: what the animator state machine resolves a tick into.  A character has up to two
// GROUPS of layers (the previous and the current state of a transition, src/skeleton_ozz.cpp:682-726); a group is
// either a BlendingJob over its layers (:556-608) or the pass-through of one of them (single-clip state :552-555,
// :617-624, or the early exit when a clip sits exactly at the blend position :588-593).
struct PoseProgram {
    uint8_t n0, n1;        // layers of group 0 = [0, n0), of group 1 = [n0, n0 + n1)
    uint8_t rest0, rest1;  // blend: index (within the group) of the rest-pose layer; pass-through: the layer to copy
    uint8_t mode0, mode1;  // kGroupBlend / kGroupPassthrough / kGroupZero
    uint8_t top;           // kTopGroup0 / kTopTransition / kTopRestPose / kTopSkip
    uint8_t reserved;
    float threshold0, threshold1;  // BlendingJob::threshold of each group
    float w0, w1;          // transition layer weights (1 - v, v); threshold is bx::kFloatSmallest
};
constexpr uint8_t kGroupBlend = 0, kGroupPassthrough = 1, kGroupZero = 2;
constexpr uint8_t kTopGroup0 = 0, kTopTransition = 1, kTopRestPose = 2, kTopSkip = 3;

struct SkeletonDev {
    const int16_t* parents;   // [J]
    const float* rest;        // [P][10] rest pose per joint: T.xyz R.xyzw S.xyz (ozz joint_rest_poses, AoS)
    const uint16_t* schedule; // [rounds][32] local-to-model tasks: joint | column << 12, kIdleTask = none
    int num_joints;           // J
    int padded_joints;        // P
    int num_soa;
    int rounds;
};

struct ArmatureDev {
    // The armature joints in the PROCESSING ORDER the host chose (dmk_armature_create): entry i is armature joint order[i], and the 32
    // entries a warp handles together name skeleton joints that differ mod 32 (their model matrices sit in different shared-memory banks)
    const int32_t* remap;     // [K] (skeleton joint & 0xffff, 0xffff = unknown name) | (armature joint order[i] << 16)
    const float* inv_bind;    // [K][12] affine 3x4 of armature joint order[i], column-major columns of 3 (c0 c1 c2 c3)
    int num_joints;           // K
};

struct PoseParams {
    ClipSetDev clips;
    SkeletonDev skel;
    ArmatureDev arm;
    int64_t n;
    int num_layers;
    // layer state, [L][N]
    const int32_t* clip;
    float* ratio;
    const float* weight;
    const float* speed;
    const uint8_t* loop;      // may be null (all loop)
    uint8_t* looped;
    float threshold;
    const PoseProgram* program;  // [N] or null: flat mode (one group of num_layers layers, uniform threshold)
    float dt;
    int advance;              // run the clip clocks before sampling
    // outputs (null = skip)
    float* palette;           // [N][K+1][16] column-major mat4
    float* palette34;         // [N][K+1][16] row-major 3x4 (rows of 4) + one zero row: 2 sectors per slot, for the skinning kernel
    float* models;            // [N][J][16] left-handed model matrices
    float* locals;            // [N][num_soa][40]
    int32_t* key_indices;     // [N][L][3][P][2]
    int32_t* error_flag;      // set to character index + 1 when a job would fail
    float* scratch;           // [N][P][16] local 3x4 matrices (sampling kernel -> model kernel)
    int fused;                // one kernel for sampling + local-to-model + palette (no scratch round trip)
};

struct SkinParams {
    const float* palette34;   // [N][PS][16] (3 rows used)
    const float* pos;         // [Vpad][3]
    const float* nrm;
    const float* tan;
    const float* weights;     // [Vpad][4]
    const uint32_t* slots;    // [Vpad] 4 palette slots packed as bytes
    float* out;               // [N][Vpad][9]
    const int32_t* char_ids;  // optional indirection (visible characters), null = identity
    const int32_t* char_count;// device count when char_ids != null
    int64_t n;
    int vpad;                 // vertices per character, multiple of 8
    int palette_size;         // PS
    int num_tiles;            // vertex tiles per character
    int tile_vertices;        // multiple of 8
    int replicate_in_smem;    // replicas 1.. are copied from replica 0 inside shared memory instead of from global memory (set from the plan)
    int num_stages;           // staging depth of the producer warp's bulk copies (set by launch_skin from the plan)
    int32_t* next_item;       // [num_tiles] counters of the warp-specialised kernel's dynamic character assignment (zeroed by launch_skin); null = static split
    int vertex_order_mode;    // 0 caller's order; 1 grouped by influence count: the gathers of absent influences are predicated off;
                              // 2 clustered by bone (DMK_MESH_CLUSTER_BONES): also the block's bone once per thread and character
                              // 3 clustered by bone in the STRIDED layout (cluster_mesh_strided): also the bulk-store output path
};

struct CullParams {
    // either world_inverse, or (position, rotation, scale) from which Transform::update's world inverse is derived
    const float* position;      // [n][3] or null
    const float* rotation;      // [n][4] quaternion x, y, z, w
    const float* scale;         // [n][3]
    float* world_inverse_out;   // [n][16] optional: the derived matrices, for other consumers
    const int32_t* entity_parent;       // [n] or null: group node each entity's Transform hangs under, -1 = none
    const float* node_world_inverse;    // [num_nodes][16] from transform_nodes_kernel
    const float* world_inverse; // [n][16]
    const float* bbox;          // [n][6]
    int64_t n;
    uint32_t* vis_bits;         // [ceil(n/32)]
    float corners[24];          // camera frustum corners (world space), Frustum::CornerType order
};

}  // namespace dmk

// ---- device-side animator: darmok's animator state machine, one thread per character --------
namespace dmk {

constexpr int kMaxStateClips = 16;

struct AnimClipDef {      // SkeletalAnimatorAnimation (assets/protobuf/skeleton.proto:18-23) resolved against the clip set
    int32_t clip;         // index into the clip set
    float blend_x, blend_y;
    float speed;          // 0 counts as 1 (src/skeleton_ozz.cpp:398-404)
    int32_t loop;
    float duration;       // clip duration (seconds)
};
struct AnimStateDef {     // SkeletalAnimatorState (:30-46)
    int32_t first_clip, num_clips;
    int32_t blend_type;   // 0 Cartesian, 1 Directional
    float threshold;
    int32_t tween_easing;
    float tween_duration;
    int32_t next_state;   // -1 = none
    float speed;
    float duration;       // OzzSkeletalAnimatorState::calcDuration (host, once per definition)
    int32_t rest_clip;    // last clip whose blend position is exactly (0, 0), else 0
    int32_t table;        // Directional states: offset into AnimatorParams::tables of [m] |blend_position| followed by
                          // [m][m] angle(blend_position j, blend_position i) at i * m + j; -1 otherwise
};
struct AnimTransitionDef {  // SkeletalAnimatorTransition (:48-53); -1 = any state
    int32_t src_state, dst_state;
    int32_t easing;
    float duration;
    float offset;
};

// Animator state lives in HBM as structure-of-arrays words so that a warp's accesses coalesce and a character only
// touches what its situation needs (a plain single-clip state reads ~8 words; a 9-clip transition ~60).
// One "slot" = OzzSkeletalAnimatorState + its OzzSkeletalAnimatorAnimationStates; every character has three:
// _state, and the transition's previous / current states.  Word w of slot s of character c: slots[(s * kSlotWords + w) * N + c].
enum AnimSlot { kSlotState = 0, kSlotPrev = 1, kSlotCur = 2, kNumSlots = 3 };
enum AnimSlotWord {
    kWDef = 0,        // int: state definition index, -1 = none
    kWClipLooped,     // uint: bit per clip
    kWBlendX, kWBlendY, kWOldBlendX, kWOldBlendY,
    kWNormalizedTime, kWTweenTime, kWSpeed,
    kWLooped,         // int
    kWClipTime,       // kMaxStateClips floats
    kSlotWords = kWClipTime + kMaxStateClips
};
struct AnimStateInfo {        // = dmk_anim_state
    int32_t state, transition, previous_state;
    float state_time, transition_time, previous_time, state_speed, previous_speed;
};
struct AnimEvents {          // listener events of one tick, in the reference's order (src/skeleton_ozz.cpp:783-856, 1044-1087)
    int16_t cmd_started;     // play() command: onAnimatorStateStarted(state), -1 = none
    int16_t cmd_transition_started;
    int16_t transition_finished;  // 1: onAnimatorTransitionFinished + onAnimatorStateFinished(prev_finished)
    int16_t prev_finished;
    int16_t looped;          // onAnimatorStateLooped(state), -1 = none
    int16_t finished;        // onAnimatorStateFinished(state), -1 = none
    int16_t auto_started;    // next_state play: onAnimatorStateStarted
    int16_t auto_transition_started;
};

struct AnimatorParams {
    const AnimStateDef* states;
    const AnimClipDef* clips;
    const AnimTransitionDef* transitions;
    int num_states, num_transitions;
    const float* tables;         // per-definition constants of calcBlendWeight (src/skeleton.cpp:306-344)
    uint32_t* slots;             // [kNumSlots][kSlotWords][N]
    int32_t* transition_def;     // [N] running transition's definition, -1 = none
    float* transition_time;      // [N] its normalized time
    int32_t* cmd_state;          // [N] pending command: -2 none, -1 stop, >= 0 play(state); reset by the kernel
    const float* cmd_speed;      // [N]
    const float* in_blend;       // [N][2] setBlendPosition
    const float* in_speed;       // [N] setPlaybackSpeed
    AnimEvents* events;          // [N]
    int32_t* status;             // [N] 0 ok, 1 "error in the blending job" (pose left untouched), written every tick
    int64_t n;
    int num_layers;              // rows of the layer arrays
    float dt;
    // outputs: the layer arrays and programs the pose kernels consume
    int32_t* clip;
    float* ratio;
    float* weight;
    PoseProgram* program;
};

// Visible-palette exchange over peer memory (exchange_kernel.cu).  The header sits at the start of the render GPU's
// allocation (kExchangeHeaderBytes reserved), the two slab buffers [2][world][capacity][palette][16] behind it.
constexpr int kExchangeMaxRanks = 64;
constexpr size_t kExchangeHeaderBytes = 4096;
constexpr uint32_t kExchangeErrTimeoutProducer = 1, kExchangeErrTimeoutOwner = 2, kExchangeErrOverflow = 3;
struct ExchangeHeader {
    uint32_t published[kExchangeMaxRanks];   // last frame rank r has completely stored (release, system scope)
    int32_t count[2][kExchangeMaxRanks];     // palettes stored by rank r for the frame that uses buffer b
    uint32_t consumed;                       // last frame the owner has finished reading
    uint32_t error;                          // kExchangeErr*: a bounded wait of the owner expired
    uint32_t overflow_frame;                 // last frame in which a rank had more visible characters than the slab holds (0: never)
};
static_assert(sizeof(ExchangeHeader) <= kExchangeHeaderBytes, "header outgrew its reservation");

}  // namespace dmk
