/*
 * dmk.h -- C ABI of libdarmok_b200.so: darmok's per-frame character hot path on NVIDIA B200 (sm_100a).
 *
 * This is the drop-in boundary.  darmok has no C ABI of its own: the seam is the ozz backend translation
 * unit (src/skeleton_ozz.cpp, selected by DARMOK_BUILD_OZZ, CMakeLists.txt:366-377) plus FrustumCuller
 * (src/culling.cpp).  A replacement TU (darmok_b200/host/skeleton_b200.cpp) provides
 * the same C++ symbols and forwards to the entry points below.  Every entry point cites the reference
 * interface it replaces (paths relative to the darmok repository).
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types; no exceptions cross this boundary
 *   - every function returns dmk_status (0 = ok); dmk_last_error(ctx) holds the message, using the
 *     reference's own error strings where one exists (src/skeleton_ozz.cpp:39,212,436,607,1038)
 *   - matrices are column-major float[16] (glm::mat4); ozz data is right-handed, outputs that leave the
 *     animator are left-handed exactly as OzzUtils::convert does (src/skeleton_ozz.cpp:21-25)
 *   - a context is single-caller (the reference calls this path from the main thread only,
 *     src/app.cpp:1050-1081); work is asynchronous on the context's CUDA stream until dmk_sync or a getter
 *   - host input arrays are copied at the call; "_dev" arguments are device pointers owned by the caller
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with DMK_ERR_CUDA
 */
#ifndef DMK_H
#define DMK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define DMK_API __declspec(dllexport)
#else
#define DMK_API __attribute__((visibility("default")))
#endif

typedef int dmk_status;
enum {
    DMK_OK = 0,
    DMK_ERR_INVALID = 1,     /* bad argument */
    DMK_ERR_FORMAT = 2,      /* archive rejected ("archive does not contain the type", ...) */
    DMK_ERR_CUDA = 3,        /* CUDA runtime failure / no device */
    DMK_ERR_UNSUPPORTED = 4, /* valid in the reference but outside this build (see message) */
    DMK_ERR_JOB = 5          /* "error in the sampling job" / "blending job" / "model job" */
};

typedef struct dmk_context_s* dmk_context_t;
typedef struct dmk_skeleton_s* dmk_skeleton_t; /* replaces SkeletonImpl          src/detail/skeleton_ozz.hpp:19-28 */
typedef struct dmk_clip_s* dmk_clip_t;         /* replaces SkeletalAnimationImpl src/detail/skeleton_ozz.hpp:30-38 */
typedef struct dmk_clipset_s* dmk_clipset_t;   /* replaces SkeletalAnimationMap  include/darmok/skeleton.hpp:82 */
typedef struct dmk_armature_s* dmk_armature_t; /* replaces Armature              include/darmok/skeleton.hpp:369-384 */
typedef struct dmk_mesh_s* dmk_mesh_t;         /* skinned vertex streams of MeshData::exportData, src/mesh_core.cpp:323-365 */
typedef struct dmk_crowd_s* dmk_crowd_t;       /* N x (SkeletalAnimatorImpl + Skinnable + BoundingBox/Transform) */

/* ---- context ------------------------------------------------------------------------------------------- */
/* cuda_stream: a cudaStream_t owned by the caller (e.g. torch's current stream) or NULL to create one. */
DMK_API dmk_status dmk_context_create(int device, void* cuda_stream, dmk_context_t* out);
DMK_API void dmk_context_destroy(dmk_context_t ctx);
DMK_API const char* dmk_last_error(dmk_context_t ctx); /* ctx may be NULL: last creation error */
DMK_API dmk_status dmk_sync(dmk_context_t ctx);
DMK_API void* dmk_context_stream(dmk_context_t ctx);
DMK_API const char* dmk_version(void);
/* number of kernels this library has launched on the context since creation (bench.py gpu_launches) */
DMK_API uint64_t dmk_context_launch_count(dmk_context_t ctx);

/* Optional per-kernel device timing (CUDA events on the context's stream around each launch; used by bench.py
 * for the roofline line).  kernel: DMK_KERNEL_*; totals accumulate until reset; the query synchronises. */
enum { DMK_KERNEL_POSE = 0, DMK_KERNEL_CULL = 1, DMK_KERNEL_SKIN = 2, DMK_KERNEL_PACK = 3, DMK_KERNEL_ANIMATOR = 4, DMK_KERNEL_COUNT = 5 };
DMK_API dmk_status dmk_context_enable_kernel_timing(dmk_context_t ctx, int on);
DMK_API dmk_status dmk_context_kernel_time(dmk_context_t ctx, int kernel, double* total_ms, uint64_t* launches);
DMK_API dmk_status dmk_context_reset_kernel_timing(dmk_context_t ctx);

/* ---- assets -------------------------------------------------------------------------------------------- */
/* OzzSkeletonLoader::operator() -> OzzUtils::loadFromData<ozz::animation::Skeleton> (src/skeleton_ozz.cpp:27-44,
 * :191-200): bytes of an ozz archive holding an "ozz-skeleton" (version 2). */
DMK_API dmk_status dmk_skeleton_load_ozz(dmk_context_t ctx, const void* bytes, size_t size, dmk_skeleton_t* out);
DMK_API void dmk_skeleton_destroy(dmk_skeleton_t skel);
DMK_API int dmk_skeleton_num_joints(dmk_skeleton_t skel);                  /* ozz Skeleton::num_joints   (:57) */
DMK_API const char* dmk_skeleton_joint_name(dmk_skeleton_t skel, int i);   /* ozz Skeleton::joint_names  (:968) */
DMK_API int dmk_skeleton_joint_parent(dmk_skeleton_t skel, int i);         /* ozz Skeleton::joint_parents (:983) */
DMK_API int dmk_skeleton_find_joint(dmk_skeleton_t skel, const char* name); /* linear scan of :968-975; -1 if absent */

/* OzzSkeletalAnimationLoader::operator() (src/skeleton_ozz.cpp:207-216): "ozz-animation" archive (version 6). */
DMK_API dmk_status dmk_clip_load_ozz(dmk_context_t ctx, const void* bytes, size_t size, dmk_clip_t* out);
DMK_API void dmk_clip_destroy(dmk_clip_t clip);
DMK_API float dmk_clip_duration(dmk_clip_t clip);     /* SkeletalAnimation::getDuration (src/skeleton_ozz.cpp:116-119) */
DMK_API const char* dmk_clip_name(dmk_clip_t clip);   /* SkeletalAnimation::getName     (src/skeleton_ozz.cpp:111-114) */
DMK_API int dmk_clip_num_tracks(dmk_clip_t clip);

/* The clips an animator may play (SkeletalAnimator::AnimationMap, src/skeleton_ozz.cpp:893-899), re-laid out
 * into per-track SoA key arrays on the device.  Every clip must have the skeleton's SoA track count. */
DMK_API dmk_status dmk_clipset_create(dmk_context_t ctx, dmk_skeleton_t skel, const dmk_clip_t* clips, int num_clips,
                                      dmk_clipset_t* out);
DMK_API void dmk_clipset_destroy(dmk_clipset_t set);

/* Armature(std::vector<ArmatureJoint>) (include/darmok/skeleton.hpp:363-384): joint names + inverse bind poses.
 * The name -> skeleton joint lookup of getJointModelMatrix (src/skeleton_ozz.cpp:962-977) is resolved here once;
 * an unknown name keeps the reference behaviour (identity model matrix).  inv_bind must be affine (bottom row
 * 0,0,0,1), otherwise DMK_ERR_UNSUPPORTED.  num_joints + 1 must be <= 256 (reference cap: 64,
 * DARMOK_SKELETON_MAX_BONES include/darmok/skeleton.hpp:21-23). */
DMK_API dmk_status dmk_armature_create(dmk_context_t ctx, dmk_skeleton_t skel, const char* const* joint_names,
                                       const float* inv_bind, int num_joints, dmk_armature_t* out);
DMK_API void dmk_armature_destroy(dmk_armature_t arm);
DMK_API int dmk_armature_palette_size(dmk_armature_t arm); /* num_joints + 1 (slot 0 = identity, src/skeleton.cpp:238) */

/* Vertex streams in the layout MeshData::exportData writes (src/mesh_core.cpp:323-365, src/vertex.cpp:228-235):
 * position/normal/tangent float3; indices float4 (bone index into the armature, -1 = unused); weights float4. */
DMK_API dmk_status dmk_mesh_create(dmk_context_t ctx, const float* position, const float* normal, const float* tangent,
                                   const float* indices, const float* weights, int num_vertices, int palette_size,
                                   dmk_mesh_t* out);
/* Same with options.  DMK_MESH_SORT_BY_INFLUENCES stores (and skins) the vertices grouped by their number of non-zero
 * weights, so that warps skip the palette fetches of absent influences; the skinned buffer is then in the order given by
 * dmk_mesh_vertex_order (out[i] = index, in the caller's arrays, of the vertex at position i) and the caller remaps its
 * index buffer once, as after any vertex-cache optimisation.  Without the flag the caller's order is kept. */
enum {
    DMK_MESH_SORT_BY_INFLUENCES = 1,
    /* with DMK_MESH_SORT_BY_INFLUENCES: group for the 4-vertices-per-thread skinning shapes of
     * DMK_OPT_SKIN_PLANAR_OUTPUT instead of the default 8-vertices-per-thread kernel (either kernel is correct on either
     * grouping; the matching one skips more gathers) */
    DMK_MESH_SORT_BLOCKS_OF_4 = 2,
    /* Store the vertices clustered by bone (supersedes the two flags above).  The skinning kernel's threads own 8 consecutive
     * stored vertices; with this flag (almost) every such block shares one bone, which the thread then reads once per character
     * instead of once per vertex, absent influences are skipped by whole quarter-warps, and heavy and light blocks are dealt
     * evenly over the kernel's tiles and warps.  On a mesh with 1..4 random influences per vertex this halves the kernel's
     * shared-memory traffic (its limiter); on a real mesh, whose neighbouring vertices share bones anyway, it does at least as
     * well.  As with DMK_MESH_SORT_BY_INFLUENCES the skinned buffer is in the order dmk_mesh_vertex_order reports, and the 4
     * (bone, weight) pairs of a vertex may be summed in another order than the caller listed them (same result within 1 ulp of
     * the 4-term sum).  The reference has no counterpart: its vertex order is whatever the asset importer wrote
     * (src/mesh_core.cpp:323-365) and the GPU's vertex fetch does not care. */
    DMK_MESH_CLUSTER_BONES = 4
};
DMK_API dmk_status dmk_mesh_create_ex(dmk_context_t ctx, const float* position, const float* normal, const float* tangent,
                                      const float* indices, const float* weights, int num_vertices, int palette_size,
                                      uint32_t flags, dmk_mesh_t* out);
DMK_API dmk_status dmk_mesh_vertex_order(dmk_mesh_t mesh, int32_t* out_order);
/* What DMK_MESH_CLUSTER_BONES would do to a mesh, without a device: the stored order (out_order[num_vertices], may be NULL)
 * and out_stats[0] = 128-bit shared-memory load wavefronts per character the clustered order needs, [1] = the same for the
 * caller's order (4 influences x 3 rows per vertex), [2] = 8-vertex blocks that share a bone, [3] = blocks in total.
 * `indices` / `weights` as for dmk_mesh_create. */
DMK_API dmk_status dmk_mesh_cluster_stats(const float* indices, const float* weights, int num_vertices, int palette_size,
                                          int32_t* out_order, int64_t* out_stats4);
/* How DMK_MESH_CLUSTER_BONES stores a mesh of num_vertices vertices (no device needed): out_layout[0] = vertex pitch (what
 * dmk_mesh_vertex_pitch will report: num_vertices rounded up to 8, or to 64 for the layout below), [1] = 1 when the blocks are
 * STRIDED, [2] = vertices per tile of the skinning plan.  A block is what one thread of the skinning kernel owns.  Plain layout:
 * 8 consecutive stored vertices.  Strided layout (the kernel's bulk-store output path: a warp's round of vertices is one contiguous
 * piece of the output): inside a tile, the thread (warp w, lane l) owns positions tile_begin + 256 w + k nl + l, k = 0..7, with
 * nl = min(32, blocks of the tile - 32 w). */
DMK_API dmk_status dmk_mesh_cluster_layout(int num_vertices, int palette_size, int64_t* out_layout3);
DMK_API void dmk_mesh_destroy(dmk_mesh_t mesh);
DMK_API int dmk_mesh_num_vertices(dmk_mesh_t mesh);
DMK_API int dmk_mesh_vertex_pitch(dmk_mesh_t mesh); /* vertices per character in the skinned buffer (V rounded up to 8) */

/* ---- crowd: N characters sharing skeleton / clip set / armature / mesh ---------------------------------- */
/* armature and mesh may be NULL (palette-only or animation-only crowds). */
DMK_API dmk_status dmk_crowd_create(dmk_context_t ctx, dmk_skeleton_t skel, dmk_clipset_t clips, dmk_armature_t arm,
                                    dmk_mesh_t mesh, int64_t num_characters, int max_layers, dmk_crowd_t* out);
DMK_API void dmk_crowd_destroy(dmk_crowd_t crowd);
DMK_API int64_t dmk_crowd_size(dmk_crowd_t crowd);
/* Tuning / test switches.  DMK_OPT_FORCE_KEY_SEARCH: sample with the per-track binary search even when the clip
 * set carries the bucket table (both paths must give identical key indices). */
enum {
    DMK_OPT_FORCE_KEY_SEARCH = 1,
    /* Leave `value` SMs out of the persistent skinning grid so that a collective enqueued on another stream (the
     * visible-palette gather over NVLink) runs concurrently with the skinning kernel instead of after it.  The grid
     * otherwise takes every SM (its blocks fill the register file: nothing else fits beside them); the copy-engine
     * transport needs one SM for its one-thread flow-control kernels, NCCL eight. */
    DMK_OPT_SKIN_RESERVED_SMS = 2,
    /* With dmk_crowd_set_transforms: also store the world inverse the cull derives for each entity (64 B/entity more
     * written per CULL step) so that dmk_crowd_get_world_inverses / dmk_crowd_world_inverse_dev can hand it on, e.g. to
     * the renderer's per-entity uniforms. */
    DMK_OPT_KEEP_WORLD_INVERSE = 3,
    /* Layout of the skinned vertices.  0 (default): interleaved [N][vertex_pitch][9].  1 or 2: nine component planes per
     * character, [N][9][vertex_pitch] (plane 0..2 position.xyz, 3..5 normal.xyz, 6..8 tangent.xyz), for consumers that
     * pull vertices from buffers; the skinning kernel then keeps 4 vertices per thread and writes whole 128-byte lines
     * (1: up to 10 warps per SM, 2: up to 14).  3: three float3 streams per character, [N][3][vertex_pitch][3] (positions,
     * normals, tangents: the usual non-interleaved vertex-buffer layout), same thread shape as the default.  4: the nine planes of 1 / 2
     * written by the default kernel's thread shape (8 vertices per thread, one whole sector of every plane per lane): the store
     * layout alone, without the change of occupancy.  The host getters (dmk_crowd_get_skinned, dmk_crowd_get_skin_skinned)
     * return [count][V][9] whatever the device layout is.  DMK_ERR_UNSUPPORTED for palettes above 226 slots (181 for 1 and 2).  Measured on B200: DESIGN.md K5. */
    DMK_OPT_SKIN_PLANAR_OUTPUT = 4,
    /* Diagnostic: 1 = skin with the round-1 kernel shape (every warp replicates the palette and skins) instead of the
     * producer-warp / compute-warps shape, for A/B timing of the two on the same crowd.  Layouts 1 and 2 only exist in the
     * round-1 shape.  Results are identical either way. */
    DMK_OPT_SKIN_ROUND1_KERNEL = 5,
    /* 1 = sampling, local-to-model and palette in ONE kernel (a block owns 4 characters; the local matrices stay in shared
     * memory).  Measured on B200 (profiles/r02_pose_fused_kernel.txt, r02_pose_path_ab.txt): DRAM traffic of the pose path
     * drops from 3.8x to 0.92x its algorithmic bytes, but the path is latency bound, not bandwidth bound, and the block-wide
     * barriers between the phases cost more than the round trip saved: 742 us against 613 us per 100k characters.  Hence
     * opt-in; the default is two kernels.  Crowds with several armatures (dmk_crowd_add_skin) always take the two-kernel
     * path.  Results are bit-identical. */
    DMK_OPT_POSE_FUSED = 6,
    /* 1 = do not write the [N][K+1][16] column-major mat4 palette every step: only the 3x4 row-major palette the skinning
     * kernel reads is kept (half of the pose path's writes).  dmk_crowd_palette_dev then returns NULL; dmk_crowd_get_palette,
     * dmk_crowd_pack_visible_palettes and the exchange still deliver mat4 palettes (expanded from the 3x4 one on the way). */
    DMK_OPT_PALETTE_3X4_ONLY = 7,
    /* Skinning kernel, how the 8 palette replicas of a character reach shared memory: 0 (default) = eight bulk copies from
     * global memory (L2); 1 = one bulk copy from global memory + seven shared-memory -> shared-memory bulk copies (an eighth
     * of the L2 reads; the kernel is then launched as a thread-block cluster of one, which that copy needs).  Measured on
     * B200 (profiles/r02_skin_variants_call8.jsonl): the shared-memory copies move 15 B/cycle/SM and compete with the
     * gathers -- 4.16 ms against 3.52 ms -- so this stays a diagnostic.  Results are identical. */
    DMK_OPT_SKIN_REPLICATE_IN_SMEM = 8
};
DMK_API dmk_status dmk_crowd_set_option(dmk_crowd_t crowd, int option, int value);
/* Launch shape the skinning kernel would use for a mesh of `vertex_pitch` vertices (multiple of 8), a palette of
 * `palette_size` slots, `num_sms` SMs and output layout `layout` (DMK_OPT_SKIN_PLANAR_OUTPUT value): out[0] threads per
 * block, [1] vertex tiles per character, [2] vertices per tile, [3] persistent blocks, [4] palette replicas in shared
 * memory, [5] shared-memory bytes, [6] vertices per thread.  Needs no device (occupancy is taken as one block per SM then). */
DMK_API dmk_status dmk_skin_plan(int vertex_pitch, int palette_size, int num_sms, int layout, int64_t* out7);

/* Optional per-character outputs that only the host-side getters need (they cost HBM writes per step):
 * models -> dmk_crowd_get_models, locals -> dmk_crowd_get_locals, key indices -> dmk_crowd_get_key_indices. */
enum { DMK_OUT_MODELS = 1, DMK_OUT_LOCALS = 2, DMK_OUT_KEY_INDICES = 4 };
DMK_API dmk_status dmk_crowd_enable_outputs(dmk_crowd_t crowd, uint32_t outputs);

/* Layer state of every character, arrays of [num_layers][N] (layer-major): what the animator state machine
 * resolves each tick into BlendingJob::Layer{weight, transform} (src/skeleton_ozz.cpp:593-603, :709-714) with
 * transform = SamplingJob(clip, ratio) (:429-437).  speed = clip speed / state speed product used by the device
 * clip clock; loop = SkeletalAnimatorAnimation.loop (assets/protobuf/skeleton.proto:18-23), NULL = all looping.
 * threshold = BlendingJob::threshold (state: definition, :559; transition: bx::kFloatSmallest, :718).
 * Rest pose of the blend = layer 0 (:719; :600-603 when no clip sits at (0,0)). */
DMK_API dmk_status dmk_crowd_set_layers(dmk_crowd_t crowd, int num_layers, const int32_t* clip, const float* ratio,
                                        const float* weight, const float* speed, const uint8_t* loop, float threshold);
/* General form: a per-character pose program, what darmok's animator state machine resolves a tick into.
 * A character has up to two GROUPS of layers -- the previous and the current state of a transition
 * (OzzSkeletalAnimatorTransition::update, src/skeleton_ozz.cpp:682-726).  A group is a BlendingJob over its layers
 * (OzzSkeletalAnimatorState::update :556-608, threshold and rest-pose layer from the state definition) or the
 * pass-through of one layer (single-clip state :552-555/:617-624; early exit when a clip sits exactly at the blend
 * position :588-593; all clips of the state are still sampled).  top: DMK_TOP_GROUP0 = a plain state or a finished
 * transition; DMK_TOP_TRANSITION = 2-layer BlendingJob (weight0, weight1) = (1 - v, v), threshold bx::kFloatSmallest,
 * rest pose = group 0 (:709-719); DMK_TOP_REST_POSE = the load-time pose (LocalToModelJob over joint_rest_poses,
 * :901-916); DMK_TOP_SKIP = animator without state: matrices are left untouched (:1029-1032).
 * Layer arrays are [num_layers][N]; group 0 uses layers [0, num_layers0), group 1 the next num_layers1.
 * Ratios are host-resolved (do not combine with DMK_STEP_ADVANCE).  Copies are staged in pinned memory and async. */
typedef struct dmk_pose_program {
    uint8_t num_layers0, num_layers1;
    uint8_t rest0, rest1;   /* blend: rest-pose layer within the group; pass-through: the layer that is the pose */
    uint8_t mode0, mode1;   /* DMK_GROUP_* */
    uint8_t top;            /* DMK_TOP_* */
    uint8_t reserved;
    float threshold0, threshold1;
    float weight0, weight1;
} dmk_pose_program;
/* DMK_GROUP_ZERO: the group's pose is darmok's zero-initialised locals (translation, rotation and scale all 0) -- what a state
 * shows whose single clip (or the clip sitting exactly at the blend position) has never been sampled, because a non-looping
 * clip that runs past its end on its very first tick is never sampled at all (src/skeleton_ozz.cpp:411-428).  The group's
 * layer rows are ignored. */
enum { DMK_GROUP_BLEND = 0, DMK_GROUP_PASSTHROUGH = 1, DMK_GROUP_ZERO = 2 };
enum { DMK_TOP_GROUP0 = 0, DMK_TOP_TRANSITION = 1, DMK_TOP_REST_POSE = 2, DMK_TOP_SKIP = 3 };
DMK_API dmk_status dmk_crowd_set_programs(dmk_crowd_t crowd, int num_layers, const int32_t* clip, const float* ratio,
                                          const float* weight, const dmk_pose_program* programs);
/* per-frame variant for a host-side state machine: pinned staging + async copy of clip/ratio/weight only.  The caller's arrays may
 * be reused as soon as the call returns.  The copy runs on a stream of the crowd's own: behind the last step that read the layer
 * arrays (i.e. under the skinning kernel of the frame before, not behind it) and in front of the next dmk_crowd_step, which waits for
 * it -- the order of results is the order of the calls, as if everything were on the context's stream. */
DMK_API dmk_status dmk_crowd_update_layers(dmk_crowd_t crowd, int num_layers, const int32_t* clip, const float* ratio,
                                           const float* weight);

/* ---- "next" row: the animator state machine itself on the device ---------------------
 * Instead of resolving every SkeletalAnimator on the host each frame and uploading layers + programs, hand over the
 * SkeletalAnimatorDefinition once; DMK_STEP_ANIMATOR then ticks one animator per character on the GPU
 * (SkeletalAnimatorImpl::update / afterUpdate, src/skeleton_ozz.cpp:1006-1087; state update :535-610; transition
 * :682-726; play :783-856; calcBlendWeights src/skeleton.cpp:306-364) and writes the layer rows and pose programs the
 * POSE step consumes.  Per frame only commands (play / stop), blend positions and playback speeds cross the bus.
 * Field names follow assets/protobuf/skeleton.proto:18-53; names are resolved to indices by the caller. */
typedef struct dmk_anim_clip_def {   /* SkeletalAnimatorAnimation */
    int32_t clip;                    /* index into the clip set */
    float blend_x, blend_y;          /* blend_position */
    float speed;                     /* 0 counts as 1 (src/skeleton_ozz.cpp:398-404) */
    int32_t loop;
} dmk_anim_clip_def;
typedef struct dmk_anim_state_def {  /* SkeletalAnimatorState */
    int32_t first_clip, num_clips;   /* range in the clip-def array, 1..16 clips */
    int32_t blend_type;              /* 0 Cartesian, 1 Directional */
    float threshold;
    int32_t tween_easing;            /* assets/protobuf/easing.proto value */
    float tween_duration;
    int32_t next_state;              /* -1 = none */
    float speed;
} dmk_anim_state_def;
typedef struct dmk_anim_transition_def {   /* SkeletalAnimatorTransition; state -1 = any ("" in the proto) */
    int32_t src_state, dst_state;
    int32_t easing;
    float duration;
    float offset;
} dmk_anim_transition_def;
/* Listener events of one tick in the reference's order: play() command, then afterUpdate (src/skeleton_ozz.cpp:1044-1087) */
typedef struct dmk_anim_events {
    int16_t cmd_started;             /* onAnimatorStateStarted(state) raised by the play command, -1 = none */
    int16_t cmd_transition_started;  /* 1: onAnimatorTransitionStarted followed it */
    int16_t transition_finished;     /* 1: onAnimatorTransitionFinished + onAnimatorStateFinished(prev_finished) */
    int16_t prev_finished;
    int16_t looped;                  /* onAnimatorStateLooped(state), -1 = none */
    int16_t finished;                /* onAnimatorStateFinished(state), -1 = none */
    int16_t auto_started;            /* next_state: onAnimatorStateStarted(state), -1 = none */
    int16_t auto_transition_started;
} dmk_anim_events;
/* Needs max_layers >= 2 * (largest num_clips): a transition blends two states.  All animators start stopped. */
DMK_API dmk_status dmk_crowd_set_animator(dmk_crowd_t crowd, const dmk_anim_state_def* states, int num_states,
                                          const dmk_anim_clip_def* clips, int num_clips,
                                          const dmk_anim_transition_def* transitions, int num_transitions);
/* Commands executed at the start of the next DMK_STEP_ANIMATOR, per character: state >= 0 = play(state, speed),
 * DMK_ANIM_STOP = stop() (clears _state only: a running transition keeps going), DMK_ANIM_NONE = nothing,
 * DMK_ANIM_CLEAR = what play() of a state that is defined but cannot be created (none of its animations is known) leaves
 * behind: _state and _transition are both gone before State::create fails, no listener is called and play returns false
 * (src/skeleton_ozz.cpp:793-845).  speed NULL = 1 (play()'s default stateSpeed). */
enum { DMK_ANIM_CLEAR = -3, DMK_ANIM_NONE = -2, DMK_ANIM_STOP = -1 };
DMK_API dmk_status dmk_crowd_animator_commands(dmk_crowd_t crowd, const int32_t* state, const float* speed);
/* setBlendPosition / setPlaybackSpeed for every character: blend_xy [N][2] and/or speed [N] (NULL = unchanged) */
DMK_API dmk_status dmk_crowd_animator_set_inputs(dmk_crowd_t crowd, const float* blend_xy, const float* speed);
/* After a step: events [N]; status [N] = 0, or 1 where update() returned "error in the blending job" (pose kept) */
DMK_API dmk_status dmk_crowd_get_animator_events(dmk_crowd_t crowd, dmk_anim_events* out_events, int32_t* out_status);
/* getCurrentState / getCurrentTransition of every character (src/skeleton_ozz.cpp:858-889) */
typedef struct dmk_anim_state {
    int32_t state;            /* the running transition's current state, else _state; -1 = none */
    int32_t transition;       /* running transition's definition index, -1 = none (then _state is unset: C10) */
    int32_t previous_state;   /* the transition's previous state, -1 = none */
    float state_time;         /* getNormalizedTime() of each */
    float transition_time;
    float previous_time;
    float state_speed;        /* state speed factor: getDuration() = definition duration * speed (C2) */
    float previous_speed;
} dmk_anim_state;
DMK_API dmk_status dmk_crowd_get_animator_state(dmk_crowd_t crowd, dmk_anim_state* out_states);
/* The layer rows and programs currently on the device (whoever wrote them): [num_layers][N] each, programs [N] */
DMK_API dmk_status dmk_crowd_get_programs(dmk_crowd_t crowd, int num_layers, int32_t* out_clip, float* out_ratio,
                                          float* out_weight, dmk_pose_program* out_programs);

/* Transform::getWorldInverse (src/transform.cpp:286-309) + BoundingBox (assets/protobuf/shape.proto:7-10) of each
 * character entity: world_inverse [N][16], bbox [N][6] = min.xyz, max.xyz (local space).  An entity without a Transform
 * is culled against the untransformed frustum (src/culling.cpp:275-278): pass the identity.  FrustumCuller only visits
 * entities that have a BoundingBox (filter at src/culling.cpp:36-44); one without is never culled there, so a caller
 * ignores its visibility bit (there is no box value that means "always visible": a NaN box is culled, as in the reference,
 * because `sd <= 0` is false for NaN). */
DMK_API dmk_status dmk_crowd_set_instances(dmk_crowd_t crowd, const float* world_inverse, const float* bbox);

/* "Next" row: instead of precomputed world inverses, hand over what Transform holds --
 * position, rotation (quaternion x, y, z, w) and scale of each parent-less entity -- and let the cull kernel run
 * Transform::update itself (src/transform.cpp:286-309: glm::inverse(translate * mat4_cast * scale), glm operation
 * order).  40 B/instance instead of 64 B, and no host loop over 10M transforms.  position [N][3], rotation [N][4],
 * scale [N][3], bbox [N][6]. */
DMK_API dmk_status dmk_crowd_set_transforms(dmk_crowd_t crowd, const float* position, const float* rotation, const float* scale,
                                            const float* bbox);

/* Entities whose Transform has a parent (src/transform.cpp:300-305): num_nodes group transforms (position [M][3],
 * rotation [M][4], scale [M][3]) forming a forest through node_parent [M] (-1 = root; any order; chains up to 32
 * deep), and for each crowd entity the node its Transform hangs under, entity_parent [N] (-1 = none).  Every CULL
 * step first runs Transform::update over the nodes on the device (world = parent.world * local, worldInverse =
 * localInverse * parent.worldInverse, parents first), then each entity's worldInverse = localInverse *
 * node.worldInverse.  Needs dmk_crowd_set_transforms.  num_nodes = 0 detaches the crowd again. */
DMK_API dmk_status dmk_crowd_set_transform_nodes(dmk_crowd_t crowd, int num_nodes, const float* position, const float* rotation,
                                                 const float* scale, const int32_t* node_parent, const int32_t* entity_parent);
/* Transform::getWorldMatrix / getWorldInverse of the nodes after the last CULL step: [M][16] each (either may be NULL) */
DMK_API dmk_status dmk_crowd_get_transform_nodes(dmk_crowd_t crowd, float* out_world, float* out_world_inverse);
/* the entities' world inverses the last CULL step derived (dmk_crowd_set_transforms path, after
 * dmk_crowd_set_option(DMK_OPT_KEEP_WORLD_INVERSE, 1)): [count][16] */
DMK_API dmk_status dmk_crowd_get_world_inverses(dmk_crowd_t crowd, int64_t first, int64_t count, float* out);
DMK_API const float* dmk_crowd_world_inverse_dev(dmk_crowd_t crowd);

/* Camera::getViewProjectionMatrix (src/camera.cpp:196-199) -> Frustum{proj * view} (src/culling.cpp:268,
 * src/shape.cpp:1141-1163), computed on the host in glm operation order.  near_depth =
 * Math::getNormalizedNearDepth() (src/math.cpp:32-36): 0, or -1 with homogeneous depth. */
DMK_API dmk_status dmk_crowd_set_camera(dmk_crowd_t crowd, const float* view_proj, float near_depth);
/* same, from 8 precomputed frustum corners in Frustum::CornerType order (include/darmok/shape.hpp:372-383) */
DMK_API dmk_status dmk_crowd_set_frustum_corners(dmk_crowd_t crowd, const float* corners);

enum {
    DMK_STEP_ADVANCE = 1,   /* clip clocks: OzzSkeletalAnimatorAnimationState::update (src/skeleton_ozz.cpp:411-428) */
    DMK_STEP_POSE = 2,      /* SamplingJob + BlendingJob + LocalToModelJob + palette (a2, a5, a6, a7, a8) */
    DMK_STEP_CULL = 4,      /* FrustumCuller::updateCulled (src/culling.cpp:264-285) */
    DMK_STEP_SKIN = 8,      /* applySkinning for every vertex (darmok_skinning.inc.slang:5-28) */
    DMK_STEP_SKIN_VISIBLE_ONLY = 16, /* with CULL+SKIN: skin only characters FrustumCuller would not cull */
    DMK_STEP_READBACK_VISIBILITY = 32, /* with CULL: also enqueue the D2H copy of the bitmask + count into pinned memory,
                                         * right behind the cull (collect it with dmk_crowd_collect_visibility) */
    DMK_STEP_ANIMATOR = 64, /* tick the device animators (dmk_crowd_set_animator) before POSE; excludes ADVANCE */
    DMK_STEP_ALL = 1 | 2 | 4 | 8
};
/* Several skinned meshes per character.  In darmok one character is an entity with the SkeletalAnimator and any number
 * of descendant entities with Renderable + Skinnable, each with ITS OWN Armature (one per assimp mesh,
 * src/scene_assimp.cpp:187-193,696-721); beforeRenderEntity finds the animator by walking the transform parents
 * (src/skeleton.cpp:214-221) and builds one palette per Skinnable from the same joint matrices (:236-249).
 * dmk_crowd_create takes the first (armature, mesh) pair = skin 0; every further pair is added here and gets its own
 * palette (and skinned vertices when mesh is not NULL) in every POSE / SKIN step.  out_skin: its index (1, 2, ...). */
DMK_API dmk_status dmk_crowd_add_skin(dmk_crowd_t crowd, dmk_armature_t armature, dmk_mesh_t mesh, int* out_skin);
DMK_API int dmk_crowd_num_skins(dmk_crowd_t crowd);
/* palette / skinned vertices of one skin, laid out like dmk_crowd_palette_dev / dmk_crowd_skinned_dev (skin 0 = those) */
DMK_API const float* dmk_crowd_skin_palette_dev(dmk_crowd_t crowd, int skin);
DMK_API const float* dmk_crowd_skin_skinned_dev(dmk_crowd_t crowd, int skin);
DMK_API dmk_status dmk_crowd_get_skin_palette(dmk_crowd_t crowd, int skin, int64_t first, int64_t count, float* out);
DMK_API dmk_status dmk_crowd_get_skin_skinned(dmk_crowd_t crowd, int skin, int64_t first, int64_t count, float* out);

/* One frame for the whole crowd: replaces the per-entity loops SkeletalAnimationSceneComponent::update
 * (src/skeleton.cpp:159-184), FrustumCuller::update (src/culling.cpp:258-262) and, per renderable,
 * SkeletalAnimationRenderComponent::beforeRenderEntity (src/skeleton.cpp:223-251) + the LBS vertex stage. */
DMK_API dmk_status dmk_crowd_step(dmk_crowd_t crowd, float dt, uint32_t flags);

/* ---- results (device pointers stay valid until the crowd is destroyed) ---------------------------------- */
/* u_skinning per character: [N][palette_size][16] column-major mat4, slot 0 identity (src/skeleton.cpp:237-249) */
DMK_API const float* dmk_crowd_palette_dev(dmk_crowd_t crowd);
/* skinned vertices: [N][vertex_pitch][9] = position.xyz, normal.xyz, tangent.xyz (forward.slang:33-38); nine planes
 * [N][9][vertex_pitch] under DMK_OPT_SKIN_PLANAR_OUTPUT */
DMK_API const float* dmk_crowd_skinned_dev(dmk_crowd_t crowd);
/* visibility bitmask, bit (i % 32) of word (i / 32) set when character i is NOT culled */
DMK_API const uint32_t* dmk_crowd_visibility_dev(dmk_crowd_t crowd);
/* ascending ids of the visible characters and their count (device side; filled by a CULL step) */
DMK_API const int32_t* dmk_crowd_visible_ids_dev(dmk_crowd_t crowd);
DMK_API const int32_t* dmk_crowd_visible_count_dev(dmk_crowd_t crowd);

/* Host getters (synchronise the stream).  first/count select a character range. */
DMK_API dmk_status dmk_crowd_get_palette(dmk_crowd_t crowd, int64_t first, int64_t count, float* out);
/* SkeletalAnimator::getJointModelMatrix for every joint (src/skeleton_ozz.cpp:962-977): [count][J][16], left-handed */
DMK_API dmk_status dmk_crowd_get_models(dmk_crowd_t crowd, int64_t first, int64_t count, float* out);
/* blended local transforms as ozz SoaTransform: [count][num_soa][40] (state/transition getLocals, :617-624, :738-741) */
DMK_API dmk_status dmk_crowd_get_locals(dmk_crowd_t crowd, int64_t first, int64_t count, float* out);
/* ozz key-stream indices (left, right) chosen by the sampler: [count][num_layers][3][4*num_soa][2] int32,
 * component order translation, rotation, scale -- the bit-exact "keyframe indices" of the parity contract */
DMK_API dmk_status dmk_crowd_get_key_indices(dmk_crowd_t crowd, int64_t first, int64_t count, int32_t* out);
DMK_API dmk_status dmk_crowd_get_skinned(dmk_crowd_t crowd, int64_t first, int64_t count, float* out); /* [count][V][9] */
DMK_API dmk_status dmk_crowd_get_visibility(dmk_crowd_t crowd, uint32_t* out_bits, int64_t* out_visible_count);
/* SkeletalAnimator::getJointModelMatrixes(dir) (src/skeleton_ozz.cpp:979-1004), the matrices RenderableSkeleton draws its
 * debug bones with (src/skeleton.cpp:62-113): for every joint that has a child, Math::transform(jointPos,
 * glm::rotation(dir, normalize(childPos - jointPos)), vec3(|childPos - jointPos|)) towards its FIRST child (the
 * reference's map is keyed by the parent's name and emplace keeps the first).  dir [3] (the reference's default is
 * (1, 0, 0)); out [count][J][16], zero for leaves; out_has_bone [J] (may be NULL): 1 where the map has an entry.
 * Needs DMK_OUT_MODELS. */
DMK_API dmk_status dmk_crowd_get_bone_matrices(dmk_crowd_t crowd, int64_t first, int64_t count, const float* dir, float* out,
                                               uint8_t* out_has_bone);
/* Pipelined frame loops: result of the OLDEST step issued with DMK_STEP_READBACK_VISIBILITY that has not been collected
 * yet (at most two may be outstanding).  Waits only for that step's cull + copy, not for work enqueued after it, so the
 * next frame can already be running: FrustumCuller's set for frame k is read while frame k + 1 executes. */
DMK_API dmk_status dmk_crowd_collect_visibility(dmk_crowd_t crowd, uint32_t* out_bits, int64_t* out_visible_count);
/* current clip clocks after ADVANCE: [num_layers][N] */
DMK_API dmk_status dmk_crowd_get_ratios(dmk_crowd_t crowd, float* out_ratio, uint8_t* out_looped);

/* Gather the palettes of the visible characters (ascending id) into a caller-owned DEVICE buffer of
 * capacity_characters * palette_size * 16 floats; count is left on the device (dmk_crowd_visible_count_dev) and is NOT
 * clamped: a count above capacity_characters means that only the first capacity_characters palettes were packed.
 * This is the payload of the one multi-GPU exchange. */
DMK_API dmk_status dmk_crowd_pack_visible_palettes(dmk_crowd_t crowd, float* out_dev, int64_t capacity_characters);

/* ---- multi-GPU: the visible-palette exchange over peer memory (NVLink / NVSwitch) ----------------------- */
/* The reference is one process on one GPU: beforeRenderEntity hands the palette to the renderer as a uniform
 * (src/skeleton.cpp:223-251).  With the crowd sharded one process per GPU the palettes of the
 * characters FrustumCuller left visible have to reach the render GPU; this is that exchange, without NCCL: the render
 * rank owns a slab [2 buffers][world][capacity][palette_size][16] in its HBM, exports it with CUDA IPC, and every rank
 * packs AND transfers in one kernel, storing its visible palettes into slab[frame & 1][rank] over NVLink and publishing
 * {count, frame} behind a system-scope release.  The render rank waits on the device for a frame of all ranks
 * (dmk_exchange_wait), reads it, and releases the buffer (dmk_exchange_release); a producer may be at most two frames
 * ahead of the last release.  Frames are numbered 1, 2, 3, ... and every rank pushes every frame.  Every device-side
 * wait is bounded (default 30 s): on expiry the exchange gives up and dmk_exchange_status reports it.
 * cuda_stream arguments: the stream to enqueue on (NULL = the context's stream); the exchange usually runs on a side
 * stream, concurrently with the skinning kernel (DMK_OPT_SKIN_RESERVED_SMS leaves it SMs). */
typedef struct dmk_exchange_s* dmk_exchange_t;
#define DMK_EXCHANGE_HANDLE_BYTES 64
#define DMK_EXCHANGE_MAX_RANKS 64
/* render rank: allocates the slab; handle_out (DMK_EXCHANGE_HANDLE_BYTES) is what the other processes open */
DMK_API dmk_status dmk_exchange_create(dmk_context_t ctx, int world, int rank, int64_t capacity_characters, int palette_size,
                                       uint8_t* handle_out, dmk_exchange_t* out);
/* another process (any GPU of the box with peer access to the render GPU): maps the render rank's slab */
DMK_API dmk_status dmk_exchange_open(dmk_context_t ctx, int world, int rank, int64_t capacity_characters, int palette_size,
                                     const uint8_t* handle, dmk_exchange_t* out);
/* same process as the owner (several logical ranks on one GPU or several GPUs driven by one process): shares the mapping */
DMK_API dmk_status dmk_exchange_attach(dmk_exchange_t owner, dmk_context_t ctx, int rank, dmk_exchange_t* out);
DMK_API void dmk_exchange_destroy(dmk_exchange_t exchange); /* owner last: the others unmap first */
/* bound of every device-side wait, milliseconds (default 30000) */
DMK_API dmk_status dmk_exchange_set_timeout_ms(dmk_exchange_t exchange, int64_t milliseconds);
/* Pack + transfer: palettes of the crowd's visible characters (ascending id, at most capacity) -> slab[frame & 1][rank]
 * on the render GPU, then publish.  frame must be the previous frame of this exchange + 1.
 * num_ctas > 0: ONE kernel of that many 256-thread blocks packs and stores over NVLink (0 = 4 per SM); keep it within the
 *   SMs left to the side stream.
 * num_ctas < 0: copy-engine transport -- the pack kernel (4 * -num_ctas blocks) fills a local buffer, one cudaMemcpyAsync
 *   moves the rank's whole slab (capacity palettes, padding included: the host never learns the count) into the render
 *   GPU's memory on a DMA engine, and two single-thread kernels do the flow control and the publication.  No SM takes part
 *   in the transfer; size the capacity close to the expected visible set.
 * The published count is the rank's number of visible characters, NOT clamped to the capacity: when it is larger, only the
 * first `capacity` palettes were transferred, dmk_exchange_wait records it and dmk_exchange_status reports error 3. */
DMK_API dmk_status dmk_crowd_push_visible_palettes(dmk_crowd_t crowd, dmk_exchange_t exchange, uint32_t frame, int num_ctas,
                                                   void* cuda_stream);
/* The copy-engine transport in two halves, for frame loops that keep the exchange on a side stream: stage = the pack kernel
 * into the rank's local buffer, on the whole GPU (tens of microseconds), typically on the frame's own stream right after the
 * cull; send = flow control + the DMA copy + publication, typically on the side stream behind an event, while the
 * skinning kernel runs.  dmk_crowd_push_visible_palettes with num_ctas < 0 is both on one stream. */
DMK_API dmk_status dmk_crowd_stage_visible_palettes(dmk_crowd_t crowd, dmk_exchange_t exchange, void* cuda_stream);
DMK_API dmk_status dmk_exchange_send_staged(dmk_exchange_t exchange, dmk_crowd_t crowd, uint32_t frame, void* cuda_stream);
/* render rank: enqueue the device-side wait for `frame` of every rank; afterwards (stream order) the frame's palettes
 * and counts are readable */
DMK_API dmk_status dmk_exchange_wait(dmk_exchange_t exchange, uint32_t frame, void* cuda_stream);
/* render rank: device pointers of a frame's buffer: slab [world][capacity][palette_size][16], counts int32 [world] */
DMK_API dmk_status dmk_exchange_frame_dev(dmk_exchange_t exchange, uint32_t frame, float** out_slab_dev, int32_t** out_counts_dev);
/* render rank: the frame has been read; its buffer may be overwritten by frame + 2 */
DMK_API dmk_status dmk_exchange_release(dmk_exchange_t exchange, uint32_t frame, void* cuda_stream);
/* synchronises the given stream and reports expired waits: *out_error = 0 ok, 1 this rank's push gave up waiting for a
 * release, 2 the render rank gave up waiting for a rank (DMK_ERR_CUDA is returned as well), 3 (render rank) some rank's
 * visible set outgrew the capacity in a frame that was waited for: its characters beyond the capacity were dropped
 * (DMK_ERR_UNSUPPORTED; the exchange itself keeps working) */
DMK_API dmk_status dmk_exchange_status(dmk_exchange_t exchange, void* cuda_stream, int* out_error);

/* ---- stand-alone batched cull (config 5: instances without animation) ---------------------------------- */
/* world_inverse_dev [n][16], bbox_dev [n][6] device pointers; vis_bits_dev ceil(n/32) words. */
DMK_API dmk_status dmk_cull_dev(dmk_context_t ctx, const float* corners_host, const float* world_inverse_dev,
                                const float* bbox_dev, int64_t n, uint32_t* vis_bits_dev);
/* same, from position / rotation (x, y, z, w) / scale device arrays; world_inverse_out_dev (optional, [n][16]) receives
 * the matrices Transform::getWorldInverse would return */
DMK_API dmk_status dmk_cull_trs_dev(dmk_context_t ctx, const float* corners_host, const float* position_dev, const float* rotation_dev,
                                    const float* scale_dev, const float* bbox_dev, int64_t n, uint32_t* vis_bits_dev,
                                    float* world_inverse_out_dev);
/* host-buffer form (copies in and out; used by the e2e measurement and the parity tests) */
DMK_API dmk_status dmk_cull_host(dmk_context_t ctx, const float* view_proj, float near_depth, const float* world_inverse,
                                 const float* bbox, int64_t n, uint32_t* vis_bits);
/* Frustum::Frustum(mtx) on the host in glm order (src/shape.cpp:1141-1163): out corners[8][3] */
DMK_API dmk_status dmk_frustum_corners(const float* view_proj, float near_depth, float* out_corners);

#ifdef __cplusplus
}
#endif
#endif /* DMK_H */
