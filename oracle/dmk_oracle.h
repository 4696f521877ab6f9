/*
 * dmk_oracle.h -- CPU oracle for the darmok/ozz character hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is product code: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * liboracle.so.  The CUDA path in darmok_b200/ never includes, links or calls this.
 *
 * PARITY STATUS
 *   - cull (Frustum/Plane/BoundingBox): pinned by the reference's own KATs
 *     (test/src/shape_test.cpp:52-59 and :61-84), see tests/test_oracle_cull.py.  The matrix conventions of the
 *     Transform restatement (translate / mat4_cast / scale) are pinned by the third KAT there (:9-50).
 *   - sampling / blending / local-to-model / palette / LBS: "parity unpinned".
 *     The arithmetic lives in ozz-animation, fetched by the reference at configure time
 *     (CMake/ozzAnimationConfig.cmake:2-5, GIT_TAG master = unpinned) and ABSENT from
 *     /root/reference; the reference holds no test, fixture or .ozz file for this path.
 *     These functions restate the published ozz-animation 0.14 runtime algorithms
 *     (archive v2 skeleton / v6 animation) with the scalar "simd_ref" semantics
 *     (RcpEst = 1/x, RSqrtEst = 1/sqrt(x)), anchored on darmok's call sites which are cited
 *     function by function.
 *
 * All matrices are column-major float[16] (glm::mat4 / ozz::math::Float4x4 memory order).
 * SoaTransform = 40 floats: T.x[4] T.y[4] T.z[4] R.x[4] R.y[4] R.z[4] R.w[4] S.x[4] S.y[4] S.z[4].
 * Build with -ffp-contract=off (reference build is ISO C++ without -mfma: no contraction).
 */
#ifndef DMK_ORACLE_H
#define DMK_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_SOA_FLOATS 40
#define ORC_MAX_JOINTS 1024 /* ozz::animation::Skeleton::kMaxJoints */
#define ORC_NO_PARENT (-1)  /* ozz::animation::Skeleton::kNoParent */

/* ---- ozz archive objects (ozz io::IArchive >> Skeleton / Animation; src/skeleton_ozz.cpp:27-44) ---- */
typedef struct orc_skeleton {
    int num_joints;
    int num_soa;     /* (num_joints + 3) / 4 */
    char* names_buf; /* NUL separated, joint order */
    int names_len;
    char** names;    /* pointers into names_buf */
    int16_t* parents;
    float* rest;     /* SoaTransform[num_soa] */
} orc_skeleton;

typedef struct orc_f3key { /* ozz Float3Key (animation_keyframe.h, v6) */
    float ratio;
    uint16_t track;
    uint16_t value[3]; /* IEEE half */
} orc_f3key;

typedef struct orc_qkey { /* ozz QuaternionKey (v6): 13/2/1 bit-field in memory, expanded here */
    float ratio;
    uint16_t track;
    uint8_t largest;
    uint8_t sign;
    int16_t value[3]; /* format 0 (archive v6): signed 16 bit; format 1 (archive v7): 15-bit magnitudes 0 .. 32767 */
    uint8_t format;
} orc_qkey;

typedef struct orc_animation {
    float duration;
    int num_tracks; /* un-padded */
    int num_soa;    /* (num_tracks + 3) / 4 */
    char* name;
    int n_t, n_r, n_s;
    orc_f3key* t;
    orc_qkey* r;
    orc_f3key* s;
} orc_animation;

/* Returns NULL and fills err on failure.  Error strings follow src/skeleton_ozz.cpp:39. */
orc_skeleton* orc_skeleton_load(const uint8_t* bytes, size_t n, char* err, size_t errn);
void orc_skeleton_free(orc_skeleton* s);
int orc_skeleton_num_joints(const orc_skeleton* s);
int orc_skeleton_num_soa(const orc_skeleton* s);
const char* orc_skeleton_joint_name(const orc_skeleton* s, int i);
void orc_skeleton_parents(const orc_skeleton* s, int16_t* out);
void orc_skeleton_rest_pose(const orc_skeleton* s, float* out_soa);

orc_animation* orc_animation_load(const uint8_t* bytes, size_t n, char* err, size_t errn);
void orc_animation_free(orc_animation* a);
float orc_animation_duration(const orc_animation* a);
const char* orc_animation_name(const orc_animation* a);
int orc_animation_num_tracks(const orc_animation* a);
int orc_animation_num_soa(const orc_animation* a);
void orc_animation_key_counts(const orc_animation* a, int* nt, int* nr, int* ns);

/* ---- ozz SamplingJob (call site src/skeleton_ozz.cpp:429-437) ---- */
typedef struct orc_sampling_context orc_sampling_context;
orc_sampling_context* orc_sampling_context_create(int max_tracks); /* Context::Resize */
void orc_sampling_context_free(orc_sampling_context* c);
void orc_sampling_context_invalidate(orc_sampling_context* c);

/* Stateful cursor form (what ozz runs).  out_soa: SoaTransform[out_soa_count].
 * out_cache (optional): int32[3][2 * 4 * num_soa] = the ozz key-stream indices (left,right) per padded
 * track, translation then rotation then scale -- the "keyframe indices" of the parity contract.
 * Returns 1 on success, 0 when ozz's Validate() would fail. */
int orc_sample(const orc_animation* a, orc_sampling_context* c, float ratio, float* out_soa,
               int out_soa_count, int32_t* out_cache);
/* Stateless per-track search (what the GPU computes); same outputs. */
int orc_sample_stateless(const orc_animation* a, float ratio, float* out_soa, int out_soa_count,
                         int32_t* out_cache);

/* ---- ozz BlendingJob, non-additive, no joint masks (src/skeleton_ozz.cpp:556-608, :709-724) ---- */
int orc_blend(int num_layers, const float* const* layer_soa, const float* weights, const float* rest_soa,
              float threshold, int num_soa, float* out_soa);

/* ---- ozz LocalToModelJob with darmok's defaults (src/skeleton_ozz.cpp:911-915, :1034-1039) ---- */
int orc_local_to_model(const orc_skeleton* s, const float* locals_soa, float* models /* 16 * J */);

/* Math::flipHandedness (src/math.cpp:21-25): zNegate * m * zNegate with real glm mat4 products. */
void orc_flip_handedness(const float* m, float* out);
/* glm::mat4 * glm::mat4, glm::inverse(mat4), mat4 * vec4 (Appendix B of SURVEY.md) */
void orc_mat4_mul(const float* a, const float* b, float* out);
void orc_mat4_inverse(const float* m, float* out);

/* Skin palette (src/skeleton.cpp:236-249 + src/skeleton_ozz.cpp:962-977): out[0] = I,
 * out[1+k] = flip(models[remap[k]]) * inv_bind[k]; remap[k] < 0 (unknown joint name) -> identity model. */
void orc_palette(const float* models, int num_joints, const int32_t* remap, const float* inv_bind, int k,
                 float* out /* 16 * (k + 1) */);
/* name -> joint index by linear string compare (src/skeleton_ozz.cpp:968-975); -1 if absent */
int orc_skeleton_find_joint(const orc_skeleton* s, const char* name);

/* SkeletalAnimator::getJointModelMatrixes(dir) (src/skeleton_ozz.cpp:979-1004): the debug-bone matrix of every joint that has
 * a child (bone to its FIRST child, keyed by the parent's name); out 16 floats per joint, has[j] = 0 for leaves.
 * models are the right-handed ozz matrices. */
void orc_bone_matrices(const orc_skeleton* s, const float* models, const float* dir, float* out, uint8_t* has);

/* ---- LBS vertex shader (assets/shaders/core/darmok_skinning.inc.slang:5-28, forward.slang:33-38) ----
 * indices/weights are float4 per vertex; out = 9 floats per vertex: pos.xyz, normal.xyz, tangent.xyz. */
void orc_skin(const float* palette, int palette_size, const float* pos, const float* nrm, const float* tan,
              const float* indices, const float* weights, int num_vertices, float* out);

/* ---- culling (src/culling.cpp:264-285, src/shape.cpp) ---- */
/* Frustum::Frustum(mtx) (src/shape.cpp:1141-1163): 8 corners = inverse(mtx) * ndc corner, / w. */
void orc_frustum_corners(const float* view_proj, float near_depth, float* corners /* 8*3 */);
/* Frustum::getPlanes (src/shape.cpp:1235-1272): out[6][4] = normal.xyz, distance */
void orc_frustum_planes(const float* corners, float* out);
/* Plane::signedDistanceTo (src/shape.cpp:449-452) */
float orc_plane_signed_distance(const float* normal, float distance, const float* point);
/* Frustum(camFrust) *= world_inverse; canSee(bbox) (src/culling.cpp:274-282). bbox = min.xyz,max.xyz */
int orc_cull_visible(const float* corners, const float* world_inverse, const float* bbox);
/* batched; vis_bits = ceil(n/32) uint32 words, bit i%32 of word i/32 set when visible */
void orc_cull_batch(const float* corners, const float* world_inverse, const float* bbox, int64_t n,
                    uint32_t* vis_bits, int num_threads);

/* "next" row 8(f)-1: Transform::update (src/transform.cpp:286-309) without parent, then the cull on its world inverse */
void orc_transform_local_matrix(const float* position, const float* rotation_xyzw, const float* scale, float* out16);
void orc_transform_world_inverse(const float* position, const float* rotation_xyzw, const float* scale, float* out16);
void orc_cull_batch_trs(const float* corners, const float* position, const float* rotation, const float* scale,
                        const float* bbox, int64_t n, uint32_t* vis_bits, float* world_inverse_out, int num_threads);

/* Transform::update with parents (src/transform.cpp:300-305): world = parent.world * local, worldInverse = localInverse *
 * parent.worldInverse.  parent[i] = -1 for a root; any node order. */
void orc_transform_nodes(int num_nodes, const float* position, const float* rotation_xyzw, const float* scale, const int32_t* parent,
                         float* out_world, float* out_world_inverse);
/* entities whose Transform has a parent among those nodes (entity_parent[i] = -1: none) */
void orc_cull_batch_trs_parented(const float* corners, const float* position, const float* rotation, const float* scale,
                                 const float* bbox, int64_t n, const int32_t* entity_parent, const float* node_world_inverse,
                                 uint32_t* vis_bits, float* world_inverse_out, int num_threads);

/* ---- clip clock (OzzSkeletalAnimatorAnimationState::update, src/skeleton_ozz.cpp:411-440) ----
 * Returns 1 when the clip is to be re-sampled this tick, 0 when it holds (finished, non looping). */
int orc_clip_clock(float* normalized_time, int* looped, float dt, float clip_duration, float speed, int loop);

/* ---- whole-crowd reference step (CPU baseline; OpenMP over characters) ----
 * Per character c and layer l: clip[l*n+c], ratio[l*n+c] (already advanced), weight[l*n+c].
 * Stateful sampling contexts are owned by the crowd, as the reference keeps one per clip state.
 */
typedef struct orc_crowd orc_crowd;
orc_crowd* orc_crowd_create(const orc_skeleton* skel, const orc_animation* const* clips, int num_clips,
                            const int32_t* remap, const float* inv_bind, int k, int64_t n, int num_layers);
void orc_crowd_free(orc_crowd* c);
/* out_palette: 16*(k+1) floats per character (may be NULL); out_models: 16*J per character (may be NULL);
 * out_locals: 40*num_soa per character (may be NULL);
 * mesh (may be NULL pointers / 0 vertices): skinned 9 floats/vertex into out_skinned (n * V * 9).
 * Returns 0 on success, else the index+1 of the first failing character. */
int64_t orc_crowd_step(orc_crowd* c, const int32_t* clip, const float* ratio, const float* weight,
                       float threshold, float* out_locals, float* out_models, float* out_palette,
                       const float* pos, const float* nrm, const float* tan, const float* indices,
                       const float* weights, int num_vertices, float* out_skinned, int num_threads);

int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
