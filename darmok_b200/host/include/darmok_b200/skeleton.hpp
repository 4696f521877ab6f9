// darmok_b200/skeleton.hpp -- darmok's skeleton / animator / skinning API over the B200 C ABI (include/dmk.h).
//
// This header mirrors include/darmok/skeleton.hpp (paths relative to the darmok repository): same class names,
// method names, argument meaning and error strings for the character hot path, so that the translation unit
// src/skeleton_ozz.cpp can be replaced by darmok_b200/host/src/skeleton_b200.cpp:
//
//   Skeleton, SkeletalAnimation                      include/darmok/skeleton.hpp:30-56
//   ISkeletalAnimatorState / ISkeletalAnimatorTransition                        :60-80
//   ISkeletalAnimatorListener (6 callbacks)                                     :100-111
//   SkeletalAnimator (play/stop/pause/update/getJointModelMatrix ...)           :259-301
//   ArmatureJoint, Armature, Skinnable                                          :363-406
//   SkeletalAnimationSceneComponent::update (per-entity tick, batched here)     src/skeleton.cpp:159-184
//   SkeletalAnimationRenderComponent::beforeRenderEntity (palette)              src/skeleton.cpp:223-251
//   FrustumCuller::update / shouldEntityBeCulled                                src/culling.cpp:258-290
//
// The definitions are plain structs with the field names of assets/protobuf/skeleton.proto (the reference uses the
// generated protobuf classes; protobuf is not part of the hot path).  What stays on the host is control flow per
// animator (clocks, blend weights, tweens, transitions, listener events: src/skeleton_ozz.cpp:411-440, 491-610,
// 682-726, 783-889, 1044-1087, src/skeleton.cpp:306-364); everything per joint / per vertex runs on the device.
#pragma once

#include "easing.hpp"
#include "expected.hpp"
#include "glm_lite.hpp"
#include "optional_ref.hpp"

#include <cstdint>
#include <filesystem>
#include <functional>
#include <memory>
#include <optional>
#include <string>
#include <string_view>
#include <unordered_map>
#include <vector>

struct dmk_context_s;
struct dmk_skeleton_s;
struct dmk_clip_s;
struct dmk_clipset_s;
struct dmk_armature_s;
struct dmk_mesh_s;
struct dmk_crowd_s;

namespace darmok {

// ---- definitions (assets/protobuf/skeleton.proto:8-64) -----------------------------------------------------------
struct SkeletalAnimatorAnimationDefinition {
    std::string name;
    glm::vec2 blend_position{0.f, 0.f};
    float speed = 0.f;   // proto3 default; 0 is treated as 1 (src/skeleton_ozz.cpp:398-404)
    bool loop = false;   // the JSON reader defaults to true (src/skeleton.cpp:473-528)
};
struct SkeletalAnimatorTweenDefinition {
    EasingType easing = EasingType::Linear;
    float duration = 0.f;
};
struct SkeletalAnimatorStateDefinition {
    enum BlendType { Cartesian = 0, Directional = 1 };
    std::string name;
    std::vector<SkeletalAnimatorAnimationDefinition> animations;
    BlendType blend = Cartesian;
    float threshold = 0.f;  // 0 makes ozz's BlendingJob fail: "error in the blending job" (C7)
    SkeletalAnimatorTweenDefinition tween;
    std::string next_state;
    float speed = 1.f;
};
struct SkeletalAnimatorTransitionDefinition {
    std::string src_state, dst_state;
    SkeletalAnimatorTweenDefinition tween;
    float offset = 0.f;
};
struct SkeletalAnimatorDefinition {
    std::string skeleton_path, animation_pattern;
    std::vector<SkeletalAnimatorStateDefinition> states;
    std::vector<SkeletalAnimatorTransitionDefinition> transitions;
};

// ---- device context shared by the assets -------------------------------------------------------------------------
class DeviceContext final {
public:
    static expected<std::shared_ptr<DeviceContext>, std::string> create(int device = 0, void* cudaStream = nullptr) noexcept;
    ~DeviceContext() noexcept;
    dmk_context_s* handle() const noexcept { return _ctx; }
    std::string lastError() const noexcept;
private:
    explicit DeviceContext(dmk_context_s* ctx) noexcept : _ctx{ctx} {}
    dmk_context_s* _ctx;
};

// ---- assets --------------------------------------------------------------------------------------------------------
class Skeleton final {
public:
    // bytes of an ozz archive holding an ozz-skeleton (OzzSkeletonLoader, src/skeleton_ozz.cpp:191-200)
    static expected<std::shared_ptr<Skeleton>, std::string> load(const std::shared_ptr<DeviceContext>& ctx, const void* bytes, size_t size) noexcept;
    ~Skeleton() noexcept;
    std::string toString() const noexcept;          // "Skeleton <n> joints" (src/skeleton_ozz.cpp:54-58)
    int getNumJoints() const noexcept;
    std::string_view getJointName(int i) const noexcept;
    int getJointParent(int i) const noexcept;
    dmk_skeleton_s* handle() const noexcept { return _h; }
    const std::shared_ptr<DeviceContext>& context() const noexcept { return _ctx; }
private:
    Skeleton(std::shared_ptr<DeviceContext> ctx, dmk_skeleton_s* h) noexcept : _ctx{std::move(ctx)}, _h{h} {}
    std::shared_ptr<DeviceContext> _ctx;
    dmk_skeleton_s* _h;
};

class SkeletalAnimation final {
public:
    static expected<std::shared_ptr<SkeletalAnimation>, std::string> load(const std::shared_ptr<DeviceContext>& ctx, const void* bytes, size_t size) noexcept;
    ~SkeletalAnimation() noexcept;
    std::string toString() const noexcept;          // "SkeletalAnimation <name>"
    std::string_view getName() const noexcept;
    float getDuration() const noexcept;
    dmk_clip_s* handle() const noexcept { return _h; }
private:
    SkeletalAnimation(std::shared_ptr<DeviceContext> ctx, dmk_clip_s* h) noexcept : _ctx{std::move(ctx)}, _h{h} {}
    std::shared_ptr<DeviceContext> _ctx;
    dmk_clip_s* _h;
};

using SkeletalAnimationMap = std::unordered_map<std::string, std::shared_ptr<SkeletalAnimation>>;

// ILoader<T>::operator()(path) -> expected<shared_ptr<T>, string> (include/darmok/loader.hpp); bytes come from a data loader
using DataLoader = std::function<expected<std::vector<uint8_t>, std::string>(const std::filesystem::path&)>;
expected<std::vector<uint8_t>, std::string> loadFileData(const std::filesystem::path& path) noexcept;

// the loader interfaces SkeletalAnimator::load reaches through its load context (include/darmok/skeleton.hpp:58, :352;
// include/darmok/asset.hpp:47-48; include/darmok/scene_serialize.hpp:370-378)
class ISkeletonLoader {
public:
    virtual ~ISkeletonLoader() = default;
    virtual expected<std::shared_ptr<Skeleton>, std::string> operator()(const std::filesystem::path& path) noexcept = 0;
};
class ISkeletalAnimationLoader {
public:
    virtual ~ISkeletalAnimationLoader() = default;
    virtual expected<std::shared_ptr<SkeletalAnimation>, std::string> operator()(const std::filesystem::path& path) noexcept = 0;
};
class IAssetContext {   // the two getters of darmok's IAssetContext that this path uses
public:
    virtual ~IAssetContext() = default;
    virtual ISkeletonLoader& getSkeletonLoader() noexcept = 0;
    virtual ISkeletalAnimationLoader& getSkeletalAnimationLoader() noexcept = 0;
};
class IComponentLoadContext {   // ... and the one getter of IComponentLoadContext
public:
    virtual ~IComponentLoadContext() = default;
    virtual IAssetContext& getAssets() noexcept = 0;
};

class B200SkeletonLoader final : public ISkeletonLoader {   // replaces OzzSkeletonLoader
public:
    B200SkeletonLoader(std::shared_ptr<DeviceContext> ctx, DataLoader dataLoader = loadFileData) noexcept;
    expected<std::shared_ptr<Skeleton>, std::string> operator()(const std::filesystem::path& path) noexcept override;
private:
    std::shared_ptr<DeviceContext> _ctx;
    DataLoader _dataLoader;
};
class B200SkeletalAnimationLoader final : public ISkeletalAnimationLoader {   // replaces OzzSkeletalAnimationLoader
public:
    B200SkeletalAnimationLoader(std::shared_ptr<DeviceContext> ctx, DataLoader dataLoader = loadFileData) noexcept;
    expected<std::shared_ptr<SkeletalAnimation>, std::string> operator()(const std::filesystem::path& path) noexcept override;
private:
    std::shared_ptr<DeviceContext> _ctx;
    DataLoader _dataLoader;
};

// ConstSkeletalAnimatorDefinitionWrapper::loadAnimations (src/skeleton.cpp:535-562): the path of an animation is the
// definition's animation_pattern with every "*" replaced by the animation's name (the name itself without a pattern);
// animations that fail to load are left out of the map (a state without any known animation cannot be created).
std::filesystem::path animationPath(const SkeletalAnimatorDefinition& def, std::string_view name) noexcept;
SkeletalAnimationMap loadAnimations(const SkeletalAnimatorDefinition& def, ISkeletalAnimationLoader& loader) noexcept;

struct ArmatureJoint final {
    std::string name;
    glm::mat4 inverseBindPose{1.f};
};
// protobuf::Armature as a plain struct (assets/protobuf/skeleton.proto:8-16): a joint's inverse_bind_pose is a Mat4 message of
// four column vectors col1..col4 (assets/protobuf/math.proto:41-46), i.e. glm's column-major mat4
struct ArmatureJointDefinition final {
    std::string name;
    float inverse_bind_pose[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};   // [column][row]
};
struct ArmatureDefinition final {
    std::string name;
    std::vector<ArmatureJointDefinition> joints;
};
class Armature final {
public:
    using Definition = ArmatureDefinition;
    explicit Armature(const Definition& def) noexcept {   // src/skeleton.cpp:662-669
        _joints.reserve(def.joints.size());
        for (const auto& joint : def.joints) {
            ArmatureJoint j;
            j.name = joint.name;
            for (int c = 0; c < 4; ++c)
                for (int r = 0; r < 4; ++r) j.inverseBindPose[c][r] = joint.inverse_bind_pose[c][r];
            _joints.push_back(std::move(j));
        }
    }
    static Definition createDefinition() noexcept { return {}; }
    explicit Armature(std::vector<ArmatureJoint> joints) noexcept : _joints{std::move(joints)} {}
    const std::vector<ArmatureJoint>& getJoints() const noexcept { return _joints; }
private:
    std::vector<ArmatureJoint> _joints;
};

// vertex streams of one skinned mesh in MeshData::exportData layout (src/mesh_core.cpp:323-365)
struct SkinnedMeshData {
    std::vector<float> position, normal, tangent;   // 3 floats per vertex
    std::vector<float> indices, weights;             // 4 floats per vertex
    int numVertices() const noexcept { return static_cast<int>(position.size() / 3); }
};

// MeshData's per-vertex skinning input (include/darmok/mesh_core.hpp:148-162): any number of (bone, weight) pairs, bone =
// index into the mesh's bone list = Armature joint order (src/skeleton_assimp.cpp:1129-1150).
struct MeshDataWeight final {
    size_t boneIndex = 0;
    float value = 0.F;
};
struct MeshDataVertex final {
    glm::vec3 position{0.f};
    glm::vec3 normal{0.f, 1.f, 0.f};
    glm::vec3 tangent{0.f};
    std::vector<MeshDataWeight> weights;
};
// The influence packing of MeshData::exportData (src/mesh_core.cpp:336-356): weights sorted by descending value (std::sort,
// like the reference: the order of equal weights is the library's), weights <= 0 dropped, the first four kept;
// indices[j] = bone index as float, unused = -1; weights start as (1, 0, 0, 0), so a vertex without any positive weight
// keeps weight 1 on index -1 (= the shader's pass-through, darmok_skinning.inc.slang:21-27).
void packVertexInfluences(const std::vector<MeshDataWeight>& weights, float outIndices[4], float outWeights[4]) noexcept;
// position / normal / tangent and the packed influences of every vertex, as exportData writes them
SkinnedMeshData makeSkinnedMeshData(const std::vector<MeshDataVertex>& vertices) noexcept;

// One Skinnable of a character: the armature whose palette is built and (optionally) the mesh it skins.  darmok hangs one
// entity with Renderable + Skinnable per assimp mesh under the entity that owns the SkeletalAnimator
// (src/scene_assimp.cpp:187-193,696-721); shared by every animator of a batch, never copied per character.
struct SkinTarget {
    std::shared_ptr<Armature> armature;
    std::shared_ptr<const SkinnedMeshData> mesh;
};

class Skinnable {
public:
    Skinnable(const std::shared_ptr<Armature>& armature = nullptr) noexcept : _armature{armature} {}
    std::shared_ptr<Armature> getArmature() const noexcept { return _armature; }
    void setArmature(const std::shared_ptr<Armature>& armature) noexcept { _armature = armature; }
private:
    std::shared_ptr<Armature> _armature;
};

// ---- animator ------------------------------------------------------------------------------------------------------
class ISkeletalAnimatorState {
public:
    virtual ~ISkeletalAnimatorState() = default;
    virtual float getNormalizedTime() const = 0;
    virtual float getDuration() const = 0;
    virtual std::string_view getName() const = 0;
};
class ISkeletalAnimatorTransition {
public:
    virtual ~ISkeletalAnimatorTransition() = default;
    virtual float getNormalizedTime() const = 0;
    virtual float getDuration() const = 0;
    virtual const ISkeletalAnimatorState& getCurrentState() const = 0;
    virtual const ISkeletalAnimatorState& getPreviousState() const = 0;
};

class SkeletalAnimator;
class ISkeletalAnimatorListener {
public:
    virtual ~ISkeletalAnimatorListener() = default;
    virtual void onAnimatorDestroyed(SkeletalAnimator&) {}
    virtual void onAnimatorStateLooped(SkeletalAnimator&, std::string_view) {}
    virtual void onAnimatorStateFinished(SkeletalAnimator&, std::string_view) {}
    virtual void onAnimatorStateStarted(SkeletalAnimator&, std::string_view) {}
    virtual void onAnimatorTransitionFinished(SkeletalAnimator&) {}
    virtual void onAnimatorTransitionStarted(SkeletalAnimator&) {}
};

class ISkeletalAnimatorListenerFilter {   // include/darmok/skeleton.hpp:123-128
public:
    virtual ~ISkeletalAnimatorListenerFilter() = default;
    virtual bool operator()(const ISkeletalAnimatorListener& listener) const = 0;
};

enum class SkeletalAnimatorPlaybackState { Stopped, Playing, Paused };

namespace detail {
class AnimatorCore;   // host state machine: resolves a tick into a pose program (skeleton_b200.cpp)
struct PoseRequest;
}
class SkeletalAnimationSceneComponent;

class SkeletalAnimator final {
public:
    using Definition = SkeletalAnimatorDefinition;
    using AnimationMap = SkeletalAnimationMap;
    using PlaybackState = SkeletalAnimatorPlaybackState;

    SkeletalAnimator() noexcept;   // no skeleton, no animations, empty definition: to be filled by load()
    SkeletalAnimator(std::shared_ptr<Skeleton> skel, AnimationMap anims, Definition def) noexcept;
    ~SkeletalAnimator() noexcept;
    SkeletalAnimator(const SkeletalAnimator&) = delete;
    SkeletalAnimator& operator=(const SkeletalAnimator&) = delete;

    SkeletalAnimator& addListener(std::unique_ptr<ISkeletalAnimatorListener> listener) noexcept;
    SkeletalAnimator& addListener(ISkeletalAnimatorListener& listener) noexcept;
    bool removeListener(const ISkeletalAnimatorListener& listener) noexcept;
    size_t removeListeners(const ISkeletalAnimatorListenerFilter& filter) noexcept;   // number removed (src/skeleton_ozz.cpp:259-262)

    SkeletalAnimator& setPlaybackSpeed(float speed) noexcept;
    float getPlaybackSpeed() const noexcept;
    SkeletalAnimator& setBlendPosition(const glm::vec2& value) noexcept;
    const glm::vec2& getBlendPosition() const noexcept;
    const Definition& getDefinition() const noexcept;
    OptionalRef<const ISkeletalAnimatorState> getCurrentState() const noexcept;            // include/darmok/skeleton.hpp:283
    OptionalRef<const ISkeletalAnimatorTransition> getCurrentTransition() const noexcept;   // :284
    float getStateDuration(const std::string& name) const noexcept;

    bool play(std::string_view name, float speedFactor = 1.F) noexcept;
    void stop() noexcept;
    void pause() noexcept;
    PlaybackState getPlaybackState() noexcept;

    // skinning targets: the armatures (and optionally the meshes) whose palettes / skinned vertices the device produces,
    // one per Skinnable entity under the animator's entity.  setSkinning defines skin 0, addSkinning every further one
    // (returns its index); both must be called before the first update().
    expected<void, std::string> setSkinning(const std::shared_ptr<Armature>& armature, const SkinnedMeshData* mesh = nullptr) noexcept;
    expected<int, std::string> addSkinning(const std::shared_ptr<Armature>& armature, const SkinnedMeshData* mesh = nullptr) noexcept;
    int getSkinCount() const noexcept { return static_cast<int>(_skins.size()); }

    glm::mat4 getJointModelMatrix(const std::string& node) const noexcept;
    // debug-bone matrices keyed by joint name (src/skeleton_ozz.cpp:979-1004), what RenderableSkeleton::update consumes
    std::unordered_map<std::string, glm::mat4> getJointModelMatrixes(const glm::vec3& dir = {1.f, 0.f, 0.f}) const noexcept;
    std::vector<glm::mat4> getSkinningPalette(int skin = 0) const noexcept;   // what beforeRenderEntity uploads as u_skinning
    std::vector<float> getSkinnedVertices(int skin = 0) const noexcept;       // [V][9] position, normal, tangent

    // stand-alone use: ticks the state machine and runs a one-character device step
    expected<void, std::string> update(float deltaTime) noexcept;
    // SkeletalAnimator::load(def, ctxt) (src/skeleton_ozz.cpp:901-933) with the two loaders darmok's load context hands out:
    // skeleton from def.skeleton_path (its error is returned), animations through loadAnimations, then the animator is reset
    // (no state, no transition, speed 1, not paused, blend position (0, 0)) and shows the new skeleton's rest pose.  The
    // skinning targets stay; not for animators owned by a SkeletalAnimationSceneComponent.
    expected<void, std::string> load(const Definition& def, ISkeletonLoader& skeletons, ISkeletalAnimationLoader& animations) noexcept;
    // the reference signature (include/darmok/skeleton.hpp:297): the two loaders come from ctxt.getAssets()
    expected<void, std::string> load(const Definition& def, IComponentLoadContext& ctxt) noexcept;
    static Definition createDefinition() noexcept { return {}; }   // src/skeleton_ozz.cpp:956-959

private:
    friend class SkeletalAnimationSceneComponent;
    struct Device;
    std::shared_ptr<Skeleton> _skeleton;
    AnimationMap _animations;
    std::unique_ptr<detail::AnimatorCore> _core;
    std::unique_ptr<Device> _device;       // own one-character crowd (stand-alone mode)
    SkeletalAnimationSceneComponent* _batch = nullptr;
    int64_t _batchIndex = -1;
    std::vector<SkinTarget> _skins;        // [0] = setSkinning's; shared pointers, the data is not copied per animator
    mutable std::vector<glm::mat4> _modelsCache;
    mutable bool _modelsValid = false;
    mutable uint64_t _modelsFrame = 0;
    expected<void, std::string> ensureDevice() noexcept;
    // AnimatorExecution::Device: the state machine lives on the GPU; these mirror what was read back last
    struct DeviceViews;
    mutable std::unique_ptr<DeviceViews> _views;
    bool onDevice() const noexcept;
};

// Where the animators of a SkeletalAnimationSceneComponent tick.
//   Host    every SkeletalAnimator's state machine runs on the CPU each frame (animator_core.cpp) and the resolved pose
//           programs are uploaded: exact reference semantics call by call.
//   Device  the state machines run in one kernel (dmk_crowd_set_animator); play() / stop() /
//           setBlendPosition() / setPlaybackSpeed() are queued and take effect at the next update(), one command per
//           animator and frame (the last one wins); listener events of a frame are delivered at the end of update().
enum class AnimatorExecution { Host, Device };

// Batched replacement of SkeletalAnimationSceneComponent::update (src/skeleton.cpp:159-184), FrustumCuller
// (src/culling.cpp:258-290) and SkeletalAnimationRenderComponent (src/skeleton.cpp:223-251) for animators that share
// skeleton, clips, armature and mesh: one device step per frame instead of one virtual call per entity.
class SkeletalAnimationSceneComponent final {
public:
    SkeletalAnimationSceneComponent(std::shared_ptr<Skeleton> skel, SkeletalAnimator::AnimationMap anims, SkeletalAnimator::Definition def,
                                    std::shared_ptr<Armature> armature, const SkinnedMeshData* mesh = nullptr) noexcept;
    ~SkeletalAnimationSceneComponent() noexcept;

    // a further Skinnable of every character (its own armature, optionally its own mesh); before init().  Returns the skin
    // index SkeletalAnimator::getSkinningPalette / getSkinnedVertices and dmk_crowd_skin_*_dev take.
    expected<int, std::string> addSkin(std::shared_ptr<Armature> armature, const SkinnedMeshData* mesh = nullptr) noexcept;

    // creates the animators (entities); must be called once before update()
    expected<void, std::string> init(int64_t numAnimators, AnimatorExecution execution = AnimatorExecution::Host) noexcept;
    SkeletalAnimator& getAnimator(int64_t index) noexcept;
    int64_t size() const noexcept;

    // per-entity Transform::getWorldInverse + BoundingBox, camera view-projection (FrustumCuller inputs)
    expected<void, std::string> setInstances(const float* worldInverse, const float* bbox) noexcept;
    // alternative to setInstances: what each entity's Transform holds (position [N][3], rotation xyzw [N][4], scale
    // [N][3]); Transform::update (src/transform.cpp:286-309) then runs on the device inside the cull
    expected<void, std::string> setTransforms(const float* position, const float* rotation, const float* scale, const float* bbox) noexcept;
    // parents of those Transforms: numNodes group transforms forming a forest (nodeParent, -1 = root) and the node each
    // entity hangs under (entityParent [N], -1 = none); numNodes = 0 detaches
    expected<void, std::string> setTransformNodes(int numNodes, const float* position, const float* rotation, const float* scale,
                                                  const int32_t* nodeParent, const int32_t* entityParent) noexcept;
    expected<void, std::string> setCamera(const glm::mat4& viewProj, float normalizedNearDepth = 0.f) noexcept;

    expected<void, std::string> update(float deltaTime) noexcept;          // ISceneComponent::update
    bool shouldEntityBeCulled(int64_t index) noexcept;                      // ICameraComponent::shouldEntityBeCulled

    dmk_crowd_s* crowd() const noexcept;
private:
    friend class SkeletalAnimator;
    struct Impl;
    std::unique_ptr<Impl> _impl;
};

}  // namespace darmok
