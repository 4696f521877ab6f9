// skeleton_b200.cpp -- the translation unit that takes the place of darmok's src/skeleton_ozz.cpp (+ the batched
// parts of src/skeleton.cpp and src/culling.cpp): the public skeleton API implemented over the C ABI of
// libdarmok_b200.so.  No arithmetic of the hot path happens here; this file owns handles, resolves the host state
// machine (animator_core.cpp) into pose programs and moves them through include/dmk.h.
#include "darmok_b200/skeleton.hpp"

#include "darmok_b200/animator_core.hpp"

#include "dmk.h"

#include <algorithm>
#include <cstring>
#include <fstream>

namespace darmok {

namespace {
constexpr int kMaxLayers = 32;

// PoseRequest of one animator -> slot `index` of the layer arrays [numLayers][n] + program
void writeRequest(const detail::PoseRequest& req, int64_t index, int64_t n, int32_t* clip, float* ratio, float* weight,
                  dmk_pose_program& prog) {
    std::memset(&prog, 0, sizeof prog);
    auto writeGroup = [&](const detail::GroupRequest& g, int first) {
        for (size_t l = 0; l < g.layers.size(); ++l) {
            const size_t at = static_cast<size_t>(first + static_cast<int>(l)) * static_cast<size_t>(n) + static_cast<size_t>(index);
            clip[at] = g.layers[l].clip;
            ratio[at] = g.layers[l].ratio;
            weight[at] = g.layers[l].weight;
        }
    };
    if (req.top == detail::PoseRequest::Skip) {
        prog.top = DMK_TOP_SKIP;
        return;
    }
    prog.top = req.top == detail::PoseRequest::Transition ? DMK_TOP_TRANSITION : DMK_TOP_GROUP0;
    prog.num_layers0 = static_cast<uint8_t>(req.group0.layers.size());
    prog.rest0 = static_cast<uint8_t>(req.group0.rest);
    prog.mode0 = req.group0.zero ? DMK_GROUP_ZERO : req.group0.passthrough ? DMK_GROUP_PASSTHROUGH : DMK_GROUP_BLEND;
    prog.threshold0 = req.group0.threshold;
    writeGroup(req.group0, 0);
    if (req.top == detail::PoseRequest::Transition) {
        prog.num_layers1 = static_cast<uint8_t>(req.group1.layers.size());
        prog.rest1 = static_cast<uint8_t>(req.group1.rest);
        prog.mode1 = req.group1.zero ? DMK_GROUP_ZERO : req.group1.passthrough ? DMK_GROUP_PASSTHROUGH : DMK_GROUP_BLEND;
        prog.threshold1 = req.group1.threshold;
        prog.weight0 = req.w0;
        prog.weight1 = req.w1;
        writeGroup(req.group1, static_cast<int>(req.group0.layers.size()));
    }
}
int layersOf(const detail::PoseRequest& req) {
    if (req.top == detail::PoseRequest::Skip) return 0;
    int n = static_cast<int>(req.group0.layers.size());
    if (req.top == detail::PoseRequest::Transition) n += static_cast<int>(req.group1.layers.size());
    return n;
}

// deterministic clip order of an AnimationMap: sorted by map key
std::vector<std::pair<std::string, std::shared_ptr<SkeletalAnimation>>> sortedClips(const SkeletalAnimationMap& anims) {
    std::vector<std::pair<std::string, std::shared_ptr<SkeletalAnimation>>> v(anims.begin(), anims.end());
    std::sort(v.begin(), v.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
    return v;
}
}  // namespace

// ---- DeviceContext ---------------------------------------------------------------------------------------------------
expected<std::shared_ptr<DeviceContext>, std::string> DeviceContext::create(int device, void* cudaStream) noexcept {
    dmk_context_t ctx = nullptr;
    if (dmk_context_create(device, cudaStream, &ctx) != DMK_OK) return unexpected<std::string>{dmk_last_error(nullptr)};
    return std::shared_ptr<DeviceContext>(new DeviceContext(ctx));
}
DeviceContext::~DeviceContext() noexcept { dmk_context_destroy(_ctx); }
std::string DeviceContext::lastError() const noexcept { return dmk_last_error(_ctx); }

// ---- Skeleton / SkeletalAnimation / loaders -----------------------------------------------------------------------------
expected<std::shared_ptr<Skeleton>, std::string> Skeleton::load(const std::shared_ptr<DeviceContext>& ctx, const void* bytes, size_t size) noexcept {
    dmk_skeleton_t h = nullptr;
    if (dmk_skeleton_load_ozz(ctx->handle(), bytes, size, &h) != DMK_OK) return unexpected<std::string>{ctx->lastError()};
    return std::shared_ptr<Skeleton>(new Skeleton(ctx, h));
}
Skeleton::~Skeleton() noexcept { dmk_skeleton_destroy(_h); }
std::string Skeleton::toString() const noexcept { return "Skeleton " + std::to_string(getNumJoints()) + " joints"; }
int Skeleton::getNumJoints() const noexcept { return dmk_skeleton_num_joints(_h); }
std::string_view Skeleton::getJointName(int i) const noexcept {
    const char* n = dmk_skeleton_joint_name(_h, i);
    return n ? std::string_view{n} : std::string_view{};
}
int Skeleton::getJointParent(int i) const noexcept { return dmk_skeleton_joint_parent(_h, i); }

expected<std::shared_ptr<SkeletalAnimation>, std::string> SkeletalAnimation::load(const std::shared_ptr<DeviceContext>& ctx, const void* bytes,
                                                                                  size_t size) noexcept {
    dmk_clip_t h = nullptr;
    if (dmk_clip_load_ozz(ctx->handle(), bytes, size, &h) != DMK_OK) return unexpected<std::string>{ctx->lastError()};
    return std::shared_ptr<SkeletalAnimation>(new SkeletalAnimation(ctx, h));
}
SkeletalAnimation::~SkeletalAnimation() noexcept { dmk_clip_destroy(_h); }
std::string SkeletalAnimation::toString() const noexcept { return "SkeletalAnimation " + std::string(getName()); }
std::string_view SkeletalAnimation::getName() const noexcept { return dmk_clip_name(_h); }
float SkeletalAnimation::getDuration() const noexcept { return dmk_clip_duration(_h); }

expected<std::vector<uint8_t>, std::string> loadFileData(const std::filesystem::path& path) noexcept {
    std::ifstream in(path, std::ios::binary);
    if (!in) return unexpected<std::string>{"could not open " + path.string()};
    std::vector<uint8_t> data((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    return data;
}
B200SkeletonLoader::B200SkeletonLoader(std::shared_ptr<DeviceContext> ctx, DataLoader dataLoader) noexcept
    : _ctx{std::move(ctx)}, _dataLoader{std::move(dataLoader)} {}
expected<std::shared_ptr<Skeleton>, std::string> B200SkeletonLoader::operator()(const std::filesystem::path& path) noexcept {
    auto data = _dataLoader(path);
    if (!data) return unexpected<std::string>{data.error()};
    return Skeleton::load(_ctx, data->data(), data->size());
}
B200SkeletalAnimationLoader::B200SkeletalAnimationLoader(std::shared_ptr<DeviceContext> ctx, DataLoader dataLoader) noexcept
    : _ctx{std::move(ctx)}, _dataLoader{std::move(dataLoader)} {}
expected<std::shared_ptr<SkeletalAnimation>, std::string> B200SkeletalAnimationLoader::operator()(const std::filesystem::path& path) noexcept {
    // OzzSkeletalAnimationLoader (src/skeleton_ozz.cpp:207-216): whatever went wrong -- the data loader, the tag, the archive -- the caller
    // gets this one text (unlike the skeleton loader, which passes loadFromData's own message on); the library's detail stays in
    // DeviceContext::lastError()
    if (auto data = _dataLoader(path))
        if (auto anim = SkeletalAnimation::load(_ctx, data->data(), data->size())) return anim;
    return unexpected<std::string>{"archive doesn't contain an animation"};
}

// ---- MeshData::exportData's influence packing (src/mesh_core.cpp:336-356) ---------------------------------------------------
void packVertexInfluences(const std::vector<MeshDataWeight>& weights, float outIndices[4], float outWeights[4]) noexcept {
    outWeights[0] = 1.f; outWeights[1] = 0.f; outWeights[2] = 0.f; outWeights[3] = 0.f;   // glm::vec4 weights{ 1, 0, 0, 0 }
    for (int k = 0; k < 4; ++k) outIndices[k] = -1.f;                                      // glm::vec4 indices{ -1 }
    int j = 0;
    auto weightVector = weights;
    std::sort(weightVector.begin(), weightVector.end(), [](auto& a, auto& b) { return a.value > b.value; });
    for (auto& weight : weightVector) {
        if (weight.value <= 0.f) continue;
        outIndices[j] = static_cast<float>(weight.boneIndex);
        outWeights[j] = weight.value;
        if (++j > 3) break;
    }
}

SkinnedMeshData makeSkinnedMeshData(const std::vector<MeshDataVertex>& vertices) noexcept {
    SkinnedMeshData out;
    const size_t n = vertices.size();
    out.position.reserve(n * 3); out.normal.reserve(n * 3); out.tangent.reserve(n * 3);
    out.indices.resize(n * 4); out.weights.resize(n * 4);
    for (size_t i = 0; i < n; ++i) {
        const MeshDataVertex& v = vertices[i];
        for (int k = 0; k < 3; ++k) {
            out.position.push_back(v.position[k]);
            out.normal.push_back(v.normal[k]);
            out.tangent.push_back(v.tangent[k]);
        }
        packVertexInfluences(v.weights, &out.indices[i * 4], &out.weights[i * 4]);
    }
    return out;
}

// ---- device objects of one (skeleton, clips, armature, mesh) combination --------------------------------------------------
struct SkeletalAnimator::Device {
    std::shared_ptr<DeviceContext> ctx;
    dmk_clipset_t clips = nullptr;
    dmk_armature_t armature = nullptr;
    dmk_mesh_t mesh = nullptr;
    dmk_crowd_t crowd = nullptr;
    int64_t n = 0;
    int paletteSize = 0, numVertices = 0;
    bool hasInstances = false, hasCamera = false;
    std::vector<int32_t> clip;
    std::vector<float> ratio, weight;
    std::vector<dmk_pose_program> programs;
    std::vector<uint32_t> visibility;
    bool visibilityValid = false;

    bool anyMesh = false;   // some skin has vertices: steps include DMK_STEP_SKIN

    ~Device() {
        dmk_crowd_destroy(crowd);
        for (auto& [arm, msh] : extraSkins) {
            dmk_mesh_destroy(msh);
            dmk_armature_destroy(arm);
        }
        dmk_mesh_destroy(mesh);
        dmk_armature_destroy(armature);
        dmk_clipset_destroy(clips);
    }

    // armature + mesh handles of the skins after the first (dmk_crowd_add_skin); destroyed after the crowd
    std::vector<std::pair<dmk_armature_t, dmk_mesh_t>> extraSkins;

    static bool makeArmature(dmk_context_t ctx, const std::shared_ptr<Skeleton>& skel, const Armature& armature, dmk_armature_t* out) noexcept {
        std::vector<const char*> names;
        std::vector<float> invBind;
        for (const auto& j : armature.getJoints()) {
            names.push_back(j.name.c_str());
            const float* m = glm::value_ptr(j.inverseBindPose);
            invBind.insert(invBind.end(), m, m + 16);
        }
        return dmk_armature_create(ctx, skel->handle(), names.data(), invBind.data(), static_cast<int>(names.size()), out) == DMK_OK;
    }
    static bool makeMesh(dmk_context_t ctx, const SkinnedMeshData& mesh, int paletteSize, dmk_mesh_t* out) noexcept {
        return dmk_mesh_create(ctx, mesh.position.data(), mesh.normal.data(), mesh.tangent.data(), mesh.indices.data(), mesh.weights.data(),
                               mesh.numVertices(), paletteSize, out) == DMK_OK;
    }

    static expected<std::unique_ptr<Device>, std::string> create(const std::shared_ptr<Skeleton>& skel, const SkeletalAnimationMap& anims,
                                                                 const std::vector<SkinTarget>& skins, int64_t n) noexcept {
        const std::shared_ptr<Armature> armature = skins.empty() ? nullptr : skins[0].armature;
        const SkinnedMeshData* mesh = skins.empty() ? nullptr : skins[0].mesh.get();
        auto dev = std::make_unique<Device>();
        dev->ctx = skel->context();
        dev->n = n;
        dmk_context_t ctx = dev->ctx->handle();
        auto err = [&]() { return unexpected<std::string>{dev->ctx->lastError()}; };
        std::vector<dmk_clip_t> handles;
        for (const auto& [name, anim] : sortedClips(anims)) handles.push_back(anim->handle());
        if (handles.empty()) return unexpected<std::string>{"animator has no animations"};
        if (dmk_clipset_create(ctx, skel->handle(), handles.data(), static_cast<int>(handles.size()), &dev->clips)) return err();
        if (armature) {
            if (!makeArmature(ctx, skel, *armature, &dev->armature)) return err();
            dev->paletteSize = dmk_armature_palette_size(dev->armature);
            if (mesh) {
                if (!makeMesh(ctx, *mesh, dev->paletteSize, &dev->mesh)) return err();
                dev->numVertices = mesh->numVertices();
            }
        }
        if (dmk_crowd_create(ctx, skel->handle(), dev->clips, dev->armature, dev->mesh, n, kMaxLayers, &dev->crowd)) return err();
        for (size_t k = 1; k < skins.size(); ++k) {
            if (!skins[k].armature) return unexpected<std::string>{"a skin needs an armature"};
            dev->extraSkins.emplace_back(nullptr, nullptr);
            auto& [arm, msh] = dev->extraSkins.back();
            if (!makeArmature(ctx, skel, *skins[k].armature, &arm)) return err();
            if (skins[k].mesh && !makeMesh(ctx, *skins[k].mesh, dmk_armature_palette_size(arm), &msh)) return err();
            if (dmk_crowd_add_skin(dev->crowd, arm, msh, nullptr)) return err();
            dev->anyMesh = dev->anyMesh || msh != nullptr;
        }
        dev->anyMesh = dev->anyMesh || dev->mesh != nullptr;
        if (dmk_crowd_enable_outputs(dev->crowd, DMK_OUT_MODELS)) return err();
        // SkeletalAnimatorImpl::load runs LocalToModelJob over the rest pose (src/skeleton_ozz.cpp:901-916, C13)
        dev->clip.assign(static_cast<size_t>(n), 0);
        dev->ratio.assign(static_cast<size_t>(n), 0.f);
        dev->weight.assign(static_cast<size_t>(n), 0.f);
        dev->programs.assign(static_cast<size_t>(n), dmk_pose_program{});
        for (auto& p : dev->programs) p.top = DMK_TOP_REST_POSE;
        if (dmk_crowd_set_programs(dev->crowd, 1, dev->clip.data(), dev->ratio.data(), dev->weight.data(), dev->programs.data()) ||
            dmk_crowd_step(dev->crowd, 0.f, DMK_STEP_POSE | (dev->anyMesh ? DMK_STEP_SKIN : 0)))
            return err();
        return dev;
    }

    // runs the programs currently held in the host arrays
    expected<void, std::string> run(int numLayers, bool cull) noexcept {
        // a zero blend threshold comes back as DMK_ERR_JOB "error in the blending job" (src/skeleton_ozz.cpp:605-608)
        if (dmk_crowd_set_programs(crowd, numLayers, clip.data(), ratio.data(), weight.data(), programs.data())) return unexpected<std::string>{ctx->lastError()};
        uint32_t flags = DMK_STEP_POSE;
        if (anyMesh) flags |= DMK_STEP_SKIN;
        if (cull && hasInstances && hasCamera) flags |= DMK_STEP_CULL;
        if (dmk_crowd_step(crowd, 0.f, flags)) return unexpected<std::string>{ctx->lastError()};
        visibilityValid = false;
        return {};
    }
};

// ---- SkeletalAnimationSceneComponent state (declared early: device-executed animators route through it) ------------------
struct SkeletalAnimationSceneComponent::Impl {
    std::shared_ptr<Skeleton> skeleton;
    SkeletalAnimator::AnimationMap anims;
    SkeletalAnimator::Definition def;
    std::vector<SkinTarget> skins;   // [0] = the constructor's armature / mesh
    std::unique_ptr<SkeletalAnimator::Device> device;
    std::vector<std::unique_ptr<SkeletalAnimator>> animators;
    std::vector<detail::PoseRequest> requests;
    std::vector<char> active;
    uint64_t frame = 0;   // bumped by every update(): invalidates the animators' cached matrices

    // AnimatorExecution::Device
    bool deviceAnimators = false;
    std::vector<std::string> stateNames;   // device state index -> name (states that can be created, in definition order)
    std::vector<float> stateDuration;      // calcDuration of each (without the state speed)
    std::vector<float> transitionDuration;
    std::vector<int32_t> cmd;
    std::vector<float> cmdSpeed, blend, speed;
    bool cmdDirty = false, inputsDirty = false;
    std::vector<dmk_anim_state> info;
    bool infoValid = false;
    std::vector<dmk_anim_events> events;
    std::vector<int32_t> status;

    int stateIndex(std::string_view name) const noexcept {
        for (size_t i = 0; i < stateNames.size(); ++i)
            if (stateNames[i] == name) return static_cast<int>(i);
        return -1;
    }
    const dmk_anim_state* stateOf(int64_t index) noexcept {
        if (!infoValid) {
            info.resize(static_cast<size_t>(device->n));
            if (dmk_crowd_get_animator_state(device->crowd, info.data()) != DMK_OK) return nullptr;
            infoValid = true;
        }
        return &info[static_cast<size_t>(index)];
    }
    expected<void, std::string> setupDeviceAnimators() noexcept;
};

struct SkeletalAnimator::DeviceViews {
    struct State final : ISkeletalAnimatorState {
        std::string name;
        float time = 0.f, duration = 0.f;
        float getNormalizedTime() const override { return time; }
        float getDuration() const override { return duration; }
        std::string_view getName() const override { return name; }
    };
    struct Transition final : ISkeletalAnimatorTransition {
        State current, previous;
        float time = 0.f, duration = 0.f;
        float getNormalizedTime() const override { return time; }
        float getDuration() const override { return duration; }
        const ISkeletalAnimatorState& getCurrentState() const override { return current; }
        const ISkeletalAnimatorState& getPreviousState() const override { return previous; }
    };
    State state;
    Transition transition;
};

// ---- SkeletalAnimator ---------------------------------------------------------------------------------------------------
namespace {
struct ListenerSet {
    std::vector<std::unique_ptr<ISkeletalAnimatorListener>> owned;
    std::vector<ISkeletalAnimatorListener*> all;
};
std::unordered_map<const SkeletalAnimator*, ListenerSet>& listenerRegistry() {
    static std::unordered_map<const SkeletalAnimator*, ListenerSet> reg;
    return reg;
}
void bindEvents(SkeletalAnimator& animator, detail::AnimatorCore& core) {
    // listeners are copied before a fan-out, as the reference does (_listeners.copy(), src/skeleton_ozz.cpp:810,1046)
    auto each = [&animator](auto fn) {
        auto itr = listenerRegistry().find(&animator);
        if (itr == listenerRegistry().end()) return;
        const auto copy = itr->second.all;
        for (auto* l : copy) fn(*l);
    };
    core.events.stateLooped = [&animator, each](std::string_view s) { each([&](auto& l) { l.onAnimatorStateLooped(animator, s); }); };
    core.events.stateFinished = [&animator, each](std::string_view s) { each([&](auto& l) { l.onAnimatorStateFinished(animator, s); }); };
    core.events.stateStarted = [&animator, each](std::string_view s) { each([&](auto& l) { l.onAnimatorStateStarted(animator, s); }); };
    core.events.transitionFinished = [&animator, each]() { each([&](auto& l) { l.onAnimatorTransitionFinished(animator); }); };
    core.events.transitionStarted = [&animator, each]() { each([&](auto& l) { l.onAnimatorTransitionStarted(animator); }); };
}
}  // namespace

SkeletalAnimator::SkeletalAnimator() noexcept : SkeletalAnimator(nullptr, AnimationMap{}, Definition{}) {}

SkeletalAnimator::SkeletalAnimator(std::shared_ptr<Skeleton> skel, AnimationMap anims, Definition def) noexcept
    : _skeleton{std::move(skel)}, _animations{std::move(anims)} {
    std::vector<detail::ClipInfo> clips;
    for (const auto& [name, anim] : sortedClips(_animations)) clips.push_back({name, anim->getDuration()});
    _core = std::make_unique<detail::AnimatorCore>(std::move(clips), std::move(def));
}

std::filesystem::path animationPath(const SkeletalAnimatorDefinition& def, std::string_view name) noexcept {
    if (def.animation_pattern.empty()) return std::string{name};
    std::string v = def.animation_pattern;   // StringUtils::replace(v, "*", name) (src/string.cpp:134-149): every occurrence
    size_t at = 0;
    while ((at = v.find('*', at)) != std::string::npos) {
        v.replace(at, 1, name);
        at += name.size();
    }
    return v;
}

SkeletalAnimationMap loadAnimations(const SkeletalAnimatorDefinition& def, ISkeletalAnimationLoader& loader) noexcept {
    SkeletalAnimationMap anims;
    for (const auto& state : def.states)
        for (const auto& anim : state.animations)
            if (auto result = loader(animationPath(def, anim.name))) anims.emplace(anim.name, *result);
    return anims;
}

expected<void, std::string> SkeletalAnimator::load(const Definition& def, IComponentLoadContext& ctxt) noexcept {
    auto& assets = ctxt.getAssets();   // src/skeleton_ozz.cpp:917-919
    return load(def, assets.getSkeletonLoader(), assets.getSkeletalAnimationLoader());
}
expected<void, std::string> SkeletalAnimator::load(const Definition& def, ISkeletonLoader& skeletons, ISkeletalAnimationLoader& animations) noexcept {
    if (_batch) return unexpected<std::string>{"this animator is owned by its SkeletalAnimationSceneComponent"};
    auto skel = skeletons(def.skeleton_path);
    if (!skel) return unexpected<std::string>{skel.error()};
    // load(skeleton, animations, def) (:901-917): reset() drops state, transition, speed, pause and blend position -- a new
    // state machine -- and the rest-pose LocalToModelJob is what a freshly created device crowd shows
    _skeleton = std::move(*skel);
    _animations = loadAnimations(def, animations);
    std::vector<detail::ClipInfo> clips;
    for (const auto& [name, anim] : sortedClips(_animations)) clips.push_back({name, anim->getDuration()});
    _core = std::make_unique<detail::AnimatorCore>(std::move(clips), def);
    bindEvents(*this, *_core);
    _device.reset();
    _views.reset();
    _modelsValid = false;
    return {};
}

SkeletalAnimator::~SkeletalAnimator() noexcept {
    auto itr = listenerRegistry().find(this);
    if (itr != listenerRegistry().end()) {
        const auto copy = itr->second.all;
        for (auto* l : copy) l->onAnimatorDestroyed(*this);   // src/skeleton_ozz.cpp:233-242
        listenerRegistry().erase(this);
    }
}

SkeletalAnimator& SkeletalAnimator::addListener(std::unique_ptr<ISkeletalAnimatorListener> listener) noexcept {
    auto& set = listenerRegistry()[this];
    set.all.push_back(listener.get());
    set.owned.push_back(std::move(listener));
    bindEvents(*this, *_core);
    return *this;
}
SkeletalAnimator& SkeletalAnimator::addListener(ISkeletalAnimatorListener& listener) noexcept {
    listenerRegistry()[this].all.push_back(&listener);
    bindEvents(*this, *_core);
    return *this;
}
bool SkeletalAnimator::removeListener(const ISkeletalAnimatorListener& listener) noexcept {
    auto itr = listenerRegistry().find(this);
    if (itr == listenerRegistry().end()) return false;
    auto& all = itr->second.all;
    const auto before = all.size();
    all.erase(std::remove(all.begin(), all.end(), &listener), all.end());
    auto& owned = itr->second.owned;
    owned.erase(std::remove_if(owned.begin(), owned.end(), [&](const auto& p) { return p.get() == &listener; }), owned.end());
    return all.size() != before;
}

size_t SkeletalAnimator::removeListeners(const ISkeletalAnimatorListenerFilter& filter) noexcept {
    auto itr = listenerRegistry().find(this);
    if (itr == listenerRegistry().end()) return 0;
    auto& all = itr->second.all;
    std::vector<ISkeletalAnimatorListener*> gone;
    for (auto* l : all)
        if (filter(*l)) gone.push_back(l);
    all.erase(std::remove_if(all.begin(), all.end(), [&](auto* l) { return std::find(gone.begin(), gone.end(), l) != gone.end(); }), all.end());
    auto& owned = itr->second.owned;   // after the filter has looked at them
    owned.erase(std::remove_if(owned.begin(), owned.end(), [&](const auto& p) { return std::find(gone.begin(), gone.end(), p.get()) != gone.end(); }), owned.end());
    return gone.size();
}

bool SkeletalAnimator::onDevice() const noexcept { return _batch && _batch->_impl->deviceAnimators; }

SkeletalAnimator& SkeletalAnimator::setPlaybackSpeed(float speed) noexcept {
    _core->setPlaybackSpeed(speed);
    if (onDevice()) {
        _batch->_impl->speed[static_cast<size_t>(_batchIndex)] = speed;
        _batch->_impl->inputsDirty = true;
    }
    return *this;
}
float SkeletalAnimator::getPlaybackSpeed() const noexcept { return _core->getPlaybackSpeed(); }
SkeletalAnimator& SkeletalAnimator::setBlendPosition(const glm::vec2& value) noexcept {
    _core->setBlendPosition(value);
    if (onDevice()) {
        _batch->_impl->blend[static_cast<size_t>(_batchIndex) * 2] = value.x;
        _batch->_impl->blend[static_cast<size_t>(_batchIndex) * 2 + 1] = value.y;
        _batch->_impl->inputsDirty = true;
    }
    return *this;
}
const glm::vec2& SkeletalAnimator::getBlendPosition() const noexcept { return _core->getBlendPosition(); }
const SkeletalAnimator::Definition& SkeletalAnimator::getDefinition() const noexcept { return _core->getDefinition(); }
OptionalRef<const ISkeletalAnimatorState> SkeletalAnimator::getCurrentState() const noexcept {
    if (!onDevice()) return _core->getCurrentState();
    auto& im = *_batch->_impl;
    const dmk_anim_state* st = im.stateOf(_batchIndex);
    if (!st || st->state < 0) return nullptr;
    if (!_views) _views = std::make_unique<DeviceViews>();
    auto& v = _views->state;
    v.name = im.stateNames[static_cast<size_t>(st->state)];
    v.time = st->state_time;
    v.duration = im.stateDuration[static_cast<size_t>(st->state)] * st->state_speed;   // multiplies (C2)
    return &v;
}
OptionalRef<const ISkeletalAnimatorTransition> SkeletalAnimator::getCurrentTransition() const noexcept {
    if (!onDevice()) return _core->getCurrentTransition();
    auto& im = *_batch->_impl;
    const dmk_anim_state* st = im.stateOf(_batchIndex);
    if (!st || st->transition < 0) return nullptr;
    if (!_views) _views = std::make_unique<DeviceViews>();
    auto& t = _views->transition;
    t.time = st->transition_time;
    t.duration = im.transitionDuration[static_cast<size_t>(st->transition)];
    t.current.name = im.stateNames[static_cast<size_t>(st->state)];
    t.current.time = st->state_time;
    t.current.duration = im.stateDuration[static_cast<size_t>(st->state)] * st->state_speed;
    t.previous.name = im.stateNames[static_cast<size_t>(st->previous_state)];
    t.previous.time = st->previous_time;
    t.previous.duration = im.stateDuration[static_cast<size_t>(st->previous_state)] * st->previous_speed;
    return &t;
}
float SkeletalAnimator::getStateDuration(const std::string& name) const noexcept { return _core->getStateDuration(name); }
bool SkeletalAnimator::play(std::string_view name, float speedFactor) noexcept {
    if (!onDevice()) return _core->play(name, speedFactor);
    auto& im = *_batch->_impl;
    const int state = im.stateIndex(name);
    const size_t i = static_cast<size_t>(_batchIndex);
    if (state < 0) {
        // an unknown name does nothing (C16); a state that IS defined but has no known animation cannot be created, and the
        // reference has dropped _state and _transition by the time it finds out (src/skeleton_ozz.cpp:793-845)
        const auto& states = im.def.states;
        if (std::any_of(states.begin(), states.end(), [&](const auto& s) { return s.name == name; })) {
            im.cmd[i] = DMK_ANIM_CLEAR;
            im.cmdDirty = true;
        }
        return false;
    }
    int current = im.cmd[i] >= 0 ? im.cmd[i] : -1;   // a command queued this frame counts as the state being played
    if (current < 0)
        if (const dmk_anim_state* st = im.stateOf(_batchIndex)) current = (im.cmd[i] == DMK_ANIM_STOP && st->transition < 0) ? -1 : st->state;
    im.cmd[i] = state;
    im.cmdSpeed[i] = speedFactor;
    im.cmdDirty = true;
    return current != state;   // playing the current state again only changes its speed and returns false (:789-794)
}
void SkeletalAnimator::stop() noexcept {
    if (!onDevice()) return _core->stop();
    _batch->_impl->cmd[static_cast<size_t>(_batchIndex)] = DMK_ANIM_STOP;
    _batch->_impl->cmdDirty = true;
}
void SkeletalAnimator::pause() noexcept { _core->pause(); }
SkeletalAnimator::PlaybackState SkeletalAnimator::getPlaybackState() noexcept {
    if (!onDevice()) return _core->getPlaybackState();
    // src/skeleton_ozz.cpp:880-889: Stopped without _state -- which is also the case while a transition runs (C10)
    const dmk_anim_state* st = _batch->_impl->stateOf(_batchIndex);
    if (!st || st->transition >= 0 || st->state < 0) return PlaybackState::Stopped;
    return _core->isPaused() ? PlaybackState::Paused : PlaybackState::Playing;
}

expected<void, std::string> SkeletalAnimator::setSkinning(const std::shared_ptr<Armature>& armature, const SkinnedMeshData* mesh) noexcept {
    if (_device || _batch) return unexpected<std::string>{"setSkinning must be called before the first update"};
    if (_skins.empty()) _skins.emplace_back();
    _skins[0].armature = armature;
    _skins[0].mesh = mesh ? std::make_shared<const SkinnedMeshData>(*mesh) : nullptr;
    return {};
}

expected<int, std::string> SkeletalAnimator::addSkinning(const std::shared_ptr<Armature>& armature, const SkinnedMeshData* mesh) noexcept {
    if (_device || _batch) return unexpected<std::string>{"addSkinning must be called before the first update"};
    if (_skins.empty() || !_skins[0].armature) return unexpected<std::string>{"call setSkinning for the first armature"};
    if (!armature) return unexpected<std::string>{"a skin needs an armature"};
    _skins.push_back({armature, mesh ? std::make_shared<const SkinnedMeshData>(*mesh) : nullptr});
    return static_cast<int>(_skins.size()) - 1;
}

expected<void, std::string> SkeletalAnimator::ensureDevice() noexcept {
    if (_device || _batch) return {};
    if (!_skeleton) return unexpected<std::string>{"the animator has no skeleton (default constructed and not loaded)"};
    auto dev = Device::create(_skeleton, _animations, _skins, 1);
    if (!dev) return unexpected<std::string>{dev.error()};
    _device = std::move(*dev);
    return {};
}

expected<void, std::string> SkeletalAnimator::update(float deltaTime) noexcept {
    if (_batch) return unexpected<std::string>{"this animator is ticked by its SkeletalAnimationSceneComponent"};
    if (auto r = ensureDevice(); !r) return r;
    const bool active = _core->hasActiveStateOrTransition();
    detail::PoseRequest req;
    if (auto r = _core->update(deltaTime, req); !r) return r;
    if (!active) return {};   // no state: SkeletalAnimatorImpl::update returns before the model job (:1029-1032)
    const int layers = std::max(1, layersOf(req));
    if (layers > kMaxLayers) return unexpected<std::string>{"state blends more clips than the device supports"};
    Device& d = *_device;
    d.clip.assign(static_cast<size_t>(layers), 0);
    d.ratio.assign(static_cast<size_t>(layers), 0.f);
    d.weight.assign(static_cast<size_t>(layers), 0.f);
    writeRequest(req, 0, 1, d.clip.data(), d.ratio.data(), d.weight.data(), d.programs[0]);
    if (auto r = d.run(layers, false); !r) return r;
    _modelsValid = false;
    _core->afterUpdate();
    return {};
}

glm::mat4 SkeletalAnimator::getJointModelMatrix(const std::string& node) const noexcept {
    // src/skeleton_ozz.cpp:962-977: identity without skeleton or for an unknown joint name
    if (!_skeleton) return glm::mat4(1.f);
    const int joint = dmk_skeleton_find_joint(_skeleton->handle(), node.c_str());
    if (joint < 0) return glm::mat4(1.f);
    if (!_batch && !_device && !const_cast<SkeletalAnimator*>(this)->ensureDevice()) return glm::mat4(1.f);
    dmk_crowd_t crowd = _batch ? _batch->crowd() : _device->crowd;
    if (!crowd) return glm::mat4(1.f);
    if (_batch && _modelsFrame != _batch->_impl->frame) _modelsValid = false;
    if (!_modelsValid) {
        if (_batch) _modelsFrame = _batch->_impl->frame;
        _modelsCache.assign(static_cast<size_t>(_skeleton->getNumJoints()), glm::mat4(1.f));
        if (dmk_crowd_get_models(crowd, _batch ? _batchIndex : 0, 1, glm::value_ptr(_modelsCache[0])) != DMK_OK) return glm::mat4(1.f);
        _modelsValid = true;
    }
    return _modelsCache[static_cast<size_t>(joint)];
}

std::unordered_map<std::string, glm::mat4> SkeletalAnimator::getJointModelMatrixes(const glm::vec3& dir) const noexcept {
    std::unordered_map<std::string, glm::mat4> out;
    if (!_skeleton) return out;
    if (!_batch && !_device && !const_cast<SkeletalAnimator*>(this)->ensureDevice()) return out;
    dmk_crowd_t crowd = _batch ? _batch->crowd() : _device->crowd;
    if (!crowd) return out;
    const int J = _skeleton->getNumJoints();
    std::vector<glm::mat4> bones(static_cast<size_t>(J));
    std::vector<uint8_t> has(static_cast<size_t>(J));
    const float d[3] = {dir.x, dir.y, dir.z};
    if (dmk_crowd_get_bone_matrices(crowd, _batch ? _batchIndex : 0, 1, d, glm::value_ptr(bones[0]), has.data()) != DMK_OK) return out;
    out.reserve(static_cast<size_t>(J));
    for (int j = 0; j < J; ++j)
        if (has[static_cast<size_t>(j)]) out.emplace(std::string(_skeleton->getJointName(j)), bones[static_cast<size_t>(j)]);
    return out;
}

std::vector<glm::mat4> SkeletalAnimator::getSkinningPalette(int skin) const noexcept {
    // SkeletalAnimationRenderComponent::beforeRenderEntity (src/skeleton.cpp:236-249): [mat4(1)] + model * inverseBindPose,
    // once per Skinnable (= skin) of the character
    if (!_batch && !_device) const_cast<SkeletalAnimator*>(this)->ensureDevice();
    dmk_crowd_t crowd = _batch ? _batch->crowd() : (_device ? _device->crowd : nullptr);
    const bool known = skin >= 0 && static_cast<size_t>(skin) < _skins.size() && _skins[static_cast<size_t>(skin)].armature;
    const int ps = known ? static_cast<int>(_skins[static_cast<size_t>(skin)].armature->getJoints().size()) + 1 : 0;
    std::vector<glm::mat4> out(static_cast<size_t>(ps), glm::mat4(1.f));
    if (crowd && ps) dmk_crowd_get_skin_palette(crowd, skin, _batch ? _batchIndex : 0, 1, glm::value_ptr(out[0]));
    return out;
}

std::vector<float> SkeletalAnimator::getSkinnedVertices(int skin) const noexcept {
    if (!_batch && !_device) const_cast<SkeletalAnimator*>(this)->ensureDevice();
    dmk_crowd_t crowd = _batch ? _batch->crowd() : (_device ? _device->crowd : nullptr);
    const bool known = skin >= 0 && static_cast<size_t>(skin) < _skins.size() && _skins[static_cast<size_t>(skin)].mesh;
    const int v = known ? _skins[static_cast<size_t>(skin)].mesh->numVertices() : 0;
    std::vector<float> out(static_cast<size_t>(v) * 9, 0.f);
    if (crowd && v) dmk_crowd_get_skin_skinned(crowd, skin, _batch ? _batchIndex : 0, 1, out.data());
    return out;
}

// ---- SkeletalAnimationSceneComponent (batched) ----------------------------------------------------------------------------
SkeletalAnimationSceneComponent::SkeletalAnimationSceneComponent(std::shared_ptr<Skeleton> skel, SkeletalAnimator::AnimationMap anims,
                                                                 SkeletalAnimator::Definition def, std::shared_ptr<Armature> armature,
                                                                 const SkinnedMeshData* mesh) noexcept
    : _impl{std::make_unique<Impl>()} {
    _impl->skeleton = std::move(skel);
    _impl->anims = std::move(anims);
    _impl->def = std::move(def);
    _impl->skins.push_back({std::move(armature), mesh ? std::make_shared<const SkinnedMeshData>(*mesh) : nullptr});
}
SkeletalAnimationSceneComponent::~SkeletalAnimationSceneComponent() noexcept = default;

expected<int, std::string> SkeletalAnimationSceneComponent::addSkin(std::shared_ptr<Armature> armature, const SkinnedMeshData* mesh) noexcept {
    if (_impl->device) return unexpected<std::string>{"addSkin must be called before init"};
    if (!_impl->skins[0].armature) return unexpected<std::string>{"the component was created without an armature"};
    if (!armature) return unexpected<std::string>{"a skin needs an armature"};
    _impl->skins.push_back({std::move(armature), mesh ? std::make_shared<const SkinnedMeshData>(*mesh) : nullptr});
    return static_cast<int>(_impl->skins.size()) - 1;
}

expected<void, std::string> SkeletalAnimationSceneComponent::Impl::setupDeviceAnimators() noexcept {
    // flatten the definition: names -> indices.  Clip ids follow the device clip set (name-sorted, as Device::create
    // builds it); animations the provider does not know are dropped and a state without any cannot be created
    // (src/skeleton_ozz.cpp:469-485) -- such a state is left out, play() of it does nothing.
    std::vector<std::string> clipNames;
    for (const auto& [name, anim] : sortedClips(anims)) clipNames.push_back(name);
    auto clipIndex = [&](const std::string& name) {
        const auto itr = std::lower_bound(clipNames.begin(), clipNames.end(), name);
        return itr != clipNames.end() && *itr == name ? static_cast<int>(itr - clipNames.begin()) : -1;
    };
    std::vector<detail::ClipInfo> clipInfo;
    for (const auto& [name, anim] : sortedClips(anims)) clipInfo.push_back({name, anim->getDuration()});
    const detail::AnimatorCore probe(clipInfo, def);
    stateNames.clear();
    for (const auto& s : def.states)
        if (std::any_of(s.animations.begin(), s.animations.end(), [&](const auto& a) { return clipIndex(a.name) >= 0; })) stateNames.push_back(s.name);
    std::vector<dmk_anim_state_def> states;
    std::vector<dmk_anim_clip_def> clips;
    std::vector<dmk_anim_transition_def> transitions;
    stateDuration.clear();
    for (const auto& s : def.states) {
        if (stateIndex(s.name) < 0) continue;
        dmk_anim_state_def d{};
        d.first_clip = static_cast<int>(clips.size());
        for (const auto& a : s.animations)
            if (const int ci = clipIndex(a.name); ci >= 0) clips.push_back({ci, a.blend_position.x, a.blend_position.y, a.speed, a.loop ? 1 : 0});
        d.num_clips = static_cast<int>(clips.size()) - d.first_clip;
        d.blend_type = static_cast<int>(s.blend);
        d.threshold = s.threshold;
        d.tween_easing = static_cast<int>(s.tween.easing);
        d.tween_duration = s.tween.duration;
        d.next_state = s.next_state.empty() ? -1 : stateIndex(s.next_state);
        d.speed = s.speed;
        states.push_back(d);
        // getStateDuration() = calcDuration * definition speed; keep calcDuration alone
        stateDuration.push_back(s.speed != 0.f ? probe.getStateDuration(s.name) / s.speed : 0.f);
    }
    transitionDuration.clear();
    for (const auto& t : def.transitions) {
        // a transition naming a state that cannot exist can never match
        const int src = t.src_state.empty() ? -1 : stateIndex(t.src_state), dst = t.dst_state.empty() ? -1 : stateIndex(t.dst_state);
        if ((!t.src_state.empty() && src < 0) || (!t.dst_state.empty() && dst < 0)) continue;
        transitions.push_back({src, dst, static_cast<int>(t.tween.easing), t.tween.duration, t.offset});
        transitionDuration.push_back(t.tween.duration);
    }
    if (states.empty()) return unexpected<std::string>{"the animator definition has no state with a known animation"};
    if (dmk_crowd_set_animator(device->crowd, states.data(), static_cast<int>(states.size()), clips.data(), static_cast<int>(clips.size()),
                               transitions.data(), static_cast<int>(transitions.size())))
        return unexpected<std::string>{device->ctx->lastError()};
    const size_t n = static_cast<size_t>(device->n);
    cmd.assign(n, DMK_ANIM_NONE);
    cmdSpeed.assign(n, 1.f);
    blend.assign(2 * n, 0.f);
    speed.assign(n, 1.f);
    deviceAnimators = true;
    return {};
}

expected<void, std::string> SkeletalAnimationSceneComponent::init(int64_t numAnimators, AnimatorExecution execution) noexcept {
    if (_impl->device) return unexpected<std::string>{"already initialised"};
    auto dev = SkeletalAnimator::Device::create(_impl->skeleton, _impl->anims, _impl->skins, numAnimators);
    if (!dev) return unexpected<std::string>{dev.error()};
    _impl->device = std::move(*dev);
    if (execution == AnimatorExecution::Device)
        if (auto r = _impl->setupDeviceAnimators(); !r) {
            _impl->device.reset();
            return r;
        }
    for (int64_t i = 0; i < numAnimators; ++i) {
        auto a = std::make_unique<SkeletalAnimator>(_impl->skeleton, _impl->anims, _impl->def);
        a->_batch = this;
        a->_batchIndex = i;
        a->_skins = _impl->skins;   // shared pointers: armatures and meshes are not copied per animator
        _impl->animators.push_back(std::move(a));
    }
    _impl->requests.resize(static_cast<size_t>(numAnimators));
    _impl->active.resize(static_cast<size_t>(numAnimators));
    return {};
}
SkeletalAnimator& SkeletalAnimationSceneComponent::getAnimator(int64_t index) noexcept { return *_impl->animators[static_cast<size_t>(index)]; }
int64_t SkeletalAnimationSceneComponent::size() const noexcept { return static_cast<int64_t>(_impl->animators.size()); }
dmk_crowd_s* SkeletalAnimationSceneComponent::crowd() const noexcept { return _impl->device ? _impl->device->crowd : nullptr; }

expected<void, std::string> SkeletalAnimationSceneComponent::setInstances(const float* worldInverse, const float* bbox) noexcept {
    if (!_impl->device) return unexpected<std::string>{"not initialised"};
    if (dmk_crowd_set_instances(_impl->device->crowd, worldInverse, bbox)) return unexpected<std::string>{_impl->device->ctx->lastError()};
    _impl->device->hasInstances = true;
    return {};
}
expected<void, std::string> SkeletalAnimationSceneComponent::setTransforms(const float* position, const float* rotation, const float* scale,
                                                                           const float* bbox) noexcept {
    if (!_impl->device) return unexpected<std::string>{"not initialised"};
    if (dmk_crowd_set_transforms(_impl->device->crowd, position, rotation, scale, bbox)) return unexpected<std::string>{_impl->device->ctx->lastError()};
    _impl->device->hasInstances = true;
    return {};
}
expected<void, std::string> SkeletalAnimationSceneComponent::setTransformNodes(int numNodes, const float* position, const float* rotation,
                                                                               const float* scale, const int32_t* nodeParent,
                                                                               const int32_t* entityParent) noexcept {
    if (!_impl->device) return unexpected<std::string>{"not initialised"};
    if (dmk_crowd_set_transform_nodes(_impl->device->crowd, numNodes, position, rotation, scale, nodeParent, entityParent))
        return unexpected<std::string>{_impl->device->ctx->lastError()};
    return {};
}
expected<void, std::string> SkeletalAnimationSceneComponent::setCamera(const glm::mat4& viewProj, float normalizedNearDepth) noexcept {
    if (!_impl->device) return unexpected<std::string>{"not initialised"};
    if (dmk_crowd_set_camera(_impl->device->crowd, glm::value_ptr(viewProj), normalizedNearDepth)) return unexpected<std::string>{_impl->device->ctx->lastError()};
    _impl->device->hasCamera = true;
    return {};
}

expected<void, std::string> SkeletalAnimationSceneComponent::update(float deltaTime) noexcept {
    // src/skeleton.cpp:159-184: every animator ticks and errors are collected; here the per-entity device work of all of
    // them is one crowd step (pose programs of the whole crowd -> one pose launch, one cull launch, one skin launch)
    if (!_impl->device) return unexpected<std::string>{"not initialised"};
    auto& d = *_impl->device;
    const int64_t n = d.n;
    std::vector<std::string> errors;
    ++_impl->frame;
    if (_impl->deviceAnimators) {
        // the state machines tick on the device: only this frame's commands and inputs are uploaded
        auto& im = *_impl;
        auto err = [&]() { return unexpected<std::string>{d.ctx->lastError()}; };
        if (im.cmdDirty) {
            if (dmk_crowd_animator_commands(d.crowd, im.cmd.data(), im.cmdSpeed.data())) return err();
            std::fill(im.cmd.begin(), im.cmd.end(), static_cast<int32_t>(DMK_ANIM_NONE));
            std::fill(im.cmdSpeed.begin(), im.cmdSpeed.end(), 1.f);
            im.cmdDirty = false;
        }
        if (im.inputsDirty) {
            if (dmk_crowd_animator_set_inputs(d.crowd, im.blend.data(), im.speed.data())) return err();
            im.inputsDirty = false;
        }
        uint32_t flags = DMK_STEP_ANIMATOR | DMK_STEP_POSE;
        if (d.anyMesh) flags |= DMK_STEP_SKIN;
        if (d.hasInstances && d.hasCamera) flags |= DMK_STEP_CULL;
        if (dmk_crowd_step(d.crowd, deltaTime, flags)) return err();
        d.visibilityValid = false;
        im.infoValid = false;
        im.events.resize(static_cast<size_t>(n));
        im.status.resize(static_cast<size_t>(n));
        if (dmk_crowd_get_animator_events(d.crowd, im.events.data(), im.status.data())) return err();
        for (int64_t i = 0; i < n; ++i) {
            if (im.status[static_cast<size_t>(i)]) errors.push_back("error in the blending job");
            const dmk_anim_events& e = im.events[static_cast<size_t>(i)];
            if (e.cmd_started < 0 && !e.transition_finished && e.looped < 0 && e.finished < 0) continue;
            const auto itr = listenerRegistry().find(im.animators[static_cast<size_t>(i)].get());
            if (itr == listenerRegistry().end()) continue;
            SkeletalAnimator& a = *im.animators[static_cast<size_t>(i)];
            const auto listeners = itr->second.all;   // copied before the fan-out, as the reference does
            auto name = [&](int16_t s) { return std::string_view{im.stateNames[static_cast<size_t>(s)]}; };
            for (auto* l : listeners) {   // the reference's order: play(), then afterUpdate (src/skeleton_ozz.cpp:783-856, 1044-1087)
                if (e.cmd_started >= 0) l->onAnimatorStateStarted(a, name(e.cmd_started));
                if (e.cmd_transition_started) l->onAnimatorTransitionStarted(a);
                if (e.transition_finished) { l->onAnimatorTransitionFinished(a); l->onAnimatorStateFinished(a, name(e.prev_finished)); }
                if (e.looped >= 0) l->onAnimatorStateLooped(a, name(e.looped));
                if (e.finished >= 0) l->onAnimatorStateFinished(a, name(e.finished));
                if (e.auto_started >= 0) l->onAnimatorStateStarted(a, name(e.auto_started));
                if (e.auto_transition_started) l->onAnimatorTransitionStarted(a);
            }
        }
    } else {
        int layers = 1;
        for (int64_t i = 0; i < n; ++i) {
            auto& a = *_impl->animators[static_cast<size_t>(i)];
            _impl->active[static_cast<size_t>(i)] = a._core->hasActiveStateOrTransition();
            auto& req = _impl->requests[static_cast<size_t>(i)];
            if (auto r = a._core->update(deltaTime, req); !r) {
                errors.push_back(r.error());
                req = detail::PoseRequest{};   // a failed animator keeps its previous matrices this frame
                _impl->active[static_cast<size_t>(i)] = 0;
            }
            layers = std::max(layers, layersOf(req));
        }
        if (layers > kMaxLayers) return unexpected<std::string>{"state blends more clips than the device supports"};
        d.clip.assign(static_cast<size_t>(layers) * static_cast<size_t>(n), 0);
        d.ratio.assign(d.clip.size(), 0.f);
        d.weight.assign(d.clip.size(), 0.f);
        for (int64_t i = 0; i < n; ++i)
            writeRequest(_impl->requests[static_cast<size_t>(i)], i, n, d.clip.data(), d.ratio.data(), d.weight.data(), d.programs[static_cast<size_t>(i)]);
        if (auto r = d.run(layers, true); !r) return r;
        for (int64_t i = 0; i < n; ++i)
            if (_impl->active[static_cast<size_t>(i)]) _impl->animators[static_cast<size_t>(i)]->_core->afterUpdate();
    }
    if (!errors.empty()) {
        std::string joined;
        for (const auto& e : errors) joined += (joined.empty() ? "" : "\n") + e;
        return unexpected<std::string>{joined};
    }
    return {};
}

bool SkeletalAnimationSceneComponent::shouldEntityBeCulled(int64_t index) noexcept {
    // FrustumCuller::shouldEntityBeCulled (src/culling.cpp:287-290): membership in the culled set of the last update
    if (!_impl->device || !_impl->device->hasInstances || !_impl->device->hasCamera) return false;
    auto& d = *_impl->device;
    if (!d.visibilityValid) {
        d.visibility.assign(static_cast<size_t>((d.n + 31) / 32), 0u);
        if (dmk_crowd_get_visibility(d.crowd, d.visibility.data(), nullptr) != DMK_OK) return false;
        d.visibilityValid = true;
    }
    return ((d.visibility[static_cast<size_t>(index >> 5)] >> (index & 31)) & 1u) == 0u;
}

}  // namespace darmok
