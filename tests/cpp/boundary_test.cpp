// boundary_test.cpp -- the drop-in boundary pieces of the host shim, built as C++20 like darmok itself
// (CMakeLists.txt:28): darmok::expected without <expected>, OptionalRef getters, the protobuf -> definition adapter against
// protoc-style accessor classes, SkeletalAnimator::load(def, IComponentLoadContext&), and FrustumCuller against the oracle.
//
//   boundary_test cpu                 no device needed
//   boundary_test gpu <assets_dir>    FrustumCuller on the device (or the CPU lockstep build) vs oracle/cull.c, load through a context
#include <darmok_b200/culling.hpp>
#include <darmok_b200/protobuf_adapter.hpp>
#include <darmok_b200/skeleton.hpp>

#include <dmk_oracle.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <string>
#include <vector>

using namespace darmok;

static int g_failures = 0;
#define CHECK(cond, ...) do { if (!(cond)) { std::printf("FAIL: "); std::printf(__VA_ARGS__); std::printf("\n"); ++g_failures; } } while (0)

#if __cplusplus >= 202302L
#error "this test must be compiled as C++20 (darmok's language level): the point is that the shim does not need <expected>"
#endif

// ---- classes with the accessors protoc generates for assets/protobuf/skeleton.proto / math.proto / easing.proto ----
namespace pb {
struct Vec2 { float x_ = 0, y_ = 0; float x() const { return x_; } float y() const { return y_; } };
struct Vec4 { float v[4] = {0, 0, 0, 0}; float x() const { return v[0]; } float y() const { return v[1]; } float z() const { return v[2]; } float w() const { return v[3]; } };
struct Mat4 { Vec4 c[4]; const Vec4& col1() const { return c[0]; } const Vec4& col2() const { return c[1]; } const Vec4& col3() const { return c[2]; } const Vec4& col4() const { return c[3]; } };
enum Easing_Type : int { Easing_Type_Linear = 0, Easing_Type_QuadraticIn = 2 };
struct Tween { Easing_Type easing_ = Easing_Type_Linear; float duration_ = 0; Easing_Type easing() const { return easing_; } float duration() const { return duration_; } };
struct Animation {
    std::string name_; Vec2 bp; float speed_ = 0; bool loop_ = false;
    const std::string& name() const { return name_; } const Vec2& blend_position() const { return bp; } float speed() const { return speed_; } bool loop() const { return loop_; }
};
enum State_BlendType : int { State_BlendType_Cartesian = 0, State_BlendType_Directional = 1 };
struct State {
    std::string name_, next_; std::vector<Animation> anims; State_BlendType blend_ = State_BlendType_Cartesian; float threshold_ = 0, speed_ = 0; Tween tween_;
    const std::string& name() const { return name_; } const std::vector<Animation>& animations() const { return anims; } State_BlendType blend() const { return blend_; }
    float threshold() const { return threshold_; } const Tween& tween() const { return tween_; } const std::string& next_state() const { return next_; } float speed() const { return speed_; }
};
struct Transition {
    std::string src, dst; Tween tween_; float offset_ = 0;
    const std::string& src_state() const { return src; } const std::string& dst_state() const { return dst; } const Tween& tween() const { return tween_; } float offset() const { return offset_; }
};
struct Animator {
    std::string skel, pattern; std::vector<State> states_; std::vector<Transition> transitions_;
    const std::string& skeleton_path() const { return skel; } const std::string& animation_pattern() const { return pattern; }
    const std::vector<State>& states() const { return states_; } const std::vector<Transition>& transitions() const { return transitions_; }
};
struct ArmatureJoint { std::string name_; Mat4 ibp; const std::string& name() const { return name_; } const Mat4& inverse_bind_pose() const { return ibp; } };
struct Armature { std::string name_; std::vector<ArmatureJoint> joints_; const std::string& name() const { return name_; } const std::vector<ArmatureJoint>& joints() const { return joints_; } };
}  // namespace pb

namespace {
struct FailingSkeletonLoader final : ISkeletonLoader {
    expected<std::shared_ptr<Skeleton>, std::string> operator()(const std::filesystem::path& path) noexcept override {
        return unexpected<std::string>{"no skeleton at " + path.string()};
    }
};
struct FailingAnimationLoader final : ISkeletalAnimationLoader {
    int calls = 0;
    expected<std::shared_ptr<SkeletalAnimation>, std::string> operator()(const std::filesystem::path&) noexcept override {
        ++calls;
        return unexpected<std::string>{"no animation"};
    }
};
template <class S, class A>
struct Context final : IComponentLoadContext, IAssetContext {
    S& skeletons; A& animations;
    Context(S& s, A& a) : skeletons(s), animations(a) {}
    IAssetContext& getAssets() noexcept override { return *this; }
    ISkeletonLoader& getSkeletonLoader() noexcept override { return skeletons; }
    ISkeletalAnimationLoader& getSkeletalAnimationLoader() noexcept override { return animations; }
};

expected<int, std::string> half(int v) {
    if (v % 2) return unexpected<std::string>{"odd"};
    return v / 2;
}
expected<void, std::string> positive(int v) {
    if (v <= 0) return unexpected<std::string>{"not positive"};
    return {};
}
}  // namespace

static int runCpu() {
    // expected: the subset the shim relies on, whichever implementation the build picked
    CHECK(half(8) && *half(8) == 4 && half(8).value() == 4 && !half(7) && half(7).error() == "odd", "expected<int, string>");
    CHECK(positive(1) && !positive(0) && positive(0).error() == "not positive", "expected<void, string>");

    // OptionalRef getters: an animator without a state has none
    SkeletalAnimator empty;
    OptionalRef<const ISkeletalAnimatorState> st = empty.getCurrentState();
    OptionalRef<const ISkeletalAnimatorTransition> tr = empty.getCurrentTransition();
    CHECK(!st && !st.has_value() && st == nullptr && !tr, "OptionalRef of an animator without a state");

    // protobuf -> definition
    pb::Animator def;
    def.skel = "skeleton.ozz";
    def.pattern = "anims/*.ozz";
    pb::State loco;
    loco.name_ = "locomotion"; loco.blend_ = pb::State_BlendType_Directional; loco.threshold_ = 0.2f; loco.speed_ = 1.5f; loco.next_ = "idle";
    loco.tween_.easing_ = pb::Easing_Type_QuadraticIn; loco.tween_.duration_ = 0.5f;
    pb::Animation run; run.name_ = "run"; run.bp.x_ = 1.f; run.bp.y_ = -1.f; run.speed_ = 2.f; run.loop_ = true;
    loco.anims.push_back(run);
    def.states_.push_back(loco);
    pb::Transition t; t.src = "*"; t.dst = "locomotion"; t.tween_.duration_ = 0.25f; t.offset_ = 0.1f;
    def.transitions_.push_back(t);
    const SkeletalAnimatorDefinition d = fromProtobuf(def);
    CHECK(d.skeleton_path == "skeleton.ozz" && d.animation_pattern == "anims/*.ozz" && d.states.size() == 1 && d.transitions.size() == 1, "definition top level");
    CHECK(d.states[0].name == "locomotion" && d.states[0].blend == SkeletalAnimatorStateDefinition::Directional && d.states[0].threshold == 0.2f &&
              d.states[0].speed == 1.5f && d.states[0].next_state == "idle" && static_cast<int>(d.states[0].tween.easing) == 2 && d.states[0].tween.duration == 0.5f,
          "state fields");
    CHECK(d.states[0].animations.size() == 1 && d.states[0].animations[0].name == "run" && d.states[0].animations[0].blend_position.x == 1.f &&
              d.states[0].animations[0].blend_position.y == -1.f && d.states[0].animations[0].speed == 2.f && d.states[0].animations[0].loop,
          "animation fields");
    CHECK(d.transitions[0].src_state == "*" && d.transitions[0].dst_state == "locomotion" && d.transitions[0].tween.duration == 0.25f && d.transitions[0].offset == 0.1f,
          "transition fields");
    pb::Armature arm;
    arm.name_ = "arm";
    pb::ArmatureJoint j; j.name_ = "hip";
    for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) j.ibp.c[c].v[r] = float(c * 4 + r);
    arm.joints_.push_back(j);
    const ArmatureDefinition ad = armatureFromProtobuf(arm);
    CHECK(ad.name == "arm" && ad.joints.size() == 1 && ad.joints[0].name == "hip" && ad.joints[0].inverse_bind_pose[2][1] == 9.f && ad.joints[0].inverse_bind_pose[3][3] == 15.f,
          "armature fields (column-major)");

    // load(def, IComponentLoadContext&): the skeleton loader's error comes back (src/skeleton_ozz.cpp:921-925), nothing else ran
    FailingSkeletonLoader skeletons;
    FailingAnimationLoader animations;
    Context<FailingSkeletonLoader, FailingAnimationLoader> ctxt(skeletons, animations);
    SkeletalAnimator animator;
    const auto loaded = animator.load(d, ctxt);
    CHECK(!loaded && loaded.error() == "no skeleton at skeleton.ozz" && animations.calls == 0, "load through a context: skeleton error");

    // a culler without a context or without entities
    FrustumCuller none(nullptr);
    CHECK(!none.init(), "FrustumCuller::init without a device context is an error");
    CHECK(none.update(0.f) && !none.shouldEntityBeCulled(5) && none.culledCount() == 0, "a culler without entities culls nothing");
    std::printf("boundary cpu: %d failures (C++%ld)\n", g_failures, __cplusplus);
    return g_failures ? 1 : 0;
}

static std::vector<uint8_t> readAll(const std::string& path) {
    std::vector<uint8_t> out;
    if (FILE* f = std::fopen(path.c_str(), "rb")) {
        uint8_t buf[4096];
        size_t n;
        while ((n = std::fread(buf, 1, sizeof buf, f)) > 0) out.insert(out.end(), buf, buf + n);
        std::fclose(f);
    }
    return out;
}

static int runGpu(const std::string& assets) {
    auto ctxr = DeviceContext::create(0);
    if (!ctxr) { std::printf("FAIL: %s\n", ctxr.error().c_str()); return 1; }
    std::shared_ptr<DeviceContext> ctx = *ctxr;

    // ---- FrustumCuller: every Renderable with a BoundingBox, animated or not, vs oracle/cull.c (bit-exact visibility) ----
    const size_t n = 20011;
    std::mt19937 rng(17);
    std::uniform_real_distribution<float> u(-1.f, 1.f);
    std::vector<Entity> entities(n);
    std::vector<float> wi(n * 16), bb(n * 6);
    for (size_t i = 0; i < n; ++i) {
        entities[i] = static_cast<Entity>(1000 + i * 3);
        float* m = &wi[i * 16];   // inverse of a translation + uniform scale, column-major
        const float s = 0.5f + 0.75f * (u(rng) + 1.f), inv = 1.f / s;
        const float tx = 60.f * u(rng), ty = 10.f * u(rng), tz = 60.f * u(rng);
        std::memset(m, 0, 64);
        m[0] = m[5] = m[10] = inv; m[15] = 1.f;
        m[12] = -tx * inv; m[13] = -ty * inv; m[14] = -tz * inv;
        for (int k = 0; k < 3; ++k) { const float c = u(rng), e = 0.2f + 0.5f * (u(rng) + 1.f); bb[i * 6 + k] = c - e; bb[i * 6 + 3 + k] = c + e; }
    }
    glm::mat4 vp(1.f);   // a perspective * view built by hand: looking down +z from (0, 2, -30)
    {
        const float f = 1.f / std::tan(0.5f * 1.0471976f), aspect = 16.f / 9.f, zn = 0.1f, zf = 200.f;
        glm::mat4 proj(0.f), view(1.f);
        proj[0][0] = f / aspect; proj[1][1] = f; proj[2][2] = zf / (zf - zn); proj[2][3] = 1.f; proj[3][2] = -zn * zf / (zf - zn);
        view[3][0] = 0.f; view[3][1] = -2.f; view[3][2] = 30.f;
        for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) { float acc = 0.f; for (int k = 0; k < 4; ++k) acc += proj[k][r] * view[c][k]; vp[c][r] = acc; }
    }
    float corners[24];
    orc_frustum_corners(&vp[0][0], 0.f, corners);
    std::vector<uint32_t> ref((n + 31) / 32, 0u);
    orc_cull_batch(corners, wi.data(), bb.data(), static_cast<int64_t>(n), ref.data(), 1);

    FrustumCuller culler(ctx);
    CHECK(bool(culler.init()) && bool(culler.load(FrustumCuller::createDefinition())), "init / load");
    CHECK(bool(culler.setEntities(entities.data(), wi.data(), bb.data(), n)), "setEntities");
    CHECK(!culler.shouldEntityBeCulled(entities[0]), "before the first update nothing is culled");
    culler.setViewProjection(vp);
    CHECK(bool(culler.update(1.f / 60.f)), "update");
    size_t culled = 0, mismatches = 0;
    for (size_t i = 0; i < n; ++i) {
        const bool refVisible = (ref[i >> 5] >> (i & 31)) & 1u;
        culled += refVisible ? 0 : 1;
        mismatches += culler.shouldEntityBeCulled(entities[i]) != !refVisible;
    }
    CHECK(mismatches == 0 && culler.culledCount() == culled && culled > n / 10 && culled < n - n / 10,
          "FrustumCuller: %zu of %zu entities differ from the oracle (culled %zu, oracle %zu)", mismatches, n, culler.culledCount(), culled);
    CHECK(!culler.shouldEntityBeCulled(7), "an entity outside the camera's filter is never culled");
    // entities without a Transform: the camera frustum itself
    CHECK(bool(culler.setEntities(entities.data(), nullptr, bb.data(), 100)) && bool(culler.update(0.f)), "entities without transforms");
    {
        std::vector<float> identity(100 * 16, 0.f);
        for (size_t i = 0; i < 100; ++i) identity[i * 16] = identity[i * 16 + 5] = identity[i * 16 + 10] = identity[i * 16 + 15] = 1.f;
        std::vector<uint32_t> r2(4, 0u);
        orc_cull_batch(corners, identity.data(), bb.data(), 100, r2.data(), 1);
        size_t bad = 0;
        for (size_t i = 0; i < 100; ++i) bad += culler.shouldEntityBeCulled(entities[i]) != !((r2[i >> 5] >> (i & 31)) & 1u);
        CHECK(bad == 0, "identity world inverse: %zu differ", bad);
    }
    CHECK(bool(culler.shutdown()) && !culler.shouldEntityBeCulled(entities[1]) && culler.culledCount() == 0, "shutdown");

    // ---- load(def, IComponentLoadContext&) with the B200 loaders behind the context ----
    B200SkeletonLoader skeletons(ctx);
    B200SkeletalAnimationLoader animations(ctx);
    Context<B200SkeletonLoader, B200SkeletalAnimationLoader> ctxt(skeletons, animations);
    SkeletalAnimatorDefinition def;
    def.skeleton_path = assets + "/skeleton.ozz";
    def.animation_pattern = assets + "/*.ozz";
    SkeletalAnimatorStateDefinition idle;
    idle.name = "idle";
    idle.threshold = 0.1f;
    SkeletalAnimatorAnimationDefinition a;
    a.name = "Idle01"; a.loop = true;
    idle.animations.push_back(a);
    def.states.push_back(idle);
    SkeletalAnimator animator;
    const auto loaded = animator.load(def, ctxt);
    CHECK(bool(loaded), "load through a context: %s", loaded ? "" : loaded.error().c_str());
    if (loaded) {
        CHECK(animator.play("idle") && animator.getCurrentState() && animator.getCurrentState()->getName() == "idle", "play after load");
        CHECK(bool(animator.update(1.f / 60.f)), "update after load");
        const glm::mat4 m = animator.getJointModelMatrix("root");
        CHECK(std::isfinite(m[3][0]), "joint matrix after load");
    }
    (void)readAll;
    std::printf("boundary gpu: %d failures\n", g_failures);
    return g_failures ? 1 : 0;
}

int main(int argc, char** argv) {
    if (argc >= 3 && std::string(argv[1]) == "gpu") return runGpu(argv[2]);
    return runCpu();
}
