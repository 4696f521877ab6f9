// animator_test.cpp -- scripted animator scenarios run twice: through the host shim (darmok_b200/host) and through the
// oracle's restatement of darmok's animator (oracle/animator_ref.cpp), compared tick by tick.
//
//   animator_test <assets_dir> cpu   host state machine only: every PoseRequest is evaluated with the oracle's jobs and
//                                    the resulting model matrices must equal the reference animator's exactly
//   animator_test <assets_dir> gpu   full path: SkeletalAnimator / SkeletalAnimationSceneComponent on the device through
//                                    the C ABI; joint matrices, palette, skinned vertices, culling against the oracle
//
// <assets_dir> is written by tests/test_host_shim.py: skeleton.ozz, <clip>.ozz, armature.bin/.txt, mesh.bin
#include "darmok_b200/animator_core.hpp"
#include "darmok_b200/skeleton.hpp"

#include "animator_kernel_host.hpp"
#include "animator_ref.hpp"
#include "dmk_oracle.h"
#include "dmk.h"

#include <algorithm>
#include <array>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <memory>
#include <sstream>

using namespace darmok;

static int g_failures = 0;
#define CHECK(cond, ...)                                 \
    do {                                                 \
        if (!(cond)) {                                   \
            std::printf("FAIL %s:%d: ", __FILE__, __LINE__); \
            std::printf(__VA_ARGS__);                    \
            std::printf("\n");                           \
            if (++g_failures > 20) std::exit(1);         \
        }                                                \
    } while (0)

static std::vector<uint8_t> readFile(const std::string& path) {
    std::ifstream in(path, std::ios::binary);
    if (!in) { std::printf("cannot open %s\n", path.c_str()); std::exit(2); }
    return std::vector<uint8_t>((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
}

static const char* kClipNames[] = {"Idle01", "Run01_Backward", "Run01_BackwardLeft", "Run01_BackwardRight", "Run01_Forward",
                                   "Run01_ForwardLeft", "Run01_ForwardRight", "Run01_Left", "Run01_Right", "Talk01"};
static const float kBlendPos[9][2] = {{0, 0}, {0, -1}, {1, -1}, {-1, -1}, {0, 1}, {-1, 1}, {1, 1}, {-1, 0}, {1, 0}};

// the animator of samples/ozz/assets/animator.json, plus a few states that exercise the quirks
static void buildDefinitions(SkeletalAnimatorDefinition& def, orc::AnimatorDef& ref) {
    auto addState = [&](const std::string& name, std::vector<std::pair<std::string, std::pair<float, float>>> anims, int blend, float threshold,
                        float tweenDuration, bool loop, const std::string& next, float speed, float clipSpeed) {
        SkeletalAnimatorStateDefinition s;
        orc::StateDef r;
        s.name = r.name = name;
        for (auto& a : anims) {
            SkeletalAnimatorAnimationDefinition ad;
            ad.name = a.first;
            ad.blend_position = glm::vec2(a.second.first, a.second.second);
            ad.loop = loop;
            ad.speed = clipSpeed;
            s.animations.push_back(ad);
            r.animations.push_back({a.first, {a.second.first, a.second.second}, clipSpeed, loop});
        }
        s.blend = static_cast<SkeletalAnimatorStateDefinition::BlendType>(blend);
        r.blend = blend;
        s.threshold = r.threshold = threshold;
        s.tween.duration = r.tween.duration = tweenDuration;
        s.next_state = r.next_state = next;
        s.speed = r.speed = speed;
        def.states.push_back(s);
        ref.states.push_back(r);
    };
    std::vector<std::pair<std::string, std::pair<float, float>>> loco;
    for (int i = 0; i < 9; ++i) loco.push_back({kClipNames[i], {kBlendPos[i][0], kBlendPos[i][1]}});
    addState("locomotion", loco, 1, 0.2f, 0.5f, true, "", 1.f, 0.f);
    addState("talk", {{"Talk01", {0, 0}}}, 0, 0.f, 0.f, true, "", 1.f, 0.f);
    addState("greet", {{"Talk01", {0, 0}}}, 0, 0.f, 0.f, false, "locomotion", 1.f, 2.f);   // non looping, auto next state
    std::vector<std::pair<std::string, std::pair<float, float>>> cart = {{"Run01_Left", {-1, 0}}, {"Run01_Right", {1, 0}}, {"Run01_Forward", {0, 1}}};
    addState("strafe", cart, 0, 1.f, 0.25f, true, "", 1.5f, 0.f);                          // Cartesian, no clip at (0, 0)
    addState("broken", cart, 0, 0.f, 0.25f, true, "", 1.f, 0.f);                           // threshold 0: BlendingJob rejects (C7)
    addState("ghost", {{"NoSuchClip", {0, 0}}}, 0, 1.f, 0.f, true, "", 1.f, 0.f);           // no known animation: cannot be created
    auto addTransition = [&](const std::string& src, const std::string& dst, float duration, float offset, int easing) {
        SkeletalAnimatorTransitionDefinition t;
        t.src_state = src; t.dst_state = dst; t.tween.duration = duration; t.offset = offset;
        t.tween.easing = static_cast<EasingType>(easing);
        def.transitions.push_back(t);
        ref.transitions.push_back({src, dst, {easing, duration}, offset});
    };
    addTransition("locomotion", "talk", 0.5f, 0.f, 0);
    addTransition("talk", "locomotion", 0.5f, 0.1f, 4);     // QuadraticInOut with an offset
    addTransition("", "strafe", 0.3f, 0.f, 16);             // *>strafe, SinusoidalInOut
    addTransition("greet", "", 0.2f, 0.f, 0);               // greet>*
}

struct Action {
    enum Kind { Tick, Play, Stop, Blend, Speed, Pause } kind;
    std::string name;
    float a = 0.f, b = 0.f;
};
static std::vector<Action> scenario() {
    std::vector<Action> s;
    auto ticks = [&](int n, float dt) { for (int i = 0; i < n; ++i) s.push_back({Action::Tick, "", dt, 0}); };
    s.push_back({Action::Play, "run", 1.f, 0});             // C16: no such state
    ticks(3, 1 / 60.f);
    s.push_back({Action::Play, "ghost", 1.f, 0});           // state without a known clip
    s.push_back({Action::Play, "locomotion", 1.f, 0});
    ticks(8, 1 / 60.f);                                     // blend position (0,0) == Idle01's: early exit
    s.push_back({Action::Blend, "", 0.3f, 0.7f});
    ticks(45, 1 / 60.f);                                    // tweened blend position, directional gradient bands
    s.push_back({Action::Blend, "", -1.f, 0.f});
    ticks(20, 1 / 30.f);
    s.push_back({Action::Play, "locomotion", 2.f, 0});      // same state: only the speed changes, returns false
    ticks(10, 1 / 30.f);
    s.push_back({Action::Play, "talk", 1.f, 0});            // locomotion>talk
    ticks(12, 1 / 30.f);
    s.push_back({Action::Play, "locomotion", 1.f, 0});      // replace a running transition: its current state becomes the source
    ticks(25, 1 / 30.f);
    s.push_back({Action::Speed, "", 0.5f, 0});
    s.push_back({Action::Play, "strafe", 1.f, 0});          // *>strafe
    s.push_back({Action::Blend, "", 0.25f, 0.5f});
    ticks(30, 1 / 30.f);
    s.push_back({Action::Blend, "", 1.f, 0.f});             // lands exactly on Run01_Right after the tween
    ticks(30, 1 / 15.f);
    s.push_back({Action::Pause, "", 0, 0});
    ticks(2, 1 / 30.f);
    s.push_back({Action::Speed, "", 1.f, 0});
    s.push_back({Action::Play, "greet", 1.f, 0});           // non looping at clip speed 2, then auto "locomotion" through greet>*
    ticks(70, 1 / 30.f);
    s.push_back({Action::Stop, "", 0, 0});
    ticks(3, 1 / 30.f);
    s.push_back({Action::Play, "broken", 1.f, 0});          // "error in the blending job"
    ticks(2, 1 / 30.f);
    s.push_back({Action::Stop, "", 0, 0});
    s.push_back({Action::Play, "talk", 1.f, 0});
    ticks(4, 0.4f);                                         // long steps: loops
    return s;
}

struct EventLog : ISkeletalAnimatorListener {
    std::vector<std::string> events;
    void onAnimatorStateLooped(SkeletalAnimator&, std::string_view s) override { events.push_back("looped:" + std::string(s)); }
    void onAnimatorStateFinished(SkeletalAnimator&, std::string_view s) override { events.push_back("finished:" + std::string(s)); }
    void onAnimatorStateStarted(SkeletalAnimator&, std::string_view s) override { events.push_back("started:" + std::string(s)); }
    void onAnimatorTransitionFinished(SkeletalAnimator&) override { events.push_back("transition_finished"); }
    void onAnimatorTransitionStarted(SkeletalAnimator&) override { events.push_back("transition_started"); }
};

struct OracleAssets {
    orc_skeleton* skel = nullptr;
    std::vector<orc_animation*> anims;
    std::vector<orc::NamedAnimation> named;   // sorted by name, like the shim's clip table
};
static OracleAssets loadOracle(const std::string& dir) {
    OracleAssets o;
    char err[256];
    auto sk = readFile(dir + "/skeleton.ozz");
    o.skel = orc_skeleton_load(sk.data(), sk.size(), err, sizeof err);
    if (!o.skel) { std::printf("oracle skeleton: %s\n", err); std::exit(2); }
    std::vector<std::string> names(std::begin(kClipNames), std::end(kClipNames));
    std::sort(names.begin(), names.end());
    for (const auto& n : names) {
        auto data = readFile(dir + "/" + n + ".ozz");
        orc_animation* a = orc_animation_load(data.data(), data.size(), err, sizeof err);
        if (!a) { std::printf("oracle animation %s: %s\n", n.c_str(), err); std::exit(2); }
        o.anims.push_back(a);
        o.named.push_back({n, a});
    }
    return o;
}

// evaluates one group of a PoseRequest with the oracle's jobs (what the device does)
static std::vector<float> evalGroup(const OracleAssets& o, const detail::GroupRequest& g) {
    const int S = o.skel->num_soa;
    if (g.zero) return std::vector<float>((size_t)S * 40, 0.f);   // a clip that was never sampled: darmok's zero-initialised locals
    std::vector<std::vector<float>> samples;
    for (const auto& l : g.layers) {
        samples.emplace_back((size_t)S * 40);
        orc_sample_stateless(o.anims[l.clip], l.ratio, samples.back().data(), S, nullptr);
    }
    if (g.passthrough) return samples[g.rest];
    std::vector<const float*> ptrs;
    std::vector<float> w;
    for (size_t i = 0; i < samples.size(); ++i) { ptrs.push_back(samples[i].data()); w.push_back(g.layers[i].weight); }
    std::vector<float> out((size_t)S * 40);
    if (!orc_blend((int)ptrs.size(), ptrs.data(), w.data(), samples[g.rest].data(), g.threshold, S, out.data())) out.clear();
    return out;
}

static void applyAction(const Action& a, auto& shim, orc::RefAnimator& ref, int step) {
    switch (a.kind) {
    case Action::Play: {
        const bool r1 = shim.play(a.name, a.a), r2 = ref.play(a.name, a.a);
        CHECK(r1 == r2, "step %d play(%s): shim %d ref %d", step, a.name.c_str(), r1, r2);
        break;
    }
    case Action::Stop: shim.stop(); ref.stop(); break;
    case Action::Blend: shim.setBlendPosition(glm::vec2(a.a, a.b)); ref.setBlendPosition({a.a, a.b}); break;
    case Action::Speed: shim.setPlaybackSpeed(a.a); ref.setPlaybackSpeed(a.a); break;
    case Action::Pause: shim.pause(); ref.pause(); break;   // only changes getPlaybackState (C10)
    default: break;
    }
}

struct ScenarioStats {
    int ticks = 0, poses = 0, transitions = 0, blends = 0, passthroughs = 0, errors = 0;
    size_t events = 0;
};
// One animator through a list of actions, twice: the shim's state machine (its pose requests evaluated with the oracle's jobs,
// as the device evaluates them) and the literal restatement of darmok's animator; everything observable must be identical.
static ScenarioStats runScenarioCpu(const OracleAssets& o, const SkeletalAnimatorDefinition& def, const orc::AnimatorDef& rdef,
                                    const std::vector<Action>& actions, const char* label, bool stopAtUnsampledHold = false) {
    std::vector<detail::ClipInfo> clips;
    for (const auto& n : o.named) clips.push_back({n.name, n.anim->duration});
    detail::AnimatorCore core(clips, def);
    std::vector<std::string> events;
    core.events.stateLooped = [&](std::string_view s) { events.push_back("looped:" + std::string(s)); };
    core.events.stateFinished = [&](std::string_view s) { events.push_back("finished:" + std::string(s)); };
    core.events.stateStarted = [&](std::string_view s) { events.push_back("started:" + std::string(s)); };
    core.events.transitionFinished = [&] { events.push_back("transition_finished"); };
    core.events.transitionStarted = [&] { events.push_back("transition_started"); };
    orc::RefAnimator ref(o.skel, o.named, rdef);

    const int J = o.skel->num_joints, S = o.skel->num_soa;
    std::vector<float> models((size_t)J * 16);
    orc_local_to_model(o.skel, o.skel->rest, models.data());
    int step = 0, ticks = 0, poses = 0, transitions = 0, blends = 0, passthroughs = 0, errors = 0;
    for (const Action& a : actions) {
        if (stopAtUnsampledHold && g_failures) break;   // fuzzing: the first difference is the interesting one
        ++step;
        if (a.kind != Action::Tick) { applyAction(a, core, ref, step); continue; }
        ++ticks;
        const bool active = core.hasActiveStateOrTransition();
        detail::PoseRequest req;
        auto r = core.update(a.a, req);
        std::string refErr;
        const int holdsBefore = orc::g_unsampledHolds;
        const bool refOk = ref.update(a.a, refErr);
        if (stopAtUnsampledHold && orc::g_unsampledHolds != holdsBefore) break;   // the documented deviation (zero locals of a never-updated state): nothing comparable follows
        CHECK(r.has_value() == refOk, "%s step %d: shim ok=%d ref ok=%d (%s)", label, step, (int)r.has_value(), (int)refOk, refErr.c_str());
        if (!r) {
            ++errors;
            CHECK(r.error() == refErr, "%s step %d: error '%s' vs '%s'", label, step, r.error().c_str(), refErr.c_str());
            continue;
        }
        if (active) {
            if (req.top != detail::PoseRequest::Skip) {
                ++poses;
                std::vector<float> locals = evalGroup(o, req.group0);
                (req.group0.passthrough ? passthroughs : blends)++;
                if (req.top == detail::PoseRequest::Transition) {
                    ++transitions;
                    std::vector<float> cur = evalGroup(o, req.group1);
                    const float* ptrs[2] = {locals.data(), cur.data()};
                    const float w[2] = {req.w0, req.w1};
                    std::vector<float> out((size_t)S * 40);
                    CHECK(orc_blend(2, ptrs, w, locals.data(), FLT_MIN, S, out.data()), "%s step %d: transition blend failed", label, step);
                    locals = out;
                }
                orc_local_to_model(o.skel, locals.data(), models.data());
            }
            core.afterUpdate();
        }
        CHECK(std::memcmp(models.data(), ref.models().data(), models.size() * 4) == 0, "%s step %d: model matrices differ from the reference animator", label, step);
        CHECK(events == ref.events, "%s step %d: listener events differ (%zu vs %zu)", label, step, events.size(), ref.events.size());
        const auto cs = core.getCurrentState();
        const auto* rs = ref.currentState();
        CHECK((cs != nullptr) == (rs != nullptr), "%s step %d: current state presence", label, step);
        if (cs && rs) {
            CHECK(cs->getName() == rs->getName(), "%s step %d: state %s vs %s", label, step, std::string(cs->getName()).c_str(), rs->getName().c_str());
            CHECK(cs->getNormalizedTime() == rs->getNormalizedTime(), "%s step %d: normalized time %g vs %g", label, step, cs->getNormalizedTime(), rs->getNormalizedTime());
            CHECK(cs->getDuration() == rs->getDuration(), "%s step %d: duration %g vs %g", label, step, cs->getDuration(), rs->getDuration());
        }
        // Stopped without _state (also while a transition runs), else Paused / Playing by the toggle (src/skeleton_ozz.cpp:880-891)
        CHECK(static_cast<int>(core.getPlaybackState()) == (ref.getPlaybackState() == 0 ? (int)SkeletalAnimatorPlaybackState::Stopped
                                                            : ref.getPlaybackState() == 1 ? (int)SkeletalAnimatorPlaybackState::Playing
                                                                                          : (int)SkeletalAnimatorPlaybackState::Paused),
              "%s step %d: playback state %d, reference %d", label, step, (int)core.getPlaybackState(), ref.getPlaybackState());
        CHECK((core.getCurrentTransition() != nullptr) == (ref.currentTransition() != nullptr), "%s step %d: transition presence", label, step);
        if (core.getCurrentTransition() && ref.currentTransition())
            CHECK(core.getCurrentTransition()->getNormalizedTime() == ref.currentTransition()->getNormalizedTime(), "%s step %d: transition time", label, step);
    }
    for (const char* n : {"locomotion", "talk", "greet", "strafe", "nope"})
        CHECK(core.getStateDuration(n) == ref.getStateDuration(n), "getStateDuration(%s): %g vs %g", n, core.getStateDuration(n), ref.getStateDuration(n));
    ScenarioStats st;
    st.ticks = ticks; st.poses = poses; st.transitions = transitions; st.blends = blends; st.passthroughs = passthroughs; st.errors = errors;
    st.events = events.size();
    return st;
}

static int runCpu(const std::string& dir) {
    OracleAssets o = loadOracle(dir);
    SkeletalAnimatorDefinition def;
    orc::AnimatorDef rdef;
    buildDefinitions(def, rdef);
    const ScenarioStats st = runScenarioCpu(o, def, rdef, scenario(), "scenario");
    const int ticks = st.ticks, poses = st.poses, transitions = st.transitions, blends = st.blends, passthroughs = st.passthroughs, errors = st.errors;
    struct { size_t n; size_t size() const { return n; } } events{st.events};
    // MeshData::exportData's influence packing (src/mesh_core.cpp:336-356): known answers derived from the reference's loop
    {
        auto pack = [](std::vector<MeshDataWeight> w) {
            std::array<float, 8> o{};
            packVertexInfluences(w, o.data(), o.data() + 4);
            return o;
        };
        using A = std::array<float, 8>;
        CHECK((pack({{3, 0.2f}, {1, 0.5f}, {7, 0.3f}}) == A{1, 7, 3, -1, 0.5f, 0.3f, 0.2f, 0}), "packing: sorted by descending weight, unused index -1");
        CHECK((pack({}) == A{-1, -1, -1, -1, 1, 0, 0, 0}), "packing: no influence keeps weights (1, 0, 0, 0) on index -1");
        CHECK((pack({{2, 0.f}, {4, -0.1f}, {5, 0.4f}}) == A{5, -1, -1, -1, 0.4f, 0, 0, 0}), "packing: weights <= 0 are dropped");
        CHECK((pack({{0, .1f}, {1, .2f}, {2, .3f}, {3, .25f}, {4, .15f}}) == A{2, 3, 1, 4, .3f, .25f, .2f, .15f}),
              "packing: the four largest are kept as they are (no renormalisation)");
        std::vector<MeshDataVertex> verts(2);
        verts[0].position = glm::vec3(1.f, 2.f, 3.f);
        verts[0].weights = {{6, 1.f}};
        verts[1].tangent = glm::vec3(0.f, 0.f, 1.f);
        const SkinnedMeshData md = makeSkinnedMeshData(verts);
        CHECK(md.numVertices() == 2 && md.position[1] == 2.f && md.normal[1] == 1.f && md.normal[4] == 1.f && md.tangent[5] == 1.f, "mesh data streams");
        CHECK(md.indices[0] == 6.f && md.indices[1] == -1.f && md.weights[0] == 1.f && md.indices[4] == -1.f && md.weights[4] == 1.f, "mesh data influences");
    }
    {   // the rest of the public surface that needs no device (include/darmok/skeleton.hpp:259-301)
        SkeletalAnimator empty;   // default constructed: nothing to play, identity matrices, refuses to tick without a skeleton
        CHECK(!empty.play("locomotion") && empty.getCurrentState() == nullptr && empty.getJointModelMatrix("root") == glm::mat4(1.f), "empty animator");
        CHECK(!empty.update(0.1f).has_value(), "an animator without a skeleton cannot tick");
        CHECK(SkeletalAnimator::createDefinition().states.empty(), "createDefinition");
        struct Tagged : ISkeletalAnimatorListener { int tag; explicit Tagged(int t) : tag(t) {} };
        struct OddOnes : ISkeletalAnimatorListenerFilter {
            bool operator()(const ISkeletalAnimatorListener& l) const override { return static_cast<const Tagged&>(l).tag % 2 == 1; }
        };
        ArmatureDefinition ad = Armature::createDefinition();
        ad.joints.resize(2);
        ad.joints[0].name = "hip";
        ad.joints[1].name = "knee";
        ad.joints[1].inverse_bind_pose[3][1] = -0.5f;   // col4.y: a translation
        const Armature fromDef(ad);
        CHECK(fromDef.getJoints().size() == 2 && fromDef.getJoints()[0].inverseBindPose == glm::mat4(1.f) && fromDef.getJoints()[1].name == "knee" &&
                  fromDef.getJoints()[1].inverseBindPose[3][1] == -0.5f && fromDef.getJoints()[1].inverseBindPose[1][3] == 0.f,
              "Armature(Definition): columns stay columns");
        Tagged borrowed(3);
        empty.addListener(std::make_unique<Tagged>(1)).addListener(std::make_unique<Tagged>(2)).addListener(borrowed);
        CHECK(empty.removeListeners(OddOnes{}) == 2, "removeListeners returns how many went");
        CHECK(!empty.removeListener(borrowed) && empty.removeListeners(OddOnes{}) == 0, "they are gone");
    }
    // ConstSkeletalAnimatorDefinitionWrapper::loadAnimations' path rule (src/skeleton.cpp:538-548, src/string.cpp:134-149)
    {
        SkeletalAnimatorDefinition d;
        CHECK(animationPath(d, "Idle01") == std::filesystem::path("Idle01"), "no pattern: the name is the path");
        d.animation_pattern = "*.ozz";   // samples/ozz/assets/animator.json: "animationNamePattern": "*.ozz"
        CHECK(animationPath(d, "Idle01") == std::filesystem::path("Idle01.ozz"), "the sample's pattern");
        d.animation_pattern = "anims/*/*_v2.bin";
        CHECK(animationPath(d, "Run") == std::filesystem::path("anims/Run/Run_v2.bin"), "every * is replaced");
        d.animation_pattern = "fixed.ozz";
        CHECK(animationPath(d, "Run") == std::filesystem::path("fixed.ozz"), "a pattern without * is used as it is");
    }
    // every easing curve, specialised (shim) against general (oracle) form
    for (int type = 0; type < 32; ++type)
        for (int i = -4; i <= 44; ++i) {
            const float p = i / 40.f;
            const float a = ease01(static_cast<EasingType>(type), p), b = orc::ease(type, p, 0.f, 1.f);
            CHECK(a == b || (std::isnan(a) && std::isnan(b)), "easing %d at %g: %.9g vs %.9g", type, p, a, b);
        }
    // ... and both against the REFERENCE's own include/darmok/easing.hpp (compiled where it lies into oracle/_ref by
    // `make -C oracle ref`; vectors written by tests/golden/make_easing_golden.py): [32 types][(0,1),(2,-3)][49 positions]
    {
        const char* gd = std::getenv("DMK_GOLDEN_DIR");
        CHECK(gd != nullptr, "DMK_GOLDEN_DIR is not set (tests/test_host_shim.py sets it to tests/golden)");
        const std::vector<uint8_t> raw = gd ? readFile(std::string(gd) + "/easing_reference.f32") : std::vector<uint8_t>();
        CHECK(raw.size() == 32u * 2 * 49 * 4, "easing_reference.f32 has %zu bytes", raw.size());
        const float* g = reinterpret_cast<const float*>(raw.data());
        const float ranges[2][2] = {{0.f, 1.f}, {2.f, -3.f}};
        int pinned = 0;
        if (raw.size() == 32u * 2 * 49 * 4)
            for (int type = 0; type < 32; ++type)
                for (int r = 0; r < 2; ++r)
                    for (int i = -4; i <= 44; ++i) {
                        const float p = i / 40.f, want = g[(type * 2 + r) * 49 + (i + 4)];
                        const float b = orc::ease(type, p, ranges[r][0], ranges[r][1]);
                        CHECK(b == want || (std::isnan(b) && std::isnan(want)), "oracle easing %d [%g,%g] at %g: %.9g, reference %.9g", type, ranges[r][0],
                              ranges[r][1], p, b, want);
                        if (r == 0) {
                            const float a = ease01(static_cast<EasingType>(type), p);
                            CHECK(a == want || (std::isnan(a) && std::isnan(want)), "shim easing %d at %g: %.9g, reference %.9g", type, p, a, want);
                        }
                        ++pinned;
                    }
        std::printf("cpu: %d easing values equal to the reference's easing.hpp\n", pinned);
    }
    std::printf("cpu: %d ticks, %d poses (%d blends, %d passthroughs, %d in transition), %d job errors, %zu events, %d failures\n", ticks, poses, blends,
                passthroughs, transitions, errors, events.size(), g_failures);
    CHECK(transitions > 20 && blends > 50 && passthroughs > 20 && errors == 2 && events.size() > 10, "scenario lost coverage");
    return g_failures ? 1 : 0;
}

// Random action sequences (seeded): the same equivalence as the scripted scenario, over situations nobody thought of.
static int runFuzz(const std::string& dir, int sequences, uint64_t seed) {
    OracleAssets o = loadOracle(dir);
    SkeletalAnimatorDefinition def;
    orc::AnimatorDef rdef;
    buildDefinitions(def, rdef);
    uint64_t state = seed * 0x9E3779B97F4A7C15ull + 0x1234567ull;
    auto next = [&]() {   // xorshift64*
        state ^= state >> 12; state ^= state << 25; state ^= state >> 27;
        return state * 0x2545F4914F6CDD1Dull;
    };
    auto pick = [&](int n) { return static_cast<int>((next() >> 33) % static_cast<uint64_t>(n)); };
    auto unit = [&]() { return static_cast<float>((next() >> 40) * (1.0 / 16777216.0)); };
    const char* states[] = {"locomotion", "talk", "greet", "strafe", "locomotion", "talk", "run", "ghost", "broken"};
    const float dts[] = {1 / 60.f, 1 / 60.f, 1 / 30.f, 1 / 30.f, 1 / 15.f, 0.1f, 0.4f, 0.f, 1.7f, 1e-4f, -1 / 60.f, -0.3f};   // C1: negative steps too
    const float speeds[] = {1.f, 1.f, 2.f, 0.5f, 0.f, 3.f, -1.f};
    ScenarioStats total;
    int cut = 0;   // sequences that ended early at the documented deviation (darmok showing the zero locals of a never-updated state)
    for (int q = 0; q < sequences; ++q) {
        std::vector<Action> actions;
        const int len = 60 + pick(120);
        for (int i = 0; i < len; ++i) {
            const int kind = pick(100);
            if (kind < 70) {
                actions.push_back({Action::Tick, "", dts[pick(12)], 0});
            } else if (kind < 82) {
                actions.push_back({Action::Play, states[pick(9)], speeds[pick(7)], 0});
            } else if (kind < 92) {
                // blend positions: random, on a clip's own position (early exit, C6), or the origin
                const int how = pick(4);
                float x = unit() * 2.4f - 1.2f, y = unit() * 2.4f - 1.2f;
                if (how == 0) { const int c = pick(9); x = kBlendPos[c][0]; y = kBlendPos[c][1]; }
                if (how == 1) { x = 0.f; y = 0.f; }
                actions.push_back({Action::Blend, "", x, y});
            } else if (kind < 96) {
                actions.push_back({Action::Speed, "", speeds[pick(7)], 0});
            } else if (kind < 98) {
                actions.push_back({Action::Stop, "", 0, 0});
            } else {
                actions.push_back({Action::Pause, "", 0, 0});
            }
        }
        char label[32];
        std::snprintf(label, sizeof label, "fuzz %d", q);
        const int holds = orc::g_unsampledHolds;
        const ScenarioStats st = runScenarioCpu(o, def, rdef, actions, label, true);
        cut += orc::g_unsampledHolds != holds;
        total.ticks += st.ticks; total.poses += st.poses; total.transitions += st.transitions; total.blends += st.blends;
        total.passthroughs += st.passthroughs; total.errors += st.errors; total.events += st.events;
        if (g_failures) {
            std::printf("fuzz: first failing sequence %d (seed %llu); its actions:\n", q, (unsigned long long)seed);
            const char* kinds[] = {"tick", "play", "stop", "blend", "speed", "pause"};
            for (size_t i = 0; i < actions.size(); ++i)
                std::printf("  %zu %s %s %g %g\n", i + 1, kinds[actions[i].kind], actions[i].name.c_str(), actions[i].a, actions[i].b);
            break;
        }
    }
    std::printf("fuzz: %d sequences (%d cut where darmok shows the zero locals of a never-updated state), %d ticks, %d poses (%d blends, %d passthroughs, %d in transition), %d job errors, %zu events, %d failures\n", sequences,
                cut, total.ticks, total.poses, total.blends, total.passthroughs, total.transitions, total.errors, total.events, g_failures);
    return g_failures ? 1 : 0;
}

static bool close(float a, float e) {
    const double err = std::fabs((double)a - (double)e);
    return err <= 1e-6 || err <= 1e-5 * std::fabs((double)e);
}


// ---- the animator state machines themselves on the device (dmk_crowd_set_animator), straight
// through the C ABI.  Every character is driven by its own script; each tick is compared with the oracle's RefAnimator
// (events, error status, current state / transition and their clocks, model matrices) and with the host shim's
// AnimatorCore (the resolved layers and pose program, bit for bit where no libm function is involved).
struct FlatDefinition {
    std::vector<dmk_anim_state_def> states;
    std::vector<dmk_anim_clip_def> clips;
    std::vector<dmk_anim_transition_def> transitions;
    std::vector<std::string> stateNames;
    int stateIndex(const std::string& name) const {
        for (size_t i = 0; i < stateNames.size(); ++i) if (stateNames[i] == name) return (int)i;
        return -1;
    }
    // the command play(name) turns into: the state's index; nothing for an unknown name (C16); DMK_ANIM_CLEAR for a state that is
    // defined but cannot be created -- the reference has dropped _state and _transition by then (src/skeleton_ozz.cpp:793-845)
    int command(const SkeletalAnimatorDefinition& def, const std::string& name) const {
        if (const int s = stateIndex(name); s >= 0) return s;
        for (const auto& s : def.states) if (s.name == name) return DMK_ANIM_CLEAR;
        return DMK_ANIM_NONE;
    }
};
// names -> indices; a state without any known animation cannot be created (src/skeleton_ozz.cpp:469-485) and is left out, so
// play() of it is "no command"
static FlatDefinition flattenDefinition(const SkeletalAnimatorDefinition& def, const OracleAssets& o) {
    FlatDefinition f;
    auto clipIndex = [&](const std::string& name) {
        for (size_t i = 0; i < o.named.size(); ++i) if (o.named[i].name == name) return (int)i;
        return -1;
    };
    for (const auto& s : def.states) {
        bool any = false;
        for (const auto& a : s.animations) any = any || clipIndex(a.name) >= 0;
        if (any) f.stateNames.push_back(s.name);
    }
    for (const auto& s : def.states) {
        if (f.stateIndex(s.name) < 0) continue;
        dmk_anim_state_def d{};
        d.first_clip = (int)f.clips.size();
        for (const auto& a : s.animations) {
            const int ci = clipIndex(a.name);
            if (ci < 0) continue;
            f.clips.push_back({ci, a.blend_position.x, a.blend_position.y, a.speed, a.loop ? 1 : 0});
        }
        d.num_clips = (int)f.clips.size() - d.first_clip;
        d.blend_type = (int)s.blend;
        d.threshold = s.threshold;
        d.tween_easing = (int)s.tween.easing;
        d.tween_duration = s.tween.duration;
        d.next_state = s.next_state.empty() ? -1 : f.stateIndex(s.next_state);
        d.speed = s.speed;
        f.states.push_back(d);
    }
    for (const auto& t : def.transitions)
        f.transitions.push_back({t.src_state.empty() ? -1 : f.stateIndex(t.src_state), t.dst_state.empty() ? -1 : f.stateIndex(t.dst_state), (int)t.tween.easing,
                                 t.tween.duration, t.offset});
    return f;
}

struct AnimatorCounters { long exactWeights = 0, closeWeights = 0, transitions = 0, errors = 0, blends = 0, keptByDevice = 0; };
// One character, one tick: what the device (or its CPU build) reports against the reference animator (events, error status,
// current state / transition and clocks) and against the host shim's resolved pose request (layers and program).
static void compareAnimatorTick(int frame, int64_t i, int64_t n, float dt, orc::RefAnimator& ref, detail::AnimatorCore& core, std::vector<std::string>& log,
                                const dmk_anim_events& e, int status, const dmk_anim_state& di, const dmk_pose_program& g, const std::vector<int32_t>& lclip,
                                const std::vector<float>& lratio, const std::vector<float>& lweight, const std::vector<std::string>& stateNames,
                                const std::vector<dmk_anim_state_def>& dstates, AnimatorCounters& counters, bool* deviated = nullptr) {
    std::string err;
    const int holds = orc::g_unsampledHolds;
    const bool refOk = ref.update(dt, err);
    if (deviated && orc::g_unsampledHolds != holds) {   // the documented deviation (zero locals of a never-updated state): nothing comparable follows for this character
        *deviated = true;
        return;
    }
    const bool active = core.hasActiveStateOrTransition();
    detail::PoseRequest req;
    const auto r = core.update(dt, req);
    if (r && active) core.afterUpdate();
    CHECK(refOk == (status == 0), "frame %d animator %ld: status %d, reference ok=%d (%s)", frame, (long)i, status, (int)refOk, err.c_str());
    counters.errors += !refOk;
    // listener events in the reference's order
    if (e.cmd_started >= 0) log.push_back("started:" + stateNames[(size_t)e.cmd_started]);
    if (e.cmd_transition_started) log.push_back("transition_started");
    if (e.transition_finished) { log.push_back("transition_finished"); log.push_back("finished:" + stateNames[(size_t)e.prev_finished]); }
    if (e.looped >= 0) log.push_back("looped:" + stateNames[(size_t)e.looped]);
    if (e.finished >= 0) log.push_back("finished:" + stateNames[(size_t)e.finished]);
    if (e.auto_started >= 0) log.push_back("started:" + stateNames[(size_t)e.auto_started]);
    if (e.auto_transition_started) log.push_back("transition_started");
    CHECK(log == ref.events, "frame %d animator %ld: listener events differ (%zu vs %zu)", frame, (long)i, log.size(), ref.events.size());
    // current state / transition and their clocks (exact: only +, *, /, fmod)
    const auto* rs = ref.currentState();
                CHECK((rs != nullptr) == (di.state >= 0), "frame %d animator %ld: current state presence", frame, (long)i);
    if (rs && di.state >= 0) {
        CHECK(stateNames[(size_t)di.state] == rs->getName(), "frame %d animator %ld: state %s vs %s", frame, (long)i, stateNames[(size_t)di.state].c_str(), rs->getName().c_str());
        CHECK(di.state_time == rs->getNormalizedTime(), "frame %d animator %ld: state clock %.9g vs %.9g", frame, (long)i, di.state_time, rs->getNormalizedTime());
        const float duration = core.getStateDuration(std::string(rs->getName())) / dstates[(size_t)di.state].speed * di.state_speed;
        CHECK(std::fabs(duration - rs->getDuration()) <= 1e-6f * std::fabs(rs->getDuration()), "frame %d animator %ld: state duration %g vs %g", frame, (long)i, duration, rs->getDuration());
    }
    auto* rt = const_cast<orc::RefTransition*>(ref.currentTransition());
    CHECK((rt != nullptr) == (di.transition >= 0), "frame %d animator %ld: transition presence", frame, (long)i);
    if (rt && di.transition >= 0) {
        CHECK(di.transition_time == rt->getNormalizedTime(), "frame %d animator %ld: transition clock", frame, (long)i);
        CHECK(di.previous_state >= 0 && stateNames[(size_t)di.previous_state] == rt->previous().getName() && di.previous_time == rt->previous().getNormalizedTime(),
              "frame %d animator %ld: previous state of the transition", frame, (long)i);
    }
    // resolved program against the host shim's PoseRequest
    // Known difference of the device-side animators: a transition whose current state fails before the
    // transition has ever blended shows, in darmok and in the host state machine, the previous state's locals as they were when
    // the transition was created; the kernel keeps the matrices (SKIP).  Only reachable through a state whose BlendingJob is
    // invalid (threshold <= 0).
    const bool keptByDevice = r && active && req.top == detail::PoseRequest::Group0 && g.top == DMK_TOP_SKIP && di.transition >= 0 && di.transition_time == 0.f;
    counters.keptByDevice += keptByDevice;
    if (keptByDevice) {
    } else if (!r || !active || req.top == detail::PoseRequest::Skip) {
        CHECK(g.top == DMK_TOP_SKIP, "frame %d animator %ld: program top %d, expected SKIP", frame, (long)i, g.top);
    } else {
        const bool tr = req.top == detail::PoseRequest::Transition;
        counters.transitions += tr;
        CHECK(g.top == (tr ? DMK_TOP_TRANSITION : DMK_TOP_GROUP0), "frame %d animator %ld: program top %d (state %d, transition %d, previous %d)", frame, (long)i, g.top,
              di.state, di.transition, di.previous_state);
        int row = 0;
        for (int k = 0; k < (tr ? 2 : 1); ++k) {
            const detail::GroupRequest& gr = k ? req.group1 : req.group0;
            const int cnt = k ? g.num_layers1 : g.num_layers0, rest = k ? g.rest1 : g.rest0, mode = k ? g.mode1 : g.mode0;
            const float th = k ? g.threshold1 : g.threshold0;
            CHECK(cnt == (int)gr.layers.size() && rest == gr.rest && mode == (gr.zero ? DMK_GROUP_ZERO : gr.passthrough ? DMK_GROUP_PASSTHROUGH : DMK_GROUP_BLEND) && th == gr.threshold,
                  "frame %d animator %ld group %d: count %d/%zu rest %d/%d mode %d", frame, (long)i, k, cnt, gr.layers.size(), rest, gr.rest, mode);
            counters.blends += !gr.passthrough;
            for (size_t l = 0; l < gr.layers.size() && (int)l < cnt; ++l, ++row) {
                const size_t at = (size_t)row * n + i;
                CHECK(lclip[at] == gr.layers[l].clip, "frame %d animator %ld row %d: clip %d vs %d", frame, (long)i, row, lclip[at], gr.layers[l].clip);
                CHECK(lratio[at] == gr.layers[l].ratio, "frame %d animator %ld row %d: ratio %.9g vs %.9g", frame, (long)i, row, lratio[at], gr.layers[l].ratio);
                if (gr.passthrough) continue;
                if (lweight[at] == gr.layers[l].weight) ++counters.exactWeights;
                else {
                    ++counters.closeWeights;
                    CHECK(std::fabs(lweight[at] - gr.layers[l].weight) <= 2e-6f * std::fmax(1.f, std::fabs(gr.layers[l].weight)),
                          "frame %d animator %ld row %d: weight %.9g vs %.9g", frame, (long)i, row, lweight[at], gr.layers[l].weight);
                }
            }
        }
        if (tr) CHECK(std::fabs(g.weight0 - req.w0) <= 1e-6f && std::fabs(g.weight1 - req.w1) <= 1e-6f, "frame %d animator %ld: transition weights %.9g/%.9g vs %.9g/%.9g",
                      frame, (long)i, g.weight0, g.weight1, req.w0, req.w1);
    }
}

// ---- the same comparison without a GPU: the animator kernel's per-character code compiled for the CPU (animator_kernel_host.cpp),
// thousands of characters, each with its own seeded random script
static int runDeviceSim(const std::string& dir, int64_t n, int frames, uint64_t seed) {
    OracleAssets o = loadOracle(dir);
    SkeletalAnimatorDefinition def;
    orc::AnimatorDef rdef;
    buildDefinitions(def, rdef);
    const FlatDefinition flat = flattenDefinition(def, o);
    std::vector<float> durations;
    std::vector<detail::ClipInfo> clipInfo;
    for (const auto& a : o.named) { durations.push_back(a.anim->duration); clipInfo.push_back({a.name, a.anim->duration}); }
    devsim::Sim sim(flat.states.data(), (int)flat.states.size(), flat.clips.data(), (int)flat.clips.size(), flat.transitions.data(), (int)flat.transitions.size(),
                    durations, n);
    const int L = sim.layers();
    std::vector<std::unique_ptr<orc::RefAnimator>> refs;
    std::vector<std::unique_ptr<detail::AnimatorCore>> cores;
    for (int64_t i = 0; i < n; ++i) {
        refs.push_back(std::make_unique<orc::RefAnimator>(o.skel, o.named, rdef));
        cores.push_back(std::make_unique<detail::AnimatorCore>(clipInfo, def));
    }
    std::vector<std::vector<std::string>> devEvents((size_t)n);
    std::vector<char> deviated((size_t)n, 0);
    uint64_t state = seed * 0x9E3779B97F4A7C15ull + 0x777ull;
    auto next = [&]() { state ^= state >> 12; state ^= state << 25; state ^= state >> 27; return state * 0x2545F4914F6CDD1Dull; };
    auto pick = [&](int m) { return static_cast<int>((next() >> 33) % static_cast<uint64_t>(m)); };
    auto unit = [&]() { return static_cast<float>((next() >> 40) * (1.0 / 16777216.0)); };
    const char* states[] = {"locomotion", "talk", "greet", "strafe", "locomotion", "talk", "run", "ghost", "broken"};
    const float dts[] = {1 / 60.f, 1 / 60.f, 1 / 30.f, 1 / 30.f, 1 / 15.f, 0.1f, 0.4f, 0.f, 1.7f, 1e-4f, -1 / 60.f, -0.3f};
    const float speeds[] = {1.f, 1.f, 2.f, 0.5f, 0.f, 3.f, -1.f};
    std::vector<int32_t> cmd((size_t)n), status((size_t)n);
    std::vector<dmk_anim_state> info((size_t)n);
    std::vector<float> cmdSpeed((size_t)n), blend((size_t)n * 2, 0.f), speed((size_t)n, 1.f);
    std::vector<dmk_anim_events> ev((size_t)n);
    std::vector<int32_t> lclip((size_t)L * n);
    std::vector<float> lratio((size_t)L * n), lweight((size_t)L * n);
    std::vector<dmk_pose_program> prog((size_t)n);
    AnimatorCounters cnt;
    long ticksCompared = 0;
    for (int frame = 0; frame < frames && g_failures == 0; ++frame) {
        const float dt = dts[pick(12)];
        for (int64_t i = 0; i < n; ++i) {
            cmd[i] = DMK_ANIM_NONE;
            cmdSpeed[i] = 1.f;
            if (deviated[(size_t)i]) continue;
            const int kind = pick(100);
            if (kind < 5) {   // one command per animator and frame
                const std::string name = states[pick(9)];
                const float sp = speeds[pick(7)];
                cmd[i] = flat.command(def, name);
                cmdSpeed[i] = sp;
                refs[i]->play(name, sp);
                cores[i]->play(name, sp);
            } else if (kind < 6) {
                cmd[i] = DMK_ANIM_STOP;
                refs[i]->stop();
                cores[i]->stop();
            }
            const int what = pick(100);
            if (what < 8) {
                const int how = pick(4);
                float x = unit() * 2.4f - 1.2f, y = unit() * 2.4f - 1.2f;
                if (how == 0) { const int c = pick(9); x = kBlendPos[c][0]; y = kBlendPos[c][1]; }
                if (how == 1) { x = 0.f; y = 0.f; }
                blend[2 * i] = x;
                blend[2 * i + 1] = y;
                refs[i]->setBlendPosition({x, y});
                cores[i]->setBlendPosition(glm::vec2(x, y));
            } else if (what < 10) {
                speed[i] = speeds[pick(7)];
                refs[i]->setPlaybackSpeed(speed[i]);
                cores[i]->setPlaybackSpeed(speed[i]);
            }
        }
        if (const char* tc = std::getenv("DMK_TRACE_CHAR")) {   // debugging aid: the script of one character
            const int64_t t = std::atoll(tc);
            if (t >= 0 && t < n) std::printf("  trace %ld frame %d: dt %g cmd %d speed %g blend (%g, %g) playback %g\n", (long)t, frame, dt, cmd[t], cmdSpeed[t], blend[2 * t], blend[2 * t + 1], speed[t]);
        }
        sim.commands(cmd.data(), cmdSpeed.data());
        sim.setInputs(blend.data(), speed.data());
        sim.step(dt);
        sim.getEvents(ev.data(), status.data());
        sim.getState(info.data());
        sim.getPrograms(L, lclip.data(), lratio.data(), lweight.data(), prog.data());
        for (int64_t i = 0; i < n; ++i) {
            if (deviated[(size_t)i]) continue;
            bool dev = false;
            compareAnimatorTick(frame, i, n, dt, *refs[i], *cores[i], devEvents[(size_t)i], ev[(size_t)i], status[i], info[(size_t)i], prog[(size_t)i], lclip, lratio,
                                lweight, flat.stateNames, flat.states, cnt, &dev);
            deviated[(size_t)i] = dev;
            ticksCompared += !dev;
        }
    }
    long cut = 0, eventsSeen = 0;
    for (char d : deviated) cut += d;
    for (const auto& l : devEvents) eventsSeen += (long)l.size();
    std::printf("devsim: kernel code on the CPU, %ld characters x %d frames (%ld cut where darmok shows the zero locals of a never-updated state): %ld ticks compared, %ld transitions, "
                "%ld blends, %ld job errors, %ld events; blend weights %ld exact, %ld within 2e-6; %ld ticks where the device keeps the matrices of a never-blended "
                "transition; %d failures\n",
                (long)n, frames, cut, ticksCompared, cnt.transitions, cnt.blends, cnt.errors, eventsSeen, cnt.exactWeights, cnt.closeWeights, cnt.keptByDevice, g_failures);
    return g_failures ? 1 : 0;
}

#define DMK_OK_OR_FAIL(expr)                                                                              \
    do {                                                                                                  \
        if ((expr) != DMK_OK) { CHECK(false, "%s: %s", #expr, dmk_last_error(ctx)); return; }             \
    } while (0)

static void runDeviceAnimators(const std::string& dir, const OracleAssets& o, const SkeletalAnimatorDefinition& def, const orc::AnimatorDef& rdef) {
    dmk_context_t ctx = nullptr;
    if (dmk_context_create(0, nullptr, &ctx) != DMK_OK) { CHECK(false, "context"); return; }
    dmk_skeleton_t skel = nullptr;
    auto skelBytes = readFile(dir + "/skeleton.ozz");
    DMK_OK_OR_FAIL(dmk_skeleton_load_ozz(ctx, skelBytes.data(), skelBytes.size(), &skel));
    std::vector<dmk_clip_t> clips;
    std::vector<detail::ClipInfo> clipInfo;
    for (const auto& n : o.named) {   // clip index = position in the name-sorted table, like the shim's
        auto bytes = readFile(dir + "/" + n.name + ".ozz");
        dmk_clip_t clip = nullptr;
        DMK_OK_OR_FAIL(dmk_clip_load_ozz(ctx, bytes.data(), bytes.size(), &clip));
        clips.push_back(clip);
        clipInfo.push_back({n.name, n.anim->duration});
    }
    dmk_clipset_t clipset = nullptr;
    DMK_OK_OR_FAIL(dmk_clipset_create(ctx, skel, clips.data(), (int)clips.size(), &clipset));

    const FlatDefinition flat = flattenDefinition(def, o);
    const std::vector<dmk_anim_state_def>& dstates = flat.states;
    const std::vector<dmk_anim_clip_def>& dclips = flat.clips;
    const std::vector<dmk_anim_transition_def>& dtrans = flat.transitions;
    const std::vector<std::string>& stateNames = flat.stateNames;
    const int64_t n = 330;
    const int J = o.skel->num_joints, L = 18;
    dmk_crowd_t crowd = nullptr;
    DMK_OK_OR_FAIL(dmk_crowd_create(ctx, skel, clipset, nullptr, nullptr, n, L, &crowd));
    DMK_OK_OR_FAIL(dmk_crowd_enable_outputs(crowd, DMK_OUT_MODELS));
    CHECK(dmk_crowd_step(crowd, 0.f, DMK_STEP_ANIMATOR) == DMK_ERR_INVALID, "animator step before dmk_crowd_set_animator must fail");
    {   // a crowd too small for a transition between two 9-clip states is refused
        dmk_crowd_t small = nullptr;
        DMK_OK_OR_FAIL(dmk_crowd_create(ctx, skel, clipset, nullptr, nullptr, 4, 9, &small));
        CHECK(dmk_crowd_set_animator(small, dstates.data(), (int)dstates.size(), dclips.data(), (int)dclips.size(), dtrans.data(), (int)dtrans.size()) == DMK_ERR_INVALID,
              "max_layers < 2 * clips per state must be refused");
        dmk_crowd_destroy(small);
    }
    DMK_OK_OR_FAIL(dmk_crowd_set_animator(crowd, dstates.data(), (int)dstates.size(), dclips.data(), (int)dclips.size(), dtrans.data(), (int)dtrans.size()));
    DMK_OK_OR_FAIL(dmk_crowd_step(crowd, 0.f, DMK_STEP_ANIMATOR | DMK_STEP_POSE));   // all stopped: every program is SKIP

    std::vector<std::unique_ptr<orc::RefAnimator>> refs;
    std::vector<std::unique_ptr<detail::AnimatorCore>> cores;
    std::vector<std::vector<std::string>> devEvents((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        refs.push_back(std::make_unique<orc::RefAnimator>(o.skel, o.named, rdef));
        cores.push_back(std::make_unique<detail::AnimatorCore>(clipInfo, def));
    }
    const char* names[] = {"locomotion", "talk", "strafe", "greet"};
    std::vector<int32_t> cmd((size_t)n), status((size_t)n);
    std::vector<dmk_anim_state> info((size_t)n);
    std::vector<float> cmdSpeed((size_t)n), blend((size_t)n * 2, 0.f), speed((size_t)n, 1.f);
    std::vector<dmk_anim_events> ev((size_t)n);
    std::vector<int32_t> lclip((size_t)L * n);
    std::vector<float> lratio((size_t)L * n), lweight((size_t)L * n), models((size_t)n * J * 16);
    std::vector<dmk_pose_program> prog((size_t)n);
    AnimatorCounters cnt;
    long matricesChecked = 0, eventsSeen = 0;
    const float dt = 1 / 30.f;
    for (int frame = 0; frame < 64; ++frame) {
        bool anyCmd = false, inputs = false;
        for (int64_t i = 0; i < n; ++i) {
            cmd[i] = DMK_ANIM_NONE;
            cmdSpeed[i] = 1.f;
            auto play = [&](const std::string& name, float sp) {
                cmd[i] = flat.command(def, name);
                cmdSpeed[i] = sp;
                refs[i]->play(name, sp);
                cores[i]->play(name, sp);
                anyCmd = true;
            };
            if (frame == (int)(i % 7)) play(names[i % 4], 1.f);
            else if (frame == 20 + (int)(i % 5)) play(names[(i + 1) % 4], 1.f);
            else if (frame == 27 && i % 11 == 0) play(names[(i + 1) % 4], 2.f);   // same state again: only its speed changes
            else if (frame == 36 + (int)(i % 3) && i % 6 == 0) { cmd[i] = DMK_ANIM_STOP; refs[i]->stop(); cores[i]->stop(); anyCmd = true; }
            else if (frame == 44 && i % 4 == 1) play("broken", 1.f);              // threshold 0: "error in the blending job"
            else if (frame == 50 && i % 4 == 1) play("talk", 1.f);
            else if (frame == 46 && i % 9 == 2) play("no_such_state", 1.f);       // C16: nothing happens
            if (frame % 10 == (int)(i % 10)) {
                const float bx = std::sin(0.1f * (float)(i + frame)), by = std::cos(0.07f * (float)(i * 3 + frame));
                blend[2 * i] = (i % 13 == 5) ? 1.f : bx;    // some land exactly on a clip's blend position after the tween
                blend[2 * i + 1] = (i % 13 == 5) ? 0.f : by;
                refs[i]->setBlendPosition({blend[2 * i], blend[2 * i + 1]});
                cores[i]->setBlendPosition(glm::vec2(blend[2 * i], blend[2 * i + 1]));
                inputs = true;
            }
            if (frame == 15 && i % 3 == 0) {
                speed[i] = 0.5f + 0.25f * (float)(i % 5);
                refs[i]->setPlaybackSpeed(speed[i]);
                cores[i]->setPlaybackSpeed(speed[i]);
                inputs = true;
            }
        }
        if (anyCmd) DMK_OK_OR_FAIL(dmk_crowd_animator_commands(crowd, cmd.data(), cmdSpeed.data()));
        if (inputs) DMK_OK_OR_FAIL(dmk_crowd_animator_set_inputs(crowd, blend.data(), speed.data()));
        DMK_OK_OR_FAIL(dmk_crowd_step(crowd, dt, DMK_STEP_ANIMATOR | DMK_STEP_POSE));
        DMK_OK_OR_FAIL(dmk_crowd_get_animator_events(crowd, ev.data(), status.data()));
        DMK_OK_OR_FAIL(dmk_crowd_get_animator_state(crowd, info.data()));
        DMK_OK_OR_FAIL(dmk_crowd_get_programs(crowd, L, lclip.data(), lratio.data(), lweight.data(), prog.data()));
        const bool readMatrices = frame % 6 == 5 || frame == 63;
        if (readMatrices) DMK_OK_OR_FAIL(dmk_crowd_get_models(crowd, 0, n, models.data()));
        for (int64_t i = 0; i < n; ++i) {
            compareAnimatorTick(frame, i, n, dt, *refs[i], *cores[i], devEvents[(size_t)i], ev[(size_t)i], status[i], info[(size_t)i], prog[(size_t)i], lclip, lratio,
                                lweight, stateNames, dstates, cnt);
            // model matrices against the reference animator
            if (readMatrices && i % 3 == 0) {
                const float* m = &models[(size_t)i * J * 16];
                const float* e16 = refs[i]->models().data();
                bool everPosed = false;
                for (const auto& s : refs[i]->events) everPosed = everPosed || s.rfind("started:", 0) == 0;
                if (!everPosed) continue;   // the reference shows its load-time rest pose; the raw crowd has not been told to
                float flipped[16];   // dmk_crowd_get_models hands out what getJointModelMatrix does: OzzUtils::convert's left-handed matrix
                for (int j = 0; j < J; ++j) {
                    orc_flip_handedness(e16 + (size_t)j * 16, flipped);
                    for (int x = 0; x < 16; ++x) CHECK(close(m[j * 16 + x], flipped[x]), "frame %d animator %ld joint %d element %d: %g vs %g", frame, (long)i, j, x, m[j * 16 + x], flipped[x]);
                }
                ++matricesChecked;
            }
        }
    }
    for (const auto& l : devEvents) eventsSeen += (long)l.size();
    std::printf("gpu: device animators, %ld characters x 64 frames: %ld transitions, %ld blends, %ld job errors, %ld events, %ld characters' matrices; "
                "blend weights %ld exact, %ld within 2e-6\n", (long)n, cnt.transitions, cnt.blends, cnt.errors, eventsSeen, matricesChecked, cnt.exactWeights, cnt.closeWeights);
    CHECK(cnt.transitions > 1000 && cnt.blends > 3000 && cnt.errors > 100 && eventsSeen > 1500 && matricesChecked > 500, "device animator scenario lost coverage");
    dmk_crowd_destroy(crowd);
    dmk_clipset_destroy(clipset);
    for (auto c : clips) dmk_clip_destroy(c);
    dmk_skeleton_destroy(skel);
    dmk_context_destroy(ctx);
}

// ---- additions made after round 1's GPU budget was spent, kept apart from the established scenario above: a second Skinnable
// of the character (dmk_crowd_add_skin through SkeletalAnimator::addSkinning) and SkeletalAnimator::load
static int runGpuExtras(const std::string& dir) {
    OracleAssets o = loadOracle(dir);
    SkeletalAnimatorDefinition def;
    orc::AnimatorDef rdef;
    buildDefinitions(def, rdef);
    auto ctxr = DeviceContext::create(0);
    if (!ctxr) { std::printf("no device: %s\n", ctxr.error().c_str()); return 3; }
    auto ctx = *ctxr;
    B200SkeletonLoader skelLoader(ctx);
    B200SkeletalAnimationLoader animLoader(ctx);
    auto skel = skelLoader(dir + "/skeleton.ozz");
    CHECK(skel.has_value(), "skeleton load: %s", skel ? "" : skel.error().c_str());
    CHECK((*skel)->toString() == "Skeleton " + std::to_string(o.skel->num_joints) + " joints", "toString: %s", (*skel)->toString().c_str());
    auto bad = skelLoader(dir + "/Idle01.ozz");
    CHECK(!bad && bad.error() == "archive does not contain the type", "wrong archive type must keep darmok's message");
    // OzzSkeletalAnimationLoader answers every failure with one text (src/skeleton_ozz.cpp:207-216): wrong archive type, missing file
    auto notAnim = animLoader(dir + "/skeleton.ozz");
    CHECK(!notAnim && notAnim.error() == "archive doesn't contain an animation", "a skeleton archive is not an animation: %s", notAnim ? "" : notAnim.error().c_str());
    auto missing = animLoader(dir + "/no_such_clip.ozz");
    CHECK(!missing && missing.error() == "archive doesn't contain an animation", "missing animation file: %s", missing ? "" : missing.error().c_str());
    auto missingSkel = skelLoader(dir + "/no_such_skeleton.ozz");
    CHECK(!missingSkel && !missingSkel.error().empty() && missingSkel.error() != "archive does not contain the type", "missing skeleton file passes the data loader's message on");
    SkeletalAnimationMap anims;
    for (const char* n : kClipNames) {
        auto a = animLoader(dir + "/" + std::string(n) + ".ozz");
        CHECK(a.has_value(), "animation %s", n);
        CHECK((*a)->getName() == n, "animation name %s", std::string((*a)->getName()).c_str());
        anims[n] = *a;
    }
    // armature + mesh
    const int J = o.skel->num_joints;
    auto armBytes = readFile(dir + "/armature.bin");
    std::ifstream namesIn(dir + "/armature.txt");
    std::vector<ArmatureJoint> joints;
    std::vector<int32_t> remap;
    std::vector<float> invBind(armBytes.size() / 4);
    std::memcpy(invBind.data(), armBytes.data(), armBytes.size());
    for (size_t k = 0; k < invBind.size() / 16; ++k) {
        ArmatureJoint j;
        std::getline(namesIn, j.name);
        std::memcpy(glm::value_ptr(j.inverseBindPose), invBind.data() + k * 16, 64);
        remap.push_back(orc_skeleton_find_joint(o.skel, j.name.c_str()));
        joints.push_back(j);
    }
    auto armature = std::make_shared<Armature>(joints);
    const int K = (int)joints.size();
    auto meshBytes = readFile(dir + "/mesh.bin");
    const int V = (int)(meshBytes.size() / 4 / 17);
    SkinnedMeshData mesh;
    {
        const float* f = reinterpret_cast<const float*>(meshBytes.data());
        mesh.position.assign(f, f + V * 3);
        mesh.normal.assign(f + V * 3, f + V * 6);
        mesh.tangent.assign(f + V * 6, f + V * 9);
        mesh.indices.assign(f + V * 9, f + V * 13);
        mesh.weights.assign(f + V * 13, f + V * 17);
    }

    SkeletalAnimator animator(*skel, anims, def);
    CHECK(animator.setSkinning(armature, &mesh).has_value(), "setSkinning");
    // a second Skinnable of the same character: its own (partial, reversed) armature, palette only
    const int K2 = std::min(K, 12);
    std::vector<ArmatureJoint> joints2(joints.rend() - K2, joints.rend());
    std::vector<int32_t> remap2;
    std::vector<float> invBind2;
    for (const auto& j : joints2) {
        remap2.push_back(orc_skeleton_find_joint(o.skel, j.name.c_str()));
        const float* m = glm::value_ptr(j.inverseBindPose);
        invBind2.insert(invBind2.end(), m, m + 16);
    }
    const auto skin2 = animator.addSkinning(std::make_shared<Armature>(joints2));
    CHECK(skin2.has_value() && *skin2 == 1 && animator.getSkinCount() == 2, "addSkinning");
    std::vector<float> palette2((size_t)(K2 + 1) * 16);
    orc::RefAnimator ref(o.skel, o.named, rdef);
    std::vector<float> flipped(16);
    CHECK(animator.play("locomotion") && ref.play("locomotion"), "play");
    animator.setBlendPosition(glm::vec2(0.3f, 0.7f));
    ref.setBlendPosition({0.3f, 0.7f});
    int checked = 0;
    for (int step = 1; step <= 24; ++step) {
        if (step == 12) { CHECK(animator.play("talk") && ref.play("talk"), "play talk"); }
        std::string refErr;
        CHECK(animator.update(1 / 30.f).has_value() && ref.update(1 / 30.f, refErr), "step %d", step);
        orc_palette(ref.models().data(), J, remap2.data(), invBind2.data(), K2, palette2.data());
        const auto pal2 = animator.getSkinningPalette(1);
        CHECK(pal2.size() == (size_t)K2 + 1, "second skin's palette has %zu slots", pal2.size());
        for (size_t e = 0; e < palette2.size() && pal2.size() == (size_t)K2 + 1; ++e)
            CHECK(close(glm::value_ptr(pal2[0])[e], palette2[e]), "step %d palette of skin 1, element %zu", step, e);
        ++checked;
    }
    {   // SkeletalAnimator::load(def, ctxt) (src/skeleton_ozz.cpp:901-933): skeleton + animations through the loaders, reset, rest pose
        SkeletalAnimatorDefinition fromFiles = def;
        fromFiles.skeleton_path = dir + "/no_such_skeleton.ozz";
        fromFiles.animation_pattern = dir + "/*.ozz";
        CHECK(!animator.load(fromFiles, skelLoader, animLoader).has_value(), "a missing skeleton is the loader's error");
        CHECK(animator.getCurrentState() != nullptr, "a failed load leaves the animator as it was");
        fromFiles.skeleton_path = dir + "/skeleton.ozz";
        const auto loaded = animator.load(fromFiles, skelLoader, animLoader);
        CHECK(loaded.has_value(), "load: %s", loaded ? "" : loaded.error().c_str());
        CHECK(animator.getCurrentState() == nullptr && animator.getCurrentTransition() == nullptr && animator.getPlaybackSpeed() == 1.f &&
                  animator.getPlaybackState() == SkeletalAnimatorPlaybackState::Stopped && animator.getBlendPosition() == glm::vec2(0.f, 0.f),
              "load resets the animator");
        orc::RefAnimator fresh(o.skel, o.named, rdef);   // its load-time LocalToModelJob over the rest pose (:909-916)
        for (int j = 0; j < J; ++j) {
            const glm::mat4 m = animator.getJointModelMatrix(o.skel->names[j]);
            orc_flip_handedness(fresh.models().data() + (size_t)j * 16, flipped.data());
            for (int e = 0; e < 16; ++e) CHECK(close(glm::value_ptr(m)[e], flipped[e]), "rest pose after load: joint %d elem %d", j, e);
        }
        CHECK(animator.play("locomotion") && fresh.play("locomotion"), "the loaded definition plays");   // every clip of the sample was found
        std::string err;
        CHECK(animator.update(1 / 30.f).has_value() && fresh.update(1 / 30.f, err), "first tick after load");
        for (int j = 0; j < J; j += 7) {
            const glm::mat4 m = animator.getJointModelMatrix(o.skel->names[j]);
            orc_flip_handedness(fresh.models().data() + (size_t)j * 16, flipped.data());
            for (int e = 0; e < 16; ++e) CHECK(close(glm::value_ptr(m)[e], flipped[e]), "pose after load: joint %d elem %d", j, e);
        }
    }
    std::printf("gpuextras: %d ticks of a second skin's palette, load / reload; %d failures\n", checked, g_failures);
    return g_failures ? 1 : 0;
}

static int runGpu(const std::string& dir) {
    OracleAssets o = loadOracle(dir);
    SkeletalAnimatorDefinition def;
    orc::AnimatorDef rdef;
    buildDefinitions(def, rdef);
    auto ctxr = DeviceContext::create(0);
    if (!ctxr) { std::printf("no device: %s\n", ctxr.error().c_str()); return 3; }
    auto ctx = *ctxr;
    B200SkeletonLoader skelLoader(ctx);
    B200SkeletalAnimationLoader animLoader(ctx);
    auto skel = skelLoader(dir + "/skeleton.ozz");
    CHECK(skel.has_value(), "skeleton load: %s", skel ? "" : skel.error().c_str());
    CHECK((*skel)->toString() == "Skeleton " + std::to_string(o.skel->num_joints) + " joints", "toString: %s", (*skel)->toString().c_str());
    auto bad = skelLoader(dir + "/Idle01.ozz");
    CHECK(!bad && bad.error() == "archive does not contain the type", "wrong archive type must keep darmok's message");
    SkeletalAnimationMap anims;
    for (const char* n : kClipNames) {
        auto a = animLoader(dir + "/" + std::string(n) + ".ozz");
        CHECK(a.has_value(), "animation %s", n);
        CHECK((*a)->getName() == n, "animation name %s", std::string((*a)->getName()).c_str());
        anims[n] = *a;
    }
    // armature + mesh
    const int J = o.skel->num_joints;
    auto armBytes = readFile(dir + "/armature.bin");
    std::ifstream namesIn(dir + "/armature.txt");
    std::vector<ArmatureJoint> joints;
    std::vector<int32_t> remap;
    std::vector<float> invBind(armBytes.size() / 4);
    std::memcpy(invBind.data(), armBytes.data(), armBytes.size());
    for (size_t k = 0; k < invBind.size() / 16; ++k) {
        ArmatureJoint j;
        std::getline(namesIn, j.name);
        std::memcpy(glm::value_ptr(j.inverseBindPose), invBind.data() + k * 16, 64);
        remap.push_back(orc_skeleton_find_joint(o.skel, j.name.c_str()));
        joints.push_back(j);
    }
    auto armature = std::make_shared<Armature>(joints);
    const int K = (int)joints.size();
    auto meshBytes = readFile(dir + "/mesh.bin");
    const int V = (int)(meshBytes.size() / 4 / 17);
    SkinnedMeshData mesh;
    {
        const float* f = reinterpret_cast<const float*>(meshBytes.data());
        mesh.position.assign(f, f + V * 3);
        mesh.normal.assign(f + V * 3, f + V * 6);
        mesh.tangent.assign(f + V * 6, f + V * 9);
        mesh.indices.assign(f + V * 9, f + V * 13);
        mesh.weights.assign(f + V * 13, f + V * 17);
    }

    // ---- one stand-alone animator through the whole scenario ----
    {
        SkeletalAnimator animator(*skel, anims, def);
        CHECK(animator.setSkinning(armature, &mesh).has_value(), "setSkinning");
        EventLog log;
        animator.addListener(log);
        orc::RefAnimator ref(o.skel, o.named, rdef);
        std::vector<float> flipped(16), palette((size_t)(K + 1) * 16), skinned((size_t)V * 9);
        int step = 0, checked = 0;
        for (const Action& a : scenario()) {
            ++step;
            if (a.kind != Action::Tick) { applyAction(a, animator, ref, step); continue; }
            auto r = animator.update(a.a);
            std::string refErr;
            const bool refOk = ref.update(a.a, refErr);
            CHECK(r.has_value() == refOk, "step %d: update ok %d vs %d", step, (int)r.has_value(), (int)refOk);
            if (!r) { CHECK(r.error() == refErr, "step %d: '%s' vs '%s'", step, r.error().c_str(), refErr.c_str()); continue; }
            CHECK(log.events == ref.events, "step %d: events differ", step);
            if (step % 3 && step > 6) continue;   // matrices are read back every third tick
            ++checked;
            for (int j = 0; j < J; ++j) {
                const glm::mat4 m = animator.getJointModelMatrix(o.skel->names[j]);
                orc_flip_handedness(ref.models().data() + (size_t)j * 16, flipped.data());
                for (int e = 0; e < 16; ++e) CHECK(close(glm::value_ptr(m)[e], flipped[e]), "step %d joint %d elem %d: %g vs %g", step, j, e, glm::value_ptr(m)[e], flipped[e]);
            }
            orc_palette(ref.models().data(), J, remap.data(), invBind.data(), K, palette.data());
            const auto pal = animator.getSkinningPalette();
            for (size_t e = 0; e < palette.size(); ++e) CHECK(close(glm::value_ptr(pal[0])[e], palette[e]), "step %d palette %zu", step, e);
            orc_skin(palette.data(), K + 1, mesh.position.data(), mesh.normal.data(), mesh.tangent.data(), mesh.indices.data(), mesh.weights.data(), V, skinned.data());
            const auto sk = animator.getSkinnedVertices();
            for (size_t e = 0; e < skinned.size(); ++e) CHECK(close(sk[e], skinned[e]), "step %d skinned %zu: %g vs %g", step, e, sk[e], skinned[e]);
        }
        CHECK(animator.getJointModelMatrix("no_such_joint") == glm::mat4(1.f), "unknown joint must give identity (C12)");
        {   // debug bones (getJointModelMatrixes, src/skeleton_ozz.cpp:979-1004) against the oracle on the reference animator's matrices
            std::vector<float> bones((size_t)J * 16);
            std::vector<uint8_t> has((size_t)J);
            const float dir[3] = {1.f, 0.f, 0.f};
            orc_bone_matrices(o.skel, ref.models().data(), dir, bones.data(), has.data());
            const auto map = animator.getJointModelMatrixes();
            size_t expected = 0;
            for (int j = 0; j < J; ++j) {
                if (!has[(size_t)j]) { CHECK(!map.contains(o.skel->names[j]), "leaf joint %d has no bone entry", j); continue; }
                ++expected;
                const auto itr = map.find(o.skel->names[j]);
                CHECK(itr != map.end(), "bone entry of joint %d", j);
                if (itr == map.end()) continue;
                for (int e = 0; e < 16; ++e) {
                    const float a = glm::value_ptr(itr->second)[e], b = bones[(size_t)j * 16 + e];
                    CHECK(std::fabs(a - b) <= 2e-5f + 2e-4f * std::fabs(b), "bone %d element %d: %g vs %g", j, e, a, b);
                }
            }
            CHECK(map.size() == expected && expected > 10, "bone map size %zu vs %zu", map.size(), expected);
        }
        CHECK(animator.getPlaybackState() == SkeletalAnimatorPlaybackState::Paused, "pause toggled once and a state is playing");
        std::printf("gpu: stand-alone animator, %d ticks compared with matrices/palette/skin, %zu events\n", checked, log.events.size());
    }

    // ---- batched scene component: many animators in different situations, one device step per frame ----
    {
        const int64_t n = 300;
        SkeletalAnimationSceneComponent scene(*skel, anims, def, armature, nullptr);
        CHECK(scene.init(n).has_value(), "scene init");
        std::vector<std::unique_ptr<orc::RefAnimator>> refs;
        for (int64_t i = 0; i < n; ++i) refs.push_back(std::make_unique<orc::RefAnimator>(o.skel, o.named, rdef));
        const char* states[] = {"locomotion", "talk", "strafe", "greet"};
        for (int frame = 0; frame < 50; ++frame) {
            for (int64_t i = 0; i < n; ++i) {
                auto& a = scene.getAnimator(i);
                if (frame == (int)(i % 7)) { a.play(states[i % 4]); refs[i]->play(states[i % 4]); }
                if (frame == 20 + (int)(i % 5)) { a.play(states[(i + 1) % 4]); refs[i]->play(states[(i + 1) % 4]); }
                if (frame % 10 == (int)(i % 10)) {
                    const float bx = std::sin(0.1f * (float)(i + frame)), by = std::cos(0.07f * (float)(i * 3 + frame));
                    a.setBlendPosition(glm::vec2(bx, by));
                    refs[i]->setBlendPosition({bx, by});
                }
            }
            auto r = scene.update(1 / 30.f);
            CHECK(r.has_value(), "scene update frame %d: %s", frame, r ? "" : r.error().c_str());
            for (int64_t i = 0; i < n; ++i) {
                std::string err;
                refs[i]->update(1 / 30.f, err);
            }
            if (frame % 7 != 6) continue;
            std::vector<float> flipped(16);
            for (int64_t i = 0; i < n; i += 13) {
                auto& a = scene.getAnimator(i);
                for (int j = 0; j < J; j += 5) {
                    const glm::mat4 m = a.getJointModelMatrix(o.skel->names[j]);
                    orc_flip_handedness(refs[i]->models().data() + (size_t)j * 16, flipped.data());
                    for (int e = 0; e < 16; ++e) CHECK(close(glm::value_ptr(m)[e], flipped[e]), "frame %d animator %ld joint %d", frame, (long)i, j);
                }
            }
        }
        // culling through the scene component against the oracle
        std::vector<float> wi((size_t)n * 16, 0.f), bbox((size_t)n * 6);
        for (int64_t i = 0; i < n; ++i) {
            float* m = &wi[(size_t)i * 16];
            m[0] = m[5] = m[10] = m[15] = 1.f;
            m[12] = -(float)((i % 40) - 20) * 1.5f;   // world = translate(x, 0, z): inverse translates back
            m[14] = -(float)(i / 40) * 3.f;
            const float b[6] = {-0.3f, 0.f, -0.3f, 0.3f, 1.8f, 0.3f};
            std::memcpy(&bbox[(size_t)i * 6], b, sizeof b);
        }
        glm::mat4 vp(1.f);
        const float vpv[16] = {0.974279f, 0, 0, 0, 0, 1.732051f, 0, 0, 0, 0, 1.0003f, 1, 0, -1.5f, 1.7f, 2.0f};
        std::memcpy(glm::value_ptr(vp), vpv, sizeof vpv);
        CHECK(scene.setInstances(wi.data(), bbox.data()).has_value(), "setInstances");
        CHECK(scene.setCamera(vp, 0.f).has_value(), "setCamera");
        CHECK(scene.update(1 / 30.f).has_value(), "scene update with cull");
        float corners[24];
        orc_frustum_corners(vpv, 0.f, corners);
        int culled = 0;
        for (int64_t i = 0; i < n; ++i) {
            const bool ref = !orc_cull_visible(corners, &wi[(size_t)i * 16], &bbox[(size_t)i * 6]);
            CHECK(scene.shouldEntityBeCulled(i) == ref, "cull of entity %ld", (long)i);
            culled += ref;
        }
        CHECK(culled > 0 && culled < n, "cull test must see both outcomes (%d of %ld culled)", culled, (long)n);
        // the same entities described by their Transforms, hanging under two group nodes (Transform::update on the device)
        std::vector<float> pos((size_t)n * 3), rot((size_t)n * 4), scl((size_t)n * 3, 1.f);
        std::vector<int32_t> eparent((size_t)n);
        for (int64_t i = 0; i < n; ++i) {
            pos[(size_t)i * 3] = (float)((i % 40) - 20) * 1.5f; pos[(size_t)i * 3 + 1] = 0.f; pos[(size_t)i * 3 + 2] = (float)(i / 40) * 3.f;
            const float h = 0.01f * (float)(i % 17);
            rot[(size_t)i * 4] = 0.f; rot[(size_t)i * 4 + 1] = std::sin(h); rot[(size_t)i * 4 + 2] = 0.f; rot[(size_t)i * 4 + 3] = std::cos(h);
            eparent[(size_t)i] = (int32_t)(i % 3) - 1;
        }
        const float npos[6] = {0.5f, 0.f, 1.f, -2.f, 0.25f, 0.f}, nrot[8] = {0, 0, 0, 1, 0, 0.0998334f, 0, 0.9950042f}, nscl[6] = {1, 1, 1, 1.25f, 1.25f, 1.25f};
        const int32_t nparent[2] = {-1, 0};
        CHECK(scene.setTransforms(pos.data(), rot.data(), scl.data(), bbox.data()).has_value(), "setTransforms");
        CHECK(scene.setTransformNodes(2, npos, nrot, nscl, nparent, eparent.data()).has_value(), "setTransformNodes");
        CHECK(scene.update(1 / 30.f).has_value(), "scene update with Transform-driven cull");
        std::vector<float> nodeWorld(32), nodeInv(32);
        orc_transform_nodes(2, npos, nrot, nscl, nparent, nodeWorld.data(), nodeInv.data());
        std::vector<uint32_t> refBits((size_t)(n + 31) / 32);
        orc_cull_batch_trs_parented(corners, pos.data(), rot.data(), scl.data(), bbox.data(), n, eparent.data(), nodeInv.data(), refBits.data(), nullptr, 1);
        int culledTrs = 0;
        for (int64_t i = 0; i < n; ++i) {
            const bool ref = !((refBits[(size_t)i >> 5] >> (i & 31)) & 1u);
            CHECK(scene.shouldEntityBeCulled(i) == ref, "Transform-driven cull of entity %ld", (long)i);
            culledTrs += ref;
        }
        CHECK(culledTrs > 0 && culledTrs < n, "Transform-driven cull must see both outcomes (%d of %ld culled)", culledTrs, (long)n);
        std::printf("gpu: scene component, %ld animators x 51 frames, %d culled\n", (long)n, culled);
    }
    runDeviceAnimators(dir, o, def, rdef);

    // ---- the same scene component with AnimatorExecution::Device: darmok's API on top of the device-side state machines ----
    {
        const int64_t n = 240;
        std::vector<EventLog> logs((size_t)n);   // listeners outlive the animators (onAnimatorDestroyed)
        SkeletalAnimationSceneComponent scene(*skel, anims, def, armature, nullptr);
        CHECK(scene.init(n, AnimatorExecution::Device).has_value(), "scene init (device animators)");
        std::vector<std::unique_ptr<orc::RefAnimator>> refs;
        for (int64_t i = 0; i < n; ++i) {
            refs.push_back(std::make_unique<orc::RefAnimator>(o.skel, o.named, rdef));
            if (i % 2 == 0) scene.getAnimator(i).addListener(logs[(size_t)i]);
        }
        const char* states[] = {"locomotion", "talk", "strafe", "greet"};
        long stateChecks = 0, transitionChecks = 0, eventCount = 0;
        for (int frame = 0; frame < 56; ++frame) {
            for (int64_t i = 0; i < n; ++i) {
                auto& a = scene.getAnimator(i);
                auto play = [&](const char* name, float sp) {
                    const bool r1 = a.play(name, sp), r2 = refs[i]->play(name, sp);
                    CHECK(r1 == r2, "frame %d animator %ld play(%s): %d vs %d", frame, (long)i, name, (int)r1, (int)r2);
                };
                if (frame == (int)(i % 7)) play(states[i % 4], 1.f);
                else if (frame == 18 + (int)(i % 5)) play(states[(i + 1) % 4], 1.f);
                else if (frame == 26 && i % 5 == 0) play(states[(i + 1) % 4], 1.75f);   // current state again: false, speed changes
                else if (frame == 30 && i % 9 == 0) play("run", 1.f);                   // C16
                else if (frame == 33 && i % 6 == 1) { a.stop(); refs[i]->stop(); }
                else if (frame == 41 && i % 4 == 2) play("broken", 1.f);
                if (frame % 8 == (int)(i % 8)) {
                    const float bx = std::sin(0.13f * (float)(i + frame)), by = std::cos(0.05f * (float)(i * 2 + frame));
                    a.setBlendPosition(glm::vec2(bx, by));
                    refs[i]->setBlendPosition({bx, by});
                }
                if (frame == 12 && i % 3 == 1) { a.setPlaybackSpeed(1.5f); refs[i]->setPlaybackSpeed(1.5f); }
                if (frame == 5 && i % 10 == 3) a.pause();
            }
            const auto r = scene.update(1 / 30.f);
            int refErrors = 0;
            for (int64_t i = 0; i < n; ++i) {
                std::string err;
                refErrors += !refs[i]->update(1 / 30.f, err);
            }
            CHECK(r.has_value() == (refErrors == 0), "frame %d: scene update ok=%d, reference errors=%d", frame, (int)r.has_value(), refErrors);
            if (!r) {
                const long lines = 1 + std::count(r.error().begin(), r.error().end(), '\n');
                CHECK(lines == refErrors && r.error().rfind("error in the blending job", 0) == 0, "frame %d: %ld error lines vs %d", frame, lines, refErrors);
            }
            for (int64_t i = 0; i < n; ++i) {
                auto& a = scene.getAnimator(i);
                if (i % 2 == 0) CHECK(logs[(size_t)i].events == refs[i]->events, "frame %d animator %ld: listener events differ (%zu vs %zu)", frame, (long)i, logs[(size_t)i].events.size(), refs[i]->events.size());
                if ((i + frame) % 3) continue;
                const auto cs = a.getCurrentState();
                const auto* rs = refs[i]->currentState();
                CHECK((cs != nullptr) == (rs != nullptr), "frame %d animator %ld: current state presence", frame, (long)i);
                if (cs && rs) {
                    ++stateChecks;
                    CHECK(cs->getName() == rs->getName() && cs->getNormalizedTime() == rs->getNormalizedTime(), "frame %d animator %ld: state %s@%g vs %s@%g", frame, (long)i,
                          std::string(cs->getName()).c_str(), cs->getNormalizedTime(), rs->getName().c_str(), rs->getNormalizedTime());
                    CHECK(std::fabs(cs->getDuration() - rs->getDuration()) <= 1e-6f * rs->getDuration(), "frame %d animator %ld: duration %g vs %g", frame, (long)i, cs->getDuration(), rs->getDuration());
                }
                const auto ct = a.getCurrentTransition();
                auto* rt = const_cast<orc::RefTransition*>(refs[i]->currentTransition());
                CHECK((ct != nullptr) == (rt != nullptr), "frame %d animator %ld: transition presence", frame, (long)i);
                if (ct && rt) {
                    ++transitionChecks;
                    CHECK(ct->getNormalizedTime() == rt->getNormalizedTime() && ct->getPreviousState().getName() == rt->previous().getName() &&
                              ct->getCurrentState().getName() == rt->current().getName() && ct->getPreviousState().getNormalizedTime() == rt->previous().getNormalizedTime(),
                          "frame %d animator %ld: transition view", frame, (long)i);
                }
                // Stopped without _state, which includes a running transition (C10)
                const bool hasState = rs && !rt;
                const auto expected = !hasState ? SkeletalAnimatorPlaybackState::Stopped : (i % 10 == 3 && frame >= 5 ? SkeletalAnimatorPlaybackState::Paused : SkeletalAnimatorPlaybackState::Playing);
                CHECK(a.getPlaybackState() == expected, "frame %d animator %ld: playback state", frame, (long)i);
            }
            if (frame % 9 != 8) continue;
            std::vector<float> flipped(16);
            for (int64_t i = 0; i < n; i += 7) {
                auto& a = scene.getAnimator(i);
                for (int j = 0; j < J; j += 4) {
                    const glm::mat4 m = a.getJointModelMatrix(o.skel->names[j]);
                    orc_flip_handedness(refs[i]->models().data() + (size_t)j * 16, flipped.data());
                    for (int e = 0; e < 16; ++e) CHECK(close(glm::value_ptr(m)[e], flipped[e]), "frame %d animator %ld joint %d (device animators)", frame, (long)i, j);
                }
            }
        }
        for (const auto& l : logs) eventCount += (long)l.events.size();
        CHECK(stateChecks > 2000 && transitionChecks > 200 && eventCount > 400, "device scene scenario lost coverage (%ld, %ld, %ld)", stateChecks, transitionChecks, eventCount);
        std::printf("gpu: scene component with device animators, %ld animators x 56 frames: %ld state views, %ld transition views, %ld listener events\n", (long)n, stateChecks,
                    transitionChecks, eventCount);
    }
    std::printf("gpu: %d failures\n", g_failures);
    return g_failures ? 1 : 0;
}

// host cost of the per-character animator tick the device-side state machines replace (no device needed):
// N AnimatorCores of the samples/ozz definition in random states, one frame = one update() + afterUpdate() each
static int runHostBench(const std::string& dir, int n) {
    OracleAssets o = loadOracle(dir);
    SkeletalAnimatorDefinition def;
    orc::AnimatorDef rdef;
    buildDefinitions(def, rdef);
    std::vector<detail::ClipInfo> clips;
    for (const auto& a : o.named) clips.push_back({a.name, a.anim->duration});
    std::vector<std::unique_ptr<detail::AnimatorCore>> cores;
    const char* states[] = {"locomotion", "talk", "strafe", "greet"};
    uint32_t rng = 12345u;
    auto next = [&]() { rng = rng * 1664525u + 1013904223u; return rng >> 8; };
    for (int i = 0; i < n; ++i) {
        cores.push_back(std::make_unique<detail::AnimatorCore>(clips, def));
        cores.back()->play(states[next() % 4]);
        cores.back()->setBlendPosition(glm::vec2((float)(next() % 2001) / 1000.f - 1.f, (float)(next() % 2001) / 1000.f - 1.f));
    }
    detail::PoseRequest req;
    auto frame = [&](int f) {
        size_t layers = 0;
        for (int i = 0; i < n; ++i) {
            auto& c = *cores[(size_t)i];
            if (f % 10 == 0 && next() % 10 == 0) c.play(states[next() % 4]);
            const bool active = c.hasActiveStateOrTransition();
            if (c.update(1 / 60.f, req) && active) c.afterUpdate();
            layers += req.group0.layers.size() + req.group1.layers.size();
        }
        return layers;
    };
    for (int f = 0; f < 5; ++f) frame(f);
    const int frames = 30;
    size_t layers = 0;
    timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int f = 0; f < frames; ++f) layers += frame(5 + f);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    const double ms = ((double)(t1.tv_sec - t0.tv_sec) * 1e3 + (double)(t1.tv_nsec - t0.tv_nsec) * 1e-6) / frames;
    std::printf("hostbench: %d animators, %.2f ms per frame on one host thread (%.0f ns per animator tick, %.1f layers per animator)\n", n, ms,
                ms * 1e6 / n, (double)layers / frames / n);
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 3) { std::printf("usage: animator_test <assets_dir> cpu|gpu|gpuextras|hostbench [n]|fuzz [sequences] [seed]|devsim [characters] [frames] [seed]\n"); return 2; }
    if (std::string(argv[2]) == "hostbench") return runHostBench(argv[1], argc > 3 ? std::atoi(argv[3]) : 100000);
    if (std::string(argv[2]) == "devsim")
        return runDeviceSim(argv[1], argc > 3 ? std::atoll(argv[3]) : 1000, argc > 4 ? std::atoi(argv[4]) : 150, argc > 5 ? std::strtoull(argv[5], nullptr, 10) : 1);
    if (std::string(argv[2]) == "fuzz") return runFuzz(argv[1], argc > 3 ? std::atoi(argv[3]) : 50, argc > 4 ? std::strtoull(argv[4], nullptr, 10) : 1);
    std::setvbuf(stdout, nullptr, _IOLBF, 0);   // keep the progress lines if a section dies
    if (std::string(argv[2]) == "gpuextras") return runGpuExtras(argv[1]);
    return std::string(argv[2]) == "gpu" ? runGpu(argv[1]) : runCpu(argv[1]);
}
