// animator_core.hpp -- host-side animator state machine.  Pure CPU control flow, no device calls: each tick it resolves
// the animator into a PoseRequest (which clips at which ratios with which weights, grouped as darmok's states and
// transition blend them) that the device executes as a dmk_pose_program.
//
// Mirrors, decision by decision (paths relative to the darmok repository; quirks C1..C16 of SURVEY.md Appendix C):
//   OzzSkeletalAnimatorAnimationState   src/skeleton_ozz.cpp:357-451    -> ClipState
//   OzzSkeletalAnimatorState            src/skeleton_ozz.cpp:453-656    -> StateMachineState
//   OzzSkeletalAnimatorTransition       src/skeleton_ozz.cpp:658-781    -> TransitionState
//   SkeletalAnimatorImpl                src/skeleton_ozz.cpp:783-1087   -> AnimatorCore
//   calcBlendWeights / getTransition    src/skeleton.cpp:306-364, 575-599
#pragma once

#include "skeleton.hpp"

#include <functional>

namespace darmok::detail {

struct LayerRequest {
    int clip = 0;        // index into the animator's clip table (order of AnimatorCore::ClipInfo)
    float ratio = 0.f;
    float weight = 0.f;
};
struct GroupRequest {
    std::vector<LayerRequest> layers;
    bool passthrough = false;   // single-clip state, or a clip sitting exactly at the blend position
    bool zero = false;          // passthrough of a clip that has never been sampled: darmok's zero-initialised locals are the pose
    int rest = 0;               // blend: rest-pose layer; passthrough: the layer that is the pose
    float threshold = 0.f;
};
struct PoseRequest {
    enum Top { Group0, Transition, Skip } top = Skip;
    GroupRequest group0, group1;
    float w0 = 1.f, w1 = 0.f;
};

struct ClipInfo {
    std::string name;
    float duration = 0.f;
};

class ClipState final {
public:
    ClipState(const SkeletalAnimatorAnimationDefinition& def, int clip, float clipDuration) noexcept;
    float getDuration() const noexcept;
    bool hasFinished() const noexcept { return _looped && !_def.loop; }
    bool hasLooped() const noexcept { return _looped; }
    void update(float deltaTime) noexcept;
    float getNormalizedTime() const noexcept { return _normalizedTime; }
    const glm::vec2& getBlendPosition() const noexcept { return _def.blend_position; }
    int clip() const noexcept { return _clip; }
    // false while the clip has never been sampled: darmok's AnimationState then still holds its zero-initialised locals
    // (a non-looping clip that runs past its end on its very first tick is never sampled at all, src/skeleton_ozz.cpp:411-428)
    bool sampledOnce() const noexcept { return _sampledOnce; }
private:
    SkeletalAnimatorAnimationDefinition _def;
    int _clip;
    float _clipDuration;
    float _normalizedTime = 0.f;
    bool _looped = false;
    bool _sampledOnce = false;
};

class StateMachineState final : public ISkeletalAnimatorState {
public:
    static std::optional<StateMachineState> create(const SkeletalAnimatorStateDefinition& def, const std::vector<ClipInfo>& clips) noexcept;
    expected<void, std::string> update(float deltaTime, const glm::vec2& blendPosition, GroupRequest& out) noexcept;
    std::string_view getName() const noexcept override { return _def.name; }
    float getNormalizedTime() const noexcept override { return _normalizedTime; }
    float getDuration() const noexcept override { return _duration * _speed; }   // multiplies (C2)
    bool hasLooped() const noexcept { return _looped; }
    bool hasFinished() const noexcept;
    const std::string& getNextState() const noexcept { return _def.next_state; }
    void setSpeed(float speed) noexcept { _speed = speed; }
    void afterFinished() noexcept { _speed = _def.speed; }
    // blend weights of the state's clips for a blend position (calcBlendWeights, src/skeleton.cpp:346-364)
    std::vector<float> calcBlendWeights(const glm::vec2& pos) const noexcept;
    void calcBlendWeights(const glm::vec2& pos, std::vector<float>& weights) const noexcept;   // same, into a reused buffer
    // what getLocals() holds (src/skeleton_ozz.cpp:617-624): the group of the last update that succeeded; null before the first
    const GroupRequest* lastGood() const noexcept { return _hasLastGood ? &_lastGood : nullptr; }
private:
    StateMachineState(const SkeletalAnimatorStateDefinition& def, std::vector<ClipState> clips) noexcept;
    expected<void, std::string> updateImpl(float deltaTime, const glm::vec2& blendPosition, GroupRequest& out) noexcept;
    GroupRequest _lastGood;
    bool _hasLastGood = false;
    std::vector<float> _animLengths, _animAngles;   // Directional: per-definition terms of calcBlendWeight (built once)
    std::vector<float> _weights;                    // scratch of update()
    float calcDuration() const noexcept;
    float calcBlendWeight(const glm::vec2& pos, const glm::vec2& animPos) const noexcept;
    SkeletalAnimatorStateDefinition _def;
    std::vector<ClipState> _clips;
    glm::vec2 _blendPos{0.f, 0.f}, _oldBlendPos{0.f, 0.f};
    float _normalizedTime = 0.f, _normalizedTweenTime = 0.f, _duration = 0.f;
    bool _looped = false;
    float _speed = 1.f;
};

class TransitionState final : public ISkeletalAnimatorTransition {
public:
    TransitionState(const SkeletalAnimatorTransitionDefinition& def, StateMachineState current, StateMachineState previous) noexcept;
    expected<void, std::string> update(float deltaTime, const glm::vec2& blendPosition, PoseRequest& out) noexcept;
    float getDuration() const noexcept override { return _def.tween.duration; }
    float getNormalizedTime() const noexcept override { return _normalizedTime; }
    bool hasFinished() const noexcept { return _normalizedTime >= 1.f; }
    StateMachineState finish() noexcept { return std::move(_current); }
    const ISkeletalAnimatorState& getCurrentState() const noexcept override { return _current; }
    const ISkeletalAnimatorState& getPreviousState() const noexcept override { return _previous; }
    StateMachineState& current() noexcept { return _current; }
    StateMachineState& previous() noexcept { return _previous; }
private:
    SkeletalAnimatorTransitionDefinition _def;
    StateMachineState _current, _previous;
    float _normalizedTime = 0.f;
    bool _everBlended = false;   // the transition's BlendingJob has run at least once: _locals holds its last output
    GroupRequest _lastCurrent;   // the current state's last resolved group (a finished transition keeps showing it)
    GroupRequest _initialPrevious;      // what _locals starts as: the previous state's locals when the transition was created (:658-665)
    bool _hasInitialPrevious = false;   // false: that state had never been updated (zero-initialised locals in darmok)
};

struct AnimatorEvents {   // listener fan-out, bound by SkeletalAnimator
    std::function<void(std::string_view)> stateLooped, stateFinished, stateStarted;
    std::function<void()> transitionFinished, transitionStarted;
};

class AnimatorCore final {
public:
    AnimatorCore(std::vector<ClipInfo> clips, SkeletalAnimatorDefinition def) noexcept;
    AnimatorEvents events;

    bool play(std::string_view name, float stateSpeed = 1.f) noexcept;
    void stop() noexcept { _state.reset(); }
    void pause() noexcept { _paused = !_paused; }
    bool isPaused() const noexcept { return _paused; }
    SkeletalAnimatorPlaybackState getPlaybackState() const noexcept;
    void setPlaybackSpeed(float speed) noexcept { _speed = speed; }
    float getPlaybackSpeed() const noexcept { return _speed; }
    void setBlendPosition(const glm::vec2& value) noexcept { _blendPosition = value; }
    const glm::vec2& getBlendPosition() const noexcept { return _blendPosition; }
    const SkeletalAnimatorDefinition& getDefinition() const noexcept { return _def; }
    const ISkeletalAnimatorState* getCurrentState() const noexcept;
    const ISkeletalAnimatorTransition* getCurrentTransition() const noexcept;
    float getStateDuration(const std::string& name) const noexcept;
    const std::vector<ClipInfo>& clips() const noexcept { return _clips; }

    // one tick: fills `out`; afterUpdate() must be called once the pose has been applied (listener events, auto next state)
    bool hasActiveStateOrTransition() const noexcept { return _transition.has_value() || _state.has_value(); }
    expected<void, std::string> update(float deltaTime, PoseRequest& out) noexcept;
    void afterUpdate() noexcept;

private:
    const SkeletalAnimatorStateDefinition* findState(std::string_view name) const noexcept;
    const SkeletalAnimatorTransitionDefinition* findTransition(std::string_view src, std::string_view dst) const noexcept;
    StateMachineState* currentStateMutable() noexcept;
    std::vector<ClipInfo> _clips;
    SkeletalAnimatorDefinition _def;
    float _speed = 1.f;
    bool _paused = false;
    glm::vec2 _blendPosition{0.f, 0.f};
    std::optional<TransitionState> _transition;
    std::optional<StateMachineState> _state;
};

}  // namespace darmok::detail
