// oracle/animator_ref.hpp -- TEST INFRASTRUCTURE (see dmk_oracle.h).
//
// CPU restatement of darmok's animator around the ozz jobs, used to check the host shim + device path:
//   OzzSkeletalAnimatorAnimationState  src/skeleton_ozz.cpp:357-451
//   OzzSkeletalAnimatorState           src/skeleton_ozz.cpp:453-656
//   OzzSkeletalAnimatorTransition      src/skeleton_ozz.cpp:658-781
//   SkeletalAnimatorImpl               src/skeleton_ozz.cpp:783-1087
//   calcBlendWeight(s), getTransition  src/skeleton.cpp:306-364, 575-599
//   Easing::apply                      include/darmok/easing.hpp
// It keeps the reference's data flow literally: every clip state owns a sampling context and a locals buffer, states
// and transitions own blended locals buffers, the animator owns the model matrices.
#pragma once

#include "dmk_oracle.h"

#include <memory>
#include <optional>
#include <string>
#include <vector>

namespace orc {

struct Vec2 {
    float x = 0.f, y = 0.f;
};

struct AnimationDef {
    std::string name;
    Vec2 blend_position;
    float speed = 0.f;
    bool loop = false;
};
struct TweenDef {
    int easing = 0;   // assets/protobuf/easing.proto
    float duration = 0.f;
};
struct StateDef {
    std::string name;
    std::vector<AnimationDef> animations;
    int blend = 0;    // 0 Cartesian, 1 Directional
    float threshold = 0.f;
    TweenDef tween;
    std::string next_state;
    float speed = 1.f;
};
struct TransitionDef {
    std::string src_state, dst_state;
    TweenDef tween;
    float offset = 0.f;
};
struct AnimatorDef {
    std::vector<StateDef> states;
    std::vector<TransitionDef> transitions;
};

float ease(int type, float position, float start, float end);
// Number of times darmok's zero-initialised locals of a state that was NEVER updated became the pose: a transition copies its
// previous state's locals when it is created (src/skeleton_ozz.cpp:658-665) and keeps showing them while its current state
// fails.  The product cannot express that pose (documented deviation); tests that generate random inputs
// stop a sequence when this counter moves.
extern int g_unsampledHolds;

struct NamedAnimation {
    std::string name;
    const orc_animation* anim;
};

class RefAnimationState {
public:
    RefAnimationState(const orc_skeleton* skel, const AnimationDef& def, const orc_animation* anim);
    RefAnimationState(RefAnimationState&&) noexcept;
    RefAnimationState& operator=(RefAnimationState&&) noexcept;
    ~RefAnimationState();
    float getDuration() const;
    bool hasFinished() const { return _looped && !_def.loop; }
    bool update(float dt, std::string& err);
    const std::vector<float>& getLocals() const { return _locals; }
    Vec2 getBlendPosition() const { return _def.blend_position; }
private:
    AnimationDef _def;
    float _normalizedTime = 0.f;
    bool _looped = false;
    const orc_animation* _animation;
    orc_sampling_context* _sampling;
    std::vector<float> _locals;   // SoaTransform[num_soa]
};

class RefState {
public:
    static std::optional<RefState> create(const orc_skeleton* skel, const StateDef& def, const std::vector<NamedAnimation>& anims);
    bool update(float dt, Vec2 blendPosition, std::string& err);
    bool updateImpl(float dt, Vec2 blendPosition, std::string& err);
    const std::string& getName() const { return _def.name; }
    const std::vector<float>& getLocals() const;
    bool hasLooped() const { return _looped; }
    bool everUpdated() const { return _everUpdated; }   // test bookkeeping only (see g_unsampledHolds)
    bool hasFinished() const;
    float getDuration() const { return _duration * _speed; }
    float getNormalizedTime() const { return _normalizedTime; }
    const std::string& getNextState() const { return _def.next_state; }
    void setSpeed(float s) { _speed = s; }
    void afterFinished() { _speed = _def.speed; }
    std::vector<float> calcBlendWeights(Vec2 pos) const;
private:
    RefState(const orc_skeleton* skel, const StateDef& def, std::vector<RefAnimationState> states);
    float calcDuration() const;
    float calcBlendWeight(Vec2 pos, Vec2 animPos) const;
    const orc_skeleton* _skel;
    StateDef _def;
    std::vector<RefAnimationState> _animationStates;
    Vec2 _blendPos, _oldBlendPos;
    float _normalizedTime = 0.f, _normalizedTweenTime = 0.f, _duration = 0.f;
    bool _looped = false;
    bool _everUpdated = false;
    float _speed = 1.f;
    std::vector<float> _locals;
};

class RefTransition {
public:
    RefTransition(const orc_skeleton* skel, const TransitionDef& def, RefState current, RefState previous);
    bool update(float dt, Vec2 blendPosition, std::string& err);
    bool hasFinished() const { return _normalizedTime >= 1.f; }
    RefState finish() { return std::move(_currentState); }
    const std::vector<float>& getLocals() const { return _locals; }
    RefState& current() { return _currentState; }
    RefState& previous() { return _previousState; }
    float getNormalizedTime() const { return _normalizedTime; }
private:
    const orc_skeleton* _skel;
    TransitionDef _def;
    RefState _currentState, _previousState;
    std::vector<float> _locals;
    float _normalizedTime = 0.f;
    bool _startedFromZeroLocals = false, _everBlended = false;   // test bookkeeping only (see g_unsampledHolds)
};

class RefAnimator {
public:
    RefAnimator(const orc_skeleton* skel, std::vector<NamedAnimation> anims, AnimatorDef def);
    bool play(const std::string& name, float stateSpeed = 1.f);
    void stop() { _state.reset(); }
    void pause() { _paused = !_paused; }                            // src/skeleton_ozz.cpp:875-878: a toggle nothing else reads
    int getPlaybackState() const { return !_state ? 0 : _paused ? 2 : 1; }   // Stopped, Playing, Paused (:880-891)
    void setPlaybackSpeed(float s) { _speed = s; }
    void setBlendPosition(Vec2 p) { _blendPosition = p; }
    bool update(float dt, std::string& err);
    const std::vector<float>& models() const { return _models; }   // 16 * J, right-handed ozz matrices
    std::vector<std::string> events;                                // "started:<state>", "looped:...", "finished:...", "transition_started", "transition_finished"
    const RefState* currentState() const;
    const RefTransition* currentTransition() const { return _transition ? &*_transition : nullptr; }
    float getStateDuration(const std::string& name) const;
private:
    const StateDef* findState(const std::string& name) const;
    const TransitionDef* findTransition(const std::string& src, const std::string& dst) const;
    void afterUpdate();
    const orc_skeleton* _skel;
    std::vector<NamedAnimation> _anims;
    AnimatorDef _def;
    float _speed = 1.f;
    bool _paused = false;
    Vec2 _blendPosition;
    std::optional<RefTransition> _transition;
    std::optional<RefState> _state;
    std::vector<float> _models;
};

}  // namespace orc
