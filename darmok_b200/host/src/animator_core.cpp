// animator_core.cpp -- see animator_core.hpp.  Host control flow only; citations are to the darmok repository.
#include "darmok_b200/animator_core.hpp"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <numeric>

namespace darmok::detail {

// ---- ClipState: OzzSkeletalAnimatorAnimationState minus the sampling (which the device does) ---------------------
ClipState::ClipState(const SkeletalAnimatorAnimationDefinition& def, int clip, float clipDuration) noexcept
    : _def{def}, _clip{clip}, _clipDuration{clipDuration} {}

float ClipState::getDuration() const noexcept {
    // src/skeleton_ozz.cpp:398-404
    float speed = _def.speed;
    speed = speed == 0.f ? 1.f : speed;
    return _clipDuration / speed;
}

void ClipState::update(float deltaTime) noexcept {
    // src/skeleton_ozz.cpp:411-428.  A finished non-looping clip is not re-sampled: its ratio simply stops moving, and
    // sampling is a pure function of the ratio, so the device reproduces the held pose.
    if (hasFinished()) return;
    const float normDeltaTime = deltaTime / getDuration();
    const float normTime = _normalizedTime + normDeltaTime;
    _looped = normTime > 1.f;
    if (hasFinished()) return;
    _normalizedTime = std::fmod(normTime, 1.f);   // setNormalizedTime (:392-396)
    _sampledOnce = true;                          // SamplingJob runs from here on (:429-437)
}

// ---- StateMachineState: OzzSkeletalAnimatorState --------------------------------------------------------------------
StateMachineState::StateMachineState(const SkeletalAnimatorStateDefinition& def, std::vector<ClipState> clips) noexcept
    : _def{def}, _clips{std::move(clips)}, _speed{def.speed} {
    _duration = calcDuration();
    if (_def.blend == SkeletalAnimatorStateDefinition::Directional) {
        // the terms of calcBlendWeight that only depend on the definition: |blend_position| and the angle between two of them
        const size_t m = _def.animations.size();
        _animLengths.resize(m);
        _animAngles.resize(m * m);
        for (size_t i = 0; i < m; ++i) {
            _animLengths[i] = glm::length(_def.animations[i].blend_position);
            for (size_t j = 0; j < m; ++j) _animAngles[i * m + j] = glm::angle(_def.animations[j].blend_position, _def.animations[i].blend_position);
        }
    }
}

std::optional<StateMachineState> StateMachineState::create(const SkeletalAnimatorStateDefinition& def,
                                                           const std::vector<ClipInfo>& clips) noexcept {
    // src/skeleton_ozz.cpp:469-485: animations the provider does not know are dropped; a state without any is invalid
    std::vector<ClipState> states;
    states.reserve(def.animations.size());
    for (const auto& animDef : def.animations) {
        const auto itr = std::find_if(clips.begin(), clips.end(), [&](const ClipInfo& c) { return c.name == animDef.name; });
        if (itr != clips.end()) states.emplace_back(animDef, static_cast<int>(itr - clips.begin()), itr->duration);
    }
    if (states.empty()) return std::nullopt;
    return StateMachineState{def, std::move(states)};
}

float StateMachineState::calcDuration() const noexcept {
    // src/skeleton_ozz.cpp:487-518: least common multiple of the clip durations in decimal fixed point.
    // For a duration with no fractional part the reference evaluates ceil(-log10(0)) = +inf and converts it to int,
    // which x86 turns into INT_MIN so that pow(10, INT_MIN) = 0 and the factor stays 1; that outcome is kept here
    // without the undefined conversion.
    if (_clips.empty()) return 0.f;
    int f = 1;
    for (const auto& clip : _clips) {
        const float d = clip.getDuration();
        const float frac = d - std::floor(d);
        if (!(frac > 0.f)) continue;
        const int decimalPlaces = static_cast<int>(std::ceil(-std::log10(frac)));
        f = std::max(f, static_cast<int>(std::pow(10, decimalPlaces)));
    }
    std::vector<int> scaled;
    for (const auto& clip : _clips) scaled.push_back(static_cast<int>(std::round(clip.getDuration() * f)));
    int d = scaled[0];
    for (size_t i = 1; i < scaled.size(); ++i) d = std::lcm(d, scaled[i]);
    return static_cast<float>(d) / f;
}

bool StateMachineState::hasFinished() const noexcept {
    for (const auto& clip : _clips)
        if (!clip.hasFinished()) return false;
    return true;
}

float StateMachineState::calcBlendWeight(const glm::vec2& pos, const glm::vec2& animPos) const noexcept {
    // src/skeleton.cpp:306-344 (gradient bands; the j == i term is 0/0 = NaN and drops out because NaN < w is false).
    // Stated as the reference states it; update() goes through calcBlendWeights, which evaluates the same expressions with
    // the per-definition terms (|blend_position|, angle between two blend positions) taken from tables built once.
    float w = FLT_MAX;   // bx::kFloatLargest
    if (_def.blend == SkeletalAnimatorStateDefinition::Directional) {
        const float lenAnim = glm::length(animPos);
        const float len = glm::length(pos);
        for (const auto& anim : _def.animations) {
            const glm::vec2 blendPos = anim.blend_position;
            const float lenAnim2 = glm::length(blendPos);
            const float p = lenAnim + lenAnim2;
            const glm::vec2 a{(len - lenAnim) * p, glm::angle(pos, animPos)};
            const glm::vec2 b{(lenAnim2 - lenAnim) * p, glm::angle(blendPos, animPos)};
            const float f = 1.F - (glm::dot(a, b) / glm::length2(b));
            if (f < w) w = f;
        }
    } else {
        const glm::vec2 a = pos - animPos;
        for (const auto& anim : _def.animations) {
            const glm::vec2 b = anim.blend_position - animPos;
            const float f = 1.F - (glm::dot(a, b) / glm::length2(b));
            if (f < w) w = f;
        }
    }
    return w;
}

void StateMachineState::calcBlendWeights(const glm::vec2& pos, std::vector<float>& weights) const noexcept {
    // calcBlendWeights (src/skeleton.cpp:346-364) over ALL animations of the definition, known to the provider or not
    const size_t m = _def.animations.size();
    weights.resize(m);
    if (_def.blend == SkeletalAnimatorStateDefinition::Directional) {
        const float len = glm::length(pos);
        for (size_t i = 0; i < m; ++i) {
            const float lenAnim = _animLengths[i];
            const float angle = glm::angle(pos, _def.animations[i].blend_position);   // the one term that depends on the tick
            const float* angles = &_animAngles[i * m];
            float w = FLT_MAX;
            for (size_t j = 0; j < m; ++j) {
                const float lenAnim2 = _animLengths[j];
                const float p = lenAnim + lenAnim2;
                const glm::vec2 a{(len - lenAnim) * p, angle};
                const glm::vec2 b{(lenAnim2 - lenAnim) * p, angles[j]};
                const float f = 1.F - (glm::dot(a, b) / glm::length2(b));
                if (f < w) w = f;
            }
            weights[i] = w;
        }
        return;
    }
    for (size_t i = 0; i < m; ++i) weights[i] = calcBlendWeight(pos, _def.animations[i].blend_position);
}

std::vector<float> StateMachineState::calcBlendWeights(const glm::vec2& pos) const noexcept {
    std::vector<float> weights;
    calcBlendWeights(pos, weights);
    return weights;
}

expected<void, std::string> StateMachineState::update(float deltaTime, const glm::vec2& blendPosition, GroupRequest& out) noexcept {
    auto r = updateImpl(deltaTime, blendPosition, out);
    if (r) {   // getLocals() now holds this group's result; a failed BlendingJob leaves the state's locals as they were
        _lastGood = out;
        _hasLastGood = true;
    }
    return r;
}

expected<void, std::string> StateMachineState::updateImpl(float deltaTime, const glm::vec2& blendPosition, GroupRequest& out) noexcept {
    // src/skeleton_ozz.cpp:535-610
    deltaTime *= _speed;
    for (auto& clip : _clips) clip.update(deltaTime);

    const float normDeltaTime = deltaTime / _duration;
    _normalizedTime += normDeltaTime;
    _looped = _normalizedTime > 1.f;
    _normalizedTime = std::fmod(_normalizedTime, 1.f);

    out.layers.clear();   // keeps the caller's capacity: no allocation per tick once a request has been used
    out.passthrough = false;
    out.zero = false;
    out.rest = 0;
    out.threshold = _def.threshold;
    for (const auto& clip : _clips) out.layers.push_back({clip.clip(), clip.getNormalizedTime(), 0.f});
    if (_clips.size() <= 1) {   // getLocals() returns the single clip's buffer (:617-624)
        out.passthrough = true;
        out.rest = 0;
        out.zero = !_clips[0].sampledOnce();
        return {};
    }
    if (_blendPos != blendPosition) {
        const float blendFactor = ease01(_def.tween.easing, _normalizedTweenTime);
        _oldBlendPos = _oldBlendPos + blendFactor * (_blendPos - _oldBlendPos);
        _normalizedTweenTime = 0.f;
        _blendPos = blendPosition;
    }
    _normalizedTweenTime += deltaTime / _def.tween.duration;
    if (_normalizedTweenTime >= 1.f) _normalizedTweenTime = 1.f;
    const float blendFactor = ease01(_def.tween.easing, _normalizedTweenTime);
    const glm::vec2 pos = _oldBlendPos + blendFactor * (_blendPos - _oldBlendPos);

    calcBlendWeights(pos, _weights);
    const std::vector<float>& weights = _weights;
    int rest = -1;
    size_t i = 0;
    for (const auto& clip : _clips) {
        if (clip.getBlendPosition() == glm::vec2(0.f)) rest = static_cast<int>(i);   // blend position (0, 0) is the rest pose
        if (clip.getBlendPosition() == pos) {   // early exit: one clip sits exactly at the blend position
            out.passthrough = true;
            out.rest = static_cast<int>(i);
            out.zero = !clip.sampledOnce();   // the copy of a buffer that was never written
            return {};
        }
        if (i < weights.size()) out.layers[i].weight = weights[i];
        ++i;
    }
    out.rest = rest < 0 ? 0 : rest;
    // BlendingJob::Validate: threshold > 0 (C7)
    if (!(out.threshold > 0.f)) return unexpected<std::string>{"error in the blending job"};
    return {};
}

// ---- TransitionState: OzzSkeletalAnimatorTransition -------------------------------------------------------------------
TransitionState::TransitionState(const SkeletalAnimatorTransitionDefinition& def, StateMachineState current, StateMachineState previous) noexcept
    : _def{def}, _current{std::move(current)}, _previous{std::move(previous)} {
    // _locals{ _previousState.getLocals() } (src/skeleton_ozz.cpp:658-665): a copy taken now, not a view
    if (const GroupRequest* g = _previous.lastGood()) {
        _initialPrevious = *g;
        _hasInitialPrevious = true;
    }
}

expected<void, std::string> TransitionState::update(float deltaTime, const glm::vec2& blendPosition, PoseRequest& out) noexcept {
    // src/skeleton_ozz.cpp:682-726
    if (hasFinished()) {   // locals = copy of the current state's locals; the states no longer tick
        if (_lastCurrent.layers.empty()) {
            out.top = PoseRequest::Skip;
        } else {
            out.top = PoseRequest::Group0;
            out.group0 = _lastCurrent;
        }
        return {};
    }
    out.top = PoseRequest::Transition;
    if (auto r = _previous.update(deltaTime, blendPosition, out.group0); !r) return r;
    float currentStateDeltaTime = deltaTime;
    if (_normalizedTime == 0.f) currentStateDeltaTime += _def.offset;
    if (auto r = _current.update(currentStateDeltaTime, blendPosition, out.group1); !r) {
        // a failing current state is swallowed (C8): the BlendingJob is not run and _locals stays what it was -- the last
        // blend's output (same matrices as last tick), or, before the first blend, the copy of the previous state's locals
        // taken at construction.  A previous state that had never been updated leaves zero-initialised locals in darmok;
        // that pose cannot be expressed here and the matrices are kept (documented deviation).
        if (!_everBlended && _hasInitialPrevious) {
            out.top = PoseRequest::Group0;
            out.group0 = _initialPrevious;
        } else {
            out.top = PoseRequest::Skip;
        }
        return {};
    }
    _everBlended = true;
    _lastCurrent = out.group1;
    _normalizedTime += deltaTime / getDuration();
    const float v = ease01(_def.tween.easing, _normalizedTime);
    out.w0 = 1.f - v;
    out.w1 = v;
    return {};
}

// ---- AnimatorCore: SkeletalAnimatorImpl ---------------------------------------------------------------------------------
AnimatorCore::AnimatorCore(std::vector<ClipInfo> clips, SkeletalAnimatorDefinition def) noexcept
    : _clips{std::move(clips)}, _def{std::move(def)} {}

const SkeletalAnimatorStateDefinition* AnimatorCore::findState(std::string_view name) const noexcept {
    const auto itr = std::find_if(_def.states.begin(), _def.states.end(), [name](const auto& s) { return s.name == name; });
    return itr == _def.states.end() ? nullptr : &*itr;
}

const SkeletalAnimatorTransitionDefinition* AnimatorCore::findTransition(std::string_view src, std::string_view dst) const noexcept {
    // src/skeleton.cpp:575-599: exact, src>*, *>dst, *>*
    const auto& ts = _def.transitions;
    auto find = [&](auto pred) -> const SkeletalAnimatorTransitionDefinition* {
        const auto itr = std::find_if(ts.begin(), ts.end(), pred);
        return itr == ts.end() ? nullptr : &*itr;
    };
    if (auto t = find([&](const auto& t) { return t.src_state == src && t.dst_state == dst; })) return t;
    if (auto t = find([&](const auto& t) { return t.src_state == src && t.dst_state.empty(); })) return t;
    if (auto t = find([&](const auto& t) { return t.src_state.empty() && t.dst_state == dst; })) return t;
    return find([&](const auto& t) { return t.src_state.empty() && t.dst_state.empty(); });
}

StateMachineState* AnimatorCore::currentStateMutable() noexcept {
    if (_transition) return &_transition->current();
    if (_state) return &*_state;
    return nullptr;
}
const ISkeletalAnimatorState* AnimatorCore::getCurrentState() const noexcept {
    if (_transition) return &_transition->getCurrentState();
    if (_state) return &*_state;
    return nullptr;
}
const ISkeletalAnimatorTransition* AnimatorCore::getCurrentTransition() const noexcept {
    return _transition ? &*_transition : nullptr;
}

float AnimatorCore::getStateDuration(const std::string& name) const noexcept {
    if (const auto* stateDef = findState(name))
        if (auto state = StateMachineState::create(*stateDef, _clips)) return state->getDuration();
    return 0.f;
}

SkeletalAnimatorPlaybackState AnimatorCore::getPlaybackState() const noexcept {
    // pausing only changes this answer: update() never reads _paused (C10)
    if (!_state) return SkeletalAnimatorPlaybackState::Stopped;
    if (_paused) return SkeletalAnimatorPlaybackState::Paused;
    return SkeletalAnimatorPlaybackState::Playing;
}

bool AnimatorCore::play(std::string_view name, float stateSpeed) noexcept {
    // src/skeleton_ozz.cpp:783-856 (C9)
    if (auto* currState = currentStateMutable(); currState && currState->getName() == name) {
        currState->setSpeed(stateSpeed);
        return false;
    }
    const auto* stateDef = findState(name);
    if (!stateDef) return false;
    std::optional<StateMachineState> prevState;
    if (_state) {
        prevState.emplace(std::move(*_state));
        _state.reset();
    }
    if (_transition) {
        prevState.emplace(_transition->finish());
        _transition.reset();
    }
    if (prevState) {
        if (const auto* transConfig = findTransition(prevState->getName(), name)) {
            auto state = StateMachineState::create(*stateDef, _clips);
            if (!state) return false;
            _transition.emplace(*transConfig, std::move(*state), std::move(*prevState));
            if (events.stateStarted) events.stateStarted(_transition->getCurrentState().getName());
            if (events.transitionStarted) events.transitionStarted();
            return true;
        }
    }
    // (the reference's "state finished" branch here tests _state right after resetting it and cannot run)
    auto state = StateMachineState::create(*stateDef, _clips);
    if (!state) return false;
    _state = std::move(state);
    if (events.stateStarted) events.stateStarted(_state->getName());
    return true;
}

expected<void, std::string> AnimatorCore::update(float deltaTime, PoseRequest& out) noexcept {
    // src/skeleton_ozz.cpp:1006-1042
    deltaTime *= _speed;
    out.top = PoseRequest::Skip;   // reset in place (the groups' layer vectors keep their capacity)
    out.group0.layers.clear();
    out.group1.layers.clear();
    out.group0.passthrough = out.group1.passthrough = false;
    out.group0.zero = out.group1.zero = false;
    out.group0.rest = out.group1.rest = 0;
    out.group0.threshold = out.group1.threshold = 0.f;
    out.w0 = 1.f;
    out.w1 = 0.f;
    if (_transition) return _transition->update(deltaTime, _blendPosition, out);
    if (_state) {
        out.top = PoseRequest::Group0;
        return _state->update(deltaTime, _blendPosition, out.group0);
    }
    out.top = PoseRequest::Skip;
    return {};
}

void AnimatorCore::afterUpdate() noexcept {
    // src/skeleton_ozz.cpp:1044-1087 (C11); only reached when a transition or a state was updated successfully
    if (_transition && _transition->hasFinished()) {
        if (events.transitionFinished) events.transitionFinished();
        if (events.stateFinished) events.stateFinished(_transition->getPreviousState().getName());
        if (_state) _state->afterFinished();
        _state.emplace(_transition->finish());
        _transition.reset();
    }
    if (_state) {
        if (_state->hasLooped() && events.stateLooped) events.stateLooped(_state->getName());
        if (_state->hasFinished()) {
            if (events.stateFinished) events.stateFinished(_state->getName());
            _state->afterFinished();
            const std::string nextState = _state->getNextState();
            if (!nextState.empty()) play(nextState);
        }
    }
}

}  // namespace darmok::detail
