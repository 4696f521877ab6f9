// oracle/animator_ref.cpp -- TEST INFRASTRUCTURE (see animator_ref.hpp and dmk_oracle.h).
#include "animator_ref.hpp"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <numeric>

namespace orc {

// ---- Easing::apply (include/darmok/easing.hpp), general (start, end) form -------------------------------------------
namespace {
const float kHalfPi = 1.57079632679489661923f, kPi = 3.14159265358979323846f, kTwoPi = 6.28318530717958647692f;

float bounceOut(float position, float start, float end) {
    const float c = end - start;
    if (position < (1 / 2.75f)) return c * (7.5625f * position * position) + start;
    if (position < (2.0f / 2.75f)) {
        const float postFix = position -= (1.5f / 2.75f);
        return c * (7.5625f * (postFix)*position + .75f) + start;
    }
    if (position < (2.5f / 2.75f)) {
        const float postFix = position -= (2.25f / 2.75f);
        return c * (7.5625f * (postFix)*position + .9375f) + start;
    }
    const float postFix = position -= (2.625f / 2.75f);
    return c * (7.5625f * (postFix)*position + .984375f) + start;
}
float bounceIn(float position, float start, float end) { return (end - start) - bounceOut((1 - position), 0.f, (end - start)) + start; }
}  // namespace

int g_unsampledHolds = 0;

float ease(int type, float position, float start, float end) {
    const float d = end - start;
    switch (type) {
    case 1: return start;                                                    // Stepped
    case 2: return d * position * position + start;                          // QuadraticIn
    case 3: return (-d) * position * (position - 2) + start;                 // QuadraticOut
    case 4:                                                                  // QuadraticInOut
        position *= 2;
        if (position < 1) return (d / 2) * position * position + start;
        --position;
        return (-d / 2) * (position * (position - 2) - 1) + start;
    case 5: return d * position * position * position + start;               // CubicIn
    case 6: --position; return d * (position * position * position + 1) + start;
    case 7:
        position *= 2;
        if (position < 1) return (d / 2) * position * position * position + start;
        position -= 2;
        return (d / 2) * (position * position * position + 2) + start;
    case 8: return d * position * position * position * position + start;    // QuarticIn
    case 9: --position; return -d * (position * position * position * position - 1) + start;
    case 10:
        position *= 2;
        if (position < 1) return (d / 2) * (position * position * position * position) + start;
        position -= 2;
        return (-d / 2) * (position * position * position * position - 2) + start;
    case 11: return d * position * position * position * position * position + start;   // QuinticIn
    case 12: position--; return d * (position * position * position * position * position + 1) + start;
    case 13:
        position *= 2;
        if (position < 1) return (d / 2) * (position * position * position * position * position) + start;
        position -= 2;
        return (d / 2) * (position * position * position * position * position + 2) + start;
    case 14: return -d * std::cos(position * kHalfPi) + d + start;           // SinusoidalIn
    case 15: return d * std::sin(position * kHalfPi) + start;
    case 16: return (-d / 2) * (std::cos(position * kPi) - 1) + start;
    case 17: return d * std::pow(2.f, 10 * (position - 1)) + start;          // ExponentialIn
    case 18: return d * (-std::pow(2.f, -10 * position) + 1) + start;
    case 19:
        position *= 2;
        if (position < 1) return (d / 2) * std::pow(2.f, 10 * (position - 1)) + start;
        --position;
        return (d / 2) * (-std::pow(2.f, -10 * position) + 2) + start;
    case 20: return -d * (std::sqrt(1 - position * position) - 1) + start;   // CircularIn
    case 21: --position; return d * (std::sqrt(1 - position * position)) + start;
    case 22:
        position *= 2;
        if (position < 1) return (-d / 2) * (std::sqrt(1 - position * position) - 1) + start;
        position -= 2;
        return (d / 2) * (std::sqrt(1 - position * position) + 1) + start;
    case 23: return bounceIn(position, start, end);
    case 24: return bounceOut(position, start, end);
    case 25:
        if (position < 0.5f) return bounceIn(position * 2, 0.f, d) * .5f + start;
        return bounceOut((position * 2 - 1), 0.f, d) * .5f + d * .5f + start;
    case 26: {                                                               // ElasticIn
        if (position <= 0.00001f) return start;
        if (position >= 0.999f) return end;
        const float p = .3f, s = p / 4;
        const float postFix = d * std::pow(2.f, 10 * (position -= 1));
        return -(postFix * std::sin((position - s) * kTwoPi / p)) + start;
    }
    case 27: {
        if (position <= 0.00001f) return start;
        if (position >= 0.999f) return end;
        const float p = .3f, s = p / 4;
        return d * std::pow(2.f, -10 * position) * std::sin((position - s) * kTwoPi / p) + end;
    }
    case 28: {
        if (position <= 0.00001f) return start;
        if (position >= 0.999f) return end;
        position *= 2;
        const float p = (.3f * 1.5f), s = p / 4;
        if (position < 1) {
            const float postFix = d * std::pow(2.f, 10 * (position -= 1));
            return -0.5f * (postFix * std::sin((position - s) * kTwoPi / p)) + start;
        }
        const float postFix = d * std::pow(2.f, -10 * (position -= 1));
        return postFix * std::sin((position - s) * kTwoPi / p) * .5f + end;
    }
    case 29: { const float s = 1.70158f; const float postFix = position; return d * (postFix)*position * ((s + 1) * position - s) + start; }
    case 30: { const float s = 1.70158f; position -= 1; return d * ((position)*position * ((s + 1) * position + s) + 1) + start; }
    case 31: {
        float s = 1.70158f, t = position;
        const float dd = 1;
        s *= (1.525f);
        if ((t /= dd / 2) < 1) return d / 2 * (t * t * (((s) + 1) * t - s)) + start;
        const float postFix = t -= 2;
        return d / 2 * ((postFix)*t * (((s) + 1) * t + s) + 2) + start;
    }
    case 0:
    default: return d * position + start;                                    // Linear
    }
}

namespace {
float tween(const TweenDef& t, float position) { return ease(t.easing, position, 0.F, 1.F); }   // calcTween, src/skeleton.cpp:269-272
float dot2(Vec2 a, Vec2 b) { return a.x * b.x + a.y * b.y; }
float length2(Vec2 v) { return dot2(v, v); }
float length(Vec2 v) { return std::sqrt(dot2(v, v)); }
float angle(Vec2 x, Vec2 y) {   // glm::angle: acos(clamp(dot(x, y), -1, 1))
    float d = dot2(x, y);
    d = std::min(std::max(d, -1.f), 1.f);
    return std::acos(d);
}
bool equal(Vec2 a, Vec2 b) { return a.x == b.x && a.y == b.y; }
}  // namespace

// ---- OzzSkeletalAnimatorAnimationState -------------------------------------------------------------------------------
RefAnimationState::RefAnimationState(const orc_skeleton* skel, const AnimationDef& def, const orc_animation* anim)
    : _def(def), _animation(anim), _sampling(orc_sampling_context_create(skel->num_joints)),
      _locals((size_t)skel->num_soa * ORC_SOA_FLOATS, 0.f) {}
RefAnimationState::RefAnimationState(RefAnimationState&& o) noexcept
    : _def(std::move(o._def)), _normalizedTime(o._normalizedTime), _looped(o._looped), _animation(o._animation), _sampling(o._sampling),
      _locals(std::move(o._locals)) {
    o._sampling = nullptr;
}
RefAnimationState& RefAnimationState::operator=(RefAnimationState&& o) noexcept {
    if (this != &o) {
        orc_sampling_context_free(_sampling);
        _def = std::move(o._def);
        _normalizedTime = o._normalizedTime;
        _looped = o._looped;
        _animation = o._animation;
        _sampling = o._sampling;
        _locals = std::move(o._locals);
        o._sampling = nullptr;
    }
    return *this;
}
RefAnimationState::~RefAnimationState() { orc_sampling_context_free(_sampling); }

float RefAnimationState::getDuration() const {
    float speed = _def.speed;
    speed = speed == 0.f ? 1.f : speed;
    return _animation->duration / speed;
}

bool RefAnimationState::update(float dt, std::string& err) {
    if (hasFinished()) return true;
    const float normDeltaTime = dt / getDuration();
    const float normTime = _normalizedTime + normDeltaTime;
    _looped = normTime > 1.f;
    if (hasFinished()) return true;   // (never sampled at all when this happens on the first tick: the locals stay zero)
    _normalizedTime = std::fmod(normTime, 1.f);
    if (!orc_sample(_animation, _sampling, _normalizedTime, _locals.data(), (int)(_locals.size() / ORC_SOA_FLOATS), nullptr)) {
        err = "error in the sampling job";
        return false;
    }
    return true;
}

// ---- OzzSkeletalAnimatorState -------------------------------------------------------------------------------------------
RefState::RefState(const orc_skeleton* skel, const StateDef& def, std::vector<RefAnimationState> states)
    : _skel(skel), _def(def), _animationStates(std::move(states)), _speed(def.speed),
      _locals((size_t)skel->num_soa * ORC_SOA_FLOATS, 0.f) {
    _duration = calcDuration();
}

std::optional<RefState> RefState::create(const orc_skeleton* skel, const StateDef& def, const std::vector<NamedAnimation>& anims) {
    std::vector<RefAnimationState> states;
    for (const auto& animDef : def.animations)
        for (const auto& a : anims)
            if (a.name == animDef.name) {
                states.emplace_back(skel, animDef, a.anim);
                break;
            }
    if (states.empty()) return std::nullopt;
    return RefState(skel, def, std::move(states));
}

float RefState::calcDuration() const {
    if (_animationStates.empty()) return 0.f;
    int f = 1;
    for (const auto& anim : _animationStates) {
        const float d = anim.getDuration();
        const float frac = d - std::floor(d);
        // ceil(-log10(0)) = +inf -> int is undefined in the reference; on x86 it ends with a factor of 1
        if (!(frac > 0.f)) continue;
        const int decimalPlaces = (int)std::ceil(-std::log10(frac));
        f = std::max(f, (int)std::pow(10, decimalPlaces));
    }
    std::vector<int> scaled;
    for (const auto& anim : _animationStates) scaled.push_back((int)std::round(anim.getDuration() * f));
    int d = scaled[0];
    for (size_t i = 1; i < scaled.size(); ++i) d = std::lcm(d, scaled[i]);
    return (float)d / f;
}

float RefState::calcBlendWeight(Vec2 pos, Vec2 animPos) const {
    std::vector<std::pair<Vec2, Vec2>> parts;
    if (_def.blend == 1) {
        const float lenAnim = length(animPos), len = length(pos);
        for (const auto& anim : _def.animations) {
            const Vec2 blendPos = anim.blend_position;
            const float lenAnim2 = length(blendPos);
            const float p = lenAnim + lenAnim2;
            parts.push_back({Vec2{(len - lenAnim) * p, angle(pos, animPos)}, Vec2{(lenAnim2 - lenAnim) * p, angle(blendPos, animPos)}});
        }
    } else {
        const Vec2 a{pos.x - animPos.x, pos.y - animPos.y};
        for (const auto& anim : _def.animations) parts.push_back({a, Vec2{anim.blend_position.x - animPos.x, anim.blend_position.y - animPos.y}});
    }
    float w = FLT_MAX;
    for (const auto& ab : parts) {
        const float f = 1.F - (dot2(ab.first, ab.second) / length2(ab.second));
        if (f < w) w = f;
    }
    return w;
}

std::vector<float> RefState::calcBlendWeights(Vec2 pos) const {
    std::vector<float> w;
    for (const auto& anim : _def.animations) w.push_back(calcBlendWeight(pos, anim.blend_position));
    return w;
}

const std::vector<float>& RefState::getLocals() const {
    if (_animationStates.size() == 1) return _animationStates.front().getLocals();
    return _locals;
}

bool RefState::hasFinished() const {
    for (const auto& anim : _animationStates)
        if (!anim.hasFinished()) return false;
    return true;
}

bool RefState::update(float dt, Vec2 blendPosition, std::string& err) {
    const bool ok = updateImpl(dt, blendPosition, err);
    if (ok) _everUpdated = true;   // the locals are no longer the zero-initialised buffer
    return ok;
}

bool RefState::updateImpl(float dt, Vec2 blendPosition, std::string& err) {
    dt *= _speed;
    for (auto& anim : _animationStates)
        if (!anim.update(dt, err)) return false;
    const float normDeltaTime = dt / _duration;
    _normalizedTime += normDeltaTime;
    _looped = _normalizedTime > 1.f;
    _normalizedTime = std::fmod(_normalizedTime, 1.f);
    if (_animationStates.size() <= 1) return true;

    if (!equal(_blendPos, blendPosition)) {
        const float blendFactor = tween(_def.tween, _normalizedTweenTime);
        _oldBlendPos = Vec2{_oldBlendPos.x + blendFactor * (_blendPos.x - _oldBlendPos.x), _oldBlendPos.y + blendFactor * (_blendPos.y - _oldBlendPos.y)};
        _normalizedTweenTime = 0.f;
        _blendPos = blendPosition;
    }
    _normalizedTweenTime += dt / _def.tween.duration;
    if (_normalizedTweenTime >= 1.f) _normalizedTweenTime = 1.f;
    const float blendFactor = tween(_def.tween, _normalizedTweenTime);
    const Vec2 pos{_oldBlendPos.x + blendFactor * (_blendPos.x - _oldBlendPos.x), _oldBlendPos.y + blendFactor * (_blendPos.y - _oldBlendPos.y)};

    std::vector<const float*> layers(_animationStates.size(), nullptr);
    std::vector<float> layerWeights(_animationStates.size(), 0.f);
    const float* restPose = nullptr;
    size_t i = 0;
    const std::vector<float> weights = calcBlendWeights(pos);
    for (auto& anim : _animationStates) {
        if (equal(anim.getBlendPosition(), Vec2{0.f, 0.f})) restPose = anim.getLocals().data();
        if (equal(anim.getBlendPosition(), pos)) {
            _locals = anim.getLocals();
            return true;
        }
        if (i < weights.size()) layerWeights[i] = weights[i];
        layers[i] = anim.getLocals().data();
        ++i;
    }
    if (!restPose) restPose = layers[0];
    if (!orc_blend((int)layers.size(), layers.data(), layerWeights.data(), restPose, _def.threshold, _skel->num_soa, _locals.data())) {
        err = "error in the blending job";
        return false;
    }
    return true;
}

// ---- OzzSkeletalAnimatorTransition ----------------------------------------------------------------------------------------
RefTransition::RefTransition(const orc_skeleton* skel, const TransitionDef& def, RefState current, RefState previous)
    : _skel(skel), _def(def), _currentState(std::move(current)), _previousState(std::move(previous)), _locals(_previousState.getLocals()),
      _startedFromZeroLocals(!_previousState.everUpdated()) {}

bool RefTransition::update(float dt, Vec2 blendPosition, std::string& err) {
    if (hasFinished()) {
        _locals = _currentState.getLocals();
        return true;
    }
    if (!_previousState.update(dt, blendPosition, err)) return false;
    float currentStateDeltaTime = dt;
    if (_normalizedTime == 0.f) currentStateDeltaTime += _def.offset;
    std::string swallowed;
    if (!_currentState.update(currentStateDeltaTime, blendPosition, swallowed)) {
        if (!_everBlended && _startedFromZeroLocals) ++g_unsampledHolds;   // the zero-initialised locals become the pose
        return true;
    }
    _everBlended = true;
    _normalizedTime += dt / _def.tween.duration;
    const float v = tween(_def.tween, _normalizedTime);
    const float* layers[2] = {_previousState.getLocals().data(), _currentState.getLocals().data()};
    const float weights[2] = {1.f - v, v};
    if (!orc_blend(2, layers, weights, layers[0], FLT_MIN, _skel->num_soa, _locals.data())) {
        err = "error in the blending job";
        return false;
    }
    return true;
}

// ---- SkeletalAnimatorImpl ---------------------------------------------------------------------------------------------------
RefAnimator::RefAnimator(const orc_skeleton* skel, std::vector<NamedAnimation> anims, AnimatorDef def)
    : _skel(skel), _anims(std::move(anims)), _def(std::move(def)), _models((size_t)skel->num_joints * 16, 0.f) {
    orc_local_to_model(skel, skel->rest, _models.data());   // load(): rest pose (src/skeleton_ozz.cpp:901-916)
}

const StateDef* RefAnimator::findState(const std::string& name) const {
    for (const auto& s : _def.states)
        if (s.name == name) return &s;
    return nullptr;
}
const TransitionDef* RefAnimator::findTransition(const std::string& src, const std::string& dst) const {
    for (const auto& t : _def.transitions) if (t.src_state == src && t.dst_state == dst) return &t;
    for (const auto& t : _def.transitions) if (t.src_state == src && t.dst_state.empty()) return &t;
    for (const auto& t : _def.transitions) if (t.src_state.empty() && t.dst_state == dst) return &t;
    for (const auto& t : _def.transitions) if (t.src_state.empty() && t.dst_state.empty()) return &t;
    return nullptr;
}
const RefState* RefAnimator::currentState() const {
    if (_transition) return &const_cast<RefTransition&>(*_transition).current();
    if (_state) return &*_state;
    return nullptr;
}
float RefAnimator::getStateDuration(const std::string& name) const {
    if (const auto* def = findState(name))
        if (auto st = RefState::create(_skel, *def, _anims)) return st->getDuration();
    return 0.f;
}

bool RefAnimator::play(const std::string& name, float stateSpeed) {
    RefState* curr = _transition ? &_transition->current() : (_state ? &*_state : nullptr);
    if (curr && curr->getName() == name) {
        curr->setSpeed(stateSpeed);
        return false;
    }
    const StateDef* stateDef = findState(name);
    if (!stateDef) return false;
    std::optional<RefState> prevState;
    if (_state) {
        prevState.emplace(std::move(*_state));
        _state.reset();
    }
    if (_transition) {
        prevState.emplace(_transition->finish());
        _transition.reset();
    }
    if (prevState) {
        if (const TransitionDef* transConfig = findTransition(prevState->getName(), name)) {
            auto state = RefState::create(_skel, *stateDef, _anims);
            if (!state) return false;
            _transition.emplace(_skel, *transConfig, std::move(*state), std::move(*prevState));
            events.push_back("started:" + _transition->current().getName());
            events.push_back("transition_started");
            return true;
        }
    }
    auto state = RefState::create(_skel, *stateDef, _anims);
    if (!state) return false;
    _state = std::move(state);
    events.push_back("started:" + _state->getName());
    return true;
}

bool RefAnimator::update(float dt, std::string& err) {
    dt *= _speed;
    const float* input = nullptr;
    if (_transition) {
        if (!_transition->update(dt, _blendPosition, err)) return false;
        input = _transition->getLocals().data();
    } else if (_state) {
        if (!_state->update(dt, _blendPosition, err)) return false;
        input = _state->getLocals().data();
    } else {
        return true;
    }
    if (!orc_local_to_model(_skel, input, _models.data())) {
        err = "error in the model job";
        return false;
    }
    afterUpdate();
    return true;
}

void RefAnimator::afterUpdate() {
    if (_transition && _transition->hasFinished()) {
        events.push_back("transition_finished");
        events.push_back("finished:" + _transition->previous().getName());
        if (_state) _state->afterFinished();
        _state.emplace(_transition->finish());
        _transition.reset();
    }
    if (_state) {
        if (_state->hasLooped()) events.push_back("looped:" + _state->getName());
        if (_state->hasFinished()) {
            events.push_back("finished:" + _state->getName());
            _state->afterFinished();
            const std::string nextState = _state->getNextState();
            if (!nextState.empty()) play(nextState);
        }
    }
}

}  // namespace orc
