// easing.hpp -- tween curves used by the animator (state blend-position tween, transition weight).
//
// darmok evaluates `Easing::apply(type, position, 0.F, 1.F)` (src/skeleton.cpp:269-272, include/darmok/easing.hpp);
// the curves are Robert Penner's equations.  Here they are specialised to start = 0, end = 1, keeping the order of the
// floating point operations (with start = 0 and end - start = 1 the dropped terms are exact identities), so the tween
// values -- and through them the layer weights handed to the device -- are the ones darmok computes.
#pragma once

#include <cmath>

#ifdef __CUDACC__
#define DMK_HD __host__ __device__
#else
#define DMK_HD
#endif

namespace darmok {

enum class EasingType : int {  // assets/protobuf/easing.proto:6-38
    Linear = 0, Stepped, QuadraticIn, QuadraticOut, QuadraticInOut, CubicIn, CubicOut, CubicInOut, QuarticIn, QuarticOut,
    QuarticInOut, QuinticIn, QuinticOut, QuinticInOut, SinusoidalIn, SinusoidalOut, SinusoidalInOut, ExponentialIn,
    ExponentialOut, ExponentialInOut, CircularIn, CircularOut, CircularInOut, BounceIn, BounceOut, BounceInOut, ElasticIn,
    ElasticOut, ElasticInOut, BackIn, BackOut, BackInOut
};

namespace easing_detail {
constexpr float kHalfPi = 1.57079632679489661923f, kPi = 3.14159265358979323846f, kTwoPi = 6.28318530717958647692f;  // constexpr: usable in device code

DMK_HD inline float bounce_out(float p, float c) {  // c = end - start, start = 0
    if (p < 1 / 2.75f) return c * (7.5625f * p * p);
    if (p < 2.0f / 2.75f) { p -= 1.5f / 2.75f; return c * (7.5625f * p * p + .75f); }
    if (p < 2.5f / 2.75f) { p -= 2.25f / 2.75f; return c * (7.5625f * p * p + .9375f); }
    p -= 2.625f / 2.75f;
    return c * (7.5625f * p * p + .984375f);
}
DMK_HD inline float bounce_in(float p, float c) { return c - bounce_out(1 - p, c); }
}  // namespace easing_detail

// Easing::apply(type, p, 0.F, 1.F)
DMK_HD inline float ease01(EasingType type, float p) {
    using namespace easing_detail;
    switch (type) {
    case EasingType::Stepped: return 0.f;
    case EasingType::QuadraticIn: return p * p;
    case EasingType::QuadraticOut: return -1.f * p * (p - 2);
    case EasingType::QuadraticInOut:
        p *= 2;
        if (p < 1) return 0.5f * p * p;
        --p;
        return -0.5f * (p * (p - 2) - 1);
    case EasingType::CubicIn: return p * p * p;
    case EasingType::CubicOut: --p; return p * p * p + 1;
    case EasingType::CubicInOut:
        p *= 2;
        if (p < 1) return 0.5f * p * p * p;
        p -= 2;
        return 0.5f * (p * p * p + 2);
    case EasingType::QuarticIn: return p * p * p * p;
    case EasingType::QuarticOut: --p; return -1.f * (p * p * p * p - 1);
    case EasingType::QuarticInOut:
        p *= 2;
        if (p < 1) return 0.5f * (p * p * p * p);
        p -= 2;
        return -0.5f * (p * p * p * p - 2);
    case EasingType::QuinticIn: return p * p * p * p * p;
    case EasingType::QuinticOut: p--; return p * p * p * p * p + 1;
    case EasingType::QuinticInOut:
        p *= 2;
        if (p < 1) return 0.5f * (p * p * p * p * p);
        p -= 2;
        return 0.5f * (p * p * p * p * p + 2);
    case EasingType::SinusoidalIn: return -1.f * std::cos(p * kHalfPi) + 1.f;
    case EasingType::SinusoidalOut: return std::sin(p * kHalfPi);
    case EasingType::SinusoidalInOut: return -0.5f * (std::cos(p * kPi) - 1);
    case EasingType::ExponentialIn: return std::pow(2.f, 10 * (p - 1));
    case EasingType::ExponentialOut: return -std::pow(2.f, -10 * p) + 1;
    case EasingType::ExponentialInOut:
        p *= 2;
        if (p < 1) return 0.5f * std::pow(2.f, 10 * (p - 1));
        --p;
        return 0.5f * (-std::pow(2.f, -10 * p) + 2);
    case EasingType::CircularIn: return -1.f * (std::sqrt(1 - p * p) - 1);
    case EasingType::CircularOut: --p; return std::sqrt(1 - p * p);
    case EasingType::CircularInOut:
        p *= 2;
        if (p < 1) return -0.5f * (std::sqrt(1 - p * p) - 1);
        p -= 2;
        return 0.5f * (std::sqrt(1 - p * p) + 1);
    case EasingType::BounceIn: return bounce_in(p, 1.f);
    case EasingType::BounceOut: return bounce_out(p, 1.f);
    case EasingType::BounceInOut:
        if (p < 0.5f) return bounce_in(p * 2, 1.f) * .5f;
        return bounce_out(p * 2 - 1, 1.f) * .5f + 1.f * .5f;
    case EasingType::ElasticIn: {
        if (p <= 0.00001f) return 0.f;
        if (p >= 0.999f) return 1.f;
        const float period = .3f, s = period / 4;
        p -= 1;
        const float post = 1.f * std::pow(2.f, 10 * p);
        return -(post * std::sin((p - s) * kTwoPi / period));
    }
    case EasingType::ElasticOut: {
        if (p <= 0.00001f) return 0.f;
        if (p >= 0.999f) return 1.f;
        const float period = .3f, s = period / 4;
        return 1.f * std::pow(2.f, -10 * p) * std::sin((p - s) * kTwoPi / period) + 1.f;
    }
    case EasingType::ElasticInOut: {
        if (p <= 0.00001f) return 0.f;
        if (p >= 0.999f) return 1.f;
        p *= 2;
        const float period = .3f * 1.5f, s = period / 4;
        if (p < 1) {
            p -= 1;
            const float post = 1.f * std::pow(2.f, 10 * p);
            return -0.5f * (post * std::sin((p - s) * kTwoPi / period));
        }
        p -= 1;
        const float post = 1.f * std::pow(2.f, -10 * p);
        return post * std::sin((p - s) * kTwoPi / period) * .5f + 1.f;
    }
    case EasingType::BackIn: { const float s = 1.70158f; return p * p * ((s + 1) * p - s); }
    case EasingType::BackOut: { const float s = 1.70158f; p -= 1; return p * p * ((s + 1) * p + s) + 1; }
    case EasingType::BackInOut: {
        const float s = 1.70158f * 1.525f;
        float t = p / (1.f / 2);
        if (t < 1) return 1.f / 2 * (t * t * ((s + 1) * t - s));
        t -= 2;
        return 1.f / 2 * (t * t * ((s + 1) * t + s) + 2);
    }
    case EasingType::Linear:
    default: return p;
    }
}

}  // namespace darmok
