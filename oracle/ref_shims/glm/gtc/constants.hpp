// oracle/ref_shims/glm/gtc/constants.hpp -- TEST INFRASTRUCTURE (see oracle/dmk_oracle.h).
//
// glm is not in the container; the reference's easing.hpp needs three constants from it.  Values are glm's own literals
// (glm/gtc/constants.inl) converted to the requested type.  libstdc++ 13 does not declare std::cosf / std::sinf /
// std::powf (the reference is built with MSVC / newer standard libraries), so they are made visible here.
#pragma once

#include <cmath>

namespace std {
using ::cosf;
using ::powf;
using ::sinf;
}  // namespace std

namespace glm {
template <typename T> constexpr T pi() { return static_cast<T>(3.14159265358979323846264338327950288); }
template <typename T> constexpr T half_pi() { return static_cast<T>(1.57079632679489661923132169163975144); }
template <typename T> constexpr T two_pi() { return static_cast<T>(6.28318530717958647692528676655900576); }
}  // namespace glm
