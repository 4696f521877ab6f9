// glm_lite.hpp -- the handful of glm types the skeleton API signatures use.  Inside darmok the real glm is used
// (the API takes glm::vec2 / glm::vec3 / glm::mat4, include/darmok/skeleton.hpp:259-301); this stand-in only exists so
// that the shim builds and is testable in an image without glm.  Column-major, same memory layout as glm.
#pragma once

#if __has_include(<glm/glm.hpp>)
#include <glm/glm.hpp>
#include <glm/gtx/norm.hpp>
#include <glm/gtx/vector_angle.hpp>
#else
#include <cmath>

namespace glm {

struct vec2 {
    float x = 0.f, y = 0.f;
    constexpr vec2() = default;
    constexpr explicit vec2(float v) : x(v), y(v) {}
    constexpr vec2(float x_, float y_) : x(x_), y(y_) {}
    friend constexpr bool operator==(const vec2& a, const vec2& b) { return a.x == b.x && a.y == b.y; }
    friend constexpr bool operator!=(const vec2& a, const vec2& b) { return !(a == b); }
    friend constexpr vec2 operator+(const vec2& a, const vec2& b) { return {a.x + b.x, a.y + b.y}; }
    friend constexpr vec2 operator-(const vec2& a, const vec2& b) { return {a.x - b.x, a.y - b.y}; }
    friend constexpr vec2 operator*(float s, const vec2& a) { return {s * a.x, s * a.y}; }
    friend constexpr vec2 operator*(const vec2& a, float s) { return {a.x * s, a.y * s}; }
};

struct vec3 {
    float x = 0.f, y = 0.f, z = 0.f;
    constexpr vec3() = default;
    constexpr explicit vec3(float v) : x(v), y(v), z(v) {}
    constexpr vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    float& operator[](int i) { return (&x)[i]; }
    const float& operator[](int i) const { return (&x)[i]; }
};

struct vec4 {
    float x = 0.f, y = 0.f, z = 0.f, w = 0.f;
    float& operator[](int i) { return (&x)[i]; }
    const float& operator[](int i) const { return (&x)[i]; }
};

struct mat4 {
    vec4 cols[4];
    constexpr mat4() = default;
    explicit mat4(float d) {
        cols[0].x = d; cols[1].y = d; cols[2].z = d; cols[3].w = d;
    }
    vec4& operator[](int c) { return cols[c]; }
    const vec4& operator[](int c) const { return cols[c]; }
    friend bool operator==(const mat4& a, const mat4& b) {
        for (int c = 0; c < 4; ++c)
            for (int r = 0; r < 4; ++r)
                if (a[c][r] != b[c][r]) return false;
        return true;
    }
};

// glm::dot(vec2): tmp = a * b; tmp.x + tmp.y
inline float dot(const vec2& a, const vec2& b) { return a.x * b.x + a.y * b.y; }
inline float length(const vec2& v) { return std::sqrt(dot(v, v)); }
inline float length2(const vec2& v) { return dot(v, v); }
// glm::angle (gtx/vector_angle): acos(clamp(dot(x, y), -1, 1)) -- darmok calls it on un-normalised vectors
inline float angle(const vec2& x, const vec2& y) {
    float d = dot(x, y);
    d = d < -1.f ? -1.f : (d > 1.f ? 1.f : d);   // glm::clamp = min(max(x, lo), hi)
    return std::acos(d);
}
inline const float* value_ptr(const mat4& m) { return &m.cols[0].x; }
inline float* value_ptr(mat4& m) { return &m.cols[0].x; }

}  // namespace glm
#endif
