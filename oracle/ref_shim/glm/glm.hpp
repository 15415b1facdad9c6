// Minimal stand-in for the parts of GLM that /root/reference/source/{utility,triangulation,io}.hpp
// use, so those headers compile UNMODIFIED into oracle/_ref (test infrastructure only).
// Arithmetic that can affect results (determinant, inverse, mat*vec, dot, length, mix) follows
// GLM's own operation order so float32 results are the ones the reference program would see.
#pragma once
#include <math.h>     // the C++ wrappers pull std::abs(float) etc. into the global namespace,
#include <stdlib.h>   // as SDL/TinyEngine's includes do in the reference build
#include <cmath>
#include <cstring>
#include <functional>
#include <iostream>
#include <string>
#include <vector>

namespace glm {

template <typename T> struct v2 {
    T x, y;
    v2() : x(0), y(0) {}
    template <typename A> explicit v2(A a) : x(T(a)), y(T(a)) {}
    template <typename A, typename B> v2(A a, B b) : x(T(a)), y(T(b)) {}
    T& operator[](int i) { return i == 0 ? x : y; }
    const T& operator[](int i) const { return i == 0 ? x : y; }
    v2& operator+=(v2 o) { x += o.x; y += o.y; return *this; }
    v2& operator-=(v2 o) { x -= o.x; y -= o.y; return *this; }
    v2& operator*=(T s) { x *= s; y *= s; return *this; }
    v2& operator/=(T s) { x /= s; y /= s; return *this; }
};
template <typename T> struct v3 {
    T x, y, z;
    v3() : x(0), y(0), z(0) {}
    template <typename A> explicit v3(A a) : x(T(a)), y(T(a)), z(T(a)) {}
    template <typename A, typename B, typename C> v3(A a, B b, C c) : x(T(a)), y(T(b)), z(T(c)) {}
    T& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    const T& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
template <typename T> struct v4 {
    T x, y, z, w;
    v4() : x(0), y(0), z(0), w(0) {}
    template <typename A> explicit v4(A a) : x(T(a)), y(T(a)), z(T(a)), w(T(a)) {}
    template <typename A, typename B, typename C, typename D> v4(A a, B b, C c, D d) : x(T(a)), y(T(b)), z(T(c)), w(T(d)) {}
    T& operator[](int i) { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
    const T& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
    v4& operator/=(T s) { x /= s; y /= s; z /= s; w /= s; return *this; }
};
using vec2 = v2<float>; using vec3 = v3<float>; using vec4 = v4<float>;
using ivec2 = v2<int>;  using ivec3 = v3<int>;  using ivec4 = v4<int>;

inline vec2 operator+(vec2 a, vec2 b) { return vec2(a.x + b.x, a.y + b.y); }
inline vec2 operator-(vec2 a, vec2 b) { return vec2(a.x - b.x, a.y - b.y); }
inline vec2 operator*(float s, vec2 a) { return vec2(s * a.x, s * a.y); }
inline vec2 operator*(vec2 a, float s) { return vec2(a.x * s, a.y * s); }
inline vec2 operator/(vec2 a, float s) { return vec2(a.x / s, a.y / s); }
inline vec3 operator/(vec3 a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
inline float dot(vec2 a, vec2 b) { float tx = a.x * b.x, ty = a.y * b.y; return tx + ty; }
inline float length(vec2 a) { return sqrtf(dot(a, a)); }
inline vec2 mix(vec2 a, vec2 b, float t) { return a * (1.0f - t) + b * t; }

struct mat3 {                      // column major, like GLM
    vec3 c[3];
    mat3() {}
    mat3(float x0, float y0, float z0, float x1, float y1, float z1, float x2, float y2, float z2)
    { c[0] = vec3(x0, y0, z0); c[1] = vec3(x1, y1, z1); c[2] = vec3(x2, y2, z2); }
    vec3& operator[](int i) { return c[i]; }
    const vec3& operator[](int i) const { return c[i]; }
};
inline float determinant(const mat3& m)
{
    return + m[0][0] * (m[1][1] * m[2][2] - m[2][1] * m[1][2])
           - m[1][0] * (m[0][1] * m[2][2] - m[2][1] * m[0][2])
           + m[2][0] * (m[0][1] * m[1][2] - m[1][1] * m[0][2]);
}
inline mat3 inverse(const mat3& m)
{
    const float ood = 1.0f / determinant(m);
    mat3 r;
    r[0][0] = +(m[1][1] * m[2][2] - m[2][1] * m[1][2]) * ood;
    r[1][0] = -(m[1][0] * m[2][2] - m[2][0] * m[1][2]) * ood;
    r[2][0] = +(m[1][0] * m[2][1] - m[2][0] * m[1][1]) * ood;
    r[0][1] = -(m[0][1] * m[2][2] - m[2][1] * m[0][2]) * ood;
    r[1][1] = +(m[0][0] * m[2][2] - m[2][0] * m[0][2]) * ood;
    r[2][1] = -(m[0][0] * m[2][1] - m[2][0] * m[0][1]) * ood;
    r[0][2] = +(m[0][1] * m[1][2] - m[1][1] * m[0][2]) * ood;
    r[1][2] = -(m[0][0] * m[1][2] - m[1][0] * m[0][2]) * ood;
    r[2][2] = +(m[0][0] * m[1][1] - m[1][0] * m[0][1]) * ood;
    return r;
}
inline vec3 operator*(const mat3& m, vec3 v)
{
    return vec3(m[0][0] * v.x + m[1][0] * v.y + m[2][0] * v.z,
                m[0][1] * v.x + m[1][1] * v.y + m[2][1] * v.z,
                m[0][2] * v.x + m[1][2] * v.y + m[2][2] * v.z);
}
inline mat3 transpose(const mat3& m)
{
    return mat3(m[0][0], m[1][0], m[2][0], m[0][1], m[1][1], m[2][1], m[0][2], m[1][2], m[2][2]);
}
struct mat4 {
    vec4 c[4];
    mat4() {}
    mat4(std::initializer_list<float> l) { int i = 0; for (float f : l) { c[i / 4][i % 4] = f; i++; } }
    vec4& operator[](int i) { return c[i]; }
    const vec4& operator[](int i) const { return c[i]; }
};
inline mat4 transpose(const mat4& m) { mat4 r; for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) r[i][j] = m[j][i]; return r; }

}  // namespace glm
