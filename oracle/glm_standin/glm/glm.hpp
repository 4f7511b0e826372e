// Minimal stand-in for the subset of GLM (g-truc/glm, un-vendored submodule of the reference:
// submodules_local/diff-gaussian-rasterization/.gitmodules:1-3, commit unpinned) that the reference
// rasterizer uses. TEST INFRASTRUCTURE ONLY: it exists so that oracle/build_ref.py can compile the
// reference's own .cu files, where they lie under /root/reference, into oracle/_ref/.
//
// Semantics follow GLM 0.9.9's generic (non-SIMD) code paths:
//   * mat3 is column-major; m[i] is column i; mat3(a0..a8) fills columns in order.
//   * (A*B)[c][r] = A[0][r]*B[c][0] + A[1][r]*B[c][1] + A[2][r]*B[c][2]   (this term order)
//   * dot(a,b) = (a.x*b.x + a.y*b.y) + a.z*b.z ; length = sqrt(dot(v,v)) ; v/s is a true division.
// "Parity unpinned" at this boundary: a stock-GLM build could differ by fp32 rounding only.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#define GLM_FUNC __host__ __device__ __forceinline__

namespace glm
{
struct vec3
{
    float x, y, z;
    GLM_FUNC vec3() {}
    GLM_FUNC vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    GLM_FUNC explicit vec3(float s) : x(s), y(s), z(s) {}
    GLM_FUNC float& operator[](int i) { return (&x)[i]; }
    GLM_FUNC const float& operator[](int i) const { return (&x)[i]; }
    GLM_FUNC vec3& operator+=(const vec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
    GLM_FUNC vec3& operator+=(float s) { x += s; y += s; z += s; return *this; }
    GLM_FUNC vec3& operator*=(float s) { x *= s; y *= s; z *= s; return *this; }
};

struct vec4
{
    float x, y, z, w;
    GLM_FUNC vec4() {}
    GLM_FUNC vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
};

GLM_FUNC vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
GLM_FUNC vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
GLM_FUNC vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
GLM_FUNC vec3 operator*(const vec3& a, const vec3& b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
GLM_FUNC vec3 operator*(float s, const vec3& v) { return vec3(s * v.x, s * v.y, s * v.z); }
GLM_FUNC vec3 operator*(const vec3& v, float s) { return vec3(v.x * s, v.y * s, v.z * s); }
GLM_FUNC vec3 operator/(const vec3& v, float s) { return vec3(v.x / s, v.y / s, v.z / s); }

GLM_FUNC float dot(const vec3& a, const vec3& b)
{
    vec3 tmp(a * b);
    return tmp.x + tmp.y + tmp.z;
}
GLM_FUNC float length(const vec3& v) { return sqrtf(dot(v, v)); }
GLM_FUNC vec3 max(const vec3& v, float s) { return vec3(fmaxf(v.x, s), fmaxf(v.y, s), fmaxf(v.z, s)); }

struct mat3
{
    vec3 c[3];
    GLM_FUNC mat3() {}
    GLM_FUNC explicit mat3(float d)
    {
        c[0] = vec3(d, 0.f, 0.f);
        c[1] = vec3(0.f, d, 0.f);
        c[2] = vec3(0.f, 0.f, d);
    }
    GLM_FUNC mat3(float x0, float y0, float z0, float x1, float y1, float z1, float x2, float y2, float z2)
    {
        c[0] = vec3(x0, y0, z0);
        c[1] = vec3(x1, y1, z1);
        c[2] = vec3(x2, y2, z2);
    }
    GLM_FUNC vec3& operator[](int i) { return c[i]; }
    GLM_FUNC const vec3& operator[](int i) const { return c[i]; }
};

GLM_FUNC mat3 operator*(const mat3& m1, const mat3& m2)
{
    mat3 r;
#pragma unroll
    for (int j = 0; j < 3; j++)
    {
        r[j][0] = m1[0][0] * m2[j][0] + m1[1][0] * m2[j][1] + m1[2][0] * m2[j][2];
        r[j][1] = m1[0][1] * m2[j][0] + m1[1][1] * m2[j][1] + m1[2][1] * m2[j][2];
        r[j][2] = m1[0][2] * m2[j][0] + m1[1][2] * m2[j][1] + m1[2][2] * m2[j][2];
    }
    return r;
}
GLM_FUNC mat3 operator*(float s, const mat3& m)
{
    mat3 r;
    r[0] = m[0] * s;
    r[1] = m[1] * s;
    r[2] = m[2] * s;
    return r;
}
GLM_FUNC mat3 transpose(const mat3& m)
{
    mat3 r;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            r[i][j] = m[j][i];
    return r;
}
} // namespace glm
