// gk_glm.h — the handful of glm operations the reference's host code performs on the way to
// the renderer boundary, restated for the host mirror (glm itself is an un-vendored vcpkg
// dependency of the reference; versions unpinned, see vcpkg_linux.sh).
//
// Conventions are the reference's (src/Utilities/Glm.hpp:3-5): right-handed, depth 0..1,
// radians, column-major mat4 with m[c][r].
//   lookAt      -> glm::lookAtRH            (glm/ext/matrix_transform.inl)
//   perspective -> glm::perspectiveRH_ZO    (glm/ext/matrix_clip_space.inl)
//   inverse     -> glm::inverse(mat4)       (glm/detail/func_matrix.inl, cofactor form)
//   quat(euler), mat4_cast, translate, scale, mat4*mat4
#pragma once
#include <cmath>
#include <cstring>

namespace gk {

struct vec2 {
    float x = 0, y = 0;
    vec2() = default;
    template <class A, class B> vec2(A a, B b) : x((float)a), y((float)b) {}
};
struct vec3 {
    float x = 0, y = 0, z = 0;
    vec3() = default;
    vec3(float a) : x(a), y(a), z(a) {}
    template <class A, class B, class C> vec3(A a, B b, C c) : x((float)a), y((float)b), z((float)c) {}
};
struct vec4 {
    float x = 0, y = 0, z = 0, w = 0;
    vec4() = default;
    vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    vec4(vec3 v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
};
struct quat {
    float w = 1, x = 0, y = 0, z = 0;
    quat() = default;
    quat(float w_, float x_, float y_, float z_) : w(w_), x(x_), y(y_), z(z_) {}
    explicit quat(vec3 e) // glm::quat(vec3 eulerAngles)
    {
        vec3 c(cosf(e.x * 0.5f), cosf(e.y * 0.5f), cosf(e.z * 0.5f));
        vec3 s(sinf(e.x * 0.5f), sinf(e.y * 0.5f), sinf(e.z * 0.5f));
        w = c.x * c.y * c.z + s.x * s.y * s.z;
        x = s.x * c.y * c.z - c.x * s.y * s.z;
        y = c.x * s.y * c.z + s.x * c.y * s.z;
        z = c.x * c.y * s.z - s.x * s.y * c.z;
    }
};

inline vec3 operator+(vec3 a, vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline vec3 operator-(vec3 a, vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline vec3 operator*(vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline vec3 operator*(vec3 a, vec3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline float dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline vec3 cross(vec3 a, vec3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
inline float length(vec3 a) { return sqrtf(dot(a, a)); }
inline vec3 normalize(vec3 a) { float i = 1.0f / sqrtf(dot(a, a)); return a * i; }
inline vec3 vmin(vec3 a, vec3 b) { return {fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)}; }
inline vec3 vmax(vec3 a, vec3 b) { return {fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)}; }

struct mat4 {
    float m[4][4]; // m[column][row]
    mat4() { identity(); }
    void identity()
    {
        memset(m, 0, sizeof(m));
        m[0][0] = m[1][1] = m[2][2] = m[3][3] = 1.f;
    }
    const float* data() const { return &m[0][0]; }
    float* data() { return &m[0][0]; }
    static mat4 from(const float* p) { mat4 r; memcpy(r.m, p, 64); return r; }
    bool operator!=(const mat4& o) const { return memcmp(m, o.m, 64) != 0; }
};

inline mat4 operator*(const mat4& a, const mat4& b)
{
    mat4 r;
    for (int c = 0; c < 4; ++c)
        for (int row = 0; row < 4; ++row)
            r.m[c][row] = a.m[0][row] * b.m[c][0] + a.m[1][row] * b.m[c][1] + a.m[2][row] * b.m[c][2] + a.m[3][row] * b.m[c][3];
    return r;
}
inline vec4 operator*(const mat4& a, vec4 v)
{
    // glm: Mov0*m[0] + Mov1*m[1] ... grouped as (m0*x + m1*y) + (m2*z + m3*w)
    vec4 r;
    r.x = (a.m[0][0] * v.x + a.m[1][0] * v.y) + (a.m[2][0] * v.z + a.m[3][0] * v.w);
    r.y = (a.m[0][1] * v.x + a.m[1][1] * v.y) + (a.m[2][1] * v.z + a.m[3][1] * v.w);
    r.z = (a.m[0][2] * v.x + a.m[1][2] * v.y) + (a.m[2][2] * v.z + a.m[3][2] * v.w);
    r.w = (a.m[0][3] * v.x + a.m[1][3] * v.y) + (a.m[2][3] * v.z + a.m[3][3] * v.w);
    return r;
}

inline mat4 transpose(const mat4& a)
{
    mat4 r;
    for (int c = 0; c < 4; ++c)
        for (int row = 0; row < 4; ++row) r.m[c][row] = a.m[row][c];
    return r;
}

inline mat4 translate(vec3 v) { mat4 r; r.m[3][0] = v.x; r.m[3][1] = v.y; r.m[3][2] = v.z; return r; }
inline mat4 scale(vec3 v) { mat4 r; r.m[0][0] = v.x; r.m[1][1] = v.y; r.m[2][2] = v.z; return r; }
inline mat4 mat4_cast(quat q)
{
    mat4 R;
    float qxx = q.x * q.x, qyy = q.y * q.y, qzz = q.z * q.z, qxz = q.x * q.z, qxy = q.x * q.y, qyz = q.y * q.z, qwx = q.w * q.x, qwy = q.w * q.y, qwz = q.w * q.z;
    R.m[0][0] = 1.f - 2.f * (qyy + qzz);
    R.m[0][1] = 2.f * (qxy + qwz);
    R.m[0][2] = 2.f * (qxz - qwy);
    R.m[1][0] = 2.f * (qxy - qwz);
    R.m[1][1] = 1.f - 2.f * (qxx + qzz);
    R.m[1][2] = 2.f * (qyz + qwx);
    R.m[2][0] = 2.f * (qxz + qwy);
    R.m[2][1] = 2.f * (qyz - qwx);
    R.m[2][2] = 1.f - 2.f * (qxx + qyy);
    return R;
}

inline mat4 lookAt(vec3 eye, vec3 center, vec3 up)
{
    const vec3 f = normalize(center - eye);
    const vec3 s = normalize(cross(f, up));
    const vec3 u = cross(s, f);
    mat4 R;
    R.m[0][0] = s.x; R.m[1][0] = s.y; R.m[2][0] = s.z;
    R.m[0][1] = u.x; R.m[1][1] = u.y; R.m[2][1] = u.z;
    R.m[0][2] = -f.x; R.m[1][2] = -f.y; R.m[2][2] = -f.z;
    R.m[3][0] = -dot(s, eye); R.m[3][1] = -dot(u, eye); R.m[3][2] = dot(f, eye);
    return R;
}

inline mat4 perspective(float fovy, float aspect, float zNear, float zFar)
{
    const float t = tanf(fovy / 2.f);
    mat4 R;
    memset(R.m, 0, sizeof(R.m));
    R.m[0][0] = 1.f / (aspect * t);
    R.m[1][1] = 1.f / t;
    R.m[2][2] = zFar / (zNear - zFar);
    R.m[2][3] = -1.f;
    R.m[3][2] = -(zFar * zNear) / (zFar - zNear);
    return R;
}

inline mat4 inverse(const mat4& M)
{
    const float(*m)[4] = M.m;
    float c00 = m[2][2] * m[3][3] - m[3][2] * m[2][3], c02 = m[1][2] * m[3][3] - m[3][2] * m[1][3], c03 = m[1][2] * m[2][3] - m[2][2] * m[1][3];
    float c04 = m[2][1] * m[3][3] - m[3][1] * m[2][3], c06 = m[1][1] * m[3][3] - m[3][1] * m[1][3], c07 = m[1][1] * m[2][3] - m[2][1] * m[1][3];
    float c08 = m[2][1] * m[3][2] - m[3][1] * m[2][2], c10 = m[1][1] * m[3][2] - m[3][1] * m[1][2], c11 = m[1][1] * m[2][2] - m[2][1] * m[1][2];
    float c12 = m[2][0] * m[3][3] - m[3][0] * m[2][3], c14 = m[1][0] * m[3][3] - m[3][0] * m[1][3], c15 = m[1][0] * m[2][3] - m[2][0] * m[1][3];
    float c16 = m[2][0] * m[3][2] - m[3][0] * m[2][2], c18 = m[1][0] * m[3][2] - m[3][0] * m[1][2], c19 = m[1][0] * m[2][2] - m[2][0] * m[1][2];
    float c20 = m[2][0] * m[3][1] - m[3][0] * m[2][1], c22 = m[1][0] * m[3][1] - m[3][0] * m[1][1], c23 = m[1][0] * m[2][1] - m[2][0] * m[1][1];
    const float F0[4] = {c00, c00, c02, c03}, F1[4] = {c04, c04, c06, c07}, F2[4] = {c08, c08, c10, c11};
    const float F3[4] = {c12, c12, c14, c15}, F4[4] = {c16, c16, c18, c19}, F5[4] = {c20, c20, c22, c23};
    const float V0[4] = {m[1][0], m[0][0], m[0][0], m[0][0]}, V1[4] = {m[1][1], m[0][1], m[0][1], m[0][1]};
    const float V2[4] = {m[1][2], m[0][2], m[0][2], m[0][2]}, V3[4] = {m[1][3], m[0][3], m[0][3], m[0][3]};
    const float sA[4] = {+1, -1, +1, -1}, sB[4] = {-1, +1, -1, +1};
    mat4 I;
    for (int i = 0; i < 4; ++i) {
        I.m[0][i] = (V1[i] * F0[i] - V2[i] * F1[i] + V3[i] * F2[i]) * sA[i];
        I.m[1][i] = (V0[i] * F0[i] - V2[i] * F3[i] + V3[i] * F4[i]) * sB[i];
        I.m[2][i] = (V0[i] * F1[i] - V1[i] * F3[i] + V3[i] * F5[i]) * sA[i];
        I.m[3][i] = (V0[i] * F2[i] - V1[i] * F4[i] + V2[i] * F5[i]) * sB[i];
    }
    const float d0 = m[0][0] * I.m[0][0], d1 = m[0][1] * I.m[1][0], d2 = m[0][2] * I.m[2][0], d3 = m[0][3] * I.m[3][0];
    const float inv = 1.f / ((d0 + d1) + (d2 + d3));
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r) I.m[c][r] *= inv;
    return I;
}

inline float radians(float deg) { return deg * 0.01745329251994329576923690768489f; }

} // namespace gk
