// glm_order.hpp — host-side float3 arithmetic in the exact operation order of glm 0.9.9.7's scalar path,
// which is what defines the reference's floats:
//   dot(a,b)      = (a.x*b.x + a.y*b.y) + a.z*b.z          include/glm/detail/func_geometric.inl:48-55
//   cross         = (x.y*y.z - y.y*x.z, ...)                :68-79
//   normalize(v)  = v * (1 / sqrt(dot(v,v)))                :82-90
//   inverse(mat3) = cofactors * (1/det)                     include/glm/detail/func_matrix.inl:269-291
// Used when a scene description is flattened for the device (areas, normals, inverse matrices, camera axes),
// so that the device consumes bit-identical derived constants to the ones the reference's constructors compute.
// Compile with -ffp-contract=off.
#pragma once
#include <cmath>

namespace ipt_host {

struct f3 {
    float x, y, z;
};
inline f3 mk(float x, float y, float z) { return f3{x, y, z}; }
inline f3 mk(const float* p) { return f3{p[0], p[1], p[2]}; }
inline void put(float* p, f3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }
inline f3 operator+(f3 a, f3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
inline f3 operator-(f3 a, f3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
inline f3 operator-(f3 a) { return mk(-a.x, -a.y, -a.z); }
inline f3 operator*(f3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
inline f3 operator*(float s, f3 a) { return mk(s * a.x, s * a.y, s * a.z); }
inline float dot(f3 a, f3 b) {
    float tx = a.x * b.x, ty = a.y * b.y, tz = a.z * b.z;
    return tx + ty + tz;
}
inline f3 cross(f3 x, f3 y) { return mk(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y); }
inline float length(f3 a) { return std::sqrt(dot(a, a)); }
inline f3 normalize(f3 a) { return a * (1.0f / std::sqrt(dot(a, a))); }

struct f33 {
    f3 c[3]; // columns
};
inline f33 inverse(const f33& m) {
    auto M = [&m](int i, int j) { return (&m.c[i].x)[j]; };
    float ood = 1.0f / (+M(0, 0) * (M(1, 1) * M(2, 2) - M(2, 1) * M(1, 2)) - M(1, 0) * (M(0, 1) * M(2, 2) - M(2, 1) * M(0, 2)) +
                        M(2, 0) * (M(0, 1) * M(1, 2) - M(1, 1) * M(0, 2)));
    f33 r;
    r.c[0].x = +(M(1, 1) * M(2, 2) - M(2, 1) * M(1, 2)) * ood;
    r.c[1].x = -(M(1, 0) * M(2, 2) - M(2, 0) * M(1, 2)) * ood;
    r.c[2].x = +(M(1, 0) * M(2, 1) - M(2, 0) * M(1, 1)) * ood;
    r.c[0].y = -(M(0, 1) * M(2, 2) - M(2, 1) * M(0, 2)) * ood;
    r.c[1].y = +(M(0, 0) * M(2, 2) - M(2, 0) * M(0, 2)) * ood;
    r.c[2].y = -(M(0, 0) * M(2, 1) - M(2, 0) * M(0, 1)) * ood;
    r.c[0].z = +(M(0, 1) * M(1, 2) - M(1, 1) * M(0, 2)) * ood;
    r.c[1].z = -(M(0, 0) * M(1, 2) - M(1, 0) * M(0, 2)) * ood;
    r.c[2].z = +(M(0, 0) * M(1, 1) - M(1, 0) * M(0, 1)) * ood;
    return r;
}

inline f3 operator*(const f33& m, f3 v) { // include/glm/detail/type_mat3x3.inl:468-474
    return mk(m.c[0].x * v.x + m.c[1].x * v.y + m.c[2].x * v.z, m.c[0].y * v.x + m.c[1].y * v.y + m.c[2].y * v.z,
              m.c[0].z * v.x + m.c[1].z * v.y + m.c[2].z * v.z);
}
// mat3(glm::rotate(identity<mat4>, angle, axis))            include/glm/ext/matrix_transform.inl:18-46
inline f33 rotation(float angle, f3 v) {
    float c = std::cos(angle), s = std::sin(angle);
    f3 a = normalize(v), t = a * (1.0f - c);
    float R[3][3];
    R[0][0] = c + t.x * a.x; R[0][1] = t.x * a.y + s * a.z; R[0][2] = t.x * a.z - s * a.y;
    R[1][0] = t.y * a.x - s * a.z; R[1][1] = c + t.y * a.y; R[1][2] = t.y * a.z + s * a.x;
    R[2][0] = t.z * a.x + s * a.y; R[2][1] = t.z * a.y - s * a.x; R[2][2] = c + t.z * a.z;
    f33 m; // Result[i] = I[0]*R[i][0] + I[1]*R[i][1] + I[2]*R[i][2]
    for (int i = 0; i < 3; ++i)
        m.c[i] = mk(1.0f * R[i][0] + 0.0f * R[i][1] + 0.0f * R[i][2], 0.0f * R[i][0] + 1.0f * R[i][1] + 0.0f * R[i][2],
                    0.0f * R[i][0] + 0.0f * R[i][1] + 1.0f * R[i][2]);
    return m;
}

} // namespace ipt_host
