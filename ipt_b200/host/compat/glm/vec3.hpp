// Minimal stand-in for <glm/vec3.hpp> used ONLY when ipt_b200's host classes are built outside the reference tree
// (the reference vendors glm 0.9.9.7 under include/glm; with -I<reference>/include ahead of this directory the real
// header is found instead). Just enough of glm::vec3 for the interface types of tracer_interfaces.h.
#pragma once
namespace glm {
struct vec3 {
    float x, y, z;
    vec3() : x(0), y(0), z(0) {}
    vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    float& operator[](int i) { return (&x)[i]; }
    const float& operator[](int i) const { return (&x)[i]; }
};
inline bool operator==(const vec3& a, const vec3& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
} // namespace glm
