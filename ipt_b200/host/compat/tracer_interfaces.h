// Interface mirror of the reference's src/tracer_interfaces.h:11-54 for builds outside the reference tree (the GPU box
// has no reference sources). Same type names, member names, virtual signatures and ownership: a translation unit
// that compiles against this header compiles against the real one (put -I<reference>/src -I<reference>/include first;
// `make -C ipt_b200/host check-reference` does exactly that when /root/reference is present).
#ifndef TRACER_INTERFACES_H
#define TRACER_INTERFACES_H
#include "libddf/ddf.h"
#include <glm/vec3.hpp>
#include <memory>
#include <optional>
#include <utility>

struct intersection {                                 // tracer_interfaces.h:11-14
    glm::vec3 position;
    glm::vec3 normal;
};
struct surface_intersection : public intersection {   // :16-20
    float curvature;
    std::unique_ptr<Ddf> sdf;
    float albedo = 1.0f;
};
struct light_intersection : public intersection {     // :22-24
    float surface_power;
};
struct Geometry {                                     // :26-29
    virtual std::optional<surface_intersection> traceRay(glm::vec3 origin, glm::vec3 direction) const = 0;
};
struct Lighting {                                     // :31-37
    virtual std::unique_ptr<Ddf> distributionInPoint(glm::vec3 pos) const = 0;
    virtual std::optional<light_intersection> traceRayToLight(glm::vec3 origin, glm::vec3 direction) const = 0;
};
struct Camera {                                       // :39-43
    virtual std::pair<glm::vec3, glm::vec3> sampleRay(float x, float y) const = 0;
};
struct Scene {                                        // :45-49
    std::shared_ptr<const Geometry> geometry;
    std::shared_ptr<const Lighting> lighting;
    std::shared_ptr<const Camera> camera;
};
struct RenderPlane {                                  // :51-54
    virtual void addRay(float x, float y, float value) = 0;
};
#endif
