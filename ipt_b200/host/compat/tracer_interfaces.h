// ipt_b200/host/compat/tracer_interfaces.h
//
// Stand-in for the reference's src/tracer_interfaces.h, used ONLY when the host classes of ipt_b200 are built where
// the reference sources do not exist (the GPU box). It declares the same six types with the same member names, member
// order and virtual signatures, so object layout and vtables agree with the real header and a translation unit that
// compiles against one compiles against the other. Inside the reference tree, put -I<reference>/src and
// -I<reference>/include ahead of this directory and the real header is found instead
// (`make -C ipt_b200/host check-reference` does that and compiles device_plugins.cpp against it).
//
// How the B200 path uses each type (reference lines in brackets):
//   intersection / surface_intersection / light_intersection  [11-24]
//       what the single-ray virtuals of DeviceGeometry / DeviceLighting hand back; the wavefront kernels keep the
//       same information in their 32-byte hit records instead.
//   Geometry, Lighting, Camera                                 [26-43]
//       implemented by DeviceGeometry, DeviceLighting, DeviceCamera (device_plugins.hpp), which additionally
//       describe themselves to the device (DeviceExportable).
//   Scene                                                      [45-49]
//       consumed as is by ipt_b200::render_sample.
//   RenderPlane                                                [51-54]
//       DevicePlane, or any foreign plane (fed one addRay per pixel).
#ifndef TRACER_INTERFACES_H
#define TRACER_INTERFACES_H

#include <memory>
#include <optional>
#include <utility>

#include <glm/vec3.hpp>

#include "libddf/ddf.h"

// A point on something a ray reached, with the unit normal there.
struct intersection {
    glm::vec3 position;
    glm::vec3 normal;
};

// ... on a surface: the reference adds the local curvature (reported, never used by the estimator), the surface's
// directional distribution function ("sdf": pdf of outgoing directions, owned by the caller) and an albedo.
struct surface_intersection : public intersection {
    float curvature;
    std::unique_ptr<Ddf> sdf;
    float albedo = 1.0f;
};

// ... on an emitter: emitted power per unit area (power / area; NaN for point lights).
struct light_intersection : public intersection {
    float surface_power;
};

// Closest surface along a ray, or nothing.
struct Geometry {
    virtual std::optional<surface_intersection> traceRay(glm::vec3 origin, glm::vec3 direction) const = 0;
};

// All emitters: their joint DDF as seen from a point, and the nearest emitter along a ray.
struct Lighting {
    virtual std::unique_ptr<Ddf> distributionInPoint(glm::vec3 pos) const = 0;
    virtual std::optional<light_intersection> traceRayToLight(glm::vec3 origin, glm::vec3 direction) const = 0;
    // (the reference also declares `static light_intersection last_sample`, a scratch static written by
    //  DdfFromLight::sample; statics do not take part in object layout and the B200 path has no use for it)
};

// Ray through frame position (x, y), both in [0,1): (origin, unit direction).
struct Camera {
    virtual std::pair<glm::vec3, glm::vec3> sampleRay(float x, float y) const = 0;
};

// What render_sample consumes.
struct Scene {
    std::shared_ptr<const Geometry> geometry;
    std::shared_ptr<const Lighting> lighting;
    std::shared_ptr<const Camera> camera;
};

// Receives one radiance sample per call, at frame position (x, y) in [0,1).
struct RenderPlane {
    virtual void addRay(float x, float y, float value) = 0;
};

#endif // TRACER_INTERFACES_H
