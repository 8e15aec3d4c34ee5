// ipt_b200/host/compat/libddf/ddf.h
//
// Stand-in for the reference's src/libddf/ddf.h where the reference sources do not exist (see the note at the top of
// ../tracer_interfaces.h): the abstract "directional distribution function" and nothing else. The reference's header
// additionally routes `operator new/delete` of Ddf objects through a per-thread boost::pool (an allocator detail:
// src/libddf/ddf.cpp:16-56) and declares unite(); neither is part of what a Geometry / Lighting plug-in must
// implement, and the B200 path never allocates DDF objects on its hot path.
#ifndef DDF_H
#define DDF_H

#include <memory>

#include <glm/vec3.hpp>

struct Ddf {
    // One direction drawn from the distribution. The ZERO vector means "this sample failed" (e.g. the emitter faces
    // away, src/lighting/lighting.cpp:55-56); the trace loop skips it but still counts it in its 1/n divisor.
    virtual glm::vec3 sample() const = 0;

    // Density over solid angle in the given direction (NaN for singular distributions such as a mirror).
    virtual float value(glm::vec3 direction) const = 0;

    virtual ~Ddf() {}
};

#endif // DDF_H
