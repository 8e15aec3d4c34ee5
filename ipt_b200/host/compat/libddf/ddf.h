// Interface mirror of the reference's src/libddf/ddf.h:12-41 for builds outside the reference tree: the abstract
// "directional distribution function". With -I<reference>/src ahead of this directory the real header is used.
// Only the abstract interface is mirrored (no pooled operator new: an allocator detail, ddf.cpp:16-56).
#ifndef DDF_H
#define DDF_H
#include <glm/vec3.hpp>
#include <memory>

struct Ddf {
    virtual glm::vec3 sample() const = 0;               // a direction; the zero vector is a FAILED sample
    virtual float value(glm::vec3 direction) const = 0; // density over solid angle; NaN if singular
    virtual ~Ddf() {}
};
#endif
