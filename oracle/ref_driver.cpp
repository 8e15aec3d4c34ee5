// oracle/ref_driver.cpp — C-ABI driver around the UNMODIFIED dimalit/ipt sources.
//
// TEST INFRASTRUCTURE ONLY. This translation unit is compiled, together with the reference's own
// .cpp files where they lie under /root/reference, into oracle/_ref/libipt_ref.so by oracle/Makefile.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs load it.
// Nothing under ipt_b200/ may include, link or dlopen it.
//
// The reference's trace loop lives in src/main.cpp next to main(); `StatsNode` is private to that
// file (src/main.cpp:26-49), so the only way to call render_sample()/ray_power_recursive() unmodified
// is to include main.cpp here with main() renamed.
//
// Besides the five reference scenes (src/sample_scenes.cpp:20-108) the driver holds ONE extension
// plug-in, `GeometryCornell` + `GlossyDdf` (BASELINE.json configs[1]; SURVEY.md §8d "C2"), written
// against the reference's own Geometry/Ddf interfaces so that the reference's own estimator
// (ray_power_recursive) remains the judge of the extension scene.

// Reference headers first: main.cpp's `using namespace glm` would make `detail::` ambiguous in them.
#include <libddf/ddf_detail.h>
#include "SimpleCamera.h"
#include "CollectionLighting.h"
#include "lighting/lighting.h"
#include "geometry/geometric_utils.h"
#include "geometry/GeometryOpenSpheres.h"
#include <glm/geometric.hpp>

#define main ipt_reference_main
#include "main.cpp"
#undef main

#include <cstdint>
#include <cstring>
#include <vector>
#include <string>

using namespace glm;
using namespace std;

// ------------------------------------------------------------------------------------------------
// Extension plug-in for config C2: power-cosine lobe + diffuse/glossy single-class mixture.
// A single (non-Union) Ddf on purpose: unite(UnionDdf, UnionDdf) in src/libddf/ddf.cpp:186-205
// moves into end() of empty vectors and is never exercised by any reference scene.
// ------------------------------------------------------------------------------------------------
namespace {

struct PowerCosineDdf : public Ddf {
    float exponent;
    explicit PowerCosineDdf(float e) : exponent(e) {}
    vec3 sample() const override {
        float u1 = randf();
        float u2 = randf();
        float cos_alpha = pow(u1, 1.0f / (exponent + 1.0f));
        float alpha = acos(cos_alpha);
        float phi = 2 * M_PI * u2;
        float r = sin(alpha);
        return vec3(r * cos(phi), r * sin(phi), cos_alpha);
    }
    float value(vec3 arg) const override {
        if (arg.z < 0.0f)
            return 0.0f;
        return (exponent + 1.0f) * pow(arg.z, exponent) / (2 * M_PI);
    }
};

// kd*Lambert(normal) + ks*PowerCosine(reflect).  Directions below the surface are FAILED samples
// (zero vector, like DdfFromLight::sample, src/lighting/lighting.cpp:55-56) and have value 0.
struct GlossyDdf : public Ddf {
    unique_ptr<Ddf> diffuse, lobe;
    vec3 normal;
    float wd, ws;
    GlossyDdf(vec3 n, vec3 refl, float kd, float ks, float exponent)
        : diffuse(make_unique<RotateDdf>(make_unique<CosineDdf>(), n)),
          lobe(make_unique<RotateDdf>(make_unique<PowerCosineDdf>(exponent), refl)),
          normal(n), wd(kd / (kd + ks)), ws(ks / (kd + ks)) {}
    vec3 sample() const override {
        float r = randf();
        vec3 w = r < wd ? diffuse->sample() : lobe->sample();
        if (dot(normal, w) < 0.0f)
            return vec3();
        return w;
    }
    float value(vec3 arg) const override {
        if (dot(normal, arg) < 0.0f)
            return 0.0f;
        return wd * diffuse->value(arg) + ws * lobe->value(arg);
    }
};

// Parameters of the C2 scene; identical literals live in ipt_b200/host/sample_scenes.cpp.
struct CornellSphere { vec3 c; float r; bool glossy; };
static const CornellSphere cornell_spheres[2] = {
    {vec3(-0.45f, 0.25f, -0.65f), 0.35f, false},
    {vec3(+0.45f, -0.2f, -0.65f), 0.35f, true},
};
static const float cornell_kd = 0.3f, cornell_ks = 0.7f, cornell_exponent = 40.0f;

struct GeometryCornell : public Geometry {
    optional<surface_intersection> traceRay(vec3 origin, vec3 direction) const override {
        vec3 planes[] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {-1, 0, 0}, {0, 0, -1}};
        float dist = numeric_limits<float>::infinity();
        int plane = -1, sphere = -1;
        for (size_t i = 0; i < 5; ++i) {
            float t = intersection_with_box_plane(planes[i], origin, direction);
            if (t < dist) { dist = t; plane = i; }
        }
        for (size_t i = 0; i < 2; ++i) {
            float t = intersection_with_sphere(cornell_spheres[i].r, origin - cornell_spheres[i].c, direction);
            if (t < dist) { dist = t; sphere = i; plane = -1; }
        }
        if (dist == numeric_limits<float>::infinity())
            return {};
        surface_intersection res;
        res.position = origin + direction * dist;
        if (plane >= 0) {
            res.curvature = 0.0f;
            res.normal = -planes[plane];
            res.sdf = make_unique<RotateDdf>(make_unique<CosineDdf>(), res.normal);
        } else {
            const CornellSphere& s = cornell_spheres[sphere];
            res.curvature = 1.0f / s.r;
            res.normal = normalize(res.position - s.c);
            if (s.glossy) {
                vec3 reflection = reflect(direction, res.normal);
                res.sdf = make_unique<GlossyDdf>(res.normal, reflection, cornell_kd, cornell_ks, cornell_exponent);
            } else {
                res.sdf = make_unique<RotateDdf>(make_unique<CosineDdf>(), res.normal);
            }
        }
        return res;
    }
};

Scene make_scene_cornell() {
    shared_ptr<CollectionLighting> lighting = make_shared<CollectionLighting>();
    lighting->addSquareLight(vec3(-0.25f, -0.25f, 0.98f), vec3(0.0f, 0.0f, -1.0f), vec3(0.0f, 0.5f, 0.0f), 4.0f);
    shared_ptr<Geometry> geometry = make_shared<GeometryCornell>();
    vec3 camera_pos(0.0f, -3.2f, 0.0f);
    vec3 camera_dir = normalize(vec3(0.0f, 1.0f, 0.0f));
    shared_ptr<SimpleCamera> camera = make_shared<SimpleCamera>(camera_pos, camera_dir);
    return Scene{geometry, lighting, camera};
}

// GeometryOpenSpheres (src/geometry/GeometryOpenSpheres.cpp) is built by no reference factory; give it one.
Scene make_scene_openspheres() {
    shared_ptr<CollectionLighting> lighting = make_shared<CollectionLighting>();
    lighting->addSquareLight(vec3(-0.05f, -0.05f, -0.2f), vec3(0, 0, -1), vec3(0, 0.1f, 0));
    shared_ptr<Geometry> geometry = make_shared<GeometryOpenSpheres>();
    vec3 camera_pos(0, -5.0f, 0);
    vec3 camera_dir = normalize(vec3(0, 0, -1.0f) - camera_pos);
    shared_ptr<SimpleCamera> camera = make_shared<SimpleCamera>(camera_pos, camera_dir * 2.0f, vec3(0, 0, 1));
    return Scene{geometry, lighting, camera};
}

// GridRenderPlane that additionally keeps double sums of v and v*v per plane pixel, so tests can
// form per-pixel z-scores (SURVEY.md §8d "Statistical parity procedure").
struct CapturePlane : public GridRenderPlane {
    vector<double> sum, sumsq;
    CapturePlane(size_t w, size_t h) : GridRenderPlane(w, h), sum(w * h), sumsq(w * h) {}
    void addRay(float x, float y, float value) override {
        size_t xi = x * width;
        size_t yi = height - y * height - 1;
        sum[yi * width + xi] += value;
        sumsq[yi * width + xi] += double(value) * value;
        GridRenderPlane::addRay(x, y, value);
    }
};

vector<Scene> g_scenes;
uint64_t g_rays_traced = 0;

// Counts Geometry::traceRay calls == "rays" in the unit of SURVEY.md §8d.
struct CountingGeometry : public Geometry {
    shared_ptr<const Geometry> inner;
    explicit CountingGeometry(shared_ptr<const Geometry> g) : inner(move(g)) {}
    optional<surface_intersection> traceRay(vec3 o, vec3 d) const override {
        ++g_rays_traced;
        return inner->traceRay(o, d);
    }
};

// Restatement of render_sample (src/main.cpp:186-223) with the hard-coded 640 replaced by W,H.
// Calls the reference's own ray_power. Equal to render_sample at W=H=640 (checked in tests).
void render_sample_wh(const Scene& scene, RenderPlane& plane, size_t W, size_t H, StatsNode* stats) {
    for (size_t iy = 0; iy < H; iy++) {
        for (size_t ix = 0; ix < W; ix++) {
            float x = (ix + randf()) / float(W);
            float y = (iy + randf()) / float(H);
            if (x == 1.0f) x = nextafter(x, 0.0f);
            if (y == 1.0f) y = nextafter(y, 0.0f);
            vec3 origin, direction;
            tie(origin, direction) = scene.camera->sampleRay(x, y);
            float value = ray_power(*scene.geometry, *scene.lighting, origin, direction, 0, n_rays, stats);
            value = value >= 0.0f ? value : 0.0f;
            if (!isfinite(value)) value = 0.0f;
            plane.addRay(x, y, value);
        }
    }
}

inline vec3 v3(const float* p) { return vec3(p[0], p[1], p[2]); }
inline void put3(float* p, vec3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }

} // namespace

extern "C" {

int iptref_scene_create(const char* name) {
    string n(name);
    Scene s;
    if (n == "box") s = make_scene_box();
    else if (n == "fractal") s = make_scene_fractal();
    else if (n == "smallpt") s = make_scene_smallpt();
    else if (n == "square") s = make_scene_square_lit_by_square();
    else if (n == "corner") s = make_scene_lit_corner();
    else if (n == "cornell") s = make_scene_cornell();
    else if (n == "openspheres") s = make_scene_openspheres();
    else return -1;
    g_scenes.push_back(s);
    return int(g_scenes.size()) - 1;
}

// Replaces the scene's lights with a rows x cols grid of square lights (config C5 shape).
int iptref_scene_set_light_grid(int scene, int rows, int cols, float side, float z, float power) {
    shared_ptr<CollectionLighting> lighting = make_shared<CollectionLighting>();
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) {
            float cx = -1.0f + (2.0f * (c + 0.5f)) / cols - 0.5f * side;
            float cy = -1.0f + (2.0f * (r + 0.5f)) / rows - 0.5f * side;
            lighting->addSquareLight(vec3(cx, cy, z), vec3(0.0f, 0.0f, -1.0f), vec3(0.0f, side, 0.0f), power);
        }
    g_scenes[scene].lighting = lighting;
    return rows * cols;
}

// Replaces the scene's lights with one of every Light class (lighting.h:16-73) through CollectionLighting's own add*
// calls; the same literals are in ipt_b200/host/sample_scenes.cpp ("mixedlights").
int iptref_scene_set_mixed_lights(int scene) {
    shared_ptr<CollectionLighting> lighting = make_shared<CollectionLighting>();
    lighting->addSquareLight(vec3{+0.1f, -0.8f - 0.1f, -0.15f}, vec3(0.0f, 0.0f, -1.0f), vec3{0.0f, 0.2f, 0.0f}, 1.0f);
    lighting->addTriangleLight(vec3(-0.8f, -0.2f, 0.6f), vec3(0.3f, 0.0f, 0.0f), vec3(0.0f, 0.0f, -0.3f), 0.5f);
    lighting->addSphereLight(vec3(-0.7f, -0.5f, -0.8f), 0.1f, 0.7f);
    lighting->addOuterLight(10.0f, 20.0f);
    lighting->addPointLight(vec3(0.9f, 0.0f, -0.8f), 0.1f, 1.0f);
    g_scenes[scene].lighting = lighting;
    return 5;
}

void iptref_set_tree(int n, int dmax) { n_rays = n; depth_max = dmax; }
void iptref_seed(long seed) { srand48(seed); }
uint64_t iptref_rays_traced() { return g_rays_traced; }

void iptref_camera_fields(int scene, float* out12) {
    const SimpleCamera* c = dynamic_cast<const SimpleCamera*>(g_scenes[scene].camera.get());
    put3(out12, c->position); put3(out12 + 3, c->direction); put3(out12 + 6, c->right); put3(out12 + 9, c->up);
}

void iptref_camera_rays(int scene, size_t n, const float* xy, float* o, float* d) {
    for (size_t i = 0; i < n; ++i) {
        auto r = g_scenes[scene].camera->sampleRay(xy[2 * i], xy[2 * i + 1]);
        put3(o + 3 * i, r.first); put3(d + 3 * i, r.second);
    }
}

void iptref_trace_geometry(int scene, size_t n, const float* o, const float* d, int32_t* hit, float* pos,
                           float* normal, float* curvature) {
    const Geometry& g = *g_scenes[scene].geometry;
    for (size_t i = 0; i < n; ++i) {
        optional<surface_intersection> si = g.traceRay(v3(o + 3 * i), v3(d + 3 * i));
        hit[i] = si.has_value();
        if (si) { put3(pos + 3 * i, si->position); put3(normal + 3 * i, si->normal); curvature[i] = si->curvature; }
        else { put3(pos + 3 * i, vec3()); put3(normal + 3 * i, vec3()); curvature[i] = 0; }
    }
}

void iptref_trace_light(int scene, size_t n, const float* o, const float* d, int32_t* hit, float* pos,
                        float* normal, float* power) {
    const Lighting& l = *g_scenes[scene].lighting;
    for (size_t i = 0; i < n; ++i) {
        optional<light_intersection> li = l.traceRayToLight(v3(o + 3 * i), v3(d + 3 * i));
        hit[i] = li.has_value();
        if (li) { put3(pos + 3 * i, li->position); put3(normal + 3 * i, li->normal); power[i] = li->surface_power; }
        else { put3(pos + 3 * i, vec3()); put3(normal + 3 * i, vec3()); power[i] = 0; }
    }
}

int iptref_light_count(int scene) {
    return int(dynamic_cast<const CollectionLighting*>(g_scenes[scene].lighting.get())->lights.size());
}
// out5 = power, area, position.xyz  (the public fields, src/lighting/lighting.h:10-12)
void iptref_light_fields(int scene, int i, float* out5) {
    const auto& l = dynamic_cast<const CollectionLighting*>(g_scenes[scene].lighting.get())->lights[i];
    out5[0] = l->power; out5[1] = l->area; put3(out5 + 2, l->position);
}

// Surface DDF returned by the geometry for ray (o,d): value(w) for n directions and n samples.
int iptref_sdf_value(int scene, const float* o, const float* d, size_t n, const float* w, float* out) {
    optional<surface_intersection> si = g_scenes[scene].geometry->traceRay(v3(o), v3(d));
    if (!si) return 0;
    for (size_t i = 0; i < n; ++i) out[i] = si->sdf->value(v3(w + 3 * i));
    return 1;
}
int iptref_sdf_sample(int scene, const float* o, const float* d, size_t n, float* w) {
    optional<surface_intersection> si = g_scenes[scene].geometry->traceRay(v3(o), v3(d));
    if (!si) return 0;
    for (size_t i = 0; i < n; ++i) put3(w + 3 * i, si->sdf->sample());
    return 1;
}
// Lighting::distributionInPoint(pos) (src/CollectionLighting.cpp:12-21): value and sample.
void iptref_light_ddf_value(int scene, const float* pos, size_t n, const float* w, float* out) {
    unique_ptr<Ddf> l = g_scenes[scene].lighting->distributionInPoint(v3(pos));
    for (size_t i = 0; i < n; ++i) out[i] = l->value(v3(w + 3 * i));
}
void iptref_light_ddf_sample(int scene, const float* pos, size_t n, float* w) {
    unique_ptr<Ddf> l = g_scenes[scene].lighting->distributionInPoint(v3(pos));
    for (size_t i = 0; i < n; ++i) put3(w + 3 * i, l->sample());
}
// The mixture the trace loop builds at a surface hit (src/main.cpp:142-143), for ray (o,d):
// n samples w_i with mix value and sdf value at each.
int iptref_mix_sample(int scene, const float* o, const float* d, size_t n, float* w, float* mixv, float* sdfv) {
    optional<surface_intersection> si = g_scenes[scene].geometry->traceRay(v3(o), v3(d));
    if (!si) return 0;
    Ddf* sdf_tmp = si->sdf.get();
    unique_ptr<Ddf> light_ddf = g_scenes[scene].lighting->distributionInPoint(si->position);
    unique_ptr<Ddf> mix = unite(move(light_ddf), 1.0f, move(si->sdf), 1.0f);
    for (size_t i = 0; i < n; ++i) {
        vec3 s = mix->sample();
        put3(w + 3 * i, s);
        mixv[i] = mix->value(s);
        sdfv[i] = sdf_tmp->value(s);
    }
    return 1;
}

// Base DDFs (src/libddf/ddf.cpp:58-108): kind 0 Spherical, 1 UpperHalf, 2 Cosine; optionally rotated
// to `to` (RotateDdf, src/libddf/ddf_detail.h:72-85) when to != NULL.
static unique_ptr<Ddf> make_base(int kind, const float* to) {
    unique_ptr<Ddf> b;
    if (kind == 0) b = make_unique<SphericalDdf>();
    else if (kind == 1) b = make_unique<UpperHalfDdf>();
    else if (kind == 2) b = make_unique<CosineDdf>();
    else b = make_unique<PowerCosineDdf>(float(kind));
    if (to) return make_unique<RotateDdf>(move(b), v3(to));
    return b;
}
void iptref_ddf_value(int kind, const float* to, size_t n, const float* w, float* out) {
    unique_ptr<Ddf> d = make_base(kind, to);
    for (size_t i = 0; i < n; ++i) out[i] = d->value(v3(w + 3 * i));
}
void iptref_ddf_sample(int kind, const float* to, size_t n, float* w) {
    unique_ptr<Ddf> d = make_base(kind, to);
    for (size_t i = 0; i < n; ++i) put3(w + 3 * i, d->sample());
}

// AreaLight known-answer test inputs (src/lighting/test_lighting.cpp:130-144).
void iptref_arealight(const float* origin, const float* xa, const float* ya, float power, int triangle,
                      const float* ro, const float* rd, float* area, int32_t* hit, float* surface_power) {
    AreaLight l(v3(origin), v3(xa), v3(ya), power, triangle ? AreaLight::TYPE_TRIANLE : AreaLight::TYPE_DIAMOND);
    *area = l.area;
    optional<light_intersection> li = l.traceRay(v3(ro), v3(rd));
    *hit = li.has_value();
    *surface_power = li ? li->surface_power : 0.0f;
}

// ray_power_preview (src/main.cpp:55-92), the reference's alternative estimator, for n rays
void iptref_preview_batch(int scene, size_t n, const float* o, const float* d, float* value) {
    StatsNode stats;
    for (size_t i = 0; i < n; ++i)
        value[i] = ray_power_preview(*g_scenes[scene].geometry, *g_scenes[scene].lighting, v3(o + 3 * i), v3(d + 3 * i), 0, 0, &stats);
}

// The reference's own estimator (ray_power_recursive, src/main.cpp:98-184) run on Geometry / Lighting objects that were
// created OUTSIDE this library (e.g. ipt_b200's host classes): the drop-in check of oracle/ab_dropin.cpp.
float iptref_ray_power_on(const void* geometry, const void* lighting, const float* o, const float* d, int depth, int n) {
    StatsNode stats;
    return ray_power(*static_cast<const Geometry*>(geometry), *static_cast<const Lighting*>(lighting), v3(o), v3(d), depth, n, &stats);
}
// ... and the reference's own objects handed out, so that the same rays can be run on both.
const void* iptref_scene_geometry(int scene) { return g_scenes[scene].geometry.get(); }
const void* iptref_scene_lighting(int scene) { return g_scenes[scene].lighting.get(); }

float iptref_ray_power(int scene, const float* o, const float* d, int depth, int n) {
    StatsNode stats;
    return ray_power(*g_scenes[scene].geometry, *g_scenes[scene].lighting, v3(o), v3(d), depth, n, &stats);
}

// `passes` calls of the reference's verbatim render_sample (640x640, src/main.cpp:186-223).
// pixels/counters: GridRenderPlane state; sum/sumsq: per plane pixel, double. Returns rays traced.
uint64_t iptref_render(int scene, int passes, float* pixels, uint64_t* counters, double* sum, double* sumsq) {
    Scene s = g_scenes[scene];
    s.geometry = make_shared<CountingGeometry>(s.geometry);
    CapturePlane plane(640, 640);
    StatsNode stats;
    uint64_t before = g_rays_traced;
    for (int p = 0; p < passes; ++p) render_sample(s, plane, &stats);
    if (pixels) memcpy(pixels, plane.pixels.data(), sizeof(float) * 640 * 640);
    if (counters) for (size_t i = 0; i < 640 * 640; ++i) counters[i] = plane.pixel_counters[i];
    if (sum) memcpy(sum, plane.sum.data(), sizeof(double) * 640 * 640);
    if (sumsq) memcpy(sumsq, plane.sumsq.data(), sizeof(double) * 640 * 640);
    return g_rays_traced - before;
}

// Same with the W,H-parametrised loop above.
uint64_t iptref_render_wh(int scene, int passes, size_t W, size_t H, float* pixels, uint64_t* counters, double* sum,
                          double* sumsq) {
    Scene s = g_scenes[scene];
    s.geometry = make_shared<CountingGeometry>(s.geometry);
    CapturePlane plane(W, H);
    StatsNode stats;
    uint64_t before = g_rays_traced;
    for (int p = 0; p < passes; ++p) render_sample_wh(s, plane, W, H, &stats);
    if (pixels) memcpy(pixels, plane.pixels.data(), sizeof(float) * W * H);
    if (counters) for (size_t i = 0; i < W * H; ++i) counters[i] = plane.pixel_counters[i];
    if (sum) memcpy(sum, plane.sum.data(), sizeof(double) * W * H);
    if (sumsq) memcpy(sumsq, plane.sumsq.data(), sizeof(double) * W * H);
    return g_rays_traced - before;
}

// GridRenderPlane::addRay alone (src/GridRenderPlane.cpp:61-75).
void iptref_plane_addray(size_t W, size_t H, size_t n, const float* x, const float* y, const float* v, float* pixels,
                         uint64_t* counters, float* max_value) {
    GridRenderPlane plane(W, H);
    for (size_t i = 0; i < n; ++i) plane.addRay(x[i], y[i], v[i]);
    memcpy(pixels, plane.pixels.data(), sizeof(float) * W * H);
    for (size_t i = 0; i < W * H; ++i) counters[i] = plane.pixel_counters[i];
    *max_value = plane.max_value;
}

} // extern "C"
