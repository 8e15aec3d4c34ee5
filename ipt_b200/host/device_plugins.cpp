// device_plugins.cpp — see device_plugins.hpp. Marshals the reference-facing C++ objects to the C ABI.
#include "device_plugins.hpp"

#include "glm_order.hpp"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <thread>
#include <tuple>

// Built against the reference's own sources (its src/ directory on the include path), the reference's concrete classes
// that expose what the device needs are accepted as Scene members and as the render plane (see render_sample).
#if defined(__has_include)
#if __has_include(<SimpleCamera.h>) && __has_include(<GridRenderPlane.h>) && __has_include(<geometry/GeometrySphereInBox.h>)
#define IPT_B200_REFERENCE_CLASSES 1
#include <GridRenderPlane.h>
#include <SimpleCamera.h>
#include <geometry/FractalSpheres.h>
#include <geometry/GeometryCorner.h>
#include <geometry/GeometryFloor.h>
#include <geometry/GeometryOpenSpheres.h>
#include <geometry/GeometrySmallPt.h>
#include <geometry/GeometrySphereInBox.h>
#endif
#endif

namespace ipt_b200 {

namespace H = ipt_host;

void check(int status) {
    if (status != IPT_OK) throw Error(status, std::string("ipt_b200: ") + ipt_last_error());
}

static H::f3 h3(glm::vec3 v) { return H::mk(v.x, v.y, v.z); }
static glm::vec3 g3(H::f3 v) { return glm::vec3(v.x, v.y, v.z); }
static void put(float* p, glm::vec3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }
static ipt_material lambert() { return ipt_material{IPT_DDF_COSINE, 1.0f, 1.0f, 0.0f, 0.0f}; }
static std::atomic<uint64_t> g_sample_counter{1};

ipt_scene_desc SceneBuilder::desc() const {
    ipt_scene_desc d{};
    d.n_prims = (uint32_t)prims.size();
    d.prims = prims.data();
    d.n_materials = (uint32_t)materials.size();
    d.materials = materials.data();
    d.n_lights = (uint32_t)lights.size();
    d.lights = lights.data();
    d.n_triangles = triangles.size() / 9;
    d.triangles = triangles.empty() ? nullptr : triangles.data();
    d.triangle_material = triangle_material;
    d.camera = camera;
    return d;
}

DeviceContext::DeviceContext(const SceneBuilder& b, int device) : desc_(b) {
    if (desc_.materials.empty()) desc_.materials.push_back(lambert());
    if (!desc_.has_camera) { // a camera is not needed for single-ray queries; give the description a valid one
        const float p[3] = {0, -3, 0}, d[3] = {0, 1, 0}, u[3] = {0, 0, 1};
        check(ipt_camera_look(p, d, u, &desc_.camera));
    }
    ipt_scene_desc d = desc_.desc();
    check(ipt_scene_create(&d, device, &scene_));
}
DeviceContext::~DeviceContext() {
    if (scene_) ipt_scene_destroy(scene_);
}

// ---- DDFs handed out through the interfaces: evaluated on the device ------------------------------------
class DeviceSdf : public Ddf {
public:
    DeviceSdf(std::shared_ptr<DeviceContext> ctx, ipt_material m, glm::vec3 n, glm::vec3 refl) : ctx_(std::move(ctx)), m_(m), n_(n), refl_(refl) {}
    glm::vec3 sample() const override {
        if (m_.ddf == IPT_DDF_COSINE) return one(2, n_);
        // kd*Lambert + ks*PowerCosine about the mirror direction; below-surface directions are failed samples
        float wd = m_.kd / (m_.kd + m_.ks);
        uint64_t c = g_sample_counter.fetch_add(1);
        float u = (float)((c * 0x9E3779B97F4A7C15ull) >> 40) * (1.0f / 16777216.0f);
        glm::vec3 w = u < wd ? one(2, n_) : one((int)m_.exponent, refl_);
        if (H::dot(h3(n_), h3(w)) < 0.0f) return glm::vec3();
        return w;
    }
    float value(glm::vec3 w) const override {
        if (m_.ddf == IPT_DDF_COSINE) return val(2, n_, w);
        if (H::dot(h3(n_), h3(w)) < 0.0f) return 0.0f;
        float wd = m_.kd / (m_.kd + m_.ks), ws = m_.ks / (m_.kd + m_.ks);
        return wd * val(2, n_, w) + ws * val((int)m_.exponent, refl_, w);
    }

private:
    glm::vec3 one(int kind, glm::vec3 to) const {
        float t[3], out[3];
        put(t, to);
        check(ipt_ddf_sample(ctx_->scene(), kind, t, g_sample_counter.fetch_add(1), 1, out));
        return glm::vec3(out[0], out[1], out[2]);
    }
    float val(int kind, glm::vec3 to, glm::vec3 w) const {
        float t[3], d[3], out;
        put(t, to);
        put(d, w);
        check(ipt_ddf_value(ctx_->scene(), kind, t, d, 1, &out));
        return out;
    }
    std::shared_ptr<DeviceContext> ctx_;
    ipt_material m_;
    glm::vec3 n_, refl_;
};

class DeviceLightDdf : public Ddf {
public:
    DeviceLightDdf(std::shared_ptr<DeviceContext> ctx, glm::vec3 pos) : ctx_(std::move(ctx)), pos_(pos) {}
    glm::vec3 sample() const override {
        float p[3], out[3];
        put(p, pos_);
        check(ipt_light_ddf_sample(ctx_->scene(), p, g_sample_counter.fetch_add(1), 1, out));
        return glm::vec3(out[0], out[1], out[2]);
    }
    float value(glm::vec3 w) const override {
        float p[3], d[3], out;
        put(p, pos_);
        put(d, w);
        check(ipt_light_ddf_value(ctx_->scene(), p, d, 1, &out));
        return out;
    }

private:
    std::shared_ptr<DeviceContext> ctx_;
    glm::vec3 pos_;
};

// ---- DeviceGeometry -------------------------------------------------------------------------------------
uint32_t DeviceGeometry::addMaterial(const ipt_material& m) {
    data_.materials.push_back(m);
    ctx_.reset();
    ++revision_;
    return (uint32_t)data_.materials.size() - 1;
}
void DeviceGeometry::addBoxPlane(glm::vec3 plane, uint32_t material) {
    ipt_prim p{};
    p.kind = IPT_PRIM_BOX_PLANE;
    p.material = material;
    put(p.p, plane);
    data_.prims.push_back(p);
    ctx_.reset();
    ++revision_;
}
void DeviceGeometry::addSphere(glm::vec3 centre, float radius, float curvature, uint32_t material) {
    ipt_prim p{};
    p.kind = IPT_PRIM_SPHERE;
    p.material = material;
    put(p.p, centre);
    p.radius = radius;
    p.curvature = curvature;
    data_.prims.push_back(p);
    ctx_.reset();
    ++revision_;
}
void DeviceGeometry::addSmallPtSphere(glm::vec3 centre, float radius, uint32_t material) {
    ipt_prim p{};
    p.kind = IPT_PRIM_SPHERE_SMALLPT;
    p.material = material;
    put(p.p, centre);
    p.radius = radius;
    p.flip_normal = radius < 100 ? 0 : 1; // GeometrySmallPt.cpp:53
    p.curvature = (float)(-1.0 / (double)radius);
    data_.prims.push_back(p);
    ctx_.reset();
    ++revision_;
}
void DeviceGeometry::setTriangles(const float* t, size_t count, uint32_t material) {
    data_.triangles.assign(t, t + 9 * count);
    data_.triangle_material = material;
    ctx_.reset();
    ++revision_;
}
std::shared_ptr<DeviceGeometry> DeviceGeometry::fromSampleScene(const char* name) {
    ipt_scene_desc* d = nullptr;
    check(ipt_sample_scene(name, &d));
    auto g = std::make_shared<DeviceGeometry>();
    g->data_.prims.assign(d->prims, d->prims + d->n_prims);
    g->data_.materials.assign(d->materials, d->materials + d->n_materials);
    if (d->n_triangles) g->data_.triangles.assign(d->triangles, d->triangles + 9 * d->n_triangles);
    g->data_.triangle_material = d->triangle_material;
    ipt_scene_desc_free(d);
    return g;
}
void DeviceGeometry::exportTo(SceneBuilder& b) const {
    b.prims = data_.prims;
    b.materials = data_.materials.empty() ? std::vector<ipt_material>{lambert()} : data_.materials;
    b.triangles = data_.triangles;
    b.triangle_material = data_.triangle_material;
}
DeviceContext& DeviceGeometry::context() const {
    std::lock_guard<std::mutex> lock(mu_);
    if (!ctx_) {
        SceneBuilder b;
        exportTo(b);
        ctx_ = std::make_shared<DeviceContext>(b);
    }
    return *ctx_;
}
std::optional<surface_intersection> DeviceGeometry::traceRay(glm::vec3 origin, glm::vec3 direction) const {
    DeviceContext& c = context();
    float o[3], d[3], t;
    uint32_t prim;
    put(o, origin);
    put(d, direction);
    check(ipt_trace_batch(c.scene(), o, d, 1, &prim, &t, nullptr, nullptr, nullptr));
    if (prim == IPT_NO_HIT) return {};
    const SceneBuilder& sb = c.description();
    surface_intersection res;
    res.position = g3(h3(origin) + h3(direction) * t); // origin + direction*dist (GeometrySphereInBox.cpp:42)
    uint32_t material;
    if (prim >= sb.prims.size()) {
        const float* tri = sb.triangles.data() + 9 * (size_t)(prim - sb.prims.size());
        res.normal = g3(H::normalize(H::cross(H::mk(tri + 3), H::mk(tri + 6))));
        res.curvature = 0.0f;
        material = sb.triangle_material;
    } else {
        const ipt_prim& p = sb.prims[prim];
        material = p.material;
        res.curvature = p.curvature;
        if (p.kind == IPT_PRIM_BOX_PLANE) res.normal = glm::vec3(-p.p[0], -p.p[1], -p.p[2]);
        else {
            H::f3 n = H::normalize(h3(res.position) - H::mk(p.p));
            res.normal = g3(p.flip_normal ? -n : n);
        }
    }
    const ipt_material& m = sb.materials[material];
    H::f3 I = h3(direction), N = h3(res.normal);
    glm::vec3 refl = g3(I - N * H::dot(N, I) * 2.0f); // glm::reflect
    res.sdf = std::make_unique<DeviceSdf>(ctx_, m, res.normal, refl);
    res.albedo = m.albedo;
    return res;
}

// ---- DeviceLighting -------------------------------------------------------------------------------------
void DeviceLighting::addPointLight(glm::vec3 position, float, float power) {
    ipt_light l{};
    l.kind = IPT_LIGHT_POINT;
    put(l.position, position);
    l.power = power;
    addLight(l);
}
void DeviceLighting::addSphereLight(glm::vec3 position, float radius, float power) {
    ipt_light l{};
    l.kind = IPT_LIGHT_SPHERE;
    put(l.position, position);
    l.radius = radius;
    l.power = power;
    addLight(l);
}
void DeviceLighting::addSquareLight(glm::vec3 corner, glm::vec3 normal, glm::vec3 x_side, float power) {
    ipt_light l{};
    l.kind = IPT_LIGHT_AREA_DIAMOND;
    put(l.position, corner);
    put(l.x_axis, x_side);
    put(l.y_axis, g3(H::cross(h3(normal), h3(x_side)))); // CollectionLighting.cpp:43
    l.power = power;
    addLight(l);
}
void DeviceLighting::addTriangleLight(glm::vec3 corner, glm::vec3 x_side, glm::vec3 y_side, float power) {
    ipt_light l{};
    l.kind = IPT_LIGHT_AREA_TRIANGLE;
    put(l.position, corner);
    put(l.x_axis, x_side);
    put(l.y_axis, y_side);
    l.power = power;
    addLight(l);
}
void DeviceLighting::addOuterLight(float radius, float power) {
    ipt_light l{};
    l.kind = IPT_LIGHT_SPHERE_INVERTED; // CollectionLighting.cpp:52-55
    l.radius = radius;
    l.power = power;
    addLight(l);
}
void DeviceLighting::exportTo(SceneBuilder& b) const { b.lights = lights_; }
DeviceContext& DeviceLighting::context() const {
    std::lock_guard<std::mutex> lock(mu_);
    if (!ctx_) {
        SceneBuilder b;
        exportTo(b);
        ctx_ = std::make_shared<DeviceContext>(b);
    }
    return *ctx_;
}
std::unique_ptr<Ddf> DeviceLighting::distributionInPoint(glm::vec3 pos) const {
    context();
    return std::make_unique<DeviceLightDdf>(ctx_, pos);
}
std::optional<light_intersection> DeviceLighting::traceRayToLight(glm::vec3 origin, glm::vec3 direction) const {
    DeviceContext& c = context();
    float o[3], d[3], lp[3];
    uint32_t light;
    put(o, origin);
    put(d, direction);
    check(ipt_trace_batch(c.scene(), o, d, 1, nullptr, nullptr, &light, lp, nullptr));
    if (light == IPT_NO_HIT) return {};
    const ipt_light& l = lights_[light];
    light_intersection res;
    res.position = glm::vec3(lp[0], lp[1], lp[2]);
    float area, sp, n[3];
    check(ipt_light_derived(&l, &area, &sp, n));
    res.surface_power = sp;
    if (l.kind <= IPT_LIGHT_AREA_TRIANGLE) res.normal = glm::vec3(n[0], n[1], n[2]);
    else {
        H::f3 nn = H::normalize(h3(res.position) - H::mk(l.position));
        res.normal = g3(l.kind == IPT_LIGHT_SPHERE_INVERTED ? -nn : nn);
    }
    return res;
}

// ---- DeviceCamera -----------------------------------------------------------------------------------------
DeviceCamera::DeviceCamera(glm::vec3 p, glm::vec3 d, glm::vec3 up_hint) : position(p), direction(d) {
    float pp[3], dd[3], uu[3];
    put(pp, p);
    put(dd, d);
    put(uu, up_hint);
    ipt_camera c;
    check(ipt_camera_look(pp, dd, uu, &c));
    right = glm::vec3(c.right[0], c.right[1], c.right[2]);
    up = glm::vec3(c.up[0], c.up[1], c.up[2]);
}
void DeviceCamera::exportTo(SceneBuilder& b) const {
    put(b.camera.position, position);
    put(b.camera.direction, direction);
    put(b.camera.right, right);
    put(b.camera.up, up);
    b.has_camera = true;
}
void DeviceCamera::orbit(int key) {
    SceneBuilder b;
    exportTo(b);
    check(ipt_camera_orbit(&b.camera, key));
    position = glm::vec3(b.camera.position[0], b.camera.position[1], b.camera.position[2]);
    direction = glm::vec3(b.camera.direction[0], b.camera.direction[1], b.camera.direction[2]);
    right = glm::vec3(b.camera.right[0], b.camera.right[1], b.camera.right[2]);
    up = glm::vec3(b.camera.up[0], b.camera.up[1], b.camera.up[2]);
}
std::pair<glm::vec3, glm::vec3> DeviceCamera::sampleRay(float x, float y) const {
    SceneBuilder b;
    exportTo(b);
    if (!ctx_) ctx_ = std::make_shared<DeviceContext>(b);
    check(ipt_scene_set_camera(ctx_->scene(), &b.camera)); // the fields are public and may have been changed
    float xy[2] = {x, y}, o[3], d[3];
    check(ipt_camera_rays(ctx_->scene(), xy, 1, o, d));
    return {glm::vec3(o[0], o[1], o[2]), glm::vec3(d[0], d[1], d[2])};
}

// ---- DevicePlane --------------------------------------------------------------------------------------------
DevicePlane::DevicePlane(size_t w, size_t h) : width(w), height(h) {
    pixels.resize(w * h);
    pixel_counters.resize(w * h);
}
DevicePlane::~DevicePlane() {
    if (plane_) ipt_plane_destroy(plane_);
}
ipt_plane* DevicePlane::attach(const std::shared_ptr<DeviceContext>& ctx) {
    if (plane_ && ctx_ == ctx) return plane_;
    std::vector<float> s, q;
    std::vector<uint32_t> c;
    bool carry = plane_ != nullptr;
    if (carry) {
        sums(s, q, c);
        ipt_plane_destroy(plane_);
        plane_ = nullptr;
    }
    ctx_ = ctx;
    check(ipt_plane_create(ctx_->scene(), (uint32_t)width, (uint32_t)height, &plane_));
    if (carry) check(ipt_plane_upload(plane_, s.data(), q.data(), c.data()));
    return plane_;
}
void DevicePlane::clear() {
    if (plane_) check(ipt_plane_clear(plane_));
}
void DevicePlane::upload(const std::vector<float>& sum, const std::vector<float>& sumsq, const std::vector<uint32_t>& count) {
    if (sum.size() != width * height || sumsq.size() != sum.size() || count.size() != sum.size()) throw Error(IPT_ERR_INVALID, "DevicePlane::upload: array sizes do not match the plane");
    if (!plane_) attach(std::make_shared<DeviceContext>(SceneBuilder())); // the next render carries the accumulators over to its scene
    check(ipt_plane_upload(plane_, sum.data(), sumsq.data(), count.data()));
}
void DevicePlane::addRay(float x, float y, float value) {
    if (!plane_) attach(std::make_shared<DeviceContext>(SceneBuilder()));
    check(ipt_plane_add_rays(plane_, plane_mode, 1, &x, &y, &value));
}
void DevicePlane::sums(std::vector<float>& sum, std::vector<float>& sumsq, std::vector<uint32_t>& count) {
    sum.assign(width * height, 0.0f);
    sumsq.assign(width * height, 0.0f);
    count.assign(width * height, 0u);
    if (plane_) check(ipt_plane_download(plane_, sum.data(), sumsq.data(), count.data()));
}
std::vector<float> DevicePlane::display(float glare_cutoff) {
    if (!plane_) throw Error{IPT_ERR_INVALID, "DevicePlane::display: nothing rendered yet"};
    std::vector<float> out(width * height);
    check(ipt_plane_display(plane_, glare_cutoff, out.data(), nullptr));
    return out;
}
void DevicePlane::save(const char* path) {
    if (!plane_) throw Error{IPT_ERR_INVALID, "DevicePlane::save: nothing rendered yet"};
    check(ipt_plane_save_png(plane_, path));
}
void DevicePlane::download() {
    if (!plane_) return;
    std::vector<uint64_t> cnt(width * height);
    check(ipt_plane_resolve(plane_, pixels.data(), cnt.data(), &max_value));
    for (size_t i = 0; i < cnt.size(); ++i) pixel_counters[i] = (size_t)cnt[i];
}

// ---- factory ------------------------------------------------------------------------------------------------
Scene make_scene(const char* name) {
    ipt_scene_desc* d = nullptr;
    check(ipt_sample_scene(name, &d));
    auto lighting = std::make_shared<DeviceLighting>();
    for (uint32_t i = 0; i < d->n_lights; ++i) lighting->addLight(d->lights[i]);
    auto camera = std::make_shared<DeviceCamera>(glm::vec3(0, 0, 0), glm::vec3(0, 1, 0));
    camera->position = glm::vec3(d->camera.position[0], d->camera.position[1], d->camera.position[2]);
    camera->direction = glm::vec3(d->camera.direction[0], d->camera.direction[1], d->camera.direction[2]);
    camera->right = glm::vec3(d->camera.right[0], d->camera.right[1], d->camera.right[2]);
    camera->up = glm::vec3(d->camera.up[0], d->camera.up[1], d->camera.up[2]);
    ipt_scene_desc_free(d);
    return Scene{DeviceGeometry::fromSampleScene(name), lighting, camera};
}

// ---- the hot path ---------------------------------------------------------------------------------------------
namespace {
// Device replicas of the scenes seen so far, keyed by the identity of the three Scene members and the device. The entry
// keeps weak references: when a member has been destroyed (and its address possibly reused by another object) the entry
// is stale and is rebuilt.
struct CachedContext {
    std::weak_ptr<const Geometry> geometry;
    std::weak_ptr<const Lighting> lighting;
    std::weak_ptr<const Camera> camera;
    std::shared_ptr<DeviceContext> ctx;
    uint64_t geometry_revision = 0, lighting_revision = 0;
    bool alive() const { return !geometry.expired() && !lighting.expired() && !camera.expired(); }
};
std::mutex g_cache_mu;
std::map<std::tuple<const void*, const void*, const void*, int, size_t>, CachedContext> g_cache;

#ifdef IPT_B200_REFERENCE_CLASSES
// The reference's data-free geometries: the class IS the scene (its primitives are literals inside traceRay).
const char* reference_geometry_name(const Geometry* g) {
    if (dynamic_cast<const GeometrySphereInBox*>(g)) return "box";       // GeometrySphereInBox.cpp:11-17
    if (dynamic_cast<const GeometryFloor*>(g)) return "square";          // GeometryFloor.cpp:10-23
    if (dynamic_cast<const GeometryCorner*>(g)) return "corner";         // GeometryCorner.cpp:10-39
    if (dynamic_cast<const GeometryOpenSpheres*>(g)) return "openspheres"; // GeometryOpenSpheres.cpp:12-68
    if (dynamic_cast<const FractalSpheres*>(g)) return "fractal";        // FractalSpheres.cpp:16-97
    if (dynamic_cast<const GeometrySmallPt*>(g)) return "smallpt";       // GeometrySmallPt.cpp:12-61
    return nullptr;
}
#endif

bool export_geometry(const Geometry* g, SceneBuilder& b) {
    if (const DeviceExportable* e = dynamic_cast<const DeviceExportable*>(g)) { e->exportTo(b); return true; }
#ifdef IPT_B200_REFERENCE_CLASSES
    if (const char* name = reference_geometry_name(g)) { DeviceGeometry::fromSampleScene(name)->exportTo(b); return true; }
#endif
    return false;
}
bool export_camera(const Camera* c, SceneBuilder& b) {
    if (const DeviceExportable* e = dynamic_cast<const DeviceExportable*>(c)) { e->exportTo(b); return true; }
#ifdef IPT_B200_REFERENCE_CLASSES
    if (const SimpleCamera* sc = dynamic_cast<const SimpleCamera*>(c)) {
        put(b.camera.position, sc->position);
        put(b.camera.direction, sc->direction);
        put(b.camera.right, sc->right);
        put(b.camera.up, sc->up);
        b.has_camera = true;
        return true;
    }
#endif
    return false;
}

// `replica`: which of the slots naming the SAME device this is (render_sample(scene, plane, params, {0, 0}) runs two host
// threads on one GPU): an ipt_scene owns one stream and one workspace, so concurrent renders need one scene replica each.
std::shared_ptr<DeviceContext> context_for(const Scene& scene, int device, size_t replica = 0) {
    SceneBuilder b;
    const DeviceExportable* l = dynamic_cast<const DeviceExportable*>(scene.lighting.get());
    if (!l || !export_geometry(scene.geometry.get(), b) || !export_camera(scene.camera.get(), b))
        throw Error(IPT_ERR_UNSUPPORTED,
                    "ipt_b200::render_sample: the Scene members must be DeviceExportable (DeviceGeometry / DeviceLighting / DeviceCamera) or, "
                    "for geometry and camera in a build against the reference's headers, the reference's own data-free Geometry classes / "
                    "SimpleCamera; other classes hide their data and there is no CPU fallback");
    std::lock_guard<std::mutex> lock(g_cache_mu);
    auto key = std::make_tuple((const void*)scene.geometry.get(), (const void*)scene.lighting.get(), (const void*)scene.camera.get(), device, replica);
    for (auto e = g_cache.begin(); e != g_cache.end();) e = e->second.alive() ? std::next(e) : g_cache.erase(e); // drop dead scenes
    const DeviceExportable* ge = dynamic_cast<const DeviceExportable*>(scene.geometry.get());
    const uint64_t grev = ge ? ge->revision() : 0, lrev = l->revision();
    auto it = g_cache.find(key);
    if (it != g_cache.end() && (it->second.geometry_revision != grev || it->second.lighting_revision != lrev)) { g_cache.erase(it); it = g_cache.end(); }
    if (it == g_cache.end()) {
        l->exportTo(b);
        it = g_cache.emplace(key, CachedContext{scene.geometry, scene.lighting, scene.camera, std::make_shared<DeviceContext>(b, device), grev, lrev}).first;
    }
    // the camera may be orbited between calls (gui.cpp:107-137): always refresh it
    check(ipt_scene_set_camera(it->second.ctx->scene(), &b.camera));
    return it->second.ctx;
}

// sum / count of loop-pixel cells -> a plane the library does not own
void flush_foreign(RenderPlane& r_plane, const ipt_render_params& p, const std::vector<float>& sum, const std::vector<uint32_t>& count) {
    for (uint32_t iy = 0; iy < p.height; ++iy)
        for (uint32_t ix = 0; ix < p.width; ++ix) {
            size_t i = (size_t)iy * p.width + ix;
            if (count[i]) r_plane.addRay((ix + 0.5f) / p.width, (iy + 0.5f) / p.height, sum[i] / count[i]);
        }
}

#ifdef IPT_B200_REFERENCE_CLASSES
// What count[i] calls of GridRenderPlane::addRay with values adding up to sum[i] leave in cell i (GridRenderPlane.cpp:68-74):
// the running mean over old and new samples, the counter, and max_value over the cells written.
void flush_grid(GridRenderPlane& g, const std::vector<float>& sum, const std::vector<uint32_t>& count) {
    for (size_t i = 0; i < count.size(); ++i) {
        if (!count[i]) continue;
        size_t before = g.pixel_counters[i];
        g.pixels[i] = before ? (g.pixels[i] * (float)before + sum[i]) / (float)(before + count[i]) : sum[i] / (float)count[i];
        g.pixel_counters[i] = before + count[i];
        if (g.pixels[i] > g.max_value) g.max_value = g.pixels[i];
    }
}
#endif

// contiguous, balanced pass ranges (the arithmetic of ipt_b200/sharding.py:shard_passes)
void shard_passes(uint32_t total, size_t world, size_t rank, uint32_t first, uint32_t& begin, uint32_t& count) {
    uint32_t base = total / (uint32_t)world, extra = total % (uint32_t)world;
    count = base + (rank < extra ? 1u : 0u);
    begin = first + (uint32_t)rank * base + (uint32_t)std::min<size_t>(rank, extra);
}

void add_stats(ipt_render_stats& a, const ipt_render_stats& b) {
    a.paths += b.paths; a.rays += b.rays;
    for (int d = 0; d < IPT_MAX_DEPTH; ++d) a.rays_at_depth[d] += b.rays_at_depth[d];
    a.surface_hits += b.surface_hits; a.light_hits += b.light_hits; a.misses += b.misses;
    a.failed_samples += b.failed_samples; a.zero_weight_pruned += b.zero_weight_pruned; a.nonfinite_dropped += b.nonfinite_dropped;
    a.bvh_nodes_visited += b.bvh_nodes_visited; a.triangles_tested += b.triangles_tested; a.lights_tested += b.lights_tested;
    a.light_bvh_nodes_visited += b.light_bvh_nodes_visited;
    a.batches += b.batches; a.kernel_launches += b.kernel_launches;
    a.ms_total = std::max(a.ms_total, b.ms_total); // the devices run side by side
    a.queue_bytes += b.queue_bytes; a.rays_resolved_in_shade += b.rays_resolved_in_shade;
}

struct PlaneHandle { // a scratch plane on some device, destroyed with the scope
    ipt_plane* p = nullptr;
    ~PlaneHandle() { if (p) ipt_plane_destroy(p); }
};
} // namespace

ipt_render_stats render_sample(const Scene& scene, RenderPlane& r_plane, const ipt_render_params& params, int device) {
    return render_sample(scene, r_plane, params, std::vector<int>{device});
}

ipt_render_stats render_sample(const Scene& scene, RenderPlane& r_plane, const ipt_render_params& params, const std::vector<int>& devices) {
    if (devices.empty()) throw Error(IPT_ERR_INVALID, "ipt_b200::render_sample: no device given");
    DevicePlane* dp = dynamic_cast<DevicePlane*>(&r_plane);
#ifdef IPT_B200_REFERENCE_CLASSES
    GridRenderPlane* gp = dynamic_cast<GridRenderPlane*>(&r_plane);
    if (gp && (gp->width != params.width || gp->height != params.height)) throw Error(IPT_ERR_INVALID, "GridRenderPlane size does not match render params");
#else
    const bool gp = false;
#endif
    ipt_render_params p = params;
    p.plane_mode = dp ? dp->plane_mode : gp ? (uint32_t)IPT_PLANE_GRID : (uint32_t)IPT_PLANE_LINEAR;
    const size_t world = devices.size();
    std::vector<std::shared_ptr<DeviceContext>> ctx(world);
    for (size_t r = 0; r < world; ++r) // scene replicas (cached per device; slots that name the same device get one each)
        ctx[r] = context_for(scene, devices[r], (size_t)std::count(devices.begin(), devices.begin() + r, devices[r]));
    // accumulators: device 0 renders into the caller's DevicePlane (or a scratch plane), the others into scratch planes
    std::vector<PlaneHandle> scratch(world);
    std::vector<ipt_plane*> plane(world, nullptr);
    for (size_t r = 0; r < world; ++r) {
        if (r == 0 && dp) plane[r] = dp->attach(ctx[0]);
        else {
            check(ipt_plane_create(ctx[r]->scene(), p.width, p.height, &scratch[r].p));
            plane[r] = scratch[r].p;
        }
    }
    std::vector<ipt_render_stats> st(world);
    std::vector<std::string> err(world);
    std::vector<int> code(world, IPT_OK);
    auto work = [&](size_t r) {
        ipt_render_params pr = p;
        shard_passes(p.pass_count, world, r, p.pass_begin, pr.pass_begin, pr.pass_count);
        std::memset(&st[r], 0, sizeof st[r]);
        if (pr.pass_count == 0) return;
        int rc = ipt_render(ctx[r]->scene(), plane[r], &pr, &st[r]);
        if (rc != IPT_OK) { code[r] = rc; err[r] = ipt_last_error(); } // ipt_last_error is per thread: fetch it here
    };
    if (world == 1) work(0);
    else {
        std::vector<std::thread> threads;
        for (size_t r = 0; r < world; ++r) threads.emplace_back(work, r);
        for (std::thread& t : threads) t.join();
    }
    for (size_t r = 0; r < world; ++r)
        if (code[r] != IPT_OK) throw Error(code[r], "ipt_b200: device " + std::to_string(devices[r]) + ": " + err[r]);
    ipt_render_stats stats = st[0];
    for (size_t r = 1; r < world; ++r) {
        check(ipt_plane_merge(plane[0], plane[r])); // the only exchange: 12 B per cell and device
        add_stats(stats, st[r]);
    }
    if (dp) return stats;
    // a plane the library does not own: read the accumulators back once and write the caller's plane
    size_t n = (size_t)p.width * p.height;
    std::vector<float> sum(n), sumsq(n);
    std::vector<uint32_t> count(n);
    check(ipt_plane_download(plane[0], sum.data(), sumsq.data(), count.data()));
#ifdef IPT_B200_REFERENCE_CLASSES
    if (gp) { flush_grid(*gp, sum, count); return stats; }
#endif
    flush_foreign(r_plane, p, sum, count);
    return stats;
}


// ---- ProgressiveSession ---------------------------------------------------------------------------------------
ProgressiveSession::ProgressiveSession(Scene scene, size_t width, size_t height, uint32_t passes_per_call, std::vector<int> devices)
    : scene_(std::move(scene)), plane_(width, height), devices_(std::move(devices)), passes_per_call_(passes_per_call ? passes_per_call : 1) {
    ipt_render_params_default(&params);
    params.width = (uint32_t)width;
    params.height = (uint32_t)height;
    plane_.plane_mode = IPT_PLANE_GUI; // the interactive window's cell mapping (Gui::addRay, gui.cpp:168-172)
}
uint64_t ProgressiveSession::step(unsigned calls) {
    for (unsigned c = 0; c < calls; ++c) {
        ipt_render_params p = params;
        if (next_pass_ + passes_per_call_ > 0xFFFFFFFFull) throw Error(IPT_ERR_UNSUPPORTED, "ProgressiveSession: pass counter exhausted");
        p.pass_begin = (uint32_t)next_pass_;
        p.pass_count = passes_per_call_;
        ipt_render_stats st = render_sample(scene_, plane_, p, devices_);
        next_pass_ += passes_per_call_;
        rays_ += st.rays;
    }
    return samples_per_pixel();
}
void ProgressiveSession::resetImage() { // pass numbers keep growing: the new image draws fresh random streams
    plane_.clear();
    first_pass_ = next_pass_;
}
void ProgressiveSession::key(int key) {
    DeviceCamera* cam = dynamic_cast<DeviceCamera*>(const_cast<Camera*>(scene_.camera.get()));
    if (!cam) throw Error(IPT_ERR_UNSUPPORTED, "ProgressiveSession::key: the scene's camera is not a DeviceCamera");
    cam->orbit(key);
    resetImage();
}
void ProgressiveSession::wheel(int clicks) { // glare_cutoff *= pow(sqrt(2.0f), wheel): float sqrt, double pow and product (gui.cpp:142)
    glare_cutoff = (float)((double)glare_cutoff * std::pow((double)std::sqrt(2.0f), clicks));
}

namespace {
struct CheckpointHeader { // little-endian, fixed layout; followed by sum[n], sumsq[n] (float) and count[n] (uint32)
    char magic[8];
    uint32_t width, height, depth_max, plane_mode, passes_per_call, schedule[IPT_MAX_DEPTH];
    uint64_t seed, next_pass, first_pass, rays;
    float glare_cutoff, camera[12];
    uint32_t has_camera, reserved;
};
const char kCheckpointMagic[8] = {'I', 'P', 'T', 'C', 'K', 'P', 'T', '1'};
} // namespace

void ProgressiveSession::checkpoint(const std::string& path) {
    CheckpointHeader h{};
    std::memcpy(h.magic, kCheckpointMagic, 8);
    h.width = params.width; h.height = params.height; h.depth_max = params.depth_max; h.plane_mode = plane_.plane_mode;
    h.passes_per_call = passes_per_call_;
    std::memcpy(h.schedule, params.schedule, sizeof h.schedule);
    h.seed = params.seed; h.next_pass = next_pass_; h.first_pass = first_pass_; h.rays = rays_;
    h.glare_cutoff = glare_cutoff;
    if (const DeviceCamera* cam = dynamic_cast<const DeviceCamera*>(scene_.camera.get())) {
        const glm::vec3* v[4] = {&cam->position, &cam->direction, &cam->right, &cam->up};
        for (int k = 0; k < 4; ++k) { h.camera[3 * k] = v[k]->x; h.camera[3 * k + 1] = v[k]->y; h.camera[3 * k + 2] = v[k]->z; }
        h.has_camera = 1;
    }
    std::vector<float> sum, sumsq;
    std::vector<uint32_t> count;
    plane_.sums(sum, sumsq, count);
    const std::string tmp = path + ".tmp";
    FILE* f = std::fopen(tmp.c_str(), "wb");
    if (!f) throw Error(IPT_ERR_INVALID, "ProgressiveSession::checkpoint: cannot write " + tmp);
    bool ok = std::fwrite(&h, sizeof h, 1, f) == 1 && std::fwrite(sum.data(), 4, sum.size(), f) == sum.size() &&
              std::fwrite(sumsq.data(), 4, sumsq.size(), f) == sumsq.size() && std::fwrite(count.data(), 4, count.size(), f) == count.size();
    ok = (std::fclose(f) == 0) && ok;
    if (!ok || std::rename(tmp.c_str(), path.c_str()) != 0) { std::remove(tmp.c_str()); throw Error(IPT_ERR_INVALID, "ProgressiveSession::checkpoint: write to " + path + " failed"); }
}
bool ProgressiveSession::resume(const std::string& path) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    CheckpointHeader h{};
    const size_t n = (size_t)params.width * params.height;
    std::vector<float> sum(n), sumsq(n);
    std::vector<uint32_t> count(n);
    bool ok = std::fread(&h, sizeof h, 1, f) == 1 && std::memcmp(h.magic, kCheckpointMagic, 8) == 0;
    // a checkpoint belongs to one estimator and one frame: resuming it under other parameters would mix two images
    const bool same = ok && h.width == params.width && h.height == params.height && h.depth_max == params.depth_max &&
                      h.plane_mode == plane_.plane_mode && h.passes_per_call == passes_per_call_ &&
                      std::memcmp(h.schedule, params.schedule, sizeof h.schedule) == 0;
    if (same) ok = std::fread(sum.data(), 4, n, f) == n && std::fread(sumsq.data(), 4, n, f) == n && std::fread(count.data(), 4, n, f) == n;
    std::fclose(f);
    if (!ok) throw Error(IPT_ERR_INVALID, "ProgressiveSession::resume: " + path + " is not a complete ipt_b200 checkpoint");
    if (!same) throw Error(IPT_ERR_INVALID, "ProgressiveSession::resume: " + path + " was written for another frame size, depth, split schedule, plane mode or step size");
    params.seed = h.seed;
    next_pass_ = h.next_pass; first_pass_ = h.first_pass; rays_ = h.rays;
    glare_cutoff = h.glare_cutoff;
    if (h.has_camera)
        if (DeviceCamera* cam = dynamic_cast<DeviceCamera*>(const_cast<Camera*>(scene_.camera.get()))) {
            cam->position = glm::vec3(h.camera[0], h.camera[1], h.camera[2]); cam->direction = glm::vec3(h.camera[3], h.camera[4], h.camera[5]);
            cam->right = glm::vec3(h.camera[6], h.camera[7], h.camera[8]); cam->up = glm::vec3(h.camera[9], h.camera[10], h.camera[11]);
        }
    plane_.upload(sum, sumsq, count);
    return true;
}

} // namespace ipt_b200
