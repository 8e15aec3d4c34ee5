// device_plugins.hpp — C++ host classes that put the CUDA trace loop behind the reference's own abstractions
// (src/tracer_interfaces.h:26-54): Geometry, Lighting, Camera, RenderPlane, Scene.
//
// The reference's loop is   render_sample(const Scene&, RenderPlane&, StatsNode*)   (src/main.cpp:186-223), called in
// a loop from main.cpp:264. The drop-in is
//
//     ipt_b200::render_sample(scene, plane, params);          // params.pass_count calls of render_sample at once
//
// with `scene` built from the classes below, which mirror the reference's factory API one to one
// (CollectionLighting::add*Light, SimpleCamera, the Geometry* classes) but, unlike the reference's classes
// (SURVEY.md S12), can describe themselves to the device (`DeviceExportable`).
//
// Every virtual of the reference interfaces is implemented by CALLING THE DEVICE through the C ABI for that one
// ray / point (there is no CPU implementation of the arithmetic anywhere in this library); they exist so that the
// objects remain usable by code written against tracer_interfaces.h (e.g. ray_power_preview, main.cpp:55-92), not
// for throughput. Include path decides which interface header is used: the reference's own
// (-I<reference>/src -I<reference>/include first) or the mirror in host/compat/.
#pragma once
#include "tracer_interfaces.h"

#include "ipt_b200.h"

#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

namespace ipt_b200 {

struct Error : public std::runtime_error {
    int code;
    Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};
void check(int status); // throws Error with ipt_last_error()

// What a scene member contributes to the flat description the device consumes.
struct SceneBuilder {
    std::vector<ipt_prim> prims;
    std::vector<ipt_material> materials;
    std::vector<ipt_light> lights;
    std::vector<float> triangles;
    uint32_t triangle_material = 0;
    ipt_camera camera{};
    bool has_camera = false;
    ipt_scene_desc desc() const;
};
struct DeviceExportable {
    virtual void exportTo(SceneBuilder& b) const = 0;
    // bumped by every mutation that changes what exportTo() writes (cameras excepted: they are re-read on every call), so
    // that render_sample's cached device replica of the scene is rebuilt when an object was edited after a render
    virtual uint64_t revision() const { return 0; }
    virtual ~DeviceExportable() {}
};

// Shared device state of one Scene (scene handle + lazily created planes), created on first use.
class DeviceContext {
public:
    explicit DeviceContext(const SceneBuilder& b, int device = 0);
    ~DeviceContext();
    ipt_scene* scene() const { return scene_; }
    const SceneBuilder& description() const { return desc_; }

private:
    SceneBuilder desc_;
    ipt_scene* scene_ = nullptr;
};

// ---- Geometry ---------------------------------------------------------------------------------------
// One class for every reference geometry (GeometrySphereInBox, GeometryFloor, GeometryCorner, GeometryOpenSpheres,
// FractalSpheres, GeometrySmallPt) and the mesh extension: an ordered primitive list.
class DeviceGeometry : public Geometry, public DeviceExportable {
public:
    uint32_t addMaterial(const ipt_material& m);
    void addBoxPlane(glm::vec3 plane, uint32_t material = 0);                              // geometric_utils.cpp:8
    void addSphere(glm::vec3 centre, float radius, float curvature, uint32_t material = 0); // geometric_utils.cpp:28
    void addSmallPtSphere(glm::vec3 centre, float radius, uint32_t material = 0);          // GeometrySmallPt.cpp:13-23
    void setTriangles(const float* v0_e1_e2, size_t count, uint32_t material = 0);
    static std::shared_ptr<DeviceGeometry> fromSampleScene(const char* name);              // "box", "cornell", ...

    std::optional<surface_intersection> traceRay(glm::vec3 origin, glm::vec3 direction) const override;
    void exportTo(SceneBuilder& b) const override;
    uint64_t revision() const override { return revision_; }

private:
    friend class DeviceSdf;
    uint64_t revision_ = 0;
    SceneBuilder data_;
    mutable std::shared_ptr<DeviceContext> ctx_; // geometry-only context used by traceRay()
    mutable std::mutex mu_;
    DeviceContext& context() const;
};

// ---- Lighting: CollectionLighting's API (src/CollectionLighting.h:10-21) ------------------------------
class DeviceLighting : public Lighting, public DeviceExportable {
public:
    void addPointLight(glm::vec3 position, float virtual_radius, float power = 1.0f);
    void addSphereLight(glm::vec3 position, float radius, float power = 1.0f);
    void addSquareLight(glm::vec3 corner, glm::vec3 normal, glm::vec3 x_side, float power = 1.0f);
    void addTriangleLight(glm::vec3 corner, glm::vec3 x_side, glm::vec3 y_side, float power = 1.0f);
    void addOuterLight(float radius, float power = 1.0f);
    void addLight(const ipt_light& l) { lights_.push_back(l); ctx_.reset(); ++revision_; }
    uint64_t revision() const override { return revision_; }
    size_t size() const { return lights_.size(); }

    std::unique_ptr<Ddf> distributionInPoint(glm::vec3 pos) const override;
    std::optional<light_intersection> traceRayToLight(glm::vec3 origin, glm::vec3 direction) const override;
    void exportTo(SceneBuilder& b) const override;

private:
    friend class DeviceLightDdf;
    uint64_t revision_ = 0;
    std::vector<ipt_light> lights_;
    mutable std::shared_ptr<DeviceContext> ctx_;
    mutable std::mutex mu_;
    DeviceContext& context() const;
};

// ---- Camera: SimpleCamera (src/SimpleCamera.h:10-17) ---------------------------------------------------
class DeviceCamera : public Camera, public DeviceExportable {
public:
    glm::vec3 position, direction, right, up; // the same public fields
    DeviceCamera(glm::vec3 position, glm::vec3 direction, glm::vec3 up_hint = glm::vec3(0, 0, 1));
    std::pair<glm::vec3, glm::vec3> sampleRay(float x, float y) const override;
    void exportTo(SceneBuilder& b) const override;
    // the arrow keys of Gui::work (gui.cpp:105-134): IPT_KEY_LEFT / RIGHT orbit about z, IPT_KEY_DOWN / UP dolly by 1.1
    void orbit(int key);

private:
    mutable std::shared_ptr<DeviceContext> ctx_;
};

// ---- RenderPlane: GridRenderPlane (src/GridRenderPlane.h:8-19) on the device --------------------------------
class DevicePlane : public RenderPlane {
public:
    std::vector<float> pixels;           // filled by download(): running mean per cell
    std::vector<size_t> pixel_counters;
    size_t width, height;
    float max_value = 0;
    uint32_t plane_mode = IPT_PLANE_GRID;

    DevicePlane(size_t width, size_t height);
    ~DevicePlane();
    void addRay(float x, float y, float value) override; // one sample -> the device accumulators
    void download();                                     // refresh pixels / pixel_counters / max_value
    void sums(std::vector<float>& sum, std::vector<float>& sumsq, std::vector<uint32_t>& count);
    // Gui's output stage on the device (gui.cpp): finalize/updateDisplay's image = normalize(glare(image, cutoff)), and
    // save(path) = normalize(image).normalize(0,255) as an 8-bit PNG (Gui::save, gui.cpp:192-194)
    std::vector<float> display(float glare_cutoff = 1.01f);
    void save(const char* path);

    void clear();                                        // Gui::resetImage (gui.cpp:152-160): accumulators back to zero
    void upload(const std::vector<float>& sum, const std::vector<float>& sumsq, const std::vector<uint32_t>& count); // resume

    // used by render_sample: (re)attach to the device context of the scene being rendered
    ipt_plane* attach(const std::shared_ptr<DeviceContext>& ctx);

private:
    std::shared_ptr<DeviceContext> ctx_;
    ipt_plane* plane_ = nullptr;
};

// ---- Scene factory: sample_scenes.h:6-11 + the benchmark scenes -------------------------------------------
Scene make_scene(const char* name); // "box", "fractal", "smallpt", "square", "corner", "openspheres", "cornell", "mesh:<n>", "lightgrid:<r>x<c>"
inline Scene make_scene_box() { return make_scene("box"); }
inline Scene make_scene_fractal() { return make_scene("fractal"); }
inline Scene make_scene_smallpt() { return make_scene("smallpt"); }
inline Scene make_scene_square_lit_by_square() { return make_scene("square"); }
inline Scene make_scene_lit_corner() { return make_scene("corner"); }

// ---- the hot path ------------------------------------------------------------------------------------------
// params.pass_count calls of the reference's render_sample(scene, r_plane, stats) (src/main.cpp:186-223).
//
// Scene members are discovered by dynamic_cast, the reference's own idiom (gui.cpp:58, ddf.cpp:177):
//   * anything DeviceExportable (the classes above);
//   * built against the REFERENCE's headers (IPT_B200_REFERENCE_CLASSES: src/ is on the include path), also the reference's
//     own objects where they expose what the device needs: SimpleCamera (public position / direction / right / up,
//     SimpleCamera.h:11-13) and the six data-free Geometry classes, whose primitive lists are literals inside traceRay()
//     (GeometrySphereInBox.cpp:11-17, ...) and are restated in ipt_sample_scene. Lighting must be a DeviceLighting:
//     AreaLight keeps its axes private (lighting.h:20-23).
// Anything else throws -- it never falls back to tracing on the CPU.
//
// The plane: a DevicePlane receives the accumulators directly. A reference GridRenderPlane (reference headers) is written
// the way addRay would have left it: pixels[i] = running mean, pixel_counters[i] += samples, max_value, cells by
// GridRenderPlane::addRay's own mapping (GridRenderPlane.cpp:61-75). Any other RenderPlane receives ONE
// addRay(x_centre, y_centre, mean) per cell (documented approximation: `count` samples collapse into one call).
ipt_render_stats render_sample(const Scene& scene, RenderPlane& r_plane, const ipt_render_params& params, int device = 0);
// The same job on several GPUs of this process: one host thread per device (the shape of the reference's own driver,
// main.cpp:258-277: N threads, one plane), each rendering a contiguous share of params.pass_count passes of the whole
// frame on its own replica of the scene; the per-device accumulators are merged into the plane of devices[0] with
// ipt_plane_merge (peer copy + add on the device). The result equals the single-device render of the same passes.
ipt_render_stats render_sample(const Scene& scene, RenderPlane& r_plane, const ipt_render_params& params, const std::vector<int>& devices);


// ---- the caller side of the interactive loop, headless -----------------------------------------
// src/main.cpp:258-269   the render threads call render_sample(scene, *gui, stats) forever: one more sample per pixel per call
// src/gui.cpp:105-137    arrow keys orbit / dolly the camera, then resetImage() + updateDisplay()
// src/gui.cpp:141-145    the mouse wheel scales glare_cutoff by sqrt(2) per click
// src/gui.cpp:83-87      updateDisplay(): normalize(glare(image, glare_cutoff));   :192-194 save()
// plus what the reference lacks: accumulator checkpoints. The Philox counters carry the pass index, so a session resumed
// from a checkpoint continues the very image the uninterrupted session would have produced.
class ProgressiveSession {
public:
    ipt_render_params params;   // depth_max, split schedule, seed, plane mode: set before the first step()
    float glare_cutoff = 1.01f; // gui.h:24
    ProgressiveSession(Scene scene, size_t width, size_t height, uint32_t passes_per_call = 4, std::vector<int> devices = {0});
    uint64_t step(unsigned calls = 1); // `calls` x passes_per_call more samples per pixel; returns samples per pixel of the image
    void key(int key);                 // IPT_KEY_*: the camera moves, the image restarts (the scene's camera must be a DeviceCamera)
    void resetImage();
    void wheel(int clicks);
    std::vector<float> display() { return plane_.display(glare_cutoff); }
    void save(const char* path) { plane_.save(path); }
    void checkpoint(const std::string& path); // atomic: written next to `path`, then renamed over it
    bool resume(const std::string& path);     // false: no such file. Throws when the file belongs to another estimator / frame
    uint64_t samples_per_pixel() const { return next_pass_ - first_pass_; }
    uint64_t rays() const { return rays_; }
    DevicePlane& plane() { return plane_; }

private:
    Scene scene_;
    DevicePlane plane_;
    std::vector<int> devices_;
    uint32_t passes_per_call_;
    uint64_t next_pass_ = 0, first_pass_ = 0, rays_ = 0;
};

} // namespace ipt_b200
