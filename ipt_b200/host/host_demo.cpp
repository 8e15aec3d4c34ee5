// host_demo.cpp — exercises the reference-facing C++ classes (device_plugins.hpp) end to end:
//   scene = ipt_b200::make_scene_box();  ipt_b200::render_sample(scene, plane, params);
// into (a) a DevicePlane and (b) a foreign RenderPlane written like the reference's GridRenderPlane, plus single-ray
// calls through the Geometry / Lighting / Camera virtuals. Prints one JSON object; tests/test_gpu_host_cpp.py checks it.
#include "device_plugins.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>

// a RenderPlane the library knows nothing about: GridRenderPlane::addRay's arithmetic (src/GridRenderPlane.cpp:61-75)
struct ForeignGridPlane : public RenderPlane {
    std::vector<float> pixels;
    std::vector<size_t> counters;
    size_t width, height;
    ForeignGridPlane(size_t w, size_t h) : pixels(w * h), counters(w * h), width(w), height(h) {}
    void addRay(float x, float y, float value) override {
        size_t xi = x * width;
        size_t yi = height - y * height - 1;
        pixels[yi * width + xi] = (pixels[yi * width + xi] * counters[yi * width + xi] + value) / (counters[yi * width + xi] + 1);
        ++counters[yi * width + xi];
    }
};

int main(int argc, char** argv) {
    const char* name = argc > 1 ? argv[1] : "box";
    int passes = argc > 2 ? atoi(argv[2]) : 4;
    try {
        Scene scene = ipt_b200::make_scene(name);
        ipt_render_params p;
        ipt_render_params_default(&p);
        p.width = p.height = 96;
        p.pass_count = passes;
        ipt_b200::DevicePlane dplane(96, 96);
        ipt_render_stats st = ipt_b200::render_sample(scene, dplane, p);
        dplane.download();
        double mean_d = 0;
        size_t cnt_d = 0;
        for (size_t i = 0; i < dplane.pixels.size(); ++i) { mean_d += (double)dplane.pixels[i] * dplane.pixel_counters[i]; cnt_d += dplane.pixel_counters[i]; }
        ForeignGridPlane fplane(96, 96);
        ipt_b200::render_sample(scene, fplane, p);
        double mean_f = 0;
        size_t cells_f = 0;
        for (size_t i = 0; i < fplane.pixels.size(); ++i) if (fplane.counters[i]) { mean_f += fplane.pixels[i]; ++cells_f; }
        // the virtuals, one ray at a time
        auto ray = scene.camera->sampleRay(0.5f, 0.3f);
        auto si = scene.geometry->traceRay(ray.first, ray.second);
        auto li = scene.lighting->traceRayToLight(glm::vec3(0.2f, -0.8f, -1.0f), glm::vec3(0, 0, 1));
        float sdf_up = si ? si->sdf->value(si->normal) : -1.0f;
        glm::vec3 smp = si ? si->sdf->sample() : glm::vec3();
        auto lddf = scene.lighting->distributionInPoint(glm::vec3(0.2f, -0.8f, -1.0f));
        float lval = lddf->value(glm::vec3(0, 0, 1));
        // (rendered before the camera keys below move the camera: an orbit round trip is exact only to rounding)
        // the multi-GPU entry: the same passes split over two device slots (two GPUs when the box has them), merged on the
        // device by ipt_plane_merge, must reproduce the single-device plane
        ipt_b200::DevicePlane mplane(96, 96);
        std::vector<int> slots = {0, ipt_device_count() > 1 ? 1 : 0};
        ipt_render_stats mst = ipt_b200::render_sample(scene, mplane, p, slots);
        mplane.download();
        std::vector<float> s1, q1, s2, q2;
        std::vector<uint32_t> c1, c2;
        dplane.sums(s1, q1, c1);
        mplane.sums(s2, q2, c2);
        size_t multi_counters_equal = 0;
        double multi_max_rel = 0;
        for (size_t i = 0; i < c1.size(); ++i) {
            multi_counters_equal += c1[i] == c2[i];
            double scale = std::fabs(s1[i]) > 1e-6 ? std::fabs(s1[i]) : 1e-6;
            double rel = std::fabs((double)s1[i] - s2[i]) / scale;
            if (rel > multi_max_rel) multi_max_rel = rel;
        }
        printf("{\"multi_devices\": [%d, %d], \"multi_paths\": %llu, \"multi_rays\": %llu, \"multi_counters_equal\": %zu, \"multi_max_rel\": %.3g}\n", slots[0],
               slots[1], (unsigned long long)mst.paths, (unsigned long long)mst.rays, multi_counters_equal, multi_max_rel);
        // Gui's output stage (gui.cpp:186-194) and camera keys (gui.cpp:105-134) through the host classes
        std::vector<float> shown = dplane.display(0.2f);
        float shown_max = 0;
        for (float v : shown) shown_max = v > shown_max ? v : shown_max;
        if (argc > 3) dplane.save(argv[3]);
        auto* cam = dynamic_cast<ipt_b200::DeviceCamera*>(const_cast<Camera*>(scene.camera.get()));
        glm::vec3 before = cam->position;
        cam->orbit(IPT_KEY_LEFT);
        cam->orbit(IPT_KEY_RIGHT);
        float orbit_err = std::fabs(cam->position.x - before.x) + std::fabs(cam->position.y - before.y) + std::fabs(cam->position.z - before.z);
        dplane.addRay(0.5f, 0.5f, 1.0f);
        // the interactive loop, headless, with a checkpoint in the middle: A renders 2 + 2 calls, B resumes A's checkpoint after
        // the first 2 and renders the other 2 — the same image (counters equal, sums equal up to the order of the float atomics);
        // then a key restarts the image and the wheel scales the glare cutoff like Gui::work does
        size_t session_counters_equal = 0, session_spp_a = 0, session_spp_b = 0, session_spp_after_key = 0;
        double session_max_rel = 0;
        float session_cutoff = 0;
        int session_refused = 0;
        if (argc > 3) {
            const std::string ckpt = std::string(argv[3]) + ".ckpt";
            ipt_b200::ProgressiveSession a(ipt_b200::make_scene(name), 64, 64, 2);
            a.step(2);
            a.checkpoint(ckpt);
            session_spp_a = a.step(2);
            ipt_b200::ProgressiveSession b(ipt_b200::make_scene(name), 64, 64, 2);
            if (!b.resume(ckpt)) throw ipt_b200::Error(IPT_ERR_INVALID, "checkpoint not found");
            session_spp_b = b.step(2);
            std::vector<float> sa, qa, sb, qb;
            std::vector<uint32_t> ca, cb;
            a.plane().sums(sa, qa, ca);
            b.plane().sums(sb, qb, cb);
            for (size_t i = 0; i < ca.size(); ++i) {
                session_counters_equal += ca[i] == cb[i];
                double scale = std::fabs(sa[i]) > 1e-6 ? std::fabs(sa[i]) : 1e-6;
                double rel = std::fabs((double)sa[i] - sb[i]) / scale;
                if (rel > session_max_rel) session_max_rel = rel;
            }
            ipt_b200::ProgressiveSession c(ipt_b200::make_scene(name), 64, 64, 3); // another step size: not the same estimator run
            try { c.resume(ckpt); } catch (const ipt_b200::Error&) { session_refused = 1; }
            b.key(IPT_KEY_LEFT);
            session_spp_after_key = b.step(1);
            b.wheel(2);
            session_cutoff = b.glare_cutoff;
            std::remove(ckpt.c_str());
        }
        printf("{\"session_counters_equal\": %zu, \"session_max_rel\": %.3g, \"session_spp\": [%zu, %zu, %zu], \"session_cutoff\": %.9g, \"session_refused\": %d}\n",
               session_counters_equal, session_max_rel, session_spp_a, session_spp_b, session_spp_after_key, session_cutoff, session_refused);
        printf("{\"display_max\": %.9g, \"orbit_round_trip_error\": %.9g}\n", shown_max, orbit_err);
        printf("{\"scene\": \"%s\", \"paths\": %llu, \"rays\": %llu, \"mean_device_plane\": %.9g, \"count_device_plane\": %zu, "
               "\"mean_foreign_plane\": %.9g, \"cells_foreign_plane\": %zu, \"hit\": %d, \"hit_pos\": [%.9g, %.9g, %.9g], "
               "\"hit_normal\": [%.9g, %.9g, %.9g], \"sdf_value_at_normal\": %.9g, \"sdf_sample_dot_normal\": %.9g, \"light_hit\": %d, "
               "\"light_power\": %.9g, \"light_ddf_value\": %.9g}\n",
               name, (unsigned long long)st.paths, (unsigned long long)st.rays, mean_d / (cnt_d ? cnt_d : 1), cnt_d, mean_f / (cells_f ? cells_f : 1),
               cells_f, si ? 1 : 0, si ? si->position.x : 0, si ? si->position.y : 0, si ? si->position.z : 0, si ? si->normal.x : 0,
               si ? si->normal.y : 0, si ? si->normal.z : 0, sdf_up, si ? smp.x * si->normal.x + smp.y * si->normal.y + smp.z * si->normal.z : 0,
               li ? 1 : 0, li ? li->surface_power : 0, lval);
    } catch (const ipt_b200::Error& e) {
        printf("{\"error\": \"%s\", \"code\": %d}\n", e.what(), e.code);
        return e.code == IPT_ERR_NO_DEVICE ? 3 : 1;
    }
    return 0;
}
