// oracle/ab_dropin.cpp — runtime drop-in check (TEST INFRASTRUCTURE; built into oracle/_ref/ab_dropin, which contains
// reference code and is therefore git-ignored like libipt_ref.so).
//
// This program is compiled against the REFERENCE's own headers (src/tracer_interfaces.h, glm, libddf/ddf.h). It
//   1. builds the default scene twice: the reference's objects (make_scene_box inside libipt_ref.so) and ipt_b200's
//      host classes (ipt_b200::make_scene_box, device_plugins.cpp compiled HERE against the reference headers);
//   2. hands BOTH object pairs to the reference's own compiled estimator ray_power_recursive (main.cpp:98-184, through
//      iptref_ray_power_on): virtual calls from reference code land in ipt_b200's classes, which evaluate every ray
//      and every DDF sample / value on the GPU through the C ABI;
//   3. compares, ray by ray, Geometry::traceRay / Lighting::traceRayToLight of the two object sets (must be equal to
//      the bit) and the mean estimate over the same camera rays (statistical: the DDF samples come from different
//      random generators).
//   4. renders through ipt_b200::render_sample a Scene made of the REFERENCE's own objects where they expose their data —
//      the reference's GeometrySphereInBox and SimpleCamera (constructed here, as sample_scenes.cpp:20-41 does) with a
//      DeviceLighting carrying the same light — into a reference GridRenderPlane, and compares that plane (pixels,
//      pixel_counters, max_value) with a DevicePlane rendered from ipt_b200's own classes: same cells, same counters, same
//      means; then the same job split over two device slots (the multi-GPU entry) must reproduce it.
// Prints one JSON object; tests/test_host_cpp.py checks it on the GPU box.
#include "device_plugins.hpp"

#include <GridRenderPlane.h>
#include <SimpleCamera.h>
#include <geometry/GeometrySphereInBox.h>

#include <glm/geometric.hpp>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

extern "C" {
int iptref_scene_create(const char* name);
void iptref_set_tree(int n, int dmax);
void iptref_seed(long seed);
float iptref_ray_power_on(const void* geometry, const void* lighting, const float* o, const float* d, int depth, int n);
const void* iptref_scene_geometry(int scene);
const void* iptref_scene_lighting(int scene);
}

int main(int argc, char** argv) {
    const char* name = argc > 1 ? argv[1] : "box";
    int rays = argc > 2 ? atoi(argv[2]) : 1500;
    try {
        int h = iptref_scene_create(name);
        const Geometry* ref_geometry = static_cast<const Geometry*>(iptref_scene_geometry(h));
        const Lighting* ref_lighting = static_cast<const Lighting*>(iptref_scene_lighting(h));
        Scene ours = ipt_b200::make_scene(name);
        iptref_set_tree(2, 3); // n_rays = 2, depth_max = 3: 1 + 2 + 2 rays per path
        srand48(5);
        std::vector<float> ox, dx;
        size_t geometry_equal = 0, light_equal = 0, hits = 0, lhits = 0;
        for (int i = 0; i < rays; ++i) {
            float x = drand48(), y = drand48();
            auto ray = ours.camera->sampleRay(x, y); // DeviceCamera -> GPU
            auto a = ref_geometry->traceRay(ray.first, ray.second);
            auto b = ours.geometry->traceRay(ray.first, ray.second); // DeviceGeometry -> GPU
            bool same = a.has_value() == b.has_value();
            if (same && a) same = std::memcmp(&a->position, &b->position, sizeof(glm::vec3)) == 0 && a->normal == b->normal && a->curvature == b->curvature;
            geometry_equal += same;
            hits += a.has_value();
            // a second ray from the hit point towards the light region exercises traceRayToLight
            glm::vec3 o2 = a ? a->position : ray.first;
            glm::vec3 d2 = glm::vec3(0.2f - o2.x + 0.1f * (float)drand48(), -0.8f - o2.y + 0.1f * (float)drand48(), -0.15f - o2.z);
            float len = std::sqrt(d2.x * d2.x + d2.y * d2.y + d2.z * d2.z);
            d2 = glm::vec3(d2.x / len, d2.y / len, d2.z / len);
            auto la = ref_lighting->traceRayToLight(o2, d2);
            auto lb = ours.lighting->traceRayToLight(o2, d2); // DeviceLighting -> GPU
            bool lsame = la.has_value() == lb.has_value();
            if (lsame && la) lsame = std::memcmp(&la->position, &lb->position, sizeof(glm::vec3)) == 0 && la->surface_power == lb->surface_power;
            light_equal += lsame;
            lhits += la.has_value();
            ox.insert(ox.end(), {ray.first.x, ray.first.y, ray.first.z});
            dx.insert(dx.end(), {ray.second.x, ray.second.y, ray.second.z});
        }
        // the reference's estimator on both object sets, same camera rays
        double sum_ref = 0, sum_ours = 0, sq_ref = 0, sq_ours = 0;
        iptref_seed(77);
        for (int i = 0; i < rays; ++i) {
            float v = iptref_ray_power_on(ref_geometry, ref_lighting, &ox[3 * i], &dx[3 * i], 0, 2);
            sum_ref += v; sq_ref += (double)v * v;
        }
        for (int i = 0; i < rays; ++i) {
            float v = iptref_ray_power_on(ours.geometry.get(), ours.lighting.get(), &ox[3 * i], &dx[3 * i], 0, 2);
            sum_ours += v; sq_ours += (double)v * v;
        }
        double m_ref = sum_ref / rays, m_ours = sum_ours / rays;
        double se = std::sqrt((sq_ref / rays - m_ref * m_ref + sq_ours / rays - m_ours * m_ours) / rays);
        // ---- 4. the reference's own objects through ipt_b200::render_sample -------------------------------------------
        size_t cells_equal = 0, cells = 0, counters_equal = 0;
        double max_rel = 0, max_rel_multi = 0;
        float max_value_grid = 0, max_value_device = 0;
        int devices_seen = ipt_device_count();
        if (std::strcmp(name, "box") == 0) {
            // make_scene_box (sample_scenes.cpp:20-41) with the reference's geometry and camera classes
            glm::vec3 camera_pos(0.0f, -3.0f, 0.1f);
            glm::vec3 camera_dir = glm::normalize(glm::vec3(0.0f, 1.0f, -1.0f) - camera_pos);
            auto lighting = std::make_shared<ipt_b200::DeviceLighting>();
            lighting->addSquareLight(glm::vec3{+0.1f, -0.8f - 0.1f, -0.15f}, glm::vec3(0.0f, 0.0f, -1.0f), glm::vec3{0.0f, 0.2f, 0.0f}, 1.0f);
            Scene mixed{std::make_shared<GeometrySphereInBox>(), lighting, std::make_shared<SimpleCamera>(camera_pos, camera_dir)};
            ipt_render_params p;
            ipt_render_params_default(&p);
            p.width = p.height = 80;
            p.pass_count = 6;
            p.seed = 17;
            GridRenderPlane grid(80, 80);
            ipt_b200::render_sample(mixed, grid, p);
            ipt_b200::DevicePlane dplane(80, 80);
            ipt_b200::render_sample(ours, dplane, p);
            dplane.download();
            GridRenderPlane grid2(80, 80); // the same passes on two device slots, merged by ipt_plane_merge
            std::vector<int> two = {0, devices_seen > 1 ? 1 : 0};
            ipt_b200::render_sample(mixed, grid2, p, two);
            for (size_t i = 0; i < grid.pixels.size(); ++i) {
                counters_equal += grid.pixel_counters[i] == dplane.pixel_counters[i] && grid2.pixel_counters[i] == dplane.pixel_counters[i];
                if (!dplane.pixel_counters[i]) continue;
                ++cells;
                double scale = std::fabs(dplane.pixels[i]) > 1e-6 ? std::fabs(dplane.pixels[i]) : 1e-6;
                double rel = std::fabs((double)grid.pixels[i] - dplane.pixels[i]) / scale;
                double rel2 = std::fabs((double)grid2.pixels[i] - dplane.pixels[i]) / scale;
                cells_equal += rel <= 1e-5;
                if (rel > max_rel) max_rel = rel;
                if (rel2 > max_rel_multi) max_rel_multi = rel2;
            }
            max_value_grid = grid.max_value;
            max_value_device = dplane.max_value;
        }
        printf("{\"scene\": \"%s\", \"rays\": %d, \"geometry_equal\": %zu, \"geometry_hits\": %zu, \"light_equal\": %zu, \"light_hits\": %zu, "
               "\"mean_reference_objects\": %.9g, \"mean_ipt_b200_objects\": %.9g, \"standard_error\": %.9g, "
               "\"grid_cells\": %zu, \"grid_cells_equal\": %zu, \"grid_counters_equal\": %zu, \"grid_max_rel\": %.3g, \"grid_max_rel_two_devices\": %.3g, "
               "\"grid_max_value\": %.9g, \"device_max_value\": %.9g, \"devices\": %d}\n",
               name, rays, geometry_equal, hits, light_equal, lhits, m_ref, m_ours, se, cells, cells_equal, counters_equal, max_rel, max_rel_multi,
               max_value_grid, max_value_device, devices_seen);
    } catch (const ipt_b200::Error& e) {
        printf("{\"error\": \"%s\", \"code\": %d}\n", e.what(), e.code);
        return e.code == IPT_ERR_NO_DEVICE ? 3 : 1;
    }
    return 0;
}
