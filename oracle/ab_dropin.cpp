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
// Prints one JSON object; tests/test_host_cpp.py checks it on the GPU box.
#include "device_plugins.hpp"

#include <SimpleCamera.h>

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
        printf("{\"scene\": \"%s\", \"rays\": %d, \"geometry_equal\": %zu, \"geometry_hits\": %zu, \"light_equal\": %zu, \"light_hits\": %zu, "
               "\"mean_reference_objects\": %.9g, \"mean_ipt_b200_objects\": %.9g, \"standard_error\": %.9g}\n",
               name, rays, geometry_equal, hits, light_equal, lhits, m_ref, m_ours, se);
    } catch (const ipt_b200::Error& e) {
        printf("{\"error\": \"%s\", \"code\": %d}\n", e.what(), e.code);
        return e.code == IPT_ERR_NO_DEVICE ? 3 : 1;
    }
    return 0;
}
