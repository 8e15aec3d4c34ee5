// oracle/ref_gui_driver.cpp — exposes the reference's OUTPUT STAGE (the file-static functions of src/gui.cpp) through a
// C ABI. TEST INFRASTRUCTURE ONLY. The reference source is compiled where it lies (`#include "gui.cpp"`, found through
// -I$(REF)/src) — this TU replaces gui.cpp in the link of oracle/_ref/libipt_ref.so, nothing of it is copied here.
// Gui itself cannot be constructed headless (its CImgDisplay member throws with cimg_display=0), so the statics
// normalize(), glare() and the expression of Gui::save are called directly.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <unistd.h>

#include "gui.cpp" // the unmodified reference: normalize() :11-16, draw_halo() :28-36, glare() :38-52, Gui::* :54-194

namespace {
CImg<float> wrap(const float* p, uint32_t w, uint32_t h) { return CImg<float>(p, w, h, 1, 1, false); } // copies
}

extern "C" {

int iptref_image_normalize(const float* image, uint32_t w, uint32_t h, float* out) {
    CImg<float> r = normalize(wrap(image, w, h));
    std::memcpy(out, r.data(), sizeof(float) * (size_t)w * h);
    return 0;
}

int iptref_image_glare(const float* image, uint32_t w, uint32_t h, float cutoff, float* out) {
    CImg<float> r = glare(wrap(image, w, h), cutoff);
    std::memcpy(out, r.data(), sizeof(float) * (size_t)w * h);
    return 0;
}

// Gui::save (gui.cpp:192-194) up to the file format: CImg built without libpng hands the image to an external
// converter as an 8-bit PGM written by save_pnm (CImg.h:60650-60667) — that PGM's payload is what ends up in result.png.
int iptref_image_save_bytes(const float* image, uint32_t w, uint32_t h, uint8_t* out) {
    char path[64];
    std::snprintf(path, sizeof path, "/tmp/iptref_%d.pgm", (int)getpid());
    try {
        normalize(wrap(image, w, h)).normalize(0, 255).save_pnm(path);
    } catch (...) { return 1; }
    FILE* f = std::fopen(path, "rb");
    if (!f) return 2;
    int pw = 0, ph = 0, maxv = 0;
    char magic[3] = {0, 0, 0};
    int ok = std::fscanf(f, "%2s %d %d %d", magic, &pw, &ph, &maxv) == 4 && std::fgetc(f) != EOF;
    ok = ok && magic[0] == 'P' && magic[1] == '5' && pw == (int)w && ph == (int)h && maxv == 255;
    ok = ok && std::fread(out, 1, (size_t)w * h, f) == (size_t)w * h;
    std::fclose(f);
    std::remove(path);
    return ok ? 0 : 3;
}

// The arrow-key branch of Gui::work (gui.cpp:105-134) cannot be reached without a display; the same glm expressions are
// evaluated here with the reference's vendored glm. key: 0 left, 1 right, 2 down, 3 up.
int iptref_camera_orbit(float* position, float* direction, float* right, float* up, int key) {
    SimpleCamera camera(vec3(position[0], position[1], position[2]), vec3(direction[0], direction[1], direction[2]));
    if (key == 0) {
        mat3 mat = glm::rotate(glm::identity<mat4>(), (float)-M_PI/12, vec3(0,0,1));
        camera.position = mat * camera.position;
        camera.direction = mat*camera.direction;
    } else if (key == 1) {
        mat3 mat = glm::rotate(glm::identity<mat4>(), (float)+M_PI/12, vec3(0,0,1));
        camera.position = mat * camera.position;
        camera.direction = mat*camera.direction;
    } else if (key == 2) camera.position *= 1.1;
    else if (key == 3) camera.position /= 1.1;
    else return 1;
    camera.right = glm::normalize(glm::cross(camera.direction, vec3(0,0,1)));
    camera.up = glm::normalize(glm::cross(camera.right, camera.direction));
    for (int i = 0; i < 3; ++i) {
        position[i] = camera.position[i]; direction[i] = camera.direction[i];
        right[i] = camera.right[i]; up[i] = camera.up[i];
    }
    return 0;
}

} // extern "C"
