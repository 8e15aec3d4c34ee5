// ipt_shading.cuh — DDF sampling / evaluation on the device (libddf + lighting DDFs of the reference).
//
// Everything downstream of a random number can only match the reference statistically (drand48 vs Philox,
// SURVEY.md S6), so this file is free to use algebraically equivalent, cheaper forms:
//   * RotateDdf's matrix (glm::rotate by acos(n.z) about z x n, ddf_detail.h:72-85) is built from
//     c = n.z, s = |z x n| without any trigonometry; it is the same rotation.
//   * TransformDdf::value(w) = origin.value(R^-1 w) only needs (R^-1 w).z = dot(R*z, w) = dot(axis_to, w).
//   * sin(acos(x)) = sqrt(1 - x*x).
// The intersection calls inside the light pdf stay on the exact routines of ipt_device.cuh.
#pragma once
#include "ipt_device.cuh"

namespace iptd {

#define IPT_PI_F 3.14159265358979323846f

// The sampling code spells out its multiply-adds (__fmaf_rn / __fmul_rn / ...): left to the compiler, the contraction
// of a*b + c into an FMA depends on the surrounding code, so the SAME source gave directions differing in the last bit
// between the instantiations of k_shade — harmless statistically, but then "fused" and "queued" runs could not be
// compared ray for ray.
__device__ __forceinline__ float pmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float pfma(float a, float b, float c) { return __fmaf_rn(a, b, c); }

struct Basis { // columns of RotateDdf::transformation; c2 == the axis rotated to
    f3 c0, c1, c2;
};

// RotateDdf::RotateDdf (src/libddf/ddf_detail.h:72-85)
__device__ __forceinline__ Basis make_basis(f3 to) {
    float c = to.z;
    float axx = -to.y, axy = to.x; // cross((0,0,1), to) = (-to.y, to.x, 0)
    float len2 = pfma(axy, axy, pmul(axx, axx));
    float s = fsqrt(len2);
    float ax, ay;
    if (s < 1e-6f) { ax = 1.0f; ay = 0.0f; } // ddf_detail.h:77-78 degenerate axis -> (1,0,0)
    else { float inv = frcp(s); ax = pmul(axx, inv); ay = pmul(axy, inv); }
    // with the degenerate axis the angle is still acos(c): c = +-1, sin = sqrt(1-c*c)
    float sn = fsqrt(fmaxf(0.0f, pfma(-c, c, 1.0f)));
    float omc = __fsub_rn(1.0f, c);
    float tx = pmul(omc, ax), ty = pmul(omc, ay);
    Basis b;
    b.c0 = mk3(pfma(tx, ax, c), pmul(tx, ay), pmul(-sn, ay));
    b.c1 = mk3(pmul(ty, ax), pfma(ty, ay, c), pmul(sn, ax));
    b.c2 = mk3(pmul(sn, ay), pmul(-sn, ax), c);
    return b;
}
__device__ __forceinline__ float dot3(f3 a, f3 b) { return pfma(a.z, b.z, pfma(a.y, b.y, pmul(a.x, b.x))); }
__device__ __forceinline__ f3 rotate(const Basis& b, f3 x) {
    return mk3(pfma(b.c2.x, x.z, pfma(b.c1.x, x.y, pmul(b.c0.x, x.x))), pfma(b.c2.y, x.z, pfma(b.c1.y, x.y, pmul(b.c0.y, x.x))),
               pfma(b.c2.z, x.z, pfma(b.c1.z, x.y, pmul(b.c0.z, x.x))));
}

// ---- reference-order sampling (EXACT = true; the SmallPt scene) ---------------------------------------------------------
// GeometrySmallPt intersects radius-1000 spheres with the unit-direction formula and eps = 1e-4 (GeometrySmallPt.cpp:17-22):
// a direction whose length is off by 1e-7 moves the computed hit point 1e-5 off the sphere, which decides whether the NEXT
// ray re-hits the wall it starts on (11 % of the rays of that scene do). The image therefore depends on the rounding errors
// of the sampled directions' lengths — scaling every sampled direction by (1 - 1e-7) darkens it by 0.7 % — and those
// errors are not zero-mean per hit: the columns of RotateDdf's float matrix are off by up to 1e-7 for a given normal. For
// that scene the sampling code below follows the reference's operation sequence (glm::rotate, acos / sin / cos of the
// angles, glm's mat3 * vec3, normalize) with IEEE operations; sin / cos / acos are evaluated in double and rounded, which
// agrees with glibc's float functions (< 0.56 ulp) in all but a few per cent of the arguments, and then by one ulp.
__device__ __forceinline__ float sin_r(float a) { return __double2float_rn(sin((double)a)); }
__device__ __forceinline__ float cos_r(float a) { return __double2float_rn(cos((double)a)); }
__device__ __forceinline__ float acos_r(float a) { return __double2float_rn(acos((double)a)); }
// RotateDdf::RotateDdf (src/libddf/ddf_detail.h:72-85) over glm::rotate(identity, angle, axis) (glm/ext/matrix_transform.inl:18-46)
__device__ __forceinline__ Basis make_basis_exact(f3 to) {
    f3 z = mk3(0.0f, 0.0f, 1.0f);
    f3 axis = xcross3(z, to);
    if (lt_1e6(xlength3(axis))) axis = mk3(1.0f, 0.0f, 0.0f); // length(axis) < 1e-6 (double literal)
    float cosinus = xdot3(z, to);
    float a = acos_r(cosinus); // ddf_detail.h:76 calls the double acos
    float c = cos_r(a), s = sin_r(a);
    f3 ax = xnormalize3(axis);
    f3 temp = xscale3(ax, xsub(1.0f, c));
    Basis b;
    b.c0 = mk3(xadd(c, xmul(temp.x, ax.x)), xadd(xmul(temp.x, ax.y), xmul(s, ax.z)), xsub(xmul(temp.x, ax.z), xmul(s, ax.y)));
    b.c1 = mk3(xsub(xmul(temp.y, ax.x), xmul(s, ax.z)), xadd(c, xmul(temp.y, ax.y)), xadd(xmul(temp.y, ax.z), xmul(s, ax.x)));
    b.c2 = mk3(xadd(xmul(temp.z, ax.x), xmul(s, ax.y)), xsub(xmul(temp.z, ax.y), xmul(s, ax.x)), xadd(c, xmul(temp.z, ax.z)));
    return b;
}
// glm mat3 * vec3 (glm/detail/type_mat3x3.inl:468-474): (m[0][i] * v.x + m[1][i] * v.y) + m[2][i] * v.z
__device__ __forceinline__ f3 rotate_exact(const Basis& b, f3 x) {
    return mk3(xadd(xadd(xmul(b.c0.x, x.x), xmul(b.c1.x, x.y)), xmul(b.c2.x, x.z)), xadd(xadd(xmul(b.c0.y, x.x), xmul(b.c1.y, x.y)), xmul(b.c2.y, x.z)),
               xadd(xadd(xmul(b.c0.z, x.x), xmul(b.c1.z, x.y)), xmul(b.c2.z, x.z)));
}
// CosineDdf::sample / the PowerCosine extension in their own frame (src/libddf/ddf.cpp:91-102): cos(alpha) = u1^(1/(n+1)),
// alpha = acos(.), phi = 2 pi u2 (double product, rounded), (sin(alpha) cos(phi), sin(alpha) sin(phi), cos(alpha))
__device__ __forceinline__ f3 lobe_sample_exact(bool lobe, float inv_np1, float u1, float u2) {
    float zc = lobe ? powf(u1, inv_np1) : xsqrt(u1);
    float alpha = acos_r(zc);
    float phi = __double2float_rn(__dmul_rn(6.283185307179586, (double)u2));
    float r = sin_r(alpha);
    return mk3(xmul(r, cos_r(phi)), xmul(r, sin_r(phi)), zc);
}

// Base DDFs in their own frame. kind: 0 Spherical (ddf.cpp:58-72), 1 UpperHalf (:74-89), 2 Cosine (:91-108),
// >=3 PowerCosine(kind) (extension, oracle/ref_driver.cpp).
// z^n for a small non-negative integer n by square-and-multiply (the glossy lobe's exponent is an integer >= 3)
__device__ __forceinline__ float ipow(float z, int n) {
    float r = 1.0f, b = z;
    while (n) {
        if (n & 1) r *= b;
        b *= b;
        n >>= 1;
    }
    return r;
}
__device__ __forceinline__ f3 base_sample(int kind, float u1, float u2) {
    float zc;
    if (kind == 0) zc = u1 * 2.0f - 1.0f;
    else if (kind == 1) zc = u1;
    else if (kind == 2) zc = fsqrt(u1);
    else zc = exp2f(__log2f(u1) * (1.0f / ((float)kind + 1.0f))); // u1^(1/(n+1)); u1 = 0 -> 0
    float r = fsqrt(fmaxf(0.0f, 1.0f - zc * zc));
    // phi = 2*pi*u2 = pi*a + pi with a in [-1,1): cos(phi) = -cos(pi*a), sin(phi) = -sin(pi*a); the SFU sine/cosine are
    // accurate to ~4e-7 absolute on [-pi, pi], far below the sampling noise
    float a = (2.0f * u2 - 1.0f) * IPT_PI_F;
    float sp = -__sinf(a), cp = -__cosf(a);
    return mk3(r * cp, r * sp, zc);
}
__device__ __forceinline__ float base_value(int kind, float z) {
    if (kind == 0) return 0.25f / IPT_PI_F;
    if (z < 0.0f) return 0.0f;
    if (kind == 1) return 0.5f / IPT_PI_F;
    if (kind == 2) return z * (1.0f / IPT_PI_F);
    return ((float)kind + 1.0f) * ipow(z, kind) * (0.5f / IPT_PI_F);
}

#ifndef IPT_LOBE_LOG2
// log2 inside cos^n of the glossy lobe: the SFU's __log2f (abs. error 2^-21.4 on [0.5, 2]) puts at most 1e-5 of relative
// error on cos^40 at the lobe's peak — statistical domain (downstream of a random direction), and the per-pixel parity
// tests against the oracle's powf hold unchanged; log2f costs 4 % of the fused kernels' instructions (C2 292 vs 300).
#define IPT_LOBE_LOG2 __log2f
#endif
// The surface DDF of a hit: RotateDdf(CosineDdf, normal) or the glossy extension. Both are members of one family —
// kd * PowerCosine(1) about the normal + ks * PowerCosine(n) about the mirror direction, with kd = 1, ks = 0 for the
// Lambert case — and are evaluated by ONE branch-free code path: hits of both materials share warps at every depth
// after the first, so a per-material branch would issue both sides for nearly every warp.
struct Sdf {
    uint32_t ddf;
    f3 normal, refl;
    float wd, ws;
    float exponent;   // n of the lobe (1 for Lambert: unused, ks = 0)
    float inv_np1;    // 1 / (n + 1): the sampling exponent of the lobe
    float lobe_norm;  // (n + 1) / (2 pi)
};
__device__ __forceinline__ f3 reflect3(f3 I, f3 N) { // glm::reflect: I - N*dot(N,I)*2
    float k = pmul(2.0f, dot3(N, I));
    return mk3(pfma(-N.x, k, I.x), pfma(-N.y, k, I.y), pfma(-N.z, k, I.z));
}
__device__ __forceinline__ Sdf make_sdf(const DevMaterial& m, f3 normal, f3 dir_in) {
    Sdf s;
    bool glossy = m.ddf == IPT_DDF_GLOSSY;
    s.ddf = m.ddf;
    s.normal = normal;
    s.wd = glossy ? m.wd : 1.0f;
    s.ws = glossy ? m.ws : 0.0f;
    s.exponent = glossy ? (float)(int)m.exponent : 1.0f;
    s.inv_np1 = m.inv_np1;
    s.lobe_norm = m.lobe_norm;
    s.refl = glossy ? reflect3(dir_in, normal) : normal;
    return s;
}
// Lambert: z/pi (ddf.cpp:104-108); glossy: kd*z/pi + ks*(n+1)/(2 pi)*cos^n about the mirror direction, 0 below the surface
// LAMBERT: the caller knows at compile time that every material of the scene is the cosine DDF
template <bool LAMBERT = false>
__device__ __forceinline__ float sdf_value(const Sdf& s, f3 w) {
    float cn = dot3(s.normal, w);
    if (LAMBERT) return cn < 0.0f ? 0.0f : pmul(cn, 1.0f / IPT_PI_F); // == the general expression with kd = 1, ks = 0
    float zr = dot3(s.refl, w);
    float lobe = zr > 0.0f ? exp2f(pmul(s.exponent, IPT_LOBE_LOG2(zr))) : 0.0f;
    float v = pfma(s.ws, pmul(s.lobe_norm, lobe), pmul(s.wd, pmul(cn, 1.0f / IPT_PI_F)));
    return cn < 0.0f ? 0.0f : v;
}
// zero vector == failed sample. ul is the lobe-selection draw (ROLE_LOBE in the oracle).
// bl: basis about the mirror direction (== bn for Lambert hits; built once per hit, not per child)
template <bool LAMBERT = false, bool EXACT = false>
__device__ __forceinline__ f3 sdf_sample(const Sdf& s, const Basis& bn, const Basis& bl, float u1, float u2, float ul) {
    bool lobe = !LAMBERT && !(ul < s.wd);   // never for Lambert (wd = 1 > ul)
    if (EXACT) { // reference-order arithmetic (see make_basis_exact)
        f3 x = lobe_sample_exact(lobe, s.inv_np1, u1, u2);
        f3 w = lobe ? rotate_exact(bl, x) : rotate_exact(bn, x);
        bool below = !LAMBERT && s.ddf == IPT_DDF_GLOSSY && xdot3(s.normal, w) < 0.0f;
        return below ? mk3(0, 0, 0) : w;
    }
    float e = lobe ? s.inv_np1 : 0.5f;      // cos(alpha) = u1^(1/(n+1)); sqrt(u1) for the cosine DDF (ddf.cpp:94)
    float zc = exp2f(pmul(__log2f(u1), e)); // u1 = 0 -> 0
    float r = fsqrt(fmaxf(0.0f, pfma(-zc, zc, 1.0f)));
    float a = pmul(pfma(2.0f, u2, -1.0f), IPT_PI_F); // see base_sample
    float sp = -__sinf(a), cp = -__cosf(a);
    f3 x = mk3(pmul(r, cp), pmul(r, sp), zc);
    Basis b;
    b.c0 = lobe ? bl.c0 : bn.c0;
    b.c1 = lobe ? bl.c1 : bn.c1;
    b.c2 = lobe ? bl.c2 : bn.c2;
    f3 w = rotate(b, x);
    bool below = !LAMBERT && s.ddf == IPT_DDF_GLOSSY && dot3(s.normal, w) < 0.0f;
    return below ? mk3(0, 0, 0) : w;
}

// DdfFromLight::value (src/lighting/lighting.cpp:61-73)
__device__ __forceinline__ float light_pdf(const DevLight& L, f3 pos, f3 w) {
    LightHit h = light_trace(L, pos, w);
    if (!h.hit) return 0.0f;
    f3 dp = mk3(h.position.x - pos.x, h.position.y - pos.y, h.position.z - pos.z);
    float decay = dot3(dp, dp);
    float inv = rsqrtf(decay);
    float cosinus = pmul(-dot3(h.normal, dp), inv);
    if (cosinus < 0.0f) return 0.0f;
    return __fdividef(decay, pmul(cosinus, L.area));
}

// The same density from an already known light hit (its position on light L along the ray from `pos`): what
// DdfFromLight::value computes after its own light->traceRay (lighting.cpp:63-72). Used by k_extend, which has just
// intersected the lights for this very ray, so the shading kernel does not have to trace the light a second time.
template <bool AREA = false>
__device__ __forceinline__ float light_pdf_at(const DevLight& L, f3 pos, f3 hit) {
    f3 dp = mk3(hit.x - pos.x, hit.y - pos.y, hit.z - pos.z);
    float decay = dot3(dp, dp);
    f3 n;
    if (AREA || L.kind <= IPT_LIGHT_AREA_TRIANGLE) n = mk3(L.nx, L.ny, L.nz);
    else if (!AREA) {
        float ir = 1.0f / L.radius;
        n = mk3(pmul(__fsub_rn(hit.x, L.px), ir), pmul(__fsub_rn(hit.y, L.py), ir), pmul(__fsub_rn(hit.z, L.pz), ir));
        if (L.kind == IPT_LIGHT_SPHERE_INVERTED) n = neg3(n);
    }
    float cosinus = pmul(-dot3(n, dp), rsqrtf(decay));
    if (cosinus < 0.0f) return 0.0f;
    return __fdividef(decay, pmul(cosinus, L.area));
}

// DdfFromLight::sample (src/lighting/lighting.cpp:50-59) over Light::sample (lighting.cpp:93-104, 172-207)
template <bool AREA = false, bool EXACT = false>
__device__ __forceinline__ f3 light_sample_dir(const DevLight& L, f3 pos, float u1, float u2) {
    f3 p, n;
    if (EXACT && (AREA || L.kind <= IPT_LIGHT_AREA_TRIANGLE)) {
        // AreaLight::sample + DdfFromLight::sample in the reference's operation order (lighting.cpp:93-104, 50-59)
        float v2 = L.kind == IPT_LIGHT_AREA_TRIANGLE ? xmul(u2, xsub(1.0f, u1)) : xmul(u2, 1.0f);
        f3 q = xadd3(xadd3(xscale3(mk3(L.xax, L.xay, L.xaz), u1), xscale3(mk3(L.yax, L.yay, L.yaz), v2)), mk3(L.px, L.py, L.pz));
        f3 dir = xnormalize3(xsub3(q, pos));
        float cosinus = xdot3(mk3(L.nx, L.ny, L.nz), neg3(dir));
        if (cosinus < 1e-5f) return mk3(0, 0, 0);
        return dir;
    }
    if (AREA || L.kind <= IPT_LIGHT_AREA_TRIANGLE) {
        float v2 = L.kind == IPT_LIGHT_AREA_TRIANGLE ? pmul(u2, __fsub_rn(1.0f, u1)) : u2;
        p = mk3(__fadd_rn(pfma(L.yax, v2, pmul(L.xax, u1)), L.px), __fadd_rn(pfma(L.yay, v2, pmul(L.xay, u1)), L.py),
                __fadd_rn(pfma(L.yaz, v2, pmul(L.xaz, u1)), L.pz));
        n = mk3(L.nx, L.ny, L.nz);
    } else {
        float z = pfma(u1, 2.0f, -1.0f);
        float r = sqrtf(fmaxf(0.0f, pfma(-z, z, 1.0f)));
        float sp, cp;
        sincospif(pmul(2.0f, u2), &sp, &cp);
        f3 unit = mk3(pmul(r, cp), pmul(r, sp), z);
        if (L.kind == IPT_LIGHT_POINT) {
            p = mk3(L.px, L.py, L.pz);
            n = unit;
        } else {
            p = mk3(pfma(unit.x, L.radius, L.px), pfma(unit.y, L.radius, L.py), pfma(unit.z, L.radius, L.pz));
            n = L.kind == IPT_LIGHT_SPHERE_INVERTED ? neg3(unit) : unit;
        }
    }
    f3 dp = mk3(__fsub_rn(p.x, pos.x), __fsub_rn(p.y, pos.y), __fsub_rn(p.z, pos.z));
    float inv = rsqrtf(dot3(dp, dp));
    f3 dir = mk3(pmul(dp.x, inv), pmul(dp.y, inv), pmul(dp.z, inv));
    float cosinus = -dot3(n, dir);
    if (cosinus < 1e-5f) return mk3(0, 0, 0); // facing back: failed sample
    return dir;
}

template <int MODE, bool X>
__device__ __forceinline__ float light_bvh_query(const DevScene& S, f3 o, f3 d, uint32_t& which, f3& lpos, float& best_len, TraceCounters& tc); // ipt_kernels.cuh

// UnionDdf::value over [lights..., sdf] (src/libddf/ddf.cpp:156-162) with the weights of main.cpp:143
__device__ __forceinline__ float mix_value(const DevScene& S, const Sdf& sdf, f3 pos, f3 w, float sdf_val) {
    float lp = 0.0f;
    if (S.light_inline) {
#pragma unroll
        for (int i = 0; i < IPT_INLINE_LIGHTS; ++i)
            if (i < (int)S.n_lights) lp += S.lights[i].weight * light_pdf(S.lights[i], pos, w);
    } else if (S.n_light_bvh) {
        uint32_t which;
        f3 lpos;
        float best_len;
        TraceCounters tc{0, 0, 0, 0};
        lp = light_bvh_query<0 /* LQ_PDF */, true>(S, pos, w, which, lpos, best_len, tc);
    } else {
        for (uint32_t i = 0; i < S.n_lights; ++i) {
            const DevLight& L = S.lights_g[i];
            lp += L.weight * light_pdf(L, pos, w);
        }
    }
    return lp + S.sdf_weight * sdf_val;
}

// UnionDdf::sample (src/libddf/ddf.cpp:138-154): scan the running float sum of weights with one draw `us`; the first
// component whose running sum exceeds it is sampled. r >= total (float rounding; uninitialised result in the
// reference) is a failed sample. Both candidate directions are formed by every lane (no light-vs-sdf divergence).
// INLINE_LIGHTS: 1 = the caller knows the lights are the inline ones (compile-time), 0 = ask the scene; AREA: ... and area lights
template <int INLINE_LIGHTS = 0, bool AREA = false, bool LAMBERT = false, bool EXACT = false>
__device__ __forceinline__ f3 mix_sample(const DevScene& S, const Sdf& sdf, const Basis& bn, const Basis& bl, f3 pos, float us, float u1, float u2, float ul) {
    f3 wl = mk3(0, 0, 0);
    float acc = 0.0f;
    bool from_light = false;
    if (INLINE_LIGHTS || S.light_inline) {
#pragma unroll
        for (int i = 0; i < IPT_INLINE_LIGHTS; ++i)
            if (i < (int)S.n_lights) {
                acc = S.lights[i].cdf;
                if (!from_light && us < acc) { from_light = true; wl = light_sample_dir<AREA, EXACT>(S.lights[i], pos, u1, u2); }
            }
    } else if (!INLINE_LIGHTS && S.n_lights) {
        // first i with us < cdf[i]; cdf is non-decreasing, so a binary search finds what the linear scan finds
        // (the guide table narrows the range to the bucket of us first, see DevScene::light_guide)
        const uint32_t bucket = (uint32_t)(us * (float)IPT_LIGHT_GUIDE);
        uint32_t lo = __ldg(&S.light_guide[bucket]), hi = min(__ldg(&S.light_guide[bucket + 1]) + 1u, S.n_lights);
        while (lo < hi) {
            uint32_t mid = (lo + hi) >> 1;
            if (us < __ldg(&S.light_cdf[mid])) hi = mid; else lo = mid + 1;
        }
#ifndef IPT_LIGHT_SAMP_RECORDS
#define IPT_LIGHT_SAMP_RECORDS 1 // 0: sample from the 112-byte DevLight record (tuning A/B only)
#endif
        if (!IPT_LIGHT_SAMP_RECORDS && lo < S.n_lights) { from_light = true; wl = light_sample_dir<false, EXACT>(S.lights_g[lo], pos, u1, u2); }
        else if (lo < S.n_lights) {
            from_light = true;
            f8 r0 = ldg256(&S.light_samp[4 * (size_t)lo]), r1 = ldg256(&S.light_samp[4 * (size_t)lo + 2]);
            DevLight L; // only the fields Light::sample reads
            L.px = r0.v[0]; L.py = r0.v[1]; L.pz = r0.v[2]; L.kind = __float_as_uint(r0.v[3]);
            L.xax = r0.v[4]; L.xay = r0.v[5]; L.xaz = r0.v[6]; L.radius = r0.v[7];
            L.yax = r1.v[0]; L.yay = r1.v[1]; L.yaz = r1.v[2];
            L.nx = r1.v[4]; L.ny = r1.v[5]; L.nz = r1.v[6];
            wl = light_sample_dir<false, EXACT>(L, pos, u1, u2);
        }
        acc = __ldg(&S.light_cdf[S.n_lights - 1]);
    }
    f3 ws = sdf_sample<LAMBERT, EXACT>(sdf, bn, bl, u1, u2, ul);
    if (from_light) return wl;
    acc += S.sdf_weight;
    if (us < acc) return ws;
    return mk3(0, 0, 0);
}

} // namespace iptd
