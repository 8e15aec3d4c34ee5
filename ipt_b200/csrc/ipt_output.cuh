// ipt_output.cuh — the OUTPUT STAGE of the reference (src/gui.cpp) on the device: the glare filter of
// Gui::updateDisplay, the tone mapping of normalize(), the 8-bit quantisation of Gui::save and a PNG writer.
// Included by ipt_capi.cu (one translation unit, shares fail()/CUDA_TRY/DevBuf and the plane type).
//
//   glare(image, cutoff)      gui.cpp:38-52  + draw_halo :28-36     k_bright_* (ordered compaction) + k_glare
//   normalize(image)          gui.cpp:11-16                           k_image_range + k_normalize
//   Gui::save pixel bytes     gui.cpp:192-194 + CImg.h:33169-33181,59204-59209   k_image_range + k_tone_bytes
//
// Parity (oracle/ipt_oracle_output.inc, pinned bit for bit to the reference's compiled gui.cpp):
//  * glare: bit-exact. The reference adds the halo of every bright pixel to every pixel in raster order of the bright
//    pixels; k_glare keeps that summation order per output pixel (one thread per output pixel walks the bright list,
//    which k_bright_scatter emits in raster order), and uses IEEE float sqrt/divide/add with the reference's
//    association. float(hypot(dx,dy)) == sqrtf(float(dx²+dy²)) for dx²+dy² < 2^24 (the integer is exact in binary32 and
//    the double rounding is innocuous: a double within 2^-53 of a binary32 midpoint would need |n - m²| < 2^-28 for a
//    25-bit m, but n·2^2e - M² is a non-zero integer); larger images take the double path.
//  * save bytes: bit-exact BY CONSTRUCTION against the host's libm. The byte of a pixel is a monotone step function of
//    its value; the 255 step positions are found on the host with the very expression of the reference (powf of the
//    box's libm, which is what the reference itself would call there) and the device only classifies pixels against
//    them. glibc's powf is not correctly rounded (0.82 ulp), so evaluating pow on the device could not be byte-exact.
//  * normalize (float image, display only): pow evaluated in double on the device, within 1 ulp of powf.
#pragma once

namespace iptd {

// order-preserving float -> uint key (NaNs are skipped by the callers, like CImg's `>` / `<` scans skip them)
__device__ __forceinline__ uint32_t f2key(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
inline float key2f(uint32_t k) {
    uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}

// range[0] = max key, range[1] = min key (initialised to 0 / 0xffffffff), range[2] = number of NaN or negative pixels
__global__ void __launch_bounds__(256) k_image_range(const float* __restrict__ img, size_t n, uint32_t* range) {
    uint32_t kmax = 0u, kmin = 0xffffffffu, bad = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float v = img[i];
        if (v != v) { ++bad; continue; }
        if (v < 0.0f) ++bad;
        uint32_t k = f2key(v == 0.0f ? 0.0f : v); // -0 == +0 for the scans
        kmax = max(kmax, k);
        kmin = min(kmin, k);
    }
    for (int o = 16; o; o >>= 1) {
        kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
        kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
        bad += __shfl_xor_sync(0xffffffffu, bad, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(&range[0], kmax);
        atomicMin(&range[1], kmin);
        if (bad) atomicAdd(&range[2], bad);
    }
}

// ---- glare ---------------------------------------------------------------------------------------------------
// one warp per image row counts its pixels brighter than the cutoff
__global__ void __launch_bounds__(256) k_bright_count(const float* __restrict__ img, uint32_t w, uint32_t h, float cutoff,
                                                      uint32_t* __restrict__ row_count) {
    uint32_t row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= h) return;
    uint32_t c = 0;
    for (uint32_t x0 = 0; x0 < w; x0 += 32) {
        uint32_t x = x0 + lane;
        bool b = x < w && !(img[(size_t)row * w + x] <= cutoff); // gui.cpp:43 `if (val <= cutoff) continue;`
        c += __popc(__ballot_sync(0xffffffffu, b));
    }
    if (lane == 0) row_count[row] = c;
}
// exclusive scan of the row counts (h <= a few thousand: one block, serial over chunks)
__global__ void __launch_bounds__(1024) k_row_scan(const uint32_t* __restrict__ row_count, uint32_t h, uint32_t* __restrict__ row_offset,
                                                   uint32_t* __restrict__ total) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < h; base += blockDim.x) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = i < h ? row_count[i] : 0u, incl = v;
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((threadIdx.x & 31) >= o) incl += t;
        }
        if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t s = warp_sums[threadIdx.x], si = s;
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, si, o);
                if (threadIdx.x >= o) si += t;
            }
            warp_sums[threadIdx.x] = si - s;
        }
        __syncthreads();
        uint32_t excl = carry + warp_sums[threadIdx.x >> 5] + incl - v;
        if (i < h) row_offset[i] = excl;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}
// bright list in RASTER order: (x, y, C, -) with C = cutoff * float(0.1 * val / cutoff)   (gui.cpp:46-47)
__global__ void __launch_bounds__(256) k_bright_scatter(const float* __restrict__ img, uint32_t w, uint32_t h, float cutoff,
                                                        const uint32_t* __restrict__ row_offset, float4* __restrict__ bright) {
    uint32_t row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= h) return;
    uint32_t pos = row_offset[row];
    for (uint32_t x0 = 0; x0 < w; x0 += 32) {
        uint32_t x = x0 + lane;
        float val = x < w ? img[(size_t)row * w + x] : 0.0f;
        bool b = x < w && !(val <= cutoff);
        uint32_t ballot = __ballot_sync(0xffffffffu, b);
        if (b) {
            float coef = __double2float_rn(__ddiv_rn(__dmul_rn(0.1, (double)val), (double)cutoff));
            float c = __fmul_rn(cutoff, coef);
            bright[pos + __popc(ballot & ((1u << lane) - 1u))] = make_float4(__int_as_float((int)x), __int_as_float((int)row), c, 0.0f);
        }
        pos += __popc(ballot);
    }
}

// (c / s) / s with IEEE rounding of both quotients: xdiv_n twice, sharing the refined reciprocal of s
__device__ __forceinline__ float xdiv2_n(float c, float s) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(s));
    float e = __fmaf_rn(-s, r, 1.0f);
    r = __fmaf_rn(r, e, r);
    float q = __fmul_rn(c, r);
    q = __fmaf_rn(r, __fmaf_rn(-s, q, c), q);
    float p = __fmul_rn(q, r);
    return __fmaf_rn(r, __fmaf_rn(-s, p, q), p);
}

#ifndef IPT_GLARE_TILE
#define IPT_GLARE_TILE 1024
#endif
// out(x,y) = cut(((img(x,y) + halo_1) + halo_2) + ..., 0, cutoff), halos in raster order of the bright pixels.
// WIDE: dx²+dy² may exceed 2^24 -> double sqrt. FASTDIV: every C is in the range where xdiv_n is exact.
template <bool WIDE, bool FASTDIV>
__global__ void __launch_bounds__(256) k_glare(const float* __restrict__ img, uint32_t w, uint32_t h, const float4* __restrict__ bright,
                                               uint32_t nb, float lo, float hi, float* __restrict__ out) {
    __shared__ float4 tile[IPT_GLARE_TILE];
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t n = (size_t)w * h;
    bool live = i < n;
    int x = live ? (int)(i % w) : 0, y = live ? (int)(i / w) : 0;
    float acc = live ? img[i] : 0.0f;
    for (uint32_t base = 0; base < nb; base += IPT_GLARE_TILE) {
        uint32_t m = min((uint32_t)IPT_GLARE_TILE, nb - base);
        __syncthreads();
        for (uint32_t k = threadIdx.x; k < m; k += blockDim.x) tile[k] = bright[base + k];
        __syncthreads();
#pragma unroll 4
        for (uint32_t k = 0; k < m; ++k) {
            float4 b = tile[k];
            int dx = x - __float_as_int(b.x), dy = y - __float_as_int(b.y);
            float r;
            if (WIDE) {
                long long n2 = (long long)dx * dx + (long long)dy * dy;
                r = __double2float_rn(__dsqrt_rn((double)n2));
            } else {
                int n2 = dx * dx + dy * dy;
                r = n2 ? xsqrt_n((float)n2) : 0.0f; // exact integer in binary32 (< 2^24), IEEE sqrt
            }
            float s = __fadd_rn(0.25f, r);        // r0 + r
            float v = FASTDIV ? xdiv2_n(b.z, s) : __fdiv_rn(__fdiv_rn(b.z, s), s);
            acc = __fadd_rn(acc, v);
        }
    }
    if (live) out[i] = acc < lo ? lo : acc > hi ? hi : acc; // cimg::cut
}

// ---- tone mapping --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_normalize(const float* __restrict__ img, size_t n, float maxv, float p, float* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float q = __fdiv_rn(img[i], maxv);
    float v = __double2float_rn(pow((double)q, (double)p));
    out[i] = v < 0.0f ? 0.0f : v > 1.0f ? 1.0f : v;
}
// byte = number of step positions <= value; thr[k-1] = bit pattern of the smallest value whose byte is >= k
__global__ void __launch_bounds__(256) k_tone_bytes(const float* __restrict__ img, size_t n, const uint32_t* __restrict__ thr,
                                                    uint8_t* __restrict__ out) {
    __shared__ uint32_t t[256];
    t[threadIdx.x] = threadIdx.x < 255 ? thr[threadIdx.x] : 0xffffffffu;
    __syncthreads();
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = img[i];
    uint32_t u = __float_as_uint(v == 0.0f ? 0.0f : v);
    uint32_t lo = 0, hi = 255; // number of thresholds <= u
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (t[mid] <= u) lo = mid + 1; else hi = mid;
    }
    out[i] = (uint8_t)lo;
}

// ---- host side -----------------------------------------------------------------------------------------------
struct ToneCurve {
    float maxv = 0, p = 0, m = 0, M = 0;
    bool constant = false;
    // Gui::save's value chain for ONE pixel value (gui.cpp:11-16,192-194; CImg.h:33169-33181,59209)
    float normalized(float v) const {
        float w = std::pow(v / maxv, p); // std::pow(float,float) == powf, CImg.h:28859
        return w < 0.0f ? 0.0f : w > 1.0f ? 1.0f : w;
    }
    unsigned byte_of(float v) const {
        if (constant) return 0;
        float t = (normalized(v) - m) / (M - m) * (255.0f - 0.0f) + 0.0f;
        return (unsigned char)t;
    }
};
static ToneCurve make_tone_curve(float vmin, float vmax) {
    ToneCurve c;
    float inv_gamma = 2.2; // gui.cpp:12
    c.maxv = vmax;
    c.p = (float)(double)(1.0f / inv_gamma);
    c.m = c.normalized(vmin);
    c.M = c.normalized(vmax);
    c.constant = c.m == c.M;
    return c;
}
static void tone_thresholds(const ToneCurve& c, float vmin, float vmax, uint32_t thr[255]) {
    uint32_t bmin, bmax;
    std::memcpy(&bmin, &vmin, 4);
    std::memcpy(&bmax, &vmax, 4);
    for (unsigned k = 1; k <= 255; ++k) {
        if (c.byte_of(vmax) < k) { thr[k - 1] = 0xffffffffu; continue; }
        if (c.byte_of(vmin) >= k) { thr[k - 1] = bmin; continue; }
        uint32_t lo = bmin, hi = bmax; // byte_of(lo) < k <= byte_of(hi); non-negative floats order like their bit patterns
        while (hi - lo > 1) {
            uint32_t mid = lo + (hi - lo) / 2;
            float v;
            std::memcpy(&v, &mid, 4);
            if (c.byte_of(v) >= k) hi = mid; else lo = mid;
        }
        thr[k - 1] = hi;
    }
}

static inline unsigned grid_for(size_t n, unsigned block) { return (unsigned)((n + block - 1) / block); }

// min / max of a device image and the number of NaN or negative pixels
static int image_range(OutputScratch& ws, const float* d_img, size_t n, cudaStream_t st, float* vmin, float* vmax, uint32_t* bad) {
    CUDA_TRY(ws.small.ensure(4096));
    uint32_t* d_range = ws.small.as<uint32_t>() + 512; // [0,255) holds the tone thresholds
    uint32_t init[3] = {0u, 0xffffffffu, 0u}, got[3];
    CUDA_TRY(cudaMemcpyAsync(d_range, init, sizeof init, cudaMemcpyHostToDevice, st));
    k_image_range<<<std::min<unsigned>(grid_for(n, 256), 148 * 8), 256, 0, st>>>(d_img, n, d_range);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(got, d_range, sizeof got, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    *bad = got[2];
    *vmax = key2f(got[0]);
    *vmin = key2f(got[1]);
    return IPT_OK;
}

static int device_glare(OutputScratch& ws, const float* d_img, uint32_t w, uint32_t h, float cutoff, float* d_out, cudaStream_t st,
                        uint32_t* n_bright) {
    size_t n = (size_t)w * h;
    CUDA_TRY(ws.rows.ensure(4 * (2 * (size_t)h + 1)));
    uint32_t *d_count = ws.rows.as<uint32_t>(), *d_offset = d_count + h, *d_total = d_count + 2 * (size_t)h;
    unsigned row_blocks = grid_for((size_t)h * 32, 256);
    k_bright_count<<<row_blocks, 256, 0, st>>>(d_img, w, h, cutoff, d_count);
    k_row_scan<<<1, 1024, 0, st>>>(d_count, h, d_offset, d_total);
    CUDA_TRY(cudaGetLastError());
    uint32_t nb = 0;
    CUDA_TRY(cudaMemcpyAsync(&nb, d_total, 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(ws.bright.ensure(sizeof(float4) * std::max<size_t>(nb, 1)));
    float4* d_bright = ws.bright.as<float4>();
    if (nb) k_bright_scatter<<<row_blocks, 256, 0, st>>>(d_img, w, h, cutoff, d_offset, d_bright);
    float lo = 0.0f < cutoff ? 0.0f : cutoff, hi = 0.0f < cutoff ? cutoff : 0.0f; // cut() orders its bounds, CImg.h:33336
    bool wide = (double)(w - 1) * (w - 1) + (double)(h - 1) * (h - 1) >= 16777216.0;
    // xdiv_n is exact while no intermediate leaves the normal range: C = 0.1*val, s in [0.25, 2^13)
    float vmin = 0, vmax = 0;
    uint32_t bad = 0;
    int rc = image_range(ws, d_img, n, st, &vmin, &vmax, &bad);
    if (rc != IPT_OK) return rc;
    bool fastdiv = !bad && cutoff > 1e-15f && vmax < 1e15f;
    unsigned g = grid_for(n, 256);
    if (wide) k_glare<true, false><<<g, 256, 0, st>>>(d_img, w, h, d_bright, nb, lo, hi, d_out);
    else if (fastdiv) k_glare<false, true><<<g, 256, 0, st>>>(d_img, w, h, d_bright, nb, lo, hi, d_out);
    else k_glare<false, false><<<g, 256, 0, st>>>(d_img, w, h, d_bright, nb, lo, hi, d_out);
    CUDA_TRY(cudaGetLastError());
    if (n_bright) *n_bright = nb;
    return IPT_OK;
}

static int device_normalize(OutputScratch& ws, const float* d_img, size_t n, float* d_out, cudaStream_t st) {
    float vmin = 0, vmax = 0;
    uint32_t bad = 0;
    int rc = image_range(ws, d_img, n, st, &vmin, &vmax, &bad);
    if (rc != IPT_OK) return rc;
    // radiance is never NaN or negative; the reference's own result for such an image is undefined (powf of a
    // negative, (unsigned char) of a NaN)
    if (bad) return fail(IPT_ERR_INVALID, "image has NaN or negative pixels");
    if (!(vmax > 0.0f)) return fail(IPT_ERR_INVALID, "image is black: normalize() divides by its maximum (gui.cpp:13-14)");
    float inv_gamma = 2.2;
    k_normalize<<<grid_for(n, 256), 256, 0, st>>>(d_img, n, vmax, (float)(double)(1.0f / inv_gamma), d_out);
    CUDA_TRY(cudaGetLastError());
    return IPT_OK;
}

static int device_save_bytes(OutputScratch& ws, const float* d_img, size_t n, uint8_t* d_out, cudaStream_t st) {
    float vmin = 0, vmax = 0;
    uint32_t bad = 0;
    int rc = image_range(ws, d_img, n, st, &vmin, &vmax, &bad);
    if (rc != IPT_OK) return rc;
    // radiance is never NaN or negative; the reference's own result for such an image is undefined (powf of a
    // negative, (unsigned char) of a NaN)
    if (bad) return fail(IPT_ERR_INVALID, "image has NaN or negative pixels");
    if (!(vmax > 0.0f)) return fail(IPT_ERR_INVALID, "image is black: normalize() divides by its maximum (gui.cpp:13-14)");
    ToneCurve curve = make_tone_curve(vmin, vmax);
    uint32_t thr[255];
    tone_thresholds(curve, vmin, vmax, thr);
    uint32_t* d_thr = ws.small.as<uint32_t>(); // ensured by image_range
    CUDA_TRY(cudaMemcpyAsync(d_thr, thr, sizeof thr, cudaMemcpyHostToDevice, st)); // pageable source: staged before return
    k_tone_bytes<<<grid_for(n, 256), 256, 0, st>>>(d_img, n, d_thr, d_out);
    CUDA_TRY(cudaGetLastError());
    return IPT_OK;
}

// ---- PNG (8-bit grey, stored deflate blocks: no zlib dependency) -----------------------------------------------
static uint32_t crc32_update(uint32_t crc, const uint8_t* p, size_t n) {
    static uint32_t table[256];
    static bool ready = false;
    if (!ready) {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xedb88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        ready = true;
    }
    for (size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 0xff] ^ (crc >> 8);
    return crc;
}
static void png_chunk(std::vector<uint8_t>& out, const char type[4], const std::vector<uint8_t>& data) {
    auto be32 = [&](uint32_t v) { for (int s = 24; s >= 0; s -= 8) out.push_back((uint8_t)(v >> s)); };
    be32((uint32_t)data.size());
    size_t start = out.size();
    out.insert(out.end(), type, type + 4);
    out.insert(out.end(), data.begin(), data.end());
    be32(crc32_update(0xffffffffu, out.data() + start, out.size() - start) ^ 0xffffffffu);
}
static std::vector<uint8_t> png_encode_gray8(const uint8_t* px, uint32_t w, uint32_t h) {
    std::vector<uint8_t> raw;
    raw.reserve(((size_t)w + 1) * h);
    for (uint32_t y = 0; y < h; ++y) {
        raw.push_back(0); // filter: none
        raw.insert(raw.end(), px + (size_t)y * w, px + (size_t)(y + 1) * w);
    }
    std::vector<uint8_t> z = {0x78, 0x01};
    uint32_t a = 1, b = 0; // adler32
    size_t pos = 0;
    do {
        size_t len = std::min<size_t>(65535, raw.size() - pos);
        z.push_back(pos + len == raw.size() ? 1 : 0);
        z.push_back((uint8_t)(len & 0xff)); z.push_back((uint8_t)(len >> 8));
        z.push_back((uint8_t)(~len & 0xff)); z.push_back((uint8_t)((~len >> 8) & 0xff));
        for (size_t i = 0; i < len; ++i) { a = (a + raw[pos + i]) % 65521u; b = (b + a) % 65521u; }
        z.insert(z.end(), raw.begin() + pos, raw.begin() + pos + len);
        pos += len;
    } while (pos < raw.size());
    uint32_t adler = (b << 16) | a;
    for (int s = 24; s >= 0; s -= 8) z.push_back((uint8_t)(adler >> s));
    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    std::vector<uint8_t> ihdr;
    for (int s = 24; s >= 0; s -= 8) ihdr.push_back((uint8_t)(w >> s));
    for (int s = 24; s >= 0; s -= 8) ihdr.push_back((uint8_t)(h >> s));
    ihdr.insert(ihdr.end(), {8, 0, 0, 0, 0}); // 8 bit, grey, deflate, adaptive filtering, no interlace
    png_chunk(out, "IHDR", ihdr);
    png_chunk(out, "IDAT", z);
    png_chunk(out, "IEND", {});
    return out;
}

} // namespace iptd

extern "C" {

// host-image entry points: upload, run, download
static int with_device_image(int device, const float* image, uint32_t w, uint32_t h, DevBuf<float>& d_img) {
    if (!image || !w || !h) return fail(IPT_ERR_INVALID, "null or empty image");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(IPT_ERR_NO_DEVICE, "no CUDA device: the output stage runs on the GPU (no CPU fallback)"); }
    if (device < 0 || device >= ndev) return fail(IPT_ERR_INVALID, "bad device index");
    CUDA_TRY(cudaSetDevice(device));
    size_t n = (size_t)w * h;
    CUDA_TRY(d_img.alloc(n));
    CUDA_TRY(cudaMemcpy(d_img.p, image, 4 * n, cudaMemcpyHostToDevice));
    return IPT_OK;
}

int ipt_image_glare(int device, const float* image, uint32_t width, uint32_t height, float cutoff, float* out, uint32_t* n_bright) {
    if (!out) return fail(IPT_ERR_INVALID, "null argument");
    DevBuf<float> d_img, d_out;
    int rc = with_device_image(device, image, width, height, d_img);
    if (rc != IPT_OK) return rc;
    size_t n = (size_t)width * height;
    CUDA_TRY(d_out.alloc(n));
    OutputScratch ws;
    rc = iptd::device_glare(ws, d_img.p, width, height, cutoff, d_out.p, 0, n_bright);
    if (rc != IPT_OK) return rc;
    CUDA_TRY(cudaMemcpy(out, d_out.p, 4 * n, cudaMemcpyDeviceToHost));
    return IPT_OK;
}

int ipt_image_normalize(int device, const float* image, uint32_t width, uint32_t height, float* out) {
    if (!out) return fail(IPT_ERR_INVALID, "null argument");
    DevBuf<float> d_img, d_out;
    int rc = with_device_image(device, image, width, height, d_img);
    if (rc != IPT_OK) return rc;
    size_t n = (size_t)width * height;
    CUDA_TRY(d_out.alloc(n));
    OutputScratch ws;
    rc = iptd::device_normalize(ws, d_img.p, n, d_out.p, 0);
    if (rc != IPT_OK) return rc;
    CUDA_TRY(cudaMemcpy(out, d_out.p, 4 * n, cudaMemcpyDeviceToHost));
    return IPT_OK;
}

int ipt_image_save_bytes(int device, const float* image, uint32_t width, uint32_t height, uint8_t* out) {
    if (!out) return fail(IPT_ERR_INVALID, "null argument");
    DevBuf<float> d_img;
    DevBuf<uint8_t> d_out;
    int rc = with_device_image(device, image, width, height, d_img);
    if (rc != IPT_OK) return rc;
    size_t n = (size_t)width * height;
    CUDA_TRY(d_out.alloc(n));
    OutputScratch ws;
    rc = iptd::device_save_bytes(ws, d_img.p, n, d_out.p, 0);
    if (rc != IPT_OK) return rc;
    CUDA_TRY(cudaMemcpy(out, d_out.p, n, cudaMemcpyDeviceToHost));
    return IPT_OK;
}

// plane entry points: the accumulators never leave the device
int ipt_plane_display(ipt_plane* p, float glare_cutoff, float* out, float* ms) {
    if (!p || !out) return fail(IPT_ERR_INVALID, "null argument");
    std::lock_guard<std::recursive_mutex> lock__(p->scene->mu);
    CUDA_TRY(cudaSetDevice(p->scene->device));
    cudaStream_t st = p->scene->stream;
    size_t n = (size_t)p->width * p->height;
    OutputScratch& ws = p->scene->out;
    CUDA_TRY(ws.mean.ensure(4 * n));
    CUDA_TRY(ws.glare.ensure(4 * n));
    CUDA_TRY(ws.shown.ensure(4 * n));
    float *d_mean = ws.mean.as<float>(), *d_glare = ws.glare.as<float>(), *d_norm = ws.shown.as<float>();
    cudaEvent_t e0 = p->scene->ev_begin, e1 = p->scene->ev_end; // the scene's own pair (the lock is held)
    CUDA_TRY(cudaEventRecord(e0, st));
    k_plane_resolve<<<iptd::grid_for(n, 256), 256, 0, st>>>(p->sum, p->count, n, d_mean);
    int rc = iptd::device_glare(ws, d_mean, p->width, p->height, glare_cutoff, d_glare, st, nullptr); // gui.cpp:84
    if (rc == IPT_OK) rc = iptd::device_normalize(ws, d_glare, n, d_norm, st);                       // gui.cpp:87
    if (rc == IPT_OK) {
        cudaEventRecord(e1, st);
        cudaError_t e = cudaMemcpyAsync(out, d_norm, 4 * n, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) rc = fail(IPT_ERR_CUDA, cudaGetErrorString(e));
        else if (ms) cudaEventElapsedTime(ms, e0, e1);
    }
    return rc;
}

int ipt_plane_save_bytes(ipt_plane* p, uint8_t* out) {
    if (!p || !out) return fail(IPT_ERR_INVALID, "null argument");
    std::lock_guard<std::recursive_mutex> lock__(p->scene->mu);
    CUDA_TRY(cudaSetDevice(p->scene->device));
    cudaStream_t st = p->scene->stream;
    size_t n = (size_t)p->width * p->height;
    OutputScratch& ws = p->scene->out;
    CUDA_TRY(ws.mean.ensure(4 * n));
    CUDA_TRY(ws.bytes.ensure(n));
    k_plane_resolve<<<iptd::grid_for(n, 256), 256, 0, st>>>(p->sum, p->count, n, ws.mean.as<float>());
    int rc = iptd::device_save_bytes(ws, ws.mean.as<float>(), n, ws.bytes.as<uint8_t>(), st);
    if (rc != IPT_OK) return rc;
    CUDA_TRY(cudaMemcpyAsync(out, ws.bytes.as<uint8_t>(), n, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return IPT_OK;
}

int ipt_write_png_gray8(const char* path, const uint8_t* bytes, uint32_t width, uint32_t height) {
    if (!path || !bytes || !width || !height) return fail(IPT_ERR_INVALID, "null or empty argument");
    std::vector<uint8_t> png = iptd::png_encode_gray8(bytes, width, height);
    FILE* f = std::fopen(path, "wb");
    if (!f) return fail(IPT_ERR_INVALID, std::string("cannot open ") + path);
    bool ok = std::fwrite(png.data(), 1, png.size(), f) == png.size();
    ok = std::fclose(f) == 0 && ok;
    return ok ? IPT_OK : fail(IPT_ERR_INVALID, std::string("short write to ") + path);
}

int ipt_plane_save_png(ipt_plane* p, const char* path) {
    if (!p || !path) return fail(IPT_ERR_INVALID, "null argument");
    std::vector<uint8_t> bytes((size_t)p->width * p->height);
    int rc = ipt_plane_save_bytes(p, bytes.data());
    if (rc != IPT_OK) return rc;
    return ipt_write_png_gray8(path, bytes.data(), p->width, p->height);
}

} // extern "C"
