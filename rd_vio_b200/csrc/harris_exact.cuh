// harris_exact.cuh -- cv::cornerHarris at ONE pixel, in the reference's float op order, and the
// constants of the integer prefilter that decides which pixels need it (harris.cu, select.cu).
//
// Reference call site: cv::GFTTDetector::detect in OpenCvImage::detect_keypoints
// (src/rdvio_extra/src/opencv_image.cpp:44; Harris selected at :184-188).
#pragma once

#include "fe_internal.cuh"

namespace rdfe {

// ---- prefilter constants ------------------------------------------------------------------------------------------
// Integer Sobel: gxi = 3060 * gx, gyi = 3060 * gy exactly in real arithmetic; IA, IB, IC = 3x3 box sums of gxi^2,
// gxi*gyi, gyi^2 (exact, < 2^24); T = IA + IC; Ru = IA*IC - IB^2 - 0.04 T^2 = R / sigma^4 in real arithmetic
// (sigma = 1/3060).  The reference evaluates R in float32; a first-order error analysis of every rounding of both
// float orders (plain / FMA) bounds |R_float - Ru * sigma^4| <= (18.5 u T'^1.5 + 3.5 u T'^2) with u = 2^-24 and
// T' = T sigma^2; the float evaluation of Ru adds <= 1.5 u T'^2.  With a 25 % margin and in units of sigma^4:
//     eps(T) = T * (kHarrisC1 * sqrt(T) + kHarrisC2 * T) + kHarrisRhoU
// tests/test_harris_prefilter_model.py checks the bound against the oracle (worst observed: 4 % of the bound).
// kHarrisRhoU covers pixels whose integer gradients vanish on the whole 3x3 block while the reference's float
// gradients do not (rounding residue, |R| <= 71 u^4 = 8.9e-28): such pixels are never flagged, and a frame whose
// threshold max(R) * qualityLevel lies below kHarrisRhoS is recomputed exactly by select_kernel (harris_exact_all).
constexpr float kHarrisC1 = 24.0f * 3060.0f * 5.9604644775390625e-08f;   // 4.377e-3
constexpr float kHarrisC2 = 7.0f * 5.9604644775390625e-08f;              // 4.17e-7
constexpr float kHarrisRhoU = 4.0e-13f;                                  // * sigma^4 = 4.6e-27
constexpr float kHarrisRhoS = 1.0e-26f;

// Sobel gradients (scaled like cornerHarris: 1 / (4 * 3 * 255)) at the IN-IMAGE position (x, y).  Border taps come
// from the materialised REFLECT_101 halo of level 0, which is Sobel's BORDER_DEFAULT.
template <bool kFma>
__device__ __forceinline__ void harris_grad(const uint8_t *__restrict__ org, int ipitch, int x, int y, float &gx, float &gy) {
    const double sc = 1.0 / (4.0 * 3.0 * 255.0);
    const float k0 = (float)sc, k1 = (float)(2.0 * sc);
    const uint8_t *r0 = org + (ptrdiff_t)(y - 1) * ipitch + x;
    const uint8_t *r1 = r0 + ipitch, *r2 = r1 + ipitch;
    const float a0 = (float)r0[-1], a1 = (float)r0[0], a2 = (float)r0[1];
    const float b0 = (float)r1[-1], b2 = (float)r1[1];
    const float c0 = (float)r2[-1], c1 = (float)r2[0], c2 = (float)r2[1];
    const float d0 = a2 - a0, d1 = b2 - b0, d2 = c2 - c0;
    float s0, s2;
    if (!kFma) {
        gx = k1 * d1 + k0 * (d0 + d2);
        s0 = ((k0 * a0) + k1 * a1) + k0 * a2;
        s2 = ((k0 * c0) + k1 * c1) + k0 * c2;
    } else {
        gx = __fmaf_rn(k0, d0 + d2, k1 * d1);
        s0 = __fmaf_rn(k0, a2, __fmaf_rn(k1, a1, k0 * a0));
        s2 = __fmaf_rn(k0, c2, __fmaf_rn(k1, c1, k0 * c0));
    }
    gy = s2 - s0;
}

// R(x, y): products on the 3x3 block (boxFilter reflects the PRODUCT maps: BORDER_REFLECT_101 by index), float64 box
// sums (exact), response in float32.
template <bool kFma>
__device__ __noinline__ float harris_exact_at(const uint8_t *__restrict__ org, int ipitch, int W, int H, int x, int y, float k) {
    double a = 0.0, b = 0.0, c = 0.0;
#pragma unroll 1
    for (int dy = -1; dy <= 1; ++dy) {
        const int ry = reflect101(y + dy, H);
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
            const int rx = reflect101(x + dx, W);
            float gx, gy;
            harris_grad<kFma>(org, ipitch, rx, ry, gx, gy);
            a += (double)(gx * gx);
            b += (double)(gx * gy);
            c += (double)(gy * gy);
        }
    }
    const float A = (float)a, B = (float)b, C = (float)c;
    if (!kFma) return (A * C - B * B) - (k * (A + C)) * (A + C);
    return (A * C - B * B) - k * ((A + C) * (A + C));
}

// float -> double for values that are zero or normal (gradient products are never subnormal unless zero: a non-zero
// Sobel term is at least one ulp of an O(1e-4) float).  Integer re-biasing on the ALU pipe instead of the slow F2F:
// exponent + 896 unless the value is zero (min(mag >> 3, 1) is 0 only for zero).
__device__ __forceinline__ double harris_f2d(float f) {
    const unsigned u = __float_as_uint(f);
    const unsigned t = (u & 0x7FFFFFFFu) >> 3;
    const unsigned hi = (u & 0x80000000u) | (t + min(t, 1u) * 0x38000000u);
    return __hiloint2double((int)hi, (int)(u << 29));
}
__device__ __forceinline__ double harris_f2d_nonneg(float f) {      // f >= +0
    const unsigned u = __float_as_uint(f);
    const unsigned t = u >> 3;
    return __hiloint2double((int)(t + min(t, 1u) * 0x38000000u), (int)(u << 29));
}

// R(x, y) for 1 <= x <= W-2, 1 <= y <= H-2: every product position is inside the image (no reflection of the product
// maps) and the 5x5 pixel patch lies inside the materialised halo.  Same arithmetic as harris_exact_at, with the row
// terms of the 5 patch rows shared by the 9 positions (two aligned word loads per row).
template <bool kFma>
__device__ __forceinline__ float harris_exact_interior(const uint8_t *__restrict__ org, int ipitch, int x, int y, float k) {
    const double sc = 1.0 / (4.0 * 3.0 * 255.0);
    const float k0 = (float)sc, k1 = (float)(2.0 * sc);
    const uint8_t *base = org + (ptrdiff_t)(y - 2) * ipitch + (x - 2);
    const unsigned sh = ((unsigned)(size_t)base & 3u) * 8u;
    const unsigned *wp = reinterpret_cast<const unsigned *>(base - (sh >> 3));
    const int wpitch = ipitch >> 2;                          // row pitch is a multiple of 64 bytes
    float d[5][3], s[5][3];                                  // [-1 0 1] and [k0 k1 k0] row terms at patch columns 1..3
#pragma unroll
    for (int r = 0; r < 5; ++r) {
        const unsigned w0 = wp[r * wpitch], w1 = wp[r * wpitch + 1];
        const unsigned lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, 0u, sh);
        // byte -> float: PRMT builds the bits of 2^23 + b, one FADD removes the 2^23 (exact)
        float p[5];
        p[0] = __uint_as_float(__byte_perm(lo, 0x4B000000u, 0x7650u)) - 8388608.0f;
        p[1] = __uint_as_float(__byte_perm(lo, 0x4B000000u, 0x7651u)) - 8388608.0f;
        p[2] = __uint_as_float(__byte_perm(lo, 0x4B000000u, 0x7652u)) - 8388608.0f;
        p[3] = __uint_as_float(__byte_perm(lo, 0x4B000000u, 0x7653u)) - 8388608.0f;
        p[4] = __uint_as_float(__byte_perm(hi, 0x4B000000u, 0x7650u)) - 8388608.0f;
        float k0p[5];
#pragma unroll
        for (int c = 0; c < 5; ++c) k0p[c] = k0 * p[c];
#pragma unroll
        for (int c = 1; c <= 3; ++c) {
            d[r][c - 1] = p[c + 1] - p[c - 1];
            if (!kFma) s[r][c - 1] = (k0p[c - 1] + k1 * p[c]) + k0p[c + 1];
            else s[r][c - 1] = __fmaf_rn(k0, p[c + 1], __fmaf_rn(k1, p[c], k0p[c - 1]));
        }
    }
    double a = 0.0, b = 0.0, c2 = 0.0;
#pragma unroll
    for (int r = 1; r <= 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float gx;
            if (!kFma) gx = k1 * d[r][c] + k0 * (d[r - 1][c] + d[r + 1][c]);
            else gx = __fmaf_rn(k0, d[r - 1][c] + d[r + 1][c], k1 * d[r][c]);
            const float gy = s[r + 1][c] - s[r - 1][c];
            a += harris_f2d_nonneg(gx * gx);
            b += harris_f2d(gx * gy);
            c2 += harris_f2d_nonneg(gy * gy);
        }
    const float A = (float)a, B = (float)b, C = (float)c2;
    if (!kFma) return (A * C - B * B) - (k * (A + C)) * (A + C);
    return (A * C - B * B) - k * ((A + C) * (A + C));
}

// dispatch: interior pixels take the patch path, the 1-px frame the general one
template <bool kFma>
__device__ __forceinline__ float harris_exact(const uint8_t *__restrict__ org, int ipitch, int W, int H, int x, int y, float k) {
    if (x >= 1 && x <= W - 2 && y >= 1 && y <= H - 2) return harris_exact_interior<kFma>(org, ipitch, x, y, k);
    return harris_exact_at<kFma>(org, ipitch, W, H, x, y, k);
}

// Whole-frame exact recomputation by ONE CTA (called by select_kernel for degenerate frames, see harris_exact.cuh):
// response map into `map` (W*H floats), then the candidate keys and the frame maximum like the kernels above.
__device__ __noinline__ static void harris_exact_all(const uint8_t *__restrict__ org, int ipitch, int W, int H, float k, bool fma, float *map,
                                 unsigned long long *out, unsigned cap, unsigned *s_count, unsigned *s_max, unsigned *overflow) {
    const int tid = threadIdx.x, nt = blockDim.x;
    if (tid == 0) { *s_count = 0u; *s_max = 0u; }
    __syncthreads();
    float tmax = 0.0f;
    for (int p = tid; p < W * H; p += nt) {
        const int y = p / W, x = p - y * W;
        const float v = fma ? harris_exact_at<true>(org, ipitch, W, H, x, y, k) : harris_exact_at<false>(org, ipitch, W, H, x, y, k);
        map[p] = v;
        tmax = fmaxf(tmax, v);
    }
    if (tmax > 0.0f) atomicMax(s_max, __float_as_uint(tmax));
    __syncthreads();
    for (int p = tid; p < W * H; p += nt) {
        const int y = p / W, x = p - y * W;
        if (x < 1 || x >= W - 1 || y < 1 || y >= H - 1) continue;
        const float v = map[p];
        if (!(v > 0.0f)) continue;
        bool ok = true;
        for (int t = 0; t < 9; ++t)
            if (t != 4 && v < map[p + (t / 3 - 1) * W + (t % 3 - 1)]) ok = false;
        if (ok) {
            const unsigned pos = atomicAdd(s_count, 1u);
            if (pos < cap) out[pos] = ((unsigned long long)__float_as_uint(v) << 32) | (unsigned)p;
            else atomicExch(overflow, 1u);
        }
    }
    __syncthreads();
}

}  // namespace rdfe
