// undistort.cu -- ingest stage, "next" rows of the scope table: the cv::undistort(img, out, K, D) that the
// reference's dataset reader applies to every frame before the Image plugin sees pixels
// (/root/reference/examples/dataset.hpp:232-236, :591; dead twin OpenCvImage::correct_distortion,
// src/rdvio_extra/src/opencv_image.cpp:163-177).
//
// OpenCV: per stripe of rows initUndistortRectifyMap(A, D, I, Ar with cy shifted by the stripe start, CV_16SC2)
// + remap(INTER_LINEAR, BORDER_CONSTANT 0).  The map depends only on the calibration, so it is built ONCE on the
// host in float64 with OpenCV's own scalar expression order (rdfe_set_undistort) and kept in HBM as
// (sx | sy << 16) words + (fy*32+fx) halves; per frame the kernel below is a pure integer gather:
//   dst = (p00*(32-fx)(32-fy)*32 + p01*fx(32-fy)*32 + p10*(32-fx)fy*32 + p11*fx*fy*32 + 2^14) >> 15.
#include <cmath>
#include <vector>

#include "fe_internal.cuh"

namespace rdfe {

// cv::cvtColor(BGR2GRAY / BGRA2GRAY) for 8-bit (rdvio.hpp:42-49): 15-bit fixed point,
// (B*3735 + G*19235 + R*9798 + 2^14) >> 15 -- verified bit-exact against cv2 4.13.
__device__ __forceinline__ int bgr_to_gray(int b, int g, int r) { return (b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15; }

// Ingest kernel: optional undistortion (per channel, like cv::undistort on the loaded image) followed by the
// optional gray conversion of Odometry::addFrame; writes the 8-bit gray frame CLAHE reads.
template <int CH, bool UND>
__global__ void __launch_bounds__(256)
ingest_kernel(const uint8_t *const *__restrict__ src, size_t src_pitch, const uint32_t *__restrict__ map_xy,
              const uint16_t *__restrict__ map_f, uint8_t *const *__restrict__ dst, size_t dst_pitch, int W, int H) {
    const int groups = (W + 3) >> 2;
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= groups * H) return;
    const int y = g / groups, x = (g - y * groups) << 2;
    const uint8_t *img = src[blockIdx.y];
    unsigned out = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int xi = x + i;
        if (xi >= W) break;
        int ch[3] = {0, 0, 0};
        if (UND) {
            const size_t o = (size_t)y * W + xi;
            const uint32_t m = __ldg(map_xy + o);
            const unsigned f = __ldg(map_f + o);
            const int sx = (int)(short)(m & 0xFFFFu), sy = (int)m >> 16;
            const int wx1 = (int)(f & 31u), wx0 = 32 - wx1, wy1 = (int)(f >> 5), wy0 = 32 - wy1;
            const bool x0 = (unsigned)sx < (unsigned)W, x1 = (unsigned)(sx + 1) < (unsigned)W;
            const bool y0 = (unsigned)sy < (unsigned)H, y1 = (unsigned)(sy + 1) < (unsigned)H;
            const uint8_t *r0 = img + (size_t)sy * src_pitch + (size_t)sx * CH;
            const uint8_t *r1 = r0 + src_pitch;
#pragma unroll
            for (int c = 0; c < (CH >= 3 ? 3 : 1); ++c) {
                // BORDER_CONSTANT, value 0
                const int p00 = (y0 && x0) ? __ldg(r0 + c) : 0, p01 = (y0 && x1) ? __ldg(r0 + CH + c) : 0;
                const int p10 = (y1 && x0) ? __ldg(r1 + c) : 0, p11 = (y1 && x1) ? __ldg(r1 + CH + c) : 0;
                const int v = ((p00 * wx0 + p01 * wx1) * wy0 + (p10 * wx0 + p11 * wx1) * wy1) * 32 + (1 << 14) >> 15;
                ch[c] = min(v, 255);
            }
        } else {
            const uint8_t *p = img + (size_t)y * src_pitch + (size_t)xi * CH;
#pragma unroll
            for (int c = 0; c < (CH >= 3 ? 3 : 1); ++c) ch[c] = __ldg(p + c);
        }
        const int v = (CH >= 3) ? bgr_to_gray(ch[0], ch[1], ch[2]) : ch[0];
        out |= (unsigned)v << (8 * i);
    }
    uint8_t *d = dst[blockIdx.y] + (size_t)y * dst_pitch + x;
    if (x + 4 <= W) *reinterpret_cast<unsigned *>(d) = out;
    else for (int i = 0; x + i < W; ++i) d[i] = (uint8_t)(out >> (8 * i));
}

// Gray frames with 4-byte aligned rows and W % 4 == 0 (the usual case): same arithmetic, fewer instructions.  A thread
// still owns 4 adjacent output pixels; their map entries come as one 16-byte and one 8-byte load.  Per pixel the two
// source rows are read as aligned 32-bit words (a second word only when the 2-pixel footprint straddles a word
// boundary) and re-aligned with a funnel shift, so that (p00, p01) and (p10, p11) are the low byte pairs of two
// registers; the blend is then two dp2a against the four 10-bit weight products:
//   ((p00 wx0 + p01 wx1) wy0 + (p10 wx0 + p11 wx1) wy1) * 32 + 2^14 >> 15  ==  (sum_k p_k w_k + 512) >> 10   (integers).
// A footprint that touches the image border takes the tap-by-tap BORDER_CONSTANT path of the general kernel.
__global__ void __launch_bounds__(256)
undistort_gray_fast_kernel(const uint8_t *const *__restrict__ src, size_t src_pitch, const uint32_t *__restrict__ map_xy,
                           const uint16_t *__restrict__ map_f, uint8_t *const *__restrict__ dst, size_t dst_pitch, int W, int H) {
    const int groups = W >> 2;
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= groups * H) return;
    const int y = g / groups, x = (g - y * groups) << 2;
    const uint8_t *img = src[blockIdx.y];
    const size_t o = (size_t)y * W + x;                                  // multiple of 4
    const uint4 m4 = __ldg(reinterpret_cast<const uint4 *>(map_xy + o));
    const uint2 f2 = __ldg(reinterpret_cast<const uint2 *>(map_f + o));
    const unsigned ms[4] = {m4.x, m4.y, m4.z, m4.w};
    const unsigned fs[4] = {f2.x & 0xFFFFu, f2.x >> 16, f2.y & 0xFFFFu, f2.y >> 16};
    unsigned out = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int sx = (int)(short)(ms[i] & 0xFFFFu), sy = (int)ms[i] >> 16;
        const unsigned wx1 = fs[i] & 31u, wx0 = 32u - wx1, wy1 = fs[i] >> 5, wy0 = 32u - wy1;
        const unsigned wt = (wx0 * wy0) | ((wx1 * wy0) << 16), wb = (wx0 * wy1) | ((wx1 * wy1) << 16);
        unsigned top, bot;                                               // low two bytes: (p00, p01) / (p10, p11)
        if ((unsigned)sx < (unsigned)(W - 1) && (unsigned)sy < (unsigned)(H - 1)) {
            const unsigned sh = ((unsigned)sx & 3u) * 8u;
            const uint8_t *r0 = img + (size_t)sy * src_pitch + (size_t)(sx & ~3);
            const uint8_t *r1 = r0 + src_pitch;
            const unsigned t0 = __ldg(reinterpret_cast<const unsigned *>(r0)), b0 = __ldg(reinterpret_cast<const unsigned *>(r1));
            unsigned t1 = 0u, b1 = 0u;
            if (sh == 24u) { t1 = __ldg(reinterpret_cast<const unsigned *>(r0 + 4)); b1 = __ldg(reinterpret_cast<const unsigned *>(r1 + 4)); }
            top = __funnelshift_r(t0, t1, sh);
            bot = __funnelshift_r(b0, b1, sh);
        } else {
            const bool x0 = (unsigned)sx < (unsigned)W, x1 = (unsigned)(sx + 1) < (unsigned)W;
            const bool y0 = (unsigned)sy < (unsigned)H, y1 = (unsigned)(sy + 1) < (unsigned)H;
            const uint8_t *r0 = img + (ptrdiff_t)sy * (ptrdiff_t)src_pitch + sx;
            const uint8_t *r1 = r0 + src_pitch;
            const unsigned p00 = (y0 && x0) ? __ldg(r0) : 0u, p01 = (y0 && x1) ? __ldg(r0 + 1) : 0u;
            const unsigned p10 = (y1 && x0) ? __ldg(r1) : 0u, p11 = (y1 && x1) ? __ldg(r1 + 1) : 0u;
            top = p00 | (p01 << 8);
            bot = p10 | (p11 << 8);
        }
        const unsigned acc = __dp2a_lo(wb, bot, __dp2a_lo(wt, top, 512u));
        out |= min(acc >> 10, 255u) << (8 * i);
    }
    *reinterpret_cast<unsigned *>(dst[blockIdx.y] + (size_t)y * dst_pitch + x) = out;
}

int launch_undistort(rdfe_ctx *ctx, int n, const uint8_t *const *d_src, size_t src_pitch, int src_vec4, uint8_t *const *d_dst,
                     size_t dst_pitch) {
    const int W = ctx->cfg.width, H = ctx->cfg.height;
    const int groups = (W + 3) >> 2;
    dim3 grid((groups * H + 255) / 256, n);
    const int ch = ctx->in_channels;
    if (ctx->und_on && ch == 1 && src_vec4 && W % 4 == 0 && dst_pitch % 4 == 0) {
        RDFE_LAUNCH(ctx, K_UNDISTORT, (undistort_gray_fast_kernel<<<grid, 256, 0, ctx->ls>>>(d_src, src_pitch, ctx->und_map_xy, ctx->und_map_f,
                                                                                           d_dst, dst_pitch, W, H)));
        return 1;
    }
#define RDFE_INGEST(CH, UND)                                                                                            \
    RDFE_LAUNCH(ctx, K_UNDISTORT, (ingest_kernel<CH, UND><<<grid, 256, 0, ctx->ls>>>(d_src, src_pitch, ctx->und_map_xy,  \
                                                                                   ctx->und_map_f, d_dst, dst_pitch, W, H)))
    if (ctx->und_on) {
        if (ch == 1) RDFE_INGEST(1, true); else if (ch == 3) RDFE_INGEST(3, true); else RDFE_INGEST(4, true);
    } else {
        if (ch == 3) RDFE_INGEST(3, false); else RDFE_INGEST(4, false);
    }
#undef RDFE_INGEST
    return 1;
}

// Host: the fixed-point map exactly as cv::undistort builds it (float64, stripe by stripe, scalar expression order
// of initUndistortRectifyMap with R = I, newCameraMatrix = cameraMatrix, 4 distortion coefficients).
void build_undistort_map(int W, int H, const float *K, const float *D, std::vector<uint32_t> &mxy, std::vector<uint16_t> &mf) {
    const double fx = (double)K[0], fy = (double)K[4], u0 = (double)K[2], v0 = (double)K[5];
    const double k1 = (double)D[0], k2 = (double)D[1], p1 = (double)D[2], p2 = (double)D[3];
    mxy.resize((size_t)W * H);
    mf.resize((size_t)W * H);
    int stripe0 = (1 << 12) / (W > 1 ? W : 1);
    stripe0 = stripe0 < 1 ? 1 : stripe0 > H ? H : stripe0;
    auto round_sat = [](double v) -> int {
        const double r = nearbyint(v);
        return r >= 2147483647.0 ? 2147483647 : r <= -2147483648.0 ? (-2147483647 - 1) : (int)r;
    };
    for (int ys = 0; ys < H; ys += stripe0) {
        const int stripe = stripe0 < H - ys ? stripe0 : H - ys;
        // inverse of Ar = [fx 0 cx; 0 fy (cy - ys); 0 0 1] (general 3x3 formula so that a skewed K also works)
        double a[9], ir[9];
        for (int i = 0; i < 9; ++i) a[i] = (double)K[i];
        a[5] = v0 - ys;
        const double det = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
        const double id = 1.0 / det;
        ir[0] = (a[4] * a[8] - a[5] * a[7]) * id; ir[1] = (a[2] * a[7] - a[1] * a[8]) * id; ir[2] = (a[1] * a[5] - a[2] * a[4]) * id;
        ir[3] = (a[5] * a[6] - a[3] * a[8]) * id; ir[4] = (a[0] * a[8] - a[2] * a[6]) * id; ir[5] = (a[2] * a[3] - a[0] * a[5]) * id;
        ir[6] = (a[3] * a[7] - a[4] * a[6]) * id; ir[7] = (a[1] * a[6] - a[0] * a[7]) * id; ir[8] = (a[0] * a[4] - a[1] * a[3]) * id;
        for (int i = 0; i < stripe; ++i) {
            double _x = i * ir[1] + ir[2], _y = i * ir[4] + ir[5], _w = i * ir[7] + ir[8];
            for (int j = 0; j < W; ++j, _x += ir[0], _y += ir[3], _w += ir[6]) {
                const double w = 1. / _w, x = _x * w, y = _y * w;
                const double x2 = x * x, y2 = y * y, r2 = x2 + y2, _2xy = 2 * x * y;
                const double kr = (1 + ((0 * r2 + k2) * r2 + k1) * r2) / (1 + ((0 * r2 + 0) * r2 + 0) * r2);
                const double xd = (x * kr + p1 * _2xy + p2 * (r2 + 2 * x2) + 0 * r2 + 0 * r2 * r2);
                const double yd = (y * kr + p1 * (r2 + 2 * y2) + p2 * _2xy + 0 * r2 + 0 * r2 * r2);
                const double u = fx * 1.0 * xd + u0, v = fy * 1.0 * yd + v0;
                const int iu = round_sat(u * 32.0), iv = round_sat(v * 32.0);
                const size_t o = (size_t)(ys + i) * W + j;
                mxy[o] = ((uint32_t)(uint16_t)(int16_t)(iu >> 5)) | ((uint32_t)(uint16_t)(int16_t)(iv >> 5) << 16);
                mf[o] = (uint16_t)((iv & 31) * 32 + (iu & 31));
            }
        }
    }
}

}  // namespace rdfe
