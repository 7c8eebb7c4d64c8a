// harris.cu -- K3a: Harris response + 3x3 non-maximum suppression + candidate emission.
// Replaces the dense half of cv::GFTTDetector::detect (reference call site:
// OpenCvImage::detect_keypoints, src/rdvio_extra/src/opencv_image.cpp:44; detector
// parameters OpenCvImage::gftt, :184-188 -- the 5th argument `true` selects HARRIS, k=0.04).
//
// cv::cornerHarris(img, blockSize=3, ksize=3, k) arithmetic ("plain" order):
//   scaled Sobel (k0 = 1/3060, k1 = 2/3060) in float32, products in float32, 3x3 box sums
//   accumulated in float64 (exact for these magnitudes, hence order independent), rounded
//   to float32, R = (A*C - B*B) - (k*(A+C))*(A+C), each op rounded (the library is built
//   with -fmad=false).  harris_fma=1 instead reproduces the FMA placement of OpenCV's
//   AVX2/AVX-512 dispatched filters (reported, not the parity target).
//
// goodFeaturesToTrack's "threshold, dilate, compare" (App. A5) is equivalent to:
//   candidate <=> R > thr  &&  R >= every in-image 8-neighbour  &&  not on the 1-px frame.
// thr = max(R)*q needs the frame maximum, so this kernel emits every POSITIVE 3x3 local
// maximum as a 64-bit key (float bits << 32 | y*W+x) plus the per-frame maximum; the
// select kernel applies thr.  The response map itself never goes to HBM on the hot path.
#include "fe_internal.cuh"

namespace rdfe {

constexpr int HT_W = 32, HT_H = 16;                 // output tile
constexpr int HP_W = HT_W + 6, HP_H = HT_H + 6;     // pixels   (apron 3)
constexpr int HG_W = HT_W + 4, HG_H = HT_H + 4;     // products (apron 2)
constexpr int HR_W = HT_W + 2, HR_H = HT_H + 2;     // response (apron 1)

template <bool kFma>
__global__ void __launch_bounds__(256)
harris_nms_kernel(Pyramid pyr, SlotList slots, float k, DetectScratch det, float *__restrict__ response) {
    __shared__ float pf[HP_H][HP_W];
    __shared__ float pa[HG_H][HG_W], pb[HG_H][HG_W], pc[HG_H][HG_W];
    __shared__ double ha[HG_H][HR_W], hbb[HG_H][HR_W], hc[HG_H][HR_W];
    __shared__ float R[HR_H][HR_W];
    __shared__ unsigned s_wcount[8];
    __shared__ unsigned s_base;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.z, slot = slots.v[b];
    const LevelGeom g = pyr.lv[0];
    const int W = g.w, H = g.h;
    const uint8_t *src = pyr.image_origin(0, slot);
    const int ox = blockIdx.x * HT_W, oy = blockIdx.y * HT_H;

    const double sc = 1.0 / (4.0 * 3.0 * 255.0);
    const float k0 = (float)sc, k1 = (float)(2.0 * sc);

    // pixels, REFLECT_101 at the image border.  Positions whose products/responses lie outside
    // the image are never consumed un-reflected (see the index mapping below).
    for (int i = tid; i < HP_H * HP_W; i += 256) {
        const int r = i / HP_W, c = i - r * HP_W;
        const int sy = reflect101(oy - 3 + r, H), sx = reflect101(ox - 3 + c, W);
        pf[r][c] = (float)src[(size_t)sy * g.ipitch + sx];
    }
    __syncthreads();
    // gradient products at (oy-2+r, ox-2+c).  boxFilter reflects the PRODUCT maps at the
    // border, so a product position outside the image must hold the product of the
    // reflected position: recompute it there from reflected pixel coordinates.
    for (int i = tid; i < HG_H * HG_W; i += 256) {
        const int r = i / HG_W, c = i - r * HG_W;
        const int gy_ = oy - 2 + r, gx_ = ox - 2 + c;
        float p00, p01, p02, p10, p12, p20, p21, p22;
        if (gy_ >= 0 && gy_ < H && gx_ >= 0 && gx_ < W) {
            p00 = pf[r][c]; p01 = pf[r][c + 1]; p02 = pf[r][c + 2];
            p10 = pf[r + 1][c]; p12 = pf[r + 1][c + 2];
            p20 = pf[r + 2][c]; p21 = pf[r + 2][c + 1]; p22 = pf[r + 2][c + 2];
        } else {
            const int cy = reflect101(gy_, H), cx = reflect101(gx_, W);
            const int y0 = reflect101(cy - 1, H), y2 = reflect101(cy + 1, H);
            const int x0 = reflect101(cx - 1, W), x2 = reflect101(cx + 1, W);
            const uint8_t *r0 = src + (size_t)y0 * g.ipitch, *r1 = src + (size_t)cy * g.ipitch,
                          *r2 = src + (size_t)y2 * g.ipitch;
            p00 = r0[x0]; p01 = r0[cx]; p02 = r0[x2];
            p10 = r1[x0]; p12 = r1[x2];
            p20 = r2[x0]; p21 = r2[cx]; p22 = r2[x2];
        }
        const float d0 = p02 - p00, d1 = p12 - p10, d2 = p22 - p20;
        float gx, s0, s2;
        if (!kFma) {
            gx = k1 * d1 + k0 * (d0 + d2);
            s0 = ((k0 * p00) + k1 * p01) + k0 * p02;
            s2 = ((k0 * p20) + k1 * p21) + k0 * p22;
        } else {
            gx = __fmaf_rn(k0, d0 + d2, k1 * d1);
            s0 = __fmaf_rn(k0, p02, __fmaf_rn(k1, p01, k0 * p00));
            s2 = __fmaf_rn(k0, p22, __fmaf_rn(k1, p21, k0 * p20));
        }
        const float gy = s2 - s0;
        pa[r][c] = gx * gx;
        pb[r][c] = gx * gy;
        pc[r][c] = gy * gy;
    }
    __syncthreads();
    // horizontal 3-sums in float64 at (oy-2+r, ox-1+c)
    for (int i = tid; i < HG_H * HR_W; i += 256) {
        const int r = i / HR_W, c = i - r * HR_W;
        ha[r][c] = ((double)pa[r][c] + (double)pa[r][c + 1]) + (double)pa[r][c + 2];
        hbb[r][c] = ((double)pb[r][c] + (double)pb[r][c + 1]) + (double)pb[r][c + 2];
        hc[r][c] = ((double)pc[r][c] + (double)pc[r][c + 1]) + (double)pc[r][c + 2];
    }
    __syncthreads();
    // vertical 3-sums -> response at (oy-1+r, ox-1+c)
    for (int i = tid; i < HR_H * HR_W; i += 256) {
        const int r = i / HR_W, c = i - r * HR_W;
        const float A = (float)((ha[r][c] + ha[r + 1][c]) + ha[r + 2][c]);
        const float B = (float)((hbb[r][c] + hbb[r + 1][c]) + hbb[r + 2][c]);
        const float C = (float)((hc[r][c] + hc[r + 1][c]) + hc[r + 2][c]);
        float v;
        if (!kFma) v = (A * C - B * B) - (k * (A + C)) * (A + C);
        else v = (A * C - B * B) - k * ((A + C) * (A + C));
        const int y = oy - 1 + r, x = ox - 1 + c;
        // dilate ignores pixels outside the image: make them lose every comparison
        R[r][c] = (y >= 0 && y < H && x >= 0 && x < W) ? v : -INFINITY;
    }
    __syncthreads();

    // NMS + emission: 2 pixels per thread (HT_W*HT_H = 512)
    unsigned long long keys[2];
    int nk = 0;
    float tmax = 0.0f;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int i = tid + j * 256;
        const int r = i / HT_W, c = i - r * HT_W;
        const int y = oy + r, x = ox + c;
        if (y < H && x < W) {
            const float v = R[r + 1][c + 1];
            if (response) response[((size_t)b * H + y) * W + x] = v;
            tmax = fmaxf(tmax, v);
            if (v > 0.0f && y >= 1 && y < H - 1 && x >= 1 && x < W - 1) {
                const float m = fmaxf(fmaxf(fmaxf(R[r][c], R[r][c + 1]), fmaxf(R[r][c + 2], R[r + 1][c])),
                                      fmaxf(fmaxf(R[r + 1][c + 2], R[r + 2][c]), fmaxf(R[r + 2][c + 1], R[r + 2][c + 2])));
                if (v >= m)
                    keys[nk++] = ((unsigned long long)__float_as_uint(v) << 32) | (unsigned)(y * W + x);
            }
        }
    }
    // frame maximum (positive floats order like unsigned ints)
    unsigned mb = __reduce_max_sync(0xffffffffu, __float_as_uint(tmax));
    if (lane == 0 && mb) atomicMax(&det.frame_max[b], mb);
    // CTA-aggregated append
    unsigned wtot = __reduce_add_sync(0xffffffffu, (unsigned)nk);
    unsigned wpre = (unsigned)nk;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned t = __shfl_up_sync(0xffffffffu, wpre, d);
        if (lane >= d) wpre += t;
    }
    wpre -= (unsigned)nk;
    if (lane == 0) s_wcount[warp] = wtot;
    __syncthreads();
    if (tid == 0) {
        unsigned tot = 0;
        for (int w = 0; w < 8; ++w) { unsigned t = s_wcount[w]; s_wcount[w] = tot; tot += t; }
        s_base = tot ? atomicAdd(&det.cand_count[b], tot) : 0u;
    }
    __syncthreads();
    unsigned pos = s_base + s_wcount[warp] + wpre;
    for (int j = 0; j < nk; ++j, ++pos) {
        if (pos < det.cand_cap) det.cand[(size_t)b * det.cand_cap + pos] = keys[j];
        else atomicExch(det.overflow, 1u);
    }
}

__global__ void detect_reset_kernel(DetectScratch det, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { det.cand_count[i] = 0u; det.frame_max[i] = 0u; }
}

int launch_harris_candidates(rdfe_ctx *ctx, const SlotList &slots, const rdfe_detect_params &p, float *d_response) {
    const LevelGeom &g = ctx->pyr.lv[0];
    detect_reset_kernel<<<1, RDFE_MAX_BATCH, 0, ctx->stream>>>(ctx->det, slots.n);
    dim3 grid((g.w + HT_W - 1) / HT_W, (g.h + HT_H - 1) / HT_H, slots.n);
    if (p.harris_fma)
        RDFE_LAUNCH(ctx, K_HARRIS, (harris_nms_kernel<true><<<grid, 256, 0, ctx->stream>>>(ctx->pyr, slots, (float)p.harris_k, ctx->det, d_response)));
    else
        RDFE_LAUNCH(ctx, K_HARRIS, (harris_nms_kernel<false><<<grid, 256, 0, ctx->stream>>>(ctx->pyr, slots, (float)p.harris_k, ctx->det, d_response)));
    return 2;
}

}  // namespace rdfe
