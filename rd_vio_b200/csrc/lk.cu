// lk.cu -- K4: forward-backward pyramidal Lucas-Kanade with the reference's gating, one launch.
// Replaces OpenCvImage::track_keypoints (src/rdvio_extra/src/opencv_image.cpp:75-154):
//   calcOpticalFlowPyrLK(curr->next, OPTFLOW_USE_INITIAL_FLOW, win, maxLevel, (COUNT+EPS,30,0.01))
//   -> status &= 20-px border on the forward result (:101-106) -> jump > rows/4 (:107-113)
//   -> calcOpticalFlowPyrLK(next->curr) seeded with curr (:117-125) -> round trip > 0.5 px (:127-134).
// Per-point arithmetic follows cv::detail::LKTrackerInvoker: Q14 bilinear
// weights, int16 template (I*32, Ix, Iy), float32 2x2 solve, eps^2 / oscillation termination.
//
// B200 mapping
//   * one warp per keypoint, all pyramid levels and both directions inside the warp;
//   * per level the warp's elected lane issues three TMA tile loads (cp.async.bulk.tensor.3d,
//     coordinates {x, y, slot}) completing on a per-warp mbarrier: the (win+1)^2 template patch
//     of I, its Scharr derivative patch (uint32 = int16 pair; the zero halo of the reference's
//     derivative buffers is TMA out-of-bounds fill) and a J search region (window + margin);
//     TMA needs the innermost box coordinate 16-byte aligned (measured on B200: an unaligned
//     start raises "illegal instruction"), so boxes start at the aligned-down column and are
//     15 bytes (3 derivative elements) wider; the residual offset is applied in shared memory;
//   * a lane owns (row, 7- or 8-pixel segment) items of the window; the template lives in
//     registers for the whole level; the J bytes a lane needs are fetched from shared memory as
//     aligned 32-bit words only when the window's INTEGER origin moves, and re-paired with
//     funnel shifts so that each bilinear sample is two IDP.2A (u16 x u8 dot products);
//   * sums (A11,A12,A22,b1,b2) are accumulated as exact integers per lane and reduced with
//     REDUX; one int64->float conversion.  (OpenCV accumulates in float32; the exact sum is the
//     value that accumulation approximates -- measured deviation < 1e-3 px.)
#include <cstdio>
#include "fe_internal.cuh"

namespace rdfe {

template <int WIN> struct LKCfg;
template <> struct LKCfg<21> {
    static constexpr int SEG = 7, NSEG = 3, JW = 48, JH = 32, DW = 28, DH = 22, MARGIN = 5, WARPS = 4, MIN_CTAS = 4;
    static constexpr bool PACKED = false;
};
template <> struct LKCfg<31> {
    static constexpr int SEG = 8, NSEG = 4, JW = 64, JH = 48, DW = 36, DH = 32, MARGIN = 8, WARPS = 4, MIN_CTAS = 3;
    static constexpr bool PACKED = true;
};

struct LKMaps {
    CUtensorMap img[RDFE_MAX_LEVELS];
    CUtensorMap der[RDFE_MAX_LEVELS];
};

struct LKParams {
    int nlevels;
    int W, H;                       // level 0
    int lw[RDFE_MAX_LEVELS], lh[RDFE_MAX_LEVELS];
    int max_count;
    double eps2;
    float min_eig_thr;
    int border;
    double max_jump;                // rows / 4 (integer division)
    double max_round_trip;
    int has_prediction;
    int stride;
};

// ------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a TMA that never completes (bad descriptor) traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
        if (spins > (1u << 20)) {
            if ((threadIdx.x & 31) == 0) printf("rdfe lk: TMA wait timed out (block %d,%d warp %d parity %u)\n", blockIdx.x, blockIdx.y, threadIdx.x >> 5, parity);
            __trap();
        }
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, int x, int y, int z, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
        : "memory");
}

// exact warp sum of int32 partials (|v| < 2^31): split so neither REDUX can overflow
__device__ __forceinline__ long long warp_sum_exact(int v) {
    const int hi = v >> 16;
    const unsigned lo = (unsigned)v & 0xFFFFu;
    const int shi = __reduce_add_sync(0xffffffffu, hi);
    const unsigned slo = __reduce_add_sync(0xffffffffu, lo);
    return ((long long)shi << 16) + (long long)slo;
}

// 9 consecutive bytes starting at byte offset `o` of a shared-memory buffer, as four
// "pair words": P0 = bytes 0..3, P1 = 1..4, P2 = 4..7, P3 = 5..8.
struct Pairs { unsigned p0, p1, p2, p3; };
__device__ __forceinline__ Pairs load_pairs(const uint8_t *base, int o) {
    const unsigned *w = reinterpret_cast<const unsigned *>(base + (o & ~3));
    const unsigned s = (unsigned)(o & 3) * 8u;
    const unsigned w0 = w[0], w1 = w[1], w2 = w[2];
    const unsigned b0 = __funnelshift_r(w0, w1, s), b1 = __funnelshift_r(w1, w2, s), b2 = w2 >> s;
    Pairs r;
    r.p0 = b0;
    r.p1 = __funnelshift_r(b0, b1, 8);
    r.p2 = b1;
    r.p3 = __funnelshift_r(b1, b2, 8);
    return r;
}
// Signed 16-bit weights x unsigned bytes.  The fourth Q14 weight is 2^14 minus the three rounded ones and is
// -1 when those round up past 2^14, so the weights must be read as signed halves (the CUDA intrinsics only
// offer same-signedness operands; PTX allows .s32.u32).
__device__ __forceinline__ int dp2a_lo_s16_u8(unsigned w, unsigned px, int c) {
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(w), "r"(px), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_hi_s16_u8(unsigned w, unsigned px, int c) {
    int d;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(w), "r"(px), "r"(c));
    return d;
}
// bilinear sample k (0..7) of a segment: top/bottom pair words, packed Q14 weights
template <int K>
__device__ __forceinline__ int sample(const Pairs &t, const Pairs &b, unsigned wt, unsigned wb) {
    const unsigned pt = (K < 4) ? ((K & 1) ? t.p1 : t.p0) : ((K & 1) ? t.p3 : t.p2);
    const unsigned pb = (K < 4) ? ((K & 1) ? b.p1 : b.p0) : ((K & 1) ? b.p3 : b.p2);
    int acc = 256;                                        // + (1 << (W_BITS1-5-1))
    if ((K & 2) == 0) { acc = dp2a_lo_s16_u8(wt, pt, acc); acc = dp2a_lo_s16_u8(wb, pb, acc); }
    else { acc = dp2a_hi_s16_u8(wt, pt, acc); acc = dp2a_hi_s16_u8(wb, pb, acc); }
    return acc >> 9;
}

template <int WIN>
struct __align__(128) WarpSmem {
    using C = LKCfg<WIN>;
    uint8_t ipatch[C::JW * C::JH];
    uint8_t jreg[C::JW * C::JH + 128];
    uint32_t dpatch[C::DW * C::DH + 32];
    uint64_t bar;
};

__device__ __forceinline__ void q14_weights(float a, float b, int &iw00, int &iw01, int &iw10, int &iw11) {
    iw00 = __float2int_rn((1.f - a) * (1.f - b) * 16384.f);
    iw01 = __float2int_rn(a * (1.f - b) * 16384.f);
    iw10 = __float2int_rn((1.f - a) * b * 16384.f);
    iw11 = 16384 - iw00 - iw01 - iw10;
}

// One pyramidal LK pass A -> B for the warp's point.  (prevx, prevy): position in A (level 0);
// (qx, qy): initial guess in / result out.  Returns the OpenCV status flag.
template <int WIN>
__device__ int lk_pyramid(WarpSmem<WIN> &ws, uint32_t &phase, const LKMaps &maps, const LKParams &P, int slotA, int slotB,
                          float prevx, float prevy, float &qx, float &qy, int win_rt) {
    using C = LKCfg<WIN>;
    constexpr int SEG = C::SEG, NSEG = C::NSEG, JW = C::JW, JH = C::JH, DW = C::DW;
    constexpr int ITEMS = WIN * NSEG, ROUNDS = (ITEMS + 31) / 32;
    const int lane = threadIdx.x & 31;
    const float half = (float)(WIN - 1) * 0.5f;
    const float FLT_SCALE = 1.f / (float)(1 << 20);
    int status = 1;
    float nxp = qx, nyp = qy;                        // nextPts[ptidx]

    int irow[ROUNDS], iseg[ROUNDS];
    bool ivalid[ROUNDS];
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        const int t = lane + 32 * r;
        ivalid[r] = t < ITEMS;
        const int tt = ivalid[r] ? t : 0;
        irow[r] = tt / NSEG;
        iseg[r] = tt - irow[r] * NSEG;
    }

    for (int level = P.nlevels - 1; level >= 0; --level) {
        const int cols = P.lw[level], rows = P.lh[level];
        const float lscale = (float)(1. / (double)(1 << level));
        float px = prevx * lscale, py = prevy * lscale;
        float nx, ny;
        if (level == P.nlevels - 1) { nx = nxp * lscale; ny = nyp * lscale; }
        else { nx = nxp * 2.f; ny = nyp * 2.f; }
        nxp = nx; nyp = ny;

        px -= half; py -= half;
        const int ipx = (int)floorf(px), ipy = (int)floorf(py);
        if (ipx < -WIN || ipx >= cols || ipy < -WIN || ipy >= rows) {
            if (level == 0) status = 0;
            continue;
        }
        // ---- stage I patch, dI patch and the J search region (16-B aligned box starts)
        const int ixg = ipx + kHaloX, ixa = ixg & ~15, ioff = ixg - ixa;
        const int dxa = ipx & ~3, doff = ipx - dxa;
        int jx0, jy0;                                    // interior coordinates of the staged J region origin
        {
            const int jxg = ((int)floorf(nx - half) - C::MARGIN) + kHaloX;
            jx0 = (jxg & ~15) - kHaloX;
            jy0 = (int)floorf(ny - half) - C::MARGIN;
        }
        __syncwarp();
        if (lane == 0) {
            mbar_expect_tx(&ws.bar, 2u * JW * JH + 4u * DW * C::DH);
            tma_load_3d(ws.ipatch, &maps.img[level], ixa, ipy + win_rt, slotA, &ws.bar);
            tma_load_3d(ws.dpatch, &maps.der[level], dxa, ipy, slotA, &ws.bar);
            tma_load_3d(ws.jreg, &maps.img[level], jx0 + kHaloX, jy0 + win_rt, slotB, &ws.bar);
        }
        mbar_wait(&ws.bar, phase);
        phase ^= 1u;

        // ---- template: Ival (I*32), Ix, Iy in registers; exact integer A sums
        int iw00, iw01, iw10, iw11;
        q14_weights(px - (float)ipx, py - (float)ipy, iw00, iw01, iw10, iw11);
        int Ival[ROUNDS][SEG];
        int Ixv[ROUNDS][SEG];                            // PACKED: (Iy << 16) | (Ix & 0xffff)
        int Iyv[C::PACKED ? 1 : ROUNDS][C::PACKED ? 1 : SEG];
        int a11 = 0, a12 = 0, a22 = 0;
        {
            const unsigned wt = (unsigned)iw00 | ((unsigned)iw01 << 16), wb = (unsigned)iw10 | ((unsigned)iw11 << 16);  // iw11 may be -1
#pragma unroll
            for (int r = 0; r < ROUNDS; ++r) {
                const int o = irow[r] * JW + ioff + iseg[r] * SEG;
                const Pairs t = load_pairs(ws.ipatch, o), b = load_pairs(ws.ipatch, o + JW);
                int iv[8];
                iv[0] = sample<0>(t, b, wt, wb); iv[1] = sample<1>(t, b, wt, wb);
                iv[2] = sample<2>(t, b, wt, wb); iv[3] = sample<3>(t, b, wt, wb);
                iv[4] = sample<4>(t, b, wt, wb); iv[5] = sample<5>(t, b, wt, wb);
                iv[6] = sample<6>(t, b, wt, wb); iv[7] = sample<7>(t, b, wt, wb);
                const uint32_t *d0 = ws.dpatch + irow[r] * DW + doff + iseg[r] * SEG;
                const uint32_t *d1 = d0 + DW;
                uint32_t dt = d0[0], db = d1[0];
#pragma unroll
                for (int k = 0; k < SEG; ++k) {
                    const uint32_t dt1 = d0[k + 1], db1 = d1[k + 1];
                    const int x00 = (short)(dt & 0xFFFFu), y00 = (int)dt >> 16;
                    const int x01 = (short)(dt1 & 0xFFFFu), y01 = (int)dt1 >> 16;
                    const int x10 = (short)(db & 0xFFFFu), y10 = (int)db >> 16;
                    const int x11 = (short)(db1 & 0xFFFFu), y11 = (int)db1 >> 16;
                    int ix = (x00 * iw00 + x01 * iw01 + x10 * iw10 + x11 * iw11 + (1 << 13)) >> 14;
                    int iy = (y00 * iw00 + y01 * iw01 + y10 * iw10 + y11 * iw11 + (1 << 13)) >> 14;
                    const bool ok = ivalid[r] && (iseg[r] * SEG + k < WIN);
                    if (!ok) { ix = 0; iy = 0; }
                    Ival[r][k] = iv[k];
                    if constexpr (C::PACKED) Ixv[r][k] = (iy << 16) | (ix & 0xFFFF);
                    else { Ixv[r][k] = ix; Iyv[r][k] = iy; }
                    a11 += ix * ix; a12 += ix * iy; a22 += iy * iy;
                    dt = dt1; db = db1;
                }
            }
        }
        const float A11 = __ll2float_rn(warp_sum_exact(a11)) * FLT_SCALE;
        const float A12 = __ll2float_rn(warp_sum_exact(a12)) * FLT_SCALE;
        const float A22 = __ll2float_rn(warp_sum_exact(a22)) * FLT_SCALE;
        float D = A11 * A22 - A12 * A12;
        const float minEig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (float)(2 * WIN * WIN);
        if (minEig < P.min_eig_thr || D < 1.1920928955078125e-07f) {
            if (level == 0) status = 0;
            continue;
        }
        D = 1.f / D;
        nx -= half; ny -= half;
        float pdx = 0.f, pdy = 0.f;
        int cinx = INT_MIN, ciny = INT_MIN;              // integer origin the cached J words belong to
        Pairs Jt[ROUNDS], Jb[ROUNDS];
        for (int j = 0; j < P.max_count; ++j) {
            const int inx = (int)floorf(nx), iny = (int)floorf(ny);
            if (inx < -WIN || inx >= cols || iny < -WIN || iny >= rows) {
                if (level == 0) status = 0;
                break;
            }
            if (inx != cinx || iny != ciny) {
                if (inx < jx0 || inx > jx0 + (JW - (WIN + 1)) || iny < jy0 || iny > jy0 + (JH - (WIN + 1))) {
                    // the window left the staged region: restage around the current position
                    jx0 = (((inx - C::MARGIN) + kHaloX) & ~15) - kHaloX;
                    jy0 = iny - C::MARGIN;
                    __syncwarp();
                    if (lane == 0) {
                        mbar_expect_tx(&ws.bar, (uint32_t)(JW * JH));
                        tma_load_3d(ws.jreg, &maps.img[level], jx0 + kHaloX, jy0 + win_rt, slotB, &ws.bar);
                    }
                    mbar_wait(&ws.bar, phase);
                    phase ^= 1u;
                }
                const int ob = (iny - jy0) * JW + (inx - jx0);
#pragma unroll
                for (int r = 0; r < ROUNDS; ++r) {
                    const int o = ob + irow[r] * JW + iseg[r] * SEG;
                    Jt[r] = load_pairs(ws.jreg, o);
                    Jb[r] = load_pairs(ws.jreg, o + JW);
                }
                cinx = inx; ciny = iny;
            }
            q14_weights(nx - (float)inx, ny - (float)iny, iw00, iw01, iw10, iw11);
            const unsigned wt = (unsigned)iw00 | ((unsigned)iw01 << 16), wb = (unsigned)iw10 | ((unsigned)iw11 << 16);  // iw11 may be -1
            int b1 = 0, b2 = 0;
#pragma unroll
            for (int r = 0; r < ROUNDS; ++r) {
                int jv[8];
                jv[0] = sample<0>(Jt[r], Jb[r], wt, wb); jv[1] = sample<1>(Jt[r], Jb[r], wt, wb);
                jv[2] = sample<2>(Jt[r], Jb[r], wt, wb); jv[3] = sample<3>(Jt[r], Jb[r], wt, wb);
                jv[4] = sample<4>(Jt[r], Jb[r], wt, wb); jv[5] = sample<5>(Jt[r], Jb[r], wt, wb);
                jv[6] = sample<6>(Jt[r], Jb[r], wt, wb); jv[7] = sample<7>(Jt[r], Jb[r], wt, wb);
#pragma unroll
                for (int k = 0; k < SEG; ++k) {
                    const int diff = jv[k] - Ival[r][k];
                    if constexpr (C::PACKED) {
                        b1 += diff * (int)(short)(Ixv[r][k] & 0xFFFF);
                        b2 += diff * (Ixv[r][k] >> 16);
                    } else {
                        b1 += diff * Ixv[r][k];
                        b2 += diff * Iyv[r][k];
                    }
                }
            }
            const float fb1 = __ll2float_rn(warp_sum_exact(b1)) * FLT_SCALE;
            const float fb2 = __ll2float_rn(warp_sum_exact(b2)) * FLT_SCALE;
            const float dx = (A12 * fb2 - A22 * fb1) * D, dy = (A12 * fb1 - A11 * fb2) * D;
            nx += dx; ny += dy;
            nxp = nx + half; nyp = ny + half;
            if ((double)dx * (double)dx + (double)dy * (double)dy <= P.eps2) break;
            if (j > 0 && fabs((double)(dx + pdx)) < 0.01 && fabs((double)(dy + pdy)) < 0.01) {
                nxp -= dx * 0.5f; nyp -= dy * 0.5f;
                break;
            }
            pdx = dx; pdy = dy;
        }
        if (status && level == 0) {
            const int fx = (int)floorf(nxp - half), fy = (int)floorf(nyp - half);
            if (fx < -WIN || fx >= cols || fy < -WIN || fy >= rows) status = 0;
        }
    }
    qx = nxp; qy = nyp;
    return status;
}

template <int WIN>
__global__ void __launch_bounds__(LKCfg<WIN>::WARPS * 32, LKCfg<WIN>::MIN_CTAS)
lk_track_kernel(const __grid_constant__ LKMaps maps, const __grid_constant__ LKParams P, SlotList curr, SlotList next,
                const double *__restrict__ curr_xy, double *__restrict__ next_xy, const int *__restrict__ counts,
                char *__restrict__ status_out) {
    using C = LKCfg<WIN>;
    __shared__ WarpSmem<WIN> smem[C::WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    const int i = blockIdx.x * C::WARPS + warp;
    if (i >= min(counts[b], P.stride)) return;
    WarpSmem<WIN> &ws = smem[warp];
    if (lane == 0) {
        mbar_init(&ws.bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t phase = 0;
    const size_t idx = (size_t)b * P.stride + i;
    const float cx = (float)curr_xy[2 * idx], cy = (float)curr_xy[2 * idx + 1];
    float qx = cx, qy = cy;
    if (P.has_prediction) { qx = (float)next_xy[2 * idx]; qy = (float)next_xy[2 * idx + 1]; }

    int st = lk_pyramid<WIN>(ws, phase, maps, P, curr.v[b], next.v[b], cx, cy, qx, qy, WIN);
    if (qx < (float)P.border || qx >= (float)(P.W - P.border) || qy < (float)P.border || qy >= (float)(P.H - P.border)) st = 0;
    if (st) {
        const float dx = qx - cx, dy = qy - cy;
        if (sqrt((double)dx * (double)dx + (double)dy * (double)dy) > P.max_jump) st = 0;
    }
    if (st) {
        float rx = cx, ry = cy;
        const int rst = lk_pyramid<WIN>(ws, phase, maps, P, next.v[b], curr.v[b], qx, qy, rx, ry, WIN);
        const float dx = cx - rx, dy = cy - ry;
        if (!rst || sqrt((double)dx * (double)dx + (double)dy * (double)dy) > P.max_round_trip) st = 0;
    }
    if (lane == 0) {
        status_out[idx] = (char)st;
        if (st) { next_xy[2 * idx] = (double)qx; next_xy[2 * idx + 1] = (double)qy; }
    }
}

int launch_lk(rdfe_ctx *ctx, const SlotList &curr, const SlotList &next, const rdfe_track_params &p,
              const double *d_curr_xy, double *d_next_xy, const int *d_counts, int stride, char *d_status) {
    LKMaps maps;
    LKParams P;
    const Pyramid &pyr = ctx->pyr;
    for (int l = 0; l < pyr.nlevels; ++l) {
        maps.img[l] = ctx->tm_img[l];
        maps.der[l] = ctx->tm_der[l];
        P.lw[l] = pyr.lv[l].w;
        P.lh[l] = pyr.lv[l].h;
    }
    P.nlevels = pyr.nlevels;
    P.W = pyr.lv[0].w; P.H = pyr.lv[0].h;
    int mc = p.max_count; mc = mc < 0 ? 0 : mc > 100 ? 100 : mc;           // calcOpticalFlowPyrLK clamps
    double eps = p.epsilon; eps = eps < 0 ? 0 : eps > 10 ? 10 : eps;
    P.max_count = mc;
    P.eps2 = eps * eps;
    P.min_eig_thr = (float)p.min_eig_threshold;
    P.border = p.border;
    P.max_jump = (double)(P.H / 4);
    P.max_round_trip = p.max_round_trip;
    P.has_prediction = p.has_prediction;
    P.stride = stride;
    if (pyr.win == 21) {
        dim3 grid((stride + LKCfg<21>::WARPS - 1) / LKCfg<21>::WARPS, curr.n);
        RDFE_LAUNCH(ctx, K_LK, (lk_track_kernel<21><<<grid, LKCfg<21>::WARPS * 32, 0, ctx->ls>>>(maps, P, curr, next, d_curr_xy,
                                                                                                      d_next_xy, d_counts, d_status)));
    } else if (pyr.win == 31) {
        dim3 grid((stride + LKCfg<31>::WARPS - 1) / LKCfg<31>::WARPS, curr.n);
        RDFE_LAUNCH(ctx, K_LK, (lk_track_kernel<31><<<grid, LKCfg<31>::WARPS * 32, 0, ctx->ls>>>(maps, P, curr, next, d_curr_xy,
                                                                                                      d_next_xy, d_counts, d_status)));
    } else {
        set_error("LK window %d unsupported (21 or 31)", pyr.win);
        return RDFE_ERR_UNSUPPORTED;
    }
    return 1;
}

}  // namespace rdfe
