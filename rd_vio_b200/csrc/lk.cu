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
//     REDUX in 16-bit halves; fma(float(hi), 65536, float(lo)) rounds the exact total once (== int64->float).
//     (OpenCV accumulates in float32; the exact sum is the value that accumulation approximates --
//     measured deviation < 1e-3 px.)
//   * forward and backward pass share ONE copy of the pyramid code (loop over the direction);
//   * template cache (rdfe_set_template_cache): the backward pass of step t builds, per level, exactly the
//     template (I*32, Ix, Iy over the window, A11/A12/A22) that the forward pass of step t+1 needs for a point
//     carried unchanged.  It is stored (packed int16, coalesced 16-byte vectors) keyed by (slot generation,
//     float x, float y); a forward pass that finds its point there loads 3 KB per level instead of staging the
//     I and dI patches and rebuilding (~420 of ~1300 warp instructions per level pass), bit-identical by
//     construction: the template is a pure function of (image, position).
#include <cstdio>
#include "fe_internal.cuh"

namespace rdfe {

template <int WIN> struct LKCfg;
// MIN_CTAS = 5 caps the 21x21 kernel at 96 registers (a few spills in the template build): alone it is 3 % slower than at
// 128 registers / 4 CTAs, but 20 warps per SM and smaller CTAs leave room beside the other streams' kernels:
// +1.1 % frames/s in the pipelined step (A/B twice on one box: 178.9 k -> 180.9 k; 80 registers / 6 CTAs: 178.5 k)
template <> struct LKCfg<21> {
    static constexpr int SEG = 7, NSEG = 3, JW = 48, JH = 32, IH = 22, DW = 28, DH = 22, MARGIN = 5, WARPS = 4, MIN_CTAS = 5;
    static constexpr bool PACKED = false;
};
template <> struct LKCfg<31> {
    static constexpr int SEG = 8, NSEG = 4, JW = 64, JH = 48, IH = 32, DW = 36, DH = 32, MARGIN = 8, WARPS = 4, MIN_CTAS = 3;
    static constexpr bool PACKED = true;
};

struct LKMaps {
    CUtensorMap img[RDFE_MAX_LEVELS];     // J search region: box JW x JH
    CUtensorMap imgT[RDFE_MAX_LEVELS];    // I template patch: box JW x (WIN+1) -- only the rows the template reads
    CUtensorMap der[RDFE_MAX_LEVELS];
};

// Template cache of one launch (all pointers null when the cache is off)
struct LKCache {
    uint4 *data;        // [slot][max_points][levels][ROUNDS*3][32] packed int16 templates
    float4 *A;          // [slot][max_points][levels] (A11, A12, A22, flag bits: 1 = valid)
    float4 *hdr;        // [slot][max_points] (x, y, generation bits, 1)
    unsigned long long *stats;   // [2] forward passes that looked a point up / found it
    int max_points, levels;
    unsigned gen_curr[RDFE_MAX_BATCH], gen_next[RDFE_MAX_BATCH];   // generation of every image of the batch
};

struct LKParams {
    int nlevels;
    int W, H;                       // level 0
    int lw[RDFE_MAX_LEVELS], lh[RDFE_MAX_LEVELS];
    int max_count;
    double eps2;
    float eps2_lo, eps2_hi;         // eps2 * (1 -+ 1e-5): float pre-test brackets (the float sum is within 2e-7 of exact)
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
// Bounded wait: a TMA that never completes (bad descriptor) traps instead of hanging the GPU (the launch then fails with
// an error the host reports; no printf here: its argument set-up costs registers in the hot kernel).
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
        if (spins > (1u << 20)) __trap();
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

// float(exact warp sum) with ONE rounding: both half sums are exact in float (< 2^21), and the fma rounds
// shi * 2^16 + slo once -- the same value as __ll2float_rn(warp_sum_exact(v)), without the 64-bit conversion.
__device__ __forceinline__ float warp_sum_float(int v) {
    const int hi = v >> 16;
    const unsigned lo = (unsigned)v & 0xFFFFu;
    const int shi = __reduce_add_sync(0xffffffffu, hi);
    const unsigned slo = __reduce_add_sync(0xffffffffu, lo);
    return __fmaf_rn((float)shi, 65536.f, (float)(int)slo);
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
// `init` = 256 gives the reference's rounded sample; the iterations pass init = 256 - (Ival << 9), which makes
// the result (sample - Ival) directly: Ival << 9 is a multiple of 2^9, so it passes through the floor shift.
template <int K>
__device__ __forceinline__ int sample(const Pairs &t, const Pairs &b, unsigned wt, unsigned wb, int init = 256) {
    const unsigned pt = (K < 4) ? ((K & 1) ? t.p1 : t.p0) : ((K & 1) ? t.p3 : t.p2);
    const unsigned pb = (K < 4) ? ((K & 1) ? b.p1 : b.p0) : ((K & 1) ? b.p3 : b.p2);
    int acc = init;                                       // 256 = (1 << (W_BITS1-5-1))
    if ((K & 2) == 0) { acc = dp2a_lo_s16_u8(wt, pt, acc); acc = dp2a_lo_s16_u8(wb, pb, acc); }
    else { acc = dp2a_hi_s16_u8(wt, pt, acc); acc = dp2a_hi_s16_u8(wb, pb, acc); }
    return acc >> 9;
}

template <int WIN>
struct __align__(128) WarpSmem {
    using C = LKCfg<WIN>;
    uint8_t ipatch[(C::JW * C::IH + 127) / 128 * 128];
    uint8_t jreg[C::JW * C::JH + 128];
    uint32_t dpatch[C::DW * C::DH + 32];
    uint64_t bar;
};

// cvRound of a float in [0, 2^22): adding 1.5 * 2^23 rounds to nearest-even in the FADD itself (same result as
// cvt.rni), on the FMA/ALU pipes instead of a quarter-rate F2I
#ifdef LK_F2I
__device__ __forceinline__ int rint_small(float v) { return __float2int_rn(v); }
#else
__device__ __forceinline__ int rint_small(float v) { return __float_as_int(__fadd_rn(v, 12582912.f)) - 0x4B400000; }
#endif
__device__ __forceinline__ void q14_weights(float a, float b, int &iw00, int &iw01, int &iw10, int &iw11) {
    iw00 = rint_small((1.f - a) * (1.f - b) * 16384.f);
    iw01 = rint_small(a * (1.f - b) * 16384.f);
    iw10 = rint_small((1.f - a) * b * 16384.f);
    iw11 = 16384 - iw00 - iw01 - iw10;
}

// ---- packed template record of one (lane, round): 12 words = 3 uint4.
//   words [0, SEG):             (Iy << 16) | (Ix & 0xffff) of pixel k
//   words [SEG, SEG + (SEG+1)/2): Ival of pixels 2q, 2q+1 as unsigned halves (0 <= Ival <= 8160)
// One pyramidal LK pass A -> B for the warp's point.  (prevx, prevy): position in A (level 0);
// (qx, qy): initial guess in / result out.  Returns the OpenCV status flag.
// tc_ld / tcA_ld: cached templates of this point in A (null: build them); tc_st / tcA_st: where to store the
// templates built here (null: do not store).
template <int WIN>
__device__ __forceinline__ int lk_pyramid(WarpSmem<WIN> &ws, uint32_t &phase, const LKMaps &maps, const LKParams &P, int slotA, int slotB,
                                          float prevx, float prevy, float &qx, float &qy, const uint4 *tc_ld,
                                          const float4 *tcA_ld, uint4 *tc_st, float4 *tcA_st) {
    using C = LKCfg<WIN>;
    constexpr int SEG = C::SEG, NSEG = C::NSEG, JW = C::JW, JH = C::JH, DW = C::DW;
    constexpr int ITEMS = WIN * NSEG, ROUNDS = (ITEMS + 31) / 32;
    constexpr int win_rt = WIN;
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
        ivalid[r] = (32 * (r + 1) <= ITEMS) || t < ITEMS;     // only the last round can hold idle lanes (compile-time for the others)
        const int tt = ivalid[r] ? t : 0;
        irow[r] = tt / NSEG;
        iseg[r] = tt - irow[r] * NSEG;
    }

#pragma unroll 1
    for (int level = P.nlevels - 1; level >= 0; --level) {
        const int cols = P.lw[level], rows = P.lh[level];
        const float lscale = __int_as_float((127 - level) << 23);      // 2^-level, exactly what (float)(1. / (1 << level)) gives
        float px = prevx * lscale, py = prevy * lscale;
        float nx, ny;
        if (level == P.nlevels - 1) { nx = nxp * lscale; ny = nyp * lscale; }
        else { nx = nxp * 2.f; ny = nyp * 2.f; }
        nxp = nx; nyp = ny;

        if (tcA_st && lane == 0) tcA_st[level] = make_float4(0.f, 0.f, 0.f, 0.f);     // no template of this level (yet)
        px -= half; py -= half;
        const int ipx = (int)floorf(px), ipy = (int)floorf(py);
        if (ipx < -WIN || ipx >= cols || ipy < -WIN || ipy >= rows) {
            if (level == 0) status = 0;
            continue;
        }
        // a cached template of this level?  (warp-uniform: every lane reads the same record)
        float A11 = 0.f, A12 = 0.f, A22 = 0.f;
        bool cached = false;
        if (tcA_ld) {
            const float4 a = __ldg(tcA_ld + level);
            cached = __float_as_uint(a.w) == 1u;
            A11 = a.x; A12 = a.y; A22 = a.z;
        }
        // ---- stage the J search region, and (if the template must be built) the I and dI patches (16-B aligned box starts)
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
            if (cached) {
                mbar_expect_tx(&ws.bar, (uint32_t)(JW * JH));
            } else {
                mbar_expect_tx(&ws.bar, (uint32_t)(JW * JH + JW * C::IH) + 4u * DW * C::DH);
                tma_load_3d(ws.ipatch, &maps.imgT[level], ixa, ipy + win_rt, slotA, &ws.bar);
                tma_load_3d(ws.dpatch, &maps.der[level], dxa, ipy, slotA, &ws.bar);
            }
            tma_load_3d(ws.jreg, &maps.img[level], jx0 + kHaloX, jy0 + win_rt, slotB, &ws.bar);
        }

        // ---- template in registers: Cv = 256 - (I*32 << 9) (so that a J sample started from it IS the difference), Ix, Iy
        int iw00, iw01, iw10, iw11;
        int Cv[ROUNDS][SEG];
        int Ixv[ROUNDS][SEG];                            // PACKED: (Iy << 16) | (Ix & 0xffff)
        int Iyv[C::PACKED ? 1 : ROUNDS][C::PACKED ? 1 : SEG];
        if (cached) {
            // 3 coalesced 16-byte loads per round while the J box is in flight
            const uint4 *src = tc_ld + (size_t)level * (ROUNDS * 3 * 32) + lane;
#pragma unroll
            for (int r = 0; r < ROUNDS; ++r) {
                unsigned w[12];
#pragma unroll
                for (int m = 0; m < 3; ++m) {
                    const uint4 v = __ldg(src + (r * 3 + m) * 32);
                    w[4 * m] = v.x; w[4 * m + 1] = v.y; w[4 * m + 2] = v.z; w[4 * m + 3] = v.w;
                }
#pragma unroll
                for (int k = 0; k < SEG; ++k) {
                    const unsigned pv = w[SEG + (k >> 1)];
                    const int iv = (k & 1) ? (int)(pv >> 16) : (int)(pv & 0xFFFFu);
                    Cv[r][k] = 256 - (iv << 9);
                    if constexpr (C::PACKED) Ixv[r][k] = (int)w[k];
                    else { Ixv[r][k] = (int)(short)(w[k] & 0xFFFFu); Iyv[r][k] = (int)w[k] >> 16; }
                }
            }
            mbar_wait(&ws.bar, phase);
            phase ^= 1u;
        } else {
            mbar_wait(&ws.bar, phase);
            phase ^= 1u;
            q14_weights(px - (float)ipx, py - (float)ipy, iw00, iw01, iw10, iw11);
            int a11 = 0, a12 = 0, a22 = 0;
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
                    const bool ok = ivalid[r] && (SEG * NSEG == WIN || iseg[r] * SEG + k < WIN);   // 21 = 3 x 7: no partial segment
                    if (!ok) { ix = 0; iy = 0; }
                    Cv[r][k] = 256 - (iv[k] << 9);
                    if constexpr (C::PACKED) Ixv[r][k] = (iy << 16) | (ix & 0xFFFF);
                    else { Ixv[r][k] = ix; Iyv[r][k] = iy; }
                    a11 += ix * ix; a12 += ix * iy; a22 += iy * iy;
                    dt = dt1; db = db1;
                }
                if (tc_st) {
                    unsigned w[12];
#pragma unroll
                    for (int k = 0; k < 12; ++k) w[k] = 0u;
#pragma unroll
                    for (int k = 0; k < SEG; ++k) {
                        if constexpr (C::PACKED) w[k] = (unsigned)Ixv[r][k];
                        else w[k] = ((unsigned)Iyv[r][k] << 16) | ((unsigned)Ixv[r][k] & 0xFFFFu);
                    }
#pragma unroll
                    for (int q = 0; q < (SEG + 1) / 2; ++q)
                        w[SEG + q] = (unsigned)iv[2 * q] | ((2 * q + 1 < SEG) ? ((unsigned)iv[2 * q + 1] << 16) : 0u);
                    uint4 *dst = tc_st + (size_t)level * (ROUNDS * 3 * 32) + lane;
#pragma unroll
                    for (int m = 0; m < 3; ++m) dst[(r * 3 + m) * 32] = make_uint4(w[4 * m], w[4 * m + 1], w[4 * m + 2], w[4 * m + 3]);
                }
            }
            A11 = warp_sum_float(a11) * FLT_SCALE;
            A12 = warp_sum_float(a12) * FLT_SCALE;
            A22 = warp_sum_float(a22) * FLT_SCALE;
        }
        float D = A11 * A22 - A12 * A12;
        if (!cached) {
            const float minEig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (float)(2 * WIN * WIN);
            if (minEig < P.min_eig_thr || D < 1.1920928955078125e-07f) {
                if (level == 0) status = 0;
                continue;
            }
            if (tcA_st && lane == 0) tcA_st[level] = make_float4(A11, A12, A22, __uint_as_float(1u));   // passed: reusable
        }
        D = 1.f / D;
        nx -= half; ny -= half;
        float pdx = 0.f, pdy = 0.f;
        int cinx = INT_MIN, ciny = INT_MIN;              // integer origin the cached J words belong to
        Pairs Jt[ROUNDS], Jb[ROUNDS];
#ifndef LK_NO_UNROLL1
#pragma unroll 1
#endif
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
                int df[8];                               // J sample - Ival, straight out of the dot products
                df[0] = sample<0>(Jt[r], Jb[r], wt, wb, Cv[r][0]); df[1] = sample<1>(Jt[r], Jb[r], wt, wb, Cv[r][1]);
                df[2] = sample<2>(Jt[r], Jb[r], wt, wb, Cv[r][2]); df[3] = sample<3>(Jt[r], Jb[r], wt, wb, Cv[r][3]);
                df[4] = sample<4>(Jt[r], Jb[r], wt, wb, Cv[r][4]); df[5] = sample<5>(Jt[r], Jb[r], wt, wb, Cv[r][5]);
                df[6] = sample<6>(Jt[r], Jb[r], wt, wb, Cv[r][6]);
                if constexpr (SEG > 7) df[7] = sample<7>(Jt[r], Jb[r], wt, wb, Cv[r][SEG - 1]);
#pragma unroll
                for (int k = 0; k < SEG; ++k) {
                    if constexpr (C::PACKED) {
                        b1 += df[k] * (int)(short)(Ixv[r][k] & 0xFFFF);
                        b2 += df[k] * (Ixv[r][k] >> 16);
                    } else {
                        b1 += df[k] * Ixv[r][k];
                        b2 += df[k] * Iyv[r][k];
                    }
                }
            }
            const float fb1 = warp_sum_float(b1) * FLT_SCALE;
            const float fb2 = warp_sum_float(b2) * FLT_SCALE;
            const float dx = (A12 * fb2 - A22 * fb1) * D, dy = (A12 * fb1 - A11 * fb2) * D;
            nx += dx; ny += dy;
            nxp = nx + half; nyp = ny + half;
            // |delta|^2 <= eps^2 is evaluated in double by the reference; the float sum is within 2e-7 of the exact
            // value, so outside the +-1e-5 bracket around eps^2 the float test decides and the doubles are skipped
            {
                const float s2 = dx * dx + dy * dy;
                bool conv;
#ifdef LK_DOUBLE_CHECK
                if (false) conv = false;
#else
                if (s2 > P.eps2_hi) conv = false;
#endif
#ifndef LK_DOUBLE_CHECK
                else if (s2 < P.eps2_lo) conv = true;
#endif
                else conv = (double)dx * (double)dx + (double)dy * (double)dy <= P.eps2;
                if (conv) break;
            }
            // fabs((double)(a + b)) < 0.01  <=>  fabsf(a + b) <= 0.01f  (0.01f is the largest float below the double 0.01)
            if (j > 0 && fabsf(dx + pdx) <= 0.01f && fabsf(dy + pdy) <= 0.01f) {
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

// CACHE = false compiles every template-cache path out (the launcher picks it whenever the cache is off)
template <int WIN, int CTAS, bool CACHE>
__global__ void __launch_bounds__(LKCfg<WIN>::WARPS * 32, CTAS)
lk_track_kernel(const __grid_constant__ LKMaps maps, const __grid_constant__ LKParams P, const __grid_constant__ SlotList curr,
                const __grid_constant__ SlotList next, const __grid_constant__ LKCache tc, const double *__restrict__ curr_xy,
                double *__restrict__ next_xy, const int *__restrict__ counts, char *__restrict__ status_out) {
    using C = LKCfg<WIN>;
    constexpr int ROUNDS = (WIN * C::NSEG + 31) / 32;
    __shared__ WarpSmem<WIN> smem[C::WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    const int i = blockIdx.x * C::WARPS + warp;
    const int npts = min(counts[b], P.stride);
    if (i >= npts) return;
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
    const int slotA = curr.v[b], slotB = next.v[b];

    // ---- template cache: is this point (same image generation, same float position) in the records the backward
    // pass of the previous step left for slot A?  Carried points sit at or after their new index (lost ones before
    // them were dropped), so the search walks upwards from i, 32 records per probe.
    const uint4 *tc_ld = nullptr;
    const float4 *tcA_ld = nullptr;
    uint4 *tc_st = nullptr;
    float4 *tcA_st = nullptr;
    const size_t rec = (size_t)tc.levels * (ROUNDS * 3 * 32);       // uint4 per point
    if (CACHE && tc.data) {
        const float4 *hdr = tc.hdr + (size_t)slotA * tc.max_points;
        const unsigned gen = tc.gen_curr[b];
        int found = -1;
        for (int base = i; base < tc.max_points && base < i + 128 && found < 0; base += 32) {
            const int j = base + lane;
            bool hit = false;
            if (j < tc.max_points) {
                const float4 h = __ldg(hdr + j);
                hit = h.x == cx && h.y == cy && __float_as_uint(h.z) == gen && __float_as_uint(h.w) == 1u;
            }
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (m) found = base + __ffs(m) - 1;
        }
        if (lane == 0) {
            atomicAdd(tc.stats, 1ull);
            if (found >= 0) atomicAdd(tc.stats + 1, 1ull);
        }
        if (found >= 0) {
            const size_t e = (size_t)slotA * tc.max_points + found;
            tc_ld = tc.data + e * rec;
            tcA_ld = tc.A + e * tc.levels;
        }
    }

    // forward (dir 0: curr -> next from the prediction) and backward (dir 1: next -> curr seeded with curr) run through
    // ONE inlined copy of the pyramid code: the arguments are selected by `dir`
    int st = 1;
#pragma unroll 1
    for (int dir = 0; dir < 2; ++dir) {
        const int sA = dir ? slotB : slotA, sB = dir ? slotA : slotB;
        const float fromx = dir ? qx : cx, fromy = dir ? qy : cy;
        float gx = dir ? cx : qx, gy = dir ? cy : qy;
        const int rst = lk_pyramid<WIN>(ws, phase, maps, P, sA, sB, fromx, fromy, gx, gy, (CACHE && !dir) ? tc_ld : nullptr,
                                        (CACHE && !dir) ? tcA_ld : nullptr, (CACHE && dir) ? tc_st : nullptr,
                                        (CACHE && dir) ? tcA_st : nullptr);
        if (dir == 0) {
            st = rst;
            qx = gx; qy = gy;
            if (qx < (float)P.border || qx >= (float)(P.W - P.border) || qy < (float)P.border || qy >= (float)(P.H - P.border)) st = 0;
            if (st) {
                const float dx = qx - cx, dy = qy - cy;
                if (sqrt((double)dx * (double)dx + (double)dy * (double)dy) > P.max_jump) st = 0;
            }
            if (!st) break;
            if (CACHE && tc.data && i < tc.max_points) {
                const size_t e = (size_t)slotB * tc.max_points + i;
                tc_st = tc.data + e * rec;
                tcA_st = tc.A + e * tc.levels;
            }
        } else {
            const float dx = cx - gx, dy = cy - gy;
            if (!rst || sqrt((double)dx * (double)dx + (double)dy * (double)dy) > P.max_round_trip) st = 0;
            // the templates just built belong to (image B, position q): next step's forward pass looks them up
            if (tc_st && lane == 0)
                tc.hdr[(size_t)slotB * tc.max_points + i] = make_float4(qx, qy, __uint_as_float(tc.gen_next[b]), __uint_as_float(1u));
        }
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
        maps.imgT[l] = ctx->tm_imgT[l];
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
    P.eps2_lo = (float)(P.eps2 * (1.0 - 1e-5));
    P.eps2_hi = (float)(P.eps2 * (1.0 + 1e-5));
    P.min_eig_thr = (float)p.min_eig_threshold;
    P.border = p.border;
    P.max_jump = (double)(P.H / 4);
    P.max_round_trip = p.max_round_trip;
    P.has_prediction = p.has_prediction;
    P.stride = stride;
    LKCache tc;
    memset(&tc, 0, sizeof tc);
    if (ctx->tc_on && ctx->tc_data) {
        tc.data = ctx->tc_data; tc.A = ctx->tc_A; tc.hdr = ctx->tc_hdr; tc.stats = ctx->tc_stats;
        tc.max_points = ctx->cfg.max_points; tc.levels = pyr.nlevels;
        for (int i = 0; i < curr.n; ++i) {
            tc.gen_curr[i] = ctx->slot_gen[curr.v[i]];
            tc.gen_next[i] = ctx->slot_gen[next.v[i]];
        }
    }
    if (pyr.win == 21) {
        dim3 grid((stride + LKCfg<21>::WARPS - 1) / LKCfg<21>::WARPS, curr.n);
        if (tc.data)
            RDFE_LAUNCH(ctx, K_LK, (lk_track_kernel<21, LKCfg<21>::MIN_CTAS, true><<<grid, LKCfg<21>::WARPS * 32, 0, ctx->ls>>>(
                                       maps, P, curr, next, tc, d_curr_xy, d_next_xy, d_counts, d_status)));
        else
            RDFE_LAUNCH(ctx, K_LK, (lk_track_kernel<21, LKCfg<21>::MIN_CTAS, false><<<grid, LKCfg<21>::WARPS * 32, 0, ctx->ls>>>(
                                       maps, P, curr, next, tc, d_curr_xy, d_next_xy, d_counts, d_status)));
    } else if (pyr.win == 31) {
        dim3 grid((stride + LKCfg<31>::WARPS - 1) / LKCfg<31>::WARPS, curr.n);
        if (tc.data)
            RDFE_LAUNCH(ctx, K_LK, (lk_track_kernel<31, LKCfg<31>::MIN_CTAS, true><<<grid, LKCfg<31>::WARPS * 32, 0, ctx->ls>>>(
                                       maps, P, curr, next, tc, d_curr_xy, d_next_xy, d_counts, d_status)));
        else
            RDFE_LAUNCH(ctx, K_LK, (lk_track_kernel<31, LKCfg<31>::MIN_CTAS, false><<<grid, LKCfg<31>::WARPS * 32, 0, ctx->ls>>>(
                                       maps, P, curr, next, tc, d_curr_xy, d_next_xy, d_counts, d_status)));
    } else {
        set_error("LK window %d unsupported (21 or 31)", pyr.win);
        return RDFE_ERR_UNSUPPORTED;
    }
    return 1;
}

// ---- caller-side prediction for keypoints that never leave the device (frame.cpp:82-93): pixel -> unit bearing through
// K^-1, rotation by the pre-integrated gyro increment, back through K_next = one 3x3 homography per stream, float64.
__global__ void predict_rotation_kernel(const double *__restrict__ H, const double *__restrict__ curr_xy, const int *__restrict__ counts,
                                        int stride, double *__restrict__ pred_xy) {
    const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= min(counts[b], stride)) return;
    const double *h = H + 9 * b;
    const size_t k = ((size_t)b * stride + i) * 2;
    const double x = curr_xy[k], y = curr_xy[k + 1];
    const double X = h[0] * x + h[1] * y + h[2], Y = h[3] * x + h[4] * y + h[5], Z = h[6] * x + h[7] * y + h[8];
    pred_xy[k] = X / Z;
    pred_xy[k + 1] = Y / Z;
}

int launch_predict_rotation(rdfe_ctx *ctx, int n, const double *d_H, const double *d_curr_xy, const int *d_counts, int stride,
                            double *d_pred_xy) {
    dim3 grid((stride + 127) / 128, n);
    RDFE_LAUNCH(ctx, K_PREDICT, (predict_rotation_kernel<<<grid, 128, 0, ctx->ls>>>(d_H, d_curr_xy, d_counts, stride, d_pred_xy)));
    return 1;
}

// bytes of template-cache storage one (slot, point) needs for this context's window / level count
size_t lk_cache_record_bytes(int win, int nlevels) {
    const int nseg = win == 21 ? LKCfg<21>::NSEG : LKCfg<31>::NSEG;
    const int rounds = (win * nseg + 31) / 32;
    return (size_t)nlevels * rounds * 3 * 32 * sizeof(uint4);
}

}  // namespace rdfe
