// pyramid.cu -- K2: what cv::buildOpticalFlowPyramid(image, pyr, Size(win,win), maxLevel,
// withDerivatives=true) produces (reference call site: OpenCvImage::preprocess,
// src/rdvio_extra/src/opencv_image.cpp:159-160; arithmetic: SURVEY.md App. A2-A3).
//
//   pyrdown_kernel   level l -> l+1: separable [1 4 6 4 1]^2, (sum+128)>>8.  A warp owns a strip of
//                    128 output columns and walks down the INPUT rows: each lane reads one aligned
//                    8-byte word per row (its 4 output pixels' 8 centre/odd taps), takes the 3 taps it
//                    lacks from lane+-1 by shuffle, keeps the horizontal sums of the last 5 input rows
//                    in registers and emits one output row every second step as a 4-byte store
//                    (128 B per warp).  The producer also writes the output level's win-px
//                    REFLECT_101 halo (store4_rows_coop / store4_with_halo), so no separate halo pass exists.
//   scharr_kernel    all levels in one launch, same rolling scheme (4 pixels per lane, 3-row window):
//                    un-normalised 3x3 Scharr, (dx,dy) int16 pairs, one 16-byte store per lane per row.
// Both read their border taps from the materialised REFLECT_101 halo of the source level -- exactly the
// border rule of pyrDown / calcSharrDeriv -- so the inner loops contain no index reflection.
// All integer arithmetic: bit-exact by construction.
#include "fe_internal.cuh"

namespace rdfe {

constexpr int PW_WARPS = 4;            // warps per CTA (independent strips)

__device__ __forceinline__ unsigned bfe8(unsigned w, int k) { return (w >> (8 * k)) & 0xFFu; }

// ---------------------------------------------------------------- pyrDown
constexpr int PD_ROWS = 8;             // output rows per warp strip at full batches (adaptive_strip_rows)

__global__ void __launch_bounds__(PW_WARPS * 32)
pyrdown_kernel(Pyramid pyr, SlotList slots, int l, int tiles_x, int n_items, int pd_rows) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wi = blockIdx.x * PW_WARPS + warp;            // flattened (image, strip): no idle warps per image
    if (wi >= n_items * slots.n) return;
    const int bimg = wi / n_items, item = wi - bimg * n_items;
    const int slot = slots.v[bimg];
    const LevelGeom gs = pyr.lv[l], gd = pyr.lv[l + 1];
    const uint8_t *src = pyr.image_origin(l, slot);
    uint8_t *dst = pyr.image_origin(l + 1, slot);
    const int j0 = (item % tiles_x) * 128 + 4 * lane;       // first of this lane's 4 output columns
    const int i0 = (item / tiles_x) * pd_rows;              // first output row of the strip
    const int rows = min(pd_rows, gd.h - i0);
    const int sx = 2 * j0;                                   // input column of this lane's 8-byte word
    const bool active = j0 < gd.w;
    // input columns sx-2 .. sx+8 are needed; all inside [-2, w+1] for active lanes (halo >= 2 px)
    const bool ld_ok = (sx + 7 <= gs.w + pyr.win - 1);       // also the first inactive lane: its taps feed lane-1
    const bool edge_l = (lane == 0), edge_r = (lane == 31);
    // halo of the output level: warp-cooperative word stores when the row geometry allows it
    const bool coop = coop_halo_ok(gd.w, pyr.win);
    const int nh = (pyr.win + 3) >> 2;
    const bool tile_l = (item % tiles_x) == 0, tile_r = (item % tiles_x) == tiles_x - 1;

    // Two outputs per 32-bit register (16-bit fields): h*01 = outputs 0 and 1 of the lane, h*23 = outputs 2, 3.
    // Horizontal sums are <= 16 * 255, the vertical sum + 128 <= 65408: every field stays within 16 bits, so the
    // packed adds/multiplies never carry across fields (bit-exact).
    unsigned h0a = 0, h1a = 0, h2a = 0, h3a = 0, h0b = 0, h1b = 0, h2b = 0, h3b = 0;   // input rows y-4 .. y-1

    const int ystart = 2 * i0 - 2, nsteps = 2 * rows + 3;    // input rows 2*i0-2 .. 2*(i0+rows-1)+2
    // software pipelining: the words of step s+1 are requested before step s is consumed.  The read after the
    // last step is input row 2*(i0+rows)+1 <= h+2: inside the halo (win >= 3 rows), so it needs no guard.
    const bool el = edge_l && active, er = edge_r && active && (sx + 8 <= gs.w + pyr.win - 4);
    const uint8_t *row = src + (ptrdiff_t)ystart * gs.ipitch + sx;
    auto load_row = [&](uint2 &w, unsigned &xl, unsigned &xr) {
        w = make_uint2(0u, 0u);
        xl = 0u; xr = 0u;
        if (ld_ok) w = *reinterpret_cast<const uint2 *>(row);
        if (el) xl = *reinterpret_cast<const unsigned *>(row - 4);
        if (er) xr = *reinterpret_cast<const unsigned *>(row + 8);
        row += gs.ipitch;
    };
    uint2 wn; unsigned xln, xrn;
    load_row(wn, xln, xrn);
    for (int s = 0; s < nsteps; ++s) {
        const uint2 w = wn;
        const unsigned xl = xln, xr = xrn;
        load_row(wn, xln, xrn);
        unsigned wl = __shfl_up_sync(0xffffffffu, w.y, 1);   // columns sx-4 .. sx-1
        unsigned wr = __shfl_down_sync(0xffffffffu, w.x, 1); // columns sx+8 .. sx+11
        if (el) wl = xl;
        if (er) wr = xr;
        // taps p[-2..8] relative to sx: p0 p1 = wl bytes 2 3, p2..p5 = w.x, p6..p9 = w.y, p10 = wr byte 0
        const unsigned Ex = w.x & 0x00FF00FFu, Ox = __byte_perm(w.x, 0u, 0x4341u);   // (p2,p4) (p3,p5)
        const unsigned Ey = w.y & 0x00FF00FFu, Oy = __byte_perm(w.y, 0u, 0x4341u);   // (p6,p8) (p7,p9)
        const unsigned p02 = __byte_perm(wl, Ex, 0x5452u), p13 = __byte_perm(wl, Ox, 0x5453u);
        const unsigned p46 = __byte_perm(Ex, Ey, 0x5432u), p57 = __byte_perm(Ox, Oy, 0x5432u);
        const unsigned p8a = __byte_perm(Ey, wr, 0x3432u);                           // (p8,p10)
        const unsigned hna = p02 + p46 + 4u * (p13 + Ox) + 6u * Ex;                  // outputs 0, 1
        const unsigned hnb = p46 + p8a + 4u * (p57 + Oy) + 6u * Ey;                  // outputs 2, 3
        if (s >= 4 && (s & 1) == 0) {
            // input row y = 2i+2  =>  output row i = (y-2)/2
            const int i = i0 + ((s - 4) >> 1);
            const unsigned va = h0a + hna + 4u * (h1a + h3a) + 6u * h2a + 0x00800080u;
            const unsigned vb = h0b + hnb + 4u * (h1b + h3b) + 6u * h2b + 0x00800080u;
            const unsigned out = __byte_perm(va, vb, 0x7531u);                       // (v + 128) >> 8 of each field
            if (coop) store4_rows_coop(dst, gd.ipitch, gd.w, gd.h, pyr.win, nh, j0 >> 2, i, active, out, tile_l, tile_r);
            else if (active) store4_with_halo(dst, gd.ipitch, gd.w, gd.h, pyr.win, j0, i, out);
        }
        h0a = h1a; h1a = h2a; h2a = h3a; h3a = hna;
        h0b = h1b; h1b = h2b; h2b = h3b; h3b = hnb;
    }
}

// ----------------------------------------------------------------- Scharr
constexpr int SC_ROWS = 16;            // output rows per warp strip at full batches (adaptive_strip_rows)

struct ItemTable {
    int first[RDFE_MAX_LEVELS + 1];   // work items (warps) of each level
    int tiles_x[RDFE_MAX_LEVELS];
    int rows[RDFE_MAX_LEVELS];        // strip height of each level
    int base[RDFE_MAX_LEVELS + 1];    // first flattened (level, image, strip) index of each level
};

__global__ void __launch_bounds__(PW_WARPS * 32)
scharr_kernel(Pyramid pyr, SlotList slots, ItemTable tt) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // flattened (level, image, strip) work items: tt.first[l] = items of level l per image, tt.base[l] = first
    // flattened index of level l
    int wi = blockIdx.x * PW_WARPS + warp;
    if (wi >= tt.base[pyr.nlevels]) return;
    int l = 0;
    while (wi >= tt.base[l + 1]) ++l;
    wi -= tt.base[l];
    const int bimg = wi / tt.first[l], item = wi - bimg * tt.first[l];
    const int slot = slots.v[bimg];
    // level geometry into registers once (the loop below only bumps pointers)
    const int gw = pyr.lv[l].w, gh = pyr.lv[l].h, ipitch = pyr.lv[l].ipitch, dpitch = pyr.lv[l].dpitch;
    const int tiles_x = tt.tiles_x[l];
    const int c0 = (item % tiles_x) * 128 + 4 * lane;
    const int y0 = (item / tiles_x) * tt.rows[l];
    const int rows = min(tt.rows[l], gh - y0);
    const bool active = c0 < gw;
    const bool ld_ok = (c0 + 3 <= gw + pyr.win - 1);          // inside the halo (also the first inactive lane)
    const bool el = (lane == 0) && active, er = (lane == 31) && active;
    const bool full = active && (c0 + 3 < gw);
    const uint8_t *rp = pyr.image_origin(l, slot) + (ptrdiff_t)(y0 - 1) * ipitch + c0;     // row y0-1
    uint8_t *op = reinterpret_cast<uint8_t *>(pyr.deriv_origin(l, slot)) + (size_t)y0 * dpitch + 4 * (size_t)c0;

    // Two pixels per 32-bit register (16-bit fields): A-registers hold the lane's pixels 0 and 2, B-registers
    // pixels 1 and 3.  Per row and pixel: d = p(x+1) - p(x-1) kept with a bias of 2048 (so the vertical 3/10/3
    // sum carries 16 * 2048 = 2^15 and never leaves its field), s = 3p(x-1) + 10p(x) + 3p(x+1) <= 4080.
    // dx = 3(d[y-1] + d[y+1]) + 10 d[y], dy = s[y+1] - s[y-1] + 2^15; flipping bit 15 of a field removes the bias
    // without carries and leaves the int16 two's complement value.  All fields stay within 16 bits, so the packed
    // 32-bit adds/multiplies never carry across fields: bit-exact.
    constexpr unsigned BD = 0x08000800u, B15 = 0x80008000u;
    unsigned dA0 = 0, dA1 = 0, dB0 = 0, dB1 = 0, sA0 = 0, sA1 = 0, sB0 = 0, sB1 = 0;   // rows y-2 (…0) and y-1 (…1)
    unsigned wn = 0, xln = 0, xrn = 0;
    if (ld_ok) wn = *reinterpret_cast<const unsigned *>(rp);
    if (el) xln = *reinterpret_cast<const unsigned *>(rp - 4);
    if (er) xrn = *reinterpret_cast<const unsigned *>(rp + 4);
    const int nsteps = rows + 2;
#pragma unroll 3
    for (int s = 0; s < nsteps; ++s) {
        const unsigned w = wn, xl = xln, xr = xrn;
        rp += ipitch;
        // software pipelining: next row in flight.  After the last step this reads row y0 + rows + 1 <= h + 1,
        // which lies inside the halo (win >= 2 rows): no guard needed.
        if (ld_ok) wn = *reinterpret_cast<const unsigned *>(rp);
        if (el) xln = *reinterpret_cast<const unsigned *>(rp - 4);
        if (er) xrn = *reinterpret_cast<const unsigned *>(rp + 4);
        unsigned wl = __shfl_up_sync(0xffffffffu, w, 1), wr = __shfl_down_sync(0xffffffffu, w, 1);
        if (el) wl = xl;
        if (er) wr = xr;
        const unsigned A = w & 0x00FF00FFu;                    // (v0, v2)
        const unsigned B = __byte_perm(w, 0u, 0x4341u);        // (v1, v3)
        const unsigned LA = __byte_perm(wl, B, 0x5453u);       // (vL, v1): left neighbours of A
        const unsigned RB = __byte_perm(A, wr, 0x3432u);       // (v2, vR): right neighbours of B
        const unsigned dAn = B + BD - LA, dBn = RB + BD - A;
        const unsigned sAn = 3u * (LA + B) + 10u * A, sBn = 3u * (A + RB) + 10u * B;
        if (s >= 2) {
            const unsigned xa = (3u * (dA0 + dAn) + 10u * dA1) ^ B15, xb = (3u * (dB0 + dBn) + 10u * dB1) ^ B15;
            const unsigned ya = (sAn + B15 - sA0) ^ B15, yb = (sBn + B15 - sB0) ^ B15;
            // (dx, dy) int16 pairs of pixels 0..3
            const unsigned o0 = __byte_perm(xa, ya, 0x5410u), o2 = __byte_perm(xa, ya, 0x7632u);
            const unsigned o1 = __byte_perm(xb, yb, 0x5410u), o3 = __byte_perm(xb, yb, 0x7632u);
            if (full) {
                *reinterpret_cast<uint4 *>(op) = make_uint4(o0, o1, o2, o3);
            } else if (active) {
                const unsigned o[4] = {o0, o1, o2, o3};
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (c0 + k < gw) reinterpret_cast<unsigned *>(op)[k] = o[k];
            }
            op += dpitch;
        }
        dA0 = dA1; dA1 = dAn; dB0 = dB1; dB1 = dBn;
        sA0 = sA1; sA1 = sAn; sB0 = sB1; sB1 = sBn;
    }
}

int launch_pyrdowns(rdfe_ctx *ctx, const SlotList &slots) {
    const Pyramid &pyr = ctx->pyr;
    int launches = 0;
    for (int l = 0; l + 1 < pyr.nlevels; ++l) {
        const LevelGeom &gd = pyr.lv[l + 1];
        const int tiles_x = (gd.w + 127) / 128;
        const int pd_rows = adaptive_strip_rows(gd.h, tiles_x * slots.n, 2, PD_ROWS);
        const int strips = (gd.h + pd_rows - 1) / pd_rows;
        const int n_items = tiles_x * strips;
        dim3 grid((n_items * slots.n + PW_WARPS - 1) / PW_WARPS);
        RDFE_LAUNCH(ctx, K_PYRDOWN, (pyrdown_kernel<<<grid, PW_WARPS * 32, 0, ctx->ls>>>(pyr, slots, l, tiles_x, n_items, pd_rows)));
        ++launches;
    }
    return launches;
}

// Scharr derivatives of levels [lo, hi) in one launch (levels outside the range get no work items)
int launch_scharr_levels(rdfe_ctx *ctx, const SlotList &slots, int lo, int hi) {
    const Pyramid &pyr = ctx->pyr;
    ItemTable tt;
    tt.base[0] = 0;
    for (int l = 0; l < pyr.nlevels; ++l) {
        const LevelGeom &g = pyr.lv[l];
        tt.tiles_x[l] = (g.w + 127) / 128;
        tt.rows[l] = adaptive_strip_rows(g.h, tt.tiles_x[l] * slots.n, 4, SC_ROWS);
        tt.first[l] = (l >= lo && l < hi) ? tt.tiles_x[l] * ((g.h + tt.rows[l] - 1) / tt.rows[l]) : 0;   // work items (warps) of level l per image
        tt.base[l + 1] = tt.base[l] + tt.first[l] * slots.n;
    }
    if (tt.base[pyr.nlevels] == 0) return 0;
    RDFE_LAUNCH(ctx, K_SCHARR, (scharr_kernel<<<(tt.base[pyr.nlevels] + PW_WARPS - 1) / PW_WARPS, PW_WARPS * 32, 0, ctx->ls>>>(pyr, slots, tt)));
    return 1;
}

int launch_pyramid(rdfe_ctx *ctx, const SlotList &slots) {
    const int a = launch_pyrdowns(ctx, slots);
    return a + launch_scharr_levels(ctx, slots, 0, ctx->pyr.nlevels);
}

}  // namespace rdfe
