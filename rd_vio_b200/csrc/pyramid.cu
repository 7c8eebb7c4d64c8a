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
//                    REFLECT_101 halo (store4_with_halo), so no separate halo pass exists.
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
constexpr int PD_ROWS = 8;             // output rows per warp strip

__global__ void __launch_bounds__(PW_WARPS * 32)
pyrdown_kernel(Pyramid pyr, SlotList slots, int l, int tiles_x, int n_items) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int item = blockIdx.x * PW_WARPS + warp;
    if (item >= n_items) return;
    const int slot = slots.v[blockIdx.y];
    const LevelGeom gs = pyr.lv[l], gd = pyr.lv[l + 1];
    const uint8_t *src = pyr.image_origin(l, slot);
    uint8_t *dst = pyr.image_origin(l + 1, slot);
    const int j0 = (item % tiles_x) * 128 + 4 * lane;       // first of this lane's 4 output columns
    const int i0 = (item / tiles_x) * PD_ROWS;              // first output row of the strip
    const int rows = min(PD_ROWS, gd.h - i0);
    const int sx = 2 * j0;                                   // input column of this lane's 8-byte word
    const bool active = j0 < gd.w;
    // input columns sx-2 .. sx+8 are needed; all inside [-2, w+1] for active lanes (halo >= 2 px)
    const bool ld_ok = (sx + 7 <= gs.w + pyr.win - 1);       // also the first inactive lane: its taps feed lane-1
    const bool edge_l = (lane == 0), edge_r = (lane == 31);

    int h0[4], h1[4], h2[4], h3[4];                          // horizontal sums of input rows y-4 .. y-1
#pragma unroll
    for (int k = 0; k < 4; ++k) h0[k] = h1[k] = h2[k] = h3[k] = 0;

    const int ystart = 2 * i0 - 2, nsteps = 2 * rows + 3;    // input rows 2*i0-2 .. 2*(i0+rows-1)+2
    // software pipelining: the words of step s+1 are requested before step s is consumed
    const bool el = edge_l && active, er = edge_r && active && (sx + 8 <= gs.w + pyr.win - 4);
    auto load_row = [&](int y, uint2 &w, unsigned &xl, unsigned &xr) {
        const uint8_t *row = src + (ptrdiff_t)y * gs.ipitch;
        w = make_uint2(0u, 0u);
        xl = 0u; xr = 0u;
        if (ld_ok) w = *reinterpret_cast<const uint2 *>(row + sx);
        if (el) xl = *reinterpret_cast<const unsigned *>(row + sx - 4);
        if (er) xr = *reinterpret_cast<const unsigned *>(row + sx + 8);
    };
    uint2 wn; unsigned xln, xrn;
    load_row(ystart, wn, xln, xrn);
    for (int s = 0; s < nsteps; ++s) {
        const uint2 w = wn;
        const unsigned xl = xln, xr = xrn;
        if (s + 1 < nsteps) load_row(ystart + s + 1, wn, xln, xrn);
        unsigned wl = __shfl_up_sync(0xffffffffu, w.y, 1);   // columns sx-4 .. sx-1
        unsigned wr = __shfl_down_sync(0xffffffffu, w.x, 1); // columns sx+8 .. sx+11
        if (el) wl = xl;
        if (er) wr = xr;
        // taps: p[-2..8] relative to sx
        unsigned p[11];
        p[0] = bfe8(wl, 2); p[1] = bfe8(wl, 3);
        p[2] = bfe8(w.x, 0); p[3] = bfe8(w.x, 1); p[4] = bfe8(w.x, 2); p[5] = bfe8(w.x, 3);
        p[6] = bfe8(w.y, 0); p[7] = bfe8(w.y, 1); p[8] = bfe8(w.y, 2); p[9] = bfe8(w.y, 3);
        p[10] = bfe8(wr, 0);
        int hn[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            hn[k] = (int)(p[2 * k] + p[2 * k + 4] + 4u * (p[2 * k + 1] + p[2 * k + 3]) + 6u * p[2 * k + 2]);
        if (s >= 4 && (s & 1) == 0) {
            // input row y = 2i+2  =>  output row i = (y-2)/2
            const int i = i0 + ((s - 4) >> 1);
            unsigned out = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int v = h0[k] + hn[k] + 4 * (h1[k] + h3[k]) + 6 * h2[k];
                out |= (unsigned)((v + 128) >> 8) << (8 * k);
            }
            if (active) store4_with_halo(dst, gd.ipitch, gd.w, gd.h, pyr.win, j0, i, out);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) { h0[k] = h1[k]; h1[k] = h2[k]; h2[k] = h3[k]; h3[k] = hn[k]; }
    }
}

// ----------------------------------------------------------------- Scharr
constexpr int SC_ROWS = 16;            // output rows per warp strip

struct ItemTable {
    int first[RDFE_MAX_LEVELS + 1];   // first flattened work item of each level
    int tiles_x[RDFE_MAX_LEVELS];
};

__global__ void __launch_bounds__(PW_WARPS * 32)
scharr_kernel(Pyramid pyr, SlotList slots, ItemTable tt) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int l = blockIdx.z;                                // level
    const int item = blockIdx.x * PW_WARPS + warp;
    if (item >= tt.first[l]) return;                         // tt.first[l] = work items of this level
    const int slot = slots.v[blockIdx.y];
    // level geometry into registers once (the loop below only bumps pointers)
    const int gw = pyr.lv[l].w, gh = pyr.lv[l].h, ipitch = pyr.lv[l].ipitch, dpitch = pyr.lv[l].dpitch;
    const int tiles_x = tt.tiles_x[l];
    const int c0 = (item % tiles_x) * 128 + 4 * lane;
    const int y0 = (item / tiles_x) * SC_ROWS;
    const int rows = min(SC_ROWS, gh - y0);
    const bool active = c0 < gw;
    const bool ld_ok = (c0 + 3 <= gw + pyr.win - 1);          // inside the halo (also the first inactive lane)
    const bool el = (lane == 0) && active, er = (lane == 31) && active;
    const bool full = active && (c0 + 3 < gw);
    const uint8_t *rp = pyr.image_origin(l, slot) + (ptrdiff_t)(y0 - 1) * ipitch + c0;     // row y0-1
    uint8_t *op = reinterpret_cast<uint8_t *>(pyr.deriv_origin(l, slot)) + (size_t)y0 * dpitch + 4 * (size_t)c0;

    // rows y-2, y-1: p(x+1)-p(x-1) and 3p(x-1)+10p(x)+3p(x+1)
    int d1A[4] = {0, 0, 0, 0}, d1B[4] = {0, 0, 0, 0}, s2A[4] = {0, 0, 0, 0}, s2B[4] = {0, 0, 0, 0};
    unsigned wn = 0, xln = 0, xrn = 0;
    if (ld_ok) wn = *reinterpret_cast<const unsigned *>(rp);
    if (el) xln = *reinterpret_cast<const unsigned *>(rp - 4);
    if (er) xrn = *reinterpret_cast<const unsigned *>(rp + 4);
    const int nsteps = rows + 2;
#pragma unroll 3
    for (int s = 0; s < nsteps; ++s) {
        const unsigned w = wn, xl = xln, xr = xrn;
        rp += ipitch;
        if (s + 1 < nsteps) {                                 // software pipelining: next row in flight
            if (ld_ok) wn = *reinterpret_cast<const unsigned *>(rp);
            if (el) xln = *reinterpret_cast<const unsigned *>(rp - 4);
            if (er) xrn = *reinterpret_cast<const unsigned *>(rp + 4);
        }
        unsigned wl = __shfl_up_sync(0xffffffffu, w, 1), wr = __shfl_down_sync(0xffffffffu, w, 1);
        if (el) wl = xl;
        if (er) wr = xr;
        int p[6];
        p[0] = (int)(wl >> 24);
        p[1] = (int)(w & 0xFFu); p[2] = (int)__byte_perm(w, 0u, 0x4441u); p[3] = (int)__byte_perm(w, 0u, 0x4442u); p[4] = (int)(w >> 24);
        p[5] = (int)(wr & 0xFFu);
        int d1N[4], s2N[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            d1N[k] = p[k + 2] - p[k];
            s2N[k] = 3 * (p[k] + p[k + 2]) + 10 * p[k + 1];
        }
        if (s >= 2) {
            unsigned o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int gx = 3 * (d1A[k] + d1N[k]) + 10 * d1B[k];
                const int gy = s2N[k] - s2A[k];
                o[k] = __byte_perm((unsigned)gx, (unsigned)gy, 0x5410u);      // (dx, dy) int16 pair
            }
            if (full) {
                *reinterpret_cast<uint4 *>(op) = make_uint4(o[0], o[1], o[2], o[3]);
            } else if (active) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (c0 + k < gw) reinterpret_cast<unsigned *>(op)[k] = o[k];
            }
            op += dpitch;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) { d1A[k] = d1B[k]; d1B[k] = d1N[k]; s2A[k] = s2B[k]; s2B[k] = s2N[k]; }
    }
}

int launch_pyramid(rdfe_ctx *ctx, const SlotList &slots) {
    const Pyramid &pyr = ctx->pyr;
    int launches = 0;
    for (int l = 0; l + 1 < pyr.nlevels; ++l) {
        const LevelGeom &gd = pyr.lv[l + 1];
        const int tiles_x = (gd.w + 127) / 128, strips = (gd.h + PD_ROWS - 1) / PD_ROWS;
        const int n_items = tiles_x * strips;
        dim3 grid((n_items + PW_WARPS - 1) / PW_WARPS, slots.n);
        RDFE_LAUNCH(ctx, K_PYRDOWN, (pyrdown_kernel<<<grid, PW_WARPS * 32, 0, ctx->ls>>>(pyr, slots, l, tiles_x, n_items)));
        ++launches;
    }
    ItemTable tt;
    int max_items = 0;
    for (int l = 0; l < pyr.nlevels; ++l) {
        const LevelGeom &g = pyr.lv[l];
        tt.tiles_x[l] = (g.w + 127) / 128;
        tt.first[l] = tt.tiles_x[l] * ((g.h + SC_ROWS - 1) / SC_ROWS);      // work items (warps) of level l
        max_items = tt.first[l] > max_items ? tt.first[l] : max_items;
    }
    RDFE_LAUNCH(ctx, K_SCHARR, (scharr_kernel<<<dim3((max_items + PW_WARPS - 1) / PW_WARPS, slots.n, pyr.nlevels), PW_WARPS * 32, 0, ctx->ls>>>(pyr, slots, tt)));
    return launches + 1;
}

}  // namespace rdfe
