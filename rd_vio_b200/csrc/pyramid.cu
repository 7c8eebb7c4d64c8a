// pyramid.cu -- K2: what cv::buildOpticalFlowPyramid(image, pyr, Size(win,win), maxLevel,
// withDerivatives=true) produces (reference call site: OpenCvImage::preprocess,
// src/rdvio_extra/src/opencv_image.cpp:159-160; arithmetic: SURVEY.md App. A2-A3).
//
//   pyrdown_kernel   level l -> l+1, separable [1 4 6 4 1]^2, (sum+128)>>8, REFLECT_101.
//   scharr_kernel    all levels in one launch: un-normalised 3x3 Scharr, (dx,dy) int16 pairs.
//   halo_kernel      all levels in one launch: materialises the win-px REFLECT_101 halo
//                    around every image plane (the LK window reads it; the derivative
//                    planes need no halo: TMA out-of-bounds fill supplies their zeros).
// All integer arithmetic: bit-exact by construction.
#include "fe_internal.cuh"

namespace rdfe {

// ---------------------------------------------------------------- pyrDown
constexpr int PD_TW = 64, PD_TH = 16;             // output tile
constexpr int PD_IW = 2 * PD_TW + 3, PD_IH = 2 * PD_TH + 3;

__global__ void __launch_bounds__(256)
pyrdown_kernel(Pyramid pyr, SlotList slots, int l) {
    __shared__ uint8_t in[PD_IH][PD_IW + 1];
    __shared__ uint16_t hb[PD_IH][PD_TW];
    const int tid = threadIdx.x;
    const int slot = slots.v[blockIdx.z];
    const LevelGeom gs = pyr.lv[l], gd = pyr.lv[l + 1];
    const uint8_t *src = pyr.image_origin(l, slot);
    uint8_t *dst = pyr.image_origin(l + 1, slot);
    const int ox = blockIdx.x * PD_TW, oy = blockIdx.y * PD_TH;
    const int sx0 = 2 * ox - 2, sy0 = 2 * oy - 2;

    for (int i = tid; i < PD_IH * PD_IW; i += 256) {
        const int r = i / PD_IW, c = i - r * PD_IW;
        const int sy = reflect101(sy0 + r, gs.h), sx = reflect101(sx0 + c, gs.w);
        in[r][c] = src[(size_t)sy * gs.ipitch + sx];
    }
    __syncthreads();
    for (int i = tid; i < PD_IH * PD_TW; i += 256) {
        const int r = i / PD_TW, c = i - r * PD_TW;
        const uint8_t *p = &in[r][2 * c];
        hb[r][c] = (uint16_t)(p[0] + 4 * p[1] + 6 * p[2] + 4 * p[3] + p[4]);
    }
    __syncthreads();
    for (int i = tid; i < PD_TH * PD_TW; i += 256) {
        const int r = i / PD_TW, c = i - r * PD_TW;
        const int x = ox + c, y = oy + r;
        if (x < gd.w && y < gd.h) {
            const int v = hb[2 * r][c] + 4 * hb[2 * r + 1][c] + 6 * hb[2 * r + 2][c] + 4 * hb[2 * r + 3][c] +
                          hb[2 * r + 4][c];
            dst[(size_t)y * gd.ipitch + x] = (uint8_t)((v + 128) >> 8);
        }
    }
}

// ----------------------------------------------------------------- Scharr
constexpr int SC_TW = 64, SC_TH = 16;

struct TileTable {
    int first[RDFE_MAX_LEVELS + 1];   // first flattened tile index of each level
    int tiles_x[RDFE_MAX_LEVELS];
};

__global__ void __launch_bounds__(256)
scharr_kernel(Pyramid pyr, SlotList slots, TileTable tt) {
    __shared__ uint8_t in[SC_TH + 2][SC_TW + 2 + 2];
    const int tid = threadIdx.x;
    int l = 0;
    while (l + 1 < pyr.nlevels && (int)blockIdx.x >= tt.first[l + 1]) ++l;
    const int t = blockIdx.x - tt.first[l];
    const int ox = (t % tt.tiles_x[l]) * SC_TW, oy = (t / tt.tiles_x[l]) * SC_TH;
    const int slot = slots.v[blockIdx.y];
    const LevelGeom g = pyr.lv[l];
    const uint8_t *src = pyr.image_origin(l, slot);
    int16_t *dst = pyr.deriv_origin(l, slot);

    for (int i = tid; i < (SC_TH + 2) * (SC_TW + 2); i += 256) {
        const int r = i / (SC_TW + 2), c = i - r * (SC_TW + 2);
        const int sy = reflect101(oy - 1 + r, g.h), sx = reflect101(ox - 1 + c, g.w);
        in[r][c] = src[(size_t)sy * g.ipitch + sx];
    }
    __syncthreads();
    for (int i = tid; i < SC_TH * SC_TW; i += 256) {
        const int r = i / SC_TW, c = i - r * SC_TW;
        const int x = ox + c, y = oy + r;
        if (x < g.w && y < g.h) {
            const int a00 = in[r][c], a01 = in[r][c + 1], a02 = in[r][c + 2];
            const int a10 = in[r + 1][c], a12 = in[r + 1][c + 2];
            const int a20 = in[r + 2][c], a21 = in[r + 2][c + 1], a22 = in[r + 2][c + 2];
            const int gx = 3 * (a02 - a00) + 10 * (a12 - a10) + 3 * (a22 - a20);
            const int gy = 3 * (a20 - a00) + 10 * (a21 - a01) + 3 * (a22 - a02);
            short2 o;
            o.x = (short)gx;
            o.y = (short)gy;
            *reinterpret_cast<short2 *>(reinterpret_cast<uint8_t *>(dst) + (size_t)y * g.dpitch + 4 * (size_t)x) = o;
        }
    }
}

// ------------------------------------------------------------------- halo
struct HaloTable {
    int first[RDFE_MAX_LEVELS + 1];   // first flattened halo-pixel block (256 px) of each level
};

__global__ void __launch_bounds__(256)
halo_kernel(Pyramid pyr, SlotList slots, HaloTable ht) {
    int l = 0;
    while (l + 1 < pyr.nlevels && (int)blockIdx.x >= ht.first[l + 1]) ++l;
    const int idx = (blockIdx.x - ht.first[l]) * 256 + threadIdx.x;
    const int slot = slots.v[blockIdx.y];
    const LevelGeom g = pyr.lv[l];
    const int win = pyr.win;
    const int fw = g.w + 2 * win;                       // full (haloed) width
    const int n_tb = win * fw;                          // top band, bottom band
    const int n_lr = g.h * win;                         // left band, right band
    int hx, hy;                                         // halo pixel in interior coordinates
    if (idx < n_tb) { hy = -win + idx / fw; hx = -win + idx % fw; }
    else if (idx < 2 * n_tb) { const int j = idx - n_tb; hy = g.h + j / fw; hx = -win + j % fw; }
    else if (idx < 2 * n_tb + n_lr) { const int j = idx - 2 * n_tb; hy = j / win; hx = -win + j % win; }
    else if (idx < 2 * n_tb + 2 * n_lr) { const int j = idx - 2 * n_tb - n_lr; hy = j / win; hx = g.w + j % win; }
    else return;
    uint8_t *org = pyr.image_origin(l, slot);
    org[(ptrdiff_t)hy * g.ipitch + hx] = org[(size_t)reflect101(hy, g.h) * g.ipitch + reflect101(hx, g.w)];
}

int launch_pyramid(rdfe_ctx *ctx, const SlotList &slots) {
    const Pyramid &pyr = ctx->pyr;
    int launches = 0;
    for (int l = 0; l + 1 < pyr.nlevels; ++l) {
        const LevelGeom &gd = pyr.lv[l + 1];
        dim3 grid((gd.w + PD_TW - 1) / PD_TW, (gd.h + PD_TH - 1) / PD_TH, slots.n);
        RDFE_LAUNCH(ctx, K_PYRDOWN, (pyrdown_kernel<<<grid, 256, 0, ctx->stream>>>(pyr, slots, l)));
        ++launches;
    }
    TileTable tt;
    HaloTable ht;
    int nt = 0, nh = 0;
    for (int l = 0; l < pyr.nlevels; ++l) {
        const LevelGeom &g = pyr.lv[l];
        tt.first[l] = nt;
        tt.tiles_x[l] = (g.w + SC_TW - 1) / SC_TW;
        nt += tt.tiles_x[l] * ((g.h + SC_TH - 1) / SC_TH);
        ht.first[l] = nh;
        const int npx = 2 * pyr.win * (g.w + 2 * pyr.win) + 2 * g.h * pyr.win;
        nh += (npx + 255) / 256;
    }
    tt.first[pyr.nlevels] = nt;
    ht.first[pyr.nlevels] = nh;
    RDFE_LAUNCH(ctx, K_SCHARR, (scharr_kernel<<<dim3(nt, slots.n), 256, 0, ctx->stream>>>(pyr, slots, tt)));
    RDFE_LAUNCH(ctx, K_HALO, (halo_kernel<<<dim3(nh, slots.n), 256, 0, ctx->stream>>>(pyr, slots, ht)));
    return launches + 2;
}

}  // namespace rdfe
