// clahe.cu -- K1: CLAHE for 8-bit images, bit-exact with cv::CLAHE::apply
// (reference call site: OpenCvImage::preprocess, src/rdvio_extra/src/opencv_image.cpp:157;
//  parameters: OpenCvImage::clahe, :179-182; arithmetic: SURVEY.md App. A1).
//
// Two launches per batch:
//   clahe_hist_lut_kernel  one CTA per (tile, image): 256-bin histogram with
//                          per-warp private copies in shared memory, integer
//                          clip + redistribute, prefix sum, LUT (float scale,
//                          round-half-even).            HBM: reads S bytes.
//   clahe_apply_kernel     one CTA per (row band, image): builds, per
//                          interpolation cell column, a 256-entry table of
//                          the FOUR tile-LUT bytes a pixel needs packed in one
//                          32-bit word (one shared-memory gather per pixel
//                          instead of four), then blends in float32 in the
//                          reference's exact op order (no FMA) and writes
//                          level 0 of the pyramid.      HBM: reads S, writes S.
#include "fe_internal.cuh"

namespace rdfe {

constexpr int kHistWarps = 8;

__global__ void __launch_bounds__(kHistWarps * 32)
clahe_hist_lut_kernel(const uint8_t *const *__restrict__ src, size_t pitch, int vec4, ClaheParams cp,
                      uint8_t *__restrict__ lut) {
    // per-warp private histograms; bins 256..511 of each are trash bins for the bytes of an edge word that lie
    // outside the tile (keeps the atomics branch-free; never zeroed, never read)
    // The rows must start at a multiple of 2048 in the shared address space (the atomics below compose
    // "row base | trash bit | 4 * byte" with an OR); the static base is not (the first KB of the window is
    // reserved), so one spare row is allocated and the base rounded up at run time.
    __shared__ __align__(16) unsigned hist_raw[(kHistWarps + 1) * 512];
    __shared__ unsigned wsum[kHistWarps];
    const unsigned hist_sa = (unsigned)__cvta_generic_to_shared(hist_raw);
    unsigned (*hist)[512] = reinterpret_cast<unsigned (*)[512]>(hist_raw + ((((hist_sa + 2047u) & ~2047u) - hist_sa) >> 2));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = blockIdx.x, b = blockIdx.y;
    const int tx = tile % cp.tiles_x, ty = tile / cp.tiles_x;
    const uint8_t *img = src[b];

    for (int i = tid; i < kHistWarps * 64; i += kHistWarps * 32) reinterpret_cast<uint4 *>(&hist[i >> 6][0])[i & 63] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();

    const int x0 = tx * cp.tw, y0 = ty * cp.th;
    if (!cp.padded && vec4) {
        // aligned 32-bit words covering the tile row.  All the row loads of a warp are issued before the first
        // atomic so that they overlap (the kernel is latency-bound).
        const int wa = x0 & ~3;                                    // first aligned column
        const int nwords = ((x0 + cp.tw + 3) & ~3) - wa >> 2;     // words per row
        constexpr int RPW = 8;                                     // rows per warp batch
        unsigned *hw = hist[warp];
        for (int yb = warp * RPW; yb < cp.th; yb += kHistWarps * RPW) {
            for (int wi = lane; wi < nwords; wi += 32) {
                unsigned wv[RPW];
                const uint8_t *rp = img + (size_t)(y0 + yb) * pitch + wa + 4 * wi;
#pragma unroll
                for (int r = 0; r < RPW; ++r, rp += pitch)
                    wv[r] = (yb + r < cp.th) ? __ldg(reinterpret_cast<const unsigned *>(rp)) : 0u;
                const int xw = wa + 4 * wi;
                // bytes of this lane's word column outside [x0, x0 + tw) go to the trash half (byte offset 1024);
                // the same for every row
                // (hist rows are 2048-byte aligned, so row base | trash bit | 4 * byte is a plain OR)
                unsigned t4[4];
                const unsigned hb = (unsigned)__cvta_generic_to_shared(hw);
#pragma unroll
                for (int k = 0; k < 4; ++k) t4[k] = hb | (((xw + k >= x0) && (xw + k < x0 + cp.tw)) ? 0u : 1024u);
                auto bump = [](unsigned addr) { asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr) : "memory"); };
#pragma unroll
                for (int r = 0; r < RPW; ++r) {
                    if (yb + r < cp.th) {                          // warp-uniform
                        const unsigned w = wv[r];
                        // byte offset of the bin = 4 * byte, extracted pre-scaled
                        bump(((w << 2) & 0x3FCu) | t4[0]);
                        bump(((w >> 6) & 0x3FCu) | t4[1]);
                        bump(((w >> 14) & 0x3FCu) | t4[2]);
                        bump(((w >> 22) & 0x3FCu) | t4[3]);
                    }
                }
            }
        }
    } else if (!cp.padded) {
        for (int y = warp; y < cp.th; y += kHistWarps) {
            const uint8_t *row = img + (size_t)(y0 + y) * pitch + x0;
            for (int x = lane; x < cp.tw; x += 32) atomicAdd(&hist[warp][__ldg(row + x)], 1u);
        }
    } else {
        // copyMakeBorder(..., BORDER_REFLECT_101) of the reference's padded copy, by index
        for (int y = warp; y < cp.th; y += kHistWarps) {
            const uint8_t *row = img + (size_t)reflect101(y0 + y, cp.H) * pitch;
            for (int x = lane; x < cp.tw; x += 32) atomicAdd(&hist[warp][__ldg(row + reflect101(x0 + x, cp.W))], 1u);
        }
    }
    __syncthreads();

    // one thread per bin from here on
    int h = 0;
#pragma unroll
    for (int w = 0; w < kHistWarps; ++w) h += (int)hist[w][tid];

    if (cp.clip > 0) {
        int excess = h > cp.clip ? h - cp.clip : 0;
        if (h > cp.clip) h = cp.clip;
        int s = __reduce_add_sync(0xffffffffu, excess);
        if (lane == 0) wsum[warp] = (unsigned)s;
        __syncthreads();
        int clipped = 0;
#pragma unroll
        for (int w = 0; w < kHistWarps; ++w) clipped += (int)wsum[w];
        __syncthreads();
        const int batch = clipped / 256;
        const int resid = clipped - batch * 256;
        h += batch;
        if (resid != 0) {
            int step = 256 / resid;
            if (step < 1) step = 1;
            // for (i = 0; i < 256 && resid > 0; i += step, --resid) hist[i]++
            if (tid % step == 0 && tid / step < resid) h += 1;
        }
    }
    // inclusive prefix sum over the 256 bins
    int v = h;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    if (lane == 31) wsum[warp] = (unsigned)v;
    __syncthreads();
    int base = 0;
#pragma unroll
    for (int w = 0; w < kHistWarps; ++w) base += (w < warp) ? (int)wsum[w] : 0;
    const int sum = base + v;
    int q = __float2int_rn((float)sum * cp.lut_scale);
    q = q < 0 ? 0 : q > 255 ? 255 : q;
    lut[((size_t)b * (cp.tiles_x * cp.tiles_y) + tile) * 256 + tid] = (uint8_t)q;
}

struct ApplyBands {
    int nbands;
    short y0[96], y1[96], cy[96];
};

// magic-number conversions: exact for 0..255 and cheaper than I2F/F2I
__device__ __forceinline__ float u8_to_float(unsigned byte) { return __uint_as_float(0x4B000000u | byte) - 8388608.0f; }
__device__ __forceinline__ unsigned float_to_u8_rn(float v) {
    // v in [0, 256): adding 1.5*2^23 rounds to nearest even at integer granularity
    unsigned r = __float_as_uint(v + 12582912.0f) & 0x1FFu;
    return r > 255u ? 255u : r;
}

__global__ void __launch_bounds__(256)
clahe_apply_kernel(const uint8_t *const *__restrict__ src, size_t pitch, int src_vec4, ClaheParams cp,
                   ApplyBands bands, const uint8_t *__restrict__ lut, Pyramid pyr, SlotList slots) {
    extern __shared__ uint32_t smem_u32[];
    // [tiles_x + 1][256] combined LUT words (l11, l12, l21, l22), then per-column xa (float) and cell index
    uint32_t *comb = smem_u32;
    const int ncx = cp.tiles_x + 1;
    const int Wg = (cp.W + 3) & ~3;
    float *xa_s = reinterpret_cast<float *>(comb + ncx * 256);          // [Wg]
    uint32_t *cb_s = reinterpret_cast<uint32_t *>(xa_s + Wg);           // [Wg] column base into comb (cell << 8)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int band = blockIdx.x, b = blockIdx.y;
    const int cy = bands.cy[band], y0 = bands.y0[band], y1 = bands.y1[band];
    const int ty1 = max(cy - 1, 0), ty2 = min(cy, cp.tiles_y - 1);
    const uint8_t *L = lut + (size_t)b * (cp.tiles_x * cp.tiles_y) * 256;
    for (int i = tid; i < ncx * 256; i += 256) {
        const int c = i >> 8, v = i & 255;
        const int tx1 = max(c - 1, 0), tx2 = min(c, cp.tiles_x - 1);
        const unsigned l11 = L[(ty1 * cp.tiles_x + tx1) * 256 + v], l12 = L[(ty1 * cp.tiles_x + tx2) * 256 + v];
        const unsigned l21 = L[(ty2 * cp.tiles_x + tx1) * 256 + v], l22 = L[(ty2 * cp.tiles_x + tx2) * 256 + v];
        comb[i] = l11 | (l12 << 8) | (l21 << 16) | (l22 << 24);
    }
    for (int x = tid; x < Wg; x += 256) {
        const float txf = (float)x * cp.inv_tw - 0.5f;
        xa_s[x] = txf - floorf(txf);
        int c = 0;
        while (c < ncx - 1 && x >= cp.xb[c + 1]) ++c;                   // interpolation cell (host-built boundaries)
        cb_s[x] = (uint32_t)c << 8;
    }
    __syncthreads();

    const uint8_t *img = src[b];
    uint8_t *dst = pyr.image_origin(0, slots.v[b]);
    const int dpitch = pyr.lv[0].ipitch;
    const int groups = Wg >> 2;
    for (int y = y0 + warp; y < y1; y += 8) {
        const float tyf = (float)y * cp.inv_th - 0.5f;
        const float ya = tyf - floorf(tyf), ya1 = 1.0f - ya;
        const uint8_t *srow = img + (size_t)y * pitch;
        for (int g = lane; g < groups; g += 32) {
            const int x = g << 2;
            unsigned px;
            if (src_vec4 && x + 4 <= cp.W) {
                px = __ldg(reinterpret_cast<const unsigned *>(srow + x));
            } else {
                px = 0;
                for (int i = 0; i < min(4, cp.W - x); ++i) px |= (unsigned)__ldg(srow + x + i) << (8 * i);
            }
            const float4 xa4 = *reinterpret_cast<const float4 *>(xa_s + x);
            const uint4 cb4 = *reinterpret_cast<const uint4 *>(cb_s + x);
            const float xas[4] = {xa4.x, xa4.y, xa4.z, xa4.w};
            const unsigned cbs[4] = {cb4.x, cb4.y, cb4.z, cb4.w};
            unsigned rb[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float xa = xas[i], xa1 = 1.0f - xa;
                const unsigned e = comb[cbs[i] + __byte_perm(px, 0u, 0x4440u | (unsigned)i)];
                // four LUT bytes -> floats: PRMT builds the bits of 2^23 + b, one FADD removes the 2^23 (exact)
                const float l11 = __uint_as_float(__byte_perm(e, 0x4B000000u, 0x7650u)) - 8388608.0f;
                const float l12 = __uint_as_float(__byte_perm(e, 0x4B000000u, 0x7651u)) - 8388608.0f;
                const float l21 = __uint_as_float(__byte_perm(e, 0x4B000000u, 0x7652u)) - 8388608.0f;
                const float l22 = __uint_as_float(__byte_perm(e, 0x4B000000u, 0x7653u)) - 8388608.0f;
                const float res = (l11 * xa1 + l12 * xa) * ya1 + (l21 * xa1 + l22 * xa) * ya;
                // res in [0, 255.5): adding 1.5 * 2^23 leaves rint(res) (ties to even) in the low mantissa byte
                rb[i] = __float_as_uint(res + 12582912.0f);
            }
            const unsigned out = __byte_perm(__byte_perm(rb[0], rb[1], 0x0040u), __byte_perm(rb[2], rb[3], 0x0040u), 0x5410u);
            store4_with_halo(dst, dpitch, cp.W, cp.H, pyr.win, x, y, out);
        }
    }
}

// Fast path of clahe_apply_kernel for the common case (4-byte aligned sources, W % 4 == 0): identical arithmetic,
// but every loop-invariant is hoisted, rows are walked with bumped pointers and the store takes the interior
// branch with two integer compares.  The general kernel above stays as the fallback.
__global__ void __launch_bounds__(256)
clahe_apply_fast_kernel(const uint8_t *const *__restrict__ src, size_t pitch, ClaheParams cp, ApplyBands bands,
                        const uint8_t *__restrict__ lut, Pyramid pyr, SlotList slots) {
    // shared memory: [ncx][256] entries of the four tile-LUT values of a pixel (l11, l12, l21, l22) as bf16 (exact
    // for 0..255; float = bits << 16), 8 bytes per entry: one 8-byte gather per pixel -- the kernel is bound by
    // shared-memory wavefronts, a float4 entry costs twice as many -- then xa, 1-xa and the cell table offset per column.
    extern __shared__ uint2 smem_u2[];
    uint2 *comb = smem_u2;
    const int ncx = cp.tiles_x + 1;
    const int W = cp.W, H = cp.H, win = pyr.win;
    float *xa_s = reinterpret_cast<float *>(comb + ncx * 256);          // [W]
    float *xa1_s = xa_s + W;                                            // [W]
    uint32_t *cb_s = reinterpret_cast<uint32_t *>(xa1_s + W);           // [W] byte offset of the cell table in comb
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int band = blockIdx.x, b = blockIdx.y;
    const int cy = bands.cy[band], y0 = bands.y0[band], y1 = bands.y1[band];
    const int ty1 = max(cy - 1, 0), ty2 = min(cy, cp.tiles_y - 1);
    const uint8_t *L1 = lut + ((size_t)b * (cp.tiles_x * cp.tiles_y) + (size_t)ty1 * cp.tiles_x) * 256;
    const uint8_t *L2 = lut + ((size_t)b * (cp.tiles_x * cp.tiles_y) + (size_t)ty2 * cp.tiles_x) * 256;
    for (int i = tid; i < ncx * 256; i += 256) {
        const int c = i >> 8, v = i & 255;
        const int o1 = max(c - 1, 0) * 256 + v, o2 = min(c, cp.tiles_x - 1) * 256 + v;
        auto bf = [](unsigned v) { return __float_as_uint(u8_to_float(v)) >> 16; };     // exact: 8 significant bits
        comb[i] = make_uint2(bf(L1[o1]) | (bf(L1[o2]) << 16), bf(L2[o1]) | (bf(L2[o2]) << 16));
    }
    for (int x = tid; x < W; x += 256) {
        const float txf = (float)x * cp.inv_tw - 0.5f;
        const float xa = txf - floorf(txf);
        xa_s[x] = xa;
        xa1_s[x] = 1.0f - xa;
        int c = 0;
        while (c < ncx - 1 && x >= cp.xb[c + 1]) ++c;
        cb_s[x] = (uint32_t)c << 11;                                    // cell * 256 entries * 8 bytes
    }
    __syncthreads();

    const int groups = W >> 2;
    const bool coop = coop_halo_ok(W, win);
    const int nh = (win + 3) >> 2;
    const int g_lo = win / 4 + 1, g_hi = (W - 5 - win) >> 2;            // interior groups: g_lo <= g <= g_hi
    uint8_t *dst = pyr.image_origin(0, slots.v[b]);
    const int dpitch = pyr.lv[0].ipitch;
    const uint8_t *combb = reinterpret_cast<const uint8_t *>(comb);
    const uint8_t *srow = src[b] + (size_t)(y0 + warp) * pitch;
    uint8_t *drow = dst + (size_t)(y0 + warp) * dpitch;
    for (int y = y0 + warp; y < y1; y += 8, srow += 8 * pitch, drow += 8 * (size_t)dpitch) {
        const float tyf = (float)y * cp.inv_th - 0.5f;
        const float ya = tyf - floorf(tyf), ya1 = 1.0f - ya;
        const bool row_int = (y > win) && (y < H - 1 - win);
        auto blend4 = [&](int g, unsigned px) -> unsigned {
            const float4 xa4 = reinterpret_cast<const float4 *>(xa_s)[g];
            const float4 xb4 = reinterpret_cast<const float4 *>(xa1_s)[g];
            const uint4 cb4 = reinterpret_cast<const uint4 *>(cb_s)[g];
            const float xas[4] = {xa4.x, xa4.y, xa4.z, xa4.w};
            const float xbs[4] = {xb4.x, xb4.y, xb4.z, xb4.w};
            const unsigned cbs[4] = {cb4.x, cb4.y, cb4.z, cb4.w};
            unsigned rb[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float xa = xas[i], xa1 = xbs[i];
                const uint2 e = *reinterpret_cast<const uint2 *>(combb + cbs[i] + 8u * __byte_perm(px, 0u, 0x4440u | (unsigned)i));
                const float l11 = __uint_as_float(e.x << 16), l12 = __uint_as_float(e.x & 0xFFFF0000u);
                const float l21 = __uint_as_float(e.y << 16), l22 = __uint_as_float(e.y & 0xFFFF0000u);
                const float res = (l11 * xa1 + l12 * xa) * ya1 + (l21 * xa1 + l22 * xa) * ya;
                rb[i] = __float_as_uint(res + 12582912.0f);
            }
            return __byte_perm(__byte_perm(rb[0], rb[1], 0x0040u), __byte_perm(rb[2], rb[3], 0x0040u), 0x5410u);
        };
        const unsigned *sw = reinterpret_cast<const unsigned *>(srow);
        unsigned *dw = reinterpret_cast<unsigned *>(drow);
        if (coop) {
            // all groups of the row in batches of 32; the first / last batch also writes the side halos from
            // registers (store4_row_coop), rows next to the top / bottom edge are stored twice
            const int last = ((groups - 1) >> 5) << 5;
            unsigned pn = (lane < groups) ? __ldg(sw + lane) : 0u;
            for (int gbase = 0; gbase < groups; gbase += 32) {
                const int g = gbase + lane;
                const bool act = g < groups;
                const unsigned px = pn;
                if (g + 32 < groups) pn = __ldg(sw + g + 32);
                const unsigned v = blend4(act ? g : groups - 1, px);
                if (row_int) store4_row_coop(drow, W, nh, g, act, v, gbase == 0, gbase == last);
                else store4_rows_coop(dst, dpitch, W, H, win, nh, g, y, act, v, gbase == 0, gbase == last);
            }
        } else if (row_int) {
            // interior groups: branch-free stores, the next group's pixels are requested before this one is blended
            int g = g_lo + lane;
            unsigned pn = (g <= g_hi) ? __ldg(sw + g) : 0u;
            for (; g <= g_hi; g += 32) {
                const unsigned px = pn;
                if (g + 32 <= g_hi) pn = __ldg(sw + g + 32);
                dw[g] = blend4(g, px);
            }
            // the few groups whose mirror images live in the left / right halo
            for (int gb = lane; gb < groups; gb += 32) {
                if (gb >= g_lo && gb <= g_hi) { gb += ((g_hi - gb) / 32) * 32; continue; }
                store4_border(dst, dpitch, W, H, win, gb << 2, y, blend4(gb, __ldg(sw + gb)));
            }
        } else if (H - 1 - win <= win) {
            // tiny image: a row may have mirror images in both halos
            for (int g = lane; g < groups; g += 32)
                store4_border(dst, dpitch, W, H, win, g << 2, y, blend4(g, __ldg(sw + g)));
        } else {
            // a row whose mirror image lies in the top / bottom halo: interior column groups store the same word
            // twice (row y and its REFLECT_101 image), only the side groups take the general path
            ptrdiff_t mirror = 0;
            if (y >= 1 && y <= win) mirror = -2 * (ptrdiff_t)y * dpitch;
            else if (y >= H - 1 - win && y <= H - 2) mirror = 2 * (ptrdiff_t)(H - 1 - y) * dpitch;
            unsigned *dm = reinterpret_cast<unsigned *>(drow + mirror);
            for (int g = g_lo + lane; g <= g_hi; g += 32) {
                const unsigned v = blend4(g, __ldg(sw + g));
                dw[g] = v;
                if (mirror) dm[g] = v;
            }
            for (int gb = lane; gb < groups; gb += 32) {
                if (gb >= g_lo && gb <= g_hi) { gb += ((g_hi - gb) / 32) * 32; continue; }
                store4_border(dst, dpitch, W, H, win, gb << 2, y, blend4(gb, __ldg(sw + gb)));
            }
        }
    }
}

int launch_clahe(rdfe_ctx *ctx, const SlotList &slots, const uint8_t *const *d_src, size_t src_pitch,
                 int src_vec4, const ClaheParams &cp) {
    for (int i = 0; i < slots.n; ++i) ctx->slot_gen[slots.v[i]] = ++ctx->gen_counter;   // level 0 is rewritten: cached LK templates of these slots are stale
    const int ntiles = cp.tiles_x * cp.tiles_y;
    dim3 g1(ntiles, slots.n);
    RDFE_LAUNCH(ctx, K_CLAHE_HIST, (clahe_hist_lut_kernel<<<g1, kHistWarps * 32, 0, ctx->ls>>>(d_src, src_pitch, src_vec4, cp, ctx->lut)));

    // row bands: split every interpolation cell row into chunks of <= RB rows
    ApplyBands bands;
    bands.nbands = 0;
    int RB = 32;
    for (;;) {
        int nb = 0;
        for (int cy = 0; cy <= cp.tiles_y; ++cy) nb += (cp.yb[cy + 1] - cp.yb[cy] + RB - 1) / RB;
        if (nb <= 96) break;
        RB *= 2;
    }
    for (int cy = 0; cy <= cp.tiles_y; ++cy)
        for (int y = cp.yb[cy]; y < cp.yb[cy + 1]; y += RB) {
            bands.y0[bands.nbands] = (short)y;
            bands.y1[bands.nbands] = (short)min(y + RB, cp.yb[cy + 1]);
            bands.cy[bands.nbands] = (short)cy;
            ++bands.nbands;
        }
    dim3 g2(bands.nbands, slots.n);
    const size_t smem = (size_t)(cp.tiles_x + 1) * 256 * sizeof(uint32_t) + (size_t)((cp.W + 3) & ~3) * 8;
    const size_t smem_fast = (size_t)(cp.tiles_x + 1) * 256 * sizeof(uint2) + (size_t)cp.W * 12;
    if (src_vec4 && cp.W % 4 == 0 && smem_fast <= 100 * 1024) {
        if (smem_fast > 48 * 1024 && smem_fast > ctx->smem_optin[2]) {      // per-device attribute: remembered per context
            if (cudaFuncSetAttribute(clahe_apply_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fast) != cudaSuccess) {
                set_error("clahe: cudaFuncSetAttribute(%zu) failed", smem_fast);
                return RDFE_ERR_CUDA;
            }
            ctx->smem_optin[2] = smem_fast;
        }
        RDFE_LAUNCH(ctx, K_CLAHE_APPLY, (clahe_apply_fast_kernel<<<g2, 256, smem_fast, ctx->ls>>>(d_src, src_pitch, cp, bands, ctx->lut,
                                                                                                   ctx->pyr, slots)));
    } else
        RDFE_LAUNCH(ctx, K_CLAHE_APPLY, (clahe_apply_kernel<<<g2, 256, smem, ctx->ls>>>(d_src, src_pitch, src_vec4, cp, bands,
                                                                                         ctx->lut, ctx->pyr, slots)));
    return 2;
}

}  // namespace rdfe
