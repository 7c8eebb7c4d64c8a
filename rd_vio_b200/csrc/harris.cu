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
#include <cstdlib>
#include "fe_internal.cuh"
#include "harris_exact.cuh"

namespace rdfe {

// Warp-rolling formulation: a warp owns a strip of HR_ROWS output rows x 120 output columns.  Lane l
// owns 4 adjacent columns (c0 = x0 - 4 + 4l; lanes 0 and 31 are apron) and walks DOWN the strip one
// pixel row per step, keeping in registers: the Sobel row terms of the last two rows, the gradient
// products of the last row (float64) and their pair sums, and the responses of the last rows.
// Horizontal neighbours come from lane+-1 shuffles.  Pixels are read once per row as one aligned
// 32-bit word per lane straight from the haloed level-0 plane: the materialised REFLECT_101 halo
// IS Sobel's border extension.  boxFilter however reflects the PRODUCT maps, and a product evaluated
// on halo pixels differs from the reflected product in two ways, both handled below:
//   * gx*gy changes sign across a mirrored row or column;
//   * the 3-tap row smoothing ((k0*p[x-1]) + k1*p[x]) + k0*p[x+1] is not associative, so on the
//     mirrored columns x = -1 and x = W it is evaluated in mirrored order.
// All float64 sums are exact (9 terms, exponent spread < 2^29), hence order independent.
constexpr int HR_ROWS = 80;            // output rows per warp strip at full batches (adaptive_strip_rows).  Taller strips
                                       // = fewer halo rows: in the pipelined step 44 -> 80 rows gives +1.5 % (120 rows +3 %, but
                                       // then the kernel alone suffers from a 1.01-wave grid)
constexpr int HR_COLS = 120;           // output columns per warp (lanes 1..30)
constexpr int HW_WARPS = 4;            // warps per CTA
constexpr int HW_BUF = 256;            // per-warp candidate staging (keys)

__device__ __forceinline__ float byte_f(unsigned w, int k) {
    // byte k of w as float: build 2^23 + b by PRMT, subtract 2^23 (exact)
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7650u | (unsigned)k)) - 8388608.0f;
}
// float -> double for values that are zero or normal (gradient products are never subnormal: a non-zero
// Sobel term is >= 1 ulp of an O(1e-3..1) float).  Pure integer re-biasing on the ALU pipe: the F2F
// conversion unit (XU, ~4 results/clk/SM measured) is the bottleneck of this kernel.
__device__ __forceinline__ double f2d_exact(float f) {
    const unsigned u = __float_as_uint(f);
    const unsigned mag = u & 0x7FFFFFFFu;
    unsigned hi = (u & 0x80000000u) | ((mag >> 3) + 0x38000000u);
    if (mag == 0u) hi = u;
    return __hiloint2double((int)hi, (int)(u << 29));
}
__device__ __forceinline__ double shfl_up_d(double v) {
    return __hiloint2double(__shfl_up_sync(0xffffffffu, __double2hiint(v), 1), __shfl_up_sync(0xffffffffu, __double2loint(v), 1));
}
__device__ __forceinline__ double shfl_down_d(double v) {
    return __hiloint2double(__shfl_down_sync(0xffffffffu, __double2hiint(v), 1), __shfl_down_sync(0xffffffffu, __double2loint(v), 1));
}

// One strip (HR_ROWS x 120 outputs) by one warp.  BORDER = false is the lean variant for strips whose
// whole 128-column x (rows+6)-row footprint lies inside the image: no mirror/sign/validity logic at all.
//
// Narrow last column tile (BORDER only): when the columns left over after the full 120-column tiles fit in 8 (14)
// output lanes, the warp is split into 3 (2) groups of GL = 10 (16) lanes -- first and last lane of a group are its
// apron -- and group g walks strip number first_strip + g of that tile, so the lanes a single strip would leave
// idle work on the strips below it (752 = 6 x 120 + 32: 38 instead of 42 warp strips per image).  The shuffles
// stay warp-wide: what crosses a group boundary only reaches apron columns whose results are never used, exactly
// like lanes 0 and 31 of the one-group case.  y0, the row count and every row predicate are then per lane; the loop
// trip count, the staging flush and the REDUX stay warp-uniform (group 0 always has the most rows).
template <bool kFma, bool BORDER, bool RESP>
__device__ __forceinline__ void harris_strip(const uint8_t *__restrict__ org, int ipitch, int W, int H, int x0, int y0w, int hr_rows, int gl_lanes,
                                             int n_groups, float k, const DetectScratch &det, int b, float *__restrict__ response,
                                             unsigned long long *buf, unsigned *cnt) {
    const int lane = threadIdx.x & 31;
    const int GL = BORDER ? gl_lanes : 32;
    const int grp = BORDER ? lane / GL : 0;
    const int gl = lane - grp * GL;                         // lane within its group
    const int y0 = y0w + grp * hr_rows;                     // first output row of this lane's strip
    const bool active = !BORDER || (grp < n_groups && y0 < H);
    const int c0 = x0 - 4 + 4 * gl;                         // first of this lane's 4 columns
    const bool ld_ok = BORDER ? (active && (c0 + 3 <= W + 20) && (c0 >= -20)) : true;   // inside the materialised halo
    const bool out_lane = active && gl >= 1 && gl <= GL - 2;   // first and last lane of a group are apron
    const double sc = 1.0 / (4.0 * 3.0 * 255.0);
    const float k0 = (float)sc, k1 = (float)(2.0 * sc);
    // per-column flags (BORDER only)
    bool mir[4], bflipx[4];
    float cinv[4], cemit[4];                                // 0 or -inf: column outside the image / not emittable
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int x = c0 + i;
        mir[i] = BORDER && ((x == -1) || (x == W));
        bflipx[i] = BORDER && ((x < 0) || (x >= W));
        cinv[i] = (BORDER && (x < 0 || x >= W)) ? -INFINITY : 0.0f;
        cemit[i] = (!out_lane || (BORDER && (x < 1 || x >= W - 1))) ? -INFINITY : 0.0f;
    }
    float dA[4] = {0, 0, 0, 0}, dB[4] = {0, 0, 0, 0};       // d(y-2), d(y-1)
    float sA[4] = {0, 0, 0, 0}, sB[4] = {0, 0, 0, 0};       // s(y-2), s(y-1)
    double pa[4] = {0, 0, 0, 0}, pb[4] = {0, 0, 0, 0}, pc[4] = {0, 0, 0, 0};   // products of row p-1
    double ta[4] = {0, 0, 0, 0}, tb[4] = {0, 0, 0, 0}, tc[4] = {0, 0, 0, 0};   // a(p-2)+a(p-1)
    float Rm[6], Rc[6];                                      // rows q-2 (hmax3 in [1..4]) and q-1 (cols -1..4)
#pragma unroll
    for (int i = 0; i < 6; ++i) { Rm[i] = -INFINITY; Rc[i] = -INFINITY; }
    float tmax = 0.0f;

    // Candidate keys are staged in the warp's shared-memory buffer.  Slots come from one shared-memory atomic per
    // lane that has candidates; the fill count is mirrored in a warp-uniform register (REDUX of the per-lane counts),
    // so the flush decision needs neither a shared-memory read nor a warp barrier.
    unsigned nstaged = 0;
    auto flush = [&]() {
        const unsigned nb = nstaged;
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(&det.cand_count[b], nb);
        base = __shfl_sync(0xffffffffu, base, 0);
        __syncwarp();                                    // the staged keys of all lanes are visible
        for (unsigned i = lane; i < nb; i += 32) {
            const unsigned pos = base + i;
            if (pos < det.cand_cap) det.cand[(size_t)b * det.cand_cap + pos] = buf[i];
            else atomicExch(det.overflow, 1u);
        }
        __syncwarp();
        if (lane == 0) *cnt = 0u;
        nstaged = 0;
        __syncwarp();
    };

    const int rows = active ? min(hr_rows, H - y0) : 0;    // of this lane's strip
    const int steps = min(hr_rows, H - y0w) + 6;            // warp-uniform trip count
    const int my_steps = BORDER ? rows + 6 : steps;         // rows this lane may read (its strip ends earlier at the bottom)
    // software pipelining: the pixel words of rows j+1, j+2 are in flight while row j is processed
    const uint8_t *rowp = org + (ptrdiff_t)(y0 - 3) * ipitch + c0;
    unsigned wq0 = 0, wq1 = 0;
    if (ld_ok) { wq0 = *reinterpret_cast<const unsigned *>(rowp); wq1 = *reinterpret_cast<const unsigned *>(rowp + ipitch); }
    rowp += 2 * (ptrdiff_t)ipitch;
    float *resp_row = RESP ? response + ((size_t)b * H + (y0 - 5)) * W + c0 : nullptr;   // row q = y0-5+j (debug tap only)
    unsigned addr_row = (unsigned)((y0 - 6) * W + c0);       // pixel address of (row n = y0-6+j, column c0)
#pragma unroll 2
    for (int j = 0; j < steps; ++j) {
        // ---- pixels c0-1 .. c0+4 of row y = y0-3+j as floats
        const unsigned w = wq0;
        wq0 = wq1;
        if (ld_ok && j + 2 < my_steps) wq1 = *reinterpret_cast<const unsigned *>(rowp);
        rowp += ipitch;
        const unsigned wl = __shfl_up_sync(0xffffffffu, w, 1), wr = __shfl_down_sync(0xffffffffu, w, 1);
        float p[6];
        p[0] = byte_f(wl, 3);
        p[1] = byte_f(w, 0); p[2] = byte_f(w, 1); p[3] = byte_f(w, 2); p[4] = byte_f(w, 3);
        p[5] = byte_f(wr, 0);
        // ---- Sobel row terms of row y
        float dN[4], sN[4], k0p[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) k0p[i] = k0 * p[i];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            dN[i] = p[i + 2] - p[i];
            if (!kFma) {
                const float k1p = k1 * p[i + 1];
                sN[i] = (BORDER && mir[i]) ? ((k0p[i + 2] + k1p) + k0p[i]) : ((k0p[i] + k1p) + k0p[i + 2]);
            } else {
                sN[i] = (BORDER && mir[i]) ? __fmaf_rn(k0, p[i], __fmaf_rn(k1, p[i + 1], k0p[i + 2]))
                                           : __fmaf_rn(k0, p[i + 2], __fmaf_rn(k1, p[i + 1], k0p[i]));
            }
        }
        // ---- gradient products of row pr = y-1 (needs rows y-2, y-1, y)
        const int pr = y0 - 4 + j;
        const bool bflipy = BORDER && ((pr < 0) || (pr >= H));
        double na[4], nb[4], nc[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float gx;
            if (!kFma) gx = k1 * dB[i] + k0 * (dA[i] + dN[i]);
            else gx = __fmaf_rn(k0, dA[i] + dN[i], k1 * dB[i]);
            const float gy = sN[i] - sA[i];
            float fb = gx * gy;
            if (BORDER && (bflipx[i] != bflipy)) fb = -fb;
            na[i] = (double)(gx * gx);
            nb[i] = (double)fb;
            nc[i] = (double)(gy * gy);
        }
        // ---- vertical 3-sums for row q = pr-1, then horizontal 3-sums, response
        const int q = pr - 1;
        double va[6], vb[6], vc[6];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            va[i + 1] = ta[i] + na[i]; vb[i + 1] = tb[i] + nb[i]; vc[i + 1] = tc[i] + nc[i];
            ta[i] = pa[i] + na[i]; tb[i] = pb[i] + nb[i]; tc[i] = pc[i] + nc[i];
            pa[i] = na[i]; pb[i] = nb[i]; pc[i] = nc[i];
        }
        va[0] = shfl_up_d(va[4]); vb[0] = shfl_up_d(vb[4]); vc[0] = shfl_up_d(vc[4]);
        va[5] = shfl_down_d(va[1]); vb[5] = shfl_down_d(vb[1]); vc[5] = shfl_down_d(vc[1]);
        const float rinv = (BORDER && (q < 0 || q >= H)) ? -INFINITY : 0.0f;
        float Rn[6];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float A = (float)((va[i] + va[i + 1]) + va[i + 2]);
            const float B = (float)((vb[i] + vb[i + 1]) + vb[i + 2]);
            const float C = (float)((vc[i] + vc[i + 1]) + vc[i + 2]);
            float v;
            if (!kFma) v = (A * C - B * B) - (k * (A + C)) * (A + C);
            else v = (A * C - B * B) - k * ((A + C) * (A + C));
            // dilate ignores pixels outside the image: -inf there (v + 0 is exact, v - inf = -inf)
            Rn[i + 1] = BORDER ? v + (cinv[i] + rinv) : v;
        }
        Rn[0] = __shfl_up_sync(0xffffffffu, Rn[4], 1);
        Rn[5] = __shfl_down_sync(0xffffffffu, Rn[1], 1);
        if (RESP) {
            if (out_lane && q >= y0 && q < y0 + rows) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (!BORDER || (c0 + i >= 0 && c0 + i < W)) resp_row[i] = Rn[i + 1];
            }
            resp_row += W;
        }
        // ---- NMS for row n = q-1: rows Rm (n-1, reduced to hmax3), Rc (n), Rn (n+1)
        const int n = q - 1;
        const bool nrow_ok = (n < y0 + rows) && (!BORDER || ((n >= 1) && (n < H - 1)));
        if (j >= 6) {                                        // n >= y0 (n < y0 + rows holds for the tallest group)
            unsigned cmask = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float hn = fmaxf(fmaxf(Rn[i], Rn[i + 1]), Rn[i + 2]);
                const float v = Rc[i + 1];
                const float m = fmaxf(fmaxf(Rm[i + 1], hn), fmaxf(Rc[i], Rc[i + 2]));
                const float ve = v + cemit[i];               // -inf where this lane/column may not emit
                if (ve > 0.0f && ve >= m) cmask |= 1u << i;
                if (out_lane) tmax = fmaxf(tmax, v);         // -inf outside the image never wins
            }
            if (!nrow_ok) cmask = 0;
            const unsigned nmine = __popc(cmask);
            if (nmine) {
                unsigned pos = atomicAdd(cnt, nmine);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (cmask & (1u << i)) buf[pos++] = ((unsigned long long)__float_as_uint(Rc[i + 1]) << 32) | (addr_row + (unsigned)i);
            }
            nstaged += __reduce_add_sync(0xffffffffu, nmine);
            if (nstaged > (unsigned)(HW_BUF - 128)) flush();    // warp-uniform: a step adds at most 120 keys
        }
        addr_row += (unsigned)W;
        // roll the response rows: Rm <- hmax3(Rc), Rc <- Rn
#pragma unroll
        for (int i = 0; i < 4; ++i) Rm[i + 1] = fmaxf(fmaxf(Rc[i], Rc[i + 1]), Rc[i + 2]);
#pragma unroll
        for (int i = 0; i < 6; ++i) Rc[i] = Rn[i];
        // ---- roll the row terms
#pragma unroll
        for (int i = 0; i < 4; ++i) { dA[i] = dB[i]; dB[i] = dN[i]; sA[i] = sB[i]; sB[i] = sN[i]; }
    }
    if (nstaged) flush();
    const unsigned mb = __reduce_max_sync(0xffffffffu, __float_as_uint(fmaxf(tmax, 0.0f)));
    if (lane == 0 && mb) atomicMax(&det.frame_max[b], mb);
}

// RESP = true only for the rdfe_harris_response debug tap: the hot path carries neither the pointer nor the branch
template <bool kFma, bool RESP>
__global__ void __launch_bounds__(HW_WARPS * 32)
harris_nms_kernel(Pyramid pyr, SlotList slots, float k, DetectScratch det, float *__restrict__ response, int tiles_x,
                  int strips, int gl_narrow, int n_items, int hr_rows) {
    __shared__ unsigned long long s_buf[HW_WARPS][HW_BUF];
    __shared__ unsigned s_cnt[HW_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int W = pyr.lv[0].w, H = pyr.lv[0].h, ipitch = pyr.lv[0].ipitch;
    // work item = (image, strip), flattened so that consecutive warps take consecutive strips of one image; written
    // as a grid-stride loop (a persistent grid of 2-3 CTAs per SM was measured: slower than one warp per item)
    const int total = n_items * slots.n;
    for (int wi = blockIdx.x * HW_WARPS + warp; wi < total; wi += gridDim.x * HW_WARPS) {
        const int b = wi / n_items, item = wi - b * n_items, slot = slots.v[b];
        // items of an image: tiles_x full-width tiles x strips, then the narrow last tile, n_groups strips per item
        int x0, y0, gl_lanes = 32, n_groups = 1;
        if (item < tiles_x * strips) { x0 = (item % tiles_x) * HR_COLS; y0 = (item / tiles_x) * hr_rows; }
        else {
            gl_lanes = gl_narrow; n_groups = gl_narrow == 10 ? 3 : 2;
            x0 = tiles_x * HR_COLS; y0 = (item - tiles_x * strips) * n_groups * hr_rows;
        }
        const uint8_t *org = pyr.image_origin(0, slot);
        if (lane == 0) s_cnt[warp] = 0u;
        __syncwarp();
        // interior strip: columns x0-5 .. x0+124 and rows y0-6 .. y0+hr_rows+2 all inside the image, and the
        // outputs stay off the 1-px frame
        const bool interior = (x0 - 5 >= 0) && (x0 + 124 < W) && (y0 - 6 >= 0) && (y0 + hr_rows + 2 < H);
        if (interior) harris_strip<kFma, false, RESP>(org, ipitch, W, H, x0, y0, hr_rows, 32, 1, k, det, b, response, s_buf[warp], &s_cnt[warp]);
        else harris_strip<kFma, true, RESP>(org, ipitch, W, H, x0, y0, hr_rows, gl_lanes, n_groups, k, det, b, response, s_buf[warp], &s_cnt[warp]);
    }
}

// =====================================================================================================================
// Prefilter path (default): the float64 chain above costs ~90 instructions per pixel although < 3 % of the pixels can
// be local maxima.  harris_flag_kernel evaluates, for ALL pixels, the integer Sobel gradients (IDP.4A on the pixel
// bytes), their products and 3x3 box sums (exact, < 2^24) and an interval [lo, hi] = Ru -+ eps(T) that provably
// contains the reference's float32 response divided by sigma^4 (harris_exact.cuh), and flags a pixel iff it may be a
// positive 3x3 local maximum (hi >= max of the neighbours' lo); the flag carries a "certain" bit when it surely is one
// (lo > max of the neighbours' hi).  harris_resolve_kernel then evaluates the reference's exact arithmetic at the
// flagged pixels only (and, for uncertain ones, at their 8 neighbours), emits the same 64-bit keys as the exact
// kernel and the exact frame maximum (the maximum pixel is always flagged).  Same strips, lane groups and halo reads
// as harris_strip.
__device__ __forceinline__ int dp4a_u8s8(unsigned px, int w, int c) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(px), "r"(w), "r"(c));
    return d;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ float sqrt_approx(float a) {
    float d;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(a));   // T is 0 or >= 1: flushing subnormals changes nothing
    return d;
}

template <int V> struct IntC { static constexpr int value = V; };

// Blackwell packed float32 pairs (FADD2 / FMUL2 / FFMA2): two lanes per instruction on the fma pipe
struct F2 { unsigned long long v; };
__device__ __forceinline__ F2 f2_pack(float lo, float hi) { F2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void f2_unpack(F2 a, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); }
__device__ __forceinline__ F2 f2_add(F2 a, F2 b) { F2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ F2 f2_mul(F2 a, F2 b) { F2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ F2 f2_fma(F2 a, F2 b, F2 c) { F2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }

// The loop body is unrolled three times over rings of three registers per history (products of the last rows, the
// -s row terms, the interval rows), indexed at compile time, so that no register is ever moved.
template <bool BORDER>
__device__ __forceinline__ void harris_flag_strip(const uint8_t *__restrict__ org, int ipitch, int W, int H, int x0, int y0w, int hr_rows,
                                                  int gl_lanes, int n_groups, uint8_t *__restrict__ bm, int bm_pitch) {
    const int lane = threadIdx.x & 31;
    const int GL = BORDER ? gl_lanes : 32;
    const int grp = BORDER ? lane / GL : 0;
    const int gl = lane - grp * GL;
    const int y0 = y0w + grp * hr_rows;
    const bool active = !BORDER || (grp < n_groups && y0 < H);
    const int c0 = x0 - 4 + 4 * gl;
    const bool ld_ok = BORDER ? (active && (c0 + 3 <= W + 20) && (c0 >= -20)) : true;
    const bool out_lane = active && gl >= 1 && gl <= GL - 2;
    bool bflipx[4];
    float cinv[4];                                          // 0 or -inf: column outside the image
    unsigned colmask = 0;                                   // columns of this lane that may be flagged
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int x = c0 + i;
        const bool outside = BORDER && (x < 0 || x >= W);
        bflipx[i] = outside;
        cinv[i] = outside ? -INFINITY : 0.0f;
        if (out_lane && !outside) colmask |= 1u << i;
    }
    // weights of the row filters on bytes (x-1, x, x+1, -): [-1 0 1], 2 * [-1 0 1], [1 2 1], -[1 2 1]
    constexpr int WD = 0x000100FF, WD2 = 0x000200FE, WS = 0x00010201, WSN = 0x00FFFEFF;
    int P[4] = {0, 0, 0, 0}, dprev[4] = {0, 0, 0, 0};       // d(y-2) + 2 d(y-1), d(y-1)
    int ns[3][4], pa[3][4], pb[3][4], pc[3][4];             // rings: -s of the last rows, products of the last rows
    float L[3][6], Hh[3][6];                                // rings: interval rows (columns -1 .. 4)
    float Lx[3][4], Hx[3][4];                               // rings: horizontal max3 of the interval rows
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { ns[r][i] = 0; pa[r][i] = 0; pb[r][i] = 0; pc[r][i] = 0; Lx[r][i] = -INFINITY; Hx[r][i] = -INFINITY; }
#pragma unroll
        for (int i = 0; i < 6; ++i) { L[r][i] = -INFINITY; Hh[r][i] = -INFINITY; }
    }

    const int rows = active ? min(hr_rows, H - y0) : 0;
    const int steps = min(hr_rows, H - y0w) + 6;
    const int my_steps = BORDER ? rows + 6 : steps;
    const uint8_t *rowp = org + (ptrdiff_t)(y0 - 3) * ipitch + c0;
    unsigned wq0 = 0, wq1 = 0;
    if (ld_ok) { wq0 = *reinterpret_cast<const unsigned *>(rowp); wq1 = *reinterpret_cast<const unsigned *>(rowp + ipitch); }
    rowp += 2 * (ptrdiff_t)ipitch;
    // flag bytes: bits 0-3 = columns c0 .. c0+3 of row n may be positive local maxima, bits 4-7 = they surely are
    uint8_t *bm_row = bm + (ptrdiff_t)(y0 - 6) * bm_pitch + (c0 >> 2);   // row n = y0-6+j
    const bool st_ok = out_lane && c0 >= 0 && c0 < W;

    // one pixel row: PH = j % 3 selects the ring slots
    auto step = [&](auto ph, int j) {
        constexpr int R0 = decltype(ph)::value, R1 = (R0 + 1) % 3, R2 = (R0 + 2) % 3;   // R0: written now; R1: oldest; R2: previous
        const unsigned w = wq0;
        wq0 = wq1;
        if (ld_ok && j + 2 < my_steps) wq1 = *reinterpret_cast<const unsigned *>(rowp);
        rowp += ipitch;
        const unsigned wl = __shfl_up_sync(0xffffffffu, w, 1), wr = __shfl_down_sync(0xffffffffu, w, 1);
        unsigned al[4];                                      // bytes (x-1, x, x+1, x+2) of column x = c0 + i
        al[0] = __funnelshift_r(wl, w, 24);
        al[1] = w;
        al[2] = __funnelshift_r(w, wr, 8);
        al[3] = __funnelshift_r(w, wr, 16);
        // ---- integer Sobel of centre row pr = y-1 (rows y-2, y-1, y), products, vertical 3-sums of row q = pr-1
        const int pr = y0 - 4 + j;
        const bool bflipy = BORDER && ((pr < 0) || (pr >= H));
        int va[6], vb[6], vc[6];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int gx = dp4a_u8s8(al[i], WD, P[i]);
            const int gy = dp4a_u8s8(al[i], WS, ns[R1][i]);   // s(y) - s(y-2)
            P[i] = dp4a_u8s8(al[i], WD2, dprev[i]);
            dprev[i] = dp4a_u8s8(al[i], WD, 0);
            ns[R0][i] = dp4a_u8s8(al[i], WSN, 0);             // slot R0 held -s(y-3): dead
            int nb = gx * gy;
            if (BORDER && (bflipx[i] != bflipy)) nb = -nb;   // the product of a mirrored row or column changes sign
            pa[R0][i] = gx * gx; pb[R0][i] = nb; pc[R0][i] = gy * gy;
            va[i + 1] = pa[R1][i] + pa[R2][i] + pa[R0][i];
            vb[i + 1] = pb[R1][i] + pb[R2][i] + pb[R0][i];
            vc[i + 1] = pc[R1][i] + pc[R2][i] + pc[R0][i];
        }
        va[0] = __shfl_up_sync(0xffffffffu, va[4], 1); vb[0] = __shfl_up_sync(0xffffffffu, vb[4], 1); vc[0] = __shfl_up_sync(0xffffffffu, vc[4], 1);
        va[5] = __shfl_down_sync(0xffffffffu, va[1], 1); vb[5] = __shfl_down_sync(0xffffffffu, vb[1], 1); vc[5] = __shfl_down_sync(0xffffffffu, vc[1], 1);
        // ---- interval of the response of row q (ring slot R0), two pixels per instruction
        const int q = pr - 1;
        const float rinv = (BORDER && (q < 0 || q >= H)) ? -INFINITY : 0.0f;
#pragma unroll
        for (int i = 0; i < 4; i += 2) {
            const F2 fA = f2_pack((float)(va[i] + va[i + 1] + va[i + 2]), (float)(va[i + 1] + va[i + 2] + va[i + 3]));
            const F2 nB = f2_pack((float)(vb[i] + vb[i + 1] + vb[i + 2]), (float)(vb[i + 1] + vb[i + 2] + vb[i + 3]));   // -IB (products stored negated)
            const F2 fC = f2_pack((float)(vc[i] + vc[i + 1] + vc[i + 2]), (float)(vc[i + 1] + vc[i + 2] + vc[i + 3]));
            const F2 T = f2_add(fA, fC);
            // ru = IA*IC - IB^2 - 0.04 T^2 = -(nB*nB) ... evaluated as fma(T, -0.04 T, fma(nB, -nB, IA*IC)): the sign of IB drops out,
            // so a second register with the negated sum is avoided by squaring through (nB * kNeg) * nB
            const F2 kNeg = f2_pack(-1.0f, -1.0f), kK = f2_pack(-0.04f, -0.04f);
            F2 ru = f2_fma(f2_mul(nB, kNeg), nB, f2_mul(fA, fC));
            ru = f2_fma(f2_mul(T, kK), T, ru);
            float t0, t1;
            f2_unpack(T, t0, t1);
            const F2 sq = f2_pack(sqrt_approx(t0), sqrt_approx(t1));
            const F2 e = f2_fma(T, f2_fma(f2_pack(kHarrisC1, kHarrisC1), sq, f2_mul(f2_pack(kHarrisC2, kHarrisC2), T)), f2_pack(kHarrisRhoU, kHarrisRhoU));
            F2 lo = f2_fma(e, kNeg, ru), hi = f2_add(ru, e);
            if (BORDER) {
                const F2 off = f2_pack(cinv[i] + rinv, cinv[i + 1] + rinv);   // -inf outside the image (dilate ignores those pixels)
                lo = f2_add(lo, off); hi = f2_add(hi, off);
            }
            f2_unpack(lo, L[R0][i + 1], L[R0][i + 2]);
            f2_unpack(hi, Hh[R0][i + 1], Hh[R0][i + 2]);
        }
        L[R0][0] = __shfl_up_sync(0xffffffffu, L[R0][4], 1); Hh[R0][0] = __shfl_up_sync(0xffffffffu, Hh[R0][4], 1);
        L[R0][5] = __shfl_down_sync(0xffffffffu, L[R0][1], 1); Hh[R0][5] = __shfl_down_sync(0xffffffffu, Hh[R0][1], 1);
#pragma unroll
        for (int i = 0; i < 4; ++i) { Lx[R0][i] = fmax3(L[R0][i], L[R0][i + 1], L[R0][i + 2]); Hx[R0][i] = fmax3(Hh[R0][i], Hh[R0][i + 1], Hh[R0][i + 2]); }
        // ---- NMS of row n = q-1 (ring slot R2) on the intervals: rows n-1 (slot R1, max3), n, n+1 (slot R0, max3)
        const int n = q - 1;
        if (j >= 6) {
            // flagged <=> hi >= max(neighbours' lo, rho+) (rho+ = next float above kHarrisRhoU turns "hi > rho" into ">=");
            // certain <=> lo > max(neighbours' hi)
            const float rho_next = __uint_as_float(__float_as_uint(kHarrisRhoU) + 1u);
            bool fl[4], ce[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float mlo = fmax3(fmax3(L[R2][i], L[R2][i + 2], rho_next), Lx[R1][i], Lx[R0][i]);
                const float mhi = fmax3(fmaxf(Hh[R2][i], Hh[R2][i + 2]), Hx[R1][i], Hx[R0][i]);
                fl[i] = Hh[R2][i + 1] >= mlo;
                ce[i] = L[R2][i + 1] > mhi;
            }
            unsigned fmask = (fl[0] ? 1u : 0u) | (fl[1] ? 2u : 0u) | (fl[2] ? 4u : 0u) | (fl[3] ? 8u : 0u);
            const unsigned cmask = (ce[0] ? 1u : 0u) | (ce[1] ? 2u : 0u) | (ce[2] ? 4u : 0u) | (ce[3] ? 8u : 0u);
            if (st_ok && n < y0 + rows) *bm_row = (uint8_t)((fmask & colmask) | (cmask << 4));   // n >= y0 >= 0, y0 + rows <= H
        }
        bm_row += bm_pitch;
    };
    const int steps3 = (steps + 2) / 3 * 3;                  // the up to 2 extra steps read nothing and flag nothing
#pragma unroll 1
    for (int j = 0; j < steps3; j += 3) {
        step(IntC<0>{}, j);
        step(IntC<1>{}, j + 1);
        step(IntC<2>{}, j + 2);
    }
}

// Flag bytes of image b: [H][bm_pitch] at the start of the image's scratch buffer (det.cand2, free between two selections).  Every (row, 4-column group) of the image is written by exactly one lane.
__host__ __device__ __forceinline__ int harris_bm_pitch(int W) { return (((W + 3) >> 2) + 7) & ~7; }

template <int MINB>
__global__ void __launch_bounds__(HW_WARPS * 32, MINB)
harris_flag_kernel(Pyramid pyr, SlotList slots, DetectScratch det, int tiles_x, int strips, int gl_narrow, int n_items, int hr_rows) {
    const int warp = threadIdx.x >> 5;
    const int W = pyr.lv[0].w, H = pyr.lv[0].h, ipitch = pyr.lv[0].ipitch;
    const int bm_pitch = harris_bm_pitch(W);
    const int total = n_items * slots.n;
    for (int wi = blockIdx.x * HW_WARPS + warp; wi < total; wi += gridDim.x * HW_WARPS) {
        const int b = wi / n_items, item = wi - b * n_items, slot = slots.v[b];
        int x0, y0, gl_lanes = 32, n_groups = 1;
        if (item < tiles_x * strips) { x0 = (item % tiles_x) * HR_COLS; y0 = (item / tiles_x) * hr_rows; }
        else {
            gl_lanes = gl_narrow; n_groups = gl_narrow == 10 ? 3 : 2;
            x0 = tiles_x * HR_COLS; y0 = (item - tiles_x * strips) * n_groups * hr_rows;
        }
        const uint8_t *org = pyr.image_origin(0, slot);
        uint8_t *bm = reinterpret_cast<uint8_t *>(det.cand2 + (size_t)b * det.cand_cap);
        // interior strip: columns x0-5 .. x0+124 and rows y0-6 .. y0+hr_rows+2 all inside the image
        const bool interior = (x0 - 5 >= 0) && (x0 + 124 < W) && (y0 - 6 >= 0) && (y0 + hr_rows + 2 < H);
        if (interior) harris_flag_strip<false>(org, ipitch, W, H, x0, y0, hr_rows, 32, 1, bm, bm_pitch);
        else harris_flag_strip<true>(org, ipitch, W, H, x0, y0, hr_rows, gl_lanes, n_groups, bm, bm_pitch);
    }
}

// Exact response at the flagged pixels: 64-bit keys of the positive 3x3 local maxima off the 1-px frame (the candidates
// of goodFeaturesToTrack before its threshold) and the frame maximum.
//   * a warp reads 32 x 8 flag bytes (1024 pixels), compacts the flagged pixels into its shared-memory buffer with a
//     shuffle scan and evaluates them 32 at a time, so that every lane always has a pixel (the remainder is carried to
//     the next chunk);
//   * pixels that are not CERTAIN maxima need their 8 neighbours' exact responses: they go to a CTA-wide queue that
//     is worked off at the end with one (pixel, neighbour) pair per thread -- handled in place they would make whole
//     warps wait for a lane that evaluates nine pixels.
constexpr int HV_THREADS = 256;
constexpr int HV_WBUF = 1024 + 32;     // per-warp buffer: one chunk (every pixel flagged) + a carried remainder
constexpr int HV_QCAP = 1024;          // uncertain pixels per CTA before the in-place fallback
template <bool kFma>
__global__ void __launch_bounds__(HV_THREADS)
harris_resolve_kernel(Pyramid pyr, SlotList slots, float k, DetectScratch det, int chunks_per_warp) {
    __shared__ unsigned s_buf[HV_THREADS / 32][HV_WBUF];
    __shared__ unsigned q_addr[HV_QCAP];
    __shared__ float q_val[HV_QCAP];
    __shared__ unsigned q_ok[HV_QCAP];
    __shared__ unsigned q_n;
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int W = pyr.lv[0].w, H = pyr.lv[0].h, ipitch = pyr.lv[0].ipitch;
    const uint8_t *org = pyr.image_origin(0, slots.v[b]);
    const int bm_pitch = harris_bm_pitch(W), ngroups = (W + 3) >> 2;
    const unsigned long long *bm = det.cand2 + (size_t)b * det.cand_cap;       // flag bytes, read as 8-byte words
    unsigned long long *out = det.cand + (size_t)b * det.cand_cap;
    const int nwords = (bm_pitch >> 3) * H, wpr = bm_pitch >> 3;
    if (threadIdx.x == 0) q_n = 0u;
    __syncthreads();
    unsigned *buf = s_buf[warp];
    float tmax = 0.0f;
    unsigned nbuf = 0, nflag = 0;                            // warp-uniform

    auto emit_keys = [&](bool emit, unsigned long long key) {   // warp-aggregated append (all lanes call)
        const unsigned m = __ballot_sync(0xffffffffu, emit);
        if (m) {
            unsigned pos = 0;
            if (lane == 0) pos = atomicAdd(&det.cand_count[b], (unsigned)__popc(m));
            pos = __shfl_sync(0xffffffffu, pos, 0) + __popc(m & ((1u << lane) - 1u));
            if (emit) {
                if (pos < det.cand_cap) out[pos] = key;
                else atomicExch(det.overflow, 1u);
            }
        }
    };
    auto eval32 = [&](unsigned e, bool have) {                // one flagged pixel per lane
        bool emit = false;
        unsigned long long key = 0ull;
        if (have) {
            const unsigned addr = e & 0x7FFFFFFFu;
            const int y = (int)(addr / (unsigned)W), x = (int)(addr - (unsigned)y * (unsigned)W);
            const float v = harris_exact<kFma>(org, ipitch, W, H, x, y, k);
            tmax = fmaxf(tmax, v);
            if (v > 0.0f && x >= 1 && x < W - 1 && y >= 1 && y < H - 1) {
                key = ((unsigned long long)__float_as_uint(v) << 32) | addr;
                if (e >> 31) emit = true;
                else {
                    const unsigned qp = atomicAdd(&q_n, 1u);
                    if (qp < HV_QCAP) { q_addr[qp] = addr; q_val[qp] = v; q_ok[qp] = 1u; }
                    else {                                   // queue full (pathological frames): in place
                        emit = true;
#pragma unroll 1
                        for (int t = 0; t < 9; ++t)
                            if (t != 4 && v < harris_exact_at<kFma>(org, ipitch, W, H, x + t % 3 - 1, y + t / 3 - 1, k)) emit = false;
                    }
                }
            }
        }
        emit_keys(emit, key);
    };

    const int warp_id = (blockIdx.x * (HV_THREADS / 32) + warp);
    const int first = warp_id * chunks_per_warp * 32;        // first 8-byte word of this warp
    for (int c = 0; c < chunks_per_warp; ++c) {
        const int wi = first + c * 32 + lane;
        if (first + c * 32 >= nwords) break;                 // warp-uniform
        unsigned long long word = 0ull;
        int row = 0, cb = 0;
        if (wi < nwords) {
            row = wi / wpr; cb = (wi - row * wpr) * 8;       // first group (= byte) of this word within the row
            word = bm[wi];
            const int valid = ngroups - cb;                  // bytes that belong to the row (pad bytes are never written)
            if (valid < 8) word &= valid <= 0 ? 0ull : (~0ull >> (8 * (8 - valid)));
        }
        unsigned long long f = word & 0x0F0F0F0F0F0F0F0Full;
        const unsigned mine = __popcll(f);
        unsigned inc = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        const unsigned tot = __shfl_sync(0xffffffffu, inc, 31);
        unsigned pos = nbuf + inc - mine;
        const unsigned rowaddr = (unsigned)row * (unsigned)W + (unsigned)cb * 4u;
        while (f) {
            const int bit = __ffsll((long long)f) - 1;       // 8 * byte + column within the group
            f &= f - 1;
            buf[pos++] = (rowaddr + (unsigned)((bit >> 3) * 4 + (bit & 7))) | ((unsigned)((word >> (bit + 4)) & 1ull) << 31);
        }
        nbuf += tot; nflag += tot;
        __syncwarp();
        while (nbuf >= 32u) {                                // full groups only: every lane has a pixel
            nbuf -= 32u;
            eval32(buf[nbuf + lane], true);
        }
        __syncwarp();
    }
    if (nbuf) eval32(lane < nbuf ? buf[lane] : 0u, lane < nbuf);
    if (lane == 0 && nflag) atomicAdd(&det.flag_count[b], nflag);
    __syncthreads();
    // ---- uncertain pixels: one (pixel, neighbour) pair per thread
    const unsigned nq = min(q_n, (unsigned)HV_QCAP);
    for (unsigned t = threadIdx.x; t < nq * 8u; t += HV_THREADS) {
        const unsigned e = t >> 3, nb = (t & 7u) < 4u ? (t & 7u) : (t & 7u) + 1u;
        const unsigned addr = q_addr[e];
        const int y = (int)(addr / (unsigned)W), x = (int)(addr - (unsigned)y * (unsigned)W);
        if (q_val[e] < harris_exact<kFma>(org, ipitch, W, H, x + (int)(nb % 3u) - 1, y + (int)(nb / 3u) - 1, k)) q_ok[e] = 0u;
    }
    __syncthreads();
    for (unsigned e = threadIdx.x; e < nq; e += HV_THREADS)
        if (q_ok[e]) {
            const unsigned pos = atomicAdd(&det.cand_count[b], 1u);
            if (pos < det.cand_cap) out[pos] = ((unsigned long long)__float_as_uint(q_val[e]) << 32) | q_addr[e];
            else atomicExch(det.overflow, 1u);
        }
    const unsigned mb = __reduce_max_sync(0xffffffffu, __float_as_uint(fmaxf(tmax, 0.0f)));
    if (lane == 0 && mb) atomicMax(&det.frame_max[b], mb);
}

// ---------------------------------------------------------------------------------------------------------------------
// harris_strip3: same arithmetic as harris_strip (every float op of the reference, individually rounded), re-mapped for
// Blackwell's issue limits:
//   * the float32 stages work on PAIRS of pixels with FADD2 / FMUL2 / FFMA2 (pixels 0 and 2 of the lane in one
//     register pair, 1 and 3 in another, so that the left / centre / right taps of a pair are again pairs);
//   * all row histories (Sobel row terms, products, responses) live in rings of three registers indexed at compile
//     time, the loop body is unrolled three times: no register is ever moved;
//   * 3-input FMNMX3 for the 3x3 maxima, box sums with shared pair sums (6 instead of 8 DADD per quantity).
// The float64 box sums are exact, so their order is free; every float32 operation is the same IEEE operation on the
// same operands as before -- the packed instructions round each half independently.
__device__ __forceinline__ F2 f2_sub(F2 a, F2 b) { F2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ F2 f2_set(float v) { return f2_pack(v, v); }
__device__ __forceinline__ F2 f2_xor(F2 a, unsigned long long m) { F2 r; r.v = a.v ^ m; return r; }

template <bool kFma, bool BORDER, bool kIntF2D>
__device__ __forceinline__ void harris_strip3(const uint8_t *__restrict__ org, int ipitch, int W, int H, int x0, int y0w, int hr_rows, int gl_lanes,
                                              int n_groups, float k, const DetectScratch &det, int b, unsigned long long *buf, unsigned *cnt) {
    const int lane = threadIdx.x & 31;
    const int GL = BORDER ? gl_lanes : 32;
    const int grp = BORDER ? lane / GL : 0;
    const int gl = lane - grp * GL;
    const int y0 = y0w + grp * hr_rows;
    const bool active = !BORDER || (grp < n_groups && y0 < H);
    const int c0 = x0 - 4 + 4 * gl;
    const bool ld_ok = BORDER ? (active && (c0 + 3 <= W + 20) && (c0 >= -20)) : true;
    const bool out_lane = active && gl >= 1 && gl <= GL - 2;
    const double sc = 1.0 / (4.0 * 3.0 * 255.0);
    const float k0 = (float)sc, k1 = (float)(2.0 * sc);
    const F2 K0 = f2_set(k0), K1 = f2_set(k1), KK = f2_set(k);
    // per-column flags (BORDER only).  Pair E = columns (c0, c0+2), pair O = (c0+1, c0+3).
    int mir_col = -1;                                        // lane-local index of a mirrored column (x == -1 or x == W), if any
    unsigned long long flipE[2] = {0ull, 0ull}, flipO[2] = {0ull, 0ull};   // sign masks of gx*gy for rows inside / outside the image
    float cinv[4], cemit[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int x = c0 + i;
        const bool outside = BORDER && (x < 0 || x >= W);
        if (BORDER && (x == -1 || x == W)) mir_col = i;
        const unsigned long long bit = 0x80000000ull << ((i >> 1) * 32);
        if (outside) { if (i & 1) flipO[0] |= bit; else flipE[0] |= bit; }   // row inside: flip where the column is mirrored
        else if (BORDER) { if (i & 1) flipO[1] |= bit; else flipE[1] |= bit; }   // row outside: flip where the column is not
        cinv[i] = outside ? -INFINITY : 0.0f;
        cemit[i] = (!out_lane || (BORDER && (x < 1 || x >= W - 1))) ? -INFINITY : 0.0f;
    }
    const F2 cinvE = f2_pack(cinv[0], cinv[2]), cinvO = f2_pack(cinv[1], cinv[3]);

    // rings (slot = step % 3): Sobel row terms d, s (pairs E, O); products as doubles; responses
    F2 dE[3], dO[3], sE[3], sO[3];
    double qa[3][4], qb[3][4], qc[3][4];
    float Rr[3][6], Rx[3][4];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        dE[r] = dO[r] = sE[r] = sO[r] = f2_set(0.0f);
#pragma unroll
        for (int i = 0; i < 4; ++i) { qa[r][i] = 0.0; qb[r][i] = 0.0; qc[r][i] = 0.0; Rx[r][i] = -INFINITY; }
#pragma unroll
        for (int i = 0; i < 6; ++i) Rr[r][i] = -INFINITY;
    }
    float tmax = 0.0f;

    unsigned nstaged = 0;
    auto flush = [&]() {
        const unsigned nb = nstaged;
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(&det.cand_count[b], nb);
        base = __shfl_sync(0xffffffffu, base, 0);
        __syncwarp();
        for (unsigned i = lane; i < nb; i += 32) {
            const unsigned pos = base + i;
            if (pos < det.cand_cap) det.cand[(size_t)b * det.cand_cap + pos] = buf[i];
            else atomicExch(det.overflow, 1u);
        }
        __syncwarp();
        if (lane == 0) *cnt = 0u;
        nstaged = 0;
        __syncwarp();
    };

    const int rows = active ? min(hr_rows, H - y0) : 0;
    const int steps = min(hr_rows, H - y0w) + 6;
    const int my_steps = BORDER ? rows + 6 : steps;
    const uint8_t *rowp = org + (ptrdiff_t)(y0 - 3) * ipitch + c0;
    unsigned wq0 = 0, wq1 = 0;
    if (ld_ok) { wq0 = *reinterpret_cast<const unsigned *>(rowp); wq1 = *reinterpret_cast<const unsigned *>(rowp + ipitch); }
    rowp += 2 * (ptrdiff_t)ipitch;
    unsigned addr_row = (unsigned)((y0 - 6) * W + c0);

    auto step = [&](auto ph, int j) {
        constexpr int R0 = decltype(ph)::value, R1 = (R0 + 1) % 3, R2 = (R0 + 2) % 3;   // R0: this row; R1: two rows up; R2: one row up
        const unsigned w = wq0;
        wq0 = wq1;
        if (ld_ok && j + 2 < my_steps) wq1 = *reinterpret_cast<const unsigned *>(rowp);
        rowp += ipitch;
        const unsigned wl = __shfl_up_sync(0xffffffffu, w, 1), wr = __shfl_down_sync(0xffffffffu, w, 1);
        // pixel pairs as floats: Pa = (p[-1], p[1]), Pb = (p[0], p[2]), Pc = (p[1], p[3]), Pd = (p[2], p[4]) relative to c0
        const F2 M23 = f2_set(-8388608.0f);
        const F2 Pa = f2_add(f2_pack(__uint_as_float(__byte_perm(wl, 0x4B000000u, 0x7653u)), __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7651u))), M23);
        const F2 Pb = f2_add(f2_pack(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7650u)), __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7652u))), M23);
        const F2 Pc = f2_add(f2_pack(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7651u)), __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7653u))), M23);
        const F2 Pd = f2_add(f2_pack(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7652u)), __uint_as_float(__byte_perm(wr, 0x4B000000u, 0x7650u))), M23);
        // ---- Sobel row terms of this row (ring slot R0): pair E has taps (Pa, Pb, Pc), pair O has (Pb, Pc, Pd)
        dE[R0] = f2_sub(Pc, Pa);
        dO[R0] = f2_sub(Pd, Pb);
        // ptxas (12.9) contracts a packed multiply feeding a packed add into FFMA2 even for explicit .rn operations and with
        // -fmad=false (scalar mul.rn / add.rn are respected), so wherever the reference rounds a product before adding
        // it the multiplies are packed and the ADDS are scalar.
        const F2 k0a = f2_mul(K0, Pa), k0b = f2_mul(K0, Pb), k0c = f2_mul(K0, Pc), k0d = f2_mul(K0, Pd);
        if (!kFma) {
            const F2 k1b = f2_mul(K1, Pb), k1c = f2_mul(K1, Pc);
            float a0, a1, b0, b1, c0_, c1, d0, d1, m0, m1, n0, n1;
            f2_unpack(k0a, a0, a1); f2_unpack(k0b, b0, b1); f2_unpack(k0c, c0_, c1); f2_unpack(k0d, d0, d1);
            f2_unpack(k1b, m0, m1); f2_unpack(k1c, n0, n1);
            sE[R0] = f2_pack(__fadd_rn(__fadd_rn(a0, m0), c0_), __fadd_rn(__fadd_rn(a1, m1), c1));
            sO[R0] = f2_pack(__fadd_rn(__fadd_rn(b0, n0), d0), __fadd_rn(__fadd_rn(b1, n1), d1));
        } else {
            sE[R0] = f2_fma(K0, Pc, f2_fma(K1, Pb, k0a));
            sO[R0] = f2_fma(K0, Pd, f2_fma(K1, Pc, k0b));
        }
        if (BORDER && mir_col >= 0) {
            // the column x == -1 / x == W: the 3-tap smoothing in mirrored order (it is not associative)
            float pl, pc, pr2, e0, e1;
            const bool odd = mir_col & 1, hi = mir_col >> 1;
            f2_unpack(odd ? Pb : Pa, e0, e1); pl = hi ? e1 : e0;
            f2_unpack(odd ? Pc : Pb, e0, e1); pc = hi ? e1 : e0;
            f2_unpack(odd ? Pd : Pc, e0, e1); pr2 = hi ? e1 : e0;
            float sv;
            if (!kFma) sv = ((k0 * pr2) + k1 * pc) + k0 * pl;
            else sv = __fmaf_rn(k0, pl, __fmaf_rn(k1, pc, k0 * pr2));
            F2 &dst = odd ? sO[R0] : sE[R0];
            f2_unpack(dst, e0, e1);
            dst = hi ? f2_pack(e0, sv) : f2_pack(sv, e1);
        }
        // ---- gradients of the row above (centre R2: rows R1, R2, R0), products, exact box sums
        const int pr = y0 - 4 + j;
        const bool bflipy = BORDER && ((pr < 0) || (pr >= H));
        F2 gxE, gxO;
        if (!kFma) {
            const F2 uE = f2_mul(K0, f2_add(dE[R1], dE[R0])), vE = f2_mul(K1, dE[R2]);
            const F2 uO = f2_mul(K0, f2_add(dO[R1], dO[R0])), vO = f2_mul(K1, dO[R2]);
            float u0, u1, v0, v1;
            f2_unpack(uE, u0, u1); f2_unpack(vE, v0, v1); gxE = f2_pack(__fadd_rn(v0, u0), __fadd_rn(v1, u1));
            f2_unpack(uO, u0, u1); f2_unpack(vO, v0, v1); gxO = f2_pack(__fadd_rn(v0, u0), __fadd_rn(v1, u1));
        } else {
            gxE = f2_fma(K0, f2_add(dE[R1], dE[R0]), f2_mul(K1, dE[R2]));
            gxO = f2_fma(K0, f2_add(dO[R1], dO[R0]), f2_mul(K1, dO[R2]));
        }
        const F2 gyE = f2_sub(sE[R0], sE[R1]), gyO = f2_sub(sO[R0], sO[R1]);
        F2 fbE = f2_mul(gxE, gyE), fbO = f2_mul(gxO, gyO);
        if (BORDER) { fbE = f2_xor(fbE, flipE[bflipy ? 1 : 0]); fbO = f2_xor(fbO, flipO[bflipy ? 1 : 0]); }
        const F2 faE = f2_mul(gxE, gxE), faO = f2_mul(gxO, gxO), fcE = f2_mul(gyE, gyE), fcO = f2_mul(gyO, gyO);
        {
            float e0, e1;
            // float -> double: F2F.F64.F32 runs at ~5 results / clk / SM and bounds the kernel; the integer re-biasing
            // (harris_exact.cuh) moves it to the ALU pipe
            auto cv = [](float f) { return kIntF2D ? harris_f2d(f) : (double)f; };
            auto cvn = [](float f) { return kIntF2D ? harris_f2d_nonneg(f) : (double)f; };
            f2_unpack(faE, e0, e1); qa[R0][0] = cvn(e0); qa[R0][2] = cvn(e1);
            f2_unpack(faO, e0, e1); qa[R0][1] = cvn(e0); qa[R0][3] = cvn(e1);
            f2_unpack(fbE, e0, e1); qb[R0][0] = cv(e0); qb[R0][2] = cv(e1);
            f2_unpack(fbO, e0, e1); qb[R0][1] = cv(e0); qb[R0][3] = cv(e1);
            f2_unpack(fcE, e0, e1); qc[R0][0] = cvn(e0); qc[R0][2] = cvn(e1);
            f2_unpack(fcO, e0, e1); qc[R0][1] = cvn(e0); qc[R0][3] = cvn(e1);
        }
        double va[6], vb[6], vc[6];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            va[i + 1] = (qa[R1][i] + qa[R2][i]) + qa[R0][i];
            vb[i + 1] = (qb[R1][i] + qb[R2][i]) + qb[R0][i];
            vc[i + 1] = (qc[R1][i] + qc[R2][i]) + qc[R0][i];
        }
        va[0] = shfl_up_d(va[4]); vb[0] = shfl_up_d(vb[4]); vc[0] = shfl_up_d(vc[4]);
        va[5] = shfl_down_d(va[1]); vb[5] = shfl_down_d(vb[1]); vc[5] = shfl_down_d(vc[1]);
        float A[4], B[4], C[4];
        {
            const double a12 = va[1] + va[2], a34 = va[3] + va[4], b12 = vb[1] + vb[2], b34 = vb[3] + vb[4];
            const double c12 = vc[1] + vc[2], c34 = vc[3] + vc[4];
            A[0] = (float)(va[0] + a12); A[1] = (float)(a12 + va[3]); A[2] = (float)(va[2] + a34); A[3] = (float)(a34 + va[5]);
            B[0] = (float)(vb[0] + b12); B[1] = (float)(b12 + vb[3]); B[2] = (float)(vb[2] + b34); B[3] = (float)(b34 + vb[5]);
            C[0] = (float)(vc[0] + c12); C[1] = (float)(c12 + vc[3]); C[2] = (float)(vc[2] + c34); C[3] = (float)(c34 + vc[5]);
        }
        // ---- response of row q = pr - 1 (ring slot R0 of the response rings)
        const int q = pr - 1;
        {
            const F2 A01 = f2_pack(A[0], A[1]), A23 = f2_pack(A[2], A[3]), B01 = f2_pack(B[0], B[1]), B23 = f2_pack(B[2], B[3]);
            const F2 C01 = f2_pack(C[0], C[1]), C23 = f2_pack(C[2], C[3]);
            const F2 T01 = f2_add(A01, C01), T23 = f2_add(A23, C23);
            // products packed, differences scalar (see above: a packed product must not feed a packed add)
            const F2 ac01 = f2_mul(A01, C01), ac23 = f2_mul(A23, C23), bb01 = f2_mul(B01, B01), bb23 = f2_mul(B23, B23);
            const F2 kt01 = kFma ? f2_mul(KK, f2_mul(T01, T01)) : f2_mul(f2_mul(KK, T01), T01);
            const F2 kt23 = kFma ? f2_mul(KK, f2_mul(T23, T23)) : f2_mul(f2_mul(KK, T23), T23);
            float x0_, x1_, y0_, y1_, z0_, z1_;
            f2_unpack(ac01, x0_, x1_); f2_unpack(bb01, y0_, y1_); f2_unpack(kt01, z0_, z1_);
            F2 r01 = f2_pack(__fsub_rn(__fsub_rn(x0_, y0_), z0_), __fsub_rn(__fsub_rn(x1_, y1_), z1_));
            f2_unpack(ac23, x0_, x1_); f2_unpack(bb23, y0_, y1_); f2_unpack(kt23, z0_, z1_);
            F2 r23 = f2_pack(__fsub_rn(__fsub_rn(x0_, y0_), z0_), __fsub_rn(__fsub_rn(x1_, y1_), z1_));
            if (BORDER) {
                // dilate ignores pixels outside the image: -inf there (v + 0 is exact, v - inf = -inf)
                const float rinv = (q < 0 || q >= H) ? -INFINITY : 0.0f;
                r01 = f2_add(r01, f2_pack(cinv[0] + rinv, cinv[1] + rinv));
                r23 = f2_add(r23, f2_pack(cinv[2] + rinv, cinv[3] + rinv));
            }
            f2_unpack(r01, Rr[R0][1], Rr[R0][2]);
            f2_unpack(r23, Rr[R0][3], Rr[R0][4]);
        }
        Rr[R0][0] = __shfl_up_sync(0xffffffffu, Rr[R0][4], 1);
        Rr[R0][5] = __shfl_down_sync(0xffffffffu, Rr[R0][1], 1);
#pragma unroll
        for (int i = 0; i < 4; ++i) Rx[R0][i] = fmax3(Rr[R0][i], Rr[R0][i + 1], Rr[R0][i + 2]);
        // ---- NMS for row n = q-1 (raw row in slot R2): rows n-1 (slot R1, max3), n, n+1 (slot R0, max3)
        const int n = q - 1;
        if (j >= 6) {
            const bool nrow_ok = (n < y0 + rows) && (!BORDER || ((n >= 1) && (n < H - 1)));
            unsigned cmask = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float v = Rr[R2][i + 1];
                const float m = fmax3(Rx[R1][i], Rx[R0][i], fmaxf(Rr[R2][i], Rr[R2][i + 2]));
                const float ve = v + cemit[i];               // -inf where this lane/column may not emit
                if (ve > 0.0f && ve >= m) cmask |= 1u << i;
            }
            // frame maximum over every row of the strip (the rounded-up trip count must not leak rows computed from stale
            // pixels); -inf outside the image never wins
            if (out_lane && n < y0 + rows) { tmax = fmax3(tmax, Rr[R2][1], Rr[R2][2]); tmax = fmax3(tmax, Rr[R2][3], Rr[R2][4]); }
            if (!nrow_ok) cmask = 0;
            const unsigned nmine = __popc(cmask);
            if (nmine) {
                unsigned pos = atomicAdd(cnt, nmine);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (cmask & (1u << i)) buf[pos] = ((unsigned long long)__float_as_uint(Rr[R2][i + 1]) << 32) | (addr_row + (unsigned)i);
                    pos += (cmask >> i) & 1u;
                }
            }
            nstaged += __reduce_add_sync(0xffffffffu, nmine);
            if (nstaged > (unsigned)(HW_BUF - 128)) flush();    // warp-uniform: a step adds at most 120 keys
        }
        addr_row += (unsigned)W;
    };
    const int steps3 = (steps + 2) / 3 * 3;                  // the up to 2 extra steps read nothing and emit nothing
#pragma unroll 1
    for (int j = 0; j < steps3; j += 3) {
        step(IntC<0>{}, j);
        step(IntC<1>{}, j + 1);
        step(IntC<2>{}, j + 2);
    }
    if (nstaged) flush();
    const unsigned mb = __reduce_max_sync(0xffffffffu, __float_as_uint(fmaxf(tmax, 0.0f)));
    if (lane == 0 && mb) atomicMax(&det.frame_max[b], mb);
}

template <bool kFma, bool kIntF2D>
__global__ void __launch_bounds__(HW_WARPS * 32)
harris_nms3_kernel(Pyramid pyr, SlotList slots, float k, DetectScratch det, int tiles_x, int strips, int gl_narrow, int n_items, int hr_rows) {
    __shared__ unsigned long long s_buf[HW_WARPS][HW_BUF];
    __shared__ unsigned s_cnt[HW_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int W = pyr.lv[0].w, H = pyr.lv[0].h, ipitch = pyr.lv[0].ipitch;
    const int total = n_items * slots.n;
    for (int wi = blockIdx.x * HW_WARPS + warp; wi < total; wi += gridDim.x * HW_WARPS) {
        const int b = wi / n_items, item = wi - b * n_items, slot = slots.v[b];
        int x0, y0, gl_lanes = 32, n_groups = 1;
        if (item < tiles_x * strips) { x0 = (item % tiles_x) * HR_COLS; y0 = (item / tiles_x) * hr_rows; }
        else {
            gl_lanes = gl_narrow; n_groups = gl_narrow == 10 ? 3 : 2;
            x0 = tiles_x * HR_COLS; y0 = (item - tiles_x * strips) * n_groups * hr_rows;
        }
        const uint8_t *org = pyr.image_origin(0, slot);
        if (lane == 0) s_cnt[warp] = 0u;
        __syncwarp();
        const bool interior = (x0 - 5 >= 0) && (x0 + 124 < W) && (y0 - 6 >= 0) && (y0 + hr_rows + 2 < H);
        if (interior) harris_strip3<kFma, false, kIntF2D>(org, ipitch, W, H, x0, y0, hr_rows, 32, 1, k, det, b, s_buf[warp], &s_cnt[warp]);
        else harris_strip3<kFma, true, kIntF2D>(org, ipitch, W, H, x0, y0, hr_rows, gl_lanes, n_groups, k, det, b, s_buf[warp], &s_cnt[warp]);
    }
}

// Experiment switches (environment, read once): RDFE_HARRIS_ROWS = strip height at full batches, RDFE_HARRIS_NARROW=0
// walks the narrow last tile one strip per warp like the full tiles.
static int harris_max_rows() {
    static const int v = [] { const char *e = getenv("RDFE_HARRIS_ROWS"); const int r = e ? atoi(e) : 0; return r >= 8 && r <= 1080 ? (r + 3) & ~3 : HR_ROWS; }();
    return v;
}
static bool harris_narrow_enabled() {
    static const bool v = [] { const char *e = getenv("RDFE_HARRIS_NARROW"); return !(e && e[0] == '0'); }();
    return v;
}
// RDFE_HARRIS_IMPL selects one of three implementations that emit identical keys (tests/test_gpu_harris_prefilter.py and the
// detect / golden / cv2 tests pass with each; measured on B200, 64 frames of 752x480, profiles/r2_harris_variants.md):
//   0 (default) harris_nms_kernel: exact float64 chain at every pixel.  75.4 M warp instructions, 127 us.
//   1 harris_flag_kernel + harris_resolve_kernel: integer prefilter, exact chain at the ~6.7 % flagged pixels only.
//     41.2 M + 33 M warp instructions, 82 + 60 us: an isolated exact evaluation costs ~600 instructions against ~90 per
//     pixel inside the rolling kernel, so the prefilter only pays below ~6 % flagged pixels (CLAHE output is above).
//   3 harris_nms3_kernel: the exact chain with packed FADD2/FMUL2, ring-3 unrolling, FMNMX3.  60 M warp instructions
//     (-20 %) but 151 us: all three are bound by dependent-issue latency at 12 warps per SM (eligible warps 0.6-1.1 per
//     scheduler), not by issue slots or any pipe (alu 41 %, xu 34 %, fma 21 %, fp64 16 % for impl 0).
static int harris_impl() {
    static const int v = [] { const char *e = getenv("RDFE_HARRIS_IMPL"); const int r = e ? atoi(e) : 0; return (r == 1 || r == 3) ? r : 0; }();
    return v;
}

__global__ void detect_reset_kernel(DetectScratch det, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { det.cand_count[i] = 0u; det.frame_max[i] = 0u; det.flag_count[i] = 0u; }
}

int launch_harris_candidates(rdfe_ctx *ctx, const SlotList &slots, const rdfe_detect_params &p, float *d_response) {
    const LevelGeom &g = ctx->pyr.lv[0];
    detect_reset_kernel<<<1, RDFE_MAX_BATCH, 0, ctx->ls>>>(ctx->det, slots.n);
    const int tiles_all = (g.w + HR_COLS - 1) / HR_COLS;
    const int hr_rows = adaptive_strip_rows(g.h, tiles_all * slots.n, 8, harris_max_rows());
    const int strips = (g.h + hr_rows - 1) / hr_rows;
    // narrow last tile: <= 32 (56) leftover columns are walked by 3 (2) lane groups, one strip each (harris_strip)
    const int left = g.w - (tiles_all - 1) * HR_COLS;
    const int gl_narrow = harris_narrow_enabled() ? (left <= 32 ? 10 : left <= 56 ? 16 : 0) : 0;
    const int n_groups = gl_narrow == 10 ? 3 : 2;
    const int tiles_x = gl_narrow ? tiles_all - 1 : tiles_all;          // full-width tiles
    const int n_items = tiles_x * strips + (gl_narrow ? (strips + n_groups - 1) / n_groups : 0);
    const int total = n_items * slots.n;
    dim3 grid((total + HW_WARPS - 1) / HW_WARPS);     // one warp per (image, strip); the kernel's loop also accepts fewer
    if (!d_response && harris_impl() == 1) {
        // RDFE_HARRIS_MB=3: 153 registers, no spills, 12 warps per SM; 4 (default): 128 registers, 16 warps per SM
        static const int minb = [] { const char *e = getenv("RDFE_HARRIS_MB"); return e ? atoi(e) : 4; }();
        if (minb == 3)
            RDFE_LAUNCH(ctx, K_HARRIS, (harris_flag_kernel<3><<<grid, HW_WARPS * 32, 0, ctx->ls>>>(ctx->pyr, slots, ctx->det, tiles_x, strips, gl_narrow, n_items, hr_rows)));
        else
            RDFE_LAUNCH(ctx, K_HARRIS, (harris_flag_kernel<4><<<grid, HW_WARPS * 32, 0, ctx->ls>>>(ctx->pyr, slots, ctx->det, tiles_x, strips, gl_narrow, n_items, hr_rows)));
        // a warp of the resolve kernel takes `cpw` chunks of 1024 pixels; enough warps to fill the GPU at small batches
        const int nwords = (harris_bm_pitch(g.w) >> 3) * g.h, nchunks = (nwords + 31) / 32;
        int cpw = (nchunks * slots.n + kSMs * 16 - 1) / (kSMs * 16);
        cpw = cpw < 1 ? 1 : cpw > 4 ? 4 : cpw;
        const int warps = (nchunks + cpw - 1) / cpw;
        dim3 g2((warps + HV_THREADS / 32 - 1) / (HV_THREADS / 32), slots.n);
        if (p.harris_fma)
            RDFE_LAUNCH(ctx, K_HARRIS_RESOLVE, (harris_resolve_kernel<true><<<g2, HV_THREADS, 0, ctx->ls>>>(ctx->pyr, slots, (float)p.harris_k, ctx->det, cpw)));
        else
            RDFE_LAUNCH(ctx, K_HARRIS_RESOLVE, (harris_resolve_kernel<false><<<g2, HV_THREADS, 0, ctx->ls>>>(ctx->pyr, slots, (float)p.harris_k, ctx->det, cpw)));
        return 3;
    }
    if (!d_response && harris_impl() == 3) {
        // RDFE_HARRIS_F2D=0: hardware F2F for float -> double (the conversion unit then bounds the kernel)
        static const bool intcv = [] { const char *e = getenv("RDFE_HARRIS_F2D"); return !(e && e[0] == '0'); }();
#define RDFE_H3(FMA, CV) RDFE_LAUNCH(ctx, K_HARRIS, (harris_nms3_kernel<FMA, CV><<<grid, HW_WARPS * 32, 0, ctx->ls>>>(ctx->pyr, slots, (float)p.harris_k, ctx->det, tiles_x, strips, gl_narrow, n_items, hr_rows)))
        if (p.harris_fma) { if (intcv) RDFE_H3(true, true); else RDFE_H3(true, false); }
        else { if (intcv) RDFE_H3(false, true); else RDFE_H3(false, false); }
#undef RDFE_H3
        return 2;
    }
    if (p.harris_fma)
        if (d_response)
            RDFE_LAUNCH(ctx, K_HARRIS, (harris_nms_kernel<true, true><<<grid, HW_WARPS * 32, 0, ctx->ls>>>(ctx->pyr, slots, (float)p.harris_k, ctx->det, d_response, tiles_x, strips, gl_narrow, n_items, hr_rows)));
        else
            RDFE_LAUNCH(ctx, K_HARRIS, (harris_nms_kernel<true, false><<<grid, HW_WARPS * 32, 0, ctx->ls>>>(ctx->pyr, slots, (float)p.harris_k, ctx->det, nullptr, tiles_x, strips, gl_narrow, n_items, hr_rows)));
    else
        if (d_response)
            RDFE_LAUNCH(ctx, K_HARRIS, (harris_nms_kernel<false, true><<<grid, HW_WARPS * 32, 0, ctx->ls>>>(ctx->pyr, slots, (float)p.harris_k, ctx->det, d_response, tiles_x, strips, gl_narrow, n_items, hr_rows)));
        else
            RDFE_LAUNCH(ctx, K_HARRIS, (harris_nms_kernel<false, false><<<grid, HW_WARPS * 32, 0, ctx->ls>>>(ctx->pyr, slots, (float)p.harris_k, ctx->det, nullptr, tiles_x, strips, gl_narrow, n_items, hr_rows)));
    return 2;
}

}  // namespace rdfe
