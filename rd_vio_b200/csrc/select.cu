// select.cu -- K3b: corner selection, one CTA per image.
// Replaces the sequential half of cv::goodFeaturesToTrack and the host
// glue of OpenCvImage::detect_keypoints (src/rdvio_extra/src/opencv_image.cpp:46-72):
//   1. thr = float(max(R) * qualityLevel); keep candidates with R > thr;
//   2. visit them in (R desc, address desc) order -- OpenCV's greaterThanPtr -- in batches:
//      an exact 64-bit radix select picks the next SEL_CAP largest keys, a shared-memory
//      bitonic sort orders them;
//   3. greedy min-distance acceptance on the round(minDistance) cell grid (3x3 cells,
//      dx*dx+dy*dy < minDistance^2 rejects), stopping at maxCorners.  One warp walks the
//      sorted batch 32 candidates at a time: grid test in parallel, then the in-chunk order
//      dependence is resolved lane by lane, which reproduces the sequential result exactly;
//   4. extra::PoissonDiskFilter<2> semantics (poisson_disk_filter.h:23-94: one point per
//      cell, last writer wins; 5x5 cell block minus its first cell plus one past the end;
//      reject iff squared distance < r^2; float64) against the caller's existing keypoints,
//      then the 20-px border reject, then append.
#include "fe_internal.cuh"
#include "harris_exact.cuh"

namespace rdfe {

constexpr int SEL_THREADS = 1024;
constexpr int SEL_CAP = 2048;
constexpr int SEL_WIN = 8;           // greedy window: words of 32 sorted candidates tested per round

struct SelectParams {
    int W, H;
    int max_corners;
    double quality;
    float min_dist2;
    int cell, gw, gh;        // GFTT grid
    int use_min_dist;
    double kp_radius;
    int border;
    int stride;              // keypoint capacity per image
    int cap_k;               // accepted-corner capacity (= max_corners)
    float harris_k;          // for the exact fallback of degenerate frames
    int harris_fma, prefiltered;
};

__device__ __forceinline__ bool poisson_block_hit(int dcx, int dcy, int span) {
    // cells visited by PoissonDiskFilter::test_point relative to the query cell
    if (dcx < -span || dcx > span) return false;
    if (dcy >= -span && dcy <= span) return !(dcx == -span && dcy == -span);
    return dcx == -span && dcy == span + 1;
}

// true iff no accepted corner lies within minDistance of (x, y): 3x3 cells, 4 in-place slots each
__device__ __forceinline__ bool grid_pass(const uint4 *cellpts, int gw, int gh, int cell, float md2, int x, int y) {
    const int xc = x / cell, yc = y / cell;
    const int x1 = max(xc - 1, 0), y1 = max(yc - 1, 0), x2 = min(xc + 1, gw - 1), y2 = min(yc + 1, gh - 1);
    bool ok = true;
    for (int yy = y1; yy <= y2; ++yy)
        for (int xx = x1; xx <= x2; ++xx) {
            const uint4 c = cellpts[yy * gw + xx];
            const unsigned pv[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int dx = x - (int)(pv[q] & 0xFFFFu), dy = y - (int)(pv[q] >> 16);
                if (pv[q] != ~0u && (float)(dx * dx + dy * dy) < md2) ok = false;
            }
        }
    return ok;
}

// 32 registers (2 CTAs of 1024 threads per SM): the kernel itself is 7 % slower than at 64, but a CTA then takes half of an
// SM's register file instead of all of it, and the kernels of the other streams run beside it: +1.3 % frames/s in the
// pipelined step (A/B on one box, 176.8 k -> 179.0 k, twice)
__global__ void __launch_bounds__(SEL_THREADS, 2)
select_kernel(DetectScratch det, SelectParams sp, Pyramid pyr, SlotList slots, float *__restrict__ gftt_xy,
              float *__restrict__ gftt_resp, int *__restrict__ gftt_counts) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long *batch = reinterpret_cast<unsigned long long *>(smem_raw);          // [SEL_CAP]
    unsigned *hist = reinterpret_cast<unsigned *>(batch + SEL_CAP);                         // [256]
    float *ax = reinterpret_cast<float *>(hist + 256);                                      // [cap_k]
    float *ay = ax + sp.cap_k;
    float *ar = ay + sp.cap_k;
    // accepted-corner grid: up to 4 corners per cell stored in place as (x | y << 16); a cell of side
    // round(minDistance) can hold at most 3 points that are pairwise >= minDistance apart.
    uint4 *cellpts = reinterpret_cast<uint4 *>(smem_raw + (((size_t)SEL_CAP * 8 + 256 * 4 + (size_t)sp.cap_k * 12 + 15) & ~(size_t)15));
    int *cellcnt = reinterpret_cast<int *>(cellpts + sp.gw * sp.gh);                        // [gw*gh]

    __shared__ unsigned s_count, s_nb;
    __shared__ unsigned long long s_prefix, s_hi;
    __shared__ unsigned s_k;
    __shared__ int s_naccepted, s_done;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x;
    unsigned n = min(det.cand_count[b], det.cand_cap);
    float maxv = __uint_as_float(det.frame_max[b]);
    float thr = (float)((double)maxv * sp.quality);
    if (sp.prefiltered && thr < kHarrisRhoS) {
        // Degenerate frame (no response above the rounding residue of vanishing integer gradients, e.g. a blank
        // frame): the prefilter's candidate list is not trustworthy below kHarrisRhoS -- recompute the frame with the
        // reference's exact arithmetic at every pixel (harris_exact.cuh).  CTA-uniform branch.
        __shared__ unsigned s_fb_count, s_fb_max;
        harris_exact_all(pyr.image_origin(0, slots.v[b]), pyr.lv[0].ipitch, sp.W, sp.H, sp.harris_k, sp.harris_fma != 0,
                         reinterpret_cast<float *>(det.cand2 + (size_t)b * det.cand_cap), det.cand + (size_t)b * det.cand_cap,
                         det.cand_cap, &s_fb_count, &s_fb_max, det.overflow);
        n = min(s_fb_count, det.cand_cap);
        maxv = __uint_as_float(s_fb_max);
        thr = (float)((double)maxv * sp.quality);
    }
    const unsigned thr_bits = __float_as_uint(thr);       // thr >= 0 here (maxv >= 0)
    const int W = sp.W;

    for (int i = tid; i < sp.gw * sp.gh; i += SEL_THREADS) { cellpts[i] = make_uint4(~0u, ~0u, ~0u, ~0u); cellcnt[i] = 0; }
    if (tid == 0) { s_naccepted = 0; s_done = 0; s_hi = ~0ull; }
    __syncthreads();

    // address bits that can be non-zero: skip radix passes above them
    const int addr_bits = 32 - __clz(max(sp.W * sp.H - 1, 1));
    const int low_passes = (addr_bits + 7) / 8;

    // The candidate list is re-compacted at the start of every round: keys that are not eligible any more
    // (R <= thr, already visited) or that lie within minDistance of an already accepted corner can never be
    // accepted later (the accepted set only grows), so dropping them in parallel is exact and shrinks the
    // sequential part to the few candidates that still matter.
    unsigned long long *bufA = det.cand + (size_t)b * det.cand_cap;
    unsigned long long *bufB = det.cand2 + (size_t)b * det.cand_cap;
    const unsigned long long *src = bufA;
    unsigned ncur = n;
    int round = 0;
    while (true) {
        const unsigned long long hi = s_hi;
        const int nacc0 = s_naccepted;
        if (tid == 0) s_count = 0;
        __syncthreads();
        unsigned long long *dst = (round & 1) ? bufA : bufB;
        // four keys per lane are requested before the first one is used: the pass is bound by the latency of these loads
        for (unsigned base0 = warp * 32; base0 < ncur; base0 += 4 * SEL_THREADS) {
            unsigned long long kq[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const unsigned i = base0 + u * SEL_THREADS + lane;
                kq[u] = (i < ncur) ? src[i] : 0ull;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const unsigned long long k = kq[u];
                bool ok = (unsigned)(k >> 32) > thr_bits && k < hi;      // k = 0 (beyond the list) fails the first test
                if (ok && sp.use_min_dist && nacc0 > 0) {
                    const unsigned addr = (unsigned)k;
                    const int y = (int)(addr / (unsigned)W), x = (int)(addr - (unsigned)y * (unsigned)W);
                    ok = grid_pass(cellpts, sp.gw, sp.gh, sp.cell, sp.min_dist2, x, y);
                }
                const unsigned m = __ballot_sync(0xffffffffu, ok);
                if (m) {
                    unsigned pos = 0;
                    if (lane == 0) pos = atomicAdd(&s_count, (unsigned)__popc(m));
                    pos = __shfl_sync(0xffffffffu, pos, 0) + __popc(m & ((1u << lane) - 1u));
                    if (ok) dst[pos] = k;
                }
            }
        }
        __syncthreads();
        const unsigned n_el = s_count;
        src = dst;
        ncur = n_el;
        ++round;
        if (n_el == 0) break;
        const unsigned n = ncur;                      // every key of the compacted list is eligible
        const unsigned long long *keys = src;

        // Bitonic sort of batch[0, cnt) (descending), used for the pivot sample and for the batch itself.
        auto sort_batch = [&](unsigned cnt) {
            unsigned N = 32;
            while (N < cnt) N <<= 1;
            for (unsigned i = cnt + tid; i < N; i += SEL_THREADS) batch[i] = 0ull;
            __syncthreads();
            // Bitonic sort, descending.  Thread t owns elements 2t and 2t+1 for the exchange distances j <= 32 (its
            // partner for distance j is lane t ^ (j/2): register shuffles, no barrier); distances >= 64 go through
            // shared memory.  All merges up to k2 = 64 stay inside one warp's 64 elements.
            {
                const unsigned half = N >> 1;
                const bool warp_on = (unsigned)(warp * 32) < half;            // warp-uniform
                const bool own = (unsigned)tid < half;
                const unsigned e_idx = 2u * (unsigned)tid;
                unsigned long long e0 = 0ull, e1 = 0ull;
                auto reg_stages = [&](unsigned k2, unsigned jstart) {
                    const bool desc = (e_idx & k2) == 0;
                    for (unsigned j = jstart; j >= 2; j >>= 1) {
                        const unsigned long long p0 = __shfl_xor_sync(0xffffffffu, e0, (int)(j >> 1));
                        const unsigned long long p1 = __shfl_xor_sync(0xffffffffu, e1, (int)(j >> 1));
                        const bool keep_max = ((e_idx & j) == 0) == desc;
                        e0 = keep_max ? (e0 > p0 ? e0 : p0) : (e0 < p0 ? e0 : p0);
                        e1 = keep_max ? (e1 > p1 ? e1 : p1) : (e1 < p1 ? e1 : p1);
                    }
                    const unsigned long long hi = e0 > e1 ? e0 : e1, lo = e0 > e1 ? e1 : e0;
                    e0 = desc ? hi : lo;
                    e1 = desc ? lo : hi;
                };
                if (warp_on) {
                    if (own) { const ulonglong2 v = reinterpret_cast<const ulonglong2 *>(batch)[tid]; e0 = v.x; e1 = v.y; }
                    for (unsigned k2 = 2; k2 <= min(N, 64u); k2 <<= 1) reg_stages(k2, k2 >> 1);
                    if (own) reinterpret_cast<ulonglong2 *>(batch)[tid] = make_ulonglong2(e0, e1);
                }
                __syncthreads();
                for (unsigned k2 = 128; k2 <= N; k2 <<= 1) {
                    for (unsigned j = k2 >> 1, lj = 31 - __clz(k2 >> 1); j >= 64; j >>= 1, --lj) {
                        if (own) {
                            const unsigned t = (unsigned)tid;
                            const unsigned i = ((t >> lj) << (lj + 1)) | (t & (j - 1u));
                            const unsigned q = i + j;
                            const unsigned long long a = batch[i], c = batch[q];
                            const bool desc = ((i & k2) == 0);
                            if (desc ? (a < c) : (a > c)) { batch[i] = c; batch[q] = a; }
                        }
                        __syncthreads();
                    }
                    if (warp_on) {
                        if (own) { const ulonglong2 v = reinterpret_cast<const ulonglong2 *>(batch)[tid]; e0 = v.x; e1 = v.y; }
                        reg_stages(k2, 32u);
                        if (own) reinterpret_cast<ulonglong2 *>(batch)[tid] = make_ulonglong2(e0, e1);
                    }
                    __syncthreads();
                }
            }
        };
        // ---- pivot for the next batch.  ANY threshold lo gives an exact batch (the keys >= lo are a prefix of the
        // descending order) as long as at most SEL_CAP keys pass it, so the pivot comes from a sorted sample of <= SEL_CAP
        // keys taken at a regular stride (one short pass + one small sort instead of four full histogram passes); the
        // gather below counts what passes, and if a pivot lets too many through the next, stricter one is tried; the exact
        // radix select remains as the last resort.
        unsigned long long lo = 0;
        unsigned long long pv[3] = {0ull, 0ull, 0ull};
        int npv = 0;
        if (n_el > SEL_CAP) {
            const unsigned stride_s = (n_el + SEL_CAP - 1) / SEL_CAP;
            const unsigned ns = (n_el + stride_s - 1) / stride_s;
            for (unsigned i = tid; i < ns; i += SEL_THREADS) batch[i] = keys[(size_t)i * stride_s];
            __syncthreads();
            sort_batch(ns);
            // expected number of keys >= sample[r] is about (r + 1) * stride: aim at 5/8, then 5/16, then 5/64 of the batch
            const unsigned r0 = (SEL_CAP * 5u / 8u) / stride_s;
            pv[0] = batch[min(max(r0, 1u), ns - 1u)];
            pv[1] = batch[min(max(r0 / 2u, 1u), ns - 1u)];
            pv[2] = batch[min(max(r0 / 8u, 1u), ns - 1u)];
            npv = 3;
            __syncthreads();
        }
        bool gathered = false;
        for (int tr = 0; tr < npv && !gathered; ++tr) {
            if (tid == 0) s_nb = 0;
            __syncthreads();
            const unsigned long long pivot = pv[tr];
            for (unsigned i0 = tid; i0 < n; i0 += 4 * SEL_THREADS) {
                unsigned long long kq[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) kq[u] = (i0 + u * SEL_THREADS < n) ? keys[i0 + u * SEL_THREADS] : 0ull;
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (kq[u] >= pivot && kq[u] != 0ull) {
                        const unsigned pos = atomicAdd(&s_nb, 1u);
                        if (pos < SEL_CAP) batch[pos] = kq[u];
                    }
            }
            __syncthreads();
            gathered = s_nb <= SEL_CAP;                    // block-uniform
            __syncthreads();
        }
        if (!gathered) {
            // ---- exact radix select of the SEL_CAP-th largest eligible key (last resort, and the whole list when it fits)
            if (n_el > SEL_CAP) {
                if (tid == 0) { s_prefix = 0; s_k = SEL_CAP; }
                unsigned long long mask = 0;
                bool early = false;
                for (int pass = 7; pass >= 0; --pass) {
                    if (pass < 4 && pass >= low_passes) continue;      // digits known to be zero
                    const int shift = pass * 8;
                    if (tid < 256) hist[tid] = 0;
                    __syncthreads();
                    const unsigned long long prefix = s_prefix;
                    // Equal digits are the norm in the high passes (responses share their exponent byte): when a whole
                    // warp agrees on the digit one lane adds the count, otherwise plain shared-memory atomics.
                    for (unsigned base = warp * 32; base < n; base += SEL_THREADS) {
                        const unsigned i = base + lane;
                        unsigned long long k = 0;
                        bool part = false;
                        if (i < n) { k = keys[i]; part = (k & mask) == prefix; }
                        const unsigned d = (unsigned)(k >> shift) & 255u;
                        // fast path: every participating lane holds the same digit (typical for the exponent bytes)
                        const unsigned pm = __ballot_sync(0xffffffffu, part);
                        if (pm) {
                            const unsigned d0 = __shfl_sync(0xffffffffu, d, __ffs(pm) - 1);
                            const bool same = __all_sync(0xffffffffu, !part || d == d0);
                            if (same) { if (lane == __ffs(pm) - 1) atomicAdd(&hist[d0], (unsigned)__popc(pm)); }
                            else if (part) atomicAdd(&hist[d], 1u);
                        }
                    }
                    __syncthreads();
                    if (warp == 0) {
                        // lane l owns digits 255-8l .. 248-8l (descending)
                        unsigned loc[8], tot = 0;
#pragma unroll
                        for (int j = 0; j < 8; ++j) { loc[j] = hist[255 - 8 * lane - j]; tot += loc[j]; }
                        unsigned inc = tot;
#pragma unroll
                        for (int d = 1; d < 32; d <<= 1) {
                            unsigned t = __shfl_up_sync(0xffffffffu, inc, d);
                            if (lane >= d) inc += t;
                        }
                        const unsigned kk = s_k;
                        const unsigned ball = __ballot_sync(0xffffffffu, inc >= kk);
                        const int owner = __ffs(ball) - 1;               // exists: total >= kk
                        if (lane == owner) {
                            unsigned cum = inc - tot;
                            int dsel = 0;
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                if (cum + loc[j] >= kk) { dsel = 255 - 8 * lane - j; break; }
                                cum += loc[j];
                            }
                            s_k = kk - cum;
                            s_prefix = prefix | ((unsigned long long)dsel << shift);
                        }
                    }
                    mask |= 255ull << shift;
                    __syncthreads();
                    // The response bits are settled after pass 4.  Any threshold gives an exact batch (a prefix of the
                    // sorted order), so unless ties on the response leave the batch less than half full, take the
                    // keys whose response is strictly larger and skip the address passes.
                    if (pass == 4 && SEL_CAP - s_k >= SEL_CAP / 2) { early = true; break; }
                }
                lo = early ? s_prefix + (1ull << 32) : s_prefix;
            }
            // ---- gather [lo, hi) and sort descending
            if (tid == 0) s_nb = 0;
            __syncthreads();
            for (unsigned i0 = tid; i0 < n; i0 += 4 * SEL_THREADS) {
                unsigned long long kq[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) kq[u] = (i0 + u * SEL_THREADS < n) ? keys[i0 + u * SEL_THREADS] : 0ull;
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (kq[u] >= lo && kq[u] != 0ull) {
                        const unsigned pos = atomicAdd(&s_nb, 1u);
                        if (pos < SEL_CAP) batch[pos] = kq[u];
                    }
            }
            __syncthreads();
        }
        const unsigned nb = min(s_nb, (unsigned)SEL_CAP);
        sort_batch(nb);
        // ---- greedy acceptance, in rounds.  Round: (1) the next SEL_WIN words (32 candidates each) of the sorted
        // batch are tested against the accepted-corner grid, one warp per word (a candidate that fails is dead
        // for good: the accepted set only grows); (2) warp 0 takes the first 32 survivors of that window IN
        // ORDER, resolves the order dependence among them lane by lane (exactly the sequential greedy) and
        // inserts the winners.  Survivors beyond those 32 are re-tested in the next round against the enlarged
        // grid; candidates beyond the window are not touched until the window reaches them.
        {
            constexpr int NMASK = SEL_CAP / 32;
            __shared__ unsigned s_mask[NMASK];
            __shared__ unsigned s_pos;
            if (tid == 0) s_pos = 0;
            // alive mask of the batch: a candidate that once failed the grid test is never looked at again
            for (unsigned wv = tid; wv < NMASK; wv += SEL_THREADS) {
                const unsigned lo_i = wv * 32u;
                s_mask[wv] = (lo_i + 32u <= nb) ? ~0u : (lo_i < nb ? ((1u << (nb - lo_i)) - 1u) : 0u);
            }
            __syncthreads();
            while (true) {
                const unsigned pos0 = s_pos;
                const int nacc_r = s_naccepted;
                if (pos0 >= nb || nacc_r >= sp.max_corners) break;
                // (1) re-test the alive, not yet visited candidates against the grown grid
                const unsigned w0 = pos0 >> 5, wend = min(w0 + (unsigned)SEL_WIN, (unsigned)NMASK);
                for (unsigned wv = w0 + warp; wv < wend; wv += SEL_THREADS / 32) {
                    const unsigned alive = s_mask[wv];
                    if (alive == 0u) continue;                                  // warp-uniform
                    const unsigned idx = wv * 32 + lane;
                    bool pass = ((alive >> lane) & 1u) && idx >= pos0;
                    if (pass && sp.use_min_dist && nacc_r > 0) {
                        const unsigned addr = (unsigned)batch[idx];
                        const int y = (int)(addr / (unsigned)W), x = (int)(addr - (unsigned)y * (unsigned)W);
                        pass = grid_pass(cellpts, sp.gw, sp.gh, sp.cell, sp.min_dist2, x, y);
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, pass);
                    if (lane == 0) s_mask[wv] = m;
                }
                __syncthreads();
                if (warp == 0) {
                    // lane r picks the r-th survivor (in batch order)
                    unsigned c0 = 0, c1 = 0;
                    // words entirely below pos0 are history; the word containing pos0 was re-masked above (idx >= pos0)
                    if (NMASK > 32) {
                        c0 = ((unsigned)lane >= w0 && (unsigned)lane < wend) ? __popc(s_mask[lane]) : 0u;
                        c1 = ((unsigned)lane + 32u >= w0 && (unsigned)lane + 32u < wend) ? __popc(s_mask[lane + 32]) : 0u;
                    } else if (lane < NMASK) c0 = ((unsigned)lane >= w0 && (unsigned)lane < wend) ? __popc(s_mask[lane]) : 0u;
                    unsigned inc0 = c0, inc1 = c1;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const unsigned t0 = __shfl_up_sync(0xffffffffu, inc0, d), t1 = __shfl_up_sync(0xffffffffu, inc1, d);
                        if (lane >= d) { inc0 += t0; inc1 += t1; }
                    }
                    const unsigned tot0 = __shfl_sync(0xffffffffu, inc0, 31);
                    inc1 += tot0;
                    const unsigned total = __shfl_sync(0xffffffffu, inc1, 31);
                    const unsigned r = (unsigned)lane;
                    int widx = -1;
                    unsigned before = 0;
                    for (int wq = (int)w0; wq < (int)wend; ++wq) {     // warp-uniform shuffles; 64 inclusive counts live across lanes
                        const unsigned incw = (wq < 32) ? __shfl_sync(0xffffffffu, inc0, wq & 31) : __shfl_sync(0xffffffffu, inc1, wq & 31);
                        const unsigned cw = (wq < 32) ? __shfl_sync(0xffffffffu, c0, wq & 31) : __shfl_sync(0xffffffffu, c1, wq & 31);
                        if (widx < 0 && r < total && incw > r) { widx = wq; before = incw - cw; }
                        if (incw >= 32u) break;               // uniform: the first 32 survivors are located
                    }
                    const bool valid = widx >= 0;
                    unsigned idx = 0;
                    if (valid) {
                        unsigned m = s_mask[widx];
                        for (unsigned t = 0; t < r - before; ++t) m &= m - 1u;      // drop the lower set bits
                        idx = (unsigned)widx * 32u + (unsigned)(__ffs(m) - 1);
                    }
                    const unsigned long long k = valid ? batch[idx] : 0ull;
                    const unsigned addr = (unsigned)k;
                    const int y = (int)(addr / (unsigned)W), x = (int)(addr - (unsigned)y * (unsigned)W);
                    // conflict matrix: bit j of conf = survivor j (higher priority, j < lane) lies within minDistance
                    unsigned conf = 0;
                    if (sp.use_min_dist) {
                        for (int j = 0; j < 31; ++j) {
                            const int jx = __shfl_sync(0xffffffffu, x, j), jy = __shfl_sync(0xffffffffu, y, j);
                            const int dx = x - jx, dy = y - jy;
                            if (j < lane && (float)(dx * dx + dy * dy) < sp.min_dist2) conf |= 1u << j;
                        }
                    }
                    // sequential greedy over the 32 survivors, done with bit operations (every lane redundantly)
                    const unsigned vmask = __ballot_sync(0xffffffffu, valid);
                    unsigned acc = 0;
                    int room = sp.max_corners - nacc_r;
                    for (int i = 0; i < 32; ++i) {
                        const unsigned ci = __shfl_sync(0xffffffffu, conf, i);
                        if (((vmask >> i) & 1u) && !(ci & acc) && room > 0) { acc |= 1u << i; --room; }
                    }
                    // accepted lanes insert themselves (output order = priority order)
                    if ((acc >> lane) & 1u) {
                        const int slot = nacc_r + __popc(acc & ((1u << lane) - 1u));
                        ax[slot] = (float)x; ay[slot] = (float)y; ar[slot] = __uint_as_float((unsigned)(k >> 32));
                        if (sp.use_min_dist) {
                            const int cidx = (y / sp.cell) * sp.gw + (x / sp.cell);
                            const int q = atomicAdd(&cellcnt[cidx], 1);
                            if (q < 4) reinterpret_cast<unsigned *>(&cellpts[cidx])[q] = (unsigned)x | ((unsigned)y << 16);
                            else atomicExch(det.overflow, 1u);    // geometrically impossible (<= 3 per cell)
                        }
                    }
                    const int nacc = nacc_r + __popc(acc);
                    // next round starts after the last survivor taken this round (all of them were decided)
                    const unsigned taken = min(total, 32u);
                    const unsigned last_idx = __shfl_sync(0xffffffffu, idx, (int)max(taken, 1u) - 1);
                    if (lane == 0) {
                        s_naccepted = nacc;
                        s_pos = (total <= 32u) ? min(wend * 32u, nb) : last_idx + 1;    // window exhausted, or resume after the last one taken
                    }
                }
                __syncthreads();
            }
            if (tid == 0) {
                s_done = (s_naccepted >= sp.max_corners || n_el <= SEL_CAP) ? 1 : 0;
                if (nb) s_hi = batch[nb - 1];
            }
        }
        __syncthreads();
        if (s_done) break;
    }
    __syncthreads();

    const int na = s_naccepted;
    if (tid == 0) gftt_counts[b] = na;
    for (int i = tid; i < na; i += SEL_THREADS) {
        gftt_xy[((size_t)b * sp.cap_k + i) * 2] = ax[i];
        gftt_xy[((size_t)b * sp.cap_k + i) * 2 + 1] = ay[i];
        gftt_resp[(size_t)b * sp.cap_k + i] = ar[i];
    }
}

// ---- extra::PoissonDiskFilter<2> + border reject + append (poisson_disk_filter.h:23-94, opencv_image.cpp:57-71).
// Separate launch: it is the only part of detect that depends on the tracked keypoints, so the GFTT
// selection above can overlap the LK kernel on another stream.
constexpr int PO_THREADS = 512;

__global__ void __launch_bounds__(PO_THREADS)
poisson_append_kernel(SelectParams sp, const float *__restrict__ gftt_xy, const int *__restrict__ gftt_counts,
                      double *__restrict__ kp_xy, int *__restrict__ kp_counts, const char *__restrict__ lk_status,
                      unsigned *__restrict__ truncated) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *pxs = reinterpret_cast<double *>(smem_raw);                                     // [stride] preset x
    double *pys = pxs + sp.stride;                                                          // [stride] preset y
    float *ax = reinterpret_cast<float *>(pys + sp.stride);                                 // [cap_k]
    float *ay = ax + sp.cap_k;
    int *pcx = reinterpret_cast<int *>(ay + sp.cap_k);                                      // [stride] preset cells
    int *pcy = pcx + sp.stride;
    int *ccx_s = pcy + sp.stride;                                                           // [cap_k] candidate cells
    int *ccy_s = ccx_s + sp.cap_k;
    int *ins = ccy_s + sp.cap_k;                                                            // [cap_k] inserted candidates
    unsigned char *pflag = reinterpret_cast<unsigned char *>(ins + sp.cap_k);               // [stride] visible presets
    unsigned char *cflag = pflag + sp.stride;                                               // [cap_k] candidate rejected
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = PO_THREADS / 32;
    const int b = blockIdx.x;
    const int na = min(gftt_counts[b], sp.cap_k);
    double *pts = kp_xy + (size_t)b * sp.stride * 2;
    int ne = min(kp_counts[b], sp.stride);
    const double radius = sp.kp_radius, r2 = radius * radius;
    const double gsz = radius / sqrt(2.0);
    const int span = (int)ceil(sqrt(2.0));
    // ---- load presets and candidates, compute their grid cells once (float64 division: reference semantics)
    for (int i = tid; i < ne; i += PO_THREADS) {
        pxs[i] = pts[2 * i]; pys[i] = pts[2 * i + 1];
        pflag[i] = lk_status ? (lk_status[(size_t)b * sp.stride + i] != 0) : 1;
    }
    if (lk_status) {
        // Frame::track_keypoints appends only the status != 0 points to the next frame, in order (frame.cpp:160-170);
        // those are the existing keypoints detect sees.  Stable in-place compaction by warp 0 (writes never pass reads).
        __shared__ int s_ne;
        __syncthreads();
        if (warp == 0) {
            int w = 0;
            for (int base = 0; base < ne; base += 32) {
                const int i = base + lane;
                const bool keep = i < ne && pflag[i];
                const double x = i < ne ? pxs[i] : 0.0, y = i < ne ? pys[i] : 0.0;
                const unsigned m = __ballot_sync(0xffffffffu, keep);
                const int pos = w + __popc(m & ((1u << lane) - 1u));
                __syncwarp();
                if (keep) { pxs[pos] = x; pys[pos] = y; }
                w += __popc(m);
            }
            if (lane == 0) s_ne = w;
        }
        __syncthreads();
        ne = s_ne;
        for (int i = tid; i < ne; i += PO_THREADS) { pts[2 * i] = pxs[i]; pts[2 * i + 1] = pys[i]; }
    }
    for (int i = tid; i < ne; i += PO_THREADS) {
        pcx[i] = (int)floor(pxs[i] / gsz);
        pcy[i] = (int)floor(pys[i] / gsz);
        pflag[i] = 1;
    }
    for (int c = tid; c < na; c += PO_THREADS) {
        const float x = gftt_xy[((size_t)b * sp.cap_k + c) * 2], y = gftt_xy[((size_t)b * sp.cap_k + c) * 2 + 1];
        ax[c] = x; ay[c] = y;
        ccx_s[c] = (int)floor((double)x / gsz);
        ccy_s[c] = (int)floor((double)y / gsz);
        cflag[c] = 0;
    }
    __syncthreads();
    // ---- preset_point: a later preset in the same cell overwrites (hides) an earlier one
    for (int j = warp; j < ne; j += NW) {
        const int jx = pcx[j], jy = pcy[j];
        for (int i = lane; i < j; i += 32)
            if (pcx[i] == jx && pcy[i] == jy) pflag[i] = 0;
    }
    __syncthreads();
    // ---- candidate x preset pairs: a warp per candidate, lanes over presets
    for (int c = warp; c < na; c += NW) {
        const double cx = (double)ax[c], cy = (double)ay[c];
        const int ccx = ccx_s[c], ccy = ccy_s[c];
        bool hit = false;
        for (int i = lane; i < ne; i += 32) {
            if (!pflag[i]) continue;
            if (!poisson_block_hit(pcx[i] - ccx, pcy[i] - ccy, span)) continue;
            const double dx = cx - pxs[i], dy = cy - pys[i];
            if (dx * dx + dy * dy < r2) hit = true;
        }
        if (__any_sync(0xffffffffu, hit) && lane == 0) cflag[c] = 1;
    }
    __syncthreads();
    if (warp != 0) return;
    int nout = ne;              // write cursor in pts
    if (sp.kp_radius * sp.kp_radius <= (double)sp.min_dist2 && sp.use_min_dist) {
        // GFTT already keeps its corners >= minDistance apart (integer-exact test), so two NEW points can
        // never be closer than the Poisson radius: insertion order no longer matters -> parallel compaction.
        for (int base = 0; base < na; base += 32) {
            const int c = base + lane;
            bool keep = false;
            double cx = 0, cy = 0;
            if (c < na && !cflag[c]) {
                cx = (double)ax[c]; cy = (double)ay[c];
                keep = !(cx < sp.border || cy < sp.border || cx >= sp.W - sp.border || cy >= sp.H - sp.border);
            }
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            const int pos = nout + __popc(m & ((1u << lane) - 1u));
            if (keep && pos < sp.stride) { pts[2 * pos] = cx; pts[2 * pos + 1] = cy; }
            if (nout + __popc(m) > sp.stride && lane == 0) atomicOr(truncated, 2u);   // the reference's vector is unbounded
            nout = min(nout + __popc(m), sp.stride);
        }
    } else {
        // general radius: sequential insertion (new points also block later ones), border reject, append
        int nins = 0;
        for (int c = 0; c < na; ++c) {
            if (cflag[c]) continue;                      // warp-uniform (shared memory flag)
            const double cx = (double)ax[c], cy = (double)ay[c];
            const int ccx = ccx_s[c], ccy = ccy_s[c];
            bool hit = false;
            for (int i = lane; i < nins; i += 32) {
                const int q = ins[i];
                if (!poisson_block_hit(ccx_s[q] - ccx, ccy_s[q] - ccy, span)) continue;
                const double dx = cx - (double)ax[q], dy = cy - (double)ay[q];
                if (dx * dx + dy * dy < r2) hit = true;
            }
            if (__any_sync(0xffffffffu, hit)) continue;
            __syncwarp();
            if (lane == 0) ins[nins] = c;
            ++nins;
            __syncwarp();
            const bool out_of_border = cx < sp.border || cy < sp.border || cx >= sp.W - sp.border || cy >= sp.H - sp.border;
            if (!out_of_border) {
                if (nout < sp.stride) {
                    if (lane == 0) { pts[2 * nout] = cx; pts[2 * nout + 1] = cy; }
                    ++nout;
                } else if (lane == 0) atomicOr(truncated, 2u);
            }
        }
    }
    if (lane == 0) kp_counts[b] = nout;
}

// ---- the same filter with the reference's own data structure: a cell grid holding ONE point index per cell (the last
// one written, poisson_disk_filter.h:23-27,43-45), kept dense in shared memory over the cells a corner inside the image
// can ever look at (test_point's 5x5 block minus its first cell plus one, :69-92).  A preset outside that range shares
// no cell with an in-range one and is farther than the radius from every candidate, so it is simply not entered.
// O(25) cell probes per candidate instead of the pairwise loops above; used whenever the grid fits shared memory.
struct PoissonGrid {
    int gw, gh;              // dense cells per axis (corner cells 0 .. gcw-1 shifted by OX = 2, OY = 2)
};

__global__ void __launch_bounds__(PO_THREADS)
poisson_grid_kernel(SelectParams sp, PoissonGrid pg, const float *__restrict__ gftt_xy, const int *__restrict__ gftt_counts,
                    double *__restrict__ kp_xy, int *__restrict__ kp_counts, const char *__restrict__ lk_status,
                    unsigned *__restrict__ truncated) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *pxs = reinterpret_cast<double *>(smem_raw);                                     // [stride] preset x
    double *pys = pxs + sp.stride;                                                          // [stride] preset y
    float *ax = reinterpret_cast<float *>(pys + sp.stride);                                 // [cap_k] candidates
    float *ay = ax + sp.cap_k;
    int *grid = reinterpret_cast<int *>(ay + sp.cap_k);                                     // [gw*gh] point index or -1
    unsigned char *pflag = reinterpret_cast<unsigned char *>(grid + pg.gw * pg.gh);         // [stride] status flags
    unsigned char *cflag = pflag + sp.stride;                                               // [cap_k] candidate rejected
    __shared__ int s_ne;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x;
    const int na = min(gftt_counts[b], sp.cap_k);
    double *pts = kp_xy + (size_t)b * sp.stride * 2;
    int ne = min(kp_counts[b], sp.stride);
    const double radius = sp.kp_radius, r2 = radius * radius;
    const double gsz = radius / sqrt(2.0);
    constexpr int span = 2, OX = 2, OY = 2;                 // ceil(sqrt(2)) = 2
    for (int i = tid; i < pg.gw * pg.gh; i += PO_THREADS) grid[i] = -1;
    for (int i = tid; i < ne; i += PO_THREADS) {
        pxs[i] = pts[2 * i]; pys[i] = pts[2 * i + 1];
        pflag[i] = lk_status ? (lk_status[(size_t)b * sp.stride + i] != 0) : 1;
    }
    for (int c = tid; c < na; c += PO_THREADS) {
        ax[c] = gftt_xy[((size_t)b * sp.cap_k + c) * 2];
        ay[c] = gftt_xy[((size_t)b * sp.cap_k + c) * 2 + 1];
    }
    __syncthreads();
    if (lk_status) {
        // Frame::track_keypoints appends only the status != 0 points to the next frame, in order (frame.cpp:160-170)
        if (warp == 0) {
            int w = 0;
            for (int base = 0; base < ne; base += 32) {
                const int i = base + lane;
                const bool keep = i < ne && pflag[i];
                const double x = i < ne ? pxs[i] : 0.0, y = i < ne ? pys[i] : 0.0;
                const unsigned m = __ballot_sync(0xffffffffu, keep);
                const int pos = w + __popc(m & ((1u << lane) - 1u));
                __syncwarp();
                if (keep) { pxs[pos] = x; pys[pos] = y; }
                w += __popc(m);
            }
            if (lane == 0) s_ne = w;
        }
        __syncthreads();
        ne = s_ne;
        for (int i = tid; i < ne; i += PO_THREADS) { pts[2 * i] = pxs[i]; pts[2 * i + 1] = pys[i]; }
    }
    // ---- preset_point in index order: the highest index written to a cell is the one that stays
    for (int i = tid; i < ne; i += PO_THREADS) {
        const int cx = (int)floor(pxs[i] / gsz) + OX, cy = (int)floor(pys[i] / gsz) + OY;
        if (cx >= 0 && cx < pg.gw && cy >= 0 && cy < pg.gh) atomicMax(&grid[cy * pg.gw + cx], i);
    }
    __syncthreads();
    // the 25 cells test_point visits, relative to the query cell: (dx, dy) in [-2, 2]^2 without (-2, -2), plus (-2, 3)
    auto cell_of = [&](int q, int &dx, int &dy) {
        if (q < 24) { const int t = q + 1; dx = t % 5 - span; dy = t / 5 - span; }
        else { dx = -span; dy = span + 1; }
    };
    // ---- candidates against the presets: one thread per candidate
    for (int c = tid; c < na; c += PO_THREADS) {
        const double cx = (double)ax[c], cy = (double)ay[c];
        const int ccx = (int)floor(cx / gsz) + OX, ccy = (int)floor(cy / gsz) + OY;
        bool hit = false;
        for (int q = 0; q < 25; ++q) {
            int dx, dy;
            cell_of(q, dx, dy);
            const int i = grid[(ccy + dy) * pg.gw + (ccx + dx)];
            if (i >= 0) {
                const double ex = cx - pxs[i], ey = cy - pys[i];
                if (ex * ex + ey * ey < r2) hit = true;
            }
        }
        cflag[c] = hit ? 1 : 0;
    }
    __syncthreads();
    if (warp != 0) return;
    int nout = ne;              // write cursor in pts
    if (sp.kp_radius * sp.kp_radius <= (double)sp.min_dist2 && sp.use_min_dist) {
        // GFTT keeps its corners >= minDistance apart, so two NEW points never block each other, and a new point can
        // only land in the cell of a preset it is closer than the radius to (cell diagonal = radius), i.e. never next to a
        // visible one: the flags above are final -> parallel compaction.
        for (int base = 0; base < na; base += 32) {
            const int c = base + lane;
            bool keep = false;
            double cx = 0, cy = 0;
            if (c < na && !cflag[c]) {
                cx = (double)ax[c]; cy = (double)ay[c];
                keep = !(cx < sp.border || cy < sp.border || cx >= sp.W - sp.border || cy >= sp.H - sp.border);
            }
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            const int pos = nout + __popc(m & ((1u << lane) - 1u));
            if (keep && pos < sp.stride) { pts[2 * pos] = cx; pts[2 * pos + 1] = cy; }
            if (nout + __popc(m) > sp.stride && lane == 0) atomicOr(truncated, 2u);   // the reference's vector is unbounded
            nout = min(nout + __popc(m), sp.stride);
        }
    } else {
        // general radius: insertion in order; an accepted corner takes over its cell (index stride + c) and blocks later
        // candidates.  Lane q probes cell q of the 25; presets need no second look (see above).
        for (int c = 0; c < na; ++c) {
            if (cflag[c]) continue;                      // warp-uniform (shared memory flag)
            const double cx = (double)ax[c], cy = (double)ay[c];
            const int ccx = (int)floor(cx / gsz) + OX, ccy = (int)floor(cy / gsz) + OY;
            bool hit = false;
            if (lane < 25) {
                int dx, dy;
                cell_of(lane, dx, dy);
                const int i = grid[(ccy + dy) * pg.gw + (ccx + dx)];
                if (i >= sp.stride) {
                    const double ex = cx - (double)ax[i - sp.stride], ey = cy - (double)ay[i - sp.stride];
                    if (ex * ex + ey * ey < r2) hit = true;
                }
            }
            if (__any_sync(0xffffffffu, hit)) continue;
            __syncwarp();
            if (lane == 0) grid[ccy * pg.gw + ccx] = sp.stride + c;
            __syncwarp();
            const bool out_of_border = cx < sp.border || cy < sp.border || cx >= sp.W - sp.border || cy >= sp.H - sp.border;
            if (!out_of_border) {
                if (nout < sp.stride) {
                    if (lane == 0) { pts[2 * nout] = cx; pts[2 * nout + 1] = cy; }
                    ++nout;
                } else if (lane == 0) atomicOr(truncated, 2u);
            }
        }
    }
    if (lane == 0) kp_counts[b] = nout;
}

static void fill_select_params(rdfe_ctx *ctx, const rdfe_detect_params &p, int stride, SelectParams &sp) {
    const LevelGeom &g = ctx->pyr.lv[0];
    sp.W = g.w; sp.H = g.h;
    sp.max_corners = p.max_points;
    sp.quality = p.quality_level;
    sp.use_min_dist = p.min_distance >= 1.0 ? 1 : 0;
    sp.cell = sp.use_min_dist ? (int)lrint(p.min_distance) : 1;
    sp.gw = sp.use_min_dist ? (g.w + sp.cell - 1) / sp.cell : 1;
    sp.gh = sp.use_min_dist ? (g.h + sp.cell - 1) / sp.cell : 1;
    sp.min_dist2 = (float)(p.min_distance * p.min_distance);
    sp.kp_radius = p.keypoint_distance;
    sp.border = p.border;
    sp.stride = stride;
    sp.cap_k = p.max_points;
    sp.harris_k = (float)p.harris_k;
    sp.harris_fma = p.harris_fma;
    sp.prefiltered = 1;
}

// GFTT selection on `stream` (may be the context's auxiliary stream): candidates -> gftt_* arrays
int launch_gftt_select(rdfe_ctx *ctx, const SlotList &slots, const rdfe_detect_params &p, float *d_gftt_xy,
                       float *d_gftt_resp, int *d_gftt_counts) {
    const int n = slots.n;
    SelectParams sp;
    fill_select_params(ctx, p, 1, sp);
    const size_t smem = (size_t)SEL_CAP * 8 + 256 * 4 + (size_t)sp.cap_k * 12 + 16 + (size_t)sp.gw * sp.gh * 20 + 64;
    if (smem > 200 * 1024) {
        set_error("select: shared memory need %zu B exceeds the CTA limit (max_points=%d, grid %dx%d)", smem,
                  p.max_points, sp.gw, sp.gh);
        return RDFE_ERR_UNSUPPORTED;
    }
    // the opt-in is a per-device attribute: remembered per context (a context is bound to one device)
    if (smem > 48 * 1024 && smem > ctx->smem_optin[0]) {
        if (cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            set_error("select: cudaFuncSetAttribute(%zu) failed", smem);
            return RDFE_ERR_CUDA;
        }
        ctx->smem_optin[0] = smem;
    }
    RDFE_LAUNCH(ctx, K_SELECT, (select_kernel<<<n, SEL_THREADS, smem, ctx->ls>>>(ctx->det, sp, ctx->pyr, slots, d_gftt_xy, d_gftt_resp, d_gftt_counts)));
    return 1;
}

int launch_poisson_append(rdfe_ctx *ctx, int n, const rdfe_detect_params &p, const float *d_gftt_xy,
                          const int *d_gftt_counts, double *d_xy, int *d_counts, int stride, const char *d_lk_status) {
    SelectParams sp;
    fill_select_params(ctx, p, stride, sp);
    // dense cell grid (the reference's sparse_grid restricted to the cells a corner inside the image can probe)
    static const int s_impl = [] { const char *e = getenv("RDFE_POISSON_IMPL"); return e ? atoi(e) : 1; }();
    if (s_impl == 1 && p.keypoint_distance > 0.0) {
        const double gsz = p.keypoint_distance / sqrt(2.0);
        PoissonGrid pg;
        pg.gw = (int)floor((double)(sp.W - 1) / gsz) + 1 + 4;       // corner cells + 2 on either side
        pg.gh = (int)floor((double)(sp.H - 1) / gsz) + 1 + 5;       // ... and the extra probe at dy = +3
        const size_t smem_g = (size_t)stride * 17 + (size_t)sp.cap_k * 9 + (size_t)pg.gw * pg.gh * 4 + 64;
        if (smem_g <= 160 * 1024) {
            if (smem_g > 48 * 1024 && smem_g > ctx->smem_optin[3]) {
                if (cudaFuncSetAttribute(poisson_grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_g) != cudaSuccess) {
                    set_error("poisson: cudaFuncSetAttribute(%zu) failed", smem_g);
                    return RDFE_ERR_CUDA;
                }
                ctx->smem_optin[3] = smem_g;
            }
            RDFE_LAUNCH(ctx, K_POISSON, (poisson_grid_kernel<<<n, PO_THREADS, smem_g, ctx->ls>>>(sp, pg, d_gftt_xy, d_gftt_counts, d_xy, d_counts,
                                                                                               d_lk_status, ctx->det.overflow)));
            return 1;
        }
    }
    const size_t smem = (size_t)sp.cap_k * 21 + (size_t)stride * 25 + 64;
    if (smem > 200 * 1024) { set_error("poisson: shared memory need %zu B too large", smem); return RDFE_ERR_UNSUPPORTED; }
    if (smem > 48 * 1024 && smem > ctx->smem_optin[1]) {
        if (cudaFuncSetAttribute(poisson_append_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            set_error("poisson: cudaFuncSetAttribute(%zu) failed", smem);
            return RDFE_ERR_CUDA;
        }
        ctx->smem_optin[1] = smem;
    }
    RDFE_LAUNCH(ctx, K_POISSON, (poisson_append_kernel<<<n, PO_THREADS, smem, ctx->ls>>>(sp, d_gftt_xy, d_gftt_counts, d_xy, d_counts, d_lk_status, ctx->det.overflow)));
    return 1;
}

int launch_select(rdfe_ctx *ctx, const SlotList &slots, const rdfe_detect_params &p, double *d_xy, int *d_counts,
                  int stride, float *d_gftt_xy, float *d_gftt_resp, int *d_gftt_counts) {
    float *gxy = d_gftt_xy ? d_gftt_xy : ctx->d_gftt_xy;
    float *gre = d_gftt_resp ? d_gftt_resp : ctx->d_gftt_resp;
    int *gcn = d_gftt_counts ? d_gftt_counts : ctx->d_gftt_counts;
    int rc = launch_gftt_select(ctx, slots, p, gxy, gre, gcn);
    if (rc < 0) return rc;
    if (!d_xy) return 1;
    rc = launch_poisson_append(ctx, slots.n, p, gxy, gcn, d_xy, d_counts, stride, nullptr);
    return rc < 0 ? rc : 2;
}

}  // namespace rdfe
