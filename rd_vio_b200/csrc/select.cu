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

namespace rdfe {

constexpr int SEL_THREADS = 1024;
constexpr int SEL_CAP = 2048;

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
};

__device__ __forceinline__ bool poisson_block_hit(int dcx, int dcy, int span) {
    // cells visited by PoissonDiskFilter::test_point relative to the query cell
    if (dcx < -span || dcx > span) return false;
    if (dcy >= -span && dcy <= span) return !(dcx == -span && dcy == -span);
    return dcx == -span && dcy == span + 1;
}

__global__ void __launch_bounds__(SEL_THREADS)
select_kernel(DetectScratch det, SelectParams sp, SlotList slots, double *__restrict__ kp_xy, int *__restrict__ kp_counts,
              float *__restrict__ gftt_xy, float *__restrict__ gftt_resp, int *__restrict__ gftt_counts) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long *batch = reinterpret_cast<unsigned long long *>(smem_raw);          // [SEL_CAP]
    unsigned *hist = reinterpret_cast<unsigned *>(batch + SEL_CAP);                         // [256]
    float *ax = reinterpret_cast<float *>(hist + 256);                                      // [cap_k]
    float *ay = ax + sp.cap_k;
    float *ar = ay + sp.cap_k;
    int *anext = reinterpret_cast<int *>(ar + sp.cap_k);
    int *head = anext + sp.cap_k;                                                           // [gw*gh]
    int *pcx = head + sp.gw * sp.gh;                                                        // [stride]
    int *pcy = pcx + sp.stride;
    unsigned char *pflag = reinterpret_cast<unsigned char *>(pcy + sp.stride);              // [stride] visible presets
    unsigned char *cflag = pflag + sp.stride;                                               // [cap_k] candidate rejected

    __shared__ unsigned s_count, s_nb;
    __shared__ unsigned long long s_prefix, s_hi;
    __shared__ unsigned s_k;
    __shared__ int s_naccepted, s_done;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x;
    const unsigned n = min(det.cand_count[b], det.cand_cap);
    const unsigned long long *keys = det.cand + (size_t)b * det.cand_cap;
    const float maxv = __uint_as_float(det.frame_max[b]);
    const float thr = (float)((double)maxv * sp.quality);
    const unsigned thr_bits = __float_as_uint(thr);       // thr >= 0 here (maxv >= 0)
    const int W = sp.W;

    for (int i = tid; i < sp.gw * sp.gh; i += SEL_THREADS) head[i] = -1;
    if (tid == 0) { s_naccepted = 0; s_done = 0; s_hi = ~0ull; }
    __syncthreads();

    // address bits that can be non-zero: skip radix passes above them
    const int addr_bits = 32 - __clz(max(sp.W * sp.H - 1, 1));
    const int low_passes = (addr_bits + 7) / 8;

    while (true) {
        const unsigned long long hi = s_hi;
        // ---- count eligible keys (R > thr, key < hi)
        if (tid == 0) s_count = 0;
        __syncthreads();
        unsigned cnt = 0;
        for (unsigned i = tid; i < n; i += SEL_THREADS) {
            const unsigned long long k = keys[i];
            cnt += ((unsigned)(k >> 32) > thr_bits && k < hi) ? 1u : 0u;
        }
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        if (lane == 0 && cnt) atomicAdd(&s_count, cnt);
        __syncthreads();
        const unsigned n_el = s_count;
        if (n_el == 0) break;

        // ---- exact radix select of the SEL_CAP-th largest eligible key
        unsigned long long lo = 0;
        if (n_el > SEL_CAP) {
            if (tid == 0) { s_prefix = 0; s_k = SEL_CAP; }
            unsigned long long mask = 0;
            for (int pass = 7; pass >= 0; --pass) {
                if (pass < 4 && pass >= low_passes) continue;      // digits known to be zero
                const int shift = pass * 8;
                if (tid < 256) hist[tid] = 0;
                __syncthreads();
                const unsigned long long prefix = s_prefix;
                for (unsigned i = tid; i < n; i += SEL_THREADS) {
                    const unsigned long long k = keys[i];
                    if ((unsigned)(k >> 32) > thr_bits && k < hi && (k & mask) == prefix)
                        atomicAdd(&hist[(unsigned)(k >> shift) & 255u], 1u);
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
            }
            lo = s_prefix;
        }
        // ---- gather [lo, hi) and sort descending
        if (tid == 0) s_nb = 0;
        __syncthreads();
        for (unsigned i = tid; i < n; i += SEL_THREADS) {
            const unsigned long long k = keys[i];
            if ((unsigned)(k >> 32) > thr_bits && k < hi && k >= lo) {
                const unsigned pos = atomicAdd(&s_nb, 1u);
                if (pos < SEL_CAP) batch[pos] = k;
            }
        }
        __syncthreads();
        const unsigned nb = min(s_nb, (unsigned)SEL_CAP);
        unsigned N = 32;
        while (N < nb) N <<= 1;
        for (unsigned i = nb + tid; i < N; i += SEL_THREADS) batch[i] = 0ull;
        __syncthreads();
        for (unsigned k2 = 2; k2 <= N; k2 <<= 1) {
            for (unsigned j = k2 >> 1; j > 0; j >>= 1) {
                for (unsigned t = tid; t < (N >> 1); t += SEL_THREADS) {
                    const unsigned i = 2 * j * (t / j) + (t % j);
                    const unsigned p = i + j;
                    const unsigned long long a = batch[i], c = batch[p];
                    const bool desc = ((i & k2) == 0);
                    if (desc ? (a < c) : (a > c)) { batch[i] = c; batch[p] = a; }
                }
                __syncthreads();
            }
        }
        // ---- greedy acceptance (warp 0)
        if (warp == 0) {
            int nacc = s_naccepted;
            for (unsigned base = 0; base < nb && nacc < sp.max_corners; base += 32) {
                const unsigned idx = base + lane;
                const bool valid = idx < nb;
                const unsigned long long k = valid ? batch[idx] : 0ull;
                const unsigned addr = (unsigned)k;
                const int y = (int)(addr / (unsigned)W), x = (int)(addr - (unsigned)y * (unsigned)W);
                const float fx = (float)x, fy = (float)y;
                bool pass = valid;
                int xc = 0, yc = 0;
                if (sp.use_min_dist) {
                    xc = x / sp.cell; yc = y / sp.cell;
                    if (pass) {
                        const int x1 = max(xc - 1, 0), y1 = max(yc - 1, 0);
                        const int x2 = min(xc + 1, sp.gw - 1), y2 = min(yc + 1, sp.gh - 1);
                        for (int yy = y1; yy <= y2 && pass; ++yy)
                            for (int xx = x1; xx <= x2 && pass; ++xx)
                                for (int j = head[yy * sp.gw + xx]; j >= 0; j = anext[j]) {
                                    const float dx = fx - ax[j], dy = fy - ay[j];
                                    if (dx * dx + dy * dy < sp.min_dist2) { pass = false; break; }
                                }
                    }
                }
                // resolve the sequential dependence inside the chunk
                unsigned pending = __ballot_sync(0xffffffffu, pass);
                while (pending && nacc < sp.max_corners) {
                    const int j = __ffs(pending) - 1;
                    const float jx = __shfl_sync(0xffffffffu, fx, j), jy = __shfl_sync(0xffffffffu, fy, j);
                    if (lane == j) {
                        ax[nacc] = fx; ay[nacc] = fy; ar[nacc] = __uint_as_float((unsigned)(k >> 32));
                        if (sp.use_min_dist) { anext[nacc] = head[yc * sp.gw + xc]; head[yc * sp.gw + xc] = nacc; }
                        pass = false;
                    }
                    ++nacc;
                    if (sp.use_min_dist && pass && lane > j) {
                        const float dx = fx - jx, dy = fy - jy;
                        if (dx * dx + dy * dy < sp.min_dist2) pass = false;
                    }
                    __syncwarp();
                    pending = __ballot_sync(0xffffffffu, pass) & ~((2u << j) - 1u);
                }
                __syncwarp();
            }
            if (lane == 0) {
                s_naccepted = nacc;
                s_done = (nacc >= sp.max_corners || n_el <= SEL_CAP) ? 1 : 0;
                if (nb) s_hi = batch[nb - 1];
            }
        }
        __syncthreads();
        if (s_done) break;
    }
    __syncthreads();

    const int na = s_naccepted;
    if (gftt_counts && tid == 0) gftt_counts[b] = na;
    if (gftt_xy)
        for (int i = tid; i < na; i += SEL_THREADS) {
            gftt_xy[((size_t)b * sp.cap_k + i) * 2] = ax[i];
            gftt_xy[((size_t)b * sp.cap_k + i) * 2 + 1] = ay[i];
            if (gftt_resp) gftt_resp[(size_t)b * sp.cap_k + i] = ar[i];
        }
    if (!kp_xy) return;

    // ---- Poisson-disk filter against the existing keypoints (float64, reference semantics)
    double *pts = kp_xy + (size_t)b * sp.stride * 2;
    const int ne = min(kp_counts[b], sp.stride);
    const double radius = sp.kp_radius, r2 = radius * radius;
    const double gsz = radius / sqrt(2.0);
    const int span = (int)ceil(sqrt(2.0));
    for (int i = tid; i < ne; i += SEL_THREADS) {
        pcx[i] = (int)floor(pts[2 * i] / gsz);
        pcy[i] = (int)floor(pts[2 * i + 1] / gsz);
    }
    __syncthreads();
    for (int i = tid; i < ne; i += SEL_THREADS) {
        bool vis = true;                                 // preset_point: later preset in the same cell overwrites
        for (int j = i + 1; j < ne; ++j)
            if (pcx[j] == pcx[i] && pcy[j] == pcy[i]) { vis = false; break; }
        pflag[i] = vis ? 1 : 0;
    }
    for (int i = tid; i < na; i += SEL_THREADS) cflag[i] = 0;
    __syncthreads();
    // candidate x preset pairs
    for (long long t = tid; t < (long long)na * ne; t += SEL_THREADS) {
        const int c = (int)(t / ne), i = (int)(t - (long long)c * ne);
        if (!pflag[i]) continue;
        const double cx = (double)ax[c], cy = (double)ay[c];
        const int ccx = (int)floor(cx / gsz), ccy = (int)floor(cy / gsz);
        if (!poisson_block_hit(pcx[i] - ccx, pcy[i] - ccy, span)) continue;
        const double dx = cx - pts[2 * i], dy = cy - pts[2 * i + 1];
        if (dx * dx + dy * dy < r2) cflag[c] = 1;
    }
    __syncthreads();
    // sequential insertion of the survivors (new points also block later ones), border reject, append
    if (warp == 0) {
        int nout = ne;          // write cursor in pts
        int nins = 0;           // inserted candidates are kept compacted in ax/ay[0..nins) -- safe: nins <= c
        for (int c = 0; c < na; ++c) {
            if (cflag[c]) continue;                      // warp-uniform (shared memory flag)
            const double cx = (double)ax[c], cy = (double)ay[c];
            const int ccx = (int)floor(cx / gsz), ccy = (int)floor(cy / gsz);
            bool hit = false;
            for (int i = lane; i < nins; i += 32) {
                const double qx = (double)ax[i], qy = (double)ay[i];
                const int qcx = (int)floor(qx / gsz), qcy = (int)floor(qy / gsz);
                if (!poisson_block_hit(qcx - ccx, qcy - ccy, span)) continue;
                const double dx = cx - qx, dy = cy - qy;
                if (dx * dx + dy * dy < r2) hit = true;
            }
            if (__any_sync(0xffffffffu, hit)) continue;
            __syncwarp();
            if (lane == 0) { ax[nins] = (float)cx; ay[nins] = (float)cy; }
            ++nins;
            __syncwarp();
            const bool out_of_border = cx < sp.border || cy < sp.border || cx >= sp.W - sp.border || cy >= sp.H - sp.border;
            if (!out_of_border && nout < sp.stride) {
                if (lane == 0) { pts[2 * nout] = cx; pts[2 * nout + 1] = cy; }
                ++nout;
            }
        }
        if (lane == 0) kp_counts[b] = nout;
    }
}

int launch_select(rdfe_ctx *ctx, const SlotList &slots, const rdfe_detect_params &p, double *d_xy, int *d_counts,
                  int stride, float *d_gftt_xy, float *d_gftt_resp, int *d_gftt_counts) {
    const LevelGeom &g = ctx->pyr.lv[0];
    SelectParams sp;
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
    const size_t smem = (size_t)SEL_CAP * 8 + 256 * 4 + (size_t)sp.cap_k * 16 + (size_t)sp.gw * sp.gh * 4 +
                        (size_t)stride * 8 + (size_t)stride + (size_t)sp.cap_k + 64;
    if (smem > 200 * 1024) {
        set_error("select: shared memory need %zu B exceeds the CTA limit (max_points=%d, grid %dx%d)", smem,
                  p.max_points, sp.gw, sp.gh);
        return RDFE_ERR_UNSUPPORTED;
    }
    static size_t s_attr = 0;
    if (smem > 48 * 1024 && smem > s_attr) {
        if (cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            set_error("select: cudaFuncSetAttribute(%zu) failed", smem);
            return RDFE_ERR_CUDA;
        }
        s_attr = smem;
    }
    RDFE_LAUNCH(ctx, K_SELECT, (select_kernel<<<slots.n, SEL_THREADS, smem, ctx->stream>>>(ctx->det, sp, slots, d_xy, d_counts, d_gftt_xy,
                                                                                            d_gftt_resp, d_gftt_counts)));
    return 1;
}

}  // namespace rdfe
