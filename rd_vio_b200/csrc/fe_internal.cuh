// fe_internal.cuh -- context layout and kernel launch prototypes shared by the
// translation units of librdvio_fe.so (sm_100a only; compiled with -fmad=false
// so every float op is individually rounded like the reference's C++ path).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "../../include/rdvio_fe.h"

namespace rdfe {
// steps the host-buffer pipeline (rdfe_frontend_step_submit/_wait) may hold in flight: with three, the upload of
// step s+2 is already queued while step s computes, so the copy engine never waits for the host to come back
constexpr int kPipeDepth = 3;


constexpr int kHaloX = 32;          // left halo columns of an image plane (>= win, keeps interior 32-B aligned)
constexpr int kMaxTiles = 16;       // CLAHE tiles per axis
constexpr int kSMs = 148;           // B200

// Geometry of one pyramid level.  Image planes carry a materialised
// REFLECT_101 halo of `win` px (what buildOpticalFlowPyramid keeps around
// every level); derivative planes carry none -- their zero halo comes from
// TMA out-of-bounds fill.
struct LevelGeom {
    int w, h;
    int ipitch;          // bytes per padded image row
    int ph;              // padded rows = h + 2*win
    size_t islot;        // bytes per slot in the image slab
    int dpitch;          // bytes per derivative row (w * 4, 64-B aligned)
    size_t dslot;        // bytes per slot in the derivative slab
};

// cv::borderInterpolate(p, len, BORDER_REFLECT_101)
__host__ __device__ __forceinline__ int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
    return p;
}

struct SlotList {
    int n;
    int v[RDFE_MAX_BATCH];
};

struct Pyramid {
    int nlevels, win;
    LevelGeom lv[RDFE_MAX_LEVELS];
    uint8_t *img[RDFE_MAX_LEVELS];      // slab: [num_slots][ph][ipitch]
    int16_t *der[RDFE_MAX_LEVELS];      // slab: [num_slots][h][dpitch/2]
    __host__ __device__ inline uint8_t *image_origin(int l, int slot) const {
        // pointer to pixel (0,0) of the level interior
        return img[l] + (size_t)slot * lv[l].islot + (size_t)win * lv[l].ipitch + kHaloX;
    }
    __host__ __device__ inline int16_t *deriv_origin(int l, int slot) const {
        return (int16_t *)((uint8_t *)der[l] + (size_t)slot * lv[l].dslot);
    }
};

#ifdef __CUDACC__
// Stores pixels x..x+3 (x % 4 == 0; only those < W) of row y of an image plane whose interior origin is
// `org`, AND their REFLECT_101 mirror images inside the win-px halo: the producer of a level materialises
// the halo that cv::buildOpticalFlowPyramid keeps around it (copyMakeBorder, BORDER_REFLECT_101).
__device__ __forceinline__ void store4_row(uint8_t *r, int W, int win, int x, int nvalid, unsigned v4, bool near_x) {
    if (nvalid == 4) *reinterpret_cast<unsigned *>(r + x) = v4;
    else for (int k = 0; k < nvalid; ++k) r[x + k] = (uint8_t)(v4 >> (8 * k));
    if (near_x) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int xx = x + k;
            const uint8_t b = (uint8_t)(v4 >> (8 * k));
            if (k < nvalid && xx >= 1 && xx <= win) r[-xx] = b;
            if (k < nvalid && xx >= W - 1 - win && xx <= W - 2) r[2 * (W - 1) - xx] = b;
        }
    }
}
// border groups only (about 15 % of level 0)
__device__ __forceinline__ void store4_border(uint8_t *org, int pitch, int W, int H, int win, int x, int y, unsigned v4) {
    const int nvalid = min(4, W - x);
    const bool near_x = (x <= win) || (x + 3 >= W - 1 - win);
    store4_row(org + (ptrdiff_t)y * pitch, W, win, x, nvalid, v4, near_x);
    if (y >= 1 && y <= win) store4_row(org - (ptrdiff_t)y * pitch, W, win, x, nvalid, v4, near_x);
    if (y >= H - 1 - win && y <= H - 2) store4_row(org + (ptrdiff_t)(2 * (H - 1) - y) * pitch, W, win, x, nvalid, v4, near_x);
}
// Warp-cooperative store of one image row in 4-pixel groups, REFLECT_101 side halos included.  Lane owns group g
// (columns 4g .. 4g+3, value v) or is inactive; ALL 32 lanes must call, with consecutive g across the lanes.
// `left` / `right` (warp-uniform): this batch of 32 groups holds the nh = ceil(win / 4) groups next to the left /
// right image edge and one more neighbour (checked by the caller, together with W % 4 == 0).  The halo is written as
// aligned words assembled from the lane's own group and its neighbour's (one shuffle + one PRMT + one store per
// side) instead of byte by byte; up to 3 columns beyond `win` are filled too (still the correct mirror values).
__device__ __forceinline__ void store4_row_coop(uint8_t *row, int W, int nh, int g, bool active, unsigned v,
                                                bool left, bool right) {
    if (active) *reinterpret_cast<unsigned *>(row + 4 * g) = v;
    if (left) {
        const unsigned nx = __shfl_down_sync(0xffffffffu, v, 1);
        // columns -4(g+1) .. -4g-1 mirror columns 4g+4, 4g+3, 4g+2, 4g+1
        if (active && g < nh) *reinterpret_cast<unsigned *>(row - 4 * (g + 1)) = __byte_perm(v, nx, 0x1234u);
    }
    if (right) {
        const unsigned pv = __shfl_up_sync(0xffffffffu, v, 1);
        const int m = (W >> 2) - 1 - g;      // columns W+4m .. W+4m+3 mirror columns W-2-4m, W-3-4m, W-4-4m, W-5-4m
        if (active && m < nh) *reinterpret_cast<unsigned *>(row + W + 4 * m) = __byte_perm(v, pv, 0x7012u);
    }
}
// ... and its mirror images in the top / bottom halo (y warp-uniform)
__device__ __forceinline__ void store4_rows_coop(uint8_t *org, int pitch, int W, int H, int win, int nh, int g, int y,
                                                 bool active, unsigned v, bool left, bool right) {
    store4_row_coop(org + (ptrdiff_t)y * pitch, W, nh, g, active, v, left, right);
    if (y >= 1 && y <= win) store4_row_coop(org - (ptrdiff_t)y * pitch, W, nh, g, active, v, left, right);
    if (y >= H - 1 - win && y <= H - 2) store4_row_coop(org + (ptrdiff_t)(2 * (H - 1) - y) * pitch, W, nh, g, active, v, left, right);
}
// whether store4_row(s)_coop applies to rows of W pixels cut into batches of 32 groups
__host__ __device__ __forceinline__ bool coop_halo_ok(int W, int win) {
    const int G = W >> 2, nh = (win + 3) >> 2;
    return (W & 3) == 0 && G >= nh + 1 && (G - 1 - nh) / 32 == (G - 1) / 32;
}
__device__ __forceinline__ void store4_with_halo(uint8_t *org, int pitch, int W, int H, int win, int x, int y,
                                                 unsigned v4) {
    // interior: no mirror image of these pixels lies in the halo (small levels may have no interior at all)
    if (x > win && x + 3 < W - 1 - win && y > win && y < H - 1 - win)
        *reinterpret_cast<unsigned *>(org + (ptrdiff_t)y * pitch + x) = v4;
    else
        store4_border(org, pitch, W, H, win, x, y, v4);
}
#endif

// CLAHE launch parameters (host-precomputed so float boundary cases are
// evaluated once, with the same float expressions as the reference).
struct ClaheParams {
    int W, H;
    int tiles_x, tiles_y;
    int tw, th;              // tile size (of the padded image, if padding applies)
    int padded;              // W % tiles_x || H % tiles_y  (copyMakeBorder REFLECT_101 quirk)
    int clip;                // integer clip limit, 0 = no clipping
    float lut_scale;         // 255.f / (tw*th)
    float inv_tw, inv_th;
    // cell boundaries: pixels x in [xb[c], xb[c+1]) interpolate between tile columns c-1 and c
    int xb[kMaxTiles + 2];
    int yb[kMaxTiles + 2];
};

// Strip height of the warp-rolling kernels: `max_rows` at full batches; smaller (never below `min_rows`) when the
// batch is too small to give every SM a few warps -- a strip is walked row by row, so its height is the kernel's
// latency in the single-stream plugin path.  warps_per_row_strip = column tiles * images.
inline int adaptive_strip_rows(int height, int warps_per_row_strip, int min_rows, int max_rows) {
    const int want_warps = 148 * 8;
    int rows = (int)(((long long)height * warps_per_row_strip + want_warps - 1) / want_warps);
    rows = (rows + 3) & ~3;
    return rows < min_rows ? min_rows : rows > max_rows ? max_rows : rows;
}

enum KernelId {
    K_CLAHE_HIST = 0, K_CLAHE_APPLY, K_PYRDOWN, K_SCHARR, K_HARRIS, K_SELECT, K_LK, K_POISSON, K_UNDISTORT, K_HARRIS_RESOLVE, K_PREDICT, K_COUNT
};
constexpr int kProfMax = 4096;      // timed launches between two rdfe_profile_collect calls

struct DetectScratch {
    unsigned long long *cand;     // [RDFE_MAX_BATCH][cand_cap] keys: (float bits << 32) | pixel address
    unsigned long long *cand2;    // same size: ping-pong buffer for the per-round compaction in select_kernel
    unsigned *cand_count;         // [RDFE_MAX_BATCH]
    unsigned *frame_max;          // [RDFE_MAX_BATCH] max response bits (responses <= 0 never win)
    unsigned *overflow;           // [1]
    unsigned *flag_count;         // [RDFE_MAX_BATCH] pixels flagged by the Harris prefilter (statistics; the flag bytes live in cand2)
    unsigned cand_cap;
};

}  // namespace rdfe

struct rdfe_ctx {
    rdfe_config cfg;
    rdfe::Pyramid pyr;
    uint8_t *raw;                 // upload staging slab [num_slots][H][raw_pitch]
    size_t raw_pitch, raw_slot;
    size_t gray_pitch, gray_slot;  // geometry of the ingest output plane (und_plane)
    CUtensorMap tm_img[RDFE_MAX_LEVELS];
    CUtensorMap tm_der[RDFE_MAX_LEVELS];
    CUtensorMap tm_imgT[RDFE_MAX_LEVELS];   // image, box = LK template patch (JW x (win+1))
    // LK template cache (rdfe_set_template_cache): templates the backward pass built, keyed by (generation, x, y)
    bool tc_on;
    uint4 *tc_data;               // [num_slots][max_points][record]
    float4 *tc_A, *tc_hdr;        // [num_slots][max_points][levels], [num_slots][max_points]
    unsigned long long *tc_stats; // device [2]: lookups, hits
    unsigned *slot_gen;           // [num_slots] host: bumped whenever level 0 of the slot is rewritten
    unsigned gen_counter;
    cudaStream_t stream;
    bool own_stream;
    cudaStream_t ls;              // stream the kernel launchers currently enqueue on (stream or aux_stream)
    bool overlap;                 // rdfe_frontend_step*: run Harris + selection on aux_stream beside pyramid + LK
    cudaStream_t aux_stream;      // detection branch of rdfe_frontend_step*
    cudaStream_t aux_stream2;     // ... of odd steps, so that Harris(t+1) may overlap select(t)
    cudaEvent_t ev_fork, ev_join, ev_join2;
    // pipelined rdfe_frontend_step*: every kernel class on its own stream (and priority), linked by events
    cudaStream_t sel_stream[2], trk_stream, post_stream;
    cudaEvent_t ev_harris_done[2], ev_lk_done[2], ev_entry;
    long long *slot_new_step;     // [num_slots] step index at which the slot was last a step's NEW slot
    // rdfe_detect_prefetch: Harris + GFTT selection started early on aux_stream2 (candidate scratch det2)
    bool pf_valid, pf_recorded;
    int pf_n;
    int pf_slots[RDFE_MAX_BATCH];
    rdfe_detect_params pf_params;
    cudaEvent_t pf_done;
    float *pf_gftt_xy, *pf_gftt_resp;
    int *pf_gftt_counts;
    int max_batch;                // min(num_slots, RDFE_MAX_BATCH): images per batched call on this context
    bool compact_tracked;         // fused step: only status != 0 points are carried into detect (frame.cpp:160-170); default on
    size_t smem_optin[4];         // per-context (= per-device) dynamic shared memory already granted: select, poisson, clahe, spare
    bool host_sync;               // host-pointer preprocess waits for completion (default) or returns after the upload
    unsigned *h_overflow;         // pinned copy of det.overflow for the host-pointer detect
    rdfe::DetectScratch det2;     // candidate buffers of odd steps (cand, cand2, count, max; overflow flag shared)
    // optional undistortion in front of preprocess (rdfe_set_undistort): fixed-point remap tables + output staging
    int in_channels;              // 1 gray (default), 3 BGR, 4 BGRA: cvtColor of Odometry::addFrame (rdvio.hpp:42-49)
    bool und_on;
    uint32_t *und_map_xy;         // [H][W] (sx | sy << 16), int16 each
    uint16_t *und_map_f;          // [H][W] fy*32 + fx
    uint8_t *und_plane;           // [num_slots][H][raw_pitch] undistorted frames
    // cross-step pipelining (rdfe_set_pipelining): preprocess of step s+1 on pre_stream beside track/detect of step s
    bool pipeline_steps;
    cudaStream_t pre_stream;
    cudaStream_t pre_stream2;     // Scharr of level 0 beside the pyrDown chain (pipelined step)
    cudaEvent_t ev_sc0_done;
    cudaEvent_t ev_apply_done, ev_pre_done, ev_step_done[2];
    int64_t step_index;
    uint8_t *last_step_slots;     // [num_slots] 1 = touched by the previous step
    uint8_t *lut2;                // second CLAHE LUT buffer (steps alternate)
    const uint8_t **d_srcptrs2;
    float *d_gftt_xy2, *d_gftt_resp2;
    int *d_gftt_counts2;
    cudaEvent_t images_ready;     // optional: event the source images depend on (set by the host pipeline)
    bool images_ready_valid;
    // pipelined host-buffer step (rdfe_frontend_step_submit / _wait): two stages
    cudaStream_t copy_stream;     // H2D of the next step's frames overlaps the current step's kernels
    cudaEvent_t ev_clahe_done;    // raw upload staging may be overwritten after this
    cudaEvent_t ev_copy_fence;    // tail of the context stream, for uploads into slots last preprocessed by the separate calls
    cudaEvent_t ev_clahe_ring[4]; // CLAHE of fused step s done: ring entry s & 3
    long long *slot_clahe_step;   // [num_slots] fused step whose CLAHE last read the slot's raw staging (-1: not by a fused step)
    cudaEvent_t ev_upload[rdfe::kPipeDepth], ev_done[rdfe::kPipeDepth];
    double *pl_curr[rdfe::kPipeDepth], *pl_next[rdfe::kPipeDepth];      // device [RDFE_MAX_BATCH][max_points][2]
    int *pl_counts[rdfe::kPipeDepth], *pl_kcounts[rdfe::kPipeDepth];
    char *pl_status[rdfe::kPipeDepth];
    unsigned *pl_ovf[rdfe::kPipeDepth];
    uint8_t *pl_host[rdfe::kPipeDepth];   // pinned result staging: next_xy | kcounts | status | overflow
    int pl_n[rdfe::kPipeDepth], pl_stride[rdfe::kPipeDepth], pl_busy[rdfe::kPipeDepth];
    int64_t pl_ticket;
    cudaEvent_t ev_t0, ev_t1;
    // scratch
    uint8_t *lut;                 // [RDFE_MAX_BATCH][tiles][256]
    rdfe::DetectScratch det;
    // host-pointer track / detect: ONE pinned block and its device twin, so that a call is one H2D, the kernels, one D2H
    // (copies straight from / to the caller's pageable vectors cost a staged driver copy each)
    uint8_t *h_io, *d_io;
    size_t io_bytes;
    // host<->device staging for the host-pointer API
    double *d_xy_a, *d_xy_b;      // [RDFE_MAX_BATCH][max_points][2]
    int *d_counts;                // [RDFE_MAX_BATCH]
    char *d_status;               // [RDFE_MAX_BATCH][max_points]
    float *d_gftt_xy, *d_gftt_resp;
    int *d_gftt_counts;
    const uint8_t **d_srcptrs;    // [RDFE_MAX_BATCH] device copy of source pointers
    uint8_t *slot_used;
    int64_t launches;
    int last_clahe_tiles;
    // optional per-kernel CUDA-event timing (rdfe_profile_*)
    bool prof_on;
    bool prof_timeline;           // events recorded without serialising the streams (rdfe_profile_enable(ctx, 2))
    int prof_used;
    cudaEvent_t *prof_ev;         // [2 * kProfMax]
    int prof_kid[rdfe::kProfMax];
    double prof_ms[rdfe::K_COUNT];
    int64_t prof_n[rdfe::K_COUNT];
};

namespace rdfe {

void set_error(const char *fmt, ...);
void build_undistort_map(int W, int H, const float *K, const float *D, std::vector<uint32_t> &mxy, std::vector<uint16_t> &mf);
#define RDFE_CUDA_OK(expr)                                                            \
    do {                                                                              \
        cudaError_t e__ = (expr);                                                     \
        if (e__ != cudaSuccess) {                                                     \
            rdfe::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                            __FILE__, __LINE__);                                      \
            return RDFE_ERR_CUDA;                                                     \
        }                                                                             \
    } while (0)

// Bracket one kernel launch with CUDA events when profiling is on.
inline int prof_begin(rdfe_ctx *ctx, int kid) {
    if (!ctx->prof_on || ctx->prof_used >= kProfMax) return -1;
    const int i = ctx->prof_used++;
    ctx->prof_kid[i] = kid;
    cudaEventRecord(ctx->prof_ev[2 * i], ctx->ls);
    return i;
}
inline void prof_end(rdfe_ctx *ctx, int i) {
    if (i >= 0) cudaEventRecord(ctx->prof_ev[2 * i + 1], ctx->ls);
}
#define RDFE_LAUNCH(ctx, KID, ...)                     \
    do {                                               \
        const int pi__ = rdfe::prof_begin(ctx, KID);   \
        __VA_ARGS__;                                   \
        rdfe::prof_end(ctx, pi__);                     \
    } while (0)

// kernel launchers (each returns the number of kernels it launched, or <0 on error)
int launch_clahe(rdfe_ctx *ctx, const SlotList &slots, const uint8_t *const *d_src, size_t src_pitch,
                 int src_vec4, const ClaheParams &cp);
int launch_pyramid(rdfe_ctx *ctx, const SlotList &slots);
int launch_pyrdowns(rdfe_ctx *ctx, const SlotList &slots);
int launch_scharr_levels(rdfe_ctx *ctx, const SlotList &slots, int lo, int hi);
int launch_harris_candidates(rdfe_ctx *ctx, const SlotList &slots, const rdfe_detect_params &p,
                             float *d_response /* optional [n][H][W] */);
int launch_select(rdfe_ctx *ctx, const SlotList &slots, const rdfe_detect_params &p, double *d_xy, int *d_counts,
                  int stride, float *d_gftt_xy, float *d_gftt_resp, int *d_gftt_counts);
int launch_gftt_select(rdfe_ctx *ctx, const SlotList &slots, const rdfe_detect_params &p, float *d_gftt_xy,
                       float *d_gftt_resp, int *d_gftt_counts);
int launch_poisson_append(rdfe_ctx *ctx, int n, const rdfe_detect_params &p, const float *d_gftt_xy,
                          const int *d_gftt_counts, double *d_xy, int *d_counts, int stride,
                          const char *d_lk_status /* optional: keep only status != 0 presets (frame.cpp:160-170) */);
int launch_undistort(rdfe_ctx *ctx, int n, const uint8_t *const *d_src, size_t src_pitch, int src_vec4, uint8_t *const *d_dst,
                     size_t dst_pitch);
size_t lk_cache_record_bytes(int win, int nlevels);
int launch_predict_rotation(rdfe_ctx *ctx, int n, const double *d_H, const double *d_curr_xy, const int *d_counts, int stride,
                            double *d_pred_xy);
int launch_lk(rdfe_ctx *ctx, const SlotList &curr, const SlotList &next, const rdfe_track_params &p,
              const double *d_curr_xy, double *d_next_xy, const int *d_counts, int stride, char *d_status);

}  // namespace rdfe
