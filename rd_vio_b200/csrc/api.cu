// api.cu -- the C ABI of librdvio_fe.so (see include/rdvio_fe.h for the contract and the
// reference member functions each entry point replaces).
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>

#include "fe_internal.cuh"
#include "harris_exact.cuh"

namespace rdfe {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

static int make_tensor_maps(rdfe_ctx *ctx) {
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return RDFE_ERR_CUDA; }
    const Pyramid &pyr = ctx->pyr;
    const int win = pyr.win;
    // box sizes must match LKCfg<win> in lk.cu (inner extents padded so the box can start 16-B aligned)
    const cuuint32_t jw = win == 21 ? 48 : 64, jh = win == 21 ? 32 : 48;
    const cuuint32_t dw = win == 21 ? 28 : 36, dh = win == 21 ? 22 : 32;
    for (int l = 0; l < pyr.nlevels; ++l) {
        const LevelGeom &g = pyr.lv[l];
        {
            cuuint64_t dims[3] = {(cuuint64_t)g.ipitch, (cuuint64_t)g.ph, (cuuint64_t)ctx->cfg.num_slots};
            cuuint64_t strides[2] = {(cuuint64_t)g.ipitch, (cuuint64_t)g.islot};
            cuuint32_t box[3] = {jw, jh, 1};
            cuuint32_t es[3] = {1, 1, 1};
            CUresult r = enc(&ctx->tm_img[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, pyr.img[l], dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(image level %d) failed: %d", l, (int)r); return RDFE_ERR_CUDA; }
            // the template patch reads only win+1 rows: its own box keeps the over-fetch of the J box out of DRAM
            cuuint32_t boxT[3] = {jw, (cuuint32_t)win + 1, 1};
            r = enc(&ctx->tm_imgT[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, pyr.img[l], dims, strides, boxT, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(template level %d) failed: %d", l, (int)r); return RDFE_ERR_CUDA; }
        }
        {
            // derivative plane: one uint32 element = (dx, dy) int16 pair; dims are the true
            // image size so that out-of-bounds reads return the reference's zero halo
            cuuint64_t dims[3] = {(cuuint64_t)g.w, (cuuint64_t)g.h, (cuuint64_t)ctx->cfg.num_slots};
            cuuint64_t strides[2] = {(cuuint64_t)g.dpitch, (cuuint64_t)g.dslot};
            cuuint32_t box[3] = {dw, dh, 1};
            cuuint32_t es[3] = {1, 1, 1};
            CUresult r = enc(&ctx->tm_der[l], CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, pyr.der[l], dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(deriv level %d) failed: %d", l, (int)r); return RDFE_ERR_CUDA; }
        }
    }
    return RDFE_OK;
}

// CLAHE parameters exactly as cv::CLAHE_Impl::apply derives them
static int make_clahe_params(const rdfe_ctx *ctx, double clip_limit, int tiles_x, int tiles_y, ClaheParams *cp) {
    const int W = ctx->cfg.width, H = ctx->cfg.height;
    if (tiles_x < 1 || tiles_y < 1 || tiles_x > kMaxTiles || tiles_y > kMaxTiles) {
        set_error("CLAHE tile grid %dx%d out of range [1,%d]", tiles_x, tiles_y, kMaxTiles);
        return RDFE_ERR_INVALID;
    }
    int PW = W, PH = H;
    cp->padded = 0;
    if (W % tiles_x != 0 || H % tiles_y != 0) {
        PW = W + (tiles_x - (W % tiles_x));
        PH = H + (tiles_y - (H % tiles_y));
        cp->padded = 1;
    }
    cp->W = W; cp->H = H;
    cp->tiles_x = tiles_x; cp->tiles_y = tiles_y;
    cp->tw = PW / tiles_x; cp->th = PH / tiles_y;
    const int area = cp->tw * cp->th;
    cp->clip = 0;
    if (clip_limit > 0.0) {
        cp->clip = (int)(clip_limit * area / 256);
        if (cp->clip < 1) cp->clip = 1;
    }
    cp->lut_scale = (float)255 / (float)area;
    cp->inv_tw = 1.0f / (float)cp->tw;
    cp->inv_th = 1.0f / (float)cp->th;
    // interpolation cells: cell c <=> floor(x * inv_tw - 0.5f) + 1 == c  (monotone in x)
    auto build = [](int n, int tiles, float inv, int *bnd) {
        int c = 0;
        bnd[0] = 0;
        for (int x = 0; x < n; ++x) {
            volatile float t = (float)x * inv;       // volatile: keep the two roundings separate
            const float tf = t - 0.5f;
            int cell = (int)floorf(tf) + 1;
            if (cell > tiles) cell = tiles;            // cannot happen for in-range x; defensive
            while (c < cell) bnd[++c] = x;
        }
        while (c < tiles + 1) bnd[++c] = n;
    };
    build(W, tiles_x, cp->inv_tw, cp->xb);
    build(H, tiles_y, cp->inv_th, cp->yb);
    return RDFE_OK;
}

static int check_slots(const rdfe_ctx *ctx, const int *slots, int n, SlotList *out, const char *what) {
    if (!ctx || !slots || n < 1 || n > ctx->max_batch) {
        set_error("%s: batch size %d out of range [1,%d] (min(num_slots, RDFE_MAX_BATCH))", what, n, ctx ? ctx->max_batch : RDFE_MAX_BATCH);
        return RDFE_ERR_INVALID;
    }
    out->n = n;
    for (int i = 0; i < n; ++i) {
        if (slots[i] < 0 || slots[i] >= ctx->cfg.num_slots || !ctx->slot_used[slots[i]]) {
            set_error("%s: slot %d (index %d) is not an acquired slot", what, slots[i], i);
            return RDFE_ERR_INVALID;
        }
        out->v[i] = slots[i];
    }
    return RDFE_OK;
}

static inline const SlotList &sl_copy(const SlotList &s) { return s; }

static int check_launch(rdfe_ctx *ctx, int launched, const char *what);

// Copies the source pointer array to the device and, if undistortion is on, remaps every frame into the
// per-slot undistorted plane first.  Returns the device pointer array / pitch the CLAHE kernels must read.
static int stage_sources(rdfe_ctx *ctx, const int *slots, int n, const uint8_t *const *dev_images, size_t pitch,
                         const uint8_t *const **d_src_out, size_t *pitch_out, int *vec4_out) {
    cudaStream_t st = ctx->ls;
    int vec4 = (pitch % 4 == 0) ? 1 : 0;
    for (int i = 0; i < n; ++i) {
        if (!dev_images[i]) { set_error("image %d is null", i); return RDFE_ERR_INVALID; }
        if ((uintptr_t)dev_images[i] % 4) vec4 = 0;
    }
    if (!ctx->und_on && ctx->in_channels == 1) {
        RDFE_CUDA_OK(cudaMemcpyAsync(ctx->d_srcptrs, dev_images, n * sizeof(uint8_t *), cudaMemcpyHostToDevice, st));
        *d_src_out = ctx->d_srcptrs; *pitch_out = pitch; *vec4_out = vec4;
        return RDFE_OK;
    }
    const uint8_t *ptrs[2 * RDFE_MAX_BATCH];
    for (int i = 0; i < n; ++i) {
        ptrs[i] = dev_images[i];
        ptrs[RDFE_MAX_BATCH + i] = ctx->und_plane + (size_t)slots[i] * ctx->gray_slot;
    }
    RDFE_CUDA_OK(cudaMemcpyAsync(ctx->d_srcptrs, ptrs, sizeof ptrs, cudaMemcpyHostToDevice, st));
    int rc = check_launch(ctx, launch_undistort(ctx, n, ctx->d_srcptrs, pitch, vec4, (uint8_t *const *)(ctx->d_srcptrs + RDFE_MAX_BATCH), ctx->gray_pitch), "ingest");
    if (rc) return rc;
    *d_src_out = ctx->d_srcptrs + RDFE_MAX_BATCH; *pitch_out = ctx->gray_pitch; *vec4_out = 1;
    return RDFE_OK;
}

static int sync_all_streams(rdfe_ctx *ctx) {
    for (cudaStream_t st : {ctx->stream, ctx->aux_stream, ctx->aux_stream2, ctx->pre_stream, ctx->pre_stream2, ctx->sel_stream[0], ctx->sel_stream[1],
                            ctx->trk_stream, ctx->post_stream})
        RDFE_CUDA_OK(cudaStreamSynchronize(st));
    return RDFE_OK;
}

// Host frames -> the upload staging of their slots.  When the host pointers and the slots are both equally spaced
// (a batch cut out of per-stream ring buffers, consecutive slots) all frames go in ONE 2-D copy (one "row" per
// frame) instead of n calls: the per-call overhead of 64 small copies costs more than the PCIe time.
static int upload_frames(rdfe_ctx *ctx, cudaStream_t st, const int *slots, int n, const uint8_t *const *images, size_t pitch,
                         std::vector<const uint8_t *> &dptr) {
    const size_t row_bytes = (size_t)ctx->cfg.width * ctx->in_channels;
    const size_t frame_bytes = pitch * (size_t)ctx->cfg.height;
    for (int i = 0; i < n; ++i) dptr[i] = ctx->raw + (size_t)slots[i] * ctx->raw_slot;
    bool strided = n > 1 && pitch == ctx->raw_pitch;
    const ptrdiff_t hstep = n > 1 ? images[1] - images[0] : 0;
    const int sstep = n > 1 ? slots[1] - slots[0] : 0;
    if (strided && (hstep < (ptrdiff_t)frame_bytes || sstep < 1)) strided = false;
    for (int i = 2; strided && i < n; ++i)
        if (images[i] - images[i - 1] != hstep || slots[i] - slots[i - 1] != sstep) strided = false;
    if (strided) {
        RDFE_CUDA_OK(cudaMemcpy2DAsync(const_cast<uint8_t *>(dptr[0]), (size_t)sstep * ctx->raw_slot, images[0], (size_t)hstep,
                                       frame_bytes, (size_t)n, cudaMemcpyHostToDevice, st));
        return RDFE_OK;
    }
    for (int i = 0; i < n; ++i) {
        uint8_t *d = const_cast<uint8_t *>(dptr[i]);
        if (pitch == ctx->raw_pitch)
            RDFE_CUDA_OK(cudaMemcpyAsync(d, images[i], frame_bytes, cudaMemcpyHostToDevice, st));
        else
            RDFE_CUDA_OK(cudaMemcpy2DAsync(d, ctx->raw_pitch, images[i], pitch, row_bytes, (size_t)ctx->cfg.height, cudaMemcpyHostToDevice, st));
    }
    return RDFE_OK;
}

static int check_launch(rdfe_ctx *ctx, int launched, const char *what) {
    if (launched < 0) return launched;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: kernel launch failed: %s", what, cudaGetErrorString(e));
        return RDFE_ERR_CUDA;
    }
    ctx->launches += launched;
    return RDFE_OK;
}

}  // namespace rdfe

using namespace rdfe;

extern "C" {

const char *rdfe_last_error(void) { return g_err; }
int rdfe_abi_version(void) { return RDFE_ABI_VERSION; }

void rdfe_default_detect_params(rdfe_detect_params *p) {
    p->max_points = 150;
    p->quality_level = 1.0e-3;
    p->min_distance = 20.0;
    p->harris_k = 0.04;
    p->keypoint_distance = 20.0;
    p->border = 20;
    p->harris_fma = 0;
}

void rdfe_default_track_params(rdfe_track_params *p) {
    p->max_count = 30;
    p->epsilon = 0.01;
    p->min_eig_threshold = 1e-4;
    p->border = 20;
    p->max_round_trip = 0.5;
    p->has_prediction = 1;
}

int rdfe_create(const rdfe_config *cfg, rdfe_ctx **out) {
    if (!cfg || !out) { set_error("rdfe_create: null argument"); return RDFE_ERR_INVALID; }
    *out = nullptr;
    if (cfg->width < 8 || cfg->height < 8 || cfg->width > 16384 || cfg->height > 16384) {
        set_error("rdfe_create: image size %dx%d unsupported", cfg->width, cfg->height);
        return RDFE_ERR_INVALID;
    }
    if (cfg->win != 21 && cfg->win != 31) {
        set_error("rdfe_create: LK window %d unsupported (21 or 31)", cfg->win);
        return RDFE_ERR_UNSUPPORTED;
    }
    if (cfg->max_level < 0 || cfg->max_level >= RDFE_MAX_LEVELS || cfg->num_slots < 1 || cfg->max_points < 1 ||
        cfg->max_points > 8192) {
        set_error("rdfe_create: max_level=%d num_slots=%d max_points=%d out of range", cfg->max_level, cfg->num_slots,
                  cfg->max_points);
        return RDFE_ERR_INVALID;
    }
    int ndev = 0;
    RDFE_CUDA_OK(cudaGetDeviceCount(&ndev));
    if (cfg->device < 0 || cfg->device >= ndev) {
        set_error("rdfe_create: device %d not present (%d CUDA devices)", cfg->device, ndev);
        return RDFE_ERR_INVALID;
    }
    RDFE_CUDA_OK(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    RDFE_CUDA_OK(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10) {
        set_error("rdfe_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", cfg->device,
                  prop.major, prop.minor);
        return RDFE_ERR_UNSUPPORTED;
    }

    rdfe_ctx *ctx = new rdfe_ctx();
    memset(ctx, 0, sizeof *ctx);
    ctx->cfg = *cfg;
    Pyramid &pyr = ctx->pyr;
    pyr.win = cfg->win;
    // levels: buildOpticalFlowPyramid stops when the next level would be <= win
    int w = cfg->width, h = cfg->height;
    for (int l = 0; l <= cfg->max_level; ++l) {
        LevelGeom &g = pyr.lv[l];
        g.w = w; g.h = h;
        g.ipitch = (int)align_up((size_t)w + 2 * kHaloX, 64);
        g.ph = h + 2 * cfg->win;
        g.islot = align_up((size_t)g.ipitch * g.ph, 256);
        g.dpitch = (int)align_up((size_t)w * 4, 64);
        g.dslot = align_up((size_t)g.dpitch * h, 256);
        pyr.nlevels = l + 1;
        const int nw = (w + 1) / 2, nh = (h + 1) / 2;
        if (nw <= cfg->win || nh <= cfg->win) break;
        w = nw; h = nh;
    }
    auto fail = [&](int rc) { rdfe_destroy(ctx); return rc; };
#define CK(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { set_error("%s failed: %s", #expr, cudaGetErrorString(e__)); return fail(RDFE_ERR_CUDA); } } while (0)
    for (int l = 0; l < pyr.nlevels; ++l) {
        CK(cudaMalloc(&pyr.img[l], pyr.lv[l].islot * cfg->num_slots));
        CK(cudaMemset(pyr.img[l], 0, pyr.lv[l].islot * cfg->num_slots));
        CK(cudaMalloc(&pyr.der[l], pyr.lv[l].dslot * cfg->num_slots));
        CK(cudaMemset(pyr.der[l], 0, pyr.lv[l].dslot * cfg->num_slots));
    }
    ctx->in_channels = 1;
    ctx->gray_pitch = align_up((size_t)cfg->width, 4);
    ctx->gray_slot = align_up(ctx->gray_pitch * cfg->height, 256);
    ctx->raw_pitch = align_up((size_t)cfg->width, 4);   // tight: a contiguous host image uploads as ONE 1-D copy
    ctx->raw_slot = align_up(ctx->raw_pitch * cfg->height, 256);
    CK(cudaMalloc(&ctx->raw, ctx->raw_slot * cfg->num_slots));
    CK(cudaMalloc(&ctx->lut, (size_t)RDFE_MAX_BATCH * kMaxTiles * kMaxTiles * 256));
    // a batch never holds more images than there are slots: the large per-image scratch is sized by that, not by
    // RDFE_MAX_BATCH (a single-stream plugin context with 4 slots needs 1/32 of the candidate memory)
    ctx->max_batch = cfg->num_slots < RDFE_MAX_BATCH ? cfg->num_slots : RDFE_MAX_BATCH;
    const size_t mb = (size_t)ctx->max_batch;
    ctx->det.cand_cap = (unsigned)((size_t)cfg->width * cfg->height / 2);
    CK(cudaMalloc(&ctx->det.cand, mb * ctx->det.cand_cap * sizeof(unsigned long long)));
    CK(cudaMalloc(&ctx->det.cand2, mb * ctx->det.cand_cap * sizeof(unsigned long long)));
    CK(cudaMalloc(&ctx->det.cand_count, RDFE_MAX_BATCH * sizeof(unsigned)));
    CK(cudaMalloc(&ctx->det.frame_max, RDFE_MAX_BATCH * sizeof(unsigned)));
    CK(cudaMalloc(&ctx->det.flag_count, RDFE_MAX_BATCH * sizeof(unsigned)));
    CK(cudaMalloc(&ctx->det.overflow, sizeof(unsigned)));
    CK(cudaMemset(ctx->det.overflow, 0, sizeof(unsigned)));
    ctx->det2 = ctx->det;
    CK(cudaMalloc(&ctx->det2.cand, mb * ctx->det.cand_cap * sizeof(unsigned long long)));
    CK(cudaMalloc(&ctx->det2.cand2, mb * ctx->det.cand_cap * sizeof(unsigned long long)));
    CK(cudaMalloc(&ctx->det2.cand_count, RDFE_MAX_BATCH * sizeof(unsigned)));
    CK(cudaMalloc(&ctx->det2.frame_max, RDFE_MAX_BATCH * sizeof(unsigned)));
    CK(cudaMalloc(&ctx->det2.flag_count, RDFE_MAX_BATCH * sizeof(unsigned)));
    const size_t npts = (size_t)RDFE_MAX_BATCH * cfg->max_points;
    ctx->io_bytes = mb * cfg->max_points * 33 + (size_t)RDFE_MAX_BATCH * 16 + 256;
    CK(cudaMallocHost(&ctx->h_io, ctx->io_bytes));
    CK(cudaMalloc(&ctx->d_io, ctx->io_bytes));
    CK(cudaMalloc(&ctx->d_xy_a, npts * 2 * sizeof(double)));
    CK(cudaMalloc(&ctx->d_xy_b, npts * 2 * sizeof(double)));
    CK(cudaMalloc(&ctx->d_counts, RDFE_MAX_BATCH * sizeof(int)));
    CK(cudaMalloc(&ctx->d_status, npts));
    CK(cudaMalloc(&ctx->d_gftt_xy, npts * 2 * sizeof(float)));
    CK(cudaMalloc(&ctx->d_gftt_resp, npts * sizeof(float)));
    CK(cudaMalloc(&ctx->d_gftt_counts, RDFE_MAX_BATCH * sizeof(int)));
    CK(cudaMalloc(&ctx->d_srcptrs, 2 * RDFE_MAX_BATCH * sizeof(uint8_t *)));
    if (cfg->stream) { ctx->stream = (cudaStream_t)cfg->stream; ctx->own_stream = false; }
    else { CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)); ctx->own_stream = true; }
    ctx->ls = ctx->stream;
    ctx->overlap = true;
    // Stream priorities (0 = default/lowest, negative = higher) of the pipelined step, in the order
    // preprocess, harris, select, lk, poisson.  Kernels of equal priority are dispatched grid after grid (a
    // large grid keeps later kernels of other streams waiting until its last CTA is placed), so the short
    // latency-critical kernels and the upstream stages get the higher levels.  RDFE_PRIO="a,b,c,d,e" overrides.
    // round 2: the Poisson append (13 us, last link of the chain LK(s) -> append(s) -> LK(s+1)) above everything else:
    // +1.2 % frames/s (180.9 k -> 183.1 k, scripts/prio_sweep.sh, twice)
    int prio[5] = {-1, 0, 0, 0, -2};
    if (const char *e = getenv("RDFE_PRIO")) sscanf(e, "%d,%d,%d,%d,%d", &prio[0], &prio[1], &prio[2], &prio[3], &prio[4]);
    const int pre_prio = prio[0];
    CK(cudaStreamCreateWithPriority(&ctx->aux_stream, cudaStreamNonBlocking, prio[1]));
    CK(cudaStreamCreateWithPriority(&ctx->aux_stream2, cudaStreamNonBlocking, prio[1]));
    CK(cudaStreamCreateWithPriority(&ctx->sel_stream[0], cudaStreamNonBlocking, prio[2]));
    CK(cudaStreamCreateWithPriority(&ctx->sel_stream[1], cudaStreamNonBlocking, prio[2]));
    CK(cudaStreamCreateWithPriority(&ctx->trk_stream, cudaStreamNonBlocking, prio[3]));
    CK(cudaStreamCreateWithPriority(&ctx->post_stream, cudaStreamNonBlocking, prio[4]));
    for (int p = 0; p < 2; ++p) {
        CK(cudaEventCreateWithFlags(&ctx->ev_harris_done[p], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ctx->ev_lk_done[p], cudaEventDisableTiming));
    }
    CK(cudaEventCreateWithFlags(&ctx->ev_entry, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->ev_join2, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
    CK(cudaStreamCreateWithPriority(&ctx->pre_stream, cudaStreamNonBlocking, pre_prio));
    CK(cudaStreamCreateWithPriority(&ctx->pre_stream2, cudaStreamNonBlocking, pre_prio));
    CK(cudaEventCreateWithFlags(&ctx->ev_sc0_done, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->ev_apply_done, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->ev_pre_done, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->ev_step_done[0], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->ev_step_done[1], cudaEventDisableTiming));
    CK(cudaMalloc(&ctx->lut2, (size_t)RDFE_MAX_BATCH * kMaxTiles * kMaxTiles * 256));
    CK(cudaMalloc(&ctx->d_srcptrs2, 2 * RDFE_MAX_BATCH * sizeof(uint8_t *)));
    CK(cudaMalloc(&ctx->d_gftt_xy2, npts * 2 * sizeof(float)));
    CK(cudaMalloc(&ctx->d_gftt_resp2, npts * sizeof(float)));
    CK(cudaMalloc(&ctx->d_gftt_counts2, RDFE_MAX_BATCH * sizeof(int)));
    ctx->last_step_slots = (uint8_t *)calloc((size_t)cfg->num_slots, 1);
    CK(cudaEventCreateWithFlags(&ctx->pf_done, cudaEventDisableTiming));
    CK(cudaMalloc(&ctx->pf_gftt_xy, npts * 2 * sizeof(float)));
    CK(cudaMalloc(&ctx->pf_gftt_resp, npts * sizeof(float)));
    CK(cudaMalloc(&ctx->pf_gftt_counts, RDFE_MAX_BATCH * sizeof(int)));
    CK(cudaMallocHost(&ctx->h_overflow, sizeof(unsigned)));
    *ctx->h_overflow = 0u;
    ctx->host_sync = true;
    ctx->compact_tracked = true;
    ctx->slot_new_step = (long long *)malloc((size_t)cfg->num_slots * sizeof(long long));
    for (int i = 0; i < cfg->num_slots; ++i) ctx->slot_new_step[i] = -16;
    CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&ctx->ev_clahe_done, cudaEventDisableTiming));
    for (int i = 0; i < 4; ++i) CK(cudaEventCreateWithFlags(&ctx->ev_clahe_ring[i], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->ev_copy_fence, cudaEventDisableTiming));
    ctx->slot_clahe_step = (long long *)malloc((size_t)cfg->num_slots * sizeof(long long));
    for (int i = 0; i < cfg->num_slots; ++i) ctx->slot_clahe_step[i] = -1;
    for (int p = 0; p < kPipeDepth; ++p) {
        CK(cudaEventCreateWithFlags(&ctx->ev_upload[p], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ctx->ev_done[p], cudaEventDisableTiming));
        CK(cudaMalloc(&ctx->pl_curr[p], npts * 2 * sizeof(double)));
        CK(cudaMalloc(&ctx->pl_next[p], npts * 2 * sizeof(double)));
        CK(cudaMalloc(&ctx->pl_counts[p], RDFE_MAX_BATCH * sizeof(int)));
        CK(cudaMalloc(&ctx->pl_kcounts[p], RDFE_MAX_BATCH * sizeof(int)));
        CK(cudaMalloc(&ctx->pl_status[p], npts));
        CK(cudaMalloc(&ctx->pl_ovf[p], sizeof(unsigned)));
        CK(cudaMallocHost(&ctx->pl_host[p], npts * 17 + RDFE_MAX_BATCH * sizeof(int) + 64));
    }
    CK(cudaEventCreate(&ctx->ev_t0));
    CK(cudaEventCreate(&ctx->ev_t1));
    ctx->prof_ev = (cudaEvent_t *)calloc(2 * kProfMax, sizeof(cudaEvent_t));
    for (int i = 0; i < 2 * kProfMax; ++i) CK(cudaEventCreate(&ctx->prof_ev[i]));
#undef CK
    ctx->slot_used = (uint8_t *)calloc((size_t)cfg->num_slots, 1);
    ctx->slot_gen = (unsigned *)calloc((size_t)cfg->num_slots, sizeof(unsigned));
    ctx->gen_counter = 1;
    int rc = make_tensor_maps(ctx);
    if (rc != RDFE_OK) return fail(rc);
    *out = ctx;
    return RDFE_OK;
}

void rdfe_destroy(rdfe_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->cfg.device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (int l = 0; l < RDFE_MAX_LEVELS; ++l) { cudaFree(ctx->pyr.img[l]); cudaFree(ctx->pyr.der[l]); }
    cudaFree(ctx->raw); cudaFree(ctx->lut); cudaFree(ctx->und_map_xy); cudaFree(ctx->und_map_f); cudaFree(ctx->und_plane);
    cudaFree(ctx->det.cand); cudaFree(ctx->det.cand2); cudaFree(ctx->det.cand_count); cudaFree(ctx->det.frame_max); cudaFree(ctx->det.flag_count); cudaFree(ctx->det.overflow);
    if (ctx->det2.cand != ctx->det.cand) { cudaFree(ctx->det2.cand); cudaFree(ctx->det2.cand2); cudaFree(ctx->det2.cand_count); cudaFree(ctx->det2.frame_max); cudaFree(ctx->det2.flag_count); }
    cudaFree(ctx->d_xy_a); cudaFree(ctx->d_xy_b); cudaFree(ctx->d_counts); cudaFree(ctx->d_status);
    cudaFree(ctx->d_gftt_xy); cudaFree(ctx->d_gftt_resp); cudaFree(ctx->d_gftt_counts); cudaFree(ctx->d_srcptrs);
    if (ctx->prof_ev) {
        for (int i = 0; i < 2 * kProfMax; ++i) if (ctx->prof_ev[i]) cudaEventDestroy(ctx->prof_ev[i]);
        free(ctx->prof_ev);
    }
    if (ctx->ev_t0) cudaEventDestroy(ctx->ev_t0);
    if (ctx->ev_t1) cudaEventDestroy(ctx->ev_t1);
    if (ctx->aux_stream) { cudaStreamSynchronize(ctx->aux_stream); cudaStreamDestroy(ctx->aux_stream); }
    if (ctx->aux_stream2) { cudaStreamSynchronize(ctx->aux_stream2); cudaStreamDestroy(ctx->aux_stream2); }
    for (cudaStream_t st : {ctx->sel_stream[0], ctx->sel_stream[1], ctx->trk_stream, ctx->post_stream})
        if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
    for (cudaEvent_t ev : {ctx->ev_harris_done[0], ctx->ev_harris_done[1], ctx->ev_lk_done[0], ctx->ev_lk_done[1], ctx->ev_entry})
        if (ev) cudaEventDestroy(ev);
    if (ctx->ev_join2) cudaEventDestroy(ctx->ev_join2);
    if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
    if (ctx->pre_stream) { cudaStreamSynchronize(ctx->pre_stream); cudaStreamDestroy(ctx->pre_stream); }
    if (ctx->pre_stream2) { cudaStreamSynchronize(ctx->pre_stream2); cudaStreamDestroy(ctx->pre_stream2); }
    if (ctx->ev_sc0_done) cudaEventDestroy(ctx->ev_sc0_done);
    if (ctx->ev_apply_done) cudaEventDestroy(ctx->ev_apply_done);
    if (ctx->ev_pre_done) cudaEventDestroy(ctx->ev_pre_done);
    for (int p = 0; p < 2; ++p) if (ctx->ev_step_done[p]) cudaEventDestroy(ctx->ev_step_done[p]);
    cudaFree(ctx->lut2); cudaFree(ctx->d_srcptrs2); cudaFree(ctx->d_gftt_xy2); cudaFree(ctx->d_gftt_resp2); cudaFree(ctx->d_gftt_counts2);
    free(ctx->last_step_slots);
    if (ctx->ev_clahe_done) cudaEventDestroy(ctx->ev_clahe_done);
    for (int i = 0; i < 4; ++i) if (ctx->ev_clahe_ring[i]) cudaEventDestroy(ctx->ev_clahe_ring[i]);
    if (ctx->ev_copy_fence) cudaEventDestroy(ctx->ev_copy_fence);
    free(ctx->slot_clahe_step);
    for (int p = 0; p < kPipeDepth; ++p) {
        if (ctx->ev_upload[p]) cudaEventDestroy(ctx->ev_upload[p]);
        if (ctx->ev_done[p]) cudaEventDestroy(ctx->ev_done[p]);
        cudaFree(ctx->pl_curr[p]); cudaFree(ctx->pl_next[p]); cudaFree(ctx->pl_counts[p]); cudaFree(ctx->pl_kcounts[p]);
        cudaFree(ctx->pl_status[p]); cudaFree(ctx->pl_ovf[p]);
        if (ctx->pl_host[p]) cudaFreeHost(ctx->pl_host[p]);
    }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->pf_done) cudaEventDestroy(ctx->pf_done);
    cudaFree(ctx->pf_gftt_xy); cudaFree(ctx->pf_gftt_resp); cudaFree(ctx->pf_gftt_counts);
    if (ctx->h_overflow) cudaFreeHost(ctx->h_overflow);
    if (ctx->h_io) cudaFreeHost(ctx->h_io);
    cudaFree(ctx->d_io);
    free(ctx->slot_used);
    free(ctx->slot_gen);
    cudaFree(ctx->tc_data); cudaFree(ctx->tc_A); cudaFree(ctx->tc_hdr); cudaFree(ctx->tc_stats);
    free(ctx->slot_new_step);
    delete ctx;
}

int rdfe_num_levels(const rdfe_ctx *ctx) { return ctx ? ctx->pyr.nlevels : RDFE_ERR_INVALID; }

// det.overflow bits: 1 = corner-candidate buffer full (harris), 2 = keypoint list cut at `stride` (poisson append)
static int report_overflow(rdfe_ctx *ctx, unsigned bits) {
    if (bits & 1u) set_error("corner-candidate buffer overflow (more than %u local maxima in one image)", ctx->det.cand_cap);
    else set_error("keypoint list truncated: existing + new keypoints exceed `stride` (the reference's vector is unbounded; pass a larger stride)");
    return RDFE_ERR_OVERFLOW;
}

int rdfe_level_size(const rdfe_ctx *ctx, int level, int *width, int *height) {
    if (!ctx || level < 0 || level >= ctx->pyr.nlevels) { set_error("rdfe_level_size: bad level %d", level); return RDFE_ERR_INVALID; }
    if (width) *width = ctx->pyr.lv[level].w;
    if (height) *height = ctx->pyr.lv[level].h;
    return RDFE_OK;
}

int rdfe_sync(rdfe_ctx *ctx) {
    if (!ctx) return RDFE_ERR_INVALID;
    { const int rc_all = sync_all_streams(ctx); if (rc_all) return rc_all; }
    unsigned ovf = 0;
    RDFE_CUDA_OK(cudaMemcpy(&ovf, ctx->det.overflow, sizeof ovf, cudaMemcpyDeviceToHost));
    if (ovf) {
        cudaMemset(ctx->det.overflow, 0, sizeof(unsigned));
        return report_overflow(ctx, ovf);
    }
    return RDFE_OK;
}

void *rdfe_stream(rdfe_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
int64_t rdfe_kernel_launches(const rdfe_ctx *ctx) { return ctx ? ctx->launches : 0; }

int rdfe_slot_acquire(rdfe_ctx *ctx, int *slot) {
    if (!ctx || !slot) return RDFE_ERR_INVALID;
    for (int i = 0; i < ctx->cfg.num_slots; ++i)
        if (!ctx->slot_used[i]) { ctx->slot_used[i] = 1; *slot = i; return RDFE_OK; }
    set_error("rdfe_slot_acquire: all %d slots in use", ctx->cfg.num_slots);
    return RDFE_ERR_NOSLOT;
}

int rdfe_slot_release(rdfe_ctx *ctx, int slot) {
    if (!ctx || slot < 0 || slot >= ctx->cfg.num_slots || !ctx->slot_used[slot]) {
        set_error("rdfe_slot_release: slot %d not acquired", slot);
        return RDFE_ERR_INVALID;
    }
    ctx->slot_used[slot] = 0;
    for (int i = 0; ctx->pf_valid && i < ctx->pf_n; ++i)
        if (ctx->pf_slots[i] == slot) ctx->pf_valid = false;
    return RDFE_OK;
}

// ------------------------------------------------------------- preprocess
int rdfe_preprocess_batch_dev(rdfe_ctx *ctx, const int *slots, int n, const uint8_t *const *dev_images, size_t pitch,
                              double clip_limit, int tiles_x, int tiles_y) {
    SlotList sl;
    int rc = check_slots(ctx, slots, n, &sl, "rdfe_preprocess_batch_dev");
    if (rc) return rc;
    if (!dev_images || pitch < (size_t)ctx->cfg.width * ctx->in_channels) { set_error("rdfe_preprocess_batch_dev: bad image pointers/pitch"); return RDFE_ERR_INVALID; }
    ClaheParams cp;
    rc = make_clahe_params(ctx, clip_limit, tiles_x, tiles_y, &cp);
    if (rc) return rc;
    RDFE_CUDA_OK(cudaSetDevice(ctx->cfg.device));
    if (ctx->pf_recorded) RDFE_CUDA_OK(cudaStreamWaitEvent(ctx->stream, ctx->pf_done, 0));   // a prefetch may still read these slots
    for (int i = 0; ctx->pf_valid && i < ctx->pf_n; ++i)
        for (int j = 0; j < n; ++j)
            if (ctx->pf_slots[i] == slots[j]) ctx->pf_valid = false;
    const uint8_t *const *d_src = nullptr;
    size_t spitch = 0;
    int vec4 = 0;
    rc = stage_sources(ctx, slots, n, dev_images, pitch, &d_src, &spitch, &vec4);
    if (rc) return rc;
    ctx->last_clahe_tiles = tiles_x * tiles_y;
    rc = check_launch(ctx, launch_clahe(ctx, sl, d_src, spitch, vec4, cp), "clahe");
    if (rc) return rc;
    return check_launch(ctx, launch_pyramid(ctx, sl), "pyramid");
}

int rdfe_preprocess_batch(rdfe_ctx *ctx, const int *slots, int n, const uint8_t *const *images, size_t pitch,
                          double clip_limit, int tiles_x, int tiles_y) {
    SlotList sl;
    int rc = check_slots(ctx, slots, n, &sl, "rdfe_preprocess_batch");
    if (rc) return rc;
    if (!images || pitch < (size_t)ctx->cfg.width * ctx->in_channels) { set_error("rdfe_preprocess_batch: bad image pointers/pitch"); return RDFE_ERR_INVALID; }
    RDFE_CUDA_OK(cudaSetDevice(ctx->cfg.device));
    std::vector<const uint8_t *> dptr(n);
    for (int i = 0; i < n; ++i)
        if (!images[i]) { set_error("rdfe_preprocess_batch: image %d is null", i); return RDFE_ERR_INVALID; }
    for (int i = 0; i < n; ++i) ctx->slot_clahe_step[slots[i]] = -1;      // raw staging now ordered on the context stream
    rc = upload_frames(ctx, ctx->stream, slots, n, images, pitch, dptr);
    if (rc) return rc;
    rc = rdfe_preprocess_batch_dev(ctx, slots, n, dptr.data(), ctx->raw_pitch, clip_limit, tiles_x, tiles_y);
    if (rc) return rc;
    if (!ctx->host_sync) return RDFE_OK;          // rdfe_set_host_sync(ctx, 0): results are consumed in stream order
    RDFE_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    return RDFE_OK;
}

// ----------------------------------------------------------------- detect
// a prefetch is usable when it covers exactly these slots with the same GFTT parameters (the Poisson radius and the
// border only matter to the append stage, which always runs in the detect call itself)
static bool prefetch_matches(const rdfe_ctx *ctx, const int *slots, int n, const rdfe_detect_params *p) {
    if (n != ctx->pf_n) return false;
    for (int i = 0; i < n; ++i)
        if (slots[i] != ctx->pf_slots[i]) return false;
    const rdfe_detect_params &q = ctx->pf_params;
    return p->max_points == q.max_points && p->quality_level == q.quality_level && p->min_distance == q.min_distance &&
           p->harris_k == q.harris_k && p->harris_fma == q.harris_fma;
}

static int check_detect(const rdfe_ctx *ctx, const rdfe_detect_params *p, int stride, const char *what) {
    if (!p || p->max_points < 1 || p->max_points > ctx->cfg.max_points || stride < 1) {
        set_error("%s: max_points=%d (capacity %d) stride=%d invalid", what, p ? p->max_points : -1, ctx->cfg.max_points, stride);
        return RDFE_ERR_INVALID;
    }
    if (!(p->quality_level > 0) || p->min_distance < 0 || !(p->keypoint_distance > 0)) {
        set_error("%s: quality_level/min_distance/keypoint_distance invalid", what);
        return RDFE_ERR_INVALID;
    }
    return RDFE_OK;
}

int rdfe_detect_batch_dev(rdfe_ctx *ctx, const int *slots, int n, const rdfe_detect_params *p, double *dev_keypoints_xy,
                          int *dev_counts, int stride, float *dev_gftt_xy, float *dev_gftt_resp, int *dev_gftt_counts) {
    SlotList sl;
    int rc = check_slots(ctx, slots, n, &sl, "rdfe_detect_batch_dev");
    if (rc) return rc;
    rc = check_detect(ctx, p, stride, "rdfe_detect_batch_dev");
    if (rc) return rc;
    RDFE_CUDA_OK(cudaSetDevice(ctx->cfg.device));
    if (ctx->pf_valid && !dev_gftt_xy && !dev_gftt_resp && !dev_gftt_counts && prefetch_matches(ctx, slots, n, p)) {
        // Harris + GFTT selection of exactly these slots already ran (or still run) on the prefetch stream
        ctx->pf_valid = false;
        RDFE_CUDA_OK(cudaStreamWaitEvent(ctx->stream, ctx->pf_done, 0));
        return check_launch(ctx, launch_poisson_append(ctx, n, *p, ctx->pf_gftt_xy, ctx->pf_gftt_counts, dev_keypoints_xy,
                                                       dev_counts, stride, nullptr), "poisson");
    }
    rc = check_launch(ctx, launch_harris_candidates(ctx, sl, *p, nullptr), "harris");
    if (rc) return rc;
    return check_launch(ctx, launch_select(ctx, sl, *p, dev_keypoints_xy, dev_counts, stride, dev_gftt_xy, dev_gftt_resp,
                                           dev_gftt_counts), "select");
}

int rdfe_detect_prefetch(rdfe_ctx *ctx, const int *slots, int n, const rdfe_detect_params *p) {
    SlotList sl;
    int rc = check_slots(ctx, slots, n, &sl, "rdfe_detect_prefetch");
    if (rc) return rc;
    rc = check_detect(ctx, p, 1, "rdfe_detect_prefetch");
    if (rc) return rc;
    RDFE_CUDA_OK(cudaSetDevice(ctx->cfg.device));
    cudaStream_t ax = ctx->aux_stream2;
    RDFE_CUDA_OK(cudaEventRecord(ctx->ev_fork, ctx->stream));          // level 0 of the slots is produced on the main stream
    RDFE_CUDA_OK(cudaStreamWaitEvent(ax, ctx->ev_fork, 0));
    const DetectScratch det_keep = ctx->det;
    ctx->det = ctx->det2;
    ctx->ls = ax;
    rc = check_launch(ctx, launch_harris_candidates(ctx, sl, *p, nullptr), "harris");
    if (rc == RDFE_OK) rc = check_launch(ctx, launch_gftt_select(ctx, sl, *p, ctx->pf_gftt_xy, ctx->pf_gftt_resp, ctx->pf_gftt_counts), "select");
    ctx->det = det_keep;
    ctx->ls = ctx->stream;
    if (rc) return rc;
    RDFE_CUDA_OK(cudaEventRecord(ctx->pf_done, ax));
    ctx->pf_recorded = true;
    ctx->pf_valid = true;
    ctx->pf_n = n;
    for (int i = 0; i < n; ++i) ctx->pf_slots[i] = slots[i];
    ctx->pf_params = *p;
    return RDFE_OK;
}

int rdfe_set_template_cache(rdfe_ctx *ctx, int on) {
    if (!ctx) return RDFE_ERR_INVALID;
    RDFE_CUDA_OK(cudaSetDevice(ctx->cfg.device));
    { const int rc_all = sync_all_streams(ctx); if (rc_all) return rc_all; }
    if (on && !ctx->tc_data) {
        const size_t entries = (size_t)ctx->cfg.num_slots * ctx->cfg.max_points;
        const size_t rec = lk_cache_record_bytes(ctx->pyr.win, ctx->pyr.nlevels);
        if (cudaMalloc(&ctx->tc_data, entries * rec) != cudaSuccess || cudaMalloc(&ctx->tc_A, entries * ctx->pyr.nlevels * sizeof(float4)) != cudaSuccess ||
            cudaMalloc(&ctx->tc_hdr, entries * sizeof(float4)) != cudaSuccess || cudaMalloc(&ctx->tc_stats, 2 * sizeof(unsigned long long)) != cudaSuccess) {
            cudaGetLastError();
            cudaFree(ctx->tc_data); cudaFree(ctx->tc_A); cudaFree(ctx->tc_hdr);
            ctx->tc_data = nullptr; ctx->tc_A = nullptr; ctx->tc_hdr = nullptr;
            set_error("rdfe_set_template_cache: %zu MB for %zu (slot, point) records not available",
                      (entries * (rec + 16 + ctx->pyr.nlevels * 16)) >> 20, entries);
            return RDFE_ERR_NOMEM;
        }
        RDFE_CUDA_OK(cudaMemset(ctx->tc_hdr, 0, entries * sizeof(float4)));    // generation 0 never matches (slots start at >= 2)
        RDFE_CUDA_OK(cudaMemset(ctx->tc_stats, 0, 2 * sizeof(unsigned long long)));
    }
    ctx->tc_on = on != 0;
    return RDFE_OK;
}

int rdfe_template_cache_stats(rdfe_ctx *ctx, unsigned long long *lookups, unsigned long long *hits, int reset) {
    if (!ctx) return RDFE_ERR_INVALID;
    unsigned long long v[2] = {0, 0};
    if (ctx->tc_stats) {
        RDFE_CUDA_OK(cudaSetDevice(ctx->cfg.device));
        { const int rc_all = sync_all_streams(ctx); if (rc_all) return rc_all; }
        RDFE_CUDA_OK(cudaMemcpy(v, ctx->tc_stats, sizeof v, cudaMemcpyDeviceToHost));
        if (reset) RDFE_CUDA_OK(cudaMemset(ctx->tc_stats, 0, sizeof v));
    }
    if (lookups) *lookups = v[0];
    if (hits) *hits = v[1];
    return RDFE_OK;
}

int rdfe_set_step_compaction(rdfe_ctx *ctx, int on) {
    if (!ctx) return RDFE_ERR_INVALID;
    ctx->compact_tracked = on != 0;
    return RDFE_OK;
}

int rdfe_set_host_sync(rdfe_ctx *ctx, int on) {
    if (!ctx) return RDFE_ERR_INVALID;
    ctx->host_sync = on != 0;
    return RDFE_OK;
}

int rdfe_detect_batch(rdfe_ctx *ctx, const int *slots, int n, const rdfe_detect_params *p, double *keypoints_xy, int *counts,
                      int stride, float *gftt_xy, float *gftt_resp, int *gftt_counts) {
    if (!ctx || !keypoints_xy || !counts) { set_error("rdfe_detect_batch: null argument"); return RDFE_ERR_INVALID; }
    if (n < 1 || n > RDFE_MAX_BATCH || stride < 1 || stride > ctx->cfg.max_points) {
        set_error("rdfe_detect_batch: n=%d stride=%d out of range (capacity %d)", n, stride, ctx->cfg.max_points);
        return RDFE_ERR_INVALID;
    }
    for (int i = 0; i < n; ++i)
        if (counts[i] < 0 || counts[i] > stride) { set_error("rdfe_detect_batch: counts[%d]=%d out of range", i, counts[i]); return RDFE_ERR_INVALID; }
    RDFE_CUDA_OK(cudaSetDevice(ctx->cfg.device));
    if (n > ctx->max_batch) { set_error("rdfe_detect_batch: batch %d exceeds the context's %d slots", n, ctx->max_batch); return RDFE_ERR_INVALID; }
    // pinned block [xy | counts]: one copy up, one copy back
    const size_t xyb = (size_t)n * stride * 2 * sizeof(double), iob = xyb + (size_t)n * sizeof(int);
    memcpy(ctx->h_io, keypoints_xy, xyb);
    memcpy(ctx->h_io + xyb, counts, (size_t)n * sizeof(int));
    double *d_xy = reinterpret_cast<double *>(ctx->d_io);
    int *d_cnt = reinterpret_cast<int *>(ctx->d_io + xyb);
    RDFE_CUDA_OK(cudaMemcpyAsync(ctx->d_io, ctx->h_io, iob, cudaMemcpyHostToDevice, ctx->stream));
    int rc = rdfe_detect_batch_dev(ctx, slots, n, p, d_xy, d_cnt, stride, gftt_xy ? ctx->d_gftt_xy : nullptr,
                                   gftt_resp ? ctx->d_gftt_resp : nullptr, gftt_counts ? ctx->d_gftt_counts : nullptr);
    if (rc) return rc;
    RDFE_CUDA_OK(cudaMemcpyAsync(ctx->h_io, ctx->d_io, iob, cudaMemcpyDeviceToHost, ctx->stream));
    const size_t k = (size_t)p->max_points;
    if (gftt_xy) RDFE_CUDA_OK(cudaMemcpyAsync(gftt_xy, ctx->d_gftt_xy, n * k * 2 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    if (gftt_resp) RDFE_CUDA_OK(cudaMemcpyAsync(gftt_resp, ctx->d_gftt_resp, n * k * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    if (gftt_counts) RDFE_CUDA_OK(cudaMemcpyAsync(gftt_counts, ctx->d_gftt_counts, n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    RDFE_CUDA_OK(cudaMemcpyAsync(ctx->h_overflow, ctx->det.overflow, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
    RDFE_CUDA_OK(cudaStreamSynchronize(ctx->stream));      // everything this call launched is ordered on the main stream
    memcpy(keypoints_xy, ctx->h_io, xyb);
    memcpy(counts, ctx->h_io + xyb, (size_t)n * sizeof(int));
    if (*ctx->h_overflow) {
        cudaMemsetAsync(ctx->det.overflow, 0, sizeof(unsigned), ctx->stream);
        return report_overflow(ctx, *ctx->h_overflow);
    }
    return RDFE_OK;
}

// ------------------------------------------------------------------ track
int rdfe_track_batch_dev(rdfe_ctx *ctx, const int *curr_slots, const int *next_slots, int n, const rdfe_track_params *p,
                         const double *dev_curr_xy, double *dev_next_xy, const int *dev_counts, int stride, char *dev_status) {
    SlotList sc, sn;
    int rc = check_slots(ctx, curr_slots, n, &sc, "rdfe_track_batch_dev(curr)");
    if (rc) return rc;
    rc = check_slots(ctx, next_slots, n, &sn, "rdfe_track_batch_dev(next)");
    if (rc) return rc;
    if (!p || !dev_curr_xy || !dev_next_xy || !dev_counts || !dev_status || stride < 1) {
        set_error("rdfe_track_batch_dev: null argument or stride < 1");
        return RDFE_ERR_INVALID;
    }
    RDFE_CUDA_OK(cudaSetDevice(ctx->cfg.device));
    return check_launch(ctx, launch_lk(ctx, sc, sn, *p, dev_curr_xy, dev_next_xy, dev_counts, stride, dev_status), "lk");
}

int rdfe_track_batch(rdfe_ctx *ctx, const int *curr_slots, const int *next_slots, int n, const rdfe_track_params *p,
                     const double *curr_xy, double *next_xy, const int *counts, int stride, char *status) {
    if (!ctx || !curr_xy || !next_xy || !counts || !status || !p) { set_error("rdfe_track_batch: null argument"); return RDFE_ERR_INVALID; }
    if (n < 1 || n > RDFE_MAX_BATCH || stride < 1 || stride > ctx->cfg.max_points) {
        set_error("rdfe_track_batch: n=%d stride=%d out of range (capacity %d)", n, stride, ctx->cfg.max_points);
        return RDFE_ERR_INVALID;
    }
    for (int i = 0; i < n; ++i)
        if (counts[i] < 0 || counts[i] > stride) { set_error("rdfe_track_batch: counts[%d]=%d out of range", i, counts[i]); return RDFE_ERR_INVALID; }
    RDFE_CUDA_OK(cudaSetDevice(ctx->cfg.device));
    if (n > ctx->max_batch) { set_error("rdfe_track_batch: batch %d exceeds the context's %d slots", n, ctx->max_batch); return RDFE_ERR_INVALID; }
    // pinned block [curr | counts | next | status]: [curr | counts (| next = prediction)] goes up in one copy,
    // [next | status] comes back in one copy.  LK writes the status of every point below counts[i]; entries beyond are 0.
    const size_t xyb = (size_t)n * stride * 2 * sizeof(double), cb = ((size_t)n * sizeof(int) + 15) & ~(size_t)15;
    const size_t off_cnt = xyb, off_next = xyb + cb, off_st = off_next + xyb, stb = (size_t)n * stride;
    memcpy(ctx->h_io, curr_xy, xyb);
    memcpy(ctx->h_io + off_cnt, counts, (size_t)n * sizeof(int));
    if (p->has_prediction) memcpy(ctx->h_io + off_next, next_xy, xyb);
    RDFE_CUDA_OK(cudaMemcpyAsync(ctx->d_io, ctx->h_io, p->has_prediction ? off_st : off_next, cudaMemcpyHostToDevice, ctx->stream));
    int rc = rdfe_track_batch_dev(ctx, curr_slots, next_slots, n, p, reinterpret_cast<const double *>(ctx->d_io),
                                  reinterpret_cast<double *>(ctx->d_io + off_next), reinterpret_cast<const int *>(ctx->d_io + off_cnt), stride,
                                  reinterpret_cast<char *>(ctx->d_io + off_st));
    if (rc) return rc;
    RDFE_CUDA_OK(cudaMemcpyAsync(ctx->h_io + off_next, ctx->d_io + off_next, xyb + stb, cudaMemcpyDeviceToHost, ctx->stream));
    RDFE_CUDA_OK(cudaStreamSynchronize(ctx->stream));      // the LK launch and the copy back are ordered on the main stream
    // only status != 0 entries of next_xy may change (opencv_image.cpp:148-153): merge on the host
    const double *hn = reinterpret_cast<const double *>(ctx->h_io + off_next);
    const char *hst = reinterpret_cast<const char *>(ctx->h_io + off_st);
    memset(status, 0, stb);
    for (int b = 0; b < n; ++b)
        for (int i = 0; i < counts[b]; ++i) {
            const size_t k = (size_t)b * stride + i;
            status[k] = hst[k];
            if (hst[k]) { next_xy[2 * k] = hn[2 * k]; next_xy[2 * k + 1] = hn[2 * k + 1]; }
        }
    return RDFE_OK;
}

// Frame::track_keypoints' keypoint prediction (frame.cpp:82-93) for keypoints that stay on the device between steps:
// next = apply_k(delta_q * bearing_i, K_next) with bearing_i = remove_k(curr_i, K) is the homography K_next R K^-1
// applied to the pixel.  dev_H: [n][9] row-major doubles, one per stream.
int rdfe_predict_rotation_dev(rdfe_ctx *ctx, int n, const double *dev_H, const double *dev_curr_xy, const int *dev_counts,
                              int stride, double *dev_pred_xy) {
    if (!ctx || !dev_H || !dev_curr_xy || !dev_counts || !dev_pred_xy || n < 1 || n > ctx->max_batch || stride < 1) {
        set_error("rdfe_predict_rotation_dev: bad argument");
        return RDFE_ERR_INVALID;
    }
    RDFE_CUDA_OK(cudaSetDevice(ctx->cfg.device));
    return check_launch(ctx, launch_predict_rotation(ctx, n, dev_H, dev_curr_xy, dev_counts, stride, dev_pred_xy), "predict");
}

// ------------------------------------------------- fused per-frame step
// FeatureTracker::run's plugin calls for one new frame per stream (feature_tracker.cpp:32-98) in ONE call:
// preprocess(new) -> track(prev->new) -> detect(new, existing = tracked positions).  The GFTT selection
// (which does not depend on the tracked points) runs on the auxiliary stream concurrently with LK; only the
// Poisson-disk append waits for both.
int rdfe_frontend_step_dev(rdfe_ctx *ctx, const int *prev_slots, const int *new_slots, int n,
                           const uint8_t *const *dev_images, size_t pitch, double clip_limit, int tiles_x, int tiles_y,
                           const rdfe_track_params *tp, const double *dev_curr_xy, double *dev_next_xy,
                           const int *dev_track_counts, char *dev_status, const rdfe_detect_params *dp,
                           int *dev_kp_counts, int stride) {
    SlotList sn, spv;
    int rc = check_slots(ctx, new_slots, n, &sn, "rdfe_frontend_step_dev(new)");
    if (rc) return rc;
    if (prev_slots) {
        rc = check_slots(ctx, prev_slots, n, &spv, "rdfe_frontend_step_dev(prev)");
        if (rc) return rc;
        if (!tp || !dev_curr_xy || !dev_track_counts || !dev_status) { set_error("rdfe_frontend_step_dev: null track argument"); return RDFE_ERR_INVALID; }
    }
    if (!dev_next_xy || !dev_kp_counts) { set_error("rdfe_frontend_step_dev: null keypoint buffers"); return RDFE_ERR_INVALID; }
    rc = check_detect(ctx, dp, stride, "rdfe_frontend_step_dev");
    if (rc) return rc;
    // CLAHE (level 0 + halo) first: both branches need it
    ClaheParams cp;
    rc = make_clahe_params(ctx, clip_limit, tiles_x, tiles_y, &cp);
    if (rc) return rc;
    if (!dev_images || pitch < (size_t)ctx->cfg.width * ctx->in_channels) { set_error("rdfe_frontend_step_dev: bad image pointers/pitch"); return RDFE_ERR_INVALID; }
    int vec4 = (pitch % 4 == 0) ? 1 : 0;
    for (int i = 0; i < n; ++i) {
        if (!dev_images[i]) { set_error("rdfe_frontend_step_dev: image %d is null", i); return RDFE_ERR_INVALID; }
        if ((uintptr_t)dev_images[i] % 4) vec4 = 0;
    }
    RDFE_CUDA_OK(cudaSetDevice(ctx->cfg.device));
    const bool ov = ctx->overlap && (!ctx->prof_on || ctx->prof_timeline);                // detection branch beside tracking branch
    const bool pipe = ov && ctx->pipeline_steps;                   // preprocess of this step beside the previous step
    const int par = (int)(ctx->step_index & 1);
    // per-step scratch alternates so that consecutive steps may overlap
    uint8_t *lut_keep = ctx->lut;
    const uint8_t **src_keep = ctx->d_srcptrs;
    float *gxy = par ? ctx->d_gftt_xy2 : ctx->d_gftt_xy, *gre = par ? ctx->d_gftt_resp2 : ctx->d_gftt_resp;
    int *gcn = par ? ctx->d_gftt_counts2 : ctx->d_gftt_counts;
    if (par) { ctx->lut = ctx->lut2; ctx->d_srcptrs = ctx->d_srcptrs2; }
    if (pipe) {
        // ---- pipelined schedule: one stream per kernel class, events for the data dependences only.
        //   pre:    hist, apply | pyrdown, scharr          waits LK(s-2) (last reader of the slot set rewritten here)
        //   harris: reset, harris                           waits apply(s), select(s-2) (candidate buffers of this parity)
        //   select: select                                  waits harris(s), poisson(s-2) (gftt buffers of this parity)
        //   track:  LK                                      waits pyramid(s), poisson(s-1) (tracked points), caller's stream
        //   post:   poisson                                 waits LK(s), select(s); its end is the step's end
        auto restore_p = [&]() { ctx->lut = lut_keep; ctx->d_srcptrs = src_keep; ctx->ls = ctx->stream; };
        cudaStream_t ps = ctx->pre_stream, hs = par ? ctx->aux_stream2 : ctx->aux_stream, ss = ctx->sel_stream[par];
        cudaStream_t ts = ctx->trk_stream, po = ctx->post_stream;
        cudaEvent_t evj = par ? ctx->ev_join2 : ctx->ev_join;
        bool clash = false;                               // new slots still read by the previous step?
        for (int i = 0; i < n; ++i) clash |= ctx->last_step_slots[new_slots[i]] != 0;
        RDFE_CUDA_OK(cudaEventRecord(ctx->ev_entry, ctx->stream));
        if (ctx->pf_recorded) RDFE_CUDA_OK(cudaStreamWaitEvent(ps, ctx->pf_done, 0));   // a detect prefetch may read these slots
        ctx->pf_valid = false;
        bool renew = false;                               // ... or the NEW slots of step s-2 (Harris(s-2) read them)?
        for (int i = 0; i < n; ++i) renew |= ctx->slot_new_step[new_slots[i]] == (long long)ctx->step_index - 2;
        RDFE_CUDA_OK(cudaStreamWaitEvent(ps, clash ? ctx->ev_step_done[par ^ 1] : ctx->ev_lk_done[par], 0));
        if (renew) RDFE_CUDA_OK(cudaStreamWaitEvent(ps, ctx->ev_step_done[par], 0));
        if (ctx->images_ready_valid) RDFE_CUDA_OK(cudaStreamWaitEvent(ps, ctx->images_ready, 0));
        ctx->ls = ps;
        const uint8_t *const *d_src = nullptr;
        size_t spitch = 0;
        rc = stage_sources(ctx, new_slots, n, dev_images, pitch, &d_src, &spitch, &vec4);
        if (rc) { restore_p(); return rc; }
        ctx->last_clahe_tiles = tiles_x * tiles_y;
        rc = check_launch(ctx, launch_clahe(ctx, sl_copy(sn), d_src, spitch, vec4, cp), "clahe");
        if (rc) { restore_p(); return rc; }
        cudaEventRecord(ctx->ev_clahe_done, ps);
        cudaEventRecord(ctx->ev_clahe_ring[ctx->step_index & 3], ps);
        for (int i = 0; i < n; ++i) ctx->slot_clahe_step[new_slots[i]] = (long long)ctx->step_index;
        cudaEventRecord(ctx->ev_apply_done, ps);
        // The preprocess chain is the longest dependent chain of the pipelined step (its small kernels wait for SM slots
        // beside LK and Harris).  Scharr of level 0 -- 3/4 of the derivative work -- needs only the CLAHE output, so it runs
        // on a second preprocess stream beside the pyrDown chain; the small levels' Scharr follows the last pyrDown.
        static const int s_split = [] { const char *e = getenv("RDFE_SCHARR_SPLIT"); return e ? atoi(e) : 1; }();
        if (s_split && ctx->pyr.nlevels > 1) {
            cudaStreamWaitEvent(ctx->pre_stream2, ctx->ev_apply_done, 0);
            ctx->ls = ctx->pre_stream2;
            rc = check_launch(ctx, launch_scharr_levels(ctx, sn, 0, 1), "scharr0");
            cudaEventRecord(ctx->ev_sc0_done, ctx->pre_stream2);
            ctx->ls = ps;
            if (rc == RDFE_OK) rc = check_launch(ctx, launch_pyrdowns(ctx, sn), "pyrdown");
            if (rc == RDFE_OK) rc = check_launch(ctx, launch_scharr_levels(ctx, sn, 1, ctx->pyr.nlevels), "scharr1");
            cudaStreamWaitEvent(ps, ctx->ev_sc0_done, 0);
        } else
            rc = check_launch(ctx, launch_pyramid(ctx, sn), "pyramid");
        if (rc) { restore_p(); return rc; }
        cudaEventRecord(ctx->ev_pre_done, ps);
        // detection branch
        const DetectScratch det_keep = ctx->det;
        if (par) ctx->det = ctx->det2;
        cudaStreamWaitEvent(hs, ctx->ev_apply_done, 0);
        cudaStreamWaitEvent(hs, evj, 0);                  // select(s-2) has read this parity's candidates
        ctx->ls = hs;
        rc = check_launch(ctx, launch_harris_candidates(ctx, sn, *dp, nullptr), "harris");
        cudaEventRecord(ctx->ev_harris_done[par], hs);
        if (rc == RDFE_OK) {
            cudaStreamWaitEvent(ss, ctx->ev_harris_done[par], 0);
            cudaStreamWaitEvent(ss, ctx->ev_step_done[par], 0);   // poisson(s-2) has read this parity's gftt output
            ctx->ls = ss;
            rc = check_launch(ctx, launch_gftt_select(ctx, sn, *dp, gxy, gre, gcn), "select");
            cudaEventRecord(evj, ss);
        }
        ctx->det = det_keep;
        if (rc) { restore_p(); return rc; }
        // tracking branch
        cudaStreamWaitEvent(ts, ctx->ev_entry, 0);
        cudaStreamWaitEvent(ts, ctx->ev_pre_done, 0);
        cudaStreamWaitEvent(ts, ctx->ev_step_done[par ^ 1], 0);
        if (prev_slots) {
            ctx->ls = ts;
            rc = check_launch(ctx, launch_lk(ctx, spv, sn, *tp, dev_curr_xy, dev_next_xy, dev_track_counts, stride, dev_status), "lk");
            if (rc) { restore_p(); return rc; }
        }
        cudaEventRecord(ctx->ev_lk_done[par], ts);
        cudaStreamWaitEvent(po, ctx->ev_lk_done[par], 0);
        cudaStreamWaitEvent(po, evj, 0);
        ctx->ls = po;
        rc = check_launch(ctx, launch_poisson_append(ctx, n, *dp, gxy, gcn, dev_next_xy, dev_kp_counts, stride,
                                                     (prev_slots && ctx->compact_tracked) ? dev_status : nullptr), "poisson");
        restore_p();
        if (rc) return rc;
        RDFE_CUDA_OK(cudaEventRecord(ctx->ev_step_done[par], po));
        RDFE_CUDA_OK(cudaStreamWaitEvent(ctx->stream, ctx->ev_step_done[par], 0));
        memset(ctx->last_step_slots, 0, (size_t)ctx->cfg.num_slots);
        for (int i = 0; i < n; ++i) {
            ctx->last_step_slots[new_slots[i]] = 1;
            ctx->slot_new_step[new_slots[i]] = (long long)ctx->step_index;
            if (prev_slots) ctx->last_step_slots[prev_slots[i]] = 1;
        }
        ctx->step_index++;
        return RDFE_OK;
    }
    cudaStream_t ps = ctx->stream;
    if (ctx->pf_recorded) RDFE_CUDA_OK(cudaStreamWaitEvent(ps, ctx->pf_done, 0));       // a detect prefetch may read these slots
    ctx->pf_valid = false;
    auto restore = [&]() { ctx->lut = lut_keep; ctx->d_srcptrs = src_keep; ctx->ls = ctx->stream; };
    // ---- preprocess: CLAHE (level 0 + halo), pyramid, Scharr
    ctx->ls = ps;
    const uint8_t *const *d_src = nullptr;
    size_t spitch = 0;
    rc = stage_sources(ctx, new_slots, n, dev_images, pitch, &d_src, &spitch, &vec4);
    if (rc) { restore(); return rc; }
    ctx->last_clahe_tiles = tiles_x * tiles_y;
    rc = check_launch(ctx, launch_clahe(ctx, sl_copy(sn), d_src, spitch, vec4, cp), "clahe");
    if (rc) { restore(); return rc; }
    cudaEventRecord(ctx->ev_clahe_done, ps);
    cudaEventRecord(ctx->ev_clahe_ring[ctx->step_index & 3], ps);
    for (int i = 0; i < n; ++i) ctx->slot_clahe_step[new_slots[i]] = (long long)ctx->step_index;
    cudaEventRecord(ctx->ev_apply_done, ps);
    // ---- detection branch (Harris needs only level 0): auxiliary stream
    cudaStream_t axs = ctx->aux_stream;
    cudaEvent_t evj = ctx->ev_join;
    if (ov) {
        cudaStreamWaitEvent(axs, ctx->ev_apply_done, 0);
        ctx->ls = axs;
    }
    rc = check_launch(ctx, launch_harris_candidates(ctx, sn, *dp, nullptr), "harris");
    if (rc == RDFE_OK) rc = check_launch(ctx, launch_gftt_select(ctx, sn, *dp, gxy, gre, gcn), "select");
    if (rc) { restore(); return rc; }
    if (ov) cudaEventRecord(evj, axs);
    // ---- tracking branch: pyramid levels + Scharr (still on the preprocess stream), then LK on the main stream
    ctx->ls = ps;
    rc = check_launch(ctx, launch_pyramid(ctx, sn), "pyramid");
    if (rc) { restore(); return rc; }
    restore();
    if (prev_slots) {
        rc = check_launch(ctx, launch_lk(ctx, spv, sn, *tp, dev_curr_xy, dev_next_xy, dev_track_counts, stride, dev_status), "lk");
        if (rc) return rc;
    }
    if (ov) RDFE_CUDA_OK(cudaStreamWaitEvent(ctx->stream, evj, 0));
    rc = check_launch(ctx, launch_poisson_append(ctx, n, *dp, gxy, gcn, dev_next_xy, dev_kp_counts, stride,
                                                     (prev_slots && ctx->compact_tracked) ? dev_status : nullptr), "poisson");
    if (rc) return rc;
    RDFE_CUDA_OK(cudaEventRecord(ctx->ev_step_done[par], ctx->stream));
    memset(ctx->last_step_slots, 0, (size_t)ctx->cfg.num_slots);
    for (int i = 0; i < n; ++i) { ctx->last_step_slots[new_slots[i]] = 1; if (prev_slots) ctx->last_step_slots[prev_slots[i]] = 1; }
    ctx->step_index++;
    return RDFE_OK;
}

int rdfe_set_undistort(rdfe_ctx *ctx, const float *K, const float *D) {
    if (!ctx) return RDFE_ERR_INVALID;
    RDFE_CUDA_OK(cudaSetDevice(ctx->cfg.device));
    int rc = rdfe_sync(ctx);
    if (rc) return rc;
    if (!K || !D) { ctx->und_on = false; return RDFE_OK; }
    const int W = ctx->cfg.width, H = ctx->cfg.height;
    std::vector<uint32_t> mxy;
    std::vector<uint16_t> mf;
    build_undistort_map(W, H, K, D, mxy, mf);
    if (!ctx->und_map_xy) {
        RDFE_CUDA_OK(cudaMalloc(&ctx->und_map_xy, (size_t)W * H * sizeof(uint32_t)));
        RDFE_CUDA_OK(cudaMalloc(&ctx->und_map_f, (size_t)W * H * sizeof(uint16_t)));
    }
    if (!ctx->und_plane) RDFE_CUDA_OK(cudaMalloc(&ctx->und_plane, ctx->gray_slot * ctx->cfg.num_slots));
    RDFE_CUDA_OK(cudaMemcpy(ctx->und_map_xy, mxy.data(), mxy.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    RDFE_CUDA_OK(cudaMemcpy(ctx->und_map_f, mf.data(), mf.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    ctx->und_on = true;
    return RDFE_OK;
}

int rdfe_set_input_format(rdfe_ctx *ctx, int channels) {
    if (!ctx || (channels != 1 && channels != 3 && channels != 4)) { set_error("rdfe_set_input_format: channels must be 1, 3 (BGR) or 4 (BGRA)"); return RDFE_ERR_INVALID; }
    RDFE_CUDA_OK(cudaSetDevice(ctx->cfg.device));
    int rc = rdfe_sync(ctx);
    if (rc) return rc;
    if (channels == ctx->in_channels) return RDFE_OK;
    // upload staging holds the frames as given (W * channels bytes per row)
    cudaFree(ctx->raw);
    ctx->raw = nullptr;
    ctx->raw_pitch = align_up((size_t)ctx->cfg.width * channels, 4);
    ctx->raw_slot = align_up(ctx->raw_pitch * ctx->cfg.height, 256);
    RDFE_CUDA_OK(cudaMalloc(&ctx->raw, ctx->raw_slot * ctx->cfg.num_slots));
    if (!ctx->und_plane) RDFE_CUDA_OK(cudaMalloc(&ctx->und_plane, ctx->gray_slot * ctx->cfg.num_slots));
    ctx->in_channels = channels;
    return RDFE_OK;
}

int rdfe_set_pipelining(rdfe_ctx *ctx, int on) {
    if (!ctx) return RDFE_ERR_INVALID;
    { const int rc_all = sync_all_streams(ctx); if (rc_all) return rc_all; }
    ctx->pipeline_steps = on != 0;
    return RDFE_OK;
}

// ------------------------------------- pipelined host-buffer step (kPipeDepth = 3 stages)
// submit(t+1) may be called before wait(t): the frames of step t+1 are uploaded on a copy stream while the
// kernels of step t run; results come back through pinned staging and are handed out by wait().
int rdfe_frontend_step_submit(rdfe_ctx *ctx, const int *prev_slots, const int *new_slots, int n,
                              const uint8_t *const *images, size_t pitch, double clip_limit, int tiles_x, int tiles_y,
                              const rdfe_track_params *tp, const double *curr_xy, const double *pred_xy, const int *counts,
                              const rdfe_detect_params *dp, int stride, int *ticket) {
    if (!ctx || !new_slots || !images || !counts || !dp || !ticket) { set_error("rdfe_frontend_step_submit: null argument"); return RDFE_ERR_INVALID; }
    if (n < 1 || n > RDFE_MAX_BATCH || stride < 1 || stride > ctx->cfg.max_points) {
        set_error("rdfe_frontend_step_submit: n=%d stride=%d out of range (capacity %d)", n, stride, ctx->cfg.max_points);
        return RDFE_ERR_INVALID;
    }
    if (prev_slots && (!tp || !curr_xy)) { set_error("rdfe_frontend_step_submit: tracking needs tp and curr_xy"); return RDFE_ERR_INVALID; }
    if (pitch < (size_t)ctx->cfg.width * ctx->in_channels) { set_error("rdfe_frontend_step_submit: pitch < width * channels"); return RDFE_ERR_INVALID; }
    const int p = (int)(ctx->pl_ticket % kPipeDepth);
    if (ctx->pl_busy[p]) { set_error("rdfe_frontend_step_submit: stage %d still holds un-waited results (at most 3 steps in flight)", p); return RDFE_ERR_INVALID; }
    for (int i = 0; i < n; ++i) {
        if (new_slots[i] < 0 || new_slots[i] >= ctx->cfg.num_slots || !ctx->slot_used[new_slots[i]] || !images[i]) {
            set_error("rdfe_frontend_step_submit: bad slot or image at index %d", i);
            return RDFE_ERR_INVALID;
        }
        if (counts[i] < 0 || counts[i] > stride) { set_error("rdfe_frontend_step_submit: counts[%d]=%d out of range", i, counts[i]); return RDFE_ERR_INVALID; }
    }
    RDFE_CUDA_OK(cudaSetDevice(ctx->cfg.device));
    const size_t xyb = (size_t)n * stride * 2 * sizeof(double);
    // ---- copy stream: frames into the upload staging of the new slots, keypoints into stage p
    // The upload overwrites the raw staging of the NEW slots: it must wait for the last CLAHE that read them.  With
    // slot sets in rotation that is the CLAHE of an older step (s-3 for three sets), not the one just enqueued, so
    // the upload of step s+1 runs while step s is still waiting for its own frames' kernels.
    {
        long long last = -1;
        bool unknown = false;
        for (int i = 0; i < n; ++i) {
            const long long v = ctx->slot_clahe_step[new_slots[i]];
            if (v < 0) unknown = true;
            if (v > last) last = v;
        }
        const long long cur = (long long)ctx->step_index;
        if (unknown || cur == 0) {                         // conservative: everything enqueued so far
            RDFE_CUDA_OK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_clahe_done, 0));
            RDFE_CUDA_OK(cudaEventRecord(ctx->ev_copy_fence, ctx->stream));
            RDFE_CUDA_OK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_copy_fence, 0));
        }
        else {
            if (last < cur - 4) last = cur - 4;            // older entries were re-recorded; CLAHEs complete in step order
            RDFE_CUDA_OK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_clahe_ring[last & 3], 0));
        }
    }
    RDFE_CUDA_OK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_done[p], 0));        // stage buffers free again
    std::vector<const uint8_t *> dptr(n);
    {
        const int rc_up = upload_frames(ctx, ctx->copy_stream, new_slots, n, images, pitch, dptr);
        if (rc_up) return rc_up;
    }
    rdfe_track_params tpl;
    if (tp) tpl = *tp; else rdfe_default_track_params(&tpl);
    tpl.has_prediction = (prev_slots && pred_xy) ? 1 : 0;
    if (prev_slots) RDFE_CUDA_OK(cudaMemcpyAsync(ctx->pl_curr[p], curr_xy, xyb, cudaMemcpyHostToDevice, ctx->copy_stream));
    // detect's existing keypoints = prediction (if any) overwritten by tracked positions, else curr (no tracking: none)
    if (prev_slots) RDFE_CUDA_OK(cudaMemcpyAsync(ctx->pl_next[p], pred_xy ? pred_xy : curr_xy, xyb, cudaMemcpyHostToDevice, ctx->copy_stream));
    RDFE_CUDA_OK(cudaMemcpyAsync(ctx->pl_counts[p], counts, n * sizeof(int), cudaMemcpyHostToDevice, ctx->copy_stream));
    if (prev_slots) RDFE_CUDA_OK(cudaMemcpyAsync(ctx->pl_kcounts[p], counts, n * sizeof(int), cudaMemcpyHostToDevice, ctx->copy_stream));
    else RDFE_CUDA_OK(cudaMemsetAsync(ctx->pl_kcounts[p], 0, n * sizeof(int), ctx->copy_stream));   // first frame: no carried keypoints
    RDFE_CUDA_OK(cudaEventRecord(ctx->ev_upload[p], ctx->copy_stream));
    // ---- main stream: the step, then results into pinned staging
    RDFE_CUDA_OK(cudaStreamWaitEvent(ctx->stream, ctx->ev_upload[p], 0));
    if (prev_slots) RDFE_CUDA_OK(cudaMemsetAsync(ctx->pl_status[p], 0, (size_t)n * stride, ctx->stream));
    ctx->images_ready = ctx->ev_upload[p];        // the preprocess stream (if pipelining) must see the uploads too
    ctx->images_ready_valid = true;
    int rc = rdfe_frontend_step_dev(ctx, prev_slots, new_slots, n, dptr.data(), ctx->raw_pitch, clip_limit, tiles_x, tiles_y,
                                    &tpl, ctx->pl_curr[p], ctx->pl_next[p], ctx->pl_counts[p], ctx->pl_status[p], dp,
                                    ctx->pl_kcounts[p], stride);
    ctx->images_ready_valid = false;
    if (rc) return rc;
    uint8_t *hs = ctx->pl_host[p];
    RDFE_CUDA_OK(cudaMemcpyAsync(hs, ctx->pl_next[p], xyb, cudaMemcpyDeviceToHost, ctx->stream));
    RDFE_CUDA_OK(cudaMemcpyAsync(hs + xyb, ctx->pl_kcounts[p], n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    RDFE_CUDA_OK(cudaMemcpyAsync(hs + xyb + n * sizeof(int), ctx->pl_status[p], (size_t)n * stride, cudaMemcpyDeviceToHost, ctx->stream));
    RDFE_CUDA_OK(cudaMemcpyAsync(hs + xyb + n * sizeof(int) + (size_t)n * stride, ctx->det.overflow, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
    RDFE_CUDA_OK(cudaEventRecord(ctx->ev_done[p], ctx->stream));
    ctx->pl_n[p] = n; ctx->pl_stride[p] = stride; ctx->pl_busy[p] = prev_slots ? 2 : 1;
    *ticket = (int)(ctx->pl_ticket % (3 * (1LL << 28)));      // stays a multiple-of-3-periodic non-negative int
    ctx->pl_ticket++;
    return RDFE_OK;
}

// Measurement aid: exactly the frame upload of rdfe_frontend_step_submit (same staging, same copy stream, same
// one-2-D-copy path), no kernels.  bench.py times it alone to show what the host side of a box can deliver.
int rdfe_upload_only(rdfe_ctx *ctx, const int *new_slots, int n, const uint8_t *const *images, size_t pitch, int sync) {
    if (!ctx || !new_slots || !images || n < 1 || n > ctx->max_batch) { set_error("rdfe_upload_only: bad argument"); return RDFE_ERR_INVALID; }
    if (pitch < (size_t)ctx->cfg.width * ctx->in_channels) { set_error("rdfe_upload_only: pitch < width * channels"); return RDFE_ERR_INVALID; }
    for (int i = 0; i < n; ++i)
        if (new_slots[i] < 0 || new_slots[i] >= ctx->cfg.num_slots || !ctx->slot_used[new_slots[i]] || !images[i]) {
            set_error("rdfe_upload_only: bad slot or image at index %d", i);
            return RDFE_ERR_INVALID;
        }
    RDFE_CUDA_OK(cudaSetDevice(ctx->cfg.device));
    std::vector<const uint8_t *> dptr(n);
    const int rc = upload_frames(ctx, ctx->copy_stream, new_slots, n, images, pitch, dptr);
    if (rc) return rc;
    if (sync) RDFE_CUDA_OK(cudaStreamSynchronize(ctx->copy_stream));
    return RDFE_OK;
}

int rdfe_frontend_step_wait(rdfe_ctx *ctx, int ticket, double *next_xy, int *kp_counts, char *status) {
    if (!ctx) return RDFE_ERR_INVALID;
    const int p = ticket % kPipeDepth;
    if (!ctx->pl_busy[p]) { set_error("rdfe_frontend_step_wait: ticket %d has no pending step", ticket); return RDFE_ERR_INVALID; }
    RDFE_CUDA_OK(cudaEventSynchronize(ctx->ev_done[p]));
    const int n = ctx->pl_n[p], stride = ctx->pl_stride[p];
    const size_t xyb = (size_t)n * stride * 2 * sizeof(double);
    const uint8_t *hs = ctx->pl_host[p];
    if (next_xy) memcpy(next_xy, hs, xyb);
    if (kp_counts) memcpy(kp_counts, hs + xyb, n * sizeof(int));
    if (status && ctx->pl_busy[p] == 2) memcpy(status, hs + xyb + n * sizeof(int), (size_t)n * stride);
    unsigned ovf;
    memcpy(&ovf, hs + xyb + n * sizeof(int) + (size_t)n * stride, sizeof ovf);
    ctx->pl_busy[p] = 0;
    if (ovf) {
        cudaMemsetAsync(ctx->det.overflow, 0, sizeof(unsigned), ctx->stream);
        return report_overflow(ctx, ovf);
    }
    return RDFE_OK;
}

// ------------------------------------------------------------ parity taps
int rdfe_download_level(rdfe_ctx *ctx, int slot, int level, int plane, void *dst, size_t dst_bytes) {
    if (!ctx || !dst || slot < 0 || slot >= ctx->cfg.num_slots || level < 0 || level >= ctx->pyr.nlevels) {
        set_error("rdfe_download_level: bad slot/level");
        return RDFE_ERR_INVALID;
    }
    RDFE_CUDA_OK(cudaSetDevice(ctx->cfg.device));
    RDFE_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    const LevelGeom &g = ctx->pyr.lv[level];
    const int win = ctx->pyr.win;
    if (plane == 0) {
        if (dst_bytes < (size_t)g.w * g.h) { set_error("rdfe_download_level: buffer too small"); return RDFE_ERR_INVALID; }
        RDFE_CUDA_OK(cudaMemcpy2D(dst, g.w, ctx->pyr.image_origin(level, slot), g.ipitch, g.w, g.h, cudaMemcpyDeviceToHost));
    } else if (plane == 1) {
        if (dst_bytes < (size_t)g.w * g.h * 4) { set_error("rdfe_download_level: buffer too small"); return RDFE_ERR_INVALID; }
        RDFE_CUDA_OK(cudaMemcpy2D(dst, (size_t)g.w * 4, ctx->pyr.deriv_origin(level, slot), g.dpitch, (size_t)g.w * 4, g.h, cudaMemcpyDeviceToHost));
    } else if (plane == 2) {
        const size_t fw = (size_t)g.w + 2 * win, fh = (size_t)g.h + 2 * win;
        if (dst_bytes < fw * fh) { set_error("rdfe_download_level: buffer too small"); return RDFE_ERR_INVALID; }
        const uint8_t *src = ctx->pyr.image_origin(level, slot) - (size_t)win * g.ipitch - win;
        RDFE_CUDA_OK(cudaMemcpy2D(dst, fw, src, g.ipitch, fw, fh, cudaMemcpyDeviceToHost));
    } else if (plane == 3 && level == 0 && (ctx->und_on || ctx->in_channels > 1)) {
        if (dst_bytes < (size_t)g.w * g.h) { set_error("rdfe_download_level: buffer too small"); return RDFE_ERR_INVALID; }
        RDFE_CUDA_OK(cudaMemcpy2D(dst, g.w, ctx->und_plane + (size_t)slot * ctx->gray_slot, ctx->gray_pitch, g.w, g.h, cudaMemcpyDeviceToHost));
    } else {
        set_error("rdfe_download_level: plane %d unknown", plane);
        return RDFE_ERR_INVALID;
    }
    return RDFE_OK;
}

int rdfe_upload_level0(rdfe_ctx *ctx, int slot, const uint8_t *image_with_halo, size_t src_bytes) {
    if (!ctx || !image_with_halo || slot < 0 || slot >= ctx->cfg.num_slots || !ctx->slot_used[slot]) {
        set_error("rdfe_upload_level0: bad slot or null image");
        return RDFE_ERR_INVALID;
    }
    const LevelGeom &g = ctx->pyr.lv[0];
    const int win = ctx->pyr.win;
    const size_t fw = (size_t)g.w + 2 * win, fh = (size_t)g.h + 2 * win;
    if (src_bytes < fw * fh) { set_error("rdfe_upload_level0: buffer too small"); return RDFE_ERR_INVALID; }
    RDFE_CUDA_OK(cudaSetDevice(ctx->cfg.device));
    { const int rc_all = sync_all_streams(ctx); if (rc_all) return rc_all; }
    ctx->pf_valid = false;
    ctx->slot_gen[slot] = ++ctx->gen_counter;           // cached LK templates of the old pixels are stale
    uint8_t *dst = ctx->pyr.image_origin(0, slot) - (size_t)win * g.ipitch - win;
    RDFE_CUDA_OK(cudaMemcpy2D(dst, g.ipitch, image_with_halo, fw, fw, fh, cudaMemcpyHostToDevice));
    return RDFE_OK;
}

int rdfe_download_clahe_lut(rdfe_ctx *ctx, int batch_index, uint8_t *dst, size_t dst_bytes) {
    if (!ctx || !dst || batch_index < 0 || batch_index >= RDFE_MAX_BATCH || ctx->last_clahe_tiles <= 0) {
        set_error("rdfe_download_clahe_lut: bad argument or no preprocess yet");
        return RDFE_ERR_INVALID;
    }
    const size_t bytes = (size_t)ctx->last_clahe_tiles * 256;
    if (dst_bytes < bytes) { set_error("rdfe_download_clahe_lut: buffer too small"); return RDFE_ERR_INVALID; }
    RDFE_CUDA_OK(cudaSetDevice(ctx->cfg.device));
    RDFE_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    RDFE_CUDA_OK(cudaMemcpy(dst, ctx->lut + (size_t)batch_index * bytes, bytes, cudaMemcpyDeviceToHost));
    return RDFE_OK;
}

int rdfe_harris_response(rdfe_ctx *ctx, int slot, const rdfe_detect_params *p, float *dst, size_t dst_bytes) {
    SlotList sl;
    int rc = check_slots(ctx, &slot, 1, &sl, "rdfe_harris_response");
    if (rc) return rc;
    const size_t bytes = (size_t)ctx->cfg.width * ctx->cfg.height * sizeof(float);
    if (!p || !dst || dst_bytes < bytes) { set_error("rdfe_harris_response: bad argument"); return RDFE_ERR_INVALID; }
    RDFE_CUDA_OK(cudaSetDevice(ctx->cfg.device));
    float *d = nullptr;
    RDFE_CUDA_OK(cudaMalloc(&d, bytes));
    rc = check_launch(ctx, launch_harris_candidates(ctx, sl, *p, d), "harris");
    if (rc == RDFE_OK) {
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpy(dst, d, bytes, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { set_error("rdfe_harris_response: %s", cudaGetErrorString(e)); rc = RDFE_ERR_CUDA; }
    }
    cudaFree(d);
    return rc;
}

int rdfe_harris_candidates(rdfe_ctx *ctx, int slot, const rdfe_detect_params *p, uint64_t *keys, size_t cap, unsigned *count,
                           float *frame_max, unsigned *flagged) {
    SlotList sl;
    int rc = check_slots(ctx, &slot, 1, &sl, "rdfe_harris_candidates");
    if (rc) return rc;
    if (!p || !keys || !count || !frame_max) { set_error("rdfe_harris_candidates: null argument"); return RDFE_ERR_INVALID; }
    RDFE_CUDA_OK(cudaSetDevice(ctx->cfg.device));
    rc = check_launch(ctx, launch_harris_candidates(ctx, sl, *p, nullptr), "harris");
    if (rc) return rc;
    RDFE_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    unsigned n = 0, fm = 0, nf = 0;
    RDFE_CUDA_OK(cudaMemcpy(&n, ctx->det.cand_count, sizeof n, cudaMemcpyDeviceToHost));
    RDFE_CUDA_OK(cudaMemcpy(&fm, ctx->det.frame_max, sizeof fm, cudaMemcpyDeviceToHost));
    RDFE_CUDA_OK(cudaMemcpy(&nf, ctx->det.flag_count, sizeof nf, cudaMemcpyDeviceToHost));
    if (n > ctx->det.cand_cap || n > cap) { set_error("rdfe_harris_candidates: %u candidates exceed the capacity", n); return RDFE_ERR_OVERFLOW; }
    RDFE_CUDA_OK(cudaMemcpy(keys, ctx->det.cand, (size_t)n * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    *count = n;
    memcpy(frame_max, &fm, sizeof fm);
    if (flagged) *flagged = nf;
    return RDFE_OK;
}

void rdfe_harris_prefilter_constants(float *out) {
    if (!out) return;
    out[0] = rdfe::kHarrisC1; out[1] = rdfe::kHarrisC2; out[2] = rdfe::kHarrisRhoU; out[3] = rdfe::kHarrisRhoS;
}

// ---------------------------------------------------------- memory helpers
int rdfe_dev_alloc(rdfe_ctx *ctx, size_t bytes, void **dev_ptr) {
    if (!ctx || !dev_ptr) return RDFE_ERR_INVALID;
    RDFE_CUDA_OK(cudaSetDevice(ctx->cfg.device));
    RDFE_CUDA_OK(cudaMalloc(dev_ptr, bytes));
    return RDFE_OK;
}
int rdfe_dev_free(rdfe_ctx *ctx, void *dev_ptr) {
    if (!ctx) return RDFE_ERR_INVALID;
    RDFE_CUDA_OK(cudaSetDevice(ctx->cfg.device));
    RDFE_CUDA_OK(cudaFree(dev_ptr));
    return RDFE_OK;
}
int rdfe_host_alloc(rdfe_ctx *ctx, size_t bytes, void **host_ptr) {
    if (!ctx || !host_ptr) return RDFE_ERR_INVALID;
    RDFE_CUDA_OK(cudaSetDevice(ctx->cfg.device));
    RDFE_CUDA_OK(cudaMallocHost(host_ptr, bytes));
    return RDFE_OK;
}
int rdfe_host_free(rdfe_ctx *ctx, void *host_ptr) {
    if (!ctx) return RDFE_ERR_INVALID;
    RDFE_CUDA_OK(cudaFreeHost(host_ptr));
    return RDFE_OK;
}
int rdfe_memcpy_h2d(rdfe_ctx *ctx, void *dev_dst, const void *host_src, size_t bytes, int async) {
    if (!ctx) return RDFE_ERR_INVALID;
    RDFE_CUDA_OK(cudaSetDevice(ctx->cfg.device));
    RDFE_CUDA_OK(cudaMemcpyAsync(dev_dst, host_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (!async) RDFE_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    return RDFE_OK;
}
int rdfe_memcpy_d2h(rdfe_ctx *ctx, void *host_dst, const void *dev_src, size_t bytes, int async) {
    if (!ctx) return RDFE_ERR_INVALID;
    RDFE_CUDA_OK(cudaSetDevice(ctx->cfg.device));
    RDFE_CUDA_OK(cudaMemcpyAsync(host_dst, dev_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (!async) RDFE_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    return RDFE_OK;
}

static const char *const kKernelNames[K_COUNT] = {"clahe_hist_lut", "clahe_apply", "pyrdown", "scharr",
                                                  "harris_nms", "select", "lk_track", "poisson_append", "undistort",
                                                  "harris_resolve", "predict_rotation"};

int rdfe_profile_num_kernels(void) { return K_COUNT; }
const char *rdfe_profile_kernel_name(int id) { return (id >= 0 && id < K_COUNT) ? kKernelNames[id] : ""; }

int rdfe_profile_enable(rdfe_ctx *ctx, int on) {
    if (!ctx) return RDFE_ERR_INVALID;
    RDFE_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    ctx->prof_on = on != 0;
    ctx->prof_timeline = on == 2;
    ctx->prof_used = 0;
    for (int k = 0; k < K_COUNT; ++k) { ctx->prof_ms[k] = 0.0; ctx->prof_n[k] = 0; }
    return RDFE_OK;
}

int rdfe_profile_collect(rdfe_ctx *ctx, double *ms, int64_t *launches) {
    if (!ctx || !ms || !launches) return RDFE_ERR_INVALID;
    RDFE_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < ctx->prof_used; ++i) {
        float t = 0.f;
        RDFE_CUDA_OK(cudaEventElapsedTime(&t, ctx->prof_ev[2 * i], ctx->prof_ev[2 * i + 1]));
        ctx->prof_ms[ctx->prof_kid[i]] += t;
        ctx->prof_n[ctx->prof_kid[i]] += 1;
    }
    ctx->prof_used = 0;
    for (int k = 0; k < K_COUNT; ++k) { ms[k] = ctx->prof_ms[k]; launches[k] = ctx->prof_n[k]; }
    return RDFE_OK;
}

int rdfe_profile_timeline(rdfe_ctx *ctx, int *kernel_ids, float *start_ms, float *end_ms, int cap, int *count) {
    if (!ctx || !kernel_ids || !start_ms || !end_ms || !count) return RDFE_ERR_INVALID;
    RDFE_CUDA_OK(cudaDeviceSynchronize());
    const int n = ctx->prof_used < cap ? ctx->prof_used : cap;
    for (int i = 0; i < n; ++i) {
        kernel_ids[i] = ctx->prof_kid[i];
        RDFE_CUDA_OK(cudaEventElapsedTime(&start_ms[i], ctx->prof_ev[0], ctx->prof_ev[2 * i]));
        RDFE_CUDA_OK(cudaEventElapsedTime(&end_ms[i], ctx->prof_ev[0], ctx->prof_ev[2 * i + 1]));
    }
    *count = n;
    ctx->prof_used = 0;
    return RDFE_OK;
}

int rdfe_timer_start(rdfe_ctx *ctx) {
    if (!ctx) return RDFE_ERR_INVALID;
    RDFE_CUDA_OK(cudaEventRecord(ctx->ev_t0, ctx->stream));
    return RDFE_OK;
}
int rdfe_timer_stop(rdfe_ctx *ctx) {
    if (!ctx) return RDFE_ERR_INVALID;
    RDFE_CUDA_OK(cudaEventRecord(ctx->ev_t1, ctx->stream));
    return RDFE_OK;
}
int rdfe_timer_elapsed_ms(rdfe_ctx *ctx, float *ms) {
    if (!ctx || !ms) return RDFE_ERR_INVALID;
    RDFE_CUDA_OK(cudaEventSynchronize(ctx->ev_t1));
    RDFE_CUDA_OK(cudaEventElapsedTime(ms, ctx->ev_t0, ctx->ev_t1));
    return RDFE_OK;
}

}  // extern "C"
