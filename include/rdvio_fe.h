/*
 * rdvio_fe.h -- C ABI of the B200-native rd_vio visual front-end (librdvio_fe.so).
 *
 * This is the drop-in boundary for ONE hot path of SummerSigh/rd_vio: the
 * rdvio::Image plugin (reference: src/rdvio/include/rdvio/types.h:153-177) as
 * implemented by rdvio::extra::OpenCvImage (src/rdvio_extra/src/opencv_image.cpp)
 * and driven once per camera frame by FeatureTracker::run
 * (src/rdvio/src/feature_tracker.cpp:26-111).  The reference has no FFI of its
 * own (it calls OpenCV in-process); each entry point below states which
 * reference member function it replaces.  include/rdvio_b200/gpu_image.hpp is
 * the C++ Image subclass that forwards to these calls; INTEGRATION.md shows
 * the one-line change at the construction site (rdvio.hpp:50-53).
 *
 * Conventions
 *  - plain C types, caller-owned buffers, no exceptions across the ABI;
 *  - every function returns 0 (RDFE_OK) on success, a negative rdfe_status
 *    otherwise; rdfe_last_error() gives a thread-local message;
 *  - "slot" = one image resident in HBM with its pyramid (what one OpenCvImage
 *    instance holds between preprocess() and release_image_buffer());
 *  - batched: every call takes n independent images (camera streams share no
 *    state); n = 1 is the per-frame plugin case;
 *  - keypoints are (x, y) pairs of doubles, the layout of
 *    std::vector<Eigen::Matrix<double,2,1>> (types.h:55-59);
 *  - *_dev variants take DEVICE pointers and only enqueue work on the context's
 *    stream (no host synchronisation); the host variants copy in/out and
 *    return when results are in the caller's buffers.
 */
#ifndef RDVIO_FE_H
#define RDVIO_FE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define RDFE_API __attribute__((visibility("default")))
#else
#define RDFE_API
#endif

#define RDFE_ABI_VERSION 1
#define RDFE_MAX_LEVELS 8        /* pyramid images (maxLevel + 1) */
#define RDFE_MAX_BATCH 128       /* images per batched call */

typedef enum rdfe_status {
    RDFE_OK = 0,
    RDFE_ERR_INVALID = -1,       /* bad argument */
    RDFE_ERR_CUDA = -2,          /* CUDA runtime/driver failure */
    RDFE_ERR_NOMEM = -3,
    RDFE_ERR_NOSLOT = -4,        /* slot pool exhausted */
    RDFE_ERR_UNSUPPORTED = -5,   /* e.g. LK window other than 21 or 31 */
    RDFE_ERR_OVERFLOW = -6       /* corner-candidate buffer overflow, or keypoints beyond `stride` were dropped */
} rdfe_status;

typedef struct rdfe_ctx rdfe_ctx;

typedef struct rdfe_config {
    int device;        /* CUDA device ordinal */
    int width;         /* level-0 image size (OpenCvImage::width()/height(), opencv_image.h:15-17) */
    int height;
    int max_level;     /* OpenCV maxLevel = Image::level_num() (3, opencv_image.h:19) => max_level+1 images */
    int win;           /* LK window, Size(21,21) in the reference (opencv_image.cpp:96,159); 21 or 31 */
    int num_slots;     /* images resident at once (>= 2 per stream: previous + new frame) */
    int max_points;    /* capacity: keypoints per image per detect/track call */
    void *stream;      /* optional cudaStream_t to run on (NULL: the context creates its own) */
} rdfe_config;

/* Harris/GFTT parameters of OpenCvImage::gftt (opencv_image.cpp:184-188):
 * GFTTDetector::create(max_points, 1e-3, 20, 3, useHarrisDetector=true), k = 0.04. */
typedef struct rdfe_detect_params {
    int max_points;             /* maxCorners (feature_tracker_max_keypoint_detection, config.cpp:25) */
    double quality_level;       /* 1e-3 */
    double min_distance;        /* 20 (GFTT's own suppression radius) */
    double harris_k;            /* 0.04 */
    double keypoint_distance;   /* PoissonDiskFilter radius (feature_tracker_min_keypoint_distance, config.cpp:23) */
    int border;                 /* 20-px border reject (opencv_image.cpp:61-68) */
    int harris_fma;             /* 0: plain C++ float order (parity target); 1: FMA order of OpenCV's AVX2 dispatch */
} rdfe_detect_params;

/* LK parameters of OpenCvImage::track_keypoints (opencv_image.cpp:94-98,101-113,130). */
typedef struct rdfe_track_params {
    int max_count;              /* TermCriteria COUNT, 30 */
    double epsilon;             /* TermCriteria EPS, 0.01 */
    double min_eig_threshold;   /* calcOpticalFlowPyrLK default 1e-4 */
    int border;                 /* 20-px border gate on the forward result */
    double max_round_trip;      /* 0.5 px forward-backward gate */
    int has_prediction;         /* 1: next_xy holds the caller's prediction (OPTFLOW_USE_INITIAL_FLOW seed);
                                   0: seed with curr_xy (opencv_image.cpp:81-86) */
} rdfe_track_params;

RDFE_API const char *rdfe_last_error(void);
RDFE_API int rdfe_abi_version(void);
RDFE_API void rdfe_default_detect_params(rdfe_detect_params *p);
RDFE_API void rdfe_default_track_params(rdfe_track_params *p);

/* ---- context ------------------------------------------------------------ */
RDFE_API int rdfe_create(const rdfe_config *cfg, rdfe_ctx **out);
RDFE_API void rdfe_destroy(rdfe_ctx *ctx);
RDFE_API int rdfe_num_levels(const rdfe_ctx *ctx);                 /* images actually built (buildOpticalFlowPyramid may stop early) */
RDFE_API int rdfe_level_size(const rdfe_ctx *ctx, int level, int *width, int *height);
RDFE_API int rdfe_sync(rdfe_ctx *ctx);                             /* wait for everything enqueued on the context */
RDFE_API void *rdfe_stream(rdfe_ctx *ctx);                         /* the cudaStream_t work is enqueued on */
RDFE_API int64_t rdfe_kernel_launches(const rdfe_ctx *ctx);        /* kernels launched by this context so far */

/* ---- slots: replace the cv::Mat members' lifetime ----------------------- */
RDFE_API int rdfe_slot_acquire(rdfe_ctx *ctx, int *slot);          /* make_shared<OpenCvImage>() (rdvio.hpp:50) */
RDFE_API int rdfe_slot_release(rdfe_ctx *ctx, int slot);           /* OpenCvImage::release_image_buffer (opencv_image.cpp:200-208) */

/* ---- OpenCvImage::preprocess (opencv_image.cpp:156-161) ----------------- */
/* CLAHE(clip_limit, tiles_x x tiles_y) in place, then the (max_level+1)-image
 * pyramid with Scharr derivatives.  images[i] is the 8-bit gray frame for
 * slots[i], `pitch` bytes per row. */
RDFE_API int rdfe_preprocess_batch(rdfe_ctx *ctx, const int *slots, int n,
                          const uint8_t *const *images, size_t pitch,
                          double clip_limit, int tiles_x, int tiles_y);
RDFE_API int rdfe_preprocess_batch_dev(rdfe_ctx *ctx, const int *slots, int n,
                              const uint8_t *const *dev_images, size_t pitch,
                              double clip_limit, int tiles_x, int tiles_y);

/* ---- OpenCvImage::detect_keypoints (opencv_image.cpp:38-73) ------------- */
/* keypoints_xy: [n][stride][2] doubles; counts[i] existing keypoints on entry
 * (they preset the Poisson-disk filter), new ones are appended and counts[i]
 * updated.  The reference's vector is unbounded: if existing + new exceed `stride` the
 * list is cut there and the next synchronising call returns RDFE_ERR_OVERFLOW.  gftt_* (optional, may be NULL) receive the
 * raw GFTT corners [n][max_points][2] float, responses, and counts. */
RDFE_API int rdfe_detect_batch(rdfe_ctx *ctx, const int *slots, int n, const rdfe_detect_params *p,
                      double *keypoints_xy, int *counts, int stride,
                      float *gftt_xy, float *gftt_resp, int *gftt_counts);
RDFE_API int rdfe_detect_batch_dev(rdfe_ctx *ctx, const int *slots, int n, const rdfe_detect_params *p,
                          double *dev_keypoints_xy, int *dev_counts, int stride,
                          float *dev_gftt_xy, float *dev_gftt_resp, int *dev_gftt_counts);

/* ---- OpenCvImage::track_keypoints (opencv_image.cpp:75-154) ------------- */
/* Optional: start Harris + GFTT selection for these (already preprocessed) slots on an internal stream.  They need
 * only level 0 of the image and none of the tracked points, so they can run while rdfe_track_batch works.  A later
 * rdfe_detect_batch[_dev] on exactly these slots with the same GFTT parameters (max_points, quality_level,
 * min_distance, harris_k, harris_fma) and no GFTT taps then only runs the Poisson append; any other call pattern
 * simply ignores the prefetch.  Results are identical with or without it. */
RDFE_API int rdfe_detect_prefetch(rdfe_ctx *ctx, const int *slots, int n, const rdfe_detect_params *p);

/* Host-pointer rdfe_preprocess_batch: on (default) = returns when the pyramid is complete; off = returns as soon as
 * the uploads are issued (pageable memory: already staged; pinned memory: the caller keeps the frames alive until
 * the next synchronising call).  Later calls are ordered behind it on the context's stream either way. */
RDFE_API int rdfe_set_host_sync(rdfe_ctx *ctx, int on);

/* Forward pyramidal LK curr->next, border/jump gating, backward LK, 0.5-px
 * round-trip gate, all in one launch.  curr_xy, next_xy: [n][stride][2]
 * doubles; counts[i] points for image i; status: [n][stride] chars in {0,1}.
 * next_xy is in/out: prediction in (if has_prediction), result out, and only
 * entries with status != 0 are overwritten (opencv_image.cpp:148-153). */
RDFE_API int rdfe_track_batch(rdfe_ctx *ctx, const int *curr_slots, const int *next_slots, int n,
                     const rdfe_track_params *p, const double *curr_xy, double *next_xy,
                     const int *counts, int stride, char *status);
RDFE_API int rdfe_track_batch_dev(rdfe_ctx *ctx, const int *curr_slots, const int *next_slots, int n,
                         const rdfe_track_params *p, const double *dev_curr_xy, double *dev_next_xy,
                         const int *dev_counts, int stride, char *dev_status);

/* ---- one call per new frame: FeatureTracker::run's plugin sequence (feature_tracker.cpp:32-98) ----
 * preprocess(new) -> track_keypoints(prev -> new) -> detect_keypoints(new), device pointers, asynchronous.
 * dev_next_xy [n][stride][2]: prediction in (if tp->has_prediction), tracked result out (status != 0 only).
 * Then, as Frame::track_keypoints does before the new frame sees detect (frame.cpp:160-170: only status != 0
 * points are appended to the next frame), the list is compacted in order to the tracked points
 * (rdfe_set_step_compaction, default on; dev_status keeps the original indexing: entry i of the input is entry
 * rank(i) of the output), and those are detect's existing keypoints: new corners are appended and
 * dev_kp_counts[i] = tracked + new (on entry: the number of carried points = dev_track_counts[i]).  The two
 * host-side RANSAC masks of frame.cpp:99-132 sit between track and detect in the reference; a caller that wants
 * them uses the separate calls (or ANDs its mask into dev_status on the context's stream before this call's
 * Poisson stage -- not offered here).  prev_slots == NULL (first frame of a stream) skips the tracking stage.
 * After CLAHE the detection branch (Harris + GFTT selection, which need only level 0) runs on an internal second
 * stream concurrently with the tracking branch (pyramid, Scharr, LK); only the Poisson-disk append joins both. */
RDFE_API int rdfe_frontend_step_dev(rdfe_ctx *ctx, const int *prev_slots, const int *new_slots, int n,
                           const uint8_t *const *dev_images, size_t pitch, double clip_limit, int tiles_x, int tiles_y,
                           const rdfe_track_params *tp, const double *dev_curr_xy, double *dev_next_xy,
                           const int *dev_track_counts, char *dev_status, const rdfe_detect_params *dp,
                           int *dev_kp_counts, int stride);

/* Caller-side helper for keypoints that stay on the device between steps: Frame::track_keypoints' prediction
 * (frame.cpp:82-93) next_i = apply_k(delta_q * remove_k(curr_i, K), K_next) is the homography H = K_next R K^-1 applied
 * to the pixel; dev_H holds one row-major 3x3 (doubles) per image.  dev_pred_xy may then be handed to
 * rdfe_track_batch_dev / rdfe_frontend_step_dev as the prediction (has_prediction = 1). */
RDFE_API int rdfe_predict_rotation_dev(rdfe_ctx *ctx, int n, const double *dev_H, const double *dev_curr_xy,
                              const int *dev_counts, int stride, double *dev_pred_xy);

/* "Next" row: the cv::undistort(img, out, K, D) the reference's dataset reader applies to
 * every frame before addFrame (examples/dataset.hpp:232-236, :591).  K: 3x3 row-major float32 camera matrix,
 * D: k1 k2 p1 p2 float32.  When set, every preprocess / frontend_step call first remaps its source frames with
 * OpenCV's exact fixed-point bilinear arithmetic (bit-identical to cv::undistort); NULL, NULL switches it off.
 * One calibration per context (all streams of a context share the camera model). */
RDFE_API int rdfe_set_undistort(rdfe_ctx *ctx, const float *K, const float *D);

/* "Next" row: Odometry::addFrame's cv::cvtColor(BGR2GRAY / BGRA2GRAY) (rdvio.hpp:42-49).
 * channels = 1 (default, gray), 3 (BGR) or 4 (BGRA): source frames of all later preprocess / frontend_step calls
 * are `width * channels` bytes per row; colour frames are converted with OpenCV's 15-bit fixed-point weights
 * (bit-identical to cv::cvtColor), after the optional undistortion (which then runs per channel). */
RDFE_API int rdfe_set_input_format(rdfe_ctx *ctx, int channels);

/* Cross-step pipelining for rdfe_frontend_step_dev (off by default).  When on, every kernel class of a step runs on
 * its own internal stream (preprocess, Harris, selection, LK, Poisson append) and events carry only the data
 * dependences, so the preprocess stage of step s overlaps step s-1: it waits for LK of step s-2, the last reader of
 * the slot set it rewrites -- provided its new slots were not touched by the previous step (use three slot sets in
 * rotation; otherwise the call silently serialises).  LK waits for the work already enqueued on the context's stream
 * (the caller's keypoint buffers) and the context's stream waits for the end of the step, so results are ordered for
 * the caller as usual.  Requirement: the source images must already be complete in device memory when the call is
 * made (they are not ordered against the context stream).  The preprocess stream runs at a higher priority than the
 * others (environment variable RDFE_PRIO="pre,harris,select,lk,poisson" overrides the five levels); a caller stream
 * created with a high priority keeps its own small copies from queueing behind the large grids. */
RDFE_API int rdfe_set_pipelining(rdfe_ctx *ctx, int on);

/* LK template cache (off by default; costs num_slots * max_points * ~3 KB * levels of HBM).  The backward pass of a
 * track call builds, at the tracked position q in the NEXT image, exactly the per-level template (I*32, Ix, Iy over the
 * window, A11/A12/A22) that the forward pass of the following track call needs when that point is carried on
 * unchanged (FeatureTracker::run: next frame's curr keypoints = this frame's tracked ones).  With the cache on it is
 * kept, keyed by (image generation of the slot, float x, float y), and a forward pass that finds its point loads it
 * instead of staging the I / dI patches and rebuilding.  Results are bit-identical with and without (the template is
 * a pure function of image and position); any call that rewrites level 0 of a slot invalidates its records. */
RDFE_API int rdfe_set_template_cache(rdfe_ctx *ctx, int on);
/* forward passes that consulted the cache / found their point there, since creation or the last reset (waits for the context) */
RDFE_API int rdfe_template_cache_stats(rdfe_ctx *ctx, unsigned long long *lookups, unsigned long long *hits, int reset);

/* Fused step only: on (default) = lost tracks (status == 0) are dropped before detect, the reference's behaviour
 * (frame.cpp:160-170); off = every carried entry (prediction where tracking failed) stays in the list and keeps
 * suppressing new corners -- the round-1 behaviour, kept for A/B measurements. */
RDFE_API int rdfe_set_step_compaction(rdfe_ctx *ctx, int on);

/* Host-buffer form of the same step, pipelined up to three deep: submit() uploads the frames and keypoints on a
 * copy stream and enqueues the step; wait() blocks until that step's results are on the host and copies them out.
 * submit(t+1) and submit(t+2) may precede wait(t) so that the uploads of the next frames overlap the kernels of the
 * current step and the copy engine never idles while the host collects results (a fourth un-waited submit is refused).  next_xy [n][stride][2]: positions of the carried keypoints (prediction where tracking failed) followed
 * by the newly detected corners (with step compaction, the default: only the tracked ones, in order, followed by
 * the new corners), kp_counts[i] entries; status [n][stride] in the input's indexing.  pred_xy may be NULL (seed with
 * curr_xy); prev_slots may be NULL (no tracking, counts[i] existing keypoints are taken from curr_xy = NULL -> 0). */
RDFE_API int rdfe_frontend_step_submit(rdfe_ctx *ctx, const int *prev_slots, const int *new_slots, int n,
                              const uint8_t *const *images, size_t pitch, double clip_limit, int tiles_x, int tiles_y,
                              const rdfe_track_params *tp, const double *curr_xy, const double *pred_xy, const int *counts,
                              const rdfe_detect_params *dp, int stride, int *ticket);
RDFE_API int rdfe_frontend_step_wait(rdfe_ctx *ctx, int ticket, double *next_xy, int *kp_counts, char *status);
/* Measurement aid: only the frame upload of rdfe_frontend_step_submit (same staging, copy stream and copy shape),
 * no kernels; sync != 0 waits for the copy stream.  bench.py times it alone ("h2d_only") next to the e2e number. */
RDFE_API int rdfe_upload_only(rdfe_ctx *ctx, const int *new_slots, int n, const uint8_t *const *images, size_t pitch, int sync);

/* ---- parity / debugging taps (not on the hot path) ---------------------- */
/* plane: 0 = 8-bit image (w*h bytes), 1 = Scharr derivative (w*h*2 int16),
 * 2 = image with its win-px REFLECT_101 halo ((w+2win)*(h+2win) bytes),
 * 3 = (level 0 only) the undistorted frame before CLAHE, if rdfe_set_undistort is active (w*h bytes). */
RDFE_API int rdfe_download_level(rdfe_ctx *ctx, int slot, int level, int plane, void *dst, size_t dst_bytes);
RDFE_API int rdfe_download_clahe_lut(rdfe_ctx *ctx, int batch_index, uint8_t *dst, size_t dst_bytes);
/* Inverse of plane 2: overwrite level 0 of a slot (the CLAHE output) with a caller-made image INCLUDING its win-px
 * REFLECT_101 halo ((w+2win)*(h+2win) bytes), so that the detection stage can be tested on arbitrary level-0 content. */
RDFE_API int rdfe_upload_level0(rdfe_ctx *ctx, int slot, const uint8_t *image_with_halo, size_t src_bytes);
/* Harris response map of the slot's level-0 image (w*h floats). */
RDFE_API int rdfe_harris_response(rdfe_ctx *ctx, int slot, const rdfe_detect_params *p, float *dst, size_t dst_bytes);
/* What the Harris stage of the hot path hands to the selection: 64-bit keys (float bits of the response << 32 | y*w+x)
 * of every positive 3x3 local maximum off the 1-px frame (goodFeaturesToTrack's candidates before its threshold), in
 * no particular order, the frame maximum of the response, and the number of pixels the integer prefilter flagged for
 * exact evaluation (0 when the exact-everywhere kernel ran: environment RDFE_HARRIS_EXACT=1). */
RDFE_API int rdfe_harris_candidates(rdfe_ctx *ctx, int slot, const rdfe_detect_params *p, uint64_t *keys, size_t cap,
                           unsigned *count, float *frame_max, unsigned *flagged);
/* Constants of the prefilter's error bound eps(T) = T * (c1 * sqrt(T) + c2 * T) + rho_u and the degenerate-frame
 * threshold rho_s (csrc/harris_exact.cuh): out[4] = {c1, c2, rho_u, rho_s}.  Needs no GPU. */
RDFE_API void rdfe_harris_prefilter_constants(float *out);

/* ---- device-buffer helpers for callers without a CUDA runtime binding --- */
RDFE_API int rdfe_dev_alloc(rdfe_ctx *ctx, size_t bytes, void **dev_ptr);
RDFE_API int rdfe_dev_free(rdfe_ctx *ctx, void *dev_ptr);
RDFE_API int rdfe_host_alloc(rdfe_ctx *ctx, size_t bytes, void **host_ptr);     /* pinned */
RDFE_API int rdfe_host_free(rdfe_ctx *ctx, void *host_ptr);
RDFE_API int rdfe_memcpy_h2d(rdfe_ctx *ctx, void *dev_dst, const void *host_src, size_t bytes, int async);
RDFE_API int rdfe_memcpy_d2h(rdfe_ctx *ctx, void *host_dst, const void *dev_src, size_t bytes, int async);

/* CUDA-event timing on the context's stream: start/stop bracket enqueued work,
 * elapsed_ms synchronises on the stop event. */
RDFE_API int rdfe_timer_start(rdfe_ctx *ctx);
RDFE_API int rdfe_timer_stop(rdfe_ctx *ctx);
RDFE_API int rdfe_timer_elapsed_ms(rdfe_ctx *ctx, float *ms);

/* Per-kernel CUDA-event timing (off by default).  While enabled every kernel launch is
 * bracketed by two events on the context's stream; collect() synchronises and returns the
 * accumulated device milliseconds and launch counts per kernel id since enable(). */
RDFE_API int rdfe_profile_num_kernels(void);
RDFE_API const char *rdfe_profile_kernel_name(int id);
RDFE_API int rdfe_profile_enable(rdfe_ctx *ctx, int on);
RDFE_API int rdfe_profile_collect(rdfe_ctx *ctx, double *ms, int64_t *launches);
/* rdfe_profile_enable(ctx, 2) keeps the streams overlapped; this returns, per launch since then, the kernel id and the
 * stream-side start/end times in ms relative to the first launch (a timeline of the pipelined schedule). */
RDFE_API int rdfe_profile_timeline(rdfe_ctx *ctx, int *kernel_ids, float *start_ms, float *end_ms, int cap, int *count);

#ifdef __cplusplus
}
#endif
#endif /* RDVIO_FE_H */
