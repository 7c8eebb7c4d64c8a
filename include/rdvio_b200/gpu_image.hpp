// gpu_image.hpp -- drop-in rdvio::Image implementation backed by librdvio_fe.so (B200, sm_100a).
//
// Replaces rdvio::extra::OpenCvImage (reference: src/rdvio_extra/include/rdvio/extra/opencv_image.h:9-56,
// src/rdvio_extra/src/opencv_image.cpp) behind the unchanged plugin interface rdvio::Image
// (src/rdvio/include/rdvio/types.h:153-177).  Every virtual keeps the reference's exact signature
// (note detect/track are const).  Header-only; needs the reference's <rdvio/types.h> (Eigen
// vector<2>, Image) and cv::Mat for the two public members the construction site assigns
// (rdvio.hpp:50-53: image_ptr->image = gray.clone(); image_ptr->raw = image.clone(); image_ptr->t = t).
//
// Semantics kept from the reference:
//  * CLAHE and GFTT parameters are frozen by the FIRST call in the process (function-local statics,
//    opencv_image.cpp:179-188);
//  * a next_image of another dynamic type, or zero keypoints, yields all-zero status and no throw
//    (opencv_image.cpp:88-92); result_status.resize(n, 0) does not clear pre-existing entries (:91);
//  * next_keypoints is only overwritten where status != 0 (:148-153);
//  * level_num() == 3 is passed as OpenCV maxLevel => 4 pyramid images (:96,159-160);
//  * evaluate() is dead code in the reference (its interpolators are never built) -- returns 0 here.
// Failure of the CUDA library never throws across the plugin boundary: it degrades to all-zero
// status / no new keypoints exactly like the reference's failed dynamic_cast, and the message is kept
// in last_error().
#pragma once

#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <map>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include <rdvio/types.h>

#include "../rdvio_fe.h"

namespace rdvio::extra {

// One CUDA context per image size, shared by all GpuImage instances of the process.  All plugin calls
// of one Odometry come from one thread (feature_tracker.cpp:26-118); the mutex only guards creation.
class GpuFrontEnd {
  public:
    static rdfe_ctx *get(int width, int height) {
        static std::mutex m;
        static std::map<std::pair<int, int>, rdfe_ctx *> ctxs;
        std::lock_guard<std::mutex> lk(m);
        auto key = std::make_pair(width, height);
        auto it = ctxs.find(key);
        if (it != ctxs.end()) return it->second;
        rdfe_config cfg{};
        cfg.device = device();
        cfg.width = width;
        cfg.height = height;
        cfg.max_level = 3;      // OpenCvImage::level_num() (opencv_image.h:19)
        cfg.win = 21;           // Size(21, 21) (opencv_image.cpp:96,159)
        cfg.num_slots = 8;      // previous + new frame (+ clones still held by the map)
        cfg.max_points = 4096;
        cfg.stream = nullptr;
        rdfe_ctx *ctx = nullptr;
        if (rdfe_create(&cfg, &ctx) != RDFE_OK) {
            std::fprintf(stderr, "rdvio GpuImage: %s\n", rdfe_last_error());
            ctx = nullptr;
        } else {
            // preprocess() returns once the frame is uploaded: the image Mat outlives the copy (it is held until
            // release_image_buffer()), and track/detect are ordered behind the pyramid on the context's stream
            rdfe_set_host_sync(ctx, 0);
            // The LK template cache (rdfe_set_template_cache) would apply here -- Frame::track_keypoints re-projects the points
            // it carried on (apply_k(remove_k(p)), frame.cpp:76-79), which returns the same float pixel -- but it stays off:
            // the kernel variant that carries the cache paths measured slower than the plain one.
        }
        ctxs[key] = ctx;
        return ctx;
    }
    static int &device() {
        static int dev = 0;
        return dev;
    }
    // OpenCvImage::gftt is a function-local static created by the first detect_keypoints call of the process
    // (opencv_image.cpp:184-188): its max_points is frozen there.  0 = not called yet.
    static size_t &frozen_max_points() {
        static size_t v = 0;
        return v;
    }
    // FeatureTracker::run calls detect_keypoints only on sliding-window frames (feature_tracker.cpp:39-41,97: every
    // frame until initialised, then every sliding_window_tracker_frequent-th).  The detection prefetch of preprocess()
    // follows that rhythm: frames_since_detect() counts preprocess() calls since the last detect_keypoints(),
    // detect_gap() is the count seen at that detect; a frame is prefetched only when it is expected to be detected.
    static int &frames_since_detect() {
        static int v = 0;
        return v;
    }
    static int &detect_gap() {
        static int v = 1;
        return v;
    }
    static rdfe_detect_params detect_params(size_t frozen) {
        rdfe_detect_params p;
        rdfe_default_detect_params(&p);
        p.max_points = (int)(frozen ? frozen : 4096);
        if (p.max_points > 2048) p.max_points = 2048;
        return p;
    }
};

class GpuImage : public Image {
  public:
    GpuImage() = default;
    ~GpuImage() override { release_slot(); }

    uchar *get_rawdata() const override { return raw.data; }
    size_t width() const override { return cols_ ? (size_t)cols_ : (size_t)image.cols; }
    size_t height() const override { return rows_ ? (size_t)rows_ : (size_t)image.rows; }
    size_t level_num() const override { return 3; }

    double evaluate(const vector<2> &, int = 0) const override { return 0.0; }
    double evaluate(const vector<2> &, vector<2> &, int = 0) const override { return 0.0; }

    void preprocess(double clipLimit, int width, int height) override {
        // static singleton semantics of OpenCvImage::clahe (opencv_image.cpp:179-182)
        static const double s_clip = clipLimit;
        static const int s_w = width, s_h = height;
        if (image.data == nullptr) return;
        cols_ = image.cols;
        rows_ = image.rows;
        ctx_ = GpuFrontEnd::get(cols_, rows_);
        if (!ctx_) return;
        if (slot_ < 0 && rdfe_slot_acquire(ctx_, &slot_) != RDFE_OK) { slot_ = -1; return; }
        const uint8_t *src = image.data;
        if (rdfe_preprocess_batch(ctx_, &slot_, 1, &src, (size_t)image.step, s_clip, s_w, s_h) != RDFE_OK) {
            std::fprintf(stderr, "rdvio GpuImage::preprocess: %s\n", rdfe_last_error());
            return;
        }
        // FeatureTracker calls track(prev -> this) and then detect(this) (feature_tracker.cpp:43-96).  Corner
        // selection needs neither the tracked points nor the pyramid levels above 0, so once the detector's
        // parameters are frozen it is started here and runs beside the tracking call.
        const bool detect_expected = ++GpuFrontEnd::frames_since_detect() == GpuFrontEnd::detect_gap();
        if (const size_t frozen = detect_expected ? GpuFrontEnd::frozen_max_points() : 0) {
            const rdfe_detect_params p = GpuFrontEnd::detect_params(frozen);
            rdfe_detect_prefetch(ctx_, &slot_, 1, &p);      // optional: a failure only means detect does the work itself
        }
    }

    void detect_keypoints(std::vector<vector<2>> &keypoints, size_t max_points = 1000,
                          double keypoint_distance = 10) const override {
        // static singleton semantics of OpenCvImage::gftt (opencv_image.cpp:184-188)
        size_t &frozen = GpuFrontEnd::frozen_max_points();
        if (frozen == 0) frozen = max_points ? max_points : 4096;
        if (GpuFrontEnd::frames_since_detect() > 0) GpuFrontEnd::detect_gap() = GpuFrontEnd::frames_since_detect();
        GpuFrontEnd::frames_since_detect() = 0;
        if (!ctx_ || slot_ < 0) return;
        rdfe_detect_params p = GpuFrontEnd::detect_params(frozen);
        p.keypoint_distance = keypoint_distance;
        int count = (int)keypoints.size();
        const int stride = count + p.max_points;
        if (stride > 4096) return;
        static_assert(sizeof(vector<2>) == 2 * sizeof(double), "vector<2> must be a packed (x, y) pair of doubles");
        keypoints.resize((size_t)stride);
        const int rc = rdfe_detect_batch(ctx_, &slot_, 1, &p, reinterpret_cast<double *>(keypoints.data()), &count, stride,
                                         nullptr, nullptr, nullptr);
        if (rc != RDFE_OK) {
            std::fprintf(stderr, "rdvio GpuImage::detect_keypoints: %s\n", rdfe_last_error());
            count = stride - p.max_points;
        }
        keypoints.resize((size_t)count);
    }

    void track_keypoints(const Image *next_image, const std::vector<vector<2>> &curr_keypoints,
                         std::vector<vector<2>> &next_keypoints, std::vector<char> &result_status) const override {
        const size_t n = curr_keypoints.size();
        const bool has_prediction = next_keypoints.size() > 0;
        if (!has_prediction) next_keypoints.resize(n);
        const GpuImage *next = dynamic_cast<const GpuImage *>(next_image);
        result_status.resize(n, 0);
        if (!next || n == 0 || !ctx_ || slot_ < 0 || next->slot_ < 0 || next->ctx_ != ctx_ || n > 4096) return;
        rdfe_track_params p;
        rdfe_default_track_params(&p);
        p.has_prediction = has_prediction ? 1 : 0;
        const int count = (int)n;
        const int rc = rdfe_track_batch(ctx_, &slot_, &next->slot_, 1, &p, reinterpret_cast<const double *>(curr_keypoints.data()),
                                        reinterpret_cast<double *>(next_keypoints.data()), &count, count, result_status.data());
        if (rc != RDFE_OK) {
            std::fprintf(stderr, "rdvio GpuImage::track_keypoints: %s\n", rdfe_last_error());
            for (auto &s : result_status) s = 0;
        }
    }

    void release_image_buffer() override {
        image.release();
        raw.release();
        release_slot();
    }

    cv::Mat image;   // public members the construction site writes (rdvio.hpp:51-52)
    cv::Mat raw;

  private:
    void release_slot() {
        if (ctx_ && slot_ >= 0) rdfe_slot_release(ctx_, slot_);
        slot_ = -1;
    }
    rdfe_ctx *ctx_ = nullptr;
    int slot_ = -1;
    int cols_ = 0, rows_ = 0;
};

}  // namespace rdvio::extra
