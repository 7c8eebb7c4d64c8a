// plugin_replay.cpp -- drives rdvio::extra::GpuImage exactly like FeatureTracker::run drives the Image plugin
// (/root/reference/src/rdvio/src/feature_tracker.cpp:32-98): preprocess(new) -> track(prev->new) -> release(prev)
// -> detect(new), over a raw frame file, and dumps every frame's keypoints for comparison with the oracle replay.
//   usage: plugin_replay frames.bin out.txt      frames.bin = int32 n, H, W then n*H*W bytes
// Also reports the per-frame front-end latency (host image in -> keypoints in host vectors, SURVEY.md 8(d)
// config 5) on stderr: "latency_us median <m> p95 <p> max <x> frames <k>" over the frames after the first 10.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>

#include <rdvio_b200/gpu_image.hpp>

using rdvio::vector;

int main(int argc, char **argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: %s frames.bin out.txt\n", argv[0]); return 2; }
    FILE *f = std::fopen(argv[1], "rb");
    if (!f) { std::perror("frames"); return 2; }
    int hdr[3];
    if (std::fread(hdr, sizeof(int), 3, f) != 3) return 2;
    const int n = hdr[0], H = hdr[1], W = hdr[2];
    FILE *out = std::fopen(argv[2], "w");
    std::shared_ptr<rdvio::Image> last;
    std::vector<vector<2>> last_kp;
    std::vector<double> lat_us, pre_us, trk_us, det_us;
    using clk = std::chrono::steady_clock;
    auto us = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
    for (int i = 0; i < n; ++i) {
        auto img = std::make_shared<rdvio::extra::GpuImage>();      // rdvio.hpp:50-53
        cv::Mat gray(H, W);
        if (std::fread(gray.data, 1, (size_t)H * W, f) != (size_t)H * W) return 2;
        img->image = gray.clone();
        img->raw = gray.clone();
        img->t = 0.05 * i;
        const auto t0 = clk::now();
        img->preprocess(6.0, 8, 8);                                  // feature_tracker.cpp:32-34
        const auto t1 = clk::now();
        std::vector<vector<2>> kp;
        if (last) {
            std::vector<vector<2>> next;                             // no IMU prediction in this replay
            std::vector<char> status;
            last->track_keypoints(img.get(), last_kp, next, status); // frame.cpp:96
            for (size_t j = 0; j < status.size(); ++j)
                if (status[j]) kp.push_back(next[j]);
            last->release_image_buffer();                            // feature_tracker.cpp:94
        }
        const auto t2 = clk::now();
        img->detect_keypoints(kp, 150, 20.0);                        // frame.cpp:61
        const auto t3 = clk::now();
        if (i >= 10) {
            lat_us.push_back(us(t0, t3));
            pre_us.push_back(us(t0, t1));
            trk_us.push_back(us(t1, t2));
            det_us.push_back(us(t2, t3));
        }
        std::fprintf(out, "frame %d %zu\n", i, kp.size());
        for (auto &p : kp) std::fprintf(out, "%.17g %.17g\n", p.x(), p.y());
        last = img;
        last_kp = kp;
    }
    std::fclose(out);
    std::fclose(f);
    if (!lat_us.empty()) {
        auto med = [](std::vector<double> v, double q) { std::sort(v.begin(), v.end()); return v[(size_t)(q * (v.size() - 1))]; };
        std::fprintf(stderr, "latency_us median %.1f p95 %.1f max %.1f frames %zu | preprocess %.1f track %.1f detect %.1f\n",
                     med(lat_us, 0.5), med(lat_us, 0.95), med(lat_us, 1.0), lat_us.size(), med(pre_us, 0.5),
                     med(trk_us, 0.5), med(det_us, 0.5));
    }
    return 0;
}
