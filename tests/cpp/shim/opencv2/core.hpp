// TEST SHIM: the handful of cv::Mat members the plugin boundary touches (rows, cols, step, data, release, clone).
#pragma once
#include <cstddef>
#include <cstring>
#include <memory>

namespace cv {
struct Mat {
    int rows = 0, cols = 0;
    size_t step = 0;
    unsigned char *data = nullptr;
    std::shared_ptr<unsigned char> store;
    Mat() {}
    Mat(int r, int c) : rows(r), cols(c), step((size_t)c), store(new unsigned char[(size_t)r * c], std::default_delete<unsigned char[]>()) {
        data = store.get();
    }
    Mat clone() const {
        Mat m(rows, cols);
        if (data) std::memcpy(m.data, data, (size_t)rows * cols);
        return m;
    }
    void release() { store.reset(); data = nullptr; rows = cols = 0; step = 0; }
};
}  // namespace cv
