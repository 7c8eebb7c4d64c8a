// TEST SHIM standing in for the reference's <rdvio/types.h> (which needs Eigen + OpenCV, absent from
// this image).  Declares only what the Image plugin boundary uses, with the signatures of
// src/rdvio/include/rdvio/types.h:153-177 and the memory layout of Eigen::Matrix<double,2,1>.
#pragma once
#include <cstddef>
#include <vector>
#include <opencv2/core.hpp>

typedef unsigned char uchar;

namespace rdvio {

struct Vec2Shim {                       // layout-compatible with Eigen::Matrix<double, 2, 1>
    double v[2];
    Vec2Shim() {}
    Vec2Shim(double x, double y) { v[0] = x; v[1] = y; }
    double &x() { return v[0]; }
    double &y() { return v[1]; }
    const double &x() const { return v[0]; }
    const double &y() const { return v[1]; }
};
template <int N> struct VecSelect;
template <> struct VecSelect<2> { typedef Vec2Shim type; };
template <int N> using vector = typename VecSelect<N>::type;

class Image {
  public:
    double t;
    virtual uchar *get_rawdata() const = 0;
    virtual size_t width() const = 0;
    virtual size_t height() const = 0;
    virtual size_t level_num() const { return 0; }
    virtual double evaluate(const vector<2> &u, int level = 0) const = 0;
    virtual double evaluate(const vector<2> &u, vector<2> &ddu, int level = 0) const = 0;
    virtual ~Image() = default;
    virtual void preprocess(double clipLimit, int width, int height) {}
    virtual void release_image_buffer() = 0;
    virtual void detect_keypoints(std::vector<vector<2>> &keypoints, size_t max_points = 0,
                                  double keypoint_distance = 0.5) const = 0;
    virtual void track_keypoints(const Image *next_image, const std::vector<vector<2>> &curr_keypoints,
                                 std::vector<vector<2>> &next_keypoints, std::vector<char> &result_status) const = 0;
};

}  // namespace rdvio
