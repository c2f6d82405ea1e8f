// oracle/shim/opencv2/opencv.hpp -- the slice of OpenCV 4.4 (Dockerfile:32) that
// src/paf.cpp + src/post_process.hpp + include/hyperpose/utility/{human,data}.hpp need,
// for compiling the reference's parser VERBATIM where OpenCV C++ is absent.
// cv::resize(INTER_AREA, upscale) and cv::GaussianBlur(17x17, sigma 3) forward to the
// restatements in oracle/paf_oracle.c, which are pinned bit-exactly against Python cv2
// (tests/test_oracle_cv_pin.py).  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cassert>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <string>
#include <memory>
#include <array>
#include "../../paf_oracle.h"

namespace cv {
struct Size {
    int width = 0, height = 0;
    Size() = default;
    Size(int w, int h) : width(w), height(h) {}
    int area() const { return width * height; }
    bool operator==(const Size& o) const { return width == o.width && height == o.height; }
    bool operator!=(const Size& o) const { return !(*this == o); }
};
// cv::Rect_<int> as OpenCV 4.4 core/types.hpp defines it (area, and a & b = the intersection, empty -> Rect()):
// what src/pose_proposal.cpp needs for its box NMS.
struct Rect {
    int x = 0, y = 0, width = 0, height = 0;
    Rect() = default;
    Rect(int x_, int y_, int w_, int h_) : x(x_), y(y_), width(w_), height(h_) {}
    int area() const { return width * height; }
};
inline Rect operator&(const Rect& a, const Rect& b)
{
    const int x1 = a.x > b.x ? a.x : b.x, y1 = a.y > b.y ? a.y : b.y;
    const int ax2 = a.x + a.width, bx2 = b.x + b.width, ay2 = a.y + a.height, by2 = b.y + b.height;
    Rect r(x1, y1, (ax2 < bx2 ? ax2 : bx2) - x1, (ay2 < by2 ? ay2 : by2) - y1);
    if (r.width <= 0 || r.height <= 0) r = Rect();
    return r;
}
struct Scalar {
    double val[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) : val{ a, b, c, d } {}
};
template <typename T> struct DataType;
template <> struct DataType<float> { static constexpr int type = 5; /* CV_32F */ };
template <> struct DataType<double> { static constexpr int type = 6; };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_CUBIC = 2, INTER_AREA = 3 };
enum { BORDER_DEFAULT = 4 };
class Mat {
public:
    Mat() = default;
    Mat(Size s, int type, void* data) : rows(s.height), cols(s.width), data(static_cast<unsigned char*>(data)), type_(type) {}
    Size size() const { return Size(cols, rows); }
    int type() const { return type_; }
    int rows = 0, cols = 0;
    unsigned char* data = nullptr;
private:
    int type_ = 0;
};
inline void resize(const Mat& src, Mat& dst, Size dsize, double = 0, double = 0, int interpolation = INTER_LINEAR)
{
    if (interpolation != INTER_AREA || src.type() != 5 || dst.size() != dsize
        || orc_resize_area(reinterpret_cast<const float*>(src.data), src.rows, src.cols,
               reinterpret_cast<float*>(dst.data), dsize.height, dsize.width)
            != 0)
        std::abort();
}
inline void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double sigmaX, double = 0, int = BORDER_DEFAULT)
{
    if (ksize.width != 17 || ksize.height != 17 || sigmaX != 3.0 || src.type() != 5) std::abort();
    orc_gaussian17(reinterpret_cast<const float*>(src.data), reinterpret_cast<float*>(dst.data), src.rows, src.cols);
}
} // namespace cv
