// shim/opencv2/opencv.hpp -- minimal stand-in for the OpenCV types that the reference's PUBLIC headers
// (include/hyperpose/utility/{human,data}.hpp, operator/dnn/tensorrt.hpp, operator/parser/*.hpp) mention,
// so that the B200 drop-in (hyperpose_api/*.cpp) and a user program can be compiled on a machine without
// OpenCV (this container, the GPU box).  With real OpenCV installed this directory is simply left off the
// include path.  Only construction / geometry / raw-pixel access is provided: no image processing.
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <iostream>
#include <string>
#include <cstring>
#include <memory>
#include <vector>

#define CV_8U 0
#define CV_32F 5
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << 3))
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)

namespace cv {
struct Size {
    int width = 0, height = 0;
    Size() = default;
    Size(int w, int h) : width(w), height(h) {}
    int area() const { return width * height; }
    bool operator==(const Size& o) const { return width == o.width && height == o.height; }
    bool operator!=(const Size& o) const { return !(*this == o); }
};
struct Point {
    int x = 0, y = 0;
    Point() = default;
    Point(int x_, int y_) : x(x_), y(y_) {}
};
struct Rect {
    int x = 0, y = 0, width = 0, height = 0;
    Rect() = default;
    Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {}
};
struct Scalar {
    double val[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) : val{ a, b, c, d } {}
};
template <typename T, int N> struct Vec {
    T val[N];
    T& operator[](int i) { return val[i]; }
    const T& operator[](int i) const { return val[i]; }
};
using Vec3b = Vec<unsigned char, 3>;

class Mat {
public:
    Mat() = default;
    Mat(int rows_, int cols_, int type) { create(rows_, cols_, type); }
    Mat(Size s, int type) { create(s.height, s.width, type); }
    Mat(Size s, int type, void* ext) : rows(s.height), cols(s.width), data(static_cast<unsigned char*>(ext)), type_(type) {}
    void create(int rows_, int cols_, int type)
    {
        rows = rows_; cols = cols_; type_ = type;
        store_ = std::shared_ptr<unsigned char>(new unsigned char[total() * elemSize()], std::default_delete<unsigned char[]>());
        data = store_.get();
    }
    Size size() const { return Size(cols, rows); }
    int type() const { return type_; }
    int channels() const { return (type_ >> 3) + 1; }
    size_t elemSize() const { return (size_t)channels() * ((type_ & 7) == CV_32F ? 4 : 1); }
    size_t total() const { return (size_t)rows * cols; }
    bool isContinuous() const { return true; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    template <typename T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + (size_t)r * cols * elemSize()); }
    template <typename T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * cols * elemSize()); }
    int rows = 0, cols = 0;
    unsigned char* data = nullptr;

private:
    int type_ = 0;
    std::shared_ptr<unsigned char> store_;
};

// ---- what the reference's stream scheduler touches (include/hyperpose/stream/stream.hpp, src/stream.cpp) ----
// Declarations + the least behaviour that lets the UNCHANGED scheduler sources compile and run on in-memory frames where OpenCV
// is absent: resize is the identity for network-sized frames (nearest-neighbour otherwise -- a stand-in, NOT cv::resize; the product's
// own resize path is hp_engine_stage_frame_u8), the capture yields nothing, the writer counts what it is given.
enum { CAP_PROP_POS_FRAMES = 1, CAP_PROP_FRAME_WIDTH = 3, CAP_PROP_FRAME_HEIGHT = 4, CAP_PROP_FPS = 5, CAP_PROP_FOURCC = 6, CAP_PROP_FRAME_COUNT = 7 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_AREA = 3, FILLED = -1 };

inline void resize(const Mat& src, Mat& dst, Size size, double = 0, double = 0, int = INTER_LINEAR)
{
    if (src.size() == size) { if (&dst != &src) dst = src; return; }
    Mat out(size, src.type());
    const size_t es = src.elemSize();
    for (int y = 0; y < size.height; ++y) {
        const int sy = (int)((long long)y * src.rows / size.height);
        for (int x = 0; x < size.width; ++x) {
            const int sx = (int)((long long)x * src.cols / size.width);
            std::memcpy(out.data + ((size_t)y * size.width + x) * es, src.data + ((size_t)sy * src.cols + sx) * es, es);
        }
    }
    dst = out;
}

class VideoCapture {
public:
    VideoCapture() = default;
    explicit VideoCapture(const std::string&) {}
    bool isOpened() const { return false; }
    double get(int) const { return 0.0; }
    VideoCapture& operator>>(Mat& m) { m = Mat(); return *this; }
};

class VideoWriter {
public:
    VideoWriter() = default;
    VideoWriter(const std::string&, int, double, Size) {}
    bool isOpened() const { return true; }
    VideoWriter& operator<<(const Mat& m) { ++frames_written; last_size = m.size(); return *this; }
    size_t frames_written = 0;
    Size last_size;
};

inline bool imwrite(const std::string&, const Mat&) { return true; }
inline std::ostream& operator<<(std::ostream& o, const Size& s) { return o << '[' << s.width << " x " << s.height << ']'; }
} // namespace cv
