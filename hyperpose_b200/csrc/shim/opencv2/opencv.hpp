// shim/opencv2/opencv.hpp -- a small stand-in for the part of OpenCV that the reference's public headers, its
// src/{stream,human,data}.cpp and its examples/*.cpp touch, so that all of them compile UNCHANGED and run on a machine
// without OpenCV (this container, the GPU box).  With real OpenCV installed this directory is left off the include path.
//
// It is NOT on the hot path: the product resizes and batches frames on the GPU (hp_engine_stage_frame_u8).  What is here:
//   * cv::Size / Point / Rect / Scalar / Vec / Mat (ref-counted storage, clone, ptr<>, external-buffer views);
//   * cv::resize for CV_8UC3: OpenCV's 11-bit fixed-point bilinear (INTER_LINEAR; an exact 2x reduction averages like
//     INTER_AREA) -- the same arithmetic resize_u8c3_kernel implements on the GPU; nearest-neighbour for other types;
//   * cv::copyMakeBorder(BORDER_CONSTANT), cv::line / circle (integer rasterisers), addWeighted; putText / imshow /
//     waitKey are accepted and do nothing (no display, no fonts);
//   * image / video I/O on a trivial container: binary PPM ("P6").  cv::imread reads a P6 file whatever its extension,
//     cv::imwrite writes one; a "video" is P6 frames back to back in one file (VideoCapture counts them on open, 25 fps,
//     VideoWriter appends).  There are no JPEG / PNG / AVI codecs here: tests write their inputs in this format.
#pragma once
#include <algorithm>
#include <cassert>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#define CV_MAJOR_VERSION 4
#define CV_MINOR_VERSION 4
#define CV_8U 0
#define CV_32F 5
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)

namespace cv {
using String = std::string;

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
    int area() const { return width * height; }
};
inline Rect operator&(const Rect& a, const Rect& b)
{
    const int x1 = std::max(a.x, b.x), y1 = std::max(a.y, b.y);
    const int x2 = std::min(a.x + a.width, b.x + b.width), y2 = std::min(a.y + a.height, b.y + b.height);
    return (x2 <= x1 || y2 <= y1) ? Rect() : Rect(x1, y1, x2 - x1, y2 - y1);
}
struct Scalar {
    double val[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) : val{ a, b, c, d } {}
    double operator[](int i) const { return val[i]; }
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
    Mat(int rows_, int cols_, int type, void* ext) : rows(rows_), cols(cols_), data(static_cast<unsigned char*>(ext)), type_(type) {}
    void create(int rows_, int cols_, int type)
    {
        rows = rows_; cols = cols_; type_ = type;
        store_ = std::shared_ptr<unsigned char>(new unsigned char[total() * elemSize() + 4](), std::default_delete<unsigned char[]>());
        data = store_.get();
    }
    Mat clone() const
    {
        Mat m;
        if (empty()) return m;
        m.create(rows, cols, type_);
        std::memcpy(m.data, data, total() * elemSize());
        return m;
    }
    void copyTo(Mat& dst) const { dst = clone(); }
    Size size() const { return Size(cols, rows); }
    int type() const { return type_; }
    int depth() const { return type_ & 7; }
    int channels() const { return (type_ >> 3) + 1; }
    size_t elemSize() const { return (size_t)channels() * ((type_ & 7) == CV_32F ? 4 : 1); }
    size_t total() const { return (size_t)rows * cols; }
    bool isContinuous() const { return true; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    template <typename T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + (size_t)r * cols * elemSize()); }
    template <typename T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * cols * elemSize()); }
    template <typename T> T& at(int r, int c) { return ptr<T>(r)[c]; }
    template <typename T> const T& at(int r, int c) const { return ptr<T>(r)[c]; }
    int rows = 0, cols = 0;
    unsigned char* data = nullptr;

private:
    int type_ = 0;
    std::shared_ptr<unsigned char> store_;
};

enum { CAP_PROP_POS_FRAMES = 1, CAP_PROP_FRAME_WIDTH = 3, CAP_PROP_FRAME_HEIGHT = 4, CAP_PROP_FPS = 5, CAP_PROP_FOURCC = 6, CAP_PROP_FRAME_COUNT = 7 };
enum { CAP_ANY = 0, CAP_V4L2 = 200 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_AREA = 3, FILLED = -1 };
enum { BORDER_CONSTANT = 0 };
enum { FONT_HERSHEY_SIMPLEX = 0 };
enum LineTypes { LINE_4 = 4, LINE_8 = 8, LINE_AA = 16 };

namespace shim_detail {
    inline unsigned char sat8(double v) { return (unsigned char)std::min(255.0, std::max(0.0, std::nearbyint(v))); }
    // one axis of OpenCV's bilinear resize: source index + two 11-bit coefficients per destination position
    inline void linear_table(int src, int dst, bool clamp, std::vector<int>& idx, std::vector<short>& coef)
    {
        idx.resize(dst); coef.resize(2 * (size_t)dst);
        const double scale = 1.0 / ((double)dst / (double)src);
        for (int d = 0; d < dst; ++d) {
            float f = (float)((d + 0.5) * scale - 0.5);
            int s = (int)std::floor(f);
            f -= (float)s;
            if (clamp) {
                if (s < 0) { f = 0.f; s = 0; }
                if (s >= src - 1) { f = 0.f; s = src - 1; }
            }
            idx[d] = s;
            coef[2 * d] = (short)std::lrint((1.f - f) * 2048.f);
            coef[2 * d + 1] = (short)std::lrint(f * 2048.f);
        }
    }
}

inline void resize(const Mat& src, Mat& dst, Size size, double = 0, double = 0, int = INTER_LINEAR)
{
    if (src.size() == size) { if (&dst != &src) dst = src; return; }
    Mat out(size, src.type());
    if (src.type() == CV_8UC3 && src.rows > 0 && src.cols > 0) {
        const int sh = src.rows, sw = src.cols, dh = size.height, dw = size.width;
        if (sh == 2 * dh && sw == 2 * dw) {   // exact 2x reduction: OpenCV switches to the area average
            for (int y = 0; y < dh; ++y)
                for (int x = 0; x < dw; ++x) {
                    const unsigned char* p = src.data + ((size_t)(2 * y) * sw + 2 * x) * 3;
                    for (int c = 0; c < 3; ++c)
                        out.data[((size_t)y * dw + x) * 3 + c] = (unsigned char)((p[c] + p[3 + c] + p[(size_t)sw * 3 + c] + p[(size_t)sw * 3 + 3 + c] + 2) >> 2);
                }
        } else {
            std::vector<int> xi, yi; std::vector<short> xa, ya;
            shim_detail::linear_table(sw, dw, true, xi, xa);
            shim_detail::linear_table(sh, dh, false, yi, ya);
            for (int y = 0; y < dh; ++y) {
                const int y0 = std::min(std::max(yi[y], 0), sh - 1), y1 = std::min(std::max(yi[y] + 1, 0), sh - 1);
                const int b0 = ya[2 * y], b1 = ya[2 * y + 1];
                const unsigned char* r0 = src.data + (size_t)y0 * sw * 3;
                const unsigned char* r1 = src.data + (size_t)y1 * sw * 3;
                for (int x = 0; x < dw; ++x) {
                    const int x0 = xi[x], x1 = std::min(x0 + 1, sw - 1), a0 = xa[2 * x], a1 = xa[2 * x + 1];
                    for (int c = 0; c < 3; ++c) {
                        const int S0 = r0[x0 * 3 + c] * a0 + r0[x1 * 3 + c] * a1;
                        const int S1 = r1[x0 * 3 + c] * a0 + r1[x1 * 3 + c] * a1;
                        const int v = (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2;
                        out.data[((size_t)y * dw + x) * 3 + c] = (unsigned char)std::min(std::max(v, 0), 255);
                    }
                }
            }
        }
    } else {
        const size_t es = src.elemSize();
        for (int y = 0; y < size.height; ++y) {
            const int sy = (int)((long long)y * src.rows / size.height);
            for (int x = 0; x < size.width; ++x) {
                const int sx = (int)((long long)x * src.cols / size.width);
                std::memcpy(out.data + ((size_t)y * size.width + x) * es, src.data + ((size_t)sy * src.cols + sx) * es, es);
            }
        }
    }
    dst = out;
}

inline void copyMakeBorder(const Mat& src, Mat& dst, int top, int bottom, int left, int right, int /*borderType*/, const Scalar& value = Scalar())
{
    Mat out(src.rows + top + bottom, src.cols + left + right, src.type());
    const size_t es = src.elemSize();
    const int cn = src.channels();
    for (int y = 0; y < out.rows; ++y)
        for (int x = 0; x < out.cols; ++x) {
            unsigned char* o = out.data + ((size_t)y * out.cols + x) * es;
            const int sy = y - top, sx = x - left;
            if (sy >= 0 && sy < src.rows && sx >= 0 && sx < src.cols) std::memcpy(o, src.data + ((size_t)sy * src.cols + sx) * es, es);
            else if (src.depth() == CV_8U) for (int c = 0; c < cn; ++c) o[c] = shim_detail::sat8(value.val[c & 3]);
        }
    dst = out;
}

inline void circle(Mat& img, Point c, int radius, const Scalar& color, int /*thickness*/ = 1, int = LINE_8, int = 0)
{
    if (img.empty() || img.depth() != CV_8U) return;
    const int cn = img.channels(), r = std::max(radius, 0);
    for (int y = std::max(0, c.y - r); y <= std::min(img.rows - 1, c.y + r); ++y)
        for (int x = std::max(0, c.x - r); x <= std::min(img.cols - 1, c.x + r); ++x)
            if ((x - c.x) * (x - c.x) + (y - c.y) * (y - c.y) <= r * r)
                for (int k = 0; k < cn; ++k) img.data[((size_t)y * img.cols + x) * cn + k] = shim_detail::sat8(color.val[k & 3]);
}
inline void line(Mat& img, Point a, Point b, const Scalar& color, int thickness = 1, int = LINE_8, int = 0)
{
    const int steps = std::max(std::abs(b.x - a.x), std::abs(b.y - a.y));
    const int r = std::max(0, thickness / 2);
    for (int i = 0; i <= steps; ++i) {
        const double t = steps ? (double)i / steps : 0.0;
        circle(img, Point((int)std::lround(a.x + t * (b.x - a.x)), (int)std::lround(a.y + t * (b.y - a.y))), r, color, FILLED);
    }
}
inline void addWeighted(const Mat& a, double alpha, const Mat& b, double beta, double gamma, Mat& dst)
{
    Mat out(a.size(), a.type());
    const size_t n = a.total() * a.elemSize();
    if (a.depth() == CV_8U && b.size() == a.size() && b.type() == a.type())
        for (size_t i = 0; i < n; ++i) out.data[i] = shim_detail::sat8(a.data[i] * alpha + b.data[i] * beta + gamma);
    dst = out;
}
inline void putText(Mat&, const std::string&, Point, int, double, Scalar, int = 1, int = LINE_8, bool = false) {}
inline void imshow(const std::string&, const Mat&) {}
inline int waitKey(int = 0) { return -1; }

// ---- binary PPM (P6) as the only image / video container -------------------------------------------------------------
namespace shim_detail {
    // reads one P6 frame at the stream position; false at end of file / malformed data
    inline bool read_p6(std::istream& in, Mat& m)
    {
        std::string magic;
        if (!(in >> magic) || magic != "P6") return false;
        int w = 0, h = 0, maxv = 0;
        auto next_int = [&](int& v) {
            for (;;) {
                in >> std::ws;
                if (in.peek() == '#') { std::string skip; std::getline(in, skip); continue; }
                return (bool)(in >> v);
            }
        };
        if (!next_int(w) || !next_int(h) || !next_int(maxv) || w <= 0 || h <= 0 || maxv != 255) return false;
        in.get();   // the single whitespace after maxval
        Mat rgb(h, w, CV_8UC3);
        if (!in.read(reinterpret_cast<char*>(rgb.data), (std::streamsize)rgb.total() * 3)) return false;
        for (size_t i = 0; i < rgb.total(); ++i) std::swap(rgb.data[3 * i], rgb.data[3 * i + 2]);   // PPM is RGB, cv::Mat is BGR
        m = rgb;
        return true;
    }
    inline void write_p6(std::ostream& out, const Mat& m)
    {
        out << "P6\n" << m.cols << ' ' << m.rows << "\n255\n";
        std::vector<unsigned char> row((size_t)m.cols * 3);
        for (int y = 0; y < m.rows; ++y) {
            const unsigned char* s = m.data + (size_t)y * m.cols * 3;
            for (int x = 0; x < m.cols; ++x) { row[3 * x] = s[3 * x + 2]; row[3 * x + 1] = s[3 * x + 1]; row[3 * x + 2] = s[3 * x]; }
            out.write(reinterpret_cast<const char*>(row.data()), (std::streamsize)row.size());
        }
    }
}

inline Mat imread(const std::string& path, int = 1)
{
    std::ifstream f(path, std::ios::binary);
    Mat m;
    if (f) shim_detail::read_p6(f, m);
    return m;
}
inline bool imwrite(const std::string& path, const Mat& m)
{
    if (m.empty() || m.type() != CV_8UC3) return false;
    std::ofstream f(path, std::ios::binary | std::ios::trunc);
    if (!f) return false;
    shim_detail::write_p6(f, m);
    return (bool)f;
}

class VideoCapture {
public:
    VideoCapture() = default;
    explicit VideoCapture(const std::string& path) { open(path); }
    bool open(const std::string& path, int = CAP_ANY)
    {
        file_ = std::make_shared<std::ifstream>(path, std::ios::binary);
        count_ = pos_ = 0; w_ = h_ = 0;
        if (!*file_) { file_.reset(); return false; }
        Mat m;
        while (shim_detail::read_p6(*file_, m)) { if (!count_) { w_ = m.cols; h_ = m.rows; } ++count_; }
        file_->clear();
        file_->seekg(0);
        if (!count_) file_.reset();
        return count_ > 0;
    }
    bool open(int /*camera index*/, int = CAP_ANY) { file_.reset(); return false; }   // no cameras here
    bool isOpened() const { return (bool)file_; }
    double get(int prop) const
    {
        switch (prop) {
        case CAP_PROP_FRAME_COUNT: return count_;
        case CAP_PROP_POS_FRAMES: return pos_;
        case CAP_PROP_FRAME_WIDTH: return w_;
        case CAP_PROP_FRAME_HEIGHT: return h_;
        case CAP_PROP_FPS: return 25.0;
        default: return 0.0;
        }
    }
    bool read(Mat& m)
    {
        m = Mat();
        if (!file_ || !shim_detail::read_p6(*file_, m)) { m = Mat(); return false; }
        ++pos_;
        return true;
    }
    VideoCapture& operator>>(Mat& m) { read(m); return *this; }

private:
    std::shared_ptr<std::ifstream> file_;
    int count_ = 0, pos_ = 0, w_ = 0, h_ = 0;
};

class VideoWriter {
public:
    VideoWriter() = default;
    VideoWriter(const std::string& path, int /*fourcc*/, double /*fps*/, Size size) : file_(std::make_shared<std::ofstream>(path, std::ios::binary | std::ios::trunc)), size_(size) {}
    bool isOpened() const { return file_ && (bool)*file_; }
    VideoWriter& operator<<(const Mat& m) { write(m); return *this; }
    void write(const Mat& m)
    {
        ++frames_written;
        last_size = m.size();
        if (file_ && *file_ && !m.empty() && m.type() == CV_8UC3) shim_detail::write_p6(*file_, m);
    }
    void release() { if (file_) file_->flush(); }
    size_t frames_written = 0;
    Size last_size;

private:
    std::shared_ptr<std::ofstream> file_;
    Size size_;
};

inline std::ostream& operator<<(std::ostream& o, const Size& s) { return o << '[' << s.width << " x " << s.height << ']'; }
} // namespace cv
