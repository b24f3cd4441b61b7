// opencv2/core/core.hpp — COMPAT LAYER (test infrastructure, not OpenCV).
//
// Exists for one purpose: to compile the reference's own sources
//   /root/reference/orb_slam3/src/ORBextractor.cc   (needs Mat, KeyPoint, Point, Size, Rect, InputArray,
//   /root/reference/orb_slam3/src/ORBmatcher.cc      OutputArray, cvRound, fastAtan2, CV_PI, ...)
// UNMODIFIED into oracle/_ref/libvsg_ref.so in an image that has no OpenCV C++ headers or libraries.
// Only the API subset those two files touch is provided.  The image-processing entry points declared
// here (FAST, resize, GaussianBlur, copyMakeBorder, fastAtan2) are implemented in ../cv_impl.cpp on top of
// the primitives of liborb_oracle.so, each of which is pinned bit-for-bit to cv2 4.13.0 by
// tests/test_oracle_cv2.py.  So: control flow, containers, std::sort, float expressions = the reference's
// own code; OpenCV arithmetic = cv2-pinned restatements.
#pragma once
#include <algorithm>  // real <opencv2/core.hpp> pulls these in; the reference relies on it (std::sort at :707)
#include <cassert>
#include <climits>
#include <cfloat>
#include <iostream>
#include <string>
#include <utility>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#define CV_PI 3.1415926535897932384626433832795
#define CV_CN_SHIFT 3
#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) (((depth) & 7) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)

typedef unsigned char uchar;
typedef unsigned short ushort;

// cvRound: round half to even (SSE2 cvtsd2si / lrint), SURVEY Appendix A5.
static inline int cvRound(double v) { return (int)std::lrint(v); }
static inline int cvRound(float v) { return (int)std::lrintf(v); }
static inline int cvRound(int v) { return v; }
static inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
static inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }

namespace cv {

typedef ::uchar uchar;
typedef ::ushort ushort;

template <typename T>
class Point_ {
public:
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    template <typename U> Point_(const Point_<U> &p) : x((T)p.x), y((T)p.y) {}
    Point_ &operator*=(float s) { x = (T)(x * s); y = (T)(y * s); return *this; }
    Point_ &operator*=(double s) { x = (T)(x * s); y = (T)(y * s); return *this; }
    Point_ &operator*=(int s) { x = (T)(x * s); y = (T)(y * s); return *this; }
    Point_ &operator+=(const Point_ &o) { x += o.x; y += o.y; return *this; }
    Point_ &operator-=(const Point_ &o) { x -= o.x; y -= o.y; return *this; }
};
template <typename T> inline Point_<T> operator+(const Point_<T> &a, const Point_<T> &b) { return Point_<T>(a.x + b.x, a.y + b.y); }
template <typename T> inline Point_<T> operator-(const Point_<T> &a, const Point_<T> &b) { return Point_<T>(a.x - b.x, a.y - b.y); }
template <typename T> inline bool operator==(const Point_<T> &a, const Point_<T> &b) { return a.x == b.x && a.y == b.y; }
typedef Point_<int> Point2i;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
typedef Point2i Point;

template <typename T>
class Point3_ {
public:
    T x, y, z;
    Point3_() : x(0), y(0), z(0) {}
    Point3_(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
};
typedef Point3_<float> Point3f;
typedef Point3_<double> Point3d;

template <typename T>
class Size_ {
public:
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
    T area() const { return width * height; }
};
typedef Size_<int> Size;

template <typename T>
class Rect_ {
public:
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T x_, T y_, T w, T h) : x(x_), y(y_), width(w), height(h) {}
};
typedef Rect_<int> Rect;

class Range {
public:
    int start, end;
    Range() : start(0), end(0) {}
    Range(int s, int e) : start(s), end(e) {}
};

// Field order and layout of cv::KeyPoint (7 x 4 bytes).
class KeyPoint {
public:
    Point2f pt;
    float size;
    float angle;
    float response;
    int octave;
    int class_id;
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(Point2f pt_, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
        : pt(pt_), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) {}
    KeyPoint(float x, float y, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
        : pt(x, y), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) {}
};

class _InputArray;
class _OutputArray;
typedef const _InputArray &InputArray;
typedef const _OutputArray &OutputArray;

// Reference-counted 2-D matrix header over a byte buffer with a row step: the cv::Mat semantics the two
// reference files rely on (shared ROI views, create() that keeps a fitting buffer, clone, at/ptr/row).
class Mat {
public:
    int flags = CV_8UC1;
    int rows = 0, cols = 0;
    uchar *data = nullptr;
    size_t step = 0;  // bytes per row (cv::MatStep converts to size_t the same way)

    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(Size sz, int type) { create(sz.height, sz.width, type); }
    Mat(int r, int c, int type, void *ext, size_t step_ = 0)
        : flags(type), rows(r), cols(c), data((uchar *)ext), step(step_ ? step_ : (size_t)c * elemSizeOf(type)) {}

    static size_t elemSizeOf(int type) {
        static const int depth_bytes[8] = {1, 1, 2, 2, 4, 4, 8, 2};
        return (size_t)depth_bytes[type & 7] * (size_t)((type >> CV_CN_SHIFT) + 1);
    }

    void create(int r, int c, int type) {
        if (data && r == rows && c == cols && type == flags) return;  // cv::Mat::create keeps a fitting buffer
        flags = type; rows = r; cols = c;
        step = (size_t)c * elemSizeOf(type);
        owner_ = std::shared_ptr<uchar>(new uchar[(size_t)r * step + 64], std::default_delete<uchar[]>());
        data = owner_.get();
    }
    void create(Size sz, int type) { create(sz.height, sz.width, type); }
    void release() { owner_.reset(); data = nullptr; rows = cols = 0; step = 0; }

    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return flags; }
    int depth() const { return flags & 7; }
    int channels() const { return (flags >> CV_CN_SHIFT) + 1; }
    size_t elemSize() const { return elemSizeOf(flags); }
    size_t elemSize1() const { return elemSizeOf(flags & 7); }
    size_t step1() const { return step / elemSize1(); }
    size_t total() const { return (size_t)rows * cols; }
    Size size() const { return Size(cols, rows); }
    bool isContinuous() const { return step == (size_t)cols * elemSize() || rows <= 1; }

    uchar *ptr(int r = 0) { return data + (size_t)r * step; }
    const uchar *ptr(int r = 0) const { return data + (size_t)r * step; }
    template <typename T> T *ptr(int r = 0) { return (T *)(data + (size_t)r * step); }
    template <typename T> const T *ptr(int r = 0) const { return (const T *)(data + (size_t)r * step); }
    template <typename T> T &at(int r, int c) { return ((T *)(data + (size_t)r * step))[c]; }
    template <typename T> const T &at(int r, int c) const { return ((const T *)(data + (size_t)r * step))[c]; }

    Mat row(int r) const { Mat m(*this); m.rows = 1; m.data = data + (size_t)r * step; return m; }
    Mat rowRange(int a, int b) const {
        assert(0 <= a && a <= b && b <= rows);
        Mat m(*this); m.rows = b - a; m.data = data + (size_t)a * step; return m;
    }
    Mat colRange(int a, int b) const {
        assert(0 <= a && a <= b && b <= cols);
        Mat m(*this); m.cols = b - a; m.data = data + (size_t)a * elemSize(); return m;
    }
    Mat operator()(const Rect &roi) const {
        assert(roi.x >= 0 && roi.y >= 0 && roi.x + roi.width <= cols && roi.y + roi.height <= rows);
        Mat m(*this);
        m.rows = roi.height; m.cols = roi.width;
        m.data = data + (size_t)roi.y * step + (size_t)roi.x * elemSize();
        return m;
    }
    Mat clone() const {
        Mat m;
        if (empty()) return m;
        m.create(rows, cols, flags);
        const size_t rb = (size_t)cols * elemSize();
        for (int r = 0; r < rows; ++r) std::memcpy(m.ptr(r), ptr(r), rb);
        return m;
    }
    void copyTo(OutputArray dst) const;

    static Mat zeros(int r, int c, int type) {
        Mat m(r, c, type);
        if (m.data) std::memset(m.data, 0, (size_t)r * m.step);
        return m;
    }
    static Mat zeros(Size sz, int type) { return zeros(sz.height, sz.width, type); }

private:
    std::shared_ptr<uchar> owner_;
};

// Proxy classes: the reference only passes cv::Mat through them.
class _InputArray {
public:
    _InputArray() : m_(nullptr) {}
    _InputArray(const Mat &m) : m_(&m) {}
    bool empty() const { return !m_ || m_->empty(); }
    Mat getMat(int = -1) const { return m_ ? *m_ : Mat(); }
protected:
    const Mat *m_;
};
class _OutputArray : public _InputArray {
public:
    _OutputArray() {}
    _OutputArray(Mat &m) : _InputArray(m) {}
    _OutputArray(const Mat &m) : _InputArray(m) {}  // header of a temporary view (e.g. m.row(i)), as in OpenCV
    void create(int r, int c, int type) const { mat()->create(r, c, type); }
    void create(Size sz, int type) const { mat()->create(sz, type); }
    void release() const { mat()->release(); }
    Mat &getMatRef() const { return *mat(); }
private:
    Mat *mat() const { return const_cast<Mat *>(m_); }
};
inline _InputArray noArray() { return _InputArray(); }

inline void Mat::copyTo(OutputArray dst) const {
    dst.create(rows, cols, flags);
    Mat d = dst.getMat();
    const size_t rb = (size_t)cols * elemSize();
    for (int r = 0; r < rows; ++r) std::memmove(d.ptr(r), ptr(r), rb);
}

enum BorderTypes {
    BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4,
    BORDER_REFLECT101 = 4, BORDER_DEFAULT = 4, BORDER_ISOLATED = 16
};
enum NormTypes { NORM_INF = 1, NORM_L1 = 2, NORM_L2 = 4, NORM_HAMMING = 6 };

float fastAtan2(float y, float x);  // ../cv_impl.cpp -> orc_fast_atan2 (cv2-pinned, Appendix A4)

}  // namespace cv
