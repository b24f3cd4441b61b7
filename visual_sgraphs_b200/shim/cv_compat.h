// cv_compat.h — the small subset of OpenCV's core types that the ORB front-end's public interface uses
// (cv::Mat for 8-bit images / descriptor matrices, cv::KeyPoint, Point2f, Size, Rect, InputArray,
// OutputArray).  It exists only so that the drop-in classes can be compiled and tested where OpenCV's C++
// headers are not installed (this build image).  Inside the reference's tree define VSG_HAVE_OPENCV and the
// real <opencv2/core.hpp> is used instead; nothing else in the shim changes.
#pragma once
#ifdef VSG_HAVE_OPENCV
#include <opencv2/core.hpp>
#else
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5

namespace cv {

typedef unsigned char uchar;

template <typename T>
struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    Point_ &operator*=(float s) { x = (T)(x * s); y = (T)(y * s); return *this; }
};
typedef Point_<float> Point2f;
typedef Point_<int> Point2i;
typedef Point2i Point;

struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
};

struct Rect {
    int x, y, width, height;
    Rect() : x(0), y(0), width(0), height(0) {}
    Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {}
};

// same field order and layout as cv::KeyPoint (28 bytes) == vsg_keypoint
struct KeyPoint {
    Point2f pt;
    float size;
    float angle;
    float response;
    int octave;
    int class_id;
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
};

// Reference-counted 2-D byte matrix with a row step; enough of cv::Mat for CV_8UC1 data.
class Mat {
public:
    int rows = 0, cols = 0;
    uchar *data = nullptr;
    size_t step = 0;

    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(Size sz, int type) { create(sz.height, sz.width, type); }
    Mat(int r, int c, int /*type*/, void *ext, size_t step_) : rows(r), cols(c), data((uchar *)ext), step(step_ ? step_ : (size_t)c) {}

    void create(int r, int c, int /*type*/) {
        if (r == rows && c == cols && data && step == (size_t)c && owner_) return;
        rows = r; cols = c; step = (size_t)c;
        owner_.reset(new std::vector<uchar>((size_t)r * c));
        data = owner_->data();
    }
    void create(Size sz, int type) { create(sz.height, sz.width, type); }
    void release() { owner_.reset(); data = nullptr; rows = cols = 0; step = 0; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return CV_8UC1; }
    Size size() const { return Size(cols, rows); }
    size_t step1() const { return step; }
    bool isContinuous() const { return step == (size_t)cols; }

    uchar *ptr(int r = 0) { return data + (size_t)r * step; }
    const uchar *ptr(int r = 0) const { return data + (size_t)r * step; }
    template <typename T> T *ptr(int r = 0) { return (T *)(data + (size_t)r * step); }
    template <typename T> const T *ptr(int r = 0) const { return (const T *)(data + (size_t)r * step); }
    template <typename T> T &at(int r, int c) { return ((T *)(data + (size_t)r * step))[c]; }
    template <typename T> const T &at(int r, int c) const { return ((const T *)(data + (size_t)r * step))[c]; }

    Mat row(int r) const { Mat m(*this); m.rows = 1; m.data = data + (size_t)r * step; return m; }
    Mat operator()(const Rect &roi) const {
        Mat m(*this);
        m.rows = roi.height; m.cols = roi.width;
        m.data = data + (size_t)roi.y * step + roi.x;
        return m;
    }
    Mat clone() const {
        Mat m(rows, cols, CV_8UC1);
        for (int r = 0; r < rows; ++r) std::memcpy(m.ptr(r), ptr(r), (size_t)cols);
        return m;
    }

private:
    std::shared_ptr<std::vector<uchar>> owner_;
};

// The proxy classes collapse to references for the purposes of this interface.
class _InputArray {
public:
    _InputArray() : m_(nullptr) {}
    _InputArray(const Mat &m) : m_(&m) {}
    bool empty() const { return !m_ || m_->empty(); }
    Mat getMat() const { return m_ ? *m_ : Mat(); }
private:
    const Mat *m_;
};
class _OutputArray {
public:
    _OutputArray(Mat &m) : m_(&m) {}
    void create(int r, int c, int type) const { m_->create(r, c, type); }
    void release() const { m_->release(); }
    Mat getMat() const { return *m_; }
    Mat &getMatRef() const { return *m_; }
private:
    Mat *m_;
};
typedef const _InputArray &InputArray;
typedef const _OutputArray &OutputArray;
inline _InputArray noArray() { return _InputArray(); }

}  // namespace cv
#endif  // VSG_HAVE_OPENCV
