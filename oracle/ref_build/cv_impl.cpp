// cv_impl.cpp — the image-processing entry points of the OpenCV COMPAT LAYER (compat/opencv2/*), used only to
// build the reference's own ORBextractor.cc / ORBmatcher.cc into oracle/_ref/libvsg_ref.so.  TEST INFRASTRUCTURE.
//
// Every function forwards to a primitive of liborb_oracle.so that tests/test_oracle_cv2.py pins bit-for-bit to
// cv2 4.13.0 (resize, GaussianBlur, FAST incl. order and responses, fastAtan2, REFLECT_101 border).  Anything the
// reference does not ask for (other depths, kernels, border modes) aborts instead of guessing.
#include <opencv2/opencv.hpp>

#include <cstdio>
#include <cstdlib>

#include "../oracle.h"

namespace cv {

static void unsupported(const char *what) {
    std::fprintf(stderr, "vsg_ref compat layer: unsupported call: %s\n", what);
    std::abort();
}

float fastAtan2(float y, float x) { return orc_fast_atan2(y, x); }

void resize(InputArray _src, OutputArray _dst, Size dsize, double fx, double fy, int interpolation) {
    Mat src = _src.getMat();
    if (src.type() != CV_8UC1 || interpolation != INTER_LINEAR || fx != 0 || fy != 0) unsupported("resize variant");
    _dst.create(dsize, src.type());  // keeps the ROI of the padded level buffer (ORBextractor.cc:1179,1184)
    Mat dst = _dst.getMat();
    orc_resize_linear(src.data, src.cols, src.rows, (int)src.step, dst.data, dst.cols, dst.rows, (int)dst.step);
}

void GaussianBlur(InputArray _src, OutputArray _dst, Size ksize, double sigmaX, double sigmaY, int borderType) {
    Mat src = _src.getMat();
    if (src.type() != CV_8UC1 || ksize.width != 7 || ksize.height != 7 || sigmaX != 2 || sigmaY != 2 ||
        borderType != BORDER_REFLECT_101)
        unsupported("GaussianBlur variant");
    Mat tmp = src.clone();  // the reference blurs in place (src == dst, :1130)
    _dst.create(src.size(), src.type());
    Mat dst = _dst.getMat();
    orc_gaussian_blur7(tmp.data, tmp.cols, tmp.rows, (int)tmp.step, dst.data, (int)dst.step);
}

void copyMakeBorder(InputArray _src, OutputArray _dst, int top, int bottom, int left, int right, int borderType) {
    Mat src = _src.getMat();
    if (src.type() != CV_8UC1 || (borderType & ~BORDER_ISOLATED) != BORDER_REFLECT_101) unsupported("copyMakeBorder variant");
    // Without BORDER_ISOLATED OpenCV would read a sub-matrix's real surroundings; the extractor is only ever
    // handed whole images at level 0 (:1191), so both flavours reflect the image itself here.
    _dst.create(src.rows + top + bottom, src.cols + left + right, src.type());
    Mat dst = _dst.getMat();
    const int w = src.cols, h = src.rows;
    auto refl = [](int i, int n) {
        if (n == 1) return 0;
        while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
        return i;
    };
    // interior first (a no-op when src already is the centre ROI of dst, :1186), then rows, then the frame
    for (int y = 0; y < h; ++y) {
        uchar *d = dst.ptr(y + top) + left;
        const uchar *s = src.ptr(y);
        if (d != s) std::memmove(d, s, (size_t)w);
    }
    for (int y = 0; y < h; ++y) {
        uchar *row = dst.ptr(y + top);
        for (int x = 0; x < left; ++x) row[x] = row[left + refl(x - left, w)];
        for (int x = 0; x < right; ++x) row[left + w + x] = row[left + refl(w + x, w)];
    }
    const size_t rb = (size_t)dst.cols;
    for (int y = 0; y < top; ++y) std::memcpy(dst.ptr(y), dst.ptr(top + refl(y - top, h)), rb);
    for (int y = 0; y < bottom; ++y) std::memcpy(dst.ptr(top + h + y), dst.ptr(top + refl(h + y, h)), rb);
}

void FAST(InputArray _image, std::vector<KeyPoint> &keypoints, int threshold, bool nonmaxSuppression) {
    Mat img = _image.getMat();
    if (img.type() != CV_8UC1 || !nonmaxSuppression) unsupported("FAST variant");
    keypoints.clear();
    if (img.cols < 7 || img.rows < 7) return;
    const int cap = ((img.cols - 6) * (img.rows - 6) + 3) / 2 + 8;
    std::vector<int32_t> xys((size_t)cap * 3);
    const int n = orc_fast(img.data, img.cols, img.rows, (int)img.step, threshold, xys.data(), cap);
    if (n > cap) unsupported("FAST capacity");
    keypoints.reserve((size_t)n);
    for (int i = 0; i < n; ++i)
        keypoints.push_back(KeyPoint((float)xys[3 * i], (float)xys[3 * i + 1], 7.f, -1, (float)xys[3 * i + 2]));
}

void KeyPointsFilter::retainBest(std::vector<KeyPoint> &, int) { unsupported("KeyPointsFilter::retainBest (dead code path)"); }

}  // namespace cv
