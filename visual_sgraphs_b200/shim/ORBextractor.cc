// ORBextractor.cc — forwards VS_GRAPHS::ORBextractor to the CUDA library (include/vsg_cuda.h).
// Reference behaviour mirrored: orb_slam3/src/ORBextractor.cc:411-470 (tables, via vsg_extractor_tables),
// :1083-1169 (operator(): -1 on empty image, _keypoints replaced, descriptors create()d or release()d,
// monoIndex returned), :1171-1195 (mvImagePyramid[l] = ROI of a (w+38)x(h+38) REFLECT_101-bordered buffer).
#include "ORBextractor.h"

#include <cassert>
#include <cstring>
#include <stdexcept>
#include <string>

#include "../../include/vsg_cuda.h"

namespace VS_GRAPHS {

static const int EDGE_THRESHOLD = 19;

static void Check(vsg_status st, const char *what) {
    // the reference has no error path here (asserts / OpenCV exceptions); a CUDA failure is fatal
    if (st != VSG_OK) throw std::runtime_error(std::string(what) + ": " + vsg_last_error());
}

ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST)
    : nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), iniThFAST(_iniThFAST), minThFAST(_minThFAST) {
    EnsureHandle();
    mvScaleFactor.resize(nlevels); mvInvScaleFactor.resize(nlevels);
    mvLevelSigma2.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
    mnFeaturesPerLevel.resize(nlevels);
    Check(vsg_extractor_tables(mpHandle, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(),
                               mvInvLevelSigma2.data(), mnFeaturesPerLevel.data()), "vsg_extractor_tables");
    mvImagePyramid.resize(nlevels);
}

ORBextractor::~ORBextractor() { vsg_extractor_destroy(mpHandle); }

void ORBextractor::EnsureHandle() {
    if (mpHandle) return;
    vsg_orb_params p;
    p.nfeatures = nfeatures; p.scale_factor = (float)scaleFactor; p.nlevels = nlevels;
    p.ini_th_fast = iniThFAST; p.min_th_fast = minThFAST;
    Check(vsg_extractor_create(&p, mnDevice, 1, &mpHandle), "vsg_extractor_create");
}

void ORBextractor::SetDevice(int device) {
    if (device == mnDevice) return;
    vsg_extractor_destroy(mpHandle);
    mpHandle = nullptr;
    mnDevice = device;
    EnsureHandle();
}

void ORBextractor::SetRectification(const float *map1, const float *map2, int width, int height) {
    Check(vsg_extractor_set_rectify_map(mpHandle, 0, map1, map2, width, height), "vsg_extractor_set_rectify_map");
    mnRectWidth = width;
    mnRectHeight = height;
}

static inline int Reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = (i < 0) ? -i : 2 * n - 2 - i;
    return i;
}

int ORBextractor::operator()(cv::InputArray _image, cv::InputArray /*_mask*/, std::vector<cv::KeyPoint> &_keypoints,
                             cv::OutputArray _descriptors, std::vector<int> &vLappingArea) {
    if (_image.empty()) return -1;
    cv::Mat image = _image.getMat();
    assert(image.type() == CV_8UC1);

    const bool rectify = mnRectWidth > 0;
    const int cap = vsg_extractor_max_keypoints(mpHandle, rectify ? mnRectWidth : image.cols, rectify ? mnRectHeight : image.rows);
    if (cap < 0) Check(cap, "vsg_extractor_max_keypoints");
    static_assert(sizeof(cv::KeyPoint) == sizeof(vsg_keypoint), "cv::KeyPoint layout");
    std::vector<cv::KeyPoint> kps(cap);
    std::vector<unsigned char> desc((size_t)cap * 32);
    int n = 0, mono = 0;
    const int lap0 = vLappingArea.size() > 0 ? vLappingArea[0] : 0, lap1 = vLappingArea.size() > 1 ? vLappingArea[1] : 0;
    if (rectify)
        Check(vsg_extract_batch_rectify(mpHandle, image.ptr(0), 1, image.cols, image.rows, (int)image.step,
                                        (size_t)image.step * image.rows, 1, lap0, lap1,
                                        reinterpret_cast<vsg_keypoint *>(kps.data()), desc.data(), cap, &n, &mono),
              "vsg_extract_batch_rectify");
    else
        Check(vsg_extract(mpHandle, image.ptr(0), image.cols, image.rows, (int)image.step, lap0, lap1,
                          reinterpret_cast<vsg_keypoint *>(kps.data()), desc.data(), cap, &n, &mono), "vsg_extract");

    if (n == 0) {
        _descriptors.release();
    } else {
        _descriptors.create(n, 32, CV_8U);
        cv::Mat d = _descriptors.getMat();
        for (int i = 0; i < n; ++i) std::memcpy(d.ptr(i), &desc[(size_t)i * 32], 32);
    }
    kps.resize(n);
    _keypoints = kps;

    if (mbDownloadPyramid) {   // mvImagePyramid as the reference leaves it (:1171-1195)
        mvPyramidStorage.resize(nlevels);
        for (int level = 0; level < nlevels; ++level) {
            int w = 0, h = 0;
            Check(vsg_pyramid_level_size(mpHandle, level, &w, &h), "vsg_pyramid_level_size");
            cv::Mat &temp = mvPyramidStorage[level];
            temp.create(h + 2 * EDGE_THRESHOLD, w + 2 * EDGE_THRESHOLD, CV_8UC1);
            mvImagePyramid[level] = temp(cv::Rect(EDGE_THRESHOLD, EDGE_THRESHOLD, w, h));
            cv::Mat &roi = mvImagePyramid[level];
            Check(vsg_pyramid_download(mpHandle, 0, level, roi.ptr(0), (int)roi.step), "vsg_pyramid_download");
            for (int y = -EDGE_THRESHOLD; y < h + EDGE_THRESHOLD; ++y) {
                const unsigned char *src = roi.ptr(0) + (ptrdiff_t)Reflect101(y, h) * (ptrdiff_t)roi.step;
                unsigned char *dst = roi.ptr(0) + (ptrdiff_t)y * (ptrdiff_t)roi.step;
                for (int x = -EDGE_THRESHOLD; x < w + EDGE_THRESHOLD; ++x) {
                    if (y >= 0 && y < h && x >= 0 && x < w) { x = w - 1; continue; }
                    dst[x] = src[Reflect101(x, w)];
                }
            }
        }
    }
    return mono;
}

}  // namespace VS_GRAPHS
