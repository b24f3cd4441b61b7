// FrameOps.h — drop-in bodies for the per-keypoint steps the reference's Frame constructors run right after the
// extractor (snt-arg/visual_sgraphs, orb_slam3/src/Frame.cc):
//   Frame::UndistortKeyPoints      :891-922    cv::undistortPoints(mat, mat, mK, mDistCoef, cv::Mat(), mK)
//   Frame::ComputeImageBounds      :924-955    the same call on the four image corners
//   Frame::ComputeStereoFromRGBD   :1129-1150  one depth look-up per keypoint
//   Frame::ComputeStereoFishEyeMatches :1181-1225  BFMatcher.knnMatch(k = 2) + Lowe ratio 0.7 + the camera's triangulation gate
// The undistortion runs on the device (vsg_undistort_keypoints, bit-identical to OpenCV's double-precision path); the
// depth association is a gather of N values from a host image and stays host code.
//
// Frame is outside the hot path (it pulls in Eigen, Sophus, PCL, DBoW2), so these are free functions taking the members
// the reference's bodies read; a maintainer replaces the bodies of the three methods with one call each:
//   void Frame::UndistortKeyPoints() { VS_GRAPHS::frame_ops::UndistortKeyPoints(mvKeys, mK, mDistCoef, mvKeysUn); }
#ifndef VSG_SHIM_FRAMEOPS_H
#define VSG_SHIM_FRAMEOPS_H

#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/vsg_cuda.h"
#include "cv_compat.h"

namespace VS_GRAPHS {
namespace frame_ops {

// One matcher workspace (stream + scratch) per calling thread: Frame objects are built concurrently by Tracking and by
// the stereo constructor's two extraction threads.
inline vsg_matcher *Workspace(int device = 0) {
    struct Holder {
        vsg_matcher *m = nullptr;
        ~Holder() { vsg_matcher_destroy(m); }
    };
    static thread_local Holder h;
    if (!h.m && vsg_matcher_create(device, &h.m) != VSG_OK)
        throw std::runtime_error(std::string("vsg_matcher_create: ") + vsg_last_error());
    return h.m;
}

// mK: 3x3 CV_32F, mDistCoef: (4|5|8|12)x1 CV_32F, read with at<float> like the reference does.
template <class MatT>
inline void UndistortPoints(const MatT &mK, const MatT &mDistCoef, int n, const float *xy_in, float *xy_out) {
    double dist[12];
    const int nd = mDistCoef.rows * mDistCoef.cols;
    if (nd > 12) throw std::runtime_error("UndistortPoints: more than 12 distortion coefficients (tilt model unsupported)");
    for (int i = 0; i < nd; ++i)
        dist[i] = mDistCoef.cols == 1 ? mDistCoef.template at<float>(i, 0) : mDistCoef.template at<float>(0, i);
    const vsg_status st = vsg_undistort_keypoints(Workspace(), n, xy_in, mK.template at<float>(0, 0), mK.template at<float>(1, 1),
                                                  mK.template at<float>(0, 2), mK.template at<float>(1, 2), dist, nd, xy_out);
    if (st != VSG_OK) throw std::runtime_error(std::string("vsg_undistort_keypoints: ") + vsg_last_error());
}

// Frame::UndistortKeyPoints: mvKeysUn = mvKeys with undistorted coordinates (everything else copied, :913-921).
template <class MatT>
inline void UndistortKeyPoints(const std::vector<cv::KeyPoint> &mvKeys, const MatT &mK, const MatT &mDistCoef,
                               std::vector<cv::KeyPoint> &mvKeysUn) {
    mvKeysUn = mvKeys;
    const int n = (int)mvKeys.size();
    if (n == 0 || mDistCoef.template at<float>(0, 0) == 0.0) return;           // :893-897
    std::vector<float> xy(2 * (size_t)n);
    for (int i = 0; i < n; ++i) { xy[2 * i] = mvKeys[i].pt.x; xy[2 * i + 1] = mvKeys[i].pt.y; }
    UndistortPoints(mK, mDistCoef, n, xy.data(), xy.data());
    for (int i = 0; i < n; ++i) { mvKeysUn[i].pt.x = xy[2 * i]; mvKeysUn[i].pt.y = xy[2 * i + 1]; }
}

// Frame::ComputeImageBounds: mnMinX / mnMaxX / mnMinY / mnMaxY of an imLeft of cols x rows pixels.
template <class MatT>
inline void ComputeImageBounds(int cols, int rows, const MatT &mK, const MatT &mDistCoef, float &mnMinX, float &mnMaxX,
                               float &mnMinY, float &mnMaxY) {
    if (mDistCoef.template at<float>(0, 0) != 0.0) {
        float c[8] = {0.f, 0.f, (float)cols, 0.f, 0.f, (float)rows, (float)cols, (float)rows};
        UndistortPoints(mK, mDistCoef, 4, c, c);
        mnMinX = c[0] < c[4] ? c[0] : c[4];     // min(mat(0,0), mat(2,0))
        mnMaxX = c[2] > c[6] ? c[2] : c[6];     // max(mat(1,0), mat(3,0))
        mnMinY = c[1] < c[3] ? c[1] : c[3];     // min(mat(0,1), mat(1,1))
        mnMaxY = c[5] > c[7] ? c[5] : c[7];     // max(mat(2,1), mat(3,1))
    } else {
        mnMinX = 0.0f; mnMaxX = (float)cols; mnMinY = 0.0f; mnMaxY = (float)rows;
    }
}

// Frame::ComputeStereoFromRGBD: imDepth is CV_32F (depth_pitch_floats = imDepth.step / 4).
inline void ComputeStereoFromRGBD(const std::vector<cv::KeyPoint> &mvKeys, const std::vector<cv::KeyPoint> &mvKeysUn,
                                  const float *imDepth, size_t depth_pitch_floats, float mbf, std::vector<float> &mvuRight,
                                  std::vector<float> &mvDepth) {
    const size_t n = mvKeys.size();
    mvuRight.assign(n, -1.f);
    mvDepth.assign(n, -1.f);
    for (size_t i = 0; i < n; ++i) {
        const float d = imDepth[(size_t)(int)mvKeys[i].pt.y * depth_pitch_floats + (int)mvKeys[i].pt.x];
        if (d > 0) {
            mvDepth[i] = d;
            mvuRight[i] = mvKeysUn[i].pt.x - mbf / d;
        }
    }
}

// Frame::ComputeStereoFishEyeMatches (Frame.cc:1181-1225) for a two-camera frame F: the brute-force kNN-2 between the
// lapping-area descriptors of the left and the right image runs on the device (vsg_knn2: ties to the lower train index,
// as cv::BFMatcher orders them); Lowe's ratio test is evaluated exactly as the reference writes it (float distance against
// float * double 0.7); the parallax / reprojection gate is the frame's own camera model — KB8T = KannalaBrandt8, the class the
// reference static_casts mpCamera to (:1212) — called per surviving match in the reference's order.
template <class KB8T, class FrameT>
inline void ComputeStereoFishEyeMatches(FrameT &F) {
    typedef typename std::decay<decltype(F.mvStereo3Dpoints[0])>::type Vec3T;
    F.mvLeftToRightMatch = std::vector<int>(F.Nleft, -1);
    F.mvRightToLeftMatch = std::vector<int>(F.Nright, -1);
    F.mvDepth = std::vector<float>(F.Nleft, -1.0f);
    F.mvuRight = std::vector<float>(F.Nleft, -1);
    F.mvStereo3Dpoints = std::vector<Vec3T>(F.Nleft);
    F.mnCloseMPs = 0;
    const int nl = F.Nleft - F.monoLeft, nr = F.Nright - F.monoRight;      // stereoDescLeft / stereoDescRight (:1187-1188)
    if (nl <= 0 || nr <= 0) return;
    std::vector<uint8_t> dl((size_t)nl * 32), dr((size_t)nr * 32);
    for (int i = 0; i < nl; ++i) std::memcpy(&dl[(size_t)i * 32], F.mDescriptors.ptr(F.monoLeft + i), 32);
    for (int i = 0; i < nr; ++i) std::memcpy(&dr[(size_t)i * 32], F.mDescriptorsRight.ptr(F.monoRight + i), 32);
    std::vector<int32_t> idx((size_t)nl * 2), dist((size_t)nl * 2);
    if (vsg_knn2(Workspace(), dl.data(), nl, dr.data(), nr, 0, idx.data(), dist.data()) != VSG_OK)
        throw std::runtime_error(std::string("vsg_knn2: ") + vsg_last_error());
    for (int q = 0; q < nl; ++q) {                                          // :1206-1224
        if (idx[2 * q + 1] < 0) continue;                                   // (*it).size() >= 2
        const float d0 = (float)dist[2 * q], d1 = (float)dist[2 * q + 1];   // cv::DMatch::distance is a float
        if (!(d0 < d1 * 0.7)) continue;
        const int iL = q + F.monoLeft, iR = idx[2 * q] + F.monoRight;
        Vec3T p3D;
        const float sigma1 = F.mvLevelSigma2[F.mvKeys[iL].octave], sigma2 = F.mvLevelSigma2[F.mvKeysRight[iR].octave];
        const float depth = static_cast<KB8T *>(F.mpCamera)->TriangulateMatches(F.mpCamera2, F.mvKeys[iL], F.mvKeysRight[iR], F.mRlr,
                                                                               F.mtlr, sigma1, sigma2, p3D);
        if (depth > 0.0001f) {
            F.mvLeftToRightMatch[iL] = iR;
            F.mvRightToLeftMatch[iR] = iL;
            F.mvStereo3Dpoints[iL] = p3D;
            F.mvDepth[iL] = depth;
        }
    }
}

}  // namespace frame_ops
}  // namespace VS_GRAPHS
#endif /* VSG_SHIM_FRAMEOPS_H */
