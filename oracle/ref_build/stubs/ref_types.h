// ref_types.h — STAND-INS (test infrastructure) for the classes the reference's ORBmatcher.cc takes as arguments:
// VS_GRAPHS::{GeometricCamera, Frame, KeyFrame, MapPoint}.  The real headers pull in Eigen, PCL, boost, DBoW2 and g2o,
// none of which exist in this image; these classes carry exactly the members ORBmatcher.cc reads or calls, with the
// reference's names and semantics (cited per member), so that ORBmatcher.cc compiles UNMODIFIED against them
// (oracle/ref_build/Makefile) and the drop-in shim's templates instantiate on the same types.
//
// Restated from the reference (small, and needed by ORBmatcher.cc at run time):
//   Frame::AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea   orb_slam3/src/Frame.cc:521-553, 870-880, 802-868
//   KeyFrame::GetFeaturesInArea / IsInImage                       orb_slam3/src/KeyFrame.cc:834-880
//   MapPoint::PredictScale / Get{Min,Max}DistanceInvariance       orb_slam3/src/MapPoint.cc:519-565
//   Pinhole::project / toK_ / epipolarConstrain                   orb_slam3/src/CameraModels/Pinhole.cpp:46-53,111-141
//   KannalaBrandt8::project                                       orb_slam3/src/CameraModels/KannalaBrandt8.cpp:66-84
#pragma once
#include <cmath>
#include <map>
#include <set>
#include <tuple>
#include <vector>

#include <opencv2/core/core.hpp>

#include "Eigen/Core"
#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"
#include "sophus/se3.hpp"
#include "sophus/sim3.hpp"

#ifndef FRAME_GRID_ROWS
#define FRAME_GRID_ROWS 48
#define FRAME_GRID_COLS 64
#endif

// The reference's ORBmatcher.h uses unqualified `pair` / `vector` (ORBmatcher.h:73,81): in the real tree Frame.h ->
// ORBVocabulary.h -> Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h leaks `using namespace std` into every includer; the
// stand-in has to leak it too.
using namespace std;

namespace VS_GRAPHS {

class GeometricCamera {
public:
    virtual ~GeometricCamera() {}
    virtual Eigen::Vector2f project(const Eigen::Vector3f &v3D) = 0;
    virtual Eigen::Matrix3f toK_() = 0;
    virtual Eigen::Vector3f unprojectEig(const cv::Point2f &p2D) = 0;
    virtual bool epipolarConstrain(GeometricCamera *pCamera2, const cv::KeyPoint &kp1, const cv::KeyPoint &kp2,
                                   const Eigen::Matrix3f &R12, const Eigen::Vector3f &t12, const float sigmaLevel,
                                   const float unc) = 0;
    std::vector<float> mvParameters;
};

class Pinhole : public GeometricCamera {
public:
    Pinhole(float fx, float fy, float cx, float cy) { mvParameters = {fx, fy, cx, cy}; }
    Eigen::Vector2f project(const Eigen::Vector3f &v3D) override {
        Eigen::Vector2f res;
        res[0] = mvParameters[0] * v3D[0] / v3D[2] + mvParameters[2];
        res[1] = mvParameters[1] * v3D[1] / v3D[2] + mvParameters[3];
        return res;
    }
    Eigen::Matrix3f toK_() override {
        Eigen::Matrix3f K;
        K << mvParameters[0], 0.f, mvParameters[2], 0.f, mvParameters[1], mvParameters[3], 0.f, 0.f, 1.f;
        return K;
    }
    Eigen::Vector3f unprojectEig(const cv::Point2f &p) override {
        return Eigen::Vector3f((p.x - mvParameters[2]) / mvParameters[0], (p.y - mvParameters[3]) / mvParameters[1], 1.f);
    }
    bool epipolarConstrain(GeometricCamera *pCamera2, const cv::KeyPoint &kp1, const cv::KeyPoint &kp2, const Eigen::Matrix3f &R12,
                           const Eigen::Vector3f &t12, const float, const float unc) override {
        Eigen::Matrix3f t12x = Sophus::SO3f::hat(t12);
        Eigen::Matrix3f K1 = this->toK_();
        Eigen::Matrix3f K2 = pCamera2->toK_();
        Eigen::Matrix3f F12 = K1.transpose().inverse() * t12x * R12 * K2.inverse();
        const float a = kp1.pt.x * F12(0, 0) + kp1.pt.y * F12(1, 0) + F12(2, 0);
        const float b = kp1.pt.x * F12(0, 1) + kp1.pt.y * F12(1, 1) + F12(2, 1);
        const float c = kp1.pt.x * F12(0, 2) + kp1.pt.y * F12(1, 2) + F12(2, 2);
        const float num = a * kp2.pt.x + b * kp2.pt.y + c;
        const float den = a * a + b * b;
        if (den == 0) return false;
        const float dsqr = num * num / den;
        return dsqr < 3.84 * unc;
    }
};

// Fisheye stand-in: the reference's KB8 projection; the epipolar test of the real class triangulates with an SVD
// (KannalaBrandt8.cpp:323-390), which is the caller's camera code, not the matcher's — here a deterministic angular
// test on the unprojected bearing rays stands in for it (both sides of a parity test call this same function).
class KannalaBrandt8 : public GeometricCamera {
public:
    KannalaBrandt8(float fx, float fy, float cx, float cy, float k0, float k1, float k2, float k3) { mvParameters = {fx, fy, cx, cy, k0, k1, k2, k3}; }
    Eigen::Vector2f project(const Eigen::Vector3f &v3D) override {
        const float x2_plus_y2 = v3D[0] * v3D[0] + v3D[1] * v3D[1];
        const float theta = atan2f(sqrtf(x2_plus_y2), v3D[2]);
        const float psi = atan2f(v3D[1], v3D[0]);
        const float theta2 = theta * theta, theta3 = theta * theta2, theta5 = theta3 * theta2, theta7 = theta5 * theta2,
                    theta9 = theta7 * theta2;
        const float r = theta + mvParameters[4] * theta3 + mvParameters[5] * theta5 + mvParameters[6] * theta7 + mvParameters[7] * theta9;
        Eigen::Vector2f res;
        res[0] = mvParameters[0] * r * cosf(psi) + mvParameters[2];
        res[1] = mvParameters[1] * r * sinf(psi) + mvParameters[3];
        return res;
    }
    Eigen::Matrix3f toK_() override {
        Eigen::Matrix3f K;
        K << mvParameters[0], 0.f, mvParameters[2], 0.f, mvParameters[1], mvParameters[3], 0.f, 0.f, 1.f;
        return K;
    }
    Eigen::Vector3f unprojectEig(const cv::Point2f &p) override {
        const float mx = (p.x - mvParameters[2]) / mvParameters[0], my = (p.y - mvParameters[3]) / mvParameters[1];
        const float th = sqrtf(mx * mx + my * my);
        if (th < 1e-6f) return Eigen::Vector3f(mx, my, 1.f);
        const float s = tanf(th) / th;
        return Eigen::Vector3f(mx * s, my * s, 1.f);
    }
    bool epipolarConstrain(GeometricCamera *pCamera2, const cv::KeyPoint &kp1, const cv::KeyPoint &kp2, const Eigen::Matrix3f &R12,
                           const Eigen::Vector3f &t12, const float, const float unc) override {
        Eigen::Vector3f r1 = unprojectEig(kp1.pt), r2 = pCamera2->unprojectEig(kp2.pt);
        Eigen::Vector3f r21 = R12 * r2;
        const float cosParallax = r1.dot(r21) / (r1.norm() * r21.norm());
        if (cosParallax > 0.9998f) return false;
        Eigen::Vector3f n = Sophus::SO3f::hat(t12) * r21;          // normal of the epipolar plane
        const float d = r1.dot(n) / (r1.norm() * n.norm() + 1e-12f);
        return d * d < 1e-3f * unc;
    }
};

class KeyFrame;
class Frame;

class MapPoint {
public:
    // members the tracking search reads directly (MapPoint.h:142-177)
    bool mbTrackInView = false, mbTrackInViewR = false;
    float mTrackProjX = 0, mTrackProjY = 0, mTrackDepth = 0, mTrackDepthR = 0;
    float mTrackProjXR = 0, mTrackProjYR = 0;
    int mnTrackScaleLevel = 0, mnTrackScaleLevelR = 0;
    float mTrackViewCos = 1, mTrackViewCosR = 1;

    // state behind the accessors
    int id = -1;
    bool bad = false;
    int nObs = 1;
    cv::Mat descriptor;                 // 1 x 32
    Eigen::Vector3f worldPos, normal = Eigen::Vector3f(0, 0, 1);
    float mfMinDistance = 0.01f, mfMaxDistance = 1e4f;
    std::map<KeyFrame *, std::tuple<int, int>> observations;
    MapPoint *replacedBy = nullptr;

    bool isBad() { return bad; }
    int Observations() { return nObs; }
    cv::Mat GetDescriptor() { return descriptor.clone(); }                      // MapPoint.cc:419-423 (a clone)
    Eigen::Vector3f GetWorldPos() { return worldPos; }
    Eigen::Vector3f GetNormal() { return normal; }
    float GetMinDistanceInvariance() { return 0.8f * mfMinDistance; }           // MapPoint.cc:519-523
    float GetMaxDistanceInvariance() { return 1.2f * mfMaxDistance; }           // MapPoint.cc:525-529
    int PredictScale(const float &currentDist, KeyFrame *pKF);
    int PredictScale(const float &currentDist, Frame *pF);
    bool IsInKeyFrame(KeyFrame *pKF) { return observations.count(pKF) != 0; }
    std::tuple<int, int> GetIndexInKeyFrame(KeyFrame *pKF) {
        auto it = observations.find(pKF);
        return it == observations.end() ? std::tuple<int, int>(-1, -1) : it->second;
    }
    void AddObservation(KeyFrame *pKF, int idx) {
        if (!observations.count(pKF)) ++nObs;
        observations[pKF] = std::tuple<int, int>(idx, -1);
    }
    void Replace(MapPoint *pMP) {       // the effects ORBmatcher's callers can observe: this point goes bad, pMP inherits
        if (pMP == this) return;
        bad = true;
        replacedBy = pMP;
        for (auto &kv : observations)
            if (!pMP->observations.count(kv.first)) { pMP->observations[kv.first] = kv.second; ++pMP->nObs; }
    }
};

inline long unsigned int &NextFrameId() { static long unsigned int n = 0; return n; }
class Frame {
public:
    long unsigned int mnId = NextFrameId()++;      // Frame.h:217 / KeyFrame.h:312 (copies keep it, like the reference's copy constructor)
    int N = 0;
    int Nleft = -1, Nright = -1;
    std::vector<cv::KeyPoint> mvKeys, mvKeysRight, mvKeysUn;
    std::vector<float> mvuRight, mvDepth;
    cv::Mat mDescriptors;
    std::vector<MapPoint *> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    std::vector<int> mvLeftToRightMatch, mvRightToLeftMatch;
    DBoW2::FeatureVector mFeatVec;
    std::vector<float> mvScaleFactors, mvInvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
    int mnScaleLevels = 8;
    float mfScaleFactor = 1.2f, mfLogScaleFactor = std::log(1.2f);
    float mnMinX = 0, mnMaxX = 640, mnMinY = 0, mnMaxY = 480;
    float mfGridElementWidthInv = 0.1f, mfGridElementHeightInv = 0.1f;
    float mb = 0.1f, mbf = 40.f;
    float fx = 500, fy = 500, cx = 320, cy = 240;
    GeometricCamera *mpCamera = nullptr, *mpCamera2 = nullptr;
    Sophus::SE3f mTcw, mTrl;
    std::vector<std::size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS], mGridRight[FRAME_GRID_COLS][FRAME_GRID_ROWS];

    Sophus::SE3f GetPose() const { return mTcw; }
    Sophus::SE3f GetRelativePoseTrl() { return mTrl; }

    void SetBounds(float minX, float maxX, float minY, float maxY) {
        mnMinX = minX; mnMaxX = maxX; mnMinY = minY; mnMaxY = maxY;
        mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(mnMaxX - mnMinX);     // Frame.cc:186-187
        mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(mnMaxY - mnMinY);
    }
    bool PosInGrid(const cv::KeyPoint &kp, int &posX, int &posY) {
        posX = round((kp.pt.x - mnMinX) * mfGridElementWidthInv);
        posY = round((kp.pt.y - mnMinY) * mfGridElementHeightInv);
        return !(posX < 0 || posX >= FRAME_GRID_COLS || posY < 0 || posY >= FRAME_GRID_ROWS);
    }
    void AssignFeaturesToGrid() {
        for (int i = 0; i < FRAME_GRID_COLS; ++i)
            for (int j = 0; j < FRAME_GRID_ROWS; ++j) { mGrid[i][j].clear(); mGridRight[i][j].clear(); }
        for (int i = 0; i < N; i++) {
            const cv::KeyPoint &kp = (Nleft == -1) ? mvKeysUn[i] : (i < Nleft) ? mvKeys[i] : mvKeysRight[i - Nleft];
            int gx, gy;
            if (PosInGrid(kp, gx, gy)) {
                if (Nleft == -1 || i < Nleft) mGrid[gx][gy].push_back(i);
                else mGridRight[gx][gy].push_back(i - Nleft);
            }
        }
    }
    std::vector<std::size_t> GetFeaturesInArea(const float &x, const float &y, const float &r, const int minLevel = -1,
                                               const int maxLevel = -1, const bool bRight = false) const {
        std::vector<std::size_t> vIndices;
        vIndices.reserve(N);
        const float factorX = r, factorY = r;
        const int nMinCellX = std::max(0, (int)floor((x - mnMinX - factorX) * mfGridElementWidthInv));
        if (nMinCellX >= FRAME_GRID_COLS) return vIndices;
        const int nMaxCellX = std::min((int)FRAME_GRID_COLS - 1, (int)ceil((x - mnMinX + factorX) * mfGridElementWidthInv));
        if (nMaxCellX < 0) return vIndices;
        const int nMinCellY = std::max(0, (int)floor((y - mnMinY - factorY) * mfGridElementHeightInv));
        if (nMinCellY >= FRAME_GRID_ROWS) return vIndices;
        const int nMaxCellY = std::min((int)FRAME_GRID_ROWS - 1, (int)ceil((y - mnMinY + factorY) * mfGridElementHeightInv));
        if (nMaxCellY < 0) return vIndices;
        const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
        for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
            for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
                const std::vector<std::size_t> &vCell = (!bRight) ? mGrid[ix][iy] : mGridRight[ix][iy];
                for (std::size_t j = 0, jend = vCell.size(); j < jend; j++) {
                    const cv::KeyPoint &kpUn = (Nleft == -1) ? mvKeysUn[vCell[j]] : (!bRight) ? mvKeys[vCell[j]] : mvKeysRight[vCell[j]];
                    if (bCheckLevels) {
                        if (kpUn.octave < minLevel) continue;
                        if (maxLevel >= 0 && kpUn.octave > maxLevel) continue;
                    }
                    const float distx = kpUn.pt.x - x, disty = kpUn.pt.y - y;
                    if (fabs(distx) < factorX && fabs(disty) < factorY) vIndices.push_back(vCell[j]);
                }
            }
        return vIndices;
    }
};

class KeyFrame : public Frame {
public:
    int NLeft = -1, NRight = -1;
    int mnGridCols = FRAME_GRID_COLS, mnGridRows = FRAME_GRID_ROWS;
    Sophus::SE3f mTlr;                              // left-from-right (KeyFrame::GetRightPose = Trl * Tcw, KeyFrame.cc)

    KeyFrame() {}
    explicit KeyFrame(const Frame &F) : Frame(F) { NLeft = F.Nleft; NRight = F.Nright; }

    Sophus::SE3f GetPose() { return mTcw; }
    Sophus::SE3f GetPoseInverse() { return mTcw.inverse(); }
    Eigen::Vector3f GetCameraCenter() { return mTcw.inverse().translation(); }
    Sophus::SE3f GetRightPose() { return mTrl * mTcw; }
    Sophus::SE3f GetRightPoseInverse() { return (mTrl * mTcw).inverse(); }
    Eigen::Vector3f GetRightCameraCenter() { return (mTrl * mTcw).inverse().translation(); }

    std::vector<MapPoint *> GetMapPointMatches() { return mvpMapPoints; }
    MapPoint *GetMapPoint(const std::size_t &idx) { return mvpMapPoints[idx]; }
    std::set<MapPoint *> GetMapPoints() {                                        // KeyFrame.cc:378-391
        std::set<MapPoint *> s;
        for (MapPoint *pMP : mvpMapPoints)
            if (pMP && !pMP->isBad()) s.insert(pMP);
        return s;
    }
    void AddMapPoint(MapPoint *pMP, const std::size_t &idx) { mvpMapPoints[idx] = pMP; }
    bool IsInImage(const float &x, const float &y) const { return (x >= mnMinX && x < mnMaxX && y >= mnMinY && y < mnMaxY); }

    std::vector<std::size_t> GetFeaturesInArea(const float &x, const float &y, const float &r, const bool bRight = false) const {
        std::vector<std::size_t> vIndices;
        vIndices.reserve(N);
        const float factorX = r, factorY = r;
        const int nMinCellX = std::max(0, (int)floor((x - mnMinX - factorX) * mfGridElementWidthInv));
        if (nMinCellX >= mnGridCols) return vIndices;
        const int nMaxCellX = std::min((int)mnGridCols - 1, (int)ceil((x - mnMinX + factorX) * mfGridElementWidthInv));
        if (nMaxCellX < 0) return vIndices;
        const int nMinCellY = std::max(0, (int)floor((y - mnMinY - factorY) * mfGridElementHeightInv));
        if (nMinCellY >= mnGridRows) return vIndices;
        const int nMaxCellY = std::min((int)mnGridRows - 1, (int)ceil((y - mnMinY + factorY) * mfGridElementHeightInv));
        if (nMaxCellY < 0) return vIndices;
        for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
            for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
                const std::vector<std::size_t> &vCell = (!bRight) ? mGrid[ix][iy] : mGridRight[ix][iy];
                for (std::size_t j = 0, jend = vCell.size(); j < jend; j++) {
                    const cv::KeyPoint &kpUn = (NLeft == -1) ? mvKeysUn[vCell[j]] : (!bRight) ? mvKeys[vCell[j]] : mvKeysRight[vCell[j]];
                    const float distx = kpUn.pt.x - x, disty = kpUn.pt.y - y;
                    if (fabs(distx) < r && fabs(disty) < r) vIndices.push_back(vCell[j]);
                }
            }
        return vIndices;
    }
};

inline int MapPoint::PredictScale(const float &currentDist, KeyFrame *pKF) {
    const float ratio = mfMaxDistance / currentDist;
    int nScale = ceil(log(ratio) / pKF->mfLogScaleFactor);
    if (nScale < 0) nScale = 0;
    else if (nScale >= pKF->mnScaleLevels) nScale = pKF->mnScaleLevels - 1;
    return nScale;
}
inline int MapPoint::PredictScale(const float &currentDist, Frame *pF) {
    const float ratio = mfMaxDistance / currentDist;
    int nScale = ceil(log(ratio) / pF->mfLogScaleFactor);
    if (nScale < 0) nScale = 0;
    else if (nScale >= pF->mnScaleLevels) nScale = pF->mnScaleLevels - 1;
    return nScale;
}

}  // namespace VS_GRAPHS
