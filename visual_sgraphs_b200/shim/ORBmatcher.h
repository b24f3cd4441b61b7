// ORBmatcher.h — drop-in for the reference's orb_slam3/include/ORBmatcher.h (snt-arg/visual_sgraphs):
// VS_GRAPHS::ORBmatcher with the same constructor, constants, static DescriptorDistance and the Search*
// methods named by the north star, forwarding to the CUDA library (include/vsg_cuda.h).
//
// The reference's methods take Frame&, KeyFrame*, MapPoint* — classes that live outside the hot path
// (they pull in Eigen, Sophus, PCL, DBoW2).  To stay source compatible without depending on those
// headers, the methods are member templates: the argument types are deduced at the reference's call
// sites (Tracking.cc:2555, 2790, 2926, 3423 ...) and only the members the reference's own
// implementation reads are touched (same names: N, Nleft, mvKeysUn, mDescriptors, mvuRight, mvpMapPoints,
// mvScaleFactors, mnMinX ..., mfGridElementWidthInv ..., mFeatVec, GetPose(), mpCamera->project(),
// MapPoint::{mbTrackInView, mTrackProjX, ..., isBad(), Observations(), GetDescriptor(), GetWorldPos()}).
// Each template flattens those members into the C ABI's view structs, calls the library and writes the
// results back exactly where the reference does.
//
// Scope: single-camera frames (Frame::Nleft == -1).  For two-camera fisheye rigs (Nleft != -1) these
// methods throw — keep the reference's CPU ORBmatcher for that configuration.
#ifndef VSG_SHIM_ORBMATCHER_H
#define VSG_SHIM_ORBMATCHER_H

#include <cstdint>
#include <cstring>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/vsg_cuda.h"
#include "cv_compat.h"

#ifndef FRAME_GRID_ROWS
#define FRAME_GRID_ROWS 48
#define FRAME_GRID_COLS 64
#endif

namespace VS_GRAPHS {

class ORBmatcher {
public:
    ORBmatcher(float nnratio = 0.6, bool checkOri = true) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}
    ~ORBmatcher() { vsg_matcher_destroy(mpWorkspace); }
    ORBmatcher(const ORBmatcher &) = delete;
    ORBmatcher &operator=(const ORBmatcher &) = delete;

    // Computes the Hamming distance between two ORB descriptors (ORBmatcher.cc:2047-2063).  Stays a cheap host
    // function: it is also called one pair at a time from MapPoint.cc:391 and Frame.cc:1032.
    static int DescriptorDistance(const cv::Mat &a, const cv::Mat &b) {
        const uint32_t *pa = a.ptr<uint32_t>(), *pb = b.ptr<uint32_t>();
        int dist = 0;
        for (int i = 0; i < 8; ++i) dist += __builtin_popcount(pa[i] ^ pb[i]);
        return dist;
    }

    // Search matches between Frame keypoints and projected MapPoints (ORBmatcher.cc:42-216). Returns #matches.
    template <class FrameT, class MapPointT>
    int SearchByProjection(FrameT &F, const std::vector<MapPointT *> &vpMapPoints, const float th = 3,
                           const bool bFarPoints = false, const float thFarPoints = 50.0f);

    // Project MapPoints tracked in the last frame into the current frame (ORBmatcher.cc:1667-1878).
    template <class FrameT>
    int SearchByProjection(FrameT &CurrentFrame, const FrameT &LastFrame, const float th, const bool bMono);

    // Brute force constrained to ORB in the same vocabulary node (ORBmatcher.cc:226-428).
    template <class KeyFrameT, class FrameT, class MapPointT>
    int SearchByBoW(KeyFrameT *pKF, FrameT &F, std::vector<MapPointT *> &vpMapPointMatches);

    // Matching for the map initialization, monocular case (ORBmatcher.cc:643-756).
    template <class FrameT>
    int SearchForInitialization(FrameT &F1, FrameT &F2, std::vector<cv::Point2f> &vbPrevMatched,
                                std::vector<int> &vnMatches12, int windowSize = 10);

public:
    static const int TH_LOW = 50;
    static const int TH_HIGH = 100;
    static const int HISTO_LENGTH = 30;

protected:
    float mfNNratio;
    bool mbCheckOrientation;

    // ---- plumbing ----
    vsg_matcher *mpWorkspace = nullptr;
    int mnDevice = 0;

    static void Check(vsg_status st, const char *what) {
        if (st != VSG_OK) throw std::runtime_error(std::string(what) + ": " + vsg_last_error());
    }
    vsg_matcher *Workspace() {
        if (!mpWorkspace) Check(vsg_matcher_create(mnDevice, &mpWorkspace), "vsg_matcher_create");
        return mpWorkspace;
    }

    // Flatten what the Search* methods read from a Frame / KeyFrame.
    struct Flat {
        std::vector<vsg_keypoint> keys;
        std::vector<uint8_t> desc;
        std::vector<float> scale;
        vsg_frame_view view;
    };
    template <class FrameT>
    static void Flatten(const FrameT &F, Flat &out) {
        static_assert(sizeof(cv::KeyPoint) == sizeof(vsg_keypoint), "cv::KeyPoint layout");
        const int n = (int)F.mvKeysUn.size();
        out.keys.resize(n);
        if (n) std::memcpy(out.keys.data(), F.mvKeysUn.data(), (size_t)n * sizeof(vsg_keypoint));
        out.desc.resize((size_t)n * 32);
        for (int i = 0; i < n; ++i) std::memcpy(&out.desc[(size_t)i * 32], F.mDescriptors.ptr(i), 32);
        out.scale.assign(F.mvScaleFactors.begin(), F.mvScaleFactors.end());
        vsg_frame_view &v = out.view;
        v.n = n;
        v.keys = out.keys.data();
        v.descriptors = out.desc.data();
        v.u_right = F.mvuRight.empty() ? nullptr : F.mvuRight.data();
        v.min_x = F.mnMinX; v.min_y = F.mnMinY; v.max_x = F.mnMaxX; v.max_y = F.mnMaxY;
        v.grid_inv_w = F.mfGridElementWidthInv; v.grid_inv_h = F.mfGridElementHeightInv;
        v.grid_cols = FRAME_GRID_COLS; v.grid_rows = FRAME_GRID_ROWS;
        v.scale_factors = out.scale.data();
        v.n_levels = (int)out.scale.size();
    }
    struct FrameGuard {   // RAII for the uploaded frame
        vsg_frame *h = nullptr;
        ~FrameGuard() { vsg_frame_destroy(h); }
    };
    template <class FrameT>
    static void RequireSingleCamera(const FrameT &F) {
        if (F.Nleft != -1)
            throw std::runtime_error("vsg ORBmatcher: two-camera frames (Nleft != -1) are not supported by the CUDA path");
    }
};

// ---------------------------------------------------------------------------------------------------
template <class FrameT, class MapPointT>
int ORBmatcher::SearchByProjection(FrameT &F, const std::vector<MapPointT *> &vpMapPoints, const float th,
                                   const bool bFarPoints, const float thFarPoints) {
    RequireSingleCamera(F);
    Flat flat;
    Flatten(F, flat);
    FrameGuard fr;
    Check(vsg_frame_create(Workspace(), &flat.view, &fr.h), "vsg_frame_create");
    const int N = flat.view.n, nMP = (int)vpMapPoints.size();
    std::vector<uint8_t> occupied(N, 0);
    for (int i = 0; i < N; ++i)
        if (F.mvpMapPoints[i] && F.mvpMapPoints[i]->Observations() > 0) occupied[i] = 1;   // :88-90
    std::vector<vsg_track_point> pts(nMP);
    std::vector<uint8_t> desc((size_t)nMP * 32, 0);
    for (int i = 0; i < nMP; ++i) {
        MapPointT *pMP = vpMapPoints[i];
        vsg_track_point &p = pts[i];
        std::memset(&p, 0, sizeof(p));
        if (!pMP->mbTrackInView) continue;            // :49-50 (mbTrackInViewR belongs to the two-camera branch)
        p.in_view = 1;
        p.proj_x = pMP->mTrackProjX; p.proj_y = pMP->mTrackProjY; p.proj_xr = pMP->mTrackProjXR;
        p.view_cos = pMP->mTrackViewCos; p.depth = pMP->mTrackDepth; p.level = pMP->mnTrackScaleLevel;
        p.bad = pMP->isBad() ? 1 : 0;
        p.blocks = pMP->Observations() > 0 ? 1 : 0;
        const cv::Mat d = pMP->GetDescriptor();
        std::memcpy(&desc[(size_t)i * 32], d.ptr(0), 32);
    }
    std::vector<int32_t> assign(N, -1);
    int nmatches = 0;
    Check(vsg_search_by_projection_map(Workspace(), fr.h, occupied.data(), nMP, pts.data(), desc.data(), th,
                                       bFarPoints ? 1 : 0, thFarPoints, mfNNratio, assign.data(), &nmatches),
          "vsg_search_by_projection_map");
    for (int i = 0; i < N; ++i)
        if (assign[i] >= 0) F.mvpMapPoints[i] = vpMapPoints[assign[i]];                    // :130
    return nmatches;
}

template <class FrameT>
int ORBmatcher::SearchByProjection(FrameT &CurrentFrame, const FrameT &LastFrame, const float th, const bool bMono) {
    RequireSingleCamera(CurrentFrame);
    RequireSingleCamera(LastFrame);
    Flat flat;
    Flatten(CurrentFrame, flat);
    FrameGuard fr;
    Check(vsg_frame_create(Workspace(), &flat.view, &fr.h), "vsg_frame_create");
    // pose arithmetic stays with the reference's Sophus / camera classes (:1677-1716)
    const auto Tcw = CurrentFrame.GetPose();
    const auto twc = Tcw.inverse().translation();
    const auto Tlw = LastFrame.GetPose();
    const auto tlc = Tlw * twc;
    const bool bForward = tlc(2) > CurrentFrame.mb && !bMono;
    const bool bBackward = -tlc(2) > CurrentFrame.mb && !bMono;
    const int nLast = LastFrame.N, N = flat.view.n;
    std::vector<vsg_proj_point> pts(nLast);
    std::vector<uint8_t> desc((size_t)nLast * 32, 0);
    for (int i = 0; i < nLast; ++i) {
        vsg_proj_point &p = pts[i];
        std::memset(&p, 0, sizeof(p));
        auto *pMP = LastFrame.mvpMapPoints[i];
        if (!pMP || LastFrame.mvbOutlier[i]) continue;
        const auto x3Dw = pMP->GetWorldPos();
        const auto x3Dc = Tcw * x3Dw;
        const float invzc = 1.0 / x3Dc(2);
        if (invzc < 0) continue;
        const auto uv = CurrentFrame.mpCamera->project(x3Dc);
        if (uv(0) < CurrentFrame.mnMinX || uv(0) > CurrentFrame.mnMaxX) continue;
        if (uv(1) < CurrentFrame.mnMinY || uv(1) > CurrentFrame.mnMaxY) continue;
        p.valid = 1;
        p.u = uv(0); p.v = uv(1);
        p.ur = uv(0) - CurrentFrame.mbf * invzc;                                           // :1747
        p.octave = LastFrame.mvKeys[i].octave;
        p.angle = LastFrame.mvKeysUn[i].angle;
        p.blocks = pMP->Observations() > 0 ? 1 : 0;
        const cv::Mat d = pMP->GetDescriptor();
        std::memcpy(&desc[(size_t)i * 32], d.ptr(0), 32);
    }
    std::vector<uint8_t> occupied(N, 0);
    for (int i = 0; i < N; ++i)
        if (CurrentFrame.mvpMapPoints[i] && CurrentFrame.mvpMapPoints[i]->Observations() > 0) occupied[i] = 1;
    std::vector<int32_t> assign(N, -1);
    int nmatches = 0;
    Check(vsg_search_by_projection_last(Workspace(), fr.h, occupied.data(), nLast, pts.data(), desc.data(), th,
                                        bForward ? 1 : (bBackward ? 2 : 0), mbCheckOrientation ? 1 : 0, assign.data(),
                                        &nmatches), "vsg_search_by_projection_last");
    for (int i = 0; i < N; ++i) {
        if (assign[i] >= 0) CurrentFrame.mvpMapPoints[i] = LastFrame.mvpMapPoints[assign[i]];   // :1763
        else if (assign[i] == -2) CurrentFrame.mvpMapPoints[i] = nullptr;                     // :1870
    }
    return nmatches;
}

template <class KeyFrameT, class FrameT, class MapPointT>
int ORBmatcher::SearchByBoW(KeyFrameT *pKF, FrameT &F, std::vector<MapPointT *> &vpMapPointMatches) {
    RequireSingleCamera(F);
    const std::vector<MapPointT *> vpMapPointsKF = pKF->GetMapPointMatches();
    vpMapPointMatches = std::vector<MapPointT *>(F.N, static_cast<MapPointT *>(nullptr));
    Flat kf, fr;
    Flatten(*pKF, kf);
    Flatten(F, fr);
    std::vector<uint8_t> valid(kf.view.n, 0);
    for (int i = 0; i < kf.view.n; ++i)
        if (vpMapPointsKF[i] && !vpMapPointsKF[i]->isBad()) valid[i] = 1;                    // :262-266
    auto flatten_fv = [](const auto &fv, std::vector<int32_t> &nodes, std::vector<int32_t> &ptr, std::vector<int32_t> &idx) {
        ptr.push_back(0);
        for (const auto &kv : fv) {           // DBoW2::FeatureVector = std::map<NodeId, std::vector<unsigned>>
            nodes.push_back((int32_t)kv.first);
            for (unsigned v : kv.second) idx.push_back((int32_t)v);
            ptr.push_back((int32_t)idx.size());
        }
    };
    std::vector<int32_t> kn, kp, ki, fn, fp, fi;
    flatten_fv(pKF->mFeatVec, kn, kp, ki);
    flatten_fv(F.mFeatVec, fn, fp, fi);
    std::vector<int32_t> matches(fr.view.n, -1);
    int nmatches = 0;
    Check(vsg_search_by_bow(Workspace(), &kf.view, valid.data(), &fr.view, (int)kn.size(), kn.data(), kp.data(), ki.data(),
                            (int)fn.size(), fn.data(), fp.data(), fi.data(), mfNNratio, mbCheckOrientation ? 1 : 0,
                            matches.data(), &nmatches), "vsg_search_by_bow");
    for (int j = 0; j < fr.view.n; ++j)
        if (matches[j] >= 0) vpMapPointMatches[j] = vpMapPointsKF[matches[j]];               // :339
    return nmatches;
}

template <class FrameT>
int ORBmatcher::SearchForInitialization(FrameT &F1, FrameT &F2, std::vector<cv::Point2f> &vbPrevMatched,
                                        std::vector<int> &vnMatches12, int windowSize) {
    Flat f1, f2;
    Flatten(F1, f1);
    Flatten(F2, f2);
    FrameGuard fr2;
    Check(vsg_frame_create(Workspace(), &f2.view, &fr2.h), "vsg_frame_create");
    static_assert(sizeof(cv::Point2f) == 2 * sizeof(float), "cv::Point2f layout");
    vnMatches12 = std::vector<int>(F1.mvKeysUn.size(), -1);
    int nmatches = 0;
    Check(vsg_search_for_initialization(Workspace(), &f1.view, fr2.h, reinterpret_cast<float *>(vbPrevMatched.data()),
                                        windowSize, mfNNratio, mbCheckOrientation ? 1 : 0, vnMatches12.data(), &nmatches),
          "vsg_search_for_initialization");
    return nmatches;
}

}  // namespace VS_GRAPHS

namespace ORB_SLAM3 {
using VS_GRAPHS::ORBmatcher;
}

#endif
