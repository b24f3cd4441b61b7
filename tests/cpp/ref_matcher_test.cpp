// ref_matcher_test.cpp — the reference's own ORBmatcher.cc (compiled UNMODIFIED into oracle/_ref/libvsg_ref.so against the
// stand-in Frame / KeyFrame / MapPoint classes of oracle/ref_build/stubs/ref_types.h) versus the drop-in shim
// (visual_sgraphs_b200/shim/ORBmatcher.h) instantiated on the SAME stand-in classes, method by method, on two identical
// copies of a synthetic world (two related frames extracted by the reference's ORBextractor, map points with 3-D
// positions, poses, cameras, feature vectors).
//
// Two builds of this one file (oracle/ref_build/Makefile):
//   ref_matcher_test_cpu  links tests/cpp/vsg_on_oracle.cpp: shim host code + CPU oracle port  == reference   (no GPU)
//   ref_matcher_test_gpu  links libvsg_cuda.so:              shim host code + CUDA kernels      == reference   (B200)
//
//   usage: ref_matcher_test_{cpu,gpu} frame_a.raw frame_b.raw      (two 640x480 8-bit frames; b = a shifted by (9, 5))
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <random>
#include <set>
#include <string>
#include <vector>

#include "ORBmatcher.h"     // oracle/ref_build/stubs/ORBmatcher.h -> stand-ins -> the reference's ORBmatcher.h
#include "ORBextractor.h"   // the reference's ORBextractor.h

// The shim declares the same class name in the same namespace (it is a drop-in); move it aside for this one program.
#define VS_GRAPHS VSG_SHIM
#define VSG_HAVE_OPENCV 1
#include "../../visual_sgraphs_b200/shim/ORBmatcher.h"
#include "../../visual_sgraphs_b200/shim/FrameOps.h"
#undef VS_GRAPHS

using VS_GRAPHS::Frame;
using VS_GRAPHS::KeyFrame;
using VS_GRAPHS::MapPoint;
typedef VS_GRAPHS::ORBmatcher RefMatcher;
typedef VSG_SHIM::ORBmatcher ShimMatcher;

static int g_fail = 0, g_checks = 0;
static void report(const char *what, int a, int b) { std::printf("  %-58s reference %5d   shim %5d\n", what, a, b); }
#define EXPECT(cond, ...)                                                                                       \
    do {                                                                                                        \
        ++g_checks;                                                                                             \
        if (!(cond)) { std::printf("FAIL %s:%d: ", __FILE__, __LINE__); std::printf(__VA_ARGS__); std::printf("\n"); ++g_fail; } \
    } while (0)

struct Extracted {
    std::vector<cv::KeyPoint> keys;
    cv::Mat desc;
    std::vector<float> scale, inv_scale, sigma2, inv_sigma2;
};

static std::vector<unsigned char> read_raw(const char *path, size_t n) {
    std::vector<unsigned char> b(n);
    FILE *f = std::fopen(path, "rb");
    if (!f || std::fread(b.data(), 1, n, f) != n) { std::printf("cannot read %s\n", path); std::exit(2); }
    std::fclose(f);
    return b;
}

static Extracted extract(std::vector<unsigned char> &img, int nfeatures) {
    VS_GRAPHS::ORBextractor ex(nfeatures, 1.2f, 8, 20, 7);   // the reference's extractor (libvsg_ref.so)
    cv::Mat m(480, 640, CV_8UC1, img.data(), 640);
    std::vector<int> lap = {0, 0};
    Extracted e;
    ex(m, cv::Mat(), e.keys, e.desc, lap);
    e.scale = ex.GetScaleFactors(); e.inv_scale = ex.GetInverseScaleFactors();
    e.sigma2 = ex.GetScaleSigmaSquares(); e.inv_sigma2 = ex.GetInverseScaleSigmaSquares();
    return e;
}

static Eigen::Matrix3f rot_xyz(float ax, float ay, float az) {
    Eigen::Matrix3f Rx, Ry, Rz;
    Rx << 1.f, 0.f, 0.f, 0.f, std::cos(ax), -std::sin(ax), 0.f, std::sin(ax), std::cos(ax);
    Ry << std::cos(ay), 0.f, std::sin(ay), 0.f, 1.f, 0.f, -std::sin(ay), 0.f, std::cos(ay);
    Rz << std::cos(az), -std::sin(az), 0.f, std::sin(az), std::cos(az), 0.f, 0.f, 0.f, 1.f;
    return Rz * Ry * Rx;
}

// ---- the synthetic world -------------------------------------------------------------------------------------
struct Options {
    unsigned seed = 1;
    bool two_cameras = false;      // frames / keyframes with Nleft != -1 (left = image a's keypoints, right = image b's)
    float dz = 0.f;                // extra z translation of B (forward / backward motion)
    bool stereo = true;            // mvuRight set on 70 % of the keypoints
    float keep_mp = 0.85f;         // fraction of keypoints that carry a map point
};

struct World {
    VS_GRAPHS::Pinhole cam{500.f, 500.f, 320.f, 240.f};
    VS_GRAPHS::KannalaBrandt8 fish{500.f, 500.f, 320.f, 240.f, 0.01f, -0.002f, 0.0005f, -0.0001f};
    VS_GRAPHS::KannalaBrandt8 fish2{498.f, 501.f, 322.f, 239.f, 0.012f, -0.002f, 0.0004f, -0.0001f};
    std::vector<std::unique_ptr<MapPoint>> mps;     // [0, nA): seen from A's keypoints, [nA, nA + nB): from B's
    Frame A, B;
    std::unique_ptr<KeyFrame> KA, KB;
    int nA = 0, nB = 0;
    MapPoint *mp(int i) { return mps[i].get(); }
};

static void fill_frame(Frame &F, const Extracted &e, std::mt19937 &rng, const Options &o, VS_GRAPHS::GeometricCamera *cam) {
    F.N = (int)e.keys.size();
    F.mvKeys = e.keys; F.mvKeysUn = e.keys;
    F.mDescriptors = e.desc.clone();
    F.mvScaleFactors = e.scale; F.mvInvScaleFactors = e.inv_scale; F.mvLevelSigma2 = e.sigma2; F.mvInvLevelSigma2 = e.inv_sigma2;
    F.mnScaleLevels = 8; F.mfScaleFactor = 1.2f; F.mfLogScaleFactor = std::log(1.2f);
    F.mvpMapPoints.assign(F.N, nullptr);
    F.mvbOutlier.assign(F.N, false);
    F.mvuRight.assign(F.N, -1.f);
    if (o.stereo)
        for (int i = 0; i < F.N; ++i)
            if (rng() % 10 < 7) F.mvuRight[i] = F.mvKeys[i].pt.x - (2 + (rng() % 380) / 10.f);
    F.SetBounds(0.f, 640.f, 0.f, 480.f);
    F.mpCamera = cam;
    for (int i = 0; i < F.N; ++i) F.mFeatVec.addFeature((F.mDescriptors.ptr(i)[0] * 7u + 3u) % 24u, (unsigned)i);
    F.AssignFeaturesToGrid();
}

// left camera = keypoints of `l`, right camera = keypoints of `r` (Frame.cc fisheye-stereo constructor layout)
static void fill_frame_two_cameras(Frame &F, const Extracted &l, const Extracted &r, World &w) {
    const int nL = (int)l.keys.size(), nR = (int)r.keys.size();
    F.Nleft = nL; F.Nright = nR; F.N = nL + nR;
    F.mvKeys = l.keys; F.mvKeysRight = r.keys; F.mvKeysUn = l.keys;
    F.mDescriptors = cv::Mat(F.N, 32, CV_8UC1);
    for (int i = 0; i < nL; ++i) std::memcpy(F.mDescriptors.ptr(i), l.desc.ptr(i), 32);
    for (int i = 0; i < nR; ++i) std::memcpy(F.mDescriptors.ptr(nL + i), r.desc.ptr(i), 32);
    F.mvScaleFactors = l.scale; F.mvInvScaleFactors = l.inv_scale; F.mvLevelSigma2 = l.sigma2; F.mvInvLevelSigma2 = l.inv_sigma2;
    F.mvpMapPoints.assign(F.N, nullptr);
    F.mvbOutlier.assign(F.N, false);
    F.mvuRight.assign(F.N, -1.f);
    F.mvLeftToRightMatch.assign(nL, -1);
    F.mvRightToLeftMatch.assign(nR, -1);
    F.SetBounds(0.f, 640.f, 0.f, 480.f);
    F.mpCamera = &w.cam; F.mpCamera2 = &w.fish2;   // the projections of the tracking calls use mpCamera only (:1704,1788)
    F.mFeatVec.clear();
    for (int i = 0; i < F.N; ++i) F.mFeatVec.addFeature((F.mDescriptors.ptr(i)[0] * 7u + 3u) % 24u, (unsigned)i);
    F.AssignFeaturesToGrid();
}

static void build_world(World &w, const Extracted &ea, const Extracted &eb, const Options &o) {
    std::mt19937 rng(o.seed);
    auto uni = [&](float a, float b) { return a + (b - a) * (float)(rng() % 100000) / 100000.f; };
    fill_frame(w.A, ea, rng, o, &w.cam);
    fill_frame(w.B, eb, rng, o, &w.cam);
    w.nA = w.A.N; w.nB = w.B.N;
    // poses: A anywhere; B = A moved so that a point at depth ~4 shifts by (+9, +5) px (frame b is frame a rolled by (9, 5))
    const Sophus::SE3f TA(rot_xyz(0.21f, -0.33f, 0.12f), Eigen::Vector3f(0.4f, -0.7f, 1.3f));
    const Sophus::SE3f Tba(rot_xyz(0.0007f, -0.0011f, 0.0005f), Eigen::Vector3f(9.f * 4.f / 500.f, 5.f * 4.f / 500.f, o.dz));
    w.A.mTcw = TA;
    w.B.mTcw = Tba * TA;
    const Sophus::SE3f TAinv = TA.inverse(), TBinv = w.B.mTcw.inverse();
    auto make_points = [&](Frame &F, const Sophus::SE3f &Twc, int id0) {
        const Eigen::Vector3f Ow = Twc.translation();
        for (int i = 0; i < F.N; ++i) {
            std::unique_ptr<MapPoint> p(new MapPoint);
            p->id = id0 + i;
            const cv::KeyPoint &kp = F.mvKeysUn[i];
            const float z = 4.f * (1.f + uni(-0.01f, 0.01f));
            const Eigen::Vector3f Xc((kp.pt.x - 320.f) * z / 500.f, (kp.pt.y - 240.f) * z / 500.f, z);
            p->worldPos = Twc * Xc;
            const Eigen::Vector3f PO = p->worldPos - Ow;
            const float dist = PO.norm();
            p->normal = PO / dist;
            p->mfMaxDistance = dist * F.mvScaleFactors[kp.octave];
            p->mfMinDistance = p->mfMaxDistance / F.mvScaleFactors[7];
            p->descriptor = F.mDescriptors.row(i).clone();
            p->bad = rng() % 33 == 0;
            p->nObs = rng() % 20 == 0 ? 0 : 1 + (int)(rng() % 5);
            w.mps.push_back(std::move(p));
        }
    };
    make_points(w.A, TAinv, 0);
    make_points(w.B, TBinv, w.nA);
    for (int i = 0; i < w.nA; ++i)
        if (uni(0, 1) < o.keep_mp) w.A.mvpMapPoints[i] = w.mp(i);
    for (int i = 0; i < w.nB; ++i)
        if (uni(0, 1) < o.keep_mp * 0.6f) w.B.mvpMapPoints[i] = w.mp(w.nA + i);
    for (int i = 0; i < w.nA; ++i) w.A.mvbOutlier[i] = rng() % 25 == 0;
    if (o.two_cameras) {
        // A and B become two-camera frames: left = their own keypoints, right = the other image's keypoints
        Frame A1 = w.A, B1 = w.B;
        fill_frame_two_cameras(w.A, ea, eb, w);
        fill_frame_two_cameras(w.B, eb, ea, w);
        w.A.mTcw = A1.mTcw; w.B.mTcw = B1.mTcw;
        w.A.mTrl = Sophus::SE3f(rot_xyz(0.0004f, 0.0006f, -0.0003f), Eigen::Vector3f(9.f * 4.f / 500.f, 5.f * 4.f / 500.f, 0.f));
        w.B.mTrl = Sophus::SE3f(rot_xyz(-0.0004f, 0.0003f, 0.0002f), Eigen::Vector3f(-9.f * 4.f / 500.f, -5.f * 4.f / 500.f, 0.f));
        for (int i = 0; i < w.nA; ++i) w.A.mvpMapPoints[i] = A1.mvpMapPoints[i];            // left features keep their points
        for (int i = 0; i < w.nB; ++i) w.B.mvpMapPoints[i] = B1.mvpMapPoints[i];
        for (int i = 0; i < w.nB; ++i)                                                       // right features of A: some of B's points
            if (rng() % 3 == 0) w.A.mvpMapPoints[w.nA + i] = w.mp(w.nA + i);
        for (int i = 0; i < w.nA; ++i)
            if (rng() % 3 == 0) w.B.mvpMapPoints[w.nB + i] = w.mp(i);
        w.A.mvbOutlier.assign(w.A.N, false);
        for (int i = 0; i < w.A.N; ++i) w.A.mvbOutlier[i] = rng() % 25 == 0;
        // left <-> right pairings on a third of the keypoints (Frame::ComputeStereoFishEyeMatches output)
        for (Frame *F : {&w.A, &w.B}) {
            const int nL = F->Nleft, nR = F->Nright, pairs = std::min(nL, nR) / 3;
            std::vector<int> li(nL), ri(nR);
            for (int i = 0; i < nL; ++i) li[i] = i;
            for (int i = 0; i < nR; ++i) ri[i] = i;
            std::shuffle(li.begin(), li.end(), rng);
            std::shuffle(ri.begin(), ri.end(), rng);
            for (int k = 0; k < pairs; ++k) { F->mvLeftToRightMatch[li[k]] = ri[k]; F->mvRightToLeftMatch[ri[k]] = li[k]; }
        }
    }
    w.KA.reset(new KeyFrame(w.A));
    w.KB.reset(new KeyFrame(w.B));
    w.KA->mTlr = w.A.mTrl.inverse(); w.KB->mTlr = w.B.mTrl.inverse();
    for (int i = 0; i < w.KA->N; ++i)
        if (w.KA->mvpMapPoints[i]) w.KA->mvpMapPoints[i]->observations[w.KA.get()] = std::tuple<int, int>(i, -1);
    for (int i = 0; i < w.KB->N; ++i)
        if (w.KB->mvpMapPoints[i]) w.KB->mvpMapPoints[i]->observations[w.KB.get()] = std::tuple<int, int>(i, -1);
}

// ---- comparison of the two worlds --------------------------------------------------------------------------------
static std::vector<int> ids(const std::vector<MapPoint *> &v) {
    std::vector<int> r(v.size());
    for (size_t i = 0; i < v.size(); ++i) r[i] = v[i] ? v[i]->id : -1;
    return r;
}
static int count_set(const std::vector<int> &v) { int n = 0; for (int x : v) n += x >= 0; return n; }

static void compare_worlds(World &a, World &b, const char *what) {
    EXPECT(ids(a.A.mvpMapPoints) == ids(b.A.mvpMapPoints), "%s: A.mvpMapPoints differ", what);
    EXPECT(ids(a.B.mvpMapPoints) == ids(b.B.mvpMapPoints), "%s: B.mvpMapPoints differ", what);
    EXPECT(ids(a.KA->mvpMapPoints) == ids(b.KA->mvpMapPoints), "%s: KA.mvpMapPoints differ", what);
    EXPECT(ids(a.KB->mvpMapPoints) == ids(b.KB->mvpMapPoints), "%s: KB.mvpMapPoints differ", what);
    int diff = 0;
    for (size_t i = 0; i < a.mps.size(); ++i) {
        MapPoint &p = *a.mps[i], &q = *b.mps[i];
        bool same = p.bad == q.bad && p.nObs == q.nObs && (p.replacedBy ? p.replacedBy->id : -1) == (q.replacedBy ? q.replacedBy->id : -1) &&
                    p.observations.size() == q.observations.size();
        for (int k = 0; k < 2 && same; ++k) {
            KeyFrame *ka = k ? a.KB.get() : a.KA.get(), *kb = k ? b.KB.get() : b.KA.get();
            same = p.IsInKeyFrame(ka) == q.IsInKeyFrame(kb) && p.GetIndexInKeyFrame(ka) == q.GetIndexInKeyFrame(kb);
        }
        diff += !same;
    }
    EXPECT(diff == 0, "%s: %d map points differ in bad / observations / replacement", what, diff);
}

struct Pair {
    World r, s;   // r: handed to the reference, s: handed to the shim
    Pair(const Extracted &ea, const Extracted &eb, const Options &o) { build_world(r, ea, eb, o); build_world(s, ea, eb, o); }
};

// Frame::isInFrustum's outputs for SearchByProjection(F, vpMapPoints) (Frame.cc:656-732), computed identically in both worlds
static void set_track_fields(World &w, Frame &F, std::mt19937 &rng, bool two_cameras) {
    const Sophus::SE3f Tcw = F.GetPose();
    const Eigen::Vector3f Ow = Tcw.inverse().translation();
    for (auto &up : w.mps) {
        MapPoint *p = up.get();
        const Eigen::Vector3f Pc = Tcw * p->worldPos;
        const Eigen::Vector2f uv = F.mpCamera->project(Pc);
        const float dist = (p->worldPos - Ow).norm();
        p->mbTrackInView = Pc(2) > 0 && uv(0) >= 0 && uv(0) < 640 && uv(1) >= 0 && uv(1) < 480 && rng() % 10 < 9;
        p->mTrackProjX = uv(0) + ((int)(rng() % 5) - 2) * 0.5f;
        p->mTrackProjY = uv(1) + ((int)(rng() % 5) - 2) * 0.5f;
        p->mTrackDepth = Pc(2) * (rng() % 4 == 0 ? 20.f : 1.f);     // some beyond thFarPoints
        p->mTrackProjXR = p->mTrackProjX - F.mbf / Pc(2);
        p->mnTrackScaleLevel = p->PredictScale(dist, &F);
        p->mTrackViewCos = 0.99f + (rng() % 100) / 10000.f;
        if (two_cameras) {
            const Eigen::Vector3f Pr = F.GetRelativePoseTrl() * Pc;
            const Eigen::Vector2f uvr = F.mpCamera->project(Pr);
            p->mbTrackInViewR = rng() % 10 < 7;
            p->mbTrackInView = p->mbTrackInView && rng() % 10 < 8;
            p->mTrackProjXR = uvr(0) + ((int)(rng() % 5) - 2) * 0.5f;
            p->mTrackProjYR = uvr(1) + ((int)(rng() % 5) - 2) * 0.5f;
            p->mnTrackScaleLevelR = rng() % 20 == 0 ? -1 : p->mnTrackScaleLevel;
            p->mTrackViewCosR = 0.99f + (rng() % 100) / 10000.f;
        }
    }
}

template <class V>
static std::vector<V *> raw(std::vector<std::unique_ptr<V>> &v, int b, int e) {
    std::vector<V *> r;
    for (int i = b; i < e; ++i) r.push_back(v[i].get());
    return r;
}

int main(int argc, char **argv) {
    if (argc < 3) { std::printf("usage: %s frame_a.raw frame_b.raw\n", argv[0]); return 2; }
    std::vector<unsigned char> img_a = read_raw(argv[1], 640 * 480), img_b = read_raw(argv[2], 640 * 480);
    const Extracted ea = extract(img_a, 1000), eb = extract(img_b, 1000);
    const Extracted ia = extract(img_a, 2500), ib = extract(img_b, 2500);   // mpIniORBextractor-sized frames
    std::printf("frames: %zu / %zu keypoints (init: %zu / %zu)\n", ea.keys.size(), eb.keys.size(), ia.keys.size(), ib.keys.size());

    // ---- DescriptorDistance + constants ----
    for (int i = 0; i < 200; ++i) {
        const cv::Mat a = ea.desc.row(i), b = eb.desc.row((i * 7) % eb.desc.rows);
        EXPECT(RefMatcher::DescriptorDistance(a, b) == ShimMatcher::DescriptorDistance(a, b), "DescriptorDistance %d", i);
    }
    EXPECT(RefMatcher::TH_LOW == ShimMatcher::TH_LOW && RefMatcher::TH_HIGH == ShimMatcher::TH_HIGH &&
               RefMatcher::HISTO_LENGTH == ShimMatcher::HISTO_LENGTH, "constants");

    // ---- SearchByProjection(Frame&, vector<MapPoint*>&, th, bFarPoints, thFarPoints)  :42-216 ----
    for (int two = 0; two < 2; ++two)
        for (int variant = 0; variant < 3; ++variant) {
            Options o; o.seed = 10 + variant; o.two_cameras = two;
            Pair P(ea, eb, o);
            const float th = variant == 0 ? 3.f : (variant == 1 ? 1.f : 5.f), nnr = variant == 2 ? 0.6f : 0.8f;
            const bool far = variant != 1;
            int got[2];
            for (int k = 0; k < 2; ++k) {
                World &w = k ? P.s : P.r;
                std::mt19937 rng(99 + variant);
                set_track_fields(w, w.B, rng, two);
                std::vector<MapPoint *> vp = raw(w.mps, 0, w.nA);
                std::shuffle(vp.begin(), vp.end(), rng);                    // the local map is not ordered like keypoints
                if (k == 0) got[0] = RefMatcher(nnr).SearchByProjection(w.B, vp, th, far, 45.f);
                else { ShimMatcher m(nnr); got[1] = m.SearchByProjection(w.B, vp, th, far, 45.f); }
            }
            char what[96]; std::snprintf(what, sizeof what, "SearchByProjection(F, MPs) two=%d variant=%d", two, variant);
            report(what, got[0], got[1]);
            EXPECT(got[0] == got[1] && got[0] > 100, "%s: %d vs %d", what, got[0], got[1]);
            compare_worlds(P.r, P.s, what);
        }

    // ---- SearchByProjection(Frame& Cur, const Frame& Last, th, bMono)  :1667-1878 ----
    for (int two = 0; two < 2; ++two)
        for (int variant = 0; variant < 4; ++variant) {
            Options o; o.seed = 20 + variant; o.two_cameras = two;
            o.dz = variant == 1 ? -0.3f : (variant == 2 ? 0.3f : 0.f);      // tlc(2) vs mb: forward / backward windows
            Pair P(ea, eb, o);
            const bool mono = variant == 3, ori = variant != 2;
            int got[2];
            for (int k = 0; k < 2; ++k) {
                World &w = k ? P.s : P.r;
                for (int i = 0; i < w.B.N; ++i) if (i % 9 != 0) w.B.mvpMapPoints[i] = nullptr;   // a few already-assigned keypoints
                if (k == 0) got[0] = RefMatcher(0.9f, ori).SearchByProjection(w.B, w.A, two ? 12.f : 7.f, mono);
                else { ShimMatcher m(0.9f, ori); got[1] = m.SearchByProjection(w.B, w.A, two ? 12.f : 7.f, mono); }
            }
            char what[96]; std::snprintf(what, sizeof what, "SearchByProjection(Cur, Last) two=%d variant=%d", two, variant);
            report(what, got[0], got[1]);
            EXPECT(got[0] == got[1] && got[0] > 100, "%s: %d vs %d", what, got[0], got[1]);
            compare_worlds(P.r, P.s, what);
        }

    // ---- one Track(): three searches on the SAME current frame — TrackWithMotionModel's search, its retry with twice the
    // window after clearing the matches (Tracking.cc:2916-2935), then SearchLocalPoints' SearchByProjection(F, MPs).  The shim
    // serves the second and third call from its per-thread frame cache; every call must still equal the reference's. ----
    for (int two = 0; two < 2; ++two) {
        Options o; o.seed = 25; o.two_cameras = two;
        Pair P(ea, eb, o);
        int got[2][3];
        for (int k = 0; k < 2; ++k) {
            World &w = k ? P.s : P.r;
            std::mt19937 rng(7);
            for (int i = 0; i < w.B.N; ++i) if (i % 9 != 0) w.B.mvpMapPoints[i] = nullptr;
            const float th = two ? 12.f : 7.f;
            if (k == 0) got[0][0] = RefMatcher(0.9f, true).SearchByProjection(w.B, w.A, th, false);
            else { ShimMatcher m(0.9f, true); got[1][0] = m.SearchByProjection(w.B, w.A, th, false); }
            std::fill(w.B.mvpMapPoints.begin(), w.B.mvpMapPoints.end(), static_cast<MapPoint *>(nullptr));
            if (k == 0) got[0][1] = RefMatcher(0.9f, true).SearchByProjection(w.B, w.A, 2 * th, false);
            else { ShimMatcher m(0.9f, true); got[1][1] = m.SearchByProjection(w.B, w.A, 2 * th, false); }
            set_track_fields(w, w.B, rng, two);
            std::vector<MapPoint *> vp = raw(w.mps, 0, w.nA);
            std::shuffle(vp.begin(), vp.end(), rng);
            if (k == 0) got[0][2] = RefMatcher(0.8f).SearchByProjection(w.B, vp, 3.f, true, 45.f);
            else { ShimMatcher m(0.8f); got[1][2] = m.SearchByProjection(w.B, vp, 3.f, true, 45.f); }
        }
        char what[96]; std::snprintf(what, sizeof what, "one Track() on a cached frame two=%d", two);
        report(what, got[0][0] + got[0][1] + got[0][2], got[1][0] + got[1][1] + got[1][2]);
        for (int c = 0; c < 3; ++c) EXPECT(got[0][c] == got[1][c] && got[0][c] > 20, "%s: call %d: %d vs %d", what, c, got[0][c], got[1][c]);
        compare_worlds(P.r, P.s, what);
    }

    // ---- SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches)  :226-428 ----
    for (int two = 0; two < 2; ++two)
        for (int variant = 0; variant < 2; ++variant) {
            Options o; o.seed = 30 + variant; o.two_cameras = two;
            Pair P(ea, eb, o);
            std::vector<MapPoint *> m0, m1;
            const float nnr = variant ? 0.9f : 0.7f;
            const int g0 = RefMatcher(nnr, variant == 0).SearchByBoW(P.r.KA.get(), P.r.B, m0);
            ShimMatcher sm(nnr, variant == 0);
            const int g1 = sm.SearchByBoW(P.s.KA.get(), P.s.B, m1);
            char what[96]; std::snprintf(what, sizeof what, "SearchByBoW(KF, F) two=%d variant=%d", two, variant);
            report(what, g0, g1);
            EXPECT(g0 == g1 && g0 > 50, "%s: %d vs %d", what, g0, g1);
            EXPECT(ids(m0) == ids(m1) && (int)m0.size() == P.r.B.N, "%s: vpMapPointMatches differ", what);
            compare_worlds(P.r, P.s, what);
        }

    // ---- SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize)  :643-756 ----
    for (int variant = 0; variant < 3; ++variant) {
        Options o; o.seed = 40 + variant; o.stereo = false;
        Pair P(ia, ib, o);
        std::vector<cv::Point2f> prev[2];
        std::vector<int> m12[2];
        int got[2];
        for (int k = 0; k < 2; ++k) {
            World &w = k ? P.s : P.r;
            prev[k].resize(w.A.mvKeysUn.size());
            for (size_t i = 0; i < prev[k].size(); ++i) prev[k][i] = w.A.mvKeysUn[i].pt;          // Tracking.cc:2514-2516
            const int win = variant == 0 ? 100 : (variant == 1 ? 30 : 12);
            if (k == 0) got[0] = RefMatcher(0.9f, variant != 2).SearchForInitialization(w.A, w.B, prev[0], m12[0], win);
            else { ShimMatcher m(0.9f, variant != 2); got[1] = m.SearchForInitialization(w.A, w.B, prev[1], m12[1], win); }
        }
        report("SearchForInitialization", got[0], got[1]);
        EXPECT(got[0] == got[1] && got[0] > 100, "SearchForInitialization variant %d: %d vs %d", variant, got[0], got[1]);
        EXPECT(m12[0] == m12[1], "SearchForInitialization variant %d: vnMatches12 differ", variant);
        bool same = prev[0].size() == prev[1].size();
        for (size_t i = 0; same && i < prev[0].size(); ++i) same = prev[0][i].x == prev[1][i].x && prev[0][i].y == prev[1][i].y;
        EXPECT(same, "SearchForInitialization variant %d: vbPrevMatched differ", variant);
    }

    // ---- SearchByProjection(Frame& Cur, KeyFrame*, sAlreadyFound, th, ORBdist)  :1880-2000 ----
    for (int two = 0; two < 2; ++two)
    for (int variant = 0; variant < 2; ++variant) {
        Options o; o.seed = 50 + variant; o.two_cameras = two;
        Pair P(ea, eb, o);
        int got[2];
        for (int k = 0; k < 2; ++k) {
            World &w = k ? P.s : P.r;
            // two cameras: the method has no code of its own for them; it reads pKF->mvKeysUn[i] for every map point slot i, so
            // only the left camera's slots may hold points (anything else is out of bounds in the reference itself)
            if (two) for (int i = w.KA->NLeft; i < w.KA->N; ++i) w.KA->mvpMapPoints[i] = nullptr;
            std::set<MapPoint *> found;
            for (int i = 0; i < w.B.N; ++i) if (i % 4 != 0) w.B.mvpMapPoints[i] = nullptr;
            for (int i = 0; i < w.nA; i += 6) found.insert(w.mp(i));
            if (k == 0) got[0] = RefMatcher(0.9f, variant == 0).SearchByProjection(w.B, w.KA.get(), found, variant ? 10.f : 3.f, variant ? 100 : 64);
            else { ShimMatcher m(0.9f, variant == 0); got[1] = m.SearchByProjection(w.B, w.KA.get(), found, variant ? 10.f : 3.f, variant ? 100 : 64); }
        }
        report(two ? "SearchByProjection(Cur, KF, sAlreadyFound) two cameras" : "SearchByProjection(Cur, KF, sAlreadyFound)", got[0], got[1]);
        EXPECT(got[0] == got[1] && got[0] > 50, "SearchByProjection(reloc) variant %d: %d vs %d", variant, got[0], got[1]);
        compare_worlds(P.r, P.s, "SearchByProjection(reloc)");
    }

    // ---- SearchByProjection(KeyFrame*, Sim3f&, vpPoints[, vpPointsKFs], vpMatched[, vpMatchedKF], th, ratioHamming)  :430-641 ----
    for (int with_kfs = 0; with_kfs < 4; ++with_kfs)       // bit 0: vpPointsKFs overload, bit 1: two-camera keyframe
        for (int variant = 0; variant < 2; ++variant) {
            const int two = with_kfs >> 1;
            Options o; o.seed = 60 + variant; o.two_cameras = two;
            Pair P(ea, eb, o);
            std::vector<int> out[2], outkf[2];
            int got[2];
            for (int k = 0; k < 2; ++k) {
                World &w = k ? P.s : P.r;
                const float s = variant ? 1.03f : 1.f;
                const Sophus::SE3f Tb = w.KB->GetPose();
                Sophus::Sim3f Scw(s, Tb.rotationMatrix(), Tb.translation() * s);
                std::vector<MapPoint *> vp = raw(w.mps, 0, w.nA);
                std::vector<KeyFrame *> vpKFs(vp.size());
                for (size_t i = 0; i < vp.size(); ++i) vpKFs[i] = i % 2 ? w.KA.get() : w.KB.get();
                std::vector<MapPoint *> matched(w.KB->N, nullptr);
                std::vector<KeyFrame *> matchedKF(w.KB->N, nullptr);
                for (int i = 0; i < w.KB->N; i += 5) matched[i] = w.mp((i * 3) % w.nA);
                const int th = variant ? 8 : 3;
                const float rh = variant ? 1.5f : 1.0f;
                if (k == 0) {
                    RefMatcher m(0.75f, true);
                    got[0] = (with_kfs & 1) ? m.SearchByProjection(w.KB.get(), Scw, vp, vpKFs, matched, matchedKF, th, rh)
                                      : m.SearchByProjection(w.KB.get(), Scw, vp, matched, th, rh);
                } else {
                    ShimMatcher m(0.75f, true);
                    got[1] = (with_kfs & 1) ? m.SearchByProjection(w.KB.get(), Scw, vp, vpKFs, matched, matchedKF, th, rh)
                                      : m.SearchByProjection(w.KB.get(), Scw, vp, matched, th, rh);
                }
                out[k] = ids(matched);
                for (KeyFrame *kf : matchedKF) outkf[k].push_back(kf == nullptr ? -1 : (kf == w.KA.get() ? 0 : 1));
            }
            report((with_kfs & 1) ? (two ? "SearchByProjection(KF, Scw, vpPoints, vpPointsKFs) two cameras" : "SearchByProjection(KF, Scw, vpPoints, vpPointsKFs)")
                                  : (two ? "SearchByProjection(KF, Scw, vpPoints) two cameras" : "SearchByProjection(KF, Scw, vpPoints)"), got[0], got[1]);
            EXPECT(got[0] == got[1] && got[0] > 50, "SearchByProjection(Sim3) kfs=%d variant=%d: %d vs %d", with_kfs, variant, got[0], got[1]);
            EXPECT(out[0] == out[1] && outkf[0] == outkf[1], "SearchByProjection(Sim3) kfs=%d variant=%d: vpMatched differ", with_kfs, variant);
            compare_worlds(P.r, P.s, "SearchByProjection(Sim3)");
        }

    // ---- SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12)  :758-900 ----
    for (int two = 0; two < 2; ++two)
        for (int variant = 0; variant < 2; ++variant) {
            Options o; o.seed = 70 + variant; o.two_cameras = two;
            Pair P(ea, eb, o);
            std::vector<MapPoint *> m0, m1;
            const int g0 = RefMatcher(variant ? 0.9f : 0.75f, variant == 0).SearchByBoW(P.r.KA.get(), P.r.KB.get(), m0);
            ShimMatcher sm(variant ? 0.9f : 0.75f, variant == 0);
            const int g1 = sm.SearchByBoW(P.s.KA.get(), P.s.KB.get(), m1);
            report(two ? "SearchByBoW(KF, KF) two cameras" : "SearchByBoW(KF, KF)", g0, g1);
            EXPECT(g0 == g1 && g0 > 30, "SearchByBoW(KF, KF) two=%d variant=%d: %d vs %d", two, variant, g0, g1);
            EXPECT(ids(m0) == ids(m1), "SearchByBoW(KF, KF) two=%d variant=%d: vpMatches12 differ", two, variant);
            compare_worlds(P.r, P.s, "SearchByBoW(KF, KF)");
        }

    // ---- SearchForTriangulation(KF1, KF2, vMatchedPairs, bOnlyStereo, bCoarse)  :902-1146 ----
    for (int two = 0; two < 2; ++two)
    for (int variant = 0; variant < 4; ++variant) {
        Options o; o.seed = 80 + variant; o.keep_mp = 0.4f; o.two_cameras = two;
        Pair P(ea, eb, o);
        const bool only_stereo = variant == 1, coarse = variant == 2, ori = variant != 3;
        std::vector<std::pair<size_t, size_t>> p0, p1;
        const int g0 = RefMatcher(0.6f, ori).SearchForTriangulation(P.r.KA.get(), P.r.KB.get(), p0, only_stereo, coarse);
        ShimMatcher sm(0.6f, ori);
        const int g1 = sm.SearchForTriangulation(P.s.KA.get(), P.s.KB.get(), p1, only_stereo, coarse);
        report(two ? "SearchForTriangulation two cameras" : "SearchForTriangulation", g0, g1);
        EXPECT(g0 == g1 && (g0 > 20 || (two && only_stereo)), "SearchForTriangulation variant %d: %d vs %d", variant, g0, g1);
        EXPECT(p0 == p1, "SearchForTriangulation variant %d: vMatchedPairs differ (%zu vs %zu)", variant, p0.size(), p1.size());
        compare_worlds(P.r, P.s, "SearchForTriangulation");
    }

    // ---- Fuse(KeyFrame*, vpMapPoints, th, bRight)  :1148-1335 ----
    for (int two = 0; two < 2; ++two)
        for (int variant = 0; variant < 2 + two; ++variant) {
            Options o; o.seed = 90 + variant; o.two_cameras = two;
            Pair P(ea, eb, o);
            const bool right = two && variant == 2;
            int got[2];
            for (int k = 0; k < 2; ++k) {
                World &w = k ? P.s : P.r;
                std::vector<MapPoint *> vp = raw(w.mps, 0, w.nA);
                for (size_t i = 0; i < vp.size(); i += 17) vp[i] = nullptr;
                vp.push_back(vp[3]);                                           // the same point twice: second time IsInKeyFrame
                const float th = variant == 1 ? 6.f : 3.f;
                if (k == 0) got[0] = RefMatcher().Fuse(w.KB.get(), vp, th, right);
                else { ShimMatcher m; got[1] = m.Fuse(w.KB.get(), vp, th, right); }
            }
            char what[64]; std::snprintf(what, sizeof what, "Fuse(KF, MPs) two=%d variant=%d", two, variant);
            report(what, got[0], got[1]);
            EXPECT(got[0] == got[1] && got[0] > (two ? 5 : 50), "%s: %d vs %d", what, got[0], got[1]);
            compare_worlds(P.r, P.s, what);
        }

    // ---- Fuse(KeyFrame*, Sim3f&, vpPoints, th, vpReplacePoint)  :1337-1446 ----
    for (int two = 0; two < 2; ++two)
    for (int variant = 0; variant < 2; ++variant) {
        Options o; o.seed = 100 + variant; o.two_cameras = two;
        Pair P(ea, eb, o);
        std::vector<int> rep[2];
        int got[2];
        for (int k = 0; k < 2; ++k) {
            World &w = k ? P.s : P.r;
            const float s = variant ? 0.98f : 1.f;
            const Sophus::SE3f Tb = w.KB->GetPose();
            Sophus::Sim3f Scw(s, Tb.rotationMatrix(), Tb.translation() * s);
            std::vector<MapPoint *> vp = raw(w.mps, 0, w.nA);
            std::vector<MapPoint *> replace(vp.size(), nullptr);
            if (k == 0) got[0] = RefMatcher().Fuse(w.KB.get(), Scw, vp, variant ? 6.f : 4.f, replace);
            else { ShimMatcher m; got[1] = m.Fuse(w.KB.get(), Scw, vp, variant ? 6.f : 4.f, replace); }
            rep[k] = ids(replace);
        }
        report(two ? "Fuse(KF, Scw) two cameras" : "Fuse(KF, Scw)", got[0], got[1]);
        EXPECT(got[0] == got[1] && got[0] > 50, "Fuse(KF, Scw) variant %d: %d vs %d", variant, got[0], got[1]);
        EXPECT(rep[0] == rep[1], "Fuse(KF, Scw) variant %d: vpReplacePoint differ", variant);
        compare_worlds(P.r, P.s, "Fuse(KF, Scw)");
    }

    // ---- SearchBySim3(KF1, KF2, vpMatches12, S12, th)  :1448-1665 ----
    for (int two = 0; two < 2; ++two)
    for (int variant = 0; variant < 2; ++variant) {
        Options o; o.seed = 110 + variant; o.two_cameras = two;
        Pair P(ea, eb, o);
        std::vector<int> out[2];
        int got[2];
        for (int k = 0; k < 2; ++k) {
            World &w = k ? P.s : P.r;
            const Sophus::SE3f T12 = w.KA->GetPose() * w.KB->GetPoseInverse();
            const float s = variant ? 1.01f : 1.f;
            const Sophus::Sim3f S12(s, T12.rotationMatrix(), T12.translation());
            std::vector<MapPoint *> m12(w.KA->N, nullptr);
            for (int i = 0; i < w.KA->N; i += 7)                                // earlier matches (from SearchByBoW)
                if (w.KA->mvpMapPoints[i]) { const int j = (i * 5) % w.KB->N; if (w.KB->mvpMapPoints[j]) m12[i] = w.KB->mvpMapPoints[j]; }
            if (k == 0) got[0] = RefMatcher().SearchBySim3(w.KA.get(), w.KB.get(), m12, S12, variant ? 10.f : 7.5f);
            else { ShimMatcher m; got[1] = m.SearchBySim3(w.KA.get(), w.KB.get(), m12, S12, variant ? 10.f : 7.5f); }
            out[k] = ids(m12);
        }
        report(two ? "SearchBySim3 two cameras" : "SearchBySim3", got[0], got[1]);
        EXPECT(got[0] == got[1] && got[0] > 20, "SearchBySim3 variant %d: %d vs %d", variant, got[0], got[1]);
        EXPECT(out[0] == out[1], "SearchBySim3 variant %d: vpMatches12 differ (%d vs %d set)", variant, count_set(out[0]), count_set(out[1]));
        compare_worlds(P.r, P.s, "SearchBySim3");
    }

    // ---- Frame::ComputeStereoFishEyeMatches (Frame.cc:1181-1225): kNN-2 + ratio 0.7 + the camera's triangulation gate ----
    {
        struct FishCamera {     // stand-in for KannalaBrandt8::TriangulateMatches: a deterministic function of its arguments
            float TriangulateMatches(FishCamera *other, const cv::KeyPoint &kp1, const cv::KeyPoint &kp2, const Eigen::Matrix3f &R12,
                                     const Eigen::Vector3f &t12, const float sigma1, const float sigma2, Eigen::Vector3f &p3D) {
                const float dx = kp2.pt.x - kp1.pt.x - t12(0) * 100.f, dy = kp1.pt.y - kp2.pt.y;   // frame b = frame a shifted by (+9, +5)
                if (std::fabs(dy) > 5.4f * sigma1 || dx <= 3.2f * sigma2 || other == nullptr) return -1.f;
                p3D = Eigen::Vector3f(kp1.pt.x * R12(0, 0), kp1.pt.y, 400.f / dx);
                return 400.f / dx;
            }
        };
        struct FishFrame {
            int Nleft = 0, Nright = 0, monoLeft = 0, monoRight = 0, mnCloseMPs = 7;
            std::vector<cv::KeyPoint> mvKeys, mvKeysRight;
            cv::Mat mDescriptors, mDescriptorsRight;
            std::vector<int> mvLeftToRightMatch, mvRightToLeftMatch;
            std::vector<float> mvDepth, mvuRight, mvLevelSigma2;
            std::vector<Eigen::Vector3f> mvStereo3Dpoints;
            FishCamera camL, camR;
            FishCamera *mpCamera = &camL, *mpCamera2 = &camR;
            Eigen::Matrix3f mRlr = Eigen::Matrix3f::Identity();
            Eigen::Vector3f mtlr = Eigen::Vector3f(0.05f, 0.f, 0.f);
        };
        FishFrame F;
        F.mvKeys = ea.keys; F.mvKeysRight = eb.keys;
        F.mDescriptors = ea.desc.clone(); F.mDescriptorsRight = eb.desc.clone();
        F.Nleft = (int)ea.keys.size(); F.Nright = (int)eb.keys.size();
        F.monoLeft = 120; F.monoRight = 77;
        F.mvLevelSigma2 = ea.sigma2;
        FishFrame G = F;
        VSG_SHIM::frame_ops::ComputeStereoFishEyeMatches<FishCamera>(F);
        // the reference's loop on the host, with cv::BFMatcher's order (distance, then train index)
        G.mvLeftToRightMatch.assign(G.Nleft, -1); G.mvRightToLeftMatch.assign(G.Nright, -1);
        G.mvDepth.assign(G.Nleft, -1.f); G.mvuRight.assign(G.Nleft, -1.f); G.mvStereo3Dpoints.assign(G.Nleft, Eigen::Vector3f());
        G.mnCloseMPs = 0;
        int nGood = 0;
        for (int q = G.monoLeft; q < G.Nleft; ++q) {
            int b0 = 1 << 30, b1 = 1 << 30, i0 = -1, i1 = -1;
            for (int t = G.monoRight; t < G.Nright; ++t) {
                const int d = RefMatcher::DescriptorDistance(G.mDescriptors.row(q), G.mDescriptorsRight.row(t));
                if (d < b0) { b1 = b0; i1 = i0; b0 = d; i0 = t; }
                else if (d < b1) { b1 = d; i1 = t; }
            }
            if (i1 < 0 || !((float)b0 < (float)b1 * 0.7)) continue;
            Eigen::Vector3f p3D;
            const float depth = G.camL.TriangulateMatches(G.mpCamera2, G.mvKeys[q], G.mvKeysRight[i0], G.mRlr, G.mtlr,
                                                          G.mvLevelSigma2[G.mvKeys[q].octave], G.mvLevelSigma2[G.mvKeysRight[i0].octave], p3D);
            if (depth > 0.0001f) { G.mvLeftToRightMatch[q] = i0; G.mvRightToLeftMatch[i0] = q; G.mvStereo3Dpoints[q] = p3D; G.mvDepth[q] = depth; ++nGood; }
        }
        bool same3d = true;
        for (int i = 0; i < F.Nleft; ++i)
            for (int k = 0; k < 3; ++k) same3d = same3d && F.mvStereo3Dpoints[i](k) == G.mvStereo3Dpoints[i](k);
        report("ComputeStereoFishEyeMatches (matches)", nGood, (int)std::count_if(F.mvLeftToRightMatch.begin(), F.mvLeftToRightMatch.end(), [](int v) { return v >= 0; }));
        EXPECT(nGood > 20, "ComputeStereoFishEyeMatches: only %d matches", nGood);
        EXPECT(F.mvLeftToRightMatch == G.mvLeftToRightMatch && F.mvRightToLeftMatch == G.mvRightToLeftMatch && F.mvDepth == G.mvDepth &&
                   F.mvuRight == G.mvuRight && same3d && F.mnCloseMPs == 0, "ComputeStereoFishEyeMatches differs from the reference loop");
    }

    std::printf("%s: %d checks, %d failed\n", g_fail ? "FAILED" : "ok", g_checks, g_fail);
    return g_fail ? 1 : 0;
}
