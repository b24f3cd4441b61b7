// shim_test.cpp — exercises the drop-in C++ classes (visual_sgraphs_b200/shim) exactly as the reference's
// callers would (Frame::ExtractORB, Tracking::SearchLocalPoints / TrackWithMotionModel / TrackReferenceKeyFrame /
// MonocularInitialization), with small stand-ins for Frame / KeyFrame / MapPoint that carry the members the
// reference's matcher reads, and checks every result against the CPU oracle.  Needs a GPU.
//   usage: shim_test frame_a.raw frame_b.raw   (two 640x480 8-bit frames written by the pytest wrapper)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <random>
#include <vector>

#include <set>
#include <tuple>

// ---- stand-ins for Eigen / Sophus (only what the shim's templates use) ----
struct Vec3 {
    float v[3];
    float operator()(int i) const { return v[i]; }
    Vec3 operator-(const Vec3 &o) const { return Vec3{{v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]}}; }
    Vec3 operator+(const Vec3 &o) const { return Vec3{{v[0] + o.v[0], v[1] + o.v[1], v[2] + o.v[2]}}; }
    Vec3 operator/(float s) const { return Vec3{{v[0] / s, v[1] / s, v[2] / s}}; }
    Vec3 operator*(float s) const { return Vec3{{v[0] * s, v[1] * s, v[2] * s}}; }
    float dot(const Vec3 &o) const { return v[0] * o.v[0] + v[1] * o.v[1] + v[2] * o.v[2]; }
    float norm() const { return std::sqrt(dot(*this)); }
};
struct Vec2 {
    float v[2];
    float operator()(int i) const { return v[i]; }
};
struct Mat3 {
    float m[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    int fill = 0;
    float operator()(int r, int c) const { return m[3 * r + c]; }
    Mat3 &operator<<(float x) { fill = 0; m[fill++] = x; return *this; }
    Mat3 &operator,(float x) { m[fill++] = x; return *this; }
    Mat3 transpose() const { Mat3 t; for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) t.m[3 * r + c] = m[3 * c + r]; return t; }
    Mat3 operator*(const Mat3 &o) const {
        Mat3 t;
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { float a = 0; for (int k = 0; k < 3; ++k) a += m[3 * r + k] * o.m[3 * k + c]; t.m[3 * r + c] = a; }
        return t;
    }
    Vec3 operator*(const Vec3 &p) const { return Vec3{{m[0] * p.v[0] + m[1] * p.v[1] + m[2] * p.v[2], m[3] * p.v[0] + m[4] * p.v[1] + m[5] * p.v[2], m[6] * p.v[0] + m[7] * p.v[1] + m[8] * p.v[2]}}; }
    Mat3 inverse() const {
        const float *a = m;
        const float det = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
        Mat3 t;
        t.m[0] = (a[4] * a[8] - a[5] * a[7]) / det; t.m[1] = (a[2] * a[7] - a[1] * a[8]) / det; t.m[2] = (a[1] * a[5] - a[2] * a[4]) / det;
        t.m[3] = (a[5] * a[6] - a[3] * a[8]) / det; t.m[4] = (a[0] * a[8] - a[2] * a[6]) / det; t.m[5] = (a[2] * a[3] - a[0] * a[5]) / det;
        t.m[6] = (a[3] * a[7] - a[4] * a[6]) / det; t.m[7] = (a[1] * a[6] - a[0] * a[7]) / det; t.m[8] = (a[0] * a[4] - a[1] * a[3]) / det;
        return t;
    }
};
namespace Sophus {
struct SE3f {   // rotation kept as a matrix; the tests use the identity rotation
    Mat3 R;
    Vec3 t{{0, 0, 0}};
    SE3f() {}
    SE3f(const Mat3 &R_, const Vec3 &t_) : R(R_), t(t_) {}
    SE3f inverse() const { Mat3 Rt = R.transpose(); Vec3 nt = Rt * t; return SE3f(Rt, Vec3{{-nt.v[0], -nt.v[1], -nt.v[2]}}); }
    Vec3 translation() const { return t; }
    Mat3 rotationMatrix() const { return R; }
    Vec3 operator*(const Vec3 &p) const { return R * p + t; }
    SE3f operator*(const SE3f &o) const { return SE3f(R * o.R, R * o.t + t); }
};
template <class T>
struct Sim3 {
    Mat3 R;
    Vec3 t{{0, 0, 0}};
    T s = 1;
    Mat3 rotationMatrix() const { return R; }
    Vec3 translation() const { return t; }
    T scale() const { return s; }
    Sim3 inverse() const { Sim3 o; o.R = R.transpose(); o.s = 1 / s; Vec3 nt = o.R * t; o.t = Vec3{{-nt.v[0] / s, -nt.v[1] / s, -nt.v[2] / s}}; return o; }
    Vec3 operator*(const Vec3 &p) const { return (R * p) * s + t; }
};
using Sim3f = Sim3<float>;
}  // namespace Sophus
#define SOPHUS_SIM3_HPP

#include "../../oracle/oracle.h"
#include "../../visual_sgraphs_b200/shim/FrameOps.h"
#include "../../visual_sgraphs_b200/shim/ORBextractor.h"
#include "../../visual_sgraphs_b200/shim/ORBmatcher.h"

static int g_fail = 0;
#define EXPECT(cond, ...)                                  \
    do {                                                   \
        if (!(cond)) { std::printf("FAIL %s:%d: ", __FILE__, __LINE__); std::printf(__VA_ARGS__); std::printf("\n"); ++g_fail; } \
    } while (0)

// ---- stand-ins for the reference's types (only what ORBmatcher reads) ----
using Pose = Sophus::SE3f;
static Pose translation_pose(float x, float y, float z) { return Pose(Mat3(), Vec3{{x, y, z}}); }
struct Camera {
    float fx = 500, fy = 500, cx = 320, cy = 240;
    Vec2 project(const Vec3 &p) const { return Vec2{{fx * p.v[0] / p.v[2] + cx, fy * p.v[1] / p.v[2] + cy}}; }
    Mat3 toK_() const { Mat3 K; K << fx, 0.f, cx, 0.f, fy, cy, 0.f, 0.f, 1.f; return K; }
    // the per-pair test of two-camera SearchForTriangulation (the caller's camera code); not reached by this program's
    // single-camera scenarios — tests/cpp/ref_matcher_test.cpp covers that branch against the reference itself
    bool epipolarConstrain(Camera *, const cv::KeyPoint &, const cv::KeyPoint &, const Mat3 &, const Vec3 &, float, float) const { return false; }
};
struct KeyFrame;
struct Frame;
struct MapPoint {
    bool mbTrackInView = false, mbTrackInViewR = false;
    float mTrackProjX = 0, mTrackProjY = 0, mTrackProjXR = 0, mTrackViewCos = 1, mTrackDepth = 1;
    float mTrackProjYR = 0, mTrackViewCosR = 1;
    int mnTrackScaleLevel = 0, mnTrackScaleLevelR = -1;
    bool bad = false;
    int obs = 1;
    cv::Mat desc;
    Vec3 pos{{0, 0, 1}};
    Vec3 normal{{0, 0, 1}};
    float max_dist = 1e4f, min_dist = 0.01f;
    std::set<const void *> in_kf;
    MapPoint *replaced_by = nullptr;
    std::map<const void *, int> index_in;
    bool isBad() const { return bad; }
    int Observations() const { return obs; }
    cv::Mat GetDescriptor() const { return desc.clone(); }
    Vec3 GetWorldPos() const { return pos; }
    Vec3 GetNormal() const { return normal; }
    float GetMaxDistanceInvariance() const { return max_dist; }
    float GetMinDistanceInvariance() const { return min_dist; }
    template <class T> int PredictScale(float dist, T *pF) const {      // MapPoint.cc:533-565
        const float ratio = max_dist / dist;
        int nScale = std::ceil(std::log(ratio) / pF->mfLogScaleFactor);
        if (nScale < 0) nScale = 0;
        else if (nScale >= pF->mnScaleLevels) nScale = pF->mnScaleLevels - 1;
        return nScale;
    }
    bool IsInKeyFrame(const void *kf) const { return in_kf.count(kf) != 0; }
    void AddObservation(const void *kf, int idx) { in_kf.insert(kf); index_in[kf] = idx; ++obs; }
    void Replace(MapPoint *o) { bad = true; replaced_by = o; o->in_kf.insert(in_kf.begin(), in_kf.end()); o->obs += obs; }
    std::tuple<int, int> GetIndexInKeyFrame(const void *kf) const { auto it = index_in.find(kf); return std::make_tuple(it == index_in.end() ? -1 : it->second, -1); }
};
static long unsigned int g_next_frame_id = 0;
struct Frame {
    long unsigned int mnId = g_next_frame_id++;      // Frame.h:217 / KeyFrame.h:312: what the shim's frame cache keys on
    int N = 0, Nleft = -1;
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn, mvKeysRight;
    std::vector<int> mvLeftToRightMatch, mvRightToLeftMatch;
    cv::Mat mDescriptors;
    std::vector<float> mvuRight;
    std::vector<float> mvScaleFactors;
    std::vector<MapPoint *> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    std::map<unsigned, std::vector<unsigned>> mFeatVec;
    float mnMinX = 0, mnMinY = 0, mnMaxX = 640, mnMaxY = 480;
    float mfGridElementWidthInv = 64.f / 640.f, mfGridElementHeightInv = 48.f / 480.f;
    float mb = 0.1f, mbf = 40.f;
    Pose pose;
    Camera cam;
    Camera *mpCamera = &cam;
    Camera *mpCamera2 = nullptr;
    float mfLogScaleFactor = std::log(1.2f);
    int mnScaleLevels = 8;
    Pose GetPose() const { return pose; }
    Pose Trl;                                        // right-from-left camera transform of a two-camera rig
    Pose GetRelativePoseTrl() const { return Trl; }
    std::vector<MapPoint *> GetMapPointMatches() const { return mvpMapPoints; }   // KeyFrame interface
};
struct KeyFrame : Frame {   // the members ORBmatcher reads from a KeyFrame (KeyFrame.h)
    int NLeft = -1;
    float fx = 500, fy = 500, cx = 320, cy = 240;
    std::vector<float> mvLevelSigma2, mvInvLevelSigma2;
    explicit KeyFrame(const Frame &f) : Frame(f) {
        mpCamera = &cam;
        for (float sc : mvScaleFactors) { mvLevelSigma2.push_back(sc * sc); mvInvLevelSigma2.push_back(1.0f / (sc * sc)); }
    }
    KeyFrame(const KeyFrame &) = delete;
    MapPoint *GetMapPoint(size_t i) const { return mvpMapPoints[i]; }
    void AddMapPoint(MapPoint *p, size_t i) { mvpMapPoints[i] = p; }
    std::set<MapPoint *> GetMapPoints() const { std::set<MapPoint *> s; for (MapPoint *p : mvpMapPoints) if (p && !p->isBad()) s.insert(p); return s; }
    bool IsInImage(float x, float y) const { return x >= mnMinX && x < mnMaxX && y >= mnMinY && y < mnMaxY; }
    Vec3 GetCameraCenter() const { return pose.inverse().translation(); }
    Pose right_pose;                                  // two-camera rigs: pose of the right camera
    Pose GetRightPose() const { return right_pose; }
    Vec3 GetRightCameraCenter() const { return right_pose.inverse().translation(); }
    Pose GetPoseInverse() const { return pose.inverse(); }
    Pose GetRightPoseInverse() const { return right_pose.inverse(); }
};

static std::vector<unsigned char> read_raw(const char *path, size_t n) {
    std::vector<unsigned char> b(n);
    FILE *f = std::fopen(path, "rb");
    if (!f || std::fread(b.data(), 1, n, f) != n) { std::printf("cannot read %s\n", path); std::exit(2); }
    std::fclose(f);
    return b;
}

static void fill_frame(Frame &F, VS_GRAPHS::ORBextractor &ex, std::vector<unsigned char> &img, bool stereo, std::mt19937 &rng) {
    cv::Mat m(480, 640, CV_8UC1, img.data(), 640);
    std::vector<int> lap = {0, 0};
    ex(m, cv::noArray(), F.mvKeys, F.mDescriptors, lap);
    F.mvKeysUn = F.mvKeys;
    F.N = (int)F.mvKeys.size();
    F.mvScaleFactors = ex.GetScaleFactors();
    F.mvpMapPoints.assign(F.N, nullptr);
    F.mvbOutlier.assign(F.N, false);
    F.mvuRight.assign(F.N, -1.f);
    if (stereo)
        for (int i = 0; i < F.N; ++i)
            if (rng() % 10 < 7) F.mvuRight[i] = F.mvKeys[i].pt.x - (2 + (rng() % 380) / 10.f);
    for (int i = 0; i < F.N; ++i) F.mFeatVec[(F.mDescriptors.ptr(i)[0] * 7u + 3u) % 24u].push_back(i);
}

static orc_frame_view view_of(const Frame &F, std::vector<unsigned char> &desc_store) {
    desc_store.resize((size_t)F.N * 32);
    for (int i = 0; i < F.N; ++i) std::memcpy(&desc_store[(size_t)i * 32], F.mDescriptors.ptr(i), 32);
    orc_frame_view v;
    v.n = F.N;
    v.keys = reinterpret_cast<const orc_keypoint *>(F.mvKeysUn.data());
    v.descriptors = desc_store.data();
    v.u_right = F.mvuRight.data();
    v.min_x = F.mnMinX; v.min_y = F.mnMinY; v.max_x = F.mnMaxX; v.max_y = F.mnMaxY;
    v.grid_inv_w = F.mfGridElementWidthInv; v.grid_inv_h = F.mfGridElementHeightInv;
    v.grid_cols = 64; v.grid_rows = 48;
    v.scale_factors = F.mvScaleFactors.data();
    v.n_levels = (int)F.mvScaleFactors.size();
    return v;
}

int main(int argc, char **argv) {
    if (argc < 3) { std::printf("usage: %s frame_a.raw frame_b.raw\n", argv[0]); return 2; }
    std::vector<unsigned char> img_a = read_raw(argv[1], 640 * 480), img_b = read_raw(argv[2], 640 * 480);
    std::mt19937 rng(7);

    // ---------------- ORBextractor: operator(), getters, mvImagePyramid ----------------
    VS_GRAPHS::ORBextractor ex(1000, 1.2f, 8, 20, 7);
    ORB_SLAM3::ORBextractor *alias = &ex;   // the namespace the north star uses
    (void)alias;
    orc_extractor *orc = orc_extractor_create(1000, 1.2f, 8, 20, 7);
    {
        cv::Mat m(480, 640, CV_8UC1, img_a.data(), 640);
        for (int lap1 : {0, 1000}) {
            std::vector<cv::KeyPoint> kps;
            cv::Mat desc;
            std::vector<int> lap = {0, lap1};
            const int mono = ex(m, cv::noArray(), kps, desc, lap);
            const int omono = orc_extract(orc, img_a.data(), 640, 480, 640, 0, lap1);
            const int n = orc_num_keypoints(orc);
            std::vector<orc_keypoint> okps(n);
            std::vector<unsigned char> odesc((size_t)n * 32);
            orc_get_keypoints(orc, okps.data(), odesc.data());
            EXPECT(mono == omono, "monoIndex %d vs %d", mono, omono);
            EXPECT((int)kps.size() == n && desc.rows == n && desc.cols == 32, "count %zu vs %d", kps.size(), n);
            long badbits = 0;
            for (int i = 0; i < n && i < (int)kps.size(); ++i) {
                EXPECT(kps[i].pt.x == okps[i].x && kps[i].pt.y == okps[i].y && kps[i].octave == okps[i].octave &&
                           kps[i].size == okps[i].size && kps[i].response == okps[i].response && kps[i].class_id == -1,
                       "keypoint %d differs", i);
                EXPECT(std::fabs(kps[i].angle - okps[i].angle) <= 1e-3f, "angle %d: %f vs %f", i, kps[i].angle, okps[i].angle);
                for (int b = 0; b < 32; ++b) badbits += __builtin_popcount(desc.ptr(i)[b] ^ odesc[(size_t)i * 32 + b]);
            }
            EXPECT(badbits <= (long)n * 256 / 1000, "descriptor bits differing: %ld", badbits);
        }
        EXPECT(ex.GetLevels() == 8 && ex.GetScaleFactors().size() == 8, "getters");
        std::vector<float> s(8), is(8), s2(8), is2(8);
        orc_scale_factors(orc, s.data(), is.data(), s2.data(), is2.data());
        EXPECT(ex.GetScaleFactors() == s && ex.GetInverseScaleFactors() == is && ex.GetScaleSigmaSquares() == s2 &&
                   ex.GetInverseScaleSigmaSquares() == is2, "scale tables");
        for (int level = 0; level < 8; ++level) {   // mvImagePyramid incl. the 19-px reflected frame
            int w, h;
            orc_level_size(orc, level, &w, &h);
            std::vector<unsigned char> pad((size_t)(w + 38) * (h + 38));
            orc_get_level_padded(orc, level, pad.data());
            const cv::Mat &L = ex.mvImagePyramid[level];
            EXPECT(L.cols == w && L.rows == h, "level %d size", level);
            long diff = 0;
            for (int y = -19; y < h + 19; ++y)
                for (int x = -19; x < w + 19; ++x)
                    diff += L.data[(ptrdiff_t)y * (ptrdiff_t)L.step + x] != pad[(size_t)(y + 19) * (w + 38) + (x + 19)];
            EXPECT(diff == 0, "level %d: %ld padded pixels differ", level, diff);
        }
        std::vector<cv::KeyPoint> kps;
        cv::Mat desc, empty;
        std::vector<int> lap = {0, 0};
        EXPECT(ex(empty, cv::noArray(), kps, desc, lap) == -1, "empty image must return -1");
    }

    // ---------------- frames for the matcher ----------------
    ex.SetPyramidDownload(false);
    Frame A, B;
    fill_frame(A, ex, img_a, true, rng);
    fill_frame(B, ex, img_b, false, rng);
    std::vector<unsigned char> desc_a, desc_b;
    orc_frame_view va = view_of(A, desc_a), vb = view_of(B, desc_b);

    // DescriptorDistance
    for (int i = 0; i < 50; ++i)
        EXPECT(VS_GRAPHS::ORBmatcher::DescriptorDistance(A.mDescriptors.row(i), B.mDescriptors.row(i)) ==
                   orc_descriptor_distance(A.mDescriptors.ptr(i), B.mDescriptors.ptr(i)), "DescriptorDistance %d", i);
    EXPECT(VS_GRAPHS::ORBmatcher::TH_LOW == 50 && VS_GRAPHS::ORBmatcher::TH_HIGH == 100 && VS_GRAPHS::ORBmatcher::HISTO_LENGTH == 30, "constants");

    // ---------------- SearchByProjection(F, vpMapPoints) ----------------
    {
        std::vector<MapPoint> store(B.N);
        std::vector<MapPoint *> vp(B.N);
        std::vector<orc_track_point> pts(B.N);
        std::vector<unsigned char> mpdesc((size_t)B.N * 32);
        for (int i = 0; i < B.N; ++i) {
            MapPoint &mp = store[i];
            mp.mbTrackInView = rng() % 10 < 9;
            mp.mTrackProjX = B.mvKeys[i].pt.x - 9 + (int)(rng() % 5) - 2;
            mp.mTrackProjY = B.mvKeys[i].pt.y - 5 + (int)(rng() % 5) - 2;
            mp.mTrackProjXR = mp.mTrackProjX - (2 + (rng() % 380) / 10.f);
            mp.mTrackViewCos = 0.99f + (rng() % 100) / 10000.f;
            mp.mTrackDepth = 1 + rng() % 60;
            mp.mnTrackScaleLevel = B.mvKeys[i].octave;
            mp.bad = rng() % 30 == 0;
            mp.obs = rng() % 10 < 9 ? 2 : 0;
            mp.desc = B.mDescriptors.row(i).clone();
            vp[i] = &mp;
            orc_track_point &p = pts[i];
            std::memset(&p, 0, sizeof(p));
            p.proj_x = mp.mTrackProjX; p.proj_y = mp.mTrackProjY; p.proj_xr = mp.mTrackProjXR;
            p.view_cos = mp.mTrackViewCos; p.depth = mp.mTrackDepth; p.level = mp.mnTrackScaleLevel;
            p.in_view = mp.mbTrackInView; p.bad = mp.bad; p.blocks = mp.obs > 0;
            std::memcpy(&mpdesc[(size_t)i * 32], mp.desc.ptr(0), 32);
        }
        MapPoint pre;   // a keypoint that is already taken by a point with observations
        pre.obs = 3;
        std::vector<unsigned char> occupied(A.N, 0);
        for (int i = 0; i < A.N; i += 11) { A.mvpMapPoints[i] = &pre; occupied[i] = 1; }
        std::vector<int32_t> assign(A.N);
        const int want = orc_search_by_projection_map(&va, occupied.data(), B.N, pts.data(), mpdesc.data(), 3.f, 1, 45.f, 0.8f, assign.data());
        VS_GRAPHS::ORBmatcher matcher(0.8f);
        const int got = matcher.SearchByProjection(A, vp, 3.f, true, 45.f);
        EXPECT(got == want && got > 100, "SearchByProjection(map): %d vs %d", got, want);
        for (int i = 0; i < A.N; ++i) {
            MapPoint *expect = assign[i] >= 0 ? vp[assign[i]] : (occupied[i] ? &pre : nullptr);
            EXPECT(A.mvpMapPoints[i] == expect, "mvpMapPoints[%d]", i);
        }
    }

    // ---------------- SearchByProjection(F, vpMapPoints) on a two-camera frame (Nleft != -1) ----------------
    {
        Frame F2;                                   // left camera = A's keypoints, right camera = B's
        F2.Nleft = A.N;
        F2.mvKeys = A.mvKeys; F2.mvKeysRight = B.mvKeys;
        F2.N = A.N + B.N;
        F2.mvScaleFactors = A.mvScaleFactors;
        F2.mDescriptors.create(F2.N, 32, CV_8U);
        for (int i = 0; i < A.N; ++i) std::memcpy(F2.mDescriptors.ptr(i), A.mDescriptors.ptr(i), 32);
        for (int i = 0; i < B.N; ++i) std::memcpy(F2.mDescriptors.ptr(A.N + i), B.mDescriptors.ptr(i), 32);
        F2.mvpMapPoints.assign(F2.N, nullptr);
        F2.mvLeftToRightMatch.assign(A.N, -1);
        F2.mvRightToLeftMatch.assign(B.N, -1);
        for (int k = 0; k < 200; ++k) {
            const int l = rng() % A.N, r = rng() % B.N;
            if (F2.mvLeftToRightMatch[l] == -1 && F2.mvRightToLeftMatch[r] == -1) { F2.mvLeftToRightMatch[l] = r; F2.mvRightToLeftMatch[r] = l; }
        }
        const int nMP = A.N + B.N;
        std::vector<MapPoint> store(nMP);
        std::vector<MapPoint *> vp(nMP);
        std::vector<orc_track_point> pl(nMP), pr(nMP);
        std::vector<unsigned char> mpdesc((size_t)nMP * 32);
        for (int i = 0; i < nMP; ++i) {
            MapPoint &mp = store[i];
            const bool from_left = i % 2 == 0;
            const int src = from_left ? (i / 2) % A.N : (i / 2) % B.N;
            const cv::KeyPoint &kl = A.mvKeys[from_left ? src : rng() % A.N], &kr = B.mvKeys[from_left ? rng() % B.N : src];
            mp.mbTrackInView = rng() % 10 < 7; mp.mbTrackInViewR = rng() % 10 < 7;
            mp.mTrackProjX = kl.pt.x + (int)(rng() % 5) - 2; mp.mTrackProjY = kl.pt.y + (int)(rng() % 5) - 2;
            mp.mTrackProjXR = kr.pt.x + (int)(rng() % 5) - 2; mp.mTrackProjYR = kr.pt.y + (int)(rng() % 5) - 2;
            mp.mTrackViewCos = 0.99f + (rng() % 100) / 10000.f; mp.mTrackViewCosR = 0.99f + (rng() % 100) / 10000.f;
            mp.mnTrackScaleLevel = kl.octave; mp.mnTrackScaleLevelR = rng() % 20 == 0 ? -1 : kr.octave;
            mp.mTrackDepth = 1 + rng() % 60;
            mp.bad = rng() % 30 == 0;
            mp.obs = rng() % 10 < 9 ? 2 : 0;
            mp.desc = (from_left ? A.mDescriptors.row(src) : B.mDescriptors.row(src)).clone();
            vp[i] = &mp;
            std::memset(&pl[i], 0, sizeof(orc_track_point)); std::memset(&pr[i], 0, sizeof(orc_track_point));
            pl[i].proj_x = mp.mTrackProjX; pl[i].proj_y = mp.mTrackProjY; pl[i].view_cos = mp.mTrackViewCos; pl[i].level = mp.mnTrackScaleLevel;
            pl[i].in_view = mp.mbTrackInView; pl[i].depth = mp.mTrackDepth; pl[i].bad = mp.bad; pl[i].blocks = mp.obs > 0;
            pr[i].proj_x = mp.mTrackProjXR; pr[i].proj_y = mp.mTrackProjYR; pr[i].view_cos = mp.mTrackViewCosR; pr[i].level = mp.mnTrackScaleLevelR;
            pr[i].in_view = mp.mbTrackInViewR;
            std::memcpy(&mpdesc[(size_t)i * 32], mp.desc.ptr(0), 32);
        }
        MapPoint pre;
        pre.obs = 3;
        std::vector<unsigned char> occupied(F2.N, 0);
        for (int i = 0; i < F2.N; i += 13) { F2.mvpMapPoints[i] = &pre; occupied[i] = 1; }
        orc_frame_view vl = va, vr = vb;
        vl.u_right = nullptr; vr.u_right = nullptr;
        std::vector<int32_t> assign(F2.N);
        const int want = orc_search_by_projection_map_2cam(&vl, &vr, occupied.data(), F2.mvLeftToRightMatch.data(), F2.mvRightToLeftMatch.data(),
                                                           nMP, pl.data(), pr.data(), mpdesc.data(), 3.f, 1, 45.f, 0.8f, assign.data());
        VS_GRAPHS::ORBmatcher matcher(0.8f);
        const int got = matcher.SearchByProjection(F2, vp, 3.f, true, 45.f);
        EXPECT(got == want && got > 200, "SearchByProjection(map, two cameras): %d vs %d", got, want);
        int right_hits = 0;
        for (int i = 0; i < F2.N; ++i) {
            MapPoint *expect = assign[i] >= 0 ? vp[assign[i]] : (occupied[i] ? &pre : nullptr);
            EXPECT(F2.mvpMapPoints[i] == expect, "two cameras: mvpMapPoints[%d]", i);
            right_hits += i >= A.N && assign[i] >= 0;
        }
        EXPECT(right_hits > 50, "two cameras: right-camera matches %d", right_hits);
    }

    // ---------------- SearchByProjection(Cur, Last) ----------------
    {
        Frame Cur = A, Last = B;
        Cur.mpCamera = &Cur.cam; Last.mpCamera = &Last.cam;
        Cur.mvpMapPoints.assign(Cur.N, nullptr);
        Cur.pose = translation_pose(0.02f, -0.01f, -0.3f);    // tlc = Tlw * (-t): z = +0.3 > mb -> forward
        Last.pose = translation_pose(0, 0, 0);
        std::vector<MapPoint> store(Last.N);
        std::vector<orc_proj_point> pts(Last.N);
        std::vector<unsigned char> pdesc((size_t)Last.N * 32, 0);
        for (int i = 0; i < Last.N; ++i) {
            MapPoint &mp = store[i];
            const float z = 2.f + (rng() % 100) / 10.f;
            // a world point whose projection in Cur lands near the shifted keypoint
            const float u = Last.mvKeys[i].pt.x - 9, v = Last.mvKeys[i].pt.y - 5;
            const Vec3 xc{{(u - 320) * z / 500, (v - 240) * z / 500, z}};
            mp.pos = xc - Cur.pose.t;
            mp.obs = rng() % 10 < 9 ? 1 : 0;
            mp.desc = Last.mDescriptors.row(i).clone();
            Last.mvpMapPoints[i] = rng() % 10 < 8 ? &mp : nullptr;
            Last.mvbOutlier[i] = rng() % 20 == 0;
            orc_proj_point &p = pts[i];
            std::memset(&p, 0, sizeof(p));
            if (!Last.mvpMapPoints[i] || Last.mvbOutlier[i]) continue;
            const Vec3 x3Dc = Cur.pose * mp.pos;
            const float invzc = 1.0 / x3Dc(2);
            if (invzc < 0) continue;
            const Vec2 uv = Cur.cam.project(x3Dc);
            if (uv(0) < Cur.mnMinX || uv(0) > Cur.mnMaxX || uv(1) < Cur.mnMinY || uv(1) > Cur.mnMaxY) continue;
            p.valid = 1; p.u = uv(0); p.v = uv(1); p.ur = uv(0) - Cur.mbf * invzc;
            p.octave = Last.mvKeys[i].octave; p.angle = Last.mvKeysUn[i].angle; p.blocks = mp.obs > 0;
            std::memcpy(&pdesc[(size_t)i * 32], mp.desc.ptr(0), 32);
        }
        std::vector<unsigned char> dc;
        orc_frame_view vc = view_of(Cur, dc);
        std::vector<unsigned char> occupied(Cur.N, 0);
        std::vector<int32_t> assign(Cur.N);
        for (bool mono : {true, false}) {
            Cur.mvpMapPoints.assign(Cur.N, nullptr);
            // bMono -> window by octave +-1 (mode 0); stereo with tlc.z = 0.3 > mb -> forward (mode 1)
            const int want = orc_search_by_projection_last(&vc, occupied.data(), Last.N, pts.data(), pdesc.data(), 15.f, mono ? 0 : 1, 1, assign.data());
            VS_GRAPHS::ORBmatcher matcher(0.9f, true);
            const int got = matcher.SearchByProjection(Cur, Last, 15.f, mono);
            EXPECT(got == want && got > 50, "SearchByProjection(last, mono=%d): %d vs %d", (int)mono, got, want);
            for (int i = 0; i < Cur.N; ++i) {
                MapPoint *expect = assign[i] >= 0 ? Last.mvpMapPoints[assign[i]] : nullptr;
                EXPECT(Cur.mvpMapPoints[i] == expect, "Cur.mvpMapPoints[%d]", i);
            }
        }
    }

    // ---------------- SearchByProjection(Cur, Last) with a two-camera current frame ----------------
    {
        Frame Cur, Last = B;                       // Cur: left camera = A's keypoints, right camera = A's shifted by the baseline
        Cur.Nleft = A.N;
        Cur.mvKeys = A.mvKeys; Cur.mvKeysRight = A.mvKeys;
        for (cv::KeyPoint &k : Cur.mvKeysRight) k.pt.x -= 12.f;
        Cur.N = 2 * A.N;
        Cur.mvScaleFactors = A.mvScaleFactors;
        Cur.mDescriptors.create(Cur.N, 32, CV_8U);
        for (int i = 0; i < A.N; ++i) { std::memcpy(Cur.mDescriptors.ptr(i), A.mDescriptors.ptr(i), 32); std::memcpy(Cur.mDescriptors.ptr(A.N + i), A.mDescriptors.ptr(i), 32); }
        Cur.mpCamera = &Cur.cam; Last.mpCamera = &Last.cam;
        Cur.pose = translation_pose(0.02f, -0.01f, -0.3f);
        Cur.Trl = translation_pose(-0.06f, 0.f, 0.f);
        Last.pose = translation_pose(0, 0, 0);
        std::vector<MapPoint> store(Last.N);
        std::vector<orc_proj_point> pl(Last.N), pr(Last.N);
        std::vector<unsigned char> pdesc((size_t)Last.N * 32, 0);
        for (int i = 0; i < Last.N; ++i) {
            MapPoint &mp = store[i];
            const float z = 2.5f;                  // -0.06 * 500 / 2.5 = -12 px: the right projection lands on the shifted keypoint
            const float u = Last.mvKeys[i].pt.x - 9, v = Last.mvKeys[i].pt.y - 5;
            const Vec3 xc{{(u - 320) * z / 500, (v - 240) * z / 500, z}};
            mp.pos = xc - Cur.pose.t;
            mp.obs = rng() % 10 < 9 ? 1 : 0;
            mp.desc = Last.mDescriptors.row(i).clone();
            Last.mvpMapPoints[i] = rng() % 10 < 8 ? &mp : nullptr;
            Last.mvbOutlier[i] = rng() % 20 == 0;
            std::memset(&pl[i], 0, sizeof(orc_proj_point)); std::memset(&pr[i], 0, sizeof(orc_proj_point));
            if (!Last.mvpMapPoints[i] || Last.mvbOutlier[i]) continue;
            const Vec3 x3Dc = Cur.pose * mp.pos;
            const float invzc = 1.0 / x3Dc(2);
            if (invzc < 0) continue;
            const Vec2 uv = Cur.cam.project(x3Dc);
            if (uv(0) < Cur.mnMinX || uv(0) > Cur.mnMaxX || uv(1) < Cur.mnMinY || uv(1) > Cur.mnMaxY) continue;
            orc_proj_point &p = pl[i];
            p.valid = 1; p.u = uv(0); p.v = uv(1); p.ur = uv(0) - Cur.mbf * invzc;
            p.octave = Last.mvKeys[i].octave; p.angle = Last.mvKeysUn[i].angle; p.blocks = mp.obs > 0;
            const Vec2 uvr = Cur.cam.project(Cur.Trl * x3Dc);
            pr[i].u = uvr(0); pr[i].v = uvr(1);
            std::memcpy(&pdesc[(size_t)i * 32], mp.desc.ptr(0), 32);
        }
        std::vector<unsigned char> dl, dr;
        Frame camL = A, camR = A;
        camR.mvKeys = Cur.mvKeysRight; camR.mvKeysUn = Cur.mvKeysRight;
        orc_frame_view vl = view_of(camL, dl), vr = view_of(camR, dr);
        vl.u_right = nullptr; vr.u_right = nullptr;
        std::vector<unsigned char> occupied(Cur.N, 0);
        std::vector<int32_t> assign(Cur.N);
        for (bool mono : {true, false}) {
            Cur.mvpMapPoints.assign(Cur.N, nullptr);
            const int want = orc_search_by_projection_last_2cam(&vl, &vr, occupied.data(), Last.N, pl.data(), pr.data(), pdesc.data(), 15.f,
                                                                mono ? 0 : 1, 1, assign.data());
            VS_GRAPHS::ORBmatcher matcher(0.9f, true);
            const int got = matcher.SearchByProjection(Cur, Last, 15.f, mono);
            EXPECT(got == want && got > 100, "SearchByProjection(last, two cameras, mono=%d): %d vs %d", (int)mono, got, want);
            int right_hits = 0;
            for (int i = 0; i < Cur.N; ++i) {
                MapPoint *expect = assign[i] >= 0 ? Last.mvpMapPoints[assign[i]] : nullptr;
                EXPECT(Cur.mvpMapPoints[i] == expect, "two cameras: Cur.mvpMapPoints[%d]", i);
                right_hits += i >= A.N && assign[i] >= 0;
            }
            EXPECT(right_hits > 30, "two cameras (last): right-camera matches %d", right_hits);
        }
    }

    // ---------------- SearchByBoW(KF, F) ----------------
    {
        Frame KF = B, F = A;
        std::vector<MapPoint> store(KF.N);
        std::vector<unsigned char> valid(KF.N, 0);
        for (int i = 0; i < KF.N; ++i) {
            store[i].bad = rng() % 25 == 0;
            KF.mvpMapPoints[i] = rng() % 10 < 9 ? &store[i] : nullptr;
            valid[i] = KF.mvpMapPoints[i] && !store[i].bad;
        }
        auto flat = [](const std::map<unsigned, std::vector<unsigned>> &fv, std::vector<int32_t> &n, std::vector<int32_t> &p, std::vector<int32_t> &x) {
            p.push_back(0);
            for (auto &kv : fv) { n.push_back(kv.first); for (unsigned v : kv.second) x.push_back(v); p.push_back((int32_t)x.size()); }
        };
        std::vector<int32_t> kn, kp, ki, fn, fp, fi;
        flat(KF.mFeatVec, kn, kp, ki);
        flat(F.mFeatVec, fn, fp, fi);
        std::vector<unsigned char> dk, df;
        orc_frame_view vk = view_of(KF, dk), vf = view_of(F, df);
        std::vector<int32_t> mf(F.N);
        const int want = orc_search_by_bow(&vk, valid.data(), &vf, (int)kn.size(), kn.data(), kp.data(), ki.data(), (int)fn.size(),
                                           fn.data(), fp.data(), fi.data(), 0.7f, 1, mf.data());
        VS_GRAPHS::ORBmatcher matcher(0.7f, true);
        std::vector<MapPoint *> matches;
        const int got = matcher.SearchByBoW(&KF, F, matches);
        EXPECT(got == want && got > 10, "SearchByBoW: %d vs %d", got, want);
        EXPECT((int)matches.size() == F.N, "SearchByBoW output size");
        for (int j = 0; j < F.N && j < (int)matches.size(); ++j)
            EXPECT(matches[j] == (mf[j] >= 0 ? KF.mvpMapPoints[mf[j]] : nullptr), "vpMapPointMatches[%d]", j);
    }

    // ---------------- SearchByBoW(KF, F) with a two-camera frame ----------------
    {
        Frame KF = B, F2;                           // F2: left camera = A's keypoints, right camera = B's
        F2.Nleft = A.N;
        F2.mvKeys = A.mvKeys; F2.mvKeysRight = B.mvKeys;
        F2.N = A.N + B.N;
        F2.mvScaleFactors = A.mvScaleFactors;
        F2.mDescriptors.create(F2.N, 32, CV_8U);
        for (int i = 0; i < A.N; ++i) std::memcpy(F2.mDescriptors.ptr(i), A.mDescriptors.ptr(i), 32);
        for (int i = 0; i < B.N; ++i) std::memcpy(F2.mDescriptors.ptr(A.N + i), B.mDescriptors.ptr(i), 32);
        for (int i = 0; i < F2.N; ++i) F2.mFeatVec[(F2.mDescriptors.ptr(i)[0] * 7u + 3u) % 24u].push_back(i);
        std::vector<MapPoint> store(KF.N);
        std::vector<unsigned char> valid(KF.N, 0);
        for (int i = 0; i < KF.N; ++i) {
            store[i].bad = rng() % 25 == 0;
            KF.mvpMapPoints[i] = rng() % 10 < 9 ? &store[i] : nullptr;
            valid[i] = KF.mvpMapPoints[i] && !store[i].bad;
        }
        auto flat = [](const std::map<unsigned, std::vector<unsigned>> &fv, std::vector<int32_t> &n, std::vector<int32_t> &p, std::vector<int32_t> &x) {
            p.push_back(0);
            for (auto &kv : fv) { n.push_back(kv.first); for (unsigned v : kv.second) x.push_back(v); p.push_back((int32_t)x.size()); }
        };
        std::vector<int32_t> kn, kp, ki, fn, fp, fi;
        flat(KF.mFeatVec, kn, kp, ki);
        flat(F2.mFeatVec, fn, fp, fi);
        std::vector<unsigned char> dk, df((size_t)F2.N * 32);
        orc_frame_view vk = view_of(KF, dk);
        std::vector<cv::KeyPoint> allkeys(F2.mvKeys);
        allkeys.insert(allkeys.end(), F2.mvKeysRight.begin(), F2.mvKeysRight.end());
        for (int i = 0; i < F2.N; ++i) std::memcpy(&df[(size_t)i * 32], F2.mDescriptors.ptr(i), 32);
        orc_frame_view vf = vk;
        vf.n = F2.N; vf.keys = reinterpret_cast<const orc_keypoint *>(allkeys.data()); vf.descriptors = df.data(); vf.u_right = nullptr;
        std::vector<int32_t> mf(F2.N);
        const int want = orc_search_by_bow_2cam(&vk, valid.data(), &vf, F2.Nleft, (int)kn.size(), kn.data(), kp.data(), ki.data(),
                                                (int)fn.size(), fn.data(), fp.data(), fi.data(), 0.7f, 1, mf.data());
        VS_GRAPHS::ORBmatcher matcher(0.7f, true);
        std::vector<MapPoint *> matches;
        const int got = matcher.SearchByBoW(&KF, F2, matches);
        EXPECT(got == want && got > 20, "SearchByBoW(two cameras): %d vs %d", got, want);
        int right_hits = 0;
        for (int j = 0; j < F2.N && j < (int)matches.size(); ++j) {
            EXPECT(matches[j] == (mf[j] >= 0 ? KF.mvpMapPoints[mf[j]] : nullptr), "two cameras: vpMapPointMatches[%d]", j);
            right_hits += j >= A.N && mf[j] >= 0;
        }
        EXPECT((int)matches.size() == F2.N && right_hits > 10, "SearchByBoW(two cameras): right-camera matches %d", right_hits);
    }

    // ---------------- SearchForInitialization ----------------
    {
        std::vector<cv::Point2f> prev(A.N);
        std::vector<float> prev_o((size_t)A.N * 2);
        for (int i = 0; i < A.N; ++i) { prev[i] = A.mvKeysUn[i].pt; prev_o[2 * i] = prev[i].x; prev_o[2 * i + 1] = prev[i].y; }
        std::vector<int32_t> m12(A.N);
        const int want = orc_search_for_initialization(&va, &vb, prev_o.data(), 100, 0.9f, 1, m12.data());
        VS_GRAPHS::ORBmatcher matcher(0.9f, true);
        std::vector<int> got12;
        const int got = matcher.SearchForInitialization(A, B, prev, got12, 100);
        EXPECT(got == want && got > 20, "SearchForInitialization: %d vs %d", got, want);
        for (int i = 0; i < A.N; ++i) {
            EXPECT(got12[i] == m12[i], "vnMatches12[%d]", i);
            EXPECT(prev[i].x == prev_o[2 * i] && prev[i].y == prev_o[2 * i + 1], "vbPrevMatched[%d]", i);
        }
    }

    // ================= keyframe-side methods =================
    // world points that project near B's keypoints shifted into A (camera at the given translation-only pose)
    auto make_points = [&](const Frame &src, const Pose &pose, float dx, float dy, std::vector<MapPoint> &store) {
        store.assign(src.N, MapPoint());
        for (int i = 0; i < src.N; ++i) {
            MapPoint &mp = store[i];
            const float z = 2.f + (rng() % 100) / 10.f;
            const float u = src.mvKeys[i].pt.x + dx + ((int)(rng() % 5) - 2) * 0.5f, v = src.mvKeys[i].pt.y + dy + ((int)(rng() % 5) - 2) * 0.5f;
            const Vec3 xc{{(u - 320) * z / 500, (v - 240) * z / 500, z}};
            mp.pos = xc - pose.t;
            mp.desc = src.mDescriptors.row(i).clone();
            mp.bad = rng() % 30 == 0;
            mp.obs = 1 + rng() % 3;
            // max distance so that PredictScale lands on the keypoint's octave (+-1)
            mp.max_dist = z * std::pow(1.2f, (float)src.mvKeys[i].octave + ((int)(rng() % 3) - 1) * 0.6f) * 0.999f;
            mp.min_dist = 0.05f;
        }
    };
    auto flat_fv = [](const std::map<unsigned, std::vector<unsigned>> &fv, std::vector<int32_t> &n, std::vector<int32_t> &p, std::vector<int32_t> &x) {
        p.push_back(0);
        for (auto &kv : fv) { n.push_back(kv.first); for (unsigned v : kv.second) x.push_back(v); p.push_back((int32_t)x.size()); }
    };
    // the caller-side projection of the reference loops, restated for the oracle's inputs
    auto project_points = [&](const std::vector<MapPoint *> &vp, const Pose &Tcw, const KeyFrame *kf, const Frame *fr, bool check_normal,
                              bool frame_bounds, std::vector<orc_search_point> &pts, std::vector<unsigned char> &desc) {
        pts.assign(vp.size(), orc_search_point());
        desc.assign(vp.size() * 32, 0);
        const Vec3 Ow = Tcw.inverse().translation();
        for (size_t i = 0; i < vp.size(); ++i) {
            orc_search_point &p = pts[i];
            std::memset(&p, 0, sizeof(p));
            MapPoint *mp = vp[i];
            if (!mp) continue;
            const Vec3 pc = Tcw * mp->pos;
            if (!frame_bounds && pc(2) < 0.0f) continue;
            const Vec2 uv = Camera().project(pc);
            if (frame_bounds) { if (uv(0) < 0 || uv(0) > 640 || uv(1) < 0 || uv(1) > 480) continue; }
            else if (!(uv(0) >= 0 && uv(0) < 640 && uv(1) >= 0 && uv(1) < 480)) continue;
            const Vec3 PO = mp->pos - Ow;
            const float dist = PO.norm();
            if (dist < mp->min_dist || dist > mp->max_dist) continue;
            if (check_normal && PO.dot(mp->normal) < 0.5 * dist) continue;
            p.level = kf ? mp->PredictScale(dist, kf) : mp->PredictScale(dist, fr);
            p.u = uv(0); p.v = uv(1); p.ur = uv(0) - 40.f * (1 / pc(2));
            p.valid = 1;
            std::memcpy(&desc[i * 32], mp->desc.ptr(0), 32);
        }
    };

    // ---------------- SearchByProjection(Cur, KF, sAlreadyFound, th, ORBdist): relocalisation ----------------
    {
        Frame Cur = A;
        Cur.mpCamera = &Cur.cam;
        Cur.pose = translation_pose(0.01f, 0.02f, 0.05f);
        Cur.mvpMapPoints.assign(Cur.N, nullptr);
        KeyFrame KF(B);
        std::vector<MapPoint> store;
        make_points(B, Cur.pose, -9, -5, store);
        for (MapPoint &mp : store) mp.normal = Vec3{{0, 0, 0}};
        std::set<MapPoint *> found;
        MapPoint pre;
        for (int i = 0; i < KF.N; ++i) {
            KF.mvpMapPoints[i] = rng() % 10 < 9 ? &store[i] : nullptr;
            if (rng() % 15 == 0) found.insert(&store[i]);
        }
        std::vector<unsigned char> occupied(Cur.N, 0);
        for (int i = 0; i < Cur.N; i += 13) { Cur.mvpMapPoints[i] = &pre; occupied[i] = 1; }
        std::vector<MapPoint *> vp = KF.mvpMapPoints;
        for (auto &q : vp) if (q && (q->bad || found.count(q))) q = nullptr;
        std::vector<orc_search_point> pts;
        std::vector<unsigned char> pdesc, dc;
        project_points(vp, Cur.pose, nullptr, &Cur, false, true, pts, pdesc);
        for (int i = 0; i < KF.N; ++i) pts[i].angle = KF.mvKeysUn[i].angle;
        orc_frame_view vc = view_of(Cur, dc);
        std::vector<int32_t> assign(Cur.N);
        const int want = orc_search_by_projection_reloc(&vc, occupied.data(), KF.N, pts.data(), pdesc.data(), 10.f, 100, 1, assign.data());
        VS_GRAPHS::ORBmatcher matcher(0.9f, true);
        const int got = matcher.SearchByProjection(Cur, &KF, found, 10.f, 100);
        EXPECT(got == want && got > 50, "SearchByProjection(reloc): %d vs %d", got, want);
        for (int i = 0; i < Cur.N; ++i) {
            MapPoint *expect = assign[i] >= 0 ? KF.mvpMapPoints[assign[i]] : (assign[i] == -2 ? nullptr : (occupied[i] ? &pre : nullptr));
            EXPECT(Cur.mvpMapPoints[i] == expect, "reloc mvpMapPoints[%d]", i);
        }
    }

    // ---------------- SearchByBoW(KF1, KF2) ----------------
    {
        KeyFrame K1(A), K2(B);
        std::vector<MapPoint> s1(K1.N), s2(K2.N);
        std::vector<unsigned char> v1(K1.N), v2(K2.N);
        for (int i = 0; i < K1.N; ++i) { s1[i].bad = rng() % 25 == 0; K1.mvpMapPoints[i] = rng() % 10 < 9 ? &s1[i] : nullptr; v1[i] = K1.mvpMapPoints[i] && !s1[i].bad; }
        for (int i = 0; i < K2.N; ++i) { s2[i].bad = rng() % 25 == 0; K2.mvpMapPoints[i] = rng() % 10 < 9 ? &s2[i] : nullptr; v2[i] = K2.mvpMapPoints[i] && !s2[i].bad; }
        std::vector<int32_t> n1, p1, i1, n2, p2, i2;
        flat_fv(K1.mFeatVec, n1, p1, i1);
        flat_fv(K2.mFeatVec, n2, p2, i2);
        std::vector<unsigned char> d1, d2;
        orc_frame_view w1 = view_of(K1, d1), w2 = view_of(K2, d2);
        std::vector<int32_t> m12(K1.N);
        const int want = orc_search_by_bow_kf(&w1, v1.data(), &w2, v2.data(), (int)n1.size(), n1.data(), p1.data(), i1.data(),
                                              (int)n2.size(), n2.data(), p2.data(), i2.data(), 0.8f, 1, m12.data());
        VS_GRAPHS::ORBmatcher matcher(0.8f, true);
        std::vector<MapPoint *> matches;
        const int got = matcher.SearchByBoW(&K1, &K2, matches);
        EXPECT(got == want && got > 10, "SearchByBoW(KF,KF): %d vs %d", got, want);
        for (int i = 0; i < K1.N && i < (int)matches.size(); ++i)
            EXPECT(matches[i] == (m12[i] >= 0 ? K2.mvpMapPoints[m12[i]] : nullptr), "vpMatches12[%d]", i);
        // two-camera keyframes (NLeft != -1): features past mvKeysUn (the right camera's) are skipped (:793-796, :814-817),
        // so right-camera entries in the FeatureVectors and map-point lists must not change the result
        K1.NLeft = K1.N; K2.NLeft = K2.N;
        for (auto &kv : K1.mFeatVec) kv.second.push_back(K1.N + kv.first % 7);
        for (auto &kv : K2.mFeatVec) kv.second.push_back(K2.N + kv.first % 5);
        K1.mvpMapPoints.resize(K1.N + 7, &s1[0]);
        K2.mvpMapPoints.resize(K2.N + 5, &s2[0]);
        const int got2 = matcher.SearchByBoW(&K1, &K2, matches);
        EXPECT(got2 == want && (int)matches.size() == K1.N + 7, "SearchByBoW(KF,KF) with two-camera keyframes: %d vs %d", got2, want);
        for (int i = 0; i < K1.N && i < (int)matches.size(); ++i)
            EXPECT(matches[i] == (m12[i] >= 0 ? K2.mvpMapPoints[m12[i]] : nullptr), "two cameras: vpMatches12[%d]", i);
    }

    // ---------------- Fuse(KF, vpMapPoints, th) ----------------
    {
        KeyFrame KF(A);
        KF.pose = translation_pose(0.f, 0.01f, 0.02f);
        std::vector<MapPoint> store, inkf(KF.N);
        make_points(B, KF.pose, -9, -5, store);
        for (int i = 0; i < KF.N; ++i) { inkf[i].obs = 1 + rng() % 4; inkf[i].bad = rng() % 20 == 0; KF.mvpMapPoints[i] = rng() % 3 == 0 ? &inkf[i] : nullptr; }
        std::vector<MapPoint *> vp(store.size());
        for (size_t i = 0; i < store.size(); ++i) vp[i] = rng() % 12 == 0 ? nullptr : &store[i];
        for (size_t i = 0; i + 1 < vp.size(); i += 17) vp[i + 1] = vp[i];                    // duplicates: the second is skipped by IsInKeyFrame
        std::vector<orc_search_point> pts;
        std::vector<unsigned char> pdesc, dk;
        project_points(vp, KF.pose, &KF, nullptr, true, false, pts, pdesc);
        orc_frame_view vk = view_of(KF, dk);
        std::vector<int32_t> best(vp.size());
        orc_fuse_search(&vk, (int)vp.size(), pts.data(), pdesc.data(), 3.f, KF.mvInvLevelSigma2.data(), 0, best.data());
        // expected bookkeeping, replayed on copies of the mock state
        std::vector<MapPoint> e_store = store, e_inkf = inkf;
        std::vector<MapPoint *> e_kfmp(KF.N);
        for (int i = 0; i < KF.N; ++i) e_kfmp[i] = KF.mvpMapPoints[i] ? &e_inkf[KF.mvpMapPoints[i] - &inkf[0]] : nullptr;
        int want = 0;
        for (size_t i = 0; i < vp.size(); ++i) {
            if (!vp[i]) continue;
            MapPoint *mp = &e_store[vp[i] - &store[0]];
            if (mp->bad || mp->IsInKeyFrame(&KF)) continue;
            if (best[i] < 0) continue;
            MapPoint *in = e_kfmp[best[i]];
            if (in) { if (!in->bad) { if (in->obs > mp->obs) mp->Replace(in); else in->Replace(mp); } }
            else { mp->AddObservation(&KF, best[i]); e_kfmp[best[i]] = mp; }
            ++want;
        }
        VS_GRAPHS::ORBmatcher matcher;
        const int got = matcher.Fuse(&KF, vp, 3.f);
        EXPECT(got == want && got > 20, "Fuse: %d vs %d", got, want);
        for (size_t i = 0; i < store.size(); ++i) EXPECT(store[i].bad == e_store[i].bad && store[i].obs == e_store[i].obs, "Fuse map point %zu state", i);
        for (int i = 0; i < KF.N; ++i) {
            const MapPoint *g = KF.mvpMapPoints[i], *e = e_kfmp[i];
            const bool same = (!g && !e) || (g && e && ((g >= &store[0] && g < &store[0] + store.size()) ? (e == &e_store[g - &store[0]]) : (e == &e_inkf[g - &inkf[0]])));
            EXPECT(same, "Fuse KF map point %d", i);
        }
    }

    // ---------------- Fuse(KF, vpMapPoints, th, bRight = true) on a two-camera keyframe ----------------
    {
        Frame F2 = A;                               // left camera = A's keypoints, right camera = B's
        F2.mvKeysRight = B.mvKeys;
        F2.N = A.N + B.N;
        F2.mDescriptors.create(F2.N, 32, CV_8U);
        for (int i = 0; i < A.N; ++i) std::memcpy(F2.mDescriptors.ptr(i), A.mDescriptors.ptr(i), 32);
        for (int i = 0; i < B.N; ++i) std::memcpy(F2.mDescriptors.ptr(A.N + i), B.mDescriptors.ptr(i), 32);
        F2.mvuRight.assign(F2.N, -1.f);
        F2.mvpMapPoints.assign(F2.N, nullptr);
        KeyFrame KF(F2);
        KF.NLeft = A.N;
        Camera cam2;
        KF.mpCamera2 = &cam2;
        KF.pose = translation_pose(0.3f, 0.2f, 0.1f);                                        // the left pose must not be used
        KF.right_pose = translation_pose(0.f, 0.01f, 0.02f);
        std::vector<MapPoint> store, inkf(F2.N);
        make_points(B, KF.right_pose, 0, 0, store);                                          // points in front of the right camera
        for (int i = 0; i < F2.N; ++i) { inkf[i].obs = 1 + rng() % 4; inkf[i].bad = rng() % 20 == 0; KF.mvpMapPoints[i] = rng() % 3 == 0 ? &inkf[i] : nullptr; }
        std::vector<MapPoint *> vp(store.size());
        for (size_t i = 0; i < store.size(); ++i) vp[i] = rng() % 12 == 0 ? nullptr : &store[i];
        std::vector<orc_search_point> pts;
        std::vector<unsigned char> pdesc, dk;
        project_points(vp, KF.right_pose, &KF, nullptr, true, false, pts, pdesc);
        Frame camR = B;
        camR.mvuRight.assign(B.N, -1.f);
        orc_frame_view vr = view_of(camR, dk);
        std::vector<int32_t> best(vp.size());
        orc_fuse_search(&vr, (int)vp.size(), pts.data(), pdesc.data(), 3.f, KF.mvInvLevelSigma2.data(), 0, best.data());
        std::vector<MapPoint> e_store = store, e_inkf = inkf;
        std::vector<MapPoint *> e_kfmp(F2.N);
        for (int i = 0; i < F2.N; ++i) e_kfmp[i] = KF.mvpMapPoints[i] ? &e_inkf[KF.mvpMapPoints[i] - &inkf[0]] : nullptr;
        int want = 0, right_slots = 0;
        for (size_t i = 0; i < vp.size(); ++i) {
            if (!vp[i]) continue;
            MapPoint *mp = &e_store[vp[i] - &store[0]];
            if (mp->bad || mp->IsInKeyFrame(&KF)) continue;
            if (best[i] < 0) continue;
            const int slot = best[i] + A.N;                                                  // idx += pKF->NLeft (:1295-1296)
            MapPoint *in = e_kfmp[slot];
            if (in) { if (!in->bad) { if (in->obs > mp->obs) mp->Replace(in); else in->Replace(mp); } }
            else { mp->AddObservation(&KF, slot); e_kfmp[slot] = mp; ++right_slots; }
            ++want;
        }
        VS_GRAPHS::ORBmatcher matcher;
        const int got = matcher.Fuse(&KF, vp, 3.f, true);
        EXPECT(got == want && got > 20 && right_slots > 5, "Fuse(bRight): %d vs %d (%d new right slots)", got, want, right_slots);
        for (size_t i = 0; i < store.size(); ++i) EXPECT(store[i].bad == e_store[i].bad && store[i].obs == e_store[i].obs, "Fuse(bRight) map point %zu state", i);
        for (int i = 0; i < F2.N; ++i) {
            const MapPoint *g = KF.mvpMapPoints[i], *e = e_kfmp[i];
            const bool same = (!g && !e) || (g && e && ((g >= &store[0] && g < &store[0] + store.size()) ? (e == &e_store[g - &store[0]]) : (e == &e_inkf[g - &inkf[0]])));
            EXPECT(same, "Fuse(bRight) KF map point %d", i);
        }
    }

    // ---------------- Sim3 family: SearchByProjection(KF, Scw, ...), SearchBySim3, Fuse(KF, Scw, ...) ----------------
    {
        KeyFrame KF(A);
        Sophus::Sim3f Scw;
        Scw.s = 2.f; Scw.t = Vec3{{0.02f, -0.02f, 0.04f}};                                  // Tcw = [I | t/s]
        const Pose Tcw(Mat3(), Scw.t / Scw.s);
        std::vector<MapPoint> store;
        make_points(B, Tcw, -9, -5, store);
        std::vector<MapPoint *> vp(store.size());
        for (size_t i = 0; i < store.size(); ++i) vp[i] = &store[i];
        std::vector<MapPoint *> vpMatched(KF.N, nullptr);
        MapPoint pre;
        for (int i = 0; i < KF.N; i += 9) vpMatched[i] = &pre;
        for (int i = 4; i < KF.N; i += 31) vpMatched[i] = vp[i % vp.size()];                // already found points
        std::set<MapPoint *> already(vpMatched.begin(), vpMatched.end());
        std::vector<MapPoint *> vq = vp;
        for (auto &q : vq) if (q->bad || already.count(q)) q = nullptr;
        std::vector<orc_search_point> pts;
        std::vector<unsigned char> pdesc, dk;
        project_points(vq, Tcw, &KF, nullptr, true, false, pts, pdesc);
        orc_frame_view vk = view_of(KF, dk);
        std::vector<unsigned char> matched(KF.N);
        for (int i = 0; i < KF.N; ++i) matched[i] = vpMatched[i] != nullptr;
        std::vector<int32_t> assign(KF.N);
        const int want = orc_search_by_projection_sim3(&vk, matched.data(), (int)vq.size(), pts.data(), pdesc.data(), 8, 1.5f, assign.data());
        const std::vector<MapPoint *> before = vpMatched;
        std::vector<KeyFrame *> pkfs(vp.size(), &KF), matched_kf(KF.N, nullptr);
        std::vector<MapPoint *> vpMatched2 = vpMatched;
        VS_GRAPHS::ORBmatcher matcher;
        const int got = matcher.SearchByProjection(&KF, Scw, vp, vpMatched, 8, 1.5f);
        const int got2 = matcher.SearchByProjection(&KF, Scw, vp, pkfs, vpMatched2, matched_kf, 8, 1.5f);
        EXPECT(got == want && got2 == want && got > 50, "SearchByProjection(Sim3): %d / %d vs %d", got, got2, want);
        for (int i = 0; i < KF.N; ++i) {
            EXPECT(vpMatched[i] == (assign[i] >= 0 ? vp[assign[i]] : before[i]) && vpMatched2[i] == vpMatched[i], "vpMatched[%d]", i);
            EXPECT(matched_kf[i] == (assign[i] >= 0 ? &KF : nullptr), "vpMatchedKF[%d]", i);
        }
        // Fuse(KF, Scw, vpPoints, th, vpReplacePoint)
        std::vector<MapPoint> inkf(KF.N);
        for (int i = 0; i < KF.N; ++i) { inkf[i].bad = rng() % 20 == 0; KF.mvpMapPoints[i] = rng() % 3 == 0 ? &inkf[i] : nullptr; }
        std::vector<MapPoint *> vr = vp;
        project_points(vr, Tcw, &KF, nullptr, true, false, pts, pdesc);
        std::vector<int32_t> best(vr.size());
        orc_fuse_search(&vk, (int)vr.size(), pts.data(), pdesc.data(), 4.f, nullptr, 1, best.data());
        std::vector<MapPoint *> e_kfmp = KF.mvpMapPoints, e_replace(vr.size(), nullptr);
        int wantf = 0;
        for (size_t i = 0; i < vr.size(); ++i) {
            if (vr[i]->bad || best[i] < 0) continue;
            MapPoint *in = e_kfmp[best[i]];
            if (in) { if (!in->bad) e_replace[i] = in; }
            else e_kfmp[best[i]] = vr[i];
            ++wantf;
        }
        std::vector<MapPoint *> replace(vr.size(), nullptr);
        const int gotf = matcher.Fuse(&KF, Scw, vr, 4.f, replace);
        EXPECT(gotf == wantf && gotf > 50, "Fuse(Sim3): %d vs %d", gotf, wantf);
        EXPECT(replace == e_replace && KF.mvpMapPoints == e_kfmp, "Fuse(Sim3) outputs");
    }
    {
        // SearchBySim3: K1 = A at the origin, K2 = B; S12 maps camera-2 coordinates to camera-1 (translation + scale 1)
        KeyFrame K1(A), K2(B);
        K1.pose = translation_pose(0, 0, 0);
        K2.pose = translation_pose(0.03f, 0.01f, 0.f);
        Sophus::Sim3f S12;
        S12.t = Vec3{{-0.03f, -0.01f, 0.f}};
        const Sophus::Sim3f S21 = S12.inverse();
        std::vector<MapPoint> s1, s2;
        make_points(A, translation_pose(0, 0, 0), 0, 0, s1);        // seen in K1 at A's keypoints
        make_points(B, translation_pose(0, 0, 0), -9, -5, s2);      // K2's points: land at A's coordinates in camera 1
        for (int i = 0; i < K1.N; ++i) K1.mvpMapPoints[i] = rng() % 10 < 8 ? &s1[i] : nullptr;
        for (int i = 0; i < K2.N; ++i) K2.mvpMapPoints[i] = rng() % 10 < 8 ? &s2[i] : nullptr;
        // make K1's points project into K2 where B's keypoints are: shift their world position
        for (int i = 0; i < K1.N; ++i) {
            const float z = s1[i].pos.v[2];
            s1[i].pos = Vec3{{(A.mvKeys[i].pt.x + 9 - 320) * z / 500, (A.mvKeys[i].pt.y + 5 - 240) * z / 500, z}} - S21.t;
        }
        std::vector<MapPoint *> vpMatches12(K1.N, nullptr);
        for (int i = 0; i < K1.N; i += 23) if (K2.mvpMapPoints[i % K2.N]) { vpMatches12[i] = K2.mvpMapPoints[i % K2.N]; vpMatches12[i]->index_in[&K2] = i % K2.N; }
        std::vector<bool> am1(K1.N, false), am2(K2.N, false);
        for (int i = 0; i < K1.N; ++i) if (vpMatches12[i]) { am1[i] = true; am2[std::get<0>(vpMatches12[i]->GetIndexInKeyFrame(&K2))] = true; }
        auto proj = [&](const std::vector<MapPoint *> &vp, const std::vector<bool> &am, const Pose &Tw, const Sophus::Sim3f &S, KeyFrame *tgt,
                        std::vector<orc_search_point> &pts, std::vector<unsigned char> &desc) {
            pts.assign(vp.size(), orc_search_point());
            desc.assign(vp.size() * 32, 0);
            for (size_t i = 0; i < vp.size(); ++i) {
                std::memset(&pts[i], 0, sizeof(orc_search_point));
                MapPoint *mp = vp[i];
                if (!mp || am[i] || mp->bad) continue;
                const Vec3 pb = S * (Tw * mp->pos);
                if (pb(2) < 0.0) continue;
                const float invz = 1.0 / pb(2);
                const float x = pb(0) * invz, y = pb(1) * invz;
                const float u = 500 * x + 320, v = 500 * y + 240;
                if (!tgt->IsInImage(u, v)) continue;
                const float d = pb.norm();
                if (d < mp->min_dist || d > mp->max_dist) continue;
                pts[i].level = mp->PredictScale(d, tgt);
                pts[i].u = u; pts[i].v = v; pts[i].valid = 1;
                std::memcpy(&desc[i * 32], mp->desc.ptr(0), 32);
            }
        };
        std::vector<orc_search_point> p1, p2;
        std::vector<unsigned char> e1, e2, d1, d2;
        proj(K1.mvpMapPoints, am1, K1.pose, S21, &K2, p1, e1);
        proj(K2.mvpMapPoints, am2, K2.pose, S12, &K1, p2, e2);
        orc_frame_view w1 = view_of(K1, d1), w2 = view_of(K2, d2);
        std::vector<int32_t> m12(K1.N);
        const int want = orc_search_by_sim3(&w1, &w2, p1.data(), e1.data(), p2.data(), e2.data(), 7.5f, m12.data());
        const std::vector<MapPoint *> before = vpMatches12;
        VS_GRAPHS::ORBmatcher matcher;
        const int got = matcher.SearchBySim3(&K1, &K2, vpMatches12, S12, 7.5f);
        EXPECT(got == want && got > 30, "SearchBySim3: %d vs %d", got, want);
        for (int i = 0; i < K1.N; ++i) EXPECT(vpMatches12[i] == (m12[i] >= 0 ? K2.mvpMapPoints[m12[i]] : before[i]), "SearchBySim3 vpMatches12[%d]", i);
    }

    // ---------------- SearchForTriangulation ----------------
    {
        KeyFrame K1(A), K2(B);
        K1.pose = translation_pose(0, 0, 0);
        K2.pose = translation_pose(-0.09f, -0.05f, 0.f);      // pure sideways motion: epipolar lines along (9, 5)
        std::vector<MapPoint> s1(K1.N), s2(K2.N);
        std::vector<unsigned char> h1(K1.N), h2(K2.N);
        for (int i = 0; i < K1.N; ++i) { K1.mvpMapPoints[i] = rng() % 10 < 3 ? &s1[i] : nullptr; h1[i] = K1.mvpMapPoints[i] != nullptr; }
        for (int i = 0; i < K2.N; ++i) { K2.mvpMapPoints[i] = rng() % 10 < 3 ? &s2[i] : nullptr; h2[i] = K2.mvpMapPoints[i] != nullptr; }
        std::vector<int32_t> n1, p1, i1, n2, p2, i2;
        flat_fv(K1.mFeatVec, n1, p1, i1);
        flat_fv(K2.mFeatVec, n2, p2, i2);
        // the same F12 / epipole the shim derives (Pinhole.cpp:120-124, ORBmatcher.cc:908-914) with the stand-in algebra
        const Pose T12 = K1.pose * K2.pose.inverse();
        const Vec3 t12 = T12.translation();
        Mat3 t12x;
        t12x << 0, -t12(2), t12(1), t12(2), 0, -t12(0), -t12(1), t12(0), 0;
        const Mat3 K = Camera().toK_();
        const Mat3 F12 = K.transpose().inverse() * t12x * T12.rotationMatrix() * K.inverse();
        const Vec3 C2 = K2.pose * K1.GetCameraCenter();
        const Vec2 ep = Camera().project(C2);
        const float epf[2] = {ep(0), ep(1)};
        std::vector<unsigned char> d1, d2;
        orc_frame_view w1 = view_of(K1, d1), w2 = view_of(K2, d2);
        for (int coarse = 0; coarse < 2; ++coarse) {
            std::vector<int32_t> m12(K1.N);
            const int want = orc_search_for_triangulation(&w1, h1.data(), &w2, h2.data(), (int)n1.size(), n1.data(), p1.data(), i1.data(),
                                                          (int)n2.size(), n2.data(), p2.data(), i2.data(), 0, coarse, F12.m, epf,
                                                          K2.mvLevelSigma2.data(), 1, m12.data());
            VS_GRAPHS::ORBmatcher matcher(0.6f, true);
            std::vector<std::pair<size_t, size_t>> pairs;
            const int got = matcher.SearchForTriangulation(&K1, &K2, pairs, false, coarse != 0);
            EXPECT(got == want && got > 5 && (int)pairs.size() == got, "SearchForTriangulation(coarse=%d): %d vs %d (%zu pairs)", coarse, got, want, pairs.size());
            size_t k = 0;
            for (int i = 0; i < K1.N; ++i)
                if (m12[i] >= 0) { EXPECT(k < pairs.size() && pairs[k].first == (size_t)i && pairs[k].second == (size_t)m12[i], "pair %zu", k); ++k; }
        }
    }

    // ---------------- FrameOps: UndistortKeyPoints, ComputeImageBounds, ComputeStereoFromRGBD (Frame.cc:891-955, 1129-1150) ----
    {
        float Kdata[9] = {517.306408f, 0.f, 318.643040f, 0.f, 516.469215f, 255.313989f, 0.f, 0.f, 1.f};   // TUM1.yaml
        float Ddata[5] = {0.262383f, -0.953104f, -0.005358f, 0.002628f, 1.163314f};
        cv::Mat mK(3, 3, CV_8UC1, Kdata, 3 * sizeof(float)), mDist(5, 1, CV_8UC1, Ddata, sizeof(float));
        cv::Mat m(480, 640, CV_8UC1, img_a.data(), 640);
        std::vector<cv::KeyPoint> keys, keysUn;
        cv::Mat desc;
        std::vector<int> lap = {0, 0};
        ex(m, cv::noArray(), keys, desc, lap);
        VS_GRAPHS::frame_ops::UndistortKeyPoints(keys, mK, mDist, keysUn);
        const int n = (int)keys.size();
        std::vector<float> xy(2 * (size_t)n), want(2 * (size_t)n);
        for (int i = 0; i < n; ++i) { xy[2 * i] = keys[i].pt.x; xy[2 * i + 1] = keys[i].pt.y; }
        const double dist[5] = {Ddata[0], Ddata[1], Ddata[2], Ddata[3], Ddata[4]};
        orc_undistort_points(n, xy.data(), Kdata[0], Kdata[4], Kdata[2], Kdata[5], dist, 5, want.data());
        int bad = 0, moved = 0;
        for (int i = 0; i < n; ++i) {
            bad += keysUn[i].pt.x != want[2 * i] || keysUn[i].pt.y != want[2 * i + 1] || keysUn[i].octave != keys[i].octave ||
                   keysUn[i].angle != keys[i].angle;
            moved += keysUn[i].pt.x != keys[i].pt.x;
        }
        EXPECT((int)keysUn.size() == n && bad == 0 && moved > n / 2, "UndistortKeyPoints: %d of %d differ (%d moved)", bad, n, moved);
        float c[8] = {0.f, 0.f, 640.f, 0.f, 0.f, 480.f, 640.f, 480.f}, wc[8];
        orc_undistort_points(4, c, Kdata[0], Kdata[4], Kdata[2], Kdata[5], dist, 5, wc);
        float minx, maxx, miny, maxy;
        VS_GRAPHS::frame_ops::ComputeImageBounds(640, 480, mK, mDist, minx, maxx, miny, maxy);
        EXPECT(minx == std::min(wc[0], wc[4]) && maxx == std::max(wc[2], wc[6]) && miny == std::min(wc[1], wc[3]) &&
                   maxy == std::max(wc[5], wc[7]), "ComputeImageBounds %f %f %f %f", minx, maxx, miny, maxy);
        float Zdata[4] = {0.f, 0.f, 0.f, 0.f};
        cv::Mat mZero(4, 1, CV_8UC1, Zdata, sizeof(float));
        VS_GRAPHS::frame_ops::ComputeImageBounds(640, 480, mK, mZero, minx, maxx, miny, maxy);
        EXPECT(minx == 0.f && maxx == 640.f && miny == 0.f && maxy == 480.f, "ComputeImageBounds without distortion");
        std::vector<float> depth(640 * 480);
        std::uniform_real_distribution<float> du(-0.5f, 6.f);
        for (auto &d : depth) d = du(rng);
        std::vector<float> ur, dz, wur(n), wdz(n);
        VS_GRAPHS::frame_ops::ComputeStereoFromRGBD(keys, keysUn, depth.data(), 640, 40.f, ur, dz);
        orc_stereo_from_rgbd(n, xy.data(), want.data(), depth.data(), 640, 40.f, wur.data(), wdz.data());
        EXPECT(ur == wur && dz == wdz, "ComputeStereoFromRGBD");
    }

    // ---------------- rectification ahead of the extractor (System.cc:284-292) ----------------
    {
        const int w = 624, h = 464;
        std::vector<float> mx((size_t)w * h), my((size_t)w * h);
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) {
                const float xn = (x - w / 2.f) / 400.f, yn = (y - h / 2.f) / 400.f, f = 1.f + 0.15f * (xn * xn + yn * yn);
                mx[(size_t)y * w + x] = xn * f * 400.f + 320.f + 1.25f;
                my[(size_t)y * w + x] = yn * f * 400.f + 240.f - 0.75f;
            }
        VS_GRAPHS::ORBextractor exr(1000, 1.2f, 8, 20, 7);
        exr.SetRectification(mx.data(), my.data(), w, h);
        cv::Mat m(480, 640, CV_8UC1, img_a.data(), 640);
        std::vector<cv::KeyPoint> kps;
        cv::Mat desc;
        std::vector<int> lap = {0, 0};
        exr(m, cv::noArray(), kps, desc, lap);
        std::vector<unsigned char> rect((size_t)w * h);
        orc_remap_bilinear(img_a.data(), 640, 480, 640, mx.data(), my.data(), w, h, rect.data(), w);
        orc_extractor *orc2 = orc_extractor_create(1000, 1.2f, 8, 20, 7);
        orc_extract(orc2, rect.data(), w, h, w, 0, 0);
        const int n = orc_num_keypoints(orc2);
        std::vector<orc_keypoint> okps(n);
        std::vector<unsigned char> odesc((size_t)n * 32);
        orc_get_keypoints(orc2, okps.data(), odesc.data());
        int bad = (int)kps.size() != n;
        for (int i = 0; i < n && i < (int)kps.size(); ++i)
            bad += kps[i].pt.x != okps[i].x || kps[i].pt.y != okps[i].y || kps[i].octave != okps[i].octave ||
                   std::memcmp(desc.ptr(i), &odesc[(size_t)i * 32], 32) != 0;
        long diff = 0;
        const cv::Mat &L0 = exr.mvImagePyramid[0];
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) diff += L0.ptr(y)[x] != rect[(size_t)y * w + x];
        EXPECT(bad == 0 && n > 500 && diff == 0 && L0.cols == w && L0.rows == h, "rectified extraction: %d bad of %d, %ld level-0 pixels differ", bad, n, diff);
        orc_extractor_destroy(orc2);
    }

    orc_extractor_destroy(orc);
    std::printf(g_fail ? "SHIM TEST FAILED (%d)\n" : "SHIM TEST OK\n", g_fail);
    return g_fail ? 1 : 0;
}
