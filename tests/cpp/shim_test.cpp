// shim_test.cpp — exercises the drop-in C++ classes (visual_sgraphs_b200/shim) exactly as the reference's
// callers would (Frame::ExtractORB, Tracking::SearchLocalPoints / TrackWithMotionModel / TrackReferenceKeyFrame /
// MonocularInitialization), with small stand-ins for Frame / KeyFrame / MapPoint that carry the members the
// reference's matcher reads, and checks every result against the CPU oracle.  Needs a GPU.
//   usage: shim_test frame_a.raw frame_b.raw   (two 640x480 8-bit frames written by the pytest wrapper)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <random>
#include <vector>

#include "../../oracle/oracle.h"
#include "../../visual_sgraphs_b200/shim/ORBextractor.h"
#include "../../visual_sgraphs_b200/shim/ORBmatcher.h"

static int g_fail = 0;
#define EXPECT(cond, ...)                                  \
    do {                                                   \
        if (!(cond)) { std::printf("FAIL %s:%d: ", __FILE__, __LINE__); std::printf(__VA_ARGS__); std::printf("\n"); ++g_fail; } \
    } while (0)

// ---- stand-ins for the reference's types (only what ORBmatcher reads) ----
struct Vec3 {
    float v[3];
    float operator()(int i) const { return v[i]; }
};
struct Vec2 {
    float v[2];
    float operator()(int i) const { return v[i]; }
};
struct Pose {   // translation-only SE3
    Vec3 t;
    Pose inverse() const { return Pose{{{-t.v[0], -t.v[1], -t.v[2]}}}; }
    Vec3 translation() const { return t; }
    Vec3 operator*(const Vec3 &p) const { return Vec3{{p.v[0] + t.v[0], p.v[1] + t.v[1], p.v[2] + t.v[2]}}; }
};
struct Camera {
    float fx = 500, fy = 500, cx = 320, cy = 240;
    Vec2 project(const Vec3 &p) const { return Vec2{{fx * p.v[0] / p.v[2] + cx, fy * p.v[1] / p.v[2] + cy}}; }
};
struct MapPoint {
    bool mbTrackInView = false, mbTrackInViewR = false;
    float mTrackProjX = 0, mTrackProjY = 0, mTrackProjXR = 0, mTrackViewCos = 1, mTrackDepth = 1;
    int mnTrackScaleLevel = 0;
    bool bad = false;
    int obs = 1;
    cv::Mat desc;
    Vec3 pos{{0, 0, 1}};
    bool isBad() const { return bad; }
    int Observations() const { return obs; }
    cv::Mat GetDescriptor() const { return desc.clone(); }
    Vec3 GetWorldPos() const { return pos; }
};
struct Frame {
    int N = 0, Nleft = -1;
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
    cv::Mat mDescriptors;
    std::vector<float> mvuRight;
    std::vector<float> mvScaleFactors;
    std::vector<MapPoint *> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    std::map<unsigned, std::vector<unsigned>> mFeatVec;
    float mnMinX = 0, mnMinY = 0, mnMaxX = 640, mnMaxY = 480;
    float mfGridElementWidthInv = 64.f / 640.f, mfGridElementHeightInv = 48.f / 480.f;
    float mb = 0.1f, mbf = 40.f;
    Pose pose{{{0, 0, 0}}};
    Camera cam;
    Camera *mpCamera = &cam;
    Pose GetPose() const { return pose; }
    std::vector<MapPoint *> GetMapPointMatches() const { return mvpMapPoints; }   // KeyFrame interface
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

    // ---------------- SearchByProjection(Cur, Last) ----------------
    {
        Frame Cur = A, Last = B;
        Cur.mpCamera = &Cur.cam; Last.mpCamera = &Last.cam;
        Cur.mvpMapPoints.assign(Cur.N, nullptr);
        Cur.pose = Pose{{{0.02f, -0.01f, -0.3f}}};    // tlc = Tlw * (-t): z = +0.3 > mb -> forward
        Last.pose = Pose{{{0, 0, 0}}};
        std::vector<MapPoint> store(Last.N);
        std::vector<orc_proj_point> pts(Last.N);
        std::vector<unsigned char> pdesc((size_t)Last.N * 32, 0);
        for (int i = 0; i < Last.N; ++i) {
            MapPoint &mp = store[i];
            const float z = 2.f + (rng() % 100) / 10.f;
            // a world point whose projection in Cur lands near the shifted keypoint
            const float u = Last.mvKeys[i].pt.x - 9, v = Last.mvKeys[i].pt.y - 5;
            const Vec3 xc{{(u - 320) * z / 500, (v - 240) * z / 500, z}};
            mp.pos = Vec3{{xc.v[0] - Cur.pose.t.v[0], xc.v[1] - Cur.pose.t.v[1], xc.v[2] - Cur.pose.t.v[2]}};
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

    orc_extractor_destroy(orc);
    std::printf(g_fail ? "SHIM TEST FAILED (%d)\n" : "SHIM TEST OK\n", g_fail);
    return g_fail ? 1 : 0;
}
