// ref_capi.cpp — C entry points around the REFERENCE's own VS_GRAPHS::ORBextractor, compiled unmodified from
// /root/reference/orb_slam3/src/ORBextractor.cc (see Makefile).  TEST INFRASTRUCTURE: used by tests/ (parity pin
// of the oracle port and of the CUDA path) and by bench.py's cpu_baseline / --impl reference legs.
//
// A derived class reaches the protected stages (ComputePyramid, ComputeKeyPointsOctTree, DistributeOctTree,
// tables) so that tests can compare intermediates, not only the final keypoints.
#include <opencv2/opencv.hpp>

#include <cstdint>
#include <cstring>
#include <thread>
#include <chrono>
#include <vector>
#include <atomic>

#include "ORBextractor.h"  // the reference's header: /root/reference/orb_slam3/include/ORBextractor.h

namespace {

struct RefExtractor : public VS_GRAPHS::ORBextractor {
    using VS_GRAPHS::ORBextractor::ORBextractor;
    std::vector<cv::KeyPoint> keys;
    cv::Mat desc;

    int run(const uint8_t *img, int w, int h, int pitch, int lap0, int lap1) {
        cv::Mat image(h, w, CV_8UC1, (void *)img, (size_t)pitch);
        if (w == 0 || h == 0) image = cv::Mat();
        std::vector<int> lap = {lap0, lap1};
        desc = cv::Mat();
        return (*this)(image, cv::Mat(), keys, desc, lap);
    }
    // stage access (protected members of the reference class)
    void pyramid_only(const cv::Mat &image) { ComputePyramid(image); }
    void keypoints_only(std::vector<std::vector<cv::KeyPoint>> &all) { ComputeKeyPointsOctTree(all); }
    std::vector<cv::KeyPoint> octree(const std::vector<cv::KeyPoint> &in, int minX, int maxX, int minY, int maxY, int N, int level) {
        return DistributeOctTree(in, minX, maxX, minY, maxY, N, level);
    }
    const std::vector<int> &quotas() const { return mnFeaturesPerLevel; }
    const std::vector<int> &umax_table() const { return umax; }
};

}  // namespace

extern "C" {

struct ref_keypoint { float x, y, size, angle, response; int32_t octave, class_id; };
static_assert(sizeof(ref_keypoint) == sizeof(cv::KeyPoint), "cv::KeyPoint layout");

void *ref_extractor_create(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th) {
    return new RefExtractor(nfeatures, scale_factor, nlevels, ini_th, min_th);
}
void ref_extractor_destroy(void *ex) { delete (RefExtractor *)ex; }

int ref_levels(void *ex) { return ((RefExtractor *)ex)->GetLevels(); }
void ref_tables(void *ex, float *scale, float *inv_scale, float *sigma2, float *inv_sigma2, int32_t *quota, int32_t *umax16) {
    RefExtractor *e = (RefExtractor *)ex;
    const int n = e->GetLevels();
    std::vector<float> a = e->GetScaleFactors(), b = e->GetInverseScaleFactors(), c = e->GetScaleSigmaSquares(),
                       d = e->GetInverseScaleSigmaSquares();
    for (int i = 0; i < n; ++i) { scale[i] = a[i]; inv_scale[i] = b[i]; sigma2[i] = c[i]; inv_sigma2[i] = d[i]; quota[i] = e->quotas()[i]; }
    for (int i = 0; i < 16; ++i) umax16[i] = e->umax_table()[i];
}

// ORBextractor::operator() — returns monoIndex (or -1); results stay in the object.
int ref_extract(void *ex, const uint8_t *img, int w, int h, int pitch, int lap0, int lap1) {
    return ((RefExtractor *)ex)->run(img, w, h, pitch, lap0, lap1);
}
int ref_num_keypoints(void *ex) { return (int)((RefExtractor *)ex)->keys.size(); }
void ref_get_keypoints(void *ex, ref_keypoint *kps, uint8_t *desc) {
    RefExtractor *e = (RefExtractor *)ex;
    const int n = (int)e->keys.size();
    if (n) std::memcpy(kps, e->keys.data(), (size_t)n * sizeof(ref_keypoint));
    for (int i = 0; i < n; ++i) std::memcpy(desc + (size_t)i * 32, e->desc.ptr(i), 32);
}
// mvImagePyramid[level] after the last call (public member, ORBextractor.h:93)
void ref_level_size(void *ex, int level, int32_t *w, int32_t *h) {
    const cv::Mat &m = ((RefExtractor *)ex)->mvImagePyramid[level];
    *w = m.cols; *h = m.rows;
}
// border = 0: the level ROI; border = 19: the ROI with its EDGE_THRESHOLD frame (the memory around the view)
void ref_get_level(void *ex, int level, int border, uint8_t *dst) {
    const cv::Mat &m = ((RefExtractor *)ex)->mvImagePyramid[level];
    const int W = m.cols + 2 * border;
    for (int y = -border; y < m.rows + border; ++y)
        std::memcpy(dst + (size_t)(y + border) * W, m.data + (ptrdiff_t)y * (ptrdiff_t)m.step - border, (size_t)W);
}

// Stage runs: pyramid, then ComputeKeyPointsOctTree alone (per-level keypoints in level coordinates, with angles).
int ref_level_keypoints(void *ex, const uint8_t *img, int w, int h, int pitch, int level, ref_keypoint *out, int cap) {
    RefExtractor *e = (RefExtractor *)ex;
    cv::Mat image(h, w, CV_8UC1, (void *)img, (size_t)pitch);
    e->pyramid_only(image);
    std::vector<std::vector<cv::KeyPoint>> all;
    e->keypoints_only(all);
    const int n = (int)all[level].size();
    for (int i = 0; i < n && i < cap; ++i) std::memcpy(&out[i], &all[level][i], sizeof(ref_keypoint));
    return n;
}

// DistributeOctTree alone: candidates (x, y, response) in the given order -> (x, y, response) of the survivors.
int ref_distribute_octree(void *ex, const float *xyr, int n, int min_x, int max_x, int min_y, int max_y, int quota,
                          float *out_xyr, int cap) {
    std::vector<cv::KeyPoint> in((size_t)n);
    for (int i = 0; i < n; ++i) in[i] = cv::KeyPoint(xyr[3 * i], xyr[3 * i + 1], 7.f, -1, xyr[3 * i + 2]);
    std::vector<cv::KeyPoint> out = ((RefExtractor *)ex)->octree(in, min_x, max_x, min_y, max_y, quota, 0);
    const int m = (int)out.size();
    for (int i = 0; i < m && i < cap; ++i) { out_xyr[3 * i] = out[i].pt.x; out_xyr[3 * i + 1] = out[i].pt.y; out_xyr[3 * i + 2] = out[i].response; }
    return m;
}

// Throughput of the reference extractor on `threads` host threads (one extractor instance per thread, frames
// dealt round-robin) — the CPU baseline bench.py reports. Returns seconds.
double ref_bench_extract(const uint8_t *frames, int nframes, int w, int h, int nfeatures, float scale_factor, int nlevels,
                         int ini_th, int min_th, int threads, int64_t *total_keypoints) {
    if (threads < 1) threads = 1;
    std::atomic<int64_t> total(0);
    std::atomic<int> next(0);
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([&]() {
            RefExtractor ex(nfeatures, scale_factor, nlevels, ini_th, min_th);
            int64_t mine = 0;
            for (;;) {
                const int f = next.fetch_add(1);
                if (f >= nframes) break;
                ex.run(frames + (size_t)f * w * h, w, h, w, 0, 0);
                mine += (int64_t)ex.keys.size();
            }
            total += mine;
        });
    for (auto &th : pool) th.join();
    auto t1 = std::chrono::steady_clock::now();
    if (total_keypoints) *total_keypoints = total.load();
    return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
