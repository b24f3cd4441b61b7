// match_internal.cuh — declarations shared by match_methods.cu and match_methods_kf.cu (not part of the ABI).
#pragma once
#include <vector>

#include "vsg_internal.cuh"

struct vsg_frame {
    int device = 0;        // the frame may outlive the matcher that uploaded it
    int n = 0, cols = 0, rows = 0, n_levels = 0;
    float min_x = 0, min_y = 0, inv_w = 0, inv_h = 0;
    bool has_right = false;
    // device: one stream-ordered allocation (`block`), carved into the arrays below
    void *block = nullptr;
    float2 *xy = nullptr;
    int *octave = nullptr;
    float *u_right = nullptr;
    uint8_t *desc = nullptr;
    int *cell_ptr = nullptr, *cell_idx = nullptr;
    // host copies the resolve loops read
    std::vector<vsg_keypoint> keys;
    std::vector<float> scale;
    std::vector<float> u_right_h;   // mvuRight (empty if monocular)
};

namespace vsg {

#ifndef CK
#endif

constexpr int TH_HIGH = 100, TH_LOW = 50, HISTO_LENGTH = 30;   // ORBmatcher.cc:34-36

struct FrameDev {
    int n, cols, rows;
    float min_x, min_y, inv_w, inv_h;
    const float2 *xy;
    const int *octave;
    const float *u_right;   // nullptr if monocular
    const uint4 *desc;
    const int *cell_ptr, *cell_idx;
};

struct AreaQuery {          // one GetFeaturesInArea call + the per-candidate stereo gate of the caller
    float x, y, r;
    int min_level, max_level;
    float xr, rr;           // right-image gate: skip if u_right[idx] > 0 && |xr - u_right[idx]| > rr; rr < 0 disables it
    int max_dist = 256;     // candidates farther than this are not listed: the caller proved they cannot change its result
    int desc_idx = -1;      // row of the uploaded query descriptors (-1: the query's own index)
};

// Candidate lists as the kernel left them: query q's (idx, dist) pairs are raw[off[q] .. off[q] + cnt[q]) (segments in
// arbitrary order).  The arrays live in the matcher's pinned staging and stay valid until its next search.
struct Chi2Gate {            // inverse level variances for the reprojection gates of the pose-based Fuse (:1273-1299)
    float inv_sigma2[kMaxLevels];
};

struct AreaLists {
    const int2 *raw = nullptr;
    const int *off = nullptr, *cnt = nullptr;
};


// Runs the area search (Frame::GetFeaturesInArea + Hamming distance, area_search_kernel) for nq queries and brings
// the lists back: ptr[nq+1] and (idx, dist) pairs in query order, each list in the reference's candidate order.
vsg_status area_search(vsg_matcher *m, const vsg_frame *f, int nq, const AreaQuery *qs, const uint8_t *qdesc,
                       std::vector<int> &ptr, std::vector<int2> &ent);
// the same without the re-pack into query order; n_qdesc rows of qdesc are uploaded (queries pick theirs by desc_idx)
// Best candidate per query on the device (window / level gates, optional chi-square gates when inv_sigma2 != nullptr): the
// whole search of the methods whose queries are independent of one another.  out[k] = (index or -1, distance or INT_MAX).
vsg_status area_best(vsg_matcher *m, const vsg_frame *f, int nq, const AreaQuery *qs, const uint8_t *qdesc, const float *inv_sigma2,
                     std::vector<int2> &out);
vsg_status area_search_raw(vsg_matcher *m, const vsg_frame *f, int nq, const AreaQuery *qs, const uint8_t *qdesc,
                           int n_qdesc, AreaLists *out);
// ORBmatcher::ComputeThreeMaxima (ORBmatcher.cc:2002-2043)
void three_maxima(const std::vector<int> *hist, int L, int &ind1, int &ind2, int &ind3);
// rotation-histogram bin (ORBmatcher.cc:351-358)
int rot_bin(float a1, float a2);

}  // namespace vsg
