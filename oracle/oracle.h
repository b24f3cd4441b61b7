/*
 * oracle/oracle.h — C interface of the CPU ORACLE (test infrastructure, NOT product code).
 *
 * The oracle is a CPU restatement of the reference's ORB front-end
 * (reference = snt-arg/visual_sgraphs, paths relative to /root/reference):
 *   orb_slam3/src/ORBextractor.cc   (pyramid, per-cell FAST, oct-tree, IC_Angle, rBRIEF, operator())
 *   orb_slam3/src/ORBmatcher.cc     (DescriptorDistance, Search* loops, ComputeThreeMaxima)
 *   orb_slam3/src/Frame.cc          (GetFeaturesInArea, ComputeStereoMatches)
 * with the un-vendored OpenCV primitives restated bit-exactly against cv2 4.13.0
 * (SURVEY.md Appendix A).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  The product library (libvsg_cuda.so) never links or calls it.
 *
 * Parity status: the reference has no tests, golden vectors or fixtures for this path and cannot be
 * compiled here (needs OpenCV C++, Eigen, PCL, boost).  The oracle is therefore pinned to
 *   (a) real OpenCV (python cv2 4.13.0) for every OpenCV primitive (tests/test_oracle_cv2.py,
 *       tests/golden/), and
 *   (b) libstdc++ std::sort/std::list for the oct-tree control flow,
 * and is "parity unpinned" with respect to reference-owned golden vectors (none exist).
 */
#ifndef VSG_ORACLE_H
#define VSG_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* same memory layout as cv::KeyPoint (7 x 4 bytes) */
typedef struct orc_keypoint {
    float x, y;
    float size;
    float angle;
    float response;
    int32_t octave;
    int32_t class_id;
} orc_keypoint;

typedef struct orc_extractor orc_extractor;

/* ---- extractor (ORBextractor.cc) ---- */
orc_extractor *orc_extractor_create(int nfeatures, float scale_factor, int nlevels, int ini_th_fast, int min_th_fast);
void orc_extractor_destroy(orc_extractor *ex);

/* tables built by the constructor (ORBextractor.cc:411-470) */
int orc_levels(const orc_extractor *ex);
void orc_scale_factors(const orc_extractor *ex, float *scale, float *inv_scale, float *sigma2, float *inv_sigma2);
void orc_quotas(const orc_extractor *ex, int32_t *quota);
void orc_umax(const orc_extractor *ex, int32_t *umax16);

/* operator() (ORBextractor.cc:1083-1169). Returns monoIndex, or -1 for an empty image. All
 * intermediates stay inside the object for the getters below. */
int orc_extract(orc_extractor *ex, const uint8_t *img, int width, int height, int pitch, int lap_x0, int lap_x1);

int orc_num_keypoints(const orc_extractor *ex);
void orc_get_keypoints(const orc_extractor *ex, orc_keypoint *kps, uint8_t *desc /* n x 32 */);

void orc_level_size(const orc_extractor *ex, int level, int32_t *w, int32_t *h);
void orc_get_level(const orc_extractor *ex, int level, uint8_t *dst /* w*h tight */);
void orc_get_level_padded(const orc_extractor *ex, int level, uint8_t *dst /* (w+38)*(h+38) tight */);
/* blurred clone of the level; returns 0 if the level had no keypoints (blur skipped, :1125-1126) */
int orc_get_blurred(const orc_extractor *ex, int level, uint8_t *dst /* w*h tight */);
int orc_num_candidates(const orc_extractor *ex, int level);
void orc_get_candidates(const orc_extractor *ex, int level, float *xyr /* n x 3: x, y (border-relative), response */);
int orc_num_level_keypoints(const orc_extractor *ex, int level);
/* per level, after oct-tree + border offset + orientation, level coordinates, list order */
void orc_get_level_keypoints(const orc_extractor *ex, int level, orc_keypoint *kps);

/* ---- OpenCV primitives restated; exposed so tests can compare them with cv2 ---- */
void orc_resize_linear(const uint8_t *src, int sw, int sh, int spitch, uint8_t *dst, int dw, int dh, int dpitch);
void orc_gaussian_blur7(const uint8_t *src, int w, int h, int spitch, uint8_t *dst, int dpitch);
void orc_border_reflect101(const uint8_t *src, int w, int h, int spitch, uint8_t *dst, int border, int dpitch);
/* FAST-9/16 with non-max suppression on a standalone image; returns count, fills up to cap (x,y,score) */
int orc_fast(const uint8_t *img, int w, int h, int pitch, int threshold, int32_t *xys, int cap);
float orc_fast_atan2(float y, float x);
int orc_cv_round_f(float v);
int orc_cv_round_d(double v);
float orc_ic_angle(const uint8_t *img, int pitch, int x, int y);
void orc_orb_descriptor(const uint8_t *img, int pitch, int x, int y, float angle_deg, uint8_t *desc32);
/* oct-tree alone: candidates (x,y,response) in reference order -> selected indices into the input */
int orc_distribute_octree(const float *xyr, int n, int min_x, int max_x, int min_y, int max_y, int quota,
                          int32_t *selected_idx, int cap);

/* cv::cvtColor(..., COLOR_{RGB,BGR,RGBA,BGRA}2GRAY) on 8-bit images (Tracking.cc:1595-1608) */
void orc_cvt_gray(const uint8_t *src, int w, int h, int pitch, int channels, int r_first, uint8_t *dst, int dst_pitch);

/* cv::remap(src, dst, map_x, map_y, INTER_LINEAR) with CV_32FC1 maps, 8UC1 images, BORDER_CONSTANT 0 (System.cc:284-292) */
void orc_remap_bilinear(const uint8_t *src, int sw, int sh, int spitch, const float *map_x, const float *map_y, int w, int h,
                        uint8_t *dst, int dst_pitch);
/* Frame::UndistortKeyPoints / ComputeImageBounds = cv::undistortPoints(pts, K, dist, R = I, P = K) (Frame.cc:891-955) */
void orc_undistort_points(int n, const float *xy_in, double fx, double fy, double cx, double cy, const double *dist,
                          int dist_n, float *xy_out);
/* Frame::ComputeStereoFromRGBD (Frame.cc:1129-1150) */
void orc_stereo_from_rgbd(int n, const float *xy, const float *xy_un, const float *depth, int depth_pitch_floats, float bf,
                          float *u_right, float *depth_out);

/* ---- matcher arithmetic (ORBmatcher.cc) on flattened views ---- */
int orc_descriptor_distance(const uint8_t *a, const uint8_t *b);
/* CPU baseline of the brute-force kNN-2 (knnMatch k = 2): variant 0 = the reference's bit-hack distance, 1 = popcount
 * instructions; all `threads` host threads; returns seconds, fills idx / dist (nq x 2) when non-null */
double orc_bench_knn2(const uint8_t *query, int nq, const uint8_t *train, int nt, int threads, int variant, int32_t *idx_out,
                      int32_t *dist_out);

/* What the Search* methods read from a Frame / KeyFrame (Frame.h:254-290,363-381). Layout-identical to
 * vsg_frame_view in include/vsg_cuda.h so tests can hand the same buffers to both. */
typedef struct orc_frame_view {
    int32_t n;                     /* N */
    const orc_keypoint *keys;      /* mvKeysUn */
    const uint8_t *descriptors;    /* mDescriptors, n x 32 */
    const float *u_right;          /* mvuRight, NULL for monocular */
    float min_x, min_y, max_x, max_y;      /* mnMinX .. mnMaxY */
    float grid_inv_w, grid_inv_h;  /* mfGridElementWidthInv / HeightInv */
    int32_t grid_cols, grid_rows;  /* FRAME_GRID_COLS / ROWS */
    const float *scale_factors;    /* mvScaleFactors */
    int32_t n_levels;
} orc_frame_view;

/* MapPoint fields read by SearchByProjection(Frame&, vector<MapPoint*>&) (MapPoint.h:142-177). */
typedef struct orc_track_point {
    float proj_x, proj_y, proj_xr; /* mTrackProjX, mTrackProjY, mTrackProjXR */
    float view_cos;                /* mTrackViewCos */
    float depth;                   /* mTrackDepth */
    int32_t level;                 /* mnTrackScaleLevel */
    uint8_t in_view;               /* mbTrackInView */
    uint8_t bad;                   /* isBad() */
    uint8_t blocks;                /* Observations() > 0 */
    uint8_t pad;
} orc_track_point;

/* A last-frame map point already projected into the current frame (ORBmatcher.cc:1690-1716). */
typedef struct orc_proj_point {
    float u, v;                    /* uv = pCamera->project(Tcw * x3Dw) */
    float ur;                      /* uv(0) - mbf * invzc */
    float angle;                   /* LastFrame.mvKeysUn[i].angle */
    int32_t octave;                /* LastFrame.mvKeys[i].octave */
    uint8_t valid;                 /* pMP && !outlier && invzc >= 0 && uv inside the image bounds */
    uint8_t blocks;                /* pMP->Observations() > 0 */
    uint8_t pad[2];
} orc_proj_point;

/* Frame::GetFeaturesInArea (Frame.cc:802-868); returns count, writes up to cap indices. */
int orc_get_features_in_area(const orc_frame_view *f, float x, float y, float r, int min_level, int max_level,
                             int32_t *out, int cap);
/* ORBmatcher::ComputeThreeMaxima (ORBmatcher.cc:2002-2043) on the bin sizes. */
void orc_three_maxima(const int32_t *sizes, int L, int32_t *ind1, int32_t *ind2, int32_t *ind3);

/* SearchByProjection(Frame&, vector<MapPoint*>&, th, bFarPoints, thFarPoints) (ORBmatcher.cc:42-144, Nleft==-1).
 * occupied[i] = F.mvpMapPoints[i] && Observations()>0.  assign[i] = index of the map point written to
 * F.mvpMapPoints[i], or -1 if untouched.  Returns nmatches. */
int orc_search_by_projection_map(const orc_frame_view *F, const uint8_t *occupied, int n_mp, const orc_track_point *pts,
                                 const uint8_t *mp_desc, float th, int far_points, float th_far, float nnratio,
                                 int32_t *assign);
/* the same on a two-camera frame (F.Nleft != -1), ORBmatcher.cc:42-216 incl. the right-camera branch :146-213 */
int orc_search_by_projection_map_2cam(const orc_frame_view *FL, const orc_frame_view *FR, const uint8_t *occupied,
                                      const int32_t *left_to_right, const int32_t *right_to_left, int n_mp,
                                      const orc_track_point *pl, const orc_track_point *pr, const uint8_t *mp_desc, float th,
                                      int far_points, float th_far, float nnratio, int32_t *assign);
/* SearchByProjection(Frame& Cur, const Frame& Last, th, bMono) (ORBmatcher.cc:1667-1878, Nleft==-1).
 * mode: 0 = octave-1..octave+1, 1 = forward (>= octave), 2 = backward (0..octave).  assign[i] = index of the
 * last-frame point written to Cur.mvpMapPoints[i], -1 untouched, -2 written and then cleared by the rotation check. */
int orc_search_by_projection_last(const orc_frame_view *Cur, const uint8_t *occupied, int n_last,
                                  const orc_proj_point *pts, const uint8_t *desc, float th, int mode, int check_ori,
                                  int32_t *assign);
/* the same with a two-camera current frame (CurrentFrame.Nleft != -1), ORBmatcher.cc:1785-1852 */
int orc_search_by_projection_last_2cam(const orc_frame_view *CurL, const orc_frame_view *CurR, const uint8_t *occupied,
                                       int n_last, const orc_proj_point *pl, const orc_proj_point *pr, const uint8_t *desc,
                                       float th, int mode, int check_ori, int32_t *assign);
/* SearchForInitialization (ORBmatcher.cc:643-756). prev_matched: n1 x 2 floats, updated in place. */
int orc_search_for_initialization(const orc_frame_view *F1, const orc_frame_view *F2, float *prev_matched,
                                  int window_size, float nnratio, int check_ori, int32_t *matches12);
/* SearchByBoW(KeyFrame*, Frame&, ...) (ORBmatcher.cc:226-428, Nleft==-1). Feature vectors are given as sorted
 * node ids + CSR lists (DBoW2::FeatureVector is a std::map<NodeId, vector<unsigned>>).  kf_mp_valid[i] = map point
 * present and not bad.  matches_f[j] = KF feature index matched to frame feature j, or -1. */
int orc_search_by_bow(const orc_frame_view *KF, const uint8_t *kf_mp_valid, const orc_frame_view *F, int kf_nnodes,
                      const int32_t *kf_nodes, const int32_t *kf_ptr, const int32_t *kf_idx, int f_nnodes,
                      const int32_t *f_nodes, const int32_t *f_ptr, const int32_t *f_idx, float nnratio, int check_ori,
                      int32_t *matches_f);
/* the same with a two-camera frame: F's features [0, f_nleft) belong to the left camera (f_nleft == -1: single camera) */
int orc_search_by_bow_2cam(const orc_frame_view *KF, const uint8_t *kf_mp_valid, const orc_frame_view *F, int f_nleft,
                           int kf_nnodes, const int32_t *kf_nodes, const int32_t *kf_ptr, const int32_t *kf_idx, int f_nnodes,
                           const int32_t *f_nodes, const int32_t *f_ptr, const int32_t *f_idx, float nnratio, int check_ori,
                           int32_t *matches_f);

/* A map point already projected into the searched Frame / KeyFrame (layout-identical to vsg_search_point). */
typedef struct orc_search_point {
    float u, v, ur, angle;
    int32_t level;                 /* nPredictedLevel */
    uint8_t valid;
    uint8_t pad[3];
} orc_search_point;

/* SearchByProjection(Frame&, KeyFrame*, sAlreadyFound, th, ORBdist) (ORBmatcher.cc:1880-2000). */
int orc_search_by_projection_reloc(const orc_frame_view *Cur, const uint8_t *occupied, int n, const orc_search_point *pts,
                                   const uint8_t *desc, float th, int orb_dist, int check_ori, int32_t *assign);
/* SearchByProjection(KeyFrame*, Sim3f&, vpPoints[, vpPointsKFs], vpMatched, th, ratioHamming) (:430-641). */
int orc_search_by_projection_sim3(const orc_frame_view *KF, const uint8_t *matched, int n, const orc_search_point *pts,
                                  const uint8_t *desc, int th, float ratio_hamming, int32_t *assign);
/* The search of Fuse (variant 0: :1148-1335 with the chi2 gates; variant 1: :1337-1446). */
int orc_fuse_search(const orc_frame_view *KF, int n, const orc_search_point *pts, const uint8_t *desc, float th,
                    const float *inv_level_sigma2, int variant, int32_t *best_idx_out);
/* SearchBySim3 (:1448-1665): pts1 has KF1->n entries, pts2 has KF2->n. */
int orc_search_by_sim3(const orc_frame_view *KF1, const orc_frame_view *KF2, const orc_search_point *pts1,
                       const uint8_t *desc1, const orc_search_point *pts2, const uint8_t *desc2, float th,
                       int32_t *matches12);
/* ... with n1 / n2 map point slots >= the feature counts of the views (two-camera keyframes: the views are the left cameras). */
int orc_search_by_sim3_n(const orc_frame_view *KF1, const orc_frame_view *KF2, int n1, const orc_search_point *pts1,
                         const uint8_t *desc1, int n2, const orc_search_point *pts2, const uint8_t *desc2, float th,
                         int32_t *matches12);
/* SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12) (:758-900). */
int orc_search_by_bow_kf(const orc_frame_view *KF1, const uint8_t *mp_valid1, const orc_frame_view *KF2,
                         const uint8_t *mp_valid2, int nn1, const int32_t *nodes1, const int32_t *ptr1,
                         const int32_t *idx1, int nn2, const int32_t *nodes2, const int32_t *ptr2, const int32_t *idx2,
                         float nnratio, int check_ori, int32_t *matches12);
/* SearchForTriangulation (:902-1146, mpCamera2 == NULL, Pinhole::epipolarConstrain with F12 given row-major). */
int orc_search_for_triangulation(const orc_frame_view *KF1, const uint8_t *has_mp1, const orc_frame_view *KF2,
                                 const uint8_t *has_mp2, int nn1, const int32_t *nodes1, const int32_t *ptr1,
                                 const int32_t *idx1, int nn2, const int32_t *nodes2, const int32_t *ptr2,
                                 const int32_t *idx2, int only_stereo, int coarse, const float *f12, const float *ep,
                                 const float *level_sigma2_2, int check_ori, int32_t *matches12);

/* DBoW2 TemplatedVocabulary::transform tree walk (TemplatedVocabulary.h:1225-1265) on a flattened tree. */
void orc_bow_transform(const int32_t *child_ptr, const int32_t *child_idx, const uint8_t *node_desc, int levels,
                       const uint8_t *desc, int n, int levelsup, int32_t *leaf, int32_t *nid);

/* MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:340-417) for one point with n observed descriptors. */
int orc_distinctive_descriptor(const uint8_t *desc, int n);

/* Frame::ComputeStereoMatches (Frame.cc:957-1127): row-band Hamming search, 11x11 SAD refinement on the
 * un-blurred pyramids of both extractors (their last orc_extract call), parabola fit, median-based outlier
 * rejection.  keys/desc are the extractors' raw outputs (mvKeys / mvKeysRight).  Fills u_right[n_l], depth[n_l]
 * (-1 = no match). */
void orc_stereo_matches(const orc_extractor *left, const orc_extractor *right, const orc_keypoint *keys_l,
                        const uint8_t *desc_l, int n_l, const orc_keypoint *keys_r, const uint8_t *desc_r, int n_r,
                        float mb, float mbf, float *u_right, float *depth);

/* ---- throughput harness for bench.py's cpu_baseline: extracts `nframes` frames (tightly packed
 * w*h each) with `threads` worker threads, one extractor instance per thread; returns seconds. ---- */
double orc_bench_extract(const uint8_t *frames, int nframes, int width, int height, int nfeatures, float scale_factor,
                         int nlevels, int ini_th, int min_th, int threads, int64_t *total_keypoints);

#ifdef __cplusplus
}
#endif
#endif
