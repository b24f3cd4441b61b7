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

/* ---- matcher arithmetic (ORBmatcher.cc) ---- */
int orc_descriptor_distance(const uint8_t *a, const uint8_t *b);

/* ---- throughput harness for bench.py's cpu_baseline: extracts `nframes` frames (tightly packed
 * w*h each) with `threads` worker threads, one extractor instance per thread; returns seconds. ---- */
double orc_bench_extract(const uint8_t *frames, int nframes, int width, int height, int nfeatures, float scale_factor,
                         int nlevels, int ini_th, int min_th, int threads, int64_t *total_keypoints);

#ifdef __cplusplus
}
#endif
#endif
