// vsg_on_oracle.cpp — TEST ADAPTER: the matcher entry points of include/vsg_cuda.h implemented on the CPU oracle
// (oracle/oracle.h).  It exists so that the drop-in shim's host code (visual_sgraphs_b200/shim/ORBmatcher.h: flattening,
// pose arithmetic, write-back, Replace / AddObservation replay) can be run WITHOUT a GPU against the reference's own
// ORBmatcher.cc (tests/cpp/ref_matcher_test.cpp, CPU build).  Never linked into the product: libvsg_cuda.so has no CPU path.
#include <cstring>
#include <string>
#include <vector>

#include "../../include/vsg_cuda.h"
#include "../../oracle/oracle.h"

static_assert(sizeof(vsg_frame_view) == sizeof(orc_frame_view), "view layout");
static_assert(sizeof(vsg_track_point) == sizeof(orc_track_point), "track point layout");
static_assert(sizeof(vsg_proj_point) == sizeof(orc_proj_point), "proj point layout");
static_assert(sizeof(vsg_search_point) == sizeof(orc_search_point), "search point layout");
static_assert(sizeof(vsg_keypoint) == sizeof(orc_keypoint), "keypoint layout");

struct vsg_matcher { int dummy; };
struct vsg_frame {   // deep copy of the view
    std::vector<orc_keypoint> keys;
    std::vector<uint8_t> desc;
    std::vector<float> u_right, scale;
    orc_frame_view v;
};

static const orc_frame_view *V(const vsg_frame_view *v) { return reinterpret_cast<const orc_frame_view *>(v); }
static const orc_frame_view *V(const vsg_frame *f) { return &f->v; }

extern "C" {

const char *vsg_last_error(void) { return "vsg_on_oracle adapter"; }
vsg_status vsg_matcher_create(int, vsg_matcher **out) { *out = new vsg_matcher{0}; return VSG_OK; }
void vsg_matcher_destroy(vsg_matcher *m) { delete m; }

vsg_status vsg_frame_create(vsg_matcher *, const vsg_frame_view *view, vsg_frame **out) {
    vsg_frame *f = new vsg_frame;
    const orc_frame_view *s = V(view);
    f->keys.assign(s->keys, s->keys + s->n);
    f->desc.assign(s->descriptors, s->descriptors + (size_t)s->n * 32);
    if (s->u_right) f->u_right.assign(s->u_right, s->u_right + s->n);
    f->scale.assign(s->scale_factors, s->scale_factors + s->n_levels);
    f->v = *s;
    f->v.keys = f->keys.data();
    f->v.descriptors = f->desc.data();
    f->v.u_right = s->u_right ? f->u_right.data() : nullptr;
    f->v.scale_factors = f->scale.data();
    *out = f;
    return VSG_OK;
}
void vsg_frame_destroy(vsg_frame *f) { delete f; }

vsg_status vsg_search_by_projection_map(vsg_matcher *, const vsg_frame *F, const uint8_t *occupied, int n_mp, const vsg_track_point *pts,
                                        const uint8_t *mp_desc, float th, int far_points, float th_far, float nnratio,
                                        int32_t *assign_out, int *nmatches_out) {
    *nmatches_out = orc_search_by_projection_map(V(F), occupied, n_mp, reinterpret_cast<const orc_track_point *>(pts), mp_desc, th,
                                                 far_points, th_far, nnratio, assign_out);
    return VSG_OK;
}
vsg_status vsg_search_by_projection_map_2cam(vsg_matcher *, const vsg_frame *FL, const vsg_frame *FR, const uint8_t *occupied,
                                             const int32_t *l2r, const int32_t *r2l, int n_mp, const vsg_track_point *pl,
                                             const vsg_track_point *pr, const uint8_t *mp_desc, float th, int far_points, float th_far,
                                             float nnratio, int32_t *assign_out, int *nmatches_out) {
    *nmatches_out = orc_search_by_projection_map_2cam(V(FL), V(FR), occupied, l2r, r2l, n_mp, reinterpret_cast<const orc_track_point *>(pl),
                                                      reinterpret_cast<const orc_track_point *>(pr), mp_desc, th, far_points, th_far,
                                                      nnratio, assign_out);
    return VSG_OK;
}
vsg_status vsg_search_by_projection_last(vsg_matcher *, const vsg_frame *Cur, const uint8_t *occupied, int n_last, const vsg_proj_point *pts,
                                         const uint8_t *desc, float th, int mode, int check_ori, int32_t *assign_out, int *nmatches_out) {
    *nmatches_out = orc_search_by_projection_last(V(Cur), occupied, n_last, reinterpret_cast<const orc_proj_point *>(pts), desc, th, mode,
                                                  check_ori, assign_out);
    return VSG_OK;
}
vsg_status vsg_search_by_projection_last_2cam(vsg_matcher *, const vsg_frame *CurL, const vsg_frame *CurR, const uint8_t *occupied,
                                              int n_last, const vsg_proj_point *pl, const vsg_proj_point *pr, const uint8_t *desc, float th,
                                              int mode, int check_ori, int32_t *assign_out, int *nmatches_out) {
    *nmatches_out = orc_search_by_projection_last_2cam(V(CurL), V(CurR), occupied, n_last, reinterpret_cast<const orc_proj_point *>(pl),
                                                       reinterpret_cast<const orc_proj_point *>(pr), desc, th, mode, check_ori, assign_out);
    return VSG_OK;
}
vsg_status vsg_search_for_initialization(vsg_matcher *, const vsg_frame_view *F1, const vsg_frame *F2, float *prev_matched, int window_size,
                                         float nnratio, int check_ori, int32_t *matches12_out, int *nmatches_out) {
    *nmatches_out = orc_search_for_initialization(V(F1), V(F2), prev_matched, window_size, nnratio, check_ori, matches12_out);
    return VSG_OK;
}
vsg_status vsg_search_by_bow_2cam(vsg_matcher *, const vsg_frame_view *KF, const uint8_t *kf_mp_valid, const vsg_frame_view *F, int f_nleft,
                                  int kf_nnodes, const int32_t *kf_nodes, const int32_t *kf_ptr, const int32_t *kf_idx, int f_nnodes,
                                  const int32_t *f_nodes, const int32_t *f_ptr, const int32_t *f_idx, float nnratio, int check_ori,
                                  int32_t *matches_f_out, int *nmatches_out) {
    *nmatches_out = orc_search_by_bow_2cam(V(KF), kf_mp_valid, V(F), f_nleft, kf_nnodes, kf_nodes, kf_ptr, kf_idx, f_nnodes, f_nodes, f_ptr,
                                           f_idx, nnratio, check_ori, matches_f_out);
    return VSG_OK;
}
vsg_status vsg_search_by_projection_reloc(vsg_matcher *, const vsg_frame *Cur, const uint8_t *occupied, int n, const vsg_search_point *pts,
                                          const uint8_t *desc, float th, int orb_dist, int check_ori, int32_t *assign_out,
                                          int *nmatches_out) {
    *nmatches_out = orc_search_by_projection_reloc(V(Cur), occupied, n, reinterpret_cast<const orc_search_point *>(pts), desc, th, orb_dist,
                                                   check_ori, assign_out);
    return VSG_OK;
}
vsg_status vsg_search_by_projection_sim3(vsg_matcher *, const vsg_frame *KF, const uint8_t *matched, int n, const vsg_search_point *pts,
                                         const uint8_t *desc, int th, float ratio_hamming, int32_t *assign_out, int *nmatches_out) {
    *nmatches_out = orc_search_by_projection_sim3(V(KF), matched, n, reinterpret_cast<const orc_search_point *>(pts), desc, th, ratio_hamming,
                                                  assign_out);
    return VSG_OK;
}
vsg_status vsg_fuse_search(vsg_matcher *, const vsg_frame *KF, int n, const vsg_search_point *pts, const uint8_t *desc, float th,
                           const float *inv_level_sigma2, int variant, int32_t *best_idx_out, int *nfused_out) {
    const int nf = orc_fuse_search(V(KF), n, reinterpret_cast<const orc_search_point *>(pts), desc, th, inv_level_sigma2, variant, best_idx_out);
    if (nfused_out) *nfused_out = nf;
    return VSG_OK;
}
vsg_status vsg_search_by_sim3(vsg_matcher *, const vsg_frame *KF1, const vsg_frame *KF2, int, const vsg_search_point *pts1, const uint8_t *desc1,
                              int, const vsg_search_point *pts2, const uint8_t *desc2, float th, int32_t *matches12_out, int *nfound_out) {
    *nfound_out = orc_search_by_sim3(V(KF1), V(KF2), reinterpret_cast<const orc_search_point *>(pts1), desc1,
                                     reinterpret_cast<const orc_search_point *>(pts2), desc2, th, matches12_out);
    return VSG_OK;
}
vsg_status vsg_search_by_bow_kf(vsg_matcher *, const vsg_frame_view *KF1, const uint8_t *mp_valid1, const vsg_frame_view *KF2,
                                const uint8_t *mp_valid2, int nn1, const int32_t *nodes1, const int32_t *ptr1, const int32_t *idx1, int nn2,
                                const int32_t *nodes2, const int32_t *ptr2, const int32_t *idx2, float nnratio, int check_ori,
                                int32_t *matches12_out, int *nmatches_out) {
    *nmatches_out = orc_search_by_bow_kf(V(KF1), mp_valid1, V(KF2), mp_valid2, nn1, nodes1, ptr1, idx1, nn2, nodes2, ptr2, idx2, nnratio,
                                         check_ori, matches12_out);
    return VSG_OK;
}
vsg_status vsg_search_for_triangulation(vsg_matcher *, const vsg_frame_view *KF1, const uint8_t *has_mp1, const vsg_frame_view *KF2,
                                        const uint8_t *has_mp2, int nn1, const int32_t *nodes1, const int32_t *ptr1, const int32_t *idx1,
                                        int nn2, const int32_t *nodes2, const int32_t *ptr2, const int32_t *idx2, int only_stereo, int coarse,
                                        const float *f12, const float *ep, const float *level_sigma2_2, int check_ori,
                                        int32_t *matches12_out, int *nmatches_out) {
    *nmatches_out = orc_search_for_triangulation(V(KF1), has_mp1, V(KF2), has_mp2, nn1, nodes1, ptr1, idx1, nn2, nodes2, ptr2, idx2,
                                                 only_stereo, coarse, f12, ep, level_sigma2_2, check_ori, matches12_out);
    return VSG_OK;
}

}  // extern "C"
