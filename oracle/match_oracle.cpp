/*
 * oracle/match_oracle.cpp — CPU ORACLE for the ORB matcher (test infrastructure, NOT product code).
 * Restates orb_slam3/src/ORBmatcher.cc and Frame::GetFeaturesInArea / AssignFeaturesToGrid / PosInGrid
 * (orb_slam3/src/Frame.cc:521-553, 802-880) of snt-arg/visual_sgraphs on flattened views.  Only the
 * single-camera branches (Frame::Nleft == -1) are restated; the two-camera fisheye branches are out of
 * scope for this round.  See oracle.h for the parity status.
 */
#include "oracle.h"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstring>
#include <chrono>
#include <thread>
#include <vector>

namespace {

const int TH_HIGH = 100;      // ORBmatcher.cc:34
const int TH_LOW = 50;        // ORBmatcher.cc:35
const int HISTO_LENGTH = 30;  // ORBmatcher.cc:36

int descriptor_distance(const uint8_t *a, const uint8_t *b) {   // ORBmatcher.cc:2047-2063
    int dist = 0;
    for (int i = 0; i < 8; ++i) {
        uint32_t wa, wb;
        std::memcpy(&wa, a + 4 * i, 4);
        std::memcpy(&wb, b + 4 * i, 4);
        uint32_t v = wa ^ wb;
        v = v - ((v >> 1) & 0x55555555u);
        v = (v & 0x33333333u) + ((v >> 2) & 0x33333333u);
        dist += (int)((((v + (v >> 4)) & 0x0F0F0F0Fu) * 0x01010101u) >> 24);
    }
    return dist;
}

// Frame::AssignFeaturesToGrid + PosInGrid (Frame.cc:521-553, 870-880)
struct Grid {
    int cols, rows;
    std::vector<std::vector<int>> cell;   // [ix * rows + iy]
    explicit Grid(const orc_frame_view *f) : cols(f->grid_cols), rows(f->grid_rows), cell((size_t)cols * rows) {
        for (int i = 0; i < f->n; ++i) {
            const int px = (int)std::round((f->keys[i].x - f->min_x) * f->grid_inv_w);
            const int py = (int)std::round((f->keys[i].y - f->min_y) * f->grid_inv_h);
            if (px < 0 || px >= cols || py < 0 || py >= rows) continue;
            cell[(size_t)px * rows + py].push_back(i);
        }
    }
};

// Frame::GetFeaturesInArea (Frame.cc:802-868)
void features_in_area(const orc_frame_view *f, const Grid &g, float x, float y, float r, int min_level, int max_level,
                      std::vector<int> &out) {
    out.clear();
    const float fx = r, fy = r;
    const int min_cx = std::max(0, (int)std::floor((x - f->min_x - fx) * f->grid_inv_w));
    if (min_cx >= g.cols) return;
    const int max_cx = std::min(g.cols - 1, (int)std::ceil((x - f->min_x + fx) * f->grid_inv_w));
    if (max_cx < 0) return;
    const int min_cy = std::max(0, (int)std::floor((y - f->min_y - fy) * f->grid_inv_h));
    if (min_cy >= g.rows) return;
    const int max_cy = std::min(g.rows - 1, (int)std::ceil((y - f->min_y + fy) * f->grid_inv_h));
    if (max_cy < 0) return;
    const bool check_levels = (min_level > 0) || (max_level >= 0);
    for (int ix = min_cx; ix <= max_cx; ++ix)
        for (int iy = min_cy; iy <= max_cy; ++iy)
            for (int idx : g.cell[(size_t)ix * g.rows + iy]) {
                const orc_keypoint &kp = f->keys[idx];
                if (check_levels) {
                    if (kp.octave < min_level) continue;
                    if (max_level >= 0 && kp.octave > max_level) continue;
                }
                const float dx = kp.x - x, dy = kp.y - y;
                if (std::fabs(dx) < fx && std::fabs(dy) < fy) out.push_back(idx);
            }
}

// ORBmatcher::ComputeThreeMaxima (ORBmatcher.cc:2002-2043)
void three_maxima(const int *sizes, int L, int &ind1, int &ind2, int &ind3) {
    int max1 = 0, max2 = 0, max3 = 0;
    for (int i = 0; i < L; ++i) {
        const int s = sizes[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}

// the rotation-histogram idiom (e.g. ORBmatcher.cc:351-358): factor = 1/HISTO_LENGTH, C round()
int rot_bin(float a1, float a2) {
    const float factor = 1.0f / HISTO_LENGTH;
    float rot = a1 - a2;
    if (rot < 0.0) rot += 360.0f;
    int bin = (int)std::round(rot * factor);
    if (bin == HISTO_LENGTH) bin = 0;
    return bin;
}

}  // namespace

extern "C" {

int orc_descriptor_distance(const uint8_t *a, const uint8_t *b) { return descriptor_distance(a, b); }

// Brute-force kNN-2 on the host cores (cv::BFMatcher(NORM_HAMMING).knnMatch(k = 2), Frame.cc:43,1200): the CPU baseline of
// the matching benchmark.  variant 0: the reference's bit-hack DescriptorDistance (ORBmatcher.cc:2047-2063);
// variant 1: the same loop with __builtin_popcountll, labelled as such.  Ties -> lower train index first.  Returns seconds.
double orc_bench_knn2(const uint8_t *query, int nq, const uint8_t *train, int nt, int threads, int variant, int32_t *idx_out,
                      int32_t *dist_out) {
    if (threads < 1) threads = 1;
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int tid = 0; tid < threads; ++tid) {
        pool.emplace_back([=]() {
            for (int q = tid; q < nq; q += threads) {
                const uint8_t *a = query + (size_t)q * 32;
                int b0 = 257, b1 = 257, i0 = -1, i1 = -1;
                for (int t = 0; t < nt; ++t) {
                    const uint8_t *b = train + (size_t)t * 32;
                    int d;
                    if (variant == 0) {
                        d = descriptor_distance(a, b);
                    } else {
                        uint64_t wa[4], wb[4];
                        std::memcpy(wa, a, 32); std::memcpy(wb, b, 32);
                        d = __builtin_popcountll(wa[0] ^ wb[0]) + __builtin_popcountll(wa[1] ^ wb[1]) +
                            __builtin_popcountll(wa[2] ^ wb[2]) + __builtin_popcountll(wa[3] ^ wb[3]);
                    }
                    if (d < b0) { b1 = b0; i1 = i0; b0 = d; i0 = t; }
                    else if (d < b1) { b1 = d; i1 = t; }
                }
                if (idx_out) { idx_out[2 * q] = i0; idx_out[2 * q + 1] = i1; }
                if (dist_out) { dist_out[2 * q] = i0 >= 0 ? b0 : -1; dist_out[2 * q + 1] = i1 >= 0 ? b1 : -1; }
            }
        });
    }
    for (auto &th : pool) th.join();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

int orc_get_features_in_area(const orc_frame_view *f, float x, float y, float r, int min_level, int max_level,
                             int32_t *out, int cap) {
    Grid g(f);
    std::vector<int> v;
    features_in_area(f, g, x, y, r, min_level, max_level, v);
    for (size_t i = 0; i < v.size() && (int)i < cap; ++i) out[i] = v[i];
    return (int)v.size();
}

void orc_three_maxima(const int32_t *sizes, int L, int32_t *ind1, int32_t *ind2, int32_t *ind3) {
    int a = -1, b = -1, c = -1;
    three_maxima(sizes, L, a, b, c);
    *ind1 = a; *ind2 = b; *ind3 = c;
}

// ORBmatcher.cc:42-144 (left / single-camera branch)
int orc_search_by_projection_map(const orc_frame_view *F, const uint8_t *occupied_in, int n_mp,
                                 const orc_track_point *pts, const uint8_t *mp_desc, float th, int far_points,
                                 float th_far, float nnratio, int32_t *assign) {
    Grid g(F);
    std::vector<uint8_t> blocked(occupied_in, occupied_in + F->n);   // F.mvpMapPoints[idx] with Observations() > 0
    for (int i = 0; i < F->n; ++i) assign[i] = -1;
    int nmatches = 0;
    const bool b_factor = th != 1.0;
    std::vector<int> cand;
    for (int i = 0; i < n_mp; ++i) {
        const orc_track_point &mp = pts[i];
        if (!mp.in_view) continue;
        if (far_points && mp.depth > th_far) continue;
        if (mp.bad) continue;
        const int level = mp.level;
        float r = (mp.view_cos > 0.998) ? 2.5f : 4.0f;   // RadiusByViewingCos :218-224
        if (b_factor) r *= th;
        features_in_area(F, g, mp.proj_x, mp.proj_y, r * F->scale_factors[level], level - 1, level, cand);
        if (cand.empty()) continue;
        const uint8_t *d_mp = mp_desc + (size_t)i * 32;
        int best = 256, best_level = -1, best2 = 256, best_level2 = -1, best_idx = -1;
        for (int idx : cand) {
            if (blocked[idx]) continue;
            if (F->u_right && F->u_right[idx] > 0) {
                const float er = std::fabs(mp.proj_xr - F->u_right[idx]);
                if (er > r * F->scale_factors[level]) continue;
            }
            const int dist = descriptor_distance(d_mp, F->descriptors + (size_t)idx * 32);
            if (dist < best) {
                best2 = best; best = dist; best_level2 = best_level; best_level = F->keys[idx].octave; best_idx = idx;
            } else if (dist < best2) {
                best_level2 = F->keys[idx].octave; best2 = dist;
            }
        }
        if (best <= TH_HIGH) {
            if (best_level == best_level2 && best > nnratio * best2) continue;
            if (best_level != best_level2 || best <= nnratio * best2) {
                assign[best_idx] = i;
                blocked[best_idx] = mp.blocks;   // later points skip it only if this one has observations
                ++nmatches;
            }
        }
    }
    return nmatches;
}

// ORBmatcher.cc:42-216 on a two-camera frame (F.Nleft != -1): FL / FR are the two cameras' keypoints with their own
// grids (Frame::GetFeaturesInArea(..., bRight) reads mGrid + mvKeys or mGridRight + mvKeysRight, Frame.cc:840-848);
// slots [0, nL) of occupied / assign are the left keypoints, [nL, nL + nR) the right ones.  pl[i] / pr[i] are the map
// point's left / right tracking members; depth, bad and blocks are read from pl[i].
int orc_search_by_projection_map_2cam(const orc_frame_view *FL, const orc_frame_view *FR, const uint8_t *occupied_in,
                                      const int32_t *left_to_right, const int32_t *right_to_left, int n_mp,
                                      const orc_track_point *pl, const orc_track_point *pr, const uint8_t *mp_desc, float th,
                                      int far_points, float th_far, float nnratio, int32_t *assign) {
    Grid gl(FL), gr(FR);
    const int nL = FL->n, N = FL->n + FR->n;
    std::vector<uint8_t> blocked(occupied_in, occupied_in + N);
    for (int i = 0; i < N; ++i) assign[i] = -1;
    int nmatches = 0;
    const bool b_factor = th != 1.0;
    std::vector<int> cand;
    for (int i = 0; i < n_mp; ++i) {
        const orc_track_point &l = pl[i], &rp = pr[i];
        if (!l.in_view && !rp.in_view) continue;
        if (far_points && l.depth > th_far) continue;
        if (l.bad) continue;
        const uint8_t *d_mp = mp_desc + (size_t)i * 32;
        if (l.in_view) {
            const int level = l.level;
            float r = (l.view_cos > 0.998) ? 2.5f : 4.0f;
            if (b_factor) r *= th;
            features_in_area(FL, gl, l.proj_x, l.proj_y, r * FL->scale_factors[level], level - 1, level, cand);
            if (!cand.empty()) {
                int best = 256, best_level = -1, best2 = 256, best_level2 = -1, best_idx = -1;
                for (int idx : cand) {
                    if (blocked[idx]) continue;
                    const int dist = descriptor_distance(d_mp, FL->descriptors + (size_t)idx * 32);
                    if (dist < best) {
                        best2 = best; best = dist; best_level2 = best_level; best_level = FL->keys[idx].octave; best_idx = idx;
                    } else if (dist < best2) {
                        best_level2 = FL->keys[idx].octave; best2 = dist;
                    }
                }
                if (best <= TH_HIGH) {
                    if (best_level == best_level2 && best > nnratio * best2) continue;      // also skips the right camera
                    if (best_level != best_level2 || best <= nnratio * best2) {
                        assign[best_idx] = i; blocked[best_idx] = l.blocks;
                        if (left_to_right[best_idx] != -1) {
                            assign[left_to_right[best_idx] + nL] = i; blocked[left_to_right[best_idx] + nL] = l.blocks;
                            ++nmatches;
                        }
                        ++nmatches;
                    }
                }
            }
        }
        if (rp.in_view) {
            const int level = rp.level;
            if (level != -1) {
                const float r = (rp.view_cos > 0.998) ? 2.5f : 4.0f;
                features_in_area(FR, gr, rp.proj_x, rp.proj_y, r * FR->scale_factors[level], level - 1, level, cand);
                if (cand.empty()) continue;
                int best = 256, best_level = -1, best2 = 256, best_level2 = -1, best_idx = -1;
                for (int idx : cand) {
                    if (blocked[idx + nL]) continue;
                    const int dist = descriptor_distance(d_mp, FR->descriptors + (size_t)idx * 32);
                    if (dist < best) {
                        best2 = best; best = dist; best_level2 = best_level; best_level = FR->keys[idx].octave; best_idx = idx;
                    } else if (dist < best2) {
                        best_level2 = FR->keys[idx].octave; best2 = dist;
                    }
                }
                if (best <= TH_HIGH) {
                    if (best_level == best_level2 && best > nnratio * best2) continue;
                    if (right_to_left[best_idx] != -1) {
                        assign[right_to_left[best_idx]] = i; blocked[right_to_left[best_idx]] = l.blocks;
                        ++nmatches;
                    }
                    assign[best_idx + nL] = i; blocked[best_idx + nL] = l.blocks;
                    ++nmatches;
                }
            }
        }
    }
    return nmatches;
}

// ORBmatcher.cc:1667-1784, 1856-1878 (single-camera branch)
int orc_search_by_projection_last(const orc_frame_view *Cur, const uint8_t *occupied_in, int n_last,
                                  const orc_proj_point *pts, const uint8_t *desc, float th, int mode, int check_ori,
                                  int32_t *assign) {
    Grid g(Cur);
    std::vector<uint8_t> blocked(occupied_in, occupied_in + Cur->n);
    for (int i = 0; i < Cur->n; ++i) assign[i] = -1;
    std::vector<int> rot_hist[HISTO_LENGTH];
    int nmatches = 0;
    std::vector<int> cand;
    for (int i = 0; i < n_last; ++i) {
        const orc_proj_point &p = pts[i];
        if (!p.valid) continue;
        const int oct = p.octave;
        const float radius = th * Cur->scale_factors[oct];
        if (mode == 1) features_in_area(Cur, g, p.u, p.v, radius, oct, -1, cand);
        else if (mode == 2) features_in_area(Cur, g, p.u, p.v, radius, 0, oct, cand);
        else features_in_area(Cur, g, p.u, p.v, radius, oct - 1, oct + 1, cand);
        if (cand.empty()) continue;
        const uint8_t *d_mp = desc + (size_t)i * 32;
        int best = 256, best_idx = -1;
        for (int i2 : cand) {
            if (blocked[i2]) continue;
            if (Cur->u_right && Cur->u_right[i2] > 0) {
                const float er = std::fabs(p.ur - Cur->u_right[i2]);
                if (er > radius) continue;
            }
            const int dist = descriptor_distance(d_mp, Cur->descriptors + (size_t)i2 * 32);
            if (dist < best) { best = dist; best_idx = i2; }
        }
        if (best <= TH_HIGH) {
            assign[best_idx] = i;
            blocked[best_idx] = p.blocks;
            ++nmatches;
            if (check_ori) rot_hist[rot_bin(p.angle, Cur->keys[best_idx].angle)].push_back(best_idx);
        }
    }
    if (check_ori) {
        int sizes[HISTO_LENGTH], ind1 = -1, ind2 = -1, ind3 = -1;
        for (int i = 0; i < HISTO_LENGTH; ++i) sizes[i] = (int)rot_hist[i].size();
        three_maxima(sizes, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int idx : rot_hist[i]) { assign[idx] = -2; --nmatches; }
        }
    }
    return nmatches;
}

// ORBmatcher.cc:1667-1878 with a two-camera current frame: pl[i] = the point projected with the left camera (valid, u, v,
// angle of the last-frame keypoint, its octave, blocks), pr[i].u / .v = its projection into the right camera (:1787-1788).
int orc_search_by_projection_last_2cam(const orc_frame_view *CurL, const orc_frame_view *CurR, const uint8_t *occupied_in,
                                       int n_last, const orc_proj_point *pl, const orc_proj_point *pr, const uint8_t *desc,
                                       float th, int mode, int check_ori, int32_t *assign) {
    Grid gl(CurL), gr(CurR);
    const int nL = CurL->n, N = CurL->n + CurR->n;
    std::vector<uint8_t> blocked(occupied_in, occupied_in + N);
    for (int i = 0; i < N; ++i) assign[i] = -1;
    std::vector<int> rot_hist[HISTO_LENGTH];
    int nmatches = 0;
    std::vector<int> cand;
    auto window = [&](const orc_frame_view *F, const Grid &g, float u, float v, float radius, int oct) {
        if (mode == 1) features_in_area(F, g, u, v, radius, oct, -1, cand);
        else if (mode == 2) features_in_area(F, g, u, v, radius, 0, oct, cand);
        else features_in_area(F, g, u, v, radius, oct - 1, oct + 1, cand);
    };
    for (int i = 0; i < n_last; ++i) {
        const orc_proj_point &p = pl[i];
        if (!p.valid) continue;
        const int oct = p.octave;
        const float radius = th * CurL->scale_factors[oct];
        const uint8_t *d_mp = desc + (size_t)i * 32;
        window(CurL, gl, p.u, p.v, radius, oct);
        if (cand.empty()) continue;
        {
            int best = 256, best_idx = -1;
            for (int i2 : cand) {
                if (blocked[i2]) continue;
                const int dist = descriptor_distance(d_mp, CurL->descriptors + (size_t)i2 * 32);
                if (dist < best) { best = dist; best_idx = i2; }
            }
            if (best <= TH_HIGH) {
                assign[best_idx] = i; blocked[best_idx] = p.blocks; ++nmatches;
                if (check_ori) rot_hist[rot_bin(p.angle, CurL->keys[best_idx].angle)].push_back(best_idx);
            }
        }
        window(CurR, gr, pr[i].u, pr[i].v, radius, oct);
        int best = 256, best_idx = -1;
        for (int i2 : cand) {
            if (blocked[i2 + nL]) continue;
            const int dist = descriptor_distance(d_mp, CurR->descriptors + (size_t)i2 * 32);
            if (dist < best) { best = dist; best_idx = i2; }
        }
        if (best <= TH_HIGH) {
            assign[best_idx + nL] = i; blocked[best_idx + nL] = p.blocks; ++nmatches;
            if (check_ori) rot_hist[rot_bin(p.angle, CurR->keys[best_idx].angle)].push_back(best_idx + nL);
        }
    }
    if (check_ori) {
        int sizes[HISTO_LENGTH], ind1 = -1, ind2 = -1, ind3 = -1;
        for (int i = 0; i < HISTO_LENGTH; ++i) sizes[i] = (int)rot_hist[i].size();
        three_maxima(sizes, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int idx : rot_hist[i]) { assign[idx] = -2; --nmatches; }
        }
    }
    return nmatches;
}

// ORBmatcher.cc:643-756
int orc_search_for_initialization(const orc_frame_view *F1, const orc_frame_view *F2, float *prev_matched,
                                  int window_size, float nnratio, int check_ori, int32_t *matches12) {
    Grid g2(F2);
    int nmatches = 0;
    for (int i = 0; i < F1->n; ++i) matches12[i] = -1;
    std::vector<int> rot_hist[HISTO_LENGTH];
    std::vector<int> matched_distance(F2->n, INT_MAX), matches21(F2->n, -1);
    std::vector<int> cand;
    for (int i1 = 0; i1 < F1->n; ++i1) {
        const int level1 = F1->keys[i1].octave;
        if (level1 > 0) continue;
        features_in_area(F2, g2, prev_matched[2 * i1], prev_matched[2 * i1 + 1], (float)window_size, level1, level1, cand);
        if (cand.empty()) continue;
        const uint8_t *d1 = F1->descriptors + (size_t)i1 * 32;
        int best = INT_MAX, best2 = INT_MAX, best_idx2 = -1;
        for (int i2 : cand) {
            const int dist = descriptor_distance(d1, F2->descriptors + (size_t)i2 * 32);
            if (matched_distance[i2] <= dist) continue;
            if (dist < best) { best2 = best; best = dist; best_idx2 = i2; }
            else if (dist < best2) best2 = dist;
        }
        if (best <= TH_LOW) {
            if (best < (float)best2 * nnratio) {
                if (matches21[best_idx2] >= 0) { matches12[matches21[best_idx2]] = -1; --nmatches; }
                matches12[i1] = best_idx2;
                matches21[best_idx2] = i1;
                matched_distance[best_idx2] = best;
                ++nmatches;
                if (check_ori) rot_hist[rot_bin(F1->keys[i1].angle, F2->keys[best_idx2].angle)].push_back(i1);
            }
        }
    }
    if (check_ori) {
        int sizes[HISTO_LENGTH], ind1 = -1, ind2 = -1, ind3 = -1;
        for (int i = 0; i < HISTO_LENGTH; ++i) sizes[i] = (int)rot_hist[i].size();
        three_maxima(sizes, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int idx1 : rot_hist[i])
                if (matches12[idx1] >= 0) { matches12[idx1] = -1; --nmatches; }
        }
    }
    for (int i1 = 0; i1 < F1->n; ++i1)
        if (matches12[i1] >= 0) {
            prev_matched[2 * i1] = F2->keys[matches12[i1]].x;
            prev_matched[2 * i1 + 1] = F2->keys[matches12[i1]].y;
        }
    return nmatches;
}

// ORBmatcher.cc:226-428 (Nleft == -1 branch)
int orc_search_by_bow(const orc_frame_view *KF, const uint8_t *kf_mp_valid, const orc_frame_view *F, int kf_nnodes,
                      const int32_t *kf_nodes, const int32_t *kf_ptr, const int32_t *kf_idx, int f_nnodes,
                      const int32_t *f_nodes, const int32_t *f_ptr, const int32_t *f_idx, float nnratio, int check_ori,
                      int32_t *matches_f) {
    return orc_search_by_bow_2cam(KF, kf_mp_valid, F, -1, kf_nnodes, kf_nodes, kf_ptr, kf_idx, f_nnodes, f_nodes, f_ptr, f_idx,
                                  nnratio, check_ori, matches_f);
}

// ORBmatcher.cc:226-428 incl. the F.Nleft != -1 branch (:298-322, :362-390): keys / descriptors of F are the left
// camera's [0, f_nleft) followed by the right camera's; f_nleft == -1 is the single-camera method.
int orc_search_by_bow_2cam(const orc_frame_view *KF, const uint8_t *kf_mp_valid, const orc_frame_view *F, int f_nleft,
                           int kf_nnodes, const int32_t *kf_nodes, const int32_t *kf_ptr, const int32_t *kf_idx, int f_nnodes,
                           const int32_t *f_nodes, const int32_t *f_ptr, const int32_t *f_idx, float nnratio, int check_ori,
                           int32_t *matches_f) {
    for (int i = 0; i < F->n; ++i) matches_f[i] = -1;
    int nmatches = 0;
    std::vector<int> rot_hist[HISTO_LENGTH];
    int a = 0, b = 0;
    while (a < kf_nnodes && b < f_nnodes) {
        if (kf_nodes[a] == f_nodes[b]) {
            for (int ik = kf_ptr[a]; ik < kf_ptr[a + 1]; ++ik) {
                const int real_kf = kf_idx[ik];
                if (!kf_mp_valid[real_kf]) continue;
                const uint8_t *d_kf = KF->descriptors + (size_t)real_kf * 32;
                int best1 = 256, best_idx_f = -1, best2 = 256;
                int best1r = 256, best_idx_fr = -1, best2r = 256;
                for (int jf = f_ptr[b]; jf < f_ptr[b + 1]; ++jf) {
                    const int real_f = f_idx[jf];
                    if (matches_f[real_f] >= 0) continue;
                    const int dist = descriptor_distance(d_kf, F->descriptors + (size_t)real_f * 32);
                    if (f_nleft == -1) {
                        if (dist < best1) { best2 = best1; best1 = dist; best_idx_f = real_f; }
                        else if (dist < best2) best2 = dist;
                    } else {
                        if (real_f < f_nleft && dist < best1) { best2 = best1; best1 = dist; best_idx_f = real_f; }
                        else if (real_f < f_nleft && dist < best2) best2 = dist;
                        if (real_f >= f_nleft && dist < best1r) { best2r = best1r; best1r = dist; best_idx_fr = real_f; }
                        else if (real_f >= f_nleft && dist < best2r) best2r = dist;
                    }
                }
                if (best1 <= TH_LOW) {
                    if (static_cast<float>(best1) < nnratio * static_cast<float>(best2)) {
                        matches_f[best_idx_f] = real_kf;
                        if (check_ori) rot_hist[rot_bin(KF->keys[real_kf].angle, F->keys[best_idx_f].angle)].push_back(best_idx_f);
                        ++nmatches;
                    }
                    if (best1r <= TH_LOW) {
                        if (static_cast<float>(best1r) < nnratio * static_cast<float>(best2r) || true) {
                            matches_f[best_idx_fr] = real_kf;
                            if (check_ori) rot_hist[rot_bin(KF->keys[real_kf].angle, F->keys[best_idx_fr].angle)].push_back(best_idx_fr);
                            ++nmatches;
                        }
                    }
                }
            }
            ++a; ++b;
        } else if (kf_nodes[a] < f_nodes[b]) {
            while (a < kf_nnodes && kf_nodes[a] < f_nodes[b]) ++a;   // lower_bound
        } else {
            while (b < f_nnodes && f_nodes[b] < kf_nodes[a]) ++b;
        }
    }
    if (check_ori) {
        int sizes[HISTO_LENGTH], ind1 = -1, ind2 = -1, ind3 = -1;
        for (int i = 0; i < HISTO_LENGTH; ++i) sizes[i] = (int)rot_hist[i].size();
        three_maxima(sizes, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int idx : rot_hist[i]) { matches_f[idx] = -1; --nmatches; }
        }
    }
    return nmatches;
}

// KeyFrame::GetFeaturesInArea (KeyFrame.cc:834-875): the Frame version without the level filter
static void kf_features_in_area(const orc_frame_view *f, const Grid &g, float x, float y, float r, std::vector<int> &out) {
    features_in_area(f, g, x, y, r, -1, -1, out);
}

// ORBmatcher.cc:1880-2000 (projection and the gates ahead of GetFeaturesInArea are the caller's: pts[i].valid)
int orc_search_by_projection_reloc(const orc_frame_view *Cur, const uint8_t *occupied_in, int n,
                                   const orc_search_point *pts, const uint8_t *desc, float th, int orb_dist,
                                   int check_ori, int32_t *assign) {
    Grid g(Cur);
    std::vector<uint8_t> has_mp(occupied_in, occupied_in + Cur->n);
    for (int i = 0; i < Cur->n; ++i) assign[i] = -1;
    std::vector<int> rot_hist[HISTO_LENGTH];
    int nmatches = 0;
    std::vector<int> cand;
    for (int i = 0; i < n; ++i) {
        const orc_search_point &p = pts[i];
        if (!p.valid) continue;
        const int level = p.level;
        const float radius = th * Cur->scale_factors[level];
        features_in_area(Cur, g, p.u, p.v, radius, level - 1, level + 1, cand);
        if (cand.empty()) continue;
        int best = 256, best_idx2 = -1;
        for (int i2 : cand) {
            if (has_mp[i2]) continue;
            const int dist = descriptor_distance(desc + (size_t)i * 32, Cur->descriptors + (size_t)i2 * 32);
            if (dist < best) { best = dist; best_idx2 = i2; }
        }
        if (best <= orb_dist && best_idx2 >= 0) {
            assign[best_idx2] = i;
            has_mp[best_idx2] = 1;
            ++nmatches;
            if (check_ori) rot_hist[rot_bin(p.angle, Cur->keys[best_idx2].angle)].push_back(best_idx2);
        }
    }
    if (check_ori) {
        int sizes[HISTO_LENGTH], ind1 = -1, ind2 = -1, ind3 = -1;
        for (int i = 0; i < HISTO_LENGTH; ++i) sizes[i] = (int)rot_hist[i].size();
        three_maxima(sizes, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; ++i)
            if (i != ind1 && i != ind2 && i != ind3)
                for (int idx : rot_hist[i]) { assign[idx] = -2; --nmatches; }
    }
    return nmatches;
}

// ORBmatcher.cc:430-528 / :530-641
int orc_search_by_projection_sim3(const orc_frame_view *KF, const uint8_t *matched_in, int n, const orc_search_point *pts,
                                  const uint8_t *desc, int th, float ratio_hamming, int32_t *assign) {
    Grid g(KF);
    std::vector<uint8_t> matched(matched_in, matched_in + KF->n);
    for (int i = 0; i < KF->n; ++i) assign[i] = -1;
    int nmatches = 0;
    std::vector<int> cand;
    for (int i = 0; i < n; ++i) {
        const orc_search_point &p = pts[i];
        if (!p.valid) continue;
        const int level = p.level;
        const float radius = th * KF->scale_factors[level];
        kf_features_in_area(KF, g, p.u, p.v, radius, cand);
        if (cand.empty()) continue;
        int best = 256, best_idx = -1;
        for (int idx : cand) {
            if (matched[idx]) continue;
            const int kp_level = KF->keys[idx].octave;
            if (kp_level < level - 1 || kp_level > level) continue;
            const int dist = descriptor_distance(desc + (size_t)i * 32, KF->descriptors + (size_t)idx * 32);
            if (dist < best) { best = dist; best_idx = idx; }
        }
        if (best <= TH_LOW * ratio_hamming && best_idx >= 0) {
            assign[best_idx] = i;
            matched[best_idx] = 1;
            ++nmatches;
        }
    }
    return nmatches;
}

// the search of ORBmatcher.cc:1148-1335 (variant 0) / :1337-1446 (variant 1); the MapPoint bookkeeping is the caller's
int orc_fuse_search(const orc_frame_view *KF, int n, const orc_search_point *pts, const uint8_t *desc, float th,
                    const float *inv_level_sigma2, int variant, int32_t *best_idx_out) {
    Grid g(KF);
    int nfused = 0;
    std::vector<int> cand;
    for (int i = 0; i < n; ++i) {
        best_idx_out[i] = -1;
        const orc_search_point &p = pts[i];
        if (!p.valid) continue;
        const int level = p.level;
        const float radius = th * KF->scale_factors[level];
        kf_features_in_area(KF, g, p.u, p.v, radius, cand);
        if (cand.empty()) continue;
        int best = variant == 0 ? 256 : INT_MAX, best_idx = -1;
        for (int idx : cand) {
            const orc_keypoint &kp = KF->keys[idx];
            const int kp_level = kp.octave;
            if (kp_level < level - 1 || kp_level > level) continue;
            if (variant == 0) {
                if (KF->u_right && KF->u_right[idx] >= 0) {
                    const float kpx = kp.x, kpy = kp.y, kpr = KF->u_right[idx];
                    const float ex = p.u - kpx, ey = p.v - kpy, er = p.ur - kpr;
                    const float e2 = ex * ex + ey * ey + er * er;
                    if (e2 * inv_level_sigma2[kp_level] > 7.8) continue;
                } else {
                    const float kpx = kp.x, kpy = kp.y;
                    const float ex = p.u - kpx, ey = p.v - kpy;
                    const float e2 = ex * ex + ey * ey;
                    if (e2 * inv_level_sigma2[kp_level] > 5.99) continue;
                }
            }
            const int dist = descriptor_distance(desc + (size_t)i * 32, KF->descriptors + (size_t)idx * 32);
            if (dist < best) { best = dist; best_idx = idx; }
        }
        if (best <= TH_LOW) {
            best_idx_out[i] = best_idx;
            ++nfused;
        }
    }
    return nfused;
}

// ORBmatcher.cc:1448-1665
int orc_search_by_sim3(const orc_frame_view *KF1, const orc_frame_view *KF2, const orc_search_point *pts1,
                       const uint8_t *desc1, const orc_search_point *pts2, const uint8_t *desc2, float th,
                       int32_t *matches12) {
    return orc_search_by_sim3_n(KF1, KF2, KF1->n, pts1, desc1, KF2->n, pts2, desc2, th, matches12);
}

// the same with n1 / n2 map point slots that may exceed the searched feature counts: two-camera keyframes, whose
// GetMapPointMatches() has Nleft + Nright entries while GetFeaturesInArea only walks the left camera (:1488-1652)
int orc_search_by_sim3_n(const orc_frame_view *KF1, const orc_frame_view *KF2, int n1, const orc_search_point *pts1,
                         const uint8_t *desc1, int n2, const orc_search_point *pts2, const uint8_t *desc2, float th,
                         int32_t *matches12) {
    std::vector<int> match1(n1, -1), match2(n2, -1);
    std::vector<int> cand;
    for (int dir = 0; dir < 2; ++dir) {
        const orc_frame_view *T = dir == 0 ? KF2 : KF1;
        const orc_search_point *pts = dir == 0 ? pts1 : pts2;
        const uint8_t *desc = dir == 0 ? desc1 : desc2;
        std::vector<int> &out = dir == 0 ? match1 : match2;
        Grid g(T);
        for (int i = 0; i < (dir == 0 ? n1 : n2); ++i) {
            if (!pts[i].valid) continue;
            const int level = pts[i].level;
            const float radius = th * T->scale_factors[level];
            kf_features_in_area(T, g, pts[i].u, pts[i].v, radius, cand);
            if (cand.empty()) continue;
            int best = INT_MAX, best_idx = -1;
            for (int idx : cand) {
                const orc_keypoint &kp = T->keys[idx];
                if (kp.octave < level - 1 || kp.octave > level) continue;
                const int dist = descriptor_distance(desc + (size_t)i * 32, T->descriptors + (size_t)idx * 32);
                if (dist < best) { best = dist; best_idx = idx; }
            }
            if (best <= TH_HIGH) out[i] = best_idx;
        }
    }
    int nfound = 0;
    for (int i1 = 0; i1 < n1; ++i1) {
        matches12[i1] = -1;
        const int idx2 = match1[i1];
        if (idx2 >= 0) {
            const int idx1 = match2[idx2];
            if (idx1 == i1) { matches12[i1] = idx2; ++nfound; }
        }
    }
    return nfound;
}

// ORBmatcher.cc:758-900 (NLeft == -1)
int orc_search_by_bow_kf(const orc_frame_view *KF1, const uint8_t *mp_valid1, const orc_frame_view *KF2,
                         const uint8_t *mp_valid2, int nn1, const int32_t *nodes1, const int32_t *ptr1,
                         const int32_t *idx1v, int nn2, const int32_t *nodes2, const int32_t *ptr2, const int32_t *idx2v,
                         float nnratio, int check_ori, int32_t *matches12) {
    for (int i = 0; i < KF1->n; ++i) matches12[i] = -1;
    std::vector<bool> matched2(KF2->n, false);
    std::vector<int> rot_hist[HISTO_LENGTH];
    int nmatches = 0, a = 0, b = 0;
    while (a < nn1 && b < nn2) {
        if (nodes1[a] == nodes2[b]) {
            for (int k1 = ptr1[a]; k1 < ptr1[a + 1]; ++k1) {
                const int idx1 = idx1v[k1];
                if (!mp_valid1[idx1]) continue;
                const uint8_t *d1 = KF1->descriptors + (size_t)idx1 * 32;
                int best1 = 256, best_idx2 = -1, best2 = 256;
                for (int k2 = ptr2[b]; k2 < ptr2[b + 1]; ++k2) {
                    const int idx2 = idx2v[k2];
                    if (matched2[idx2] || !mp_valid2[idx2]) continue;
                    const int dist = descriptor_distance(d1, KF2->descriptors + (size_t)idx2 * 32);
                    if (dist < best1) { best2 = best1; best1 = dist; best_idx2 = idx2; }
                    else if (dist < best2) best2 = dist;
                }
                if (best1 < TH_LOW) {
                    if (static_cast<float>(best1) < nnratio * static_cast<float>(best2)) {
                        matches12[idx1] = best_idx2;
                        matched2[best_idx2] = true;
                        if (check_ori) rot_hist[rot_bin(KF1->keys[idx1].angle, KF2->keys[best_idx2].angle)].push_back(idx1);
                        ++nmatches;
                    }
                }
            }
            ++a; ++b;
        } else if (nodes1[a] < nodes2[b]) {
            while (a < nn1 && nodes1[a] < nodes2[b]) ++a;
        } else {
            while (b < nn2 && nodes2[b] < nodes1[a]) ++b;
        }
    }
    if (check_ori) {
        int sizes[HISTO_LENGTH], ind1 = -1, ind2 = -1, ind3 = -1;
        for (int i = 0; i < HISTO_LENGTH; ++i) sizes[i] = (int)rot_hist[i].size();
        three_maxima(sizes, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int i1 : rot_hist[i]) { matches12[i1] = -1; --nmatches; }
        }
    }
    return nmatches;
}

// Pinhole::epipolarConstrain (Pinhole.cpp:118-141) with F12 precomputed by the caller
static bool epipolar_constrain(const float *F12, const orc_keypoint &kp1, const orc_keypoint &kp2, float unc) {
    const float a = kp1.x * F12[0] + kp1.y * F12[3] + F12[6];
    const float b = kp1.x * F12[1] + kp1.y * F12[4] + F12[7];
    const float c = kp1.x * F12[2] + kp1.y * F12[5] + F12[8];
    const float num = a * kp2.x + b * kp2.y + c;
    const float den = a * a + b * b;
    if (den == 0) return false;
    const float dsqr = num * num / den;
    return dsqr < 3.84 * unc;
}

// ORBmatcher.cc:902-1146 (mpCamera2 == NULL)
int orc_search_for_triangulation(const orc_frame_view *KF1, const uint8_t *has_mp1, const orc_frame_view *KF2,
                                 const uint8_t *has_mp2, int nn1, const int32_t *nodes1, const int32_t *ptr1,
                                 const int32_t *idx1v, int nn2, const int32_t *nodes2, const int32_t *ptr2,
                                 const int32_t *idx2v, int only_stereo, int coarse, const float *f12, const float *ep,
                                 const float *level_sigma2_2, int check_ori, int32_t *matches12) {
    for (int i = 0; i < KF1->n; ++i) matches12[i] = -1;
    std::vector<bool> matched2(KF2->n, false);   // declared and tested by the reference, never set (SURVEY C#3)
    std::vector<int> rot_hist[HISTO_LENGTH];
    int nmatches = 0, a = 0, b = 0;
    while (a < nn1 && b < nn2) {
        if (nodes1[a] == nodes2[b]) {
            for (int k1 = ptr1[a]; k1 < ptr1[a + 1]; ++k1) {
                const int idx1 = idx1v[k1];
                if (has_mp1[idx1]) continue;
                const bool stereo1 = KF1->u_right && KF1->u_right[idx1] >= 0;
                if (only_stereo && !stereo1) continue;
                const orc_keypoint &kp1 = KF1->keys[idx1];
                const uint8_t *d1 = KF1->descriptors + (size_t)idx1 * 32;
                int best = TH_LOW, best_idx2 = -1;
                for (int k2 = ptr2[b]; k2 < ptr2[b + 1]; ++k2) {
                    const int idx2 = idx2v[k2];
                    if (matched2[idx2] || has_mp2[idx2]) continue;
                    const bool stereo2 = KF2->u_right && KF2->u_right[idx2] >= 0;
                    if (only_stereo && !stereo2) continue;
                    const int dist = descriptor_distance(d1, KF2->descriptors + (size_t)idx2 * 32);
                    if (dist > TH_LOW || dist > best) continue;
                    const orc_keypoint &kp2 = KF2->keys[idx2];
                    if (!stereo1 && !stereo2) {
                        const float distex = ep[0] - kp2.x, distey = ep[1] - kp2.y;
                        if (distex * distex + distey * distey < 100 * KF2->scale_factors[kp2.octave]) continue;
                    }
                    if (coarse || epipolar_constrain(f12, kp1, kp2, level_sigma2_2[kp2.octave])) {
                        best_idx2 = idx2;
                        best = dist;
                    }
                }
                if (best_idx2 >= 0) {
                    matches12[idx1] = best_idx2;
                    ++nmatches;
                    if (check_ori) rot_hist[rot_bin(kp1.angle, KF2->keys[best_idx2].angle)].push_back(idx1);
                }
            }
            ++a; ++b;
        } else if (nodes1[a] < nodes2[b]) {
            while (a < nn1 && nodes1[a] < nodes2[b]) ++a;
        } else {
            while (b < nn2 && nodes2[b] < nodes1[a]) ++b;
        }
    }
    if (check_ori) {
        int sizes[HISTO_LENGTH], ind1 = -1, ind2 = -1, ind3 = -1;
        for (int i = 0; i < HISTO_LENGTH; ++i) sizes[i] = (int)rot_hist[i].size();
        three_maxima(sizes, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int i1 : rot_hist[i]) { matches12[i1] = -1; --nmatches; }
        }
    }
    return nmatches;
}


// MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:383-409) for one point: N descriptors -> BestIdx
int orc_distinctive_descriptor(const uint8_t *desc, int n) {
    if (n <= 0) return -1;
    std::vector<float> distances((size_t)n * n);
    for (int i = 0; i < n; ++i) {
        distances[(size_t)i * n + i] = 0;
        for (int j = i + 1; j < n; ++j) {
            const int distij = descriptor_distance(desc + (size_t)i * 32, desc + (size_t)j * 32);
            distances[(size_t)i * n + j] = (float)distij;
            distances[(size_t)j * n + i] = (float)distij;
        }
    }
    int best_median = INT_MAX, best_idx = 0;
    for (int i = 0; i < n; ++i) {
        std::vector<int> v(distances.begin() + (size_t)i * n, distances.begin() + (size_t)(i + 1) * n);
        std::sort(v.begin(), v.end());
        const int median = v[(size_t)(0.5 * (n - 1))];
        if (median < best_median) { best_median = median; best_idx = i; }
    }
    return best_idx;
}

// TemplatedVocabulary::transform(feature, word_id, weight, nid, levelsup) (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:
// 1225-1265) on a flattened tree (children of node i = child_idx[child_ptr[i] .. child_ptr[i+1]), FORB::distance =
// the 256-bit Hamming distance, FORB.cpp:81-101).  leaf[i] = final node, nid[i] = node at level L - levelsup.
void orc_bow_transform(const int32_t *child_ptr, const int32_t *child_idx, const uint8_t *node_desc, int levels,
                       const uint8_t *desc, int n, int levelsup, int32_t *leaf, int32_t *nid_out) {
    for (int f = 0; f < n; ++f) {
        const uint8_t *feature = desc + (size_t)f * 32;
        const int nid_level = levels - levelsup;
        int nid = 0;                                  // `if (nid_level <= 0 && nid != NULL) *nid = 0;  // root`
        int final_id = 0, current_level = 0;
        do {
            ++current_level;
            const int cb = child_ptr[final_id], ce = child_ptr[final_id + 1];
            final_id = child_idx[cb];
            double best_d = (double)descriptor_distance(feature, node_desc + (size_t)final_id * 32);
            for (int c = cb + 1; c < ce; ++c) {
                const int id = child_idx[c];
                const double d = (double)descriptor_distance(feature, node_desc + (size_t)id * 32);
                if (d < best_d) { best_d = d; final_id = id; }
            }
            if (current_level == nid_level) nid = final_id;
        } while (child_ptr[final_id] != child_ptr[final_id + 1]);   // !isLeaf()
        leaf[f] = final_id;
        nid_out[f] = nid;
    }
}

}  // extern "C"
