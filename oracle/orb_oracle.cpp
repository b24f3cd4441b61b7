/*
 * oracle/orb_oracle.cpp — CPU ORACLE for the ORB extractor (test infrastructure, NOT product code).
 *
 * A from-scratch restatement of the reference algorithm in orb_slam3/src/ORBextractor.cc
 * (snt-arg/visual_sgraphs).  Every function names the reference lines it follows.  OpenCV is not
 * available as a C++ library here, so each OpenCV primitive the reference delegates to is restated
 * from its published fixed-point algorithm and checked bit-for-bit against
 * python cv2 4.13.0 by tests/test_oracle_cv2.py and the fixtures under tests/golden/.
 *
 * Parity status: pinned to cv2 4.13.0 for the primitives and to libstdc++ (std::sort / std::list)
 * for the oct-tree; unpinned with respect to reference-owned golden vectors because the reference
 * ships none.
 *
 * Build: see oracle/Makefile (g++ -O3 -ffp-contract=off, no -march=native — the reference is built
 * with plain -O3, CMakeLists.txt:10-14).
 */
#include "oracle.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cfloat>
#include <climits>
#include <cstdlib>
#include <cstring>
#include <list>
#include <thread>
#include <utility>
#include <vector>

namespace {

constexpr int kPatch = 31;      // PATCH_SIZE        ORBextractor.cc:69
constexpr int kHalfPatch = 15;  // HALF_PATCH_SIZE   ORBextractor.cc:70
constexpr int kEdge = 19;       // EDGE_THRESHOLD    ORBextractor.cc:71

const int8_t kPattern[1024] = {
#include "orb_pattern.inc"
};

// ------------------------------------------------------------------------------------------------
// OpenCV scalar helpers: cvRound = round-half-to-even.
// ------------------------------------------------------------------------------------------------
inline int cv_round(float v) { return (int)lrintf(v); }
inline int cv_round(double v) { return (int)lrint(v); }

// ------------------------------------------------------------------------------------------------
// A plain 8-bit image with a pitch.
// ------------------------------------------------------------------------------------------------
struct Image {
    int w = 0, h = 0, pitch = 0;
    std::vector<uint8_t> buf;
    uint8_t *origin = nullptr;  // pixel (0,0); may sit inside buf (bordered pyramid level)
    void alloc(int w_, int h_) {
        w = w_; h = h_; pitch = w_;
        buf.assign((size_t)w * h, 0);
        origin = buf.data();
    }
    void alloc_bordered(int w_, int h_, int border) {
        w = w_; h = h_; pitch = w_ + 2 * border;
        buf.assign((size_t)pitch * (h_ + 2 * border), 0);
        origin = buf.data() + (size_t)border * pitch + border;
    }
    const uint8_t *row(int y) const { return origin + (ptrdiff_t)y * pitch; }
    uint8_t *row(int y) { return origin + (ptrdiff_t)y * pitch; }
};

// ------------------------------------------------------------------------------------------------
// cv::resize(src, dst, dsize, 0, 0, INTER_LINEAR) for CV_8UC1 — SURVEY Appendix A1.
// Call site: ORBextractor.cc:1184.  11-bit fixed-point coefficients; the horizontal pass keeps
// int32 rows, the vertical pass is ((b0*(r0>>4))>>16) + ((b1*(r1>>4))>>16) + 2 >> 2.
// ------------------------------------------------------------------------------------------------
void resize_linear(const uint8_t *src, int sw, int sh, int spitch, uint8_t *dst, int dw, int dh, int dpitch) {
    const double scale_x = 1.0 / ((double)dw / sw);
    const double scale_y = 1.0 / ((double)dh / sh);
    std::vector<int> xofs(dw), yofs(dh);
    std::vector<short> xa(2 * dw), ya(2 * dh);
    for (int dx = 0; dx < dw; ++dx) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = (int)std::floor(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        xofs[dx] = sx;
        xa[2 * dx] = (short)cv_round((1.f - fx) * 2048.f);
        xa[2 * dx + 1] = (short)cv_round(fx * 2048.f);
    }
    for (int dy = 0; dy < dh; ++dy) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = (int)std::floor(fy);
        fy -= sy;
        yofs[dy] = sy;  // rows are clipped below, the weight is not reset (OpenCV resizeGeneric_)
        ya[2 * dy] = (short)cv_round((1.f - fy) * 2048.f);
        ya[2 * dy + 1] = (short)cv_round(fy * 2048.f);
    }
    std::vector<int> r0(dw), r1(dw);
    auto hrow = [&](int sy, std::vector<int> &out) {
        sy = std::min(std::max(sy, 0), sh - 1);
        const uint8_t *s = src + (size_t)sy * spitch;
        for (int dx = 0; dx < dw; ++dx) {
            int sx = xofs[dx];
            int sx1 = std::min(sx + 1, sw - 1);
            out[dx] = s[sx] * xa[2 * dx] + s[sx1] * xa[2 * dx + 1];
        }
    };
    for (int dy = 0; dy < dh; ++dy) {
        hrow(yofs[dy], r0);
        hrow(yofs[dy] + 1, r1);
        const int b0 = ya[2 * dy], b1 = ya[2 * dy + 1];
        uint8_t *d = dst + (size_t)dy * dpitch;
        for (int dx = 0; dx < dw; ++dx) {
            int v = (((b0 * (r0[dx] >> 4)) >> 16) + ((b1 * (r1[dx] >> 4)) >> 16) + 2) >> 2;
            d[dx] = (uint8_t)std::min(std::max(v, 0), 255);
        }
    }
}

// cv::copyMakeBorder(..., BORDER_REFLECT_101) index map — SURVEY Appendix A3 (ORBextractor.cc:1186-1192)
inline int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = (i < 0) ? -i : 2 * n - 2 - i;
    return i;
}

// Fill the `border`-pixel frame around a w x h image that already sits in the middle of dst.
void fill_border_reflect101(Image &im, int border) {
    for (int y = -border; y < im.h + border; ++y) {
        const uint8_t *s = im.row(reflect101(y, im.h));
        uint8_t *d = im.row(y);
        for (int x = -border; x < im.w + border; ++x) {
            if (y >= 0 && y < im.h && x >= 0 && x < im.w) continue;
            d[x] = s[reflect101(x, im.w)];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// cv::GaussianBlur(m, m, Size(7,7), 2, 2, BORDER_REFLECT_101) on a standalone CV_8UC1 image —
// SURVEY Appendix A2 (ORBextractor.cc:1129-1130).  8.8 fixed-point taps, single rounding at the end.
// ------------------------------------------------------------------------------------------------
const int kGauss[7] = {18, 34, 48, 56, 48, 34, 18};

void gaussian_blur7(const uint8_t *src, int w, int h, int spitch, uint8_t *dst, int dpitch) {
    std::vector<uint16_t> hbuf((size_t)w * h);
    for (int y = 0; y < h; ++y) {
        const uint8_t *s = src + (size_t)y * spitch;
        uint16_t *o = hbuf.data() + (size_t)y * w;
        for (int x = 0; x < w; ++x) {
            int acc = 0;
            for (int k = -3; k <= 3; ++k) acc += kGauss[k + 3] * s[reflect101(x + k, w)];
            o[x] = (uint16_t)acc;  // <= 255*256
        }
    }
    for (int y = 0; y < h; ++y) {
        uint8_t *d = dst + (size_t)y * dpitch;
        const uint16_t *rows[7];
        for (int k = -3; k <= 3; ++k) rows[k + 3] = hbuf.data() + (size_t)reflect101(y + k, h) * w;
        for (int x = 0; x < w; ++x) {
            uint32_t acc = 0;
            for (int k = 0; k < 7; ++k) acc += (uint32_t)kGauss[k] * rows[k][x];
            d[x] = (uint8_t)((acc + 32768u) >> 16);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// cv::FAST(img, kps, t, true) == FAST-9/16 with 3x3 non-max suppression — SURVEY Appendix A6, row a5.
// Call sites ORBextractor.cc:832-833, 850-851.  A pixel is a corner at threshold t iff its strength
// K (largest over the sixteen 9-arcs of the smallest same-signed |v-p| on the arc) exceeds t; its
// score is K-1.  Only [3,w-3) x [3,h-3) is examined; everything else scores 0 for the suppression.
// ------------------------------------------------------------------------------------------------
const int kRingDx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
const int kRingDy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

struct FastPoint { int x, y, score; };

inline int fast_strength(const int *d /* 25 entries: v - ring[k], wrapped */) {
    int best = 0;
    for (int k = 0; k < 16; ++k) {
        int mn = d[k], mx = d[k];
        for (int j = 1; j < 9; ++j) {
            mn = std::min(mn, d[k + j]);
            mx = std::max(mx, d[k + j]);
        }
        best = std::max(best, mn);
        best = std::max(best, -mx);
    }
    return best;
}

void fast_detect(const uint8_t *img, int w, int h, int pitch, int threshold, std::vector<FastPoint> &out) {
    out.clear();
    if (w < 7 || h < 7) return;
    threshold = std::min(std::max(threshold, 0), 255);
    int off[16];
    for (int k = 0; k < 16; ++k) off[k] = kRingDy[k] * pitch + kRingDx[k];
    // score plane for the interior, 0 elsewhere
    std::vector<int> score((size_t)w * h, 0);
    for (int y = 3; y < h - 3; ++y) {
        const uint8_t *p = img + (size_t)y * pitch;
        for (int x = 3; x < w - 3; ++x) {
            const uint8_t *c = p + x;
            const int v = c[0];
            const int lo = v - threshold, hi = v + threshold;
            // high-speed rejection: a 9-arc always contains one pixel of each opposite pair
            auto cls = [&](int k) { int q = c[off[k]]; return (q < lo ? 1 : 0) | (q > hi ? 2 : 0); };
            int m = cls(0) | cls(8);
            if (!m) continue;
            m &= cls(2) | cls(10); if (!m) continue;
            m &= cls(4) | cls(12); if (!m) continue;
            m &= cls(6) | cls(14); if (!m) continue;
            m &= cls(1) | cls(9);  if (!m) continue;
            m &= cls(3) | cls(11); if (!m) continue;
            m &= cls(5) | cls(13); if (!m) continue;
            m &= cls(7) | cls(15); if (!m) continue;
            int d[25];
            for (int k = 0; k < 16; ++k) d[k] = v - c[off[k]];
            for (int k = 0; k < 9; ++k) d[16 + k] = d[k];
            const int K = fast_strength(d);
            if (K > threshold) score[(size_t)y * w + x] = K - 1;
        }
    }
    for (int y = 3; y < h - 3; ++y) {
        const int *s = score.data() + (size_t)y * w;
        for (int x = 3; x < w - 3; ++x) {
            const int v = s[x];
            // non-corners score 0; a corner scoring 0 (only possible at t == 0) can never be a strict
            // maximum over zeros either — same outcome as OpenCV's score buffers.
            if (v == 0) continue;
            if (v > s[x - 1] && v > s[x + 1] && v > s[x - w - 1] && v > s[x - w] && v > s[x - w + 1] &&
                v > s[x + w - 1] && v > s[x + w] && v > s[x + w + 1])
                out.push_back({x, y, v});
        }
    }
}

// ------------------------------------------------------------------------------------------------
// cv::fastAtan2 — SURVEY Appendix A4 (called at ORBextractor.cc:99).  float32, no FMA.
// ------------------------------------------------------------------------------------------------
float fast_atan2(float y, float x) {
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * scale;
    const float p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale;
    const float p7 = -0.04432655554792128f * scale;
    const float eps = (float)DBL_EPSILON;
    const float ax = std::fabs(x), ay = std::fabs(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + eps);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + eps);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

// ------------------------------------------------------------------------------------------------
// ORBextractor tables — ORBextractor.cc:411-470
// ------------------------------------------------------------------------------------------------
struct Tables {
    int nfeatures, nlevels, ini_th, min_th;
    double scale_factor;  // the member is a double initialised from the float argument (ORBextractor.h:105)
    std::vector<float> scale, inv_scale, sigma2, inv_sigma2;
    std::vector<int> quota;
    int umax[kHalfPatch + 1];
};

void build_tables(Tables &t, int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th) {
    t.nfeatures = nfeatures; t.nlevels = nlevels; t.ini_th = ini_th; t.min_th = min_th;
    t.scale_factor = scale_factor;
    t.scale.assign(nlevels, 1.f); t.sigma2.assign(nlevels, 1.f);
    for (int i = 1; i < nlevels; ++i) {                      // :415-423
        t.scale[i] = (float)(t.scale[i - 1] * t.scale_factor);  // float * double -> double -> float
        t.sigma2[i] = t.scale[i] * t.scale[i];
    }
    t.inv_scale.resize(nlevels); t.inv_sigma2.resize(nlevels);
    for (int i = 0; i < nlevels; ++i) {                      // :425-431
        t.inv_scale[i] = 1.0f / t.scale[i];
        t.inv_sigma2[i] = 1.0f / t.sigma2[i];
    }
    t.quota.assign(nlevels, 0);                              // :435-446
    const float factor = (float)(1.0f / t.scale_factor);
    float desired = nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; ++l) {
        t.quota[l] = cv_round(desired);
        sum += t.quota[l];
        desired *= factor;
    }
    t.quota[nlevels - 1] = std::max(nfeatures - sum, 0);
    // circular patch row ends, :454-469
    const int vmax = (int)std::floor(kHalfPatch * std::sqrt(2.f) / 2 + 1);
    const int vmin = (int)std::ceil(kHalfPatch * std::sqrt(2.f) / 2);
    const double hp2 = kHalfPatch * kHalfPatch;
    for (int v = 0; v <= vmax; ++v) t.umax[v] = cv_round(std::sqrt(hp2 - v * v));
    for (int v = kHalfPatch, v0 = 0; v >= vmin; --v) {
        while (t.umax[v0] == t.umax[v0 + 1]) ++v0;
        t.umax[v] = v0;
        ++v0;
    }
}

// ------------------------------------------------------------------------------------------------
// IC_Angle — ORBextractor.cc:73-100 (intensity centroid over the radius-15 disc, un-blurred level)
// ------------------------------------------------------------------------------------------------
float ic_angle(const uint8_t *center, int step, const int *umax) {
    int m01 = 0, m10 = 0;
    for (int u = -kHalfPatch; u <= kHalfPatch; ++u) m10 += u * center[u];
    for (int v = 1; v <= kHalfPatch; ++v) {
        int vsum = 0;
        const int d = umax[v];
        for (int u = -d; u <= d; ++u) {
            const int below = center[u + v * step], above = center[u - v * step];
            vsum += below - above;
            m10 += u * (below + above);
        }
        m01 += v * vsum;
    }
    return fast_atan2((float)m01, (float)m10);
}

// ------------------------------------------------------------------------------------------------
// computeOrbDescriptor — ORBextractor.cc:103-149 (steered BRIEF-256 on the blurred level)
// ------------------------------------------------------------------------------------------------
void orb_descriptor(const uint8_t *center, int step, float angle_deg, uint8_t *desc) {
    const float factor_pi = (float)(3.14159265358979323846 / 180.f);   // :102
    const float angle = angle_deg * factor_pi;
    const float a = cosf(angle), b = sinf(angle);                      // float overloads (:108)
    auto sample = [&](int px, int py) -> int {
        const int ry = cv_round(px * b + py * a);
        const int rx = cv_round(px * a - py * b);
        return center[ry * step + rx];
    };
    const int8_t *p = kPattern;
    for (int i = 0; i < 32; ++i) {
        int val = 0;
        for (int bit = 0; bit < 8; ++bit, p += 4) {
            const int t0 = sample(p[0], p[1]);
            const int t1 = sample(p[2], p[3]);
            val |= (t0 < t1) << bit;
        }
        desc[i] = (uint8_t)val;
    }
}

// ------------------------------------------------------------------------------------------------
// Oct-tree distribution — ExtractorNode::DivideNode :482-537, compareNodes :539-560,
// DistributeOctTree :562-785.  Same containers as the reference (std::list with push_front,
// std::sort) because list order and the introsort tie order are part of the result.
// ------------------------------------------------------------------------------------------------
struct Cand { float x, y, response; int src; };

struct Node {
    std::vector<Cand> keys;
    int ulx = 0, uly = 0, urx = 0, ury = 0, blx = 0, bly = 0, brx = 0, bry = 0;
    std::list<Node>::iterator self;
    bool no_more = false;

    void divide(Node &n1, Node &n2, Node &n3, Node &n4) const {
        const int half_x = (int)std::ceil(static_cast<float>(urx - ulx) / 2);
        const int half_y = (int)std::ceil(static_cast<float>(bry - uly) / 2);
        n1.ulx = ulx;          n1.uly = uly;
        n1.urx = ulx + half_x; n1.ury = uly;
        n1.blx = ulx;          n1.bly = uly + half_y;
        n1.brx = ulx + half_x; n1.bry = uly + half_y;
        n2.ulx = n1.urx; n2.uly = n1.ury;
        n2.urx = urx;    n2.ury = ury;
        n2.blx = n1.brx; n2.bly = n1.bry;
        n2.brx = urx;    n2.bry = uly + half_y;
        n3.ulx = n1.blx; n3.uly = n1.bly;
        n3.urx = n1.brx; n3.ury = n1.bry;
        n3.blx = blx;    n3.bly = bly;
        n3.brx = n1.brx; n3.bry = bly;
        n4.ulx = n3.urx; n4.uly = n3.ury;
        n4.urx = n2.brx; n4.ury = n2.bry;
        n4.blx = n3.brx; n4.bly = n3.bry;
        n4.brx = brx;    n4.bry = bry;
        for (const Cand &k : keys) {
            if (k.x < n1.urx) {
                if (k.y < n1.bry) n1.keys.push_back(k); else n3.keys.push_back(k);
            } else if (k.y < n1.bry) n2.keys.push_back(k);
            else n4.keys.push_back(k);
        }
        if (n1.keys.size() == 1) n1.no_more = true;
        if (n2.keys.size() == 1) n2.no_more = true;
        if (n3.keys.size() == 1) n3.no_more = true;
        if (n4.keys.size() == 1) n4.no_more = true;
    }
};

typedef std::pair<int, Node *> SizedNode;

bool sized_node_less(SizedNode &e1, SizedNode &e2) {   // compareNodes :539-560
    if (e1.first < e2.first) return true;
    if (e1.first > e2.first) return false;
    return e1.second->ulx < e2.second->ulx;
}

std::vector<Cand> distribute_octree(const std::vector<Cand> &cands, int min_x, int max_x, int min_y, int max_y, int N) {
    const int n_ini = (int)std::round(static_cast<float>(max_x - min_x) / (max_y - min_y));
    const float hx = static_cast<float>(max_x - min_x) / n_ini;
    std::list<Node> nodes;
    std::vector<Node *> roots(n_ini);
    for (int i = 0; i < n_ini; ++i) {
        Node r;
        r.ulx = (int)(hx * static_cast<float>(i));       r.uly = 0;
        r.urx = (int)(hx * static_cast<float>(i + 1));   r.ury = 0;
        r.blx = r.ulx; r.bly = max_y - min_y;
        r.brx = r.urx; r.bry = max_y - min_y;
        nodes.push_back(r);
        roots[i] = &nodes.back();
    }
    for (const Cand &k : cands) roots[(size_t)(k.x / hx)]->keys.push_back(k);

    for (auto it = nodes.begin(); it != nodes.end();) {
        if (it->keys.size() == 1) { it->no_more = true; ++it; }
        else if (it->keys.empty()) it = nodes.erase(it);
        else ++it;
    }

    auto push_children = [&](Node *kids[4], std::vector<SizedNode> &expandable, int *n_expand) {
        for (int c = 0; c < 4; ++c) {
            if (kids[c]->keys.empty()) continue;
            nodes.push_front(*kids[c]);
            if (kids[c]->keys.size() > 1) {
                if (n_expand) ++*n_expand;
                expandable.push_back(std::make_pair((int)kids[c]->keys.size(), &nodes.front()));
                nodes.front().self = nodes.begin();
            }
        }
    };

    bool finished = false;
    std::vector<SizedNode> expandable;
    while (!finished) {
        int prev_size = (int)nodes.size();
        int n_expand = 0;
        expandable.clear();
        for (auto it = nodes.begin(); it != nodes.end();) {      // :626-684
            if (it->no_more) { ++it; continue; }
            Node n1, n2, n3, n4;
            it->divide(n1, n2, n3, n4);
            Node *kids[4] = {&n1, &n2, &n3, &n4};
            push_children(kids, expandable, &n_expand);
            it = nodes.erase(it);
        }
        if ((int)nodes.size() >= N || (int)nodes.size() == prev_size) {
            finished = true;
        } else if ((int)nodes.size() + n_expand * 3 > N) {        // :696-759
            while (!finished) {
                prev_size = (int)nodes.size();
                std::vector<SizedNode> prev = expandable;
                expandable.clear();
                std::sort(prev.begin(), prev.end(), sized_node_less);
                for (int j = (int)prev.size() - 1; j >= 0; --j) {
                    Node n1, n2, n3, n4;
                    prev[j].second->divide(n1, n2, n3, n4);
                    Node *kids[4] = {&n1, &n2, &n3, &n4};
                    push_children(kids, expandable, nullptr);
                    nodes.erase(prev[j].second->self);
                    if ((int)nodes.size() >= N) break;
                }
                if ((int)nodes.size() >= N || (int)nodes.size() == prev_size) finished = true;
            }
        }
    }

    std::vector<Cand> result;                                     // :766-782, first maximum wins
    result.reserve(nodes.size());
    for (const Node &n : nodes) {
        const Cand *best = &n.keys[0];
        float max_response = best->response;
        for (size_t k = 1; k < n.keys.size(); ++k)
            if (n.keys[k].response > max_response) { best = &n.keys[k]; max_response = n.keys[k].response; }
        result.push_back(*best);
    }
    return result;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// The extractor object: ORBextractor::operator() and its helpers
// ------------------------------------------------------------------------------------------------
struct orc_extractor {
    Tables t;
    std::vector<Image> pyramid;                 // mvImagePyramid, each with its 19-px border
    std::vector<Image> blurred;                 // the per-level workingMat (w==0 if skipped)
    std::vector<std::vector<Cand>> candidates;  // vToDistributeKeys per level
    std::vector<std::vector<orc_keypoint>> level_kps;
    std::vector<orc_keypoint> kps;
    std::vector<uint8_t> desc;
    int mono_index = 0;
    std::vector<FastPoint> cell_pts;

    // ComputePyramid — ORBextractor.cc:1171-1195
    void compute_pyramid(const uint8_t *img, int w, int h, int pitch) {
        pyramid.resize(t.nlevels);
        for (int level = 0; level < t.nlevels; ++level) {
            const float scale = t.inv_scale[level];
            const int lw = cv_round((float)w * scale), lh = cv_round((float)h * scale);
            Image &L = pyramid[level];
            L.alloc_bordered(lw, lh, kEdge);
            if (level != 0) {
                const Image &P = pyramid[level - 1];
                resize_linear(P.origin, P.w, P.h, P.pitch, L.origin, lw, lh, L.pitch);
            } else {
                for (int y = 0; y < h; ++y) std::memcpy(L.row(y), img + (size_t)y * pitch, w);
            }
            fill_border_reflect101(L, kEdge);
        }
    }

    // ComputeKeyPointsOctTree — ORBextractor.cc:787-900
    void compute_keypoints() {
        candidates.assign(t.nlevels, {});
        level_kps.assign(t.nlevels, {});
        const float W = 35;
        for (int level = 0; level < t.nlevels; ++level) {
            const Image &L = pyramid[level];
            const int min_bx = kEdge - 3, min_by = min_bx;
            const int max_bx = L.w - kEdge + 3, max_by = L.h - kEdge + 3;
            std::vector<Cand> &cands = candidates[level];
            const float width = (float)(max_bx - min_bx), height = (float)(max_by - min_by);
            const int n_cols = (int)(width / W), n_rows = (int)(height / W);
            const int w_cell = (int)std::ceil(width / n_cols), h_cell = (int)std::ceil(height / n_rows);
            for (int i = 0; i < n_rows; ++i) {
                const float ini_y = (float)(min_by + i * h_cell);
                float max_y = ini_y + h_cell + 6;
                if (ini_y >= max_by - 3) continue;
                if (max_y > max_by) max_y = (float)max_by;
                for (int j = 0; j < n_cols; ++j) {
                    const float ini_x = (float)(min_bx + j * w_cell);
                    float max_x = ini_x + w_cell + 6;
                    if (ini_x >= max_bx - 6) continue;
                    if (max_x > max_bx) max_x = (float)max_bx;
                    const int x0 = (int)ini_x, y0 = (int)ini_y, cw = (int)max_x - x0, ch = (int)max_y - y0;
                    const uint8_t *cell = L.row(y0) + x0;
                    fast_detect(cell, cw, ch, L.pitch, t.ini_th, cell_pts);
                    if (cell_pts.empty()) fast_detect(cell, cw, ch, L.pitch, t.min_th, cell_pts);
                    for (const FastPoint &p : cell_pts)
                        cands.push_back({(float)p.x + j * w_cell, (float)p.y + i * h_cell, (float)p.score,
                                         (int)cands.size()});
                }
            }
            std::vector<Cand> sel = distribute_octree(cands, min_bx, max_bx, min_by, max_by, t.quota[level]);
            const int scaled_patch = (int)(kPatch * t.scale[level]);
            std::vector<orc_keypoint> &out = level_kps[level];
            out.reserve(sel.size());
            for (const Cand &c : sel) {
                orc_keypoint k;
                k.x = c.x + min_bx; k.y = c.y + min_by;
                k.size = (float)scaled_patch; k.angle = -1.f; k.response = c.response;
                k.octave = level; k.class_id = -1;
                out.push_back(k);
            }
        }
        for (int level = 0; level < t.nlevels; ++level) {          // computeOrientation :472-480
            const Image &L = pyramid[level];
            for (orc_keypoint &k : level_kps[level])
                k.angle = ic_angle(L.row(cv_round(k.y)) + cv_round(k.x), L.pitch, t.umax);
        }
    }

    // operator() — ORBextractor.cc:1083-1169
    int run(const uint8_t *img, int w, int h, int pitch, int lap0, int lap1) {
        kps.clear(); desc.clear(); mono_index = 0;
        if (!img || w <= 0 || h <= 0) return -1;
        compute_pyramid(img, w, h, pitch);
        compute_keypoints();
        int total = 0;
        for (auto &v : level_kps) total += (int)v.size();
        kps.assign(total, orc_keypoint());
        desc.assign((size_t)total * 32, 0);
        blurred.assign(t.nlevels, Image());
        int mono = 0, stereo = total - 1;
        for (int level = 0; level < t.nlevels; ++level) {
            std::vector<orc_keypoint> lk = level_kps[level];
            if (lk.empty()) continue;
            const Image &L = pyramid[level];
            Image &B = blurred[level];
            B.alloc(L.w, L.h);
            gaussian_blur7(L.origin, L.w, L.h, L.pitch, B.origin, B.pitch);
            const float scale = t.scale[level];
            for (orc_keypoint &k : lk) {
                uint8_t d[32];
                orb_descriptor(B.row(cv_round(k.y)) + cv_round(k.x), B.pitch, k.angle, d);
                if (level != 0) { k.x *= scale; k.y *= scale; }
                int slot;
                if (k.x >= (float)lap0 && k.x <= (float)lap1) slot = stereo--; else slot = mono++;
                kps[slot] = k;
                std::memcpy(&desc[(size_t)slot * 32], d, 32);
            }
        }
        mono_index = mono;
        return mono;
    }
};

// ------------------------------------------------------------------------------------------------
// C interface
// ------------------------------------------------------------------------------------------------
extern "C" {

orc_extractor *orc_extractor_create(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th) {
    orc_extractor *ex = new orc_extractor();
    build_tables(ex->t, nfeatures, scale_factor, nlevels, ini_th, min_th);
    return ex;
}
void orc_extractor_destroy(orc_extractor *ex) { delete ex; }
int orc_levels(const orc_extractor *ex) { return ex->t.nlevels; }
void orc_scale_factors(const orc_extractor *ex, float *s, float *is, float *s2, float *is2) {
    for (int i = 0; i < ex->t.nlevels; ++i) {
        if (s) s[i] = ex->t.scale[i];
        if (is) is[i] = ex->t.inv_scale[i];
        if (s2) s2[i] = ex->t.sigma2[i];
        if (is2) is2[i] = ex->t.inv_sigma2[i];
    }
}
void orc_quotas(const orc_extractor *ex, int32_t *q) { for (int i = 0; i < ex->t.nlevels; ++i) q[i] = ex->t.quota[i]; }
void orc_umax(const orc_extractor *ex, int32_t *u) { for (int i = 0; i < 16; ++i) u[i] = ex->t.umax[i]; }

int orc_extract(orc_extractor *ex, const uint8_t *img, int w, int h, int pitch, int lap0, int lap1) {
    return ex->run(img, w, h, pitch, lap0, lap1);
}
int orc_num_keypoints(const orc_extractor *ex) { return (int)ex->kps.size(); }
void orc_get_keypoints(const orc_extractor *ex, orc_keypoint *kps, uint8_t *desc) {
    if (kps) std::memcpy(kps, ex->kps.data(), ex->kps.size() * sizeof(orc_keypoint));
    if (desc) std::memcpy(desc, ex->desc.data(), ex->desc.size());
}
void orc_level_size(const orc_extractor *ex, int level, int32_t *w, int32_t *h) {
    *w = ex->pyramid[level].w; *h = ex->pyramid[level].h;
}
void orc_get_level(const orc_extractor *ex, int level, uint8_t *dst) {
    const Image &L = ex->pyramid[level];
    for (int y = 0; y < L.h; ++y) std::memcpy(dst + (size_t)y * L.w, L.row(y), L.w);
}
void orc_get_level_padded(const orc_extractor *ex, int level, uint8_t *dst) {
    const Image &L = ex->pyramid[level];
    std::memcpy(dst, L.buf.data(), L.buf.size());
}
int orc_get_blurred(const orc_extractor *ex, int level, uint8_t *dst) {
    const Image &B = ex->blurred[level];
    if (B.w == 0) return 0;
    std::memcpy(dst, B.buf.data(), B.buf.size());
    return 1;
}
int orc_num_candidates(const orc_extractor *ex, int level) { return (int)ex->candidates[level].size(); }
void orc_get_candidates(const orc_extractor *ex, int level, float *xyr) {
    for (const Cand &c : ex->candidates[level]) { *xyr++ = c.x; *xyr++ = c.y; *xyr++ = c.response; }
}
int orc_num_level_keypoints(const orc_extractor *ex, int level) { return (int)ex->level_kps[level].size(); }
void orc_get_level_keypoints(const orc_extractor *ex, int level, orc_keypoint *kps) {
    std::memcpy(kps, ex->level_kps[level].data(), ex->level_kps[level].size() * sizeof(orc_keypoint));
}

void orc_resize_linear(const uint8_t *src, int sw, int sh, int sp, uint8_t *dst, int dw, int dh, int dp) {
    resize_linear(src, sw, sh, sp, dst, dw, dh, dp);
}
void orc_gaussian_blur7(const uint8_t *src, int w, int h, int sp, uint8_t *dst, int dp) {
    gaussian_blur7(src, w, h, sp, dst, dp);
}
void orc_border_reflect101(const uint8_t *src, int w, int h, int sp, uint8_t *dst, int border, int dp) {
    for (int y = -border; y < h + border; ++y) {
        const uint8_t *s = src + (size_t)reflect101(y, h) * sp;
        uint8_t *d = dst + (size_t)(y + border) * dp + border;
        for (int x = -border; x < w + border; ++x) d[x] = s[reflect101(x, w)];
    }
}
int orc_fast(const uint8_t *img, int w, int h, int pitch, int threshold, int32_t *xys, int cap) {
    std::vector<FastPoint> pts;
    fast_detect(img, w, h, pitch, threshold, pts);
    int n = 0;
    for (const FastPoint &p : pts) {
        if (n < cap) { xys[3 * n] = p.x; xys[3 * n + 1] = p.y; xys[3 * n + 2] = p.score; }
        ++n;
    }
    return n;
}
float orc_fast_atan2(float y, float x) { return fast_atan2(y, x); }
int orc_cv_round_f(float v) { return cv_round(v); }
int orc_cv_round_d(double v) { return cv_round(v); }
float orc_ic_angle(const uint8_t *img, int pitch, int x, int y) {
    Tables t;
    build_tables(t, 1000, 1.2f, 8, 20, 7);
    return ic_angle(img + (size_t)y * pitch + x, pitch, t.umax);
}
void orc_orb_descriptor(const uint8_t *img, int pitch, int x, int y, float angle_deg, uint8_t *desc32) {
    orb_descriptor(img + (size_t)y * pitch + x, pitch, angle_deg, desc32);
}
int orc_distribute_octree(const float *xyr, int n, int min_x, int max_x, int min_y, int max_y, int quota,
                          int32_t *selected_idx, int cap) {
    std::vector<Cand> c(n);
    for (int i = 0; i < n; ++i) c[i] = {xyr[3 * i], xyr[3 * i + 1], xyr[3 * i + 2], i};
    std::vector<Cand> sel = distribute_octree(c, min_x, max_x, min_y, max_y, quota);
    int m = 0;
    for (const Cand &s : sel) { if (m < cap) selected_idx[m] = s.src; ++m; }
    return m;
}

// Frame::ComputeStereoMatches — orb_slam3/src/Frame.cc:957-1127
void orc_stereo_matches(const orc_extractor *left, const orc_extractor *right, const orc_keypoint *keys_l,
                        const uint8_t *desc_l, int n_l, const orc_keypoint *keys_r, const uint8_t *desc_r, int n_r,
                        float mb, float mbf, float *u_right, float *depth) {
    const int TH_HIGH = 100, TH_LOW = 50;
    for (int i = 0; i < n_l; ++i) { u_right[i] = -1.0f; depth[i] = -1.0f; }
    const int th_orb_dist = (TH_HIGH + TH_LOW) / 2;
    const int n_rows = left->pyramid[0].h;
    const std::vector<float> &scale = left->t.scale, &inv_scale = left->t.inv_scale;
    std::vector<std::vector<int>> row_indices(n_rows);
    for (int ir = 0; ir < n_r; ++ir) {                                       // :973-984
        const float kp_y = keys_r[ir].y;
        const float r = 2.0f * scale[keys_r[ir].octave];
        const int maxr = (int)std::ceil(kp_y + r), minr = (int)std::floor(kp_y - r);
        for (int yi = minr; yi <= maxr; ++yi)
            if (yi >= 0 && yi < n_rows) row_indices[yi].push_back(ir);       // the reference has no bounds check
    }
    const float min_z = mb, min_d = 0, max_d = mbf / min_z;
    std::vector<std::pair<int, int>> dist_idx;
    for (int il = 0; il < n_l; ++il) {
        const orc_keypoint &kpl = keys_l[il];
        const int level_l = kpl.octave;
        const float vl = kpl.y, ul = kpl.x;
        const size_t row = (size_t)vl;
        if (row >= (size_t)n_rows) continue;
        const std::vector<int> &cands = row_indices[row];
        if (cands.empty()) continue;
        const float min_u = ul - max_d, max_u = ul - min_d;
        if (max_u < 0) continue;
        int best_dist = TH_HIGH;
        size_t best_idx_r = 0;
        for (int ir : cands) {                                               // :1017-1040
            const orc_keypoint &kpr = keys_r[ir];
            if (kpr.octave < level_l - 1 || kpr.octave > level_l + 1) continue;
            const float ur = kpr.x;
            if (ur >= min_u && ur <= max_u) {
                const int dist = orc_descriptor_distance(desc_l + (size_t)il * 32, desc_r + (size_t)ir * 32);
                if (dist < best_dist) { best_dist = dist; best_idx_r = ir; }
            }
        }
        if (best_dist < th_orb_dist) {                                       // :1043-1111
            const float ur0 = keys_r[best_idx_r].x;
            const float scale_factor = inv_scale[kpl.octave];
            const float scaled_ul = std::round(kpl.x * scale_factor);
            const float scaled_vl = std::round(kpl.y * scale_factor);
            const float scaled_ur0 = std::round(ur0 * scale_factor);
            const int w = 5, L = 5;
            const Image &PL = left->pyramid[kpl.octave], &PR = right->pyramid[kpl.octave];
            int best_sad = INT_MAX, best_inc_r = 0;
            float dists[2 * 5 + 1];
            const float iniu = scaled_ur0 + L - w, endu = scaled_ur0 + L + w + 1;
            if (iniu < 0 || endu >= PR.w) continue;
            const int y0 = (int)(scaled_vl - w), xl0 = (int)(scaled_ul - w);
            for (int inc = -L; inc <= +L; ++inc) {
                const int xr0 = (int)(scaled_ur0 + inc - w);
                long sad = 0;                                                // cv::norm(IL, IR, NORM_L1)
                for (int dy = 0; dy < 2 * w + 1; ++dy) {
                    const uint8_t *a = PL.row(y0 + dy) + xl0, *b = PR.row(y0 + dy) + xr0;
                    for (int dx = 0; dx < 2 * w + 1; ++dx) sad += std::abs((int)a[dx] - (int)b[dx]);
                }
                const float dist = (float)(double)sad;
                if (dist < best_sad) { best_sad = (int)dist; best_inc_r = inc; }
                dists[L + inc] = dist;
            }
            if (best_inc_r == -L || best_inc_r == L) continue;
            const float d1 = dists[L + best_inc_r - 1], d2 = dists[L + best_inc_r], d3 = dists[L + best_inc_r + 1];
            const float delta_r = (d1 - d3) / (2.0f * (d1 + d3 - 2.0f * d2));
            if (delta_r < -1 || delta_r > 1) continue;
            float best_ur = scale[kpl.octave] * ((float)scaled_ur0 + (float)best_inc_r + delta_r);
            float disparity = (ul - best_ur);
            if (disparity >= min_d && disparity < max_d) {
                if (disparity <= 0) { disparity = 0.01; best_ur = ul - 0.01; }
                depth[il] = mbf / disparity;
                u_right[il] = best_ur;
                dist_idx.push_back(std::pair<int, int>(best_sad, il));
            }
        }
    }
    if (dist_idx.empty()) return;                                            // the reference reads [0] here (SURVEY C#12)
    std::sort(dist_idx.begin(), dist_idx.end());
    const float median = dist_idx[dist_idx.size() / 2].first;
    const float th_dist = 1.5f * 1.4f * median;
    for (int i = (int)dist_idx.size() - 1; i >= 0; --i) {
        if (dist_idx[i].first < th_dist) break;
        u_right[dist_idx[i].second] = -1;
        depth[dist_idx[i].second] = -1;
    }
}

// cv::cvtColor(src, dst, COLOR_{RGB,BGR,RGBA,BGRA}2GRAY) for 8-bit images, as run by Tracking::GrabImageRGBD /
// GrabImageMonocular / GrabImageStereo ahead of the extractor (Tracking.cc:1526-1551, 1595-1608, 1646-1660).
// OpenCV's fixed-point path: 15-bit coefficients R 9798, G 19235, B 3735, round to nearest (pinned against cv2 4.13).
void orc_cvt_gray(const uint8_t *src, int w, int h, int pitch, int channels, int r_first, uint8_t *dst, int dst_pitch) {
    const int ri = r_first ? 0 : 2, bi = r_first ? 2 : 0;
    for (int y = 0; y < h; ++y) {
        const uint8_t *s = src + (size_t)y * pitch;
        uint8_t *d = dst + (size_t)y * dst_pitch;
        for (int x = 0; x < w; ++x, s += channels)
            d[x] = (uint8_t)((s[ri] * 9798 + s[1] * 19235 + s[bi] * 3735 + (1 << 14)) >> 15);
    }
}

// cv::remap(src, dst, map1, map2, cv::INTER_LINEAR) with CV_32FC1 maps, 8-bit single-channel images and the default
// BORDER_CONSTANT / 0 border, as System::TrackStereo / TrackMonocular run it on every incoming image when the settings
// ask for rectification (System.cc:284-292, 352-…; maps from cv::initUndistortRectifyMap, Settings.cc:571-574).
// OpenCV (un-vendored; imgproc remap, restated from its published algorithm and pinned against cv2 4.13 in
// tests/test_oracle_cv2.py): the map is quantised to 1/32 pixel, sx = cvRound(mapx * 32), integer part sx >> 5
// (saturated to short), fraction sx & 31; the four bilinear weights are the exact products (32-fy)(32-fx) * 32 ... of the
// 15-bit table; dst = (sum of tap * weight + 2^14) >> 15; taps outside the source count as 0.
void orc_remap_bilinear(const uint8_t *src, int sw, int sh, int spitch, const float *map_x, const float *map_y, int w, int h,
                        uint8_t *dst, int dst_pitch) {
    auto sat_short = [](int v) { return v < -32768 ? -32768 : (v > 32767 ? 32767 : v); };
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const int sx = cv_round(map_x[(size_t)y * w + x] * 32.f), sy = cv_round(map_y[(size_t)y * w + x] * 32.f);
            const int ix = sat_short(sx >> 5), iy = sat_short(sy >> 5), fx = sx & 31, fy = sy & 31;
            const int w00 = (32 - fy) * (32 - fx) * 32, w01 = (32 - fy) * fx * 32, w10 = fy * (32 - fx) * 32, w11 = fy * fx * 32;
            auto tap = [&](int px, int py) -> int {
                return (px >= 0 && px < sw && py >= 0 && py < sh) ? src[(size_t)py * spitch + px] : 0;
            };
            const int v = tap(ix, iy) * w00 + tap(ix + 1, iy) * w01 + tap(ix, iy + 1) * w10 + tap(ix + 1, iy + 1) * w11;
            dst[(size_t)y * dst_pitch + x] = (uint8_t)((v + (1 << 14)) >> 15);
        }
}

// Frame::UndistortKeyPoints (Frame.cc:891-922) = cv::undistortPoints(mat, mat, mK, mDistCoef, cv::Mat(), mK) on the N x 2
// float keypoint coordinates; also what Frame::ComputeImageBounds (:924-955) runs on the four image corners.
// OpenCV (un-vendored; calib3d cvUndistortPointsInternal, restated from its published algorithm and pinned against
// cv2 4.13 in tests/test_oracle_cv2.py): all arithmetic in double, normalise with 1/fx and 1/fy, five fixed-point
// iterations of the Brown-Conrady model (the C++ wrapper's default TermCriteria(MAX_ITER, 5, 0.01)), k[0..13] =
// (k1 k2 p1 p2 k3 k4 k5 k6 s1 s2 s3 s4 tx ty) zero-padded (the tilt terms must be zero here), a negative icdist
// falls back to the normalised input, re-projection with P = K, results stored as float.  dist_n == 0 or k1 == 0
// follows the reference's shortcut (:893-897): the coordinates are copied.
void orc_undistort_points(int n, const float *xy_in, double fx, double fy, double cx, double cy, const double *dist,
                          int dist_n, float *xy_out) {
    if (dist_n <= 0 || dist[0] == 0.0) {
        for (int i = 0; i < 2 * n; ++i) xy_out[i] = xy_in[i];
        return;
    }
    double k[14] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < dist_n && i < 12; ++i) k[i] = dist[i];
    const double ifx = 1. / fx, ify = 1. / fy;
    for (int i = 0; i < n; ++i) {
        const double u = xy_in[2 * i], v = xy_in[2 * i + 1];
        double x = (u - cx) * ifx, y = (v - cy) * ify;
        const double x0 = x, y0 = y;
        for (int j = 0; j < 5; ++j) {
            const double r2 = x * x + y * y;
            const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
            if (icdist < 0) {
                x = (u - cx) * ifx;
                y = (v - cy) * ify;
                break;
            }
            const double dx = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
            const double dy = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
            x = (x0 - dx) * icdist;
            y = (y0 - dy) * icdist;
        }
        const double xx = fx * x + 0. * y + cx, yy = 0. * x + fy * y + cy, ww = 1. / (0. * x + 0. * y + 1.);
        xy_out[2 * i] = (float)(xx * ww);
        xy_out[2 * i + 1] = (float)(yy * ww);
    }
}

// Frame::ComputeStereoFromRGBD (Frame.cc:1129-1150): depth at the (truncated) distorted keypoint position; where it is
// positive, mvDepth = d and mvuRight = undistorted x - bf / d; -1 elsewhere.
void orc_stereo_from_rgbd(int n, const float *xy, const float *xy_un, const float *depth, int depth_pitch_floats, float bf,
                          float *u_right, float *depth_out) {
    for (int i = 0; i < n; ++i) {
        u_right[i] = -1.f;
        depth_out[i] = -1.f;
        const float d = depth[(size_t)(int)xy[2 * i + 1] * depth_pitch_floats + (int)xy[2 * i]];
        if (d > 0) {
            depth_out[i] = d;
            u_right[i] = xy_un[2 * i] - bf / d;
        }
    }
}

double orc_bench_extract(const uint8_t *frames, int nframes, int w, int h, int nfeatures, float scale_factor,
                         int nlevels, int ini_th, int min_th, int threads, int64_t *total_keypoints) {
    if (threads < 1) threads = 1;
    std::vector<int64_t> counts(threads, 0);
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int tid = 0; tid < threads; ++tid) {
        pool.emplace_back([&, tid]() {
            orc_extractor *ex = orc_extractor_create(nfeatures, scale_factor, nlevels, ini_th, min_th);
            for (int f = tid; f < nframes; f += threads) {
                ex->run(frames + (size_t)f * w * h, w, h, w, 0, 0);
                counts[tid] += (int64_t)ex->kps.size();
            }
            orc_extractor_destroy(ex);
        });
    }
    for (auto &th : pool) th.join();
    auto t1 = std::chrono::steady_clock::now();
    if (total_keypoints) { *total_keypoints = 0; for (auto c : counts) *total_keypoints += c; }
    return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
