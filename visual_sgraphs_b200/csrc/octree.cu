// octree.cu — DistributeOctTree on the device, node-for-node, one CTA per (pyramid level, frame).
//
// Reference (snt-arg/visual_sgraphs):
//   ExtractorNode::DivideNode            orb_slam3/src/ORBextractor.cc:482-537
//   compareNodes                         :539-560
//   ORBextractor::DistributeOctTree      :562-785
//
// How the sequential std::list algorithm is restated for a CTA:
//   * A node's key set is a pure function of the node's box path: DivideNode sends a key to a child
//     by comparing its coordinates with the box midpoint only.  So keys never move in memory; each
//     key carries the list position of its node (`node_of`) and every pass re-labels it.
//   * One pass of the reference's main loop divides EVERY multi-key node.  All threads histogram
//     their keys into the four children of their node (shared-memory atomics); the list is then
//     rebuilt in parallel exactly as push_front/erase would leave it — block-wide prefix sums over the
//     list give every child its position: children of the i-th divided node, in n1..n4 order, end up
//     in front of everything pushed before them; single-key nodes keep their relative order behind.
//   * The "careful" phase (:696-759) sorts the expandable nodes with the libstdc++ introsort
//     emulation (introsort.cuh — tie order is part of the result; one thread, a few hundred items),
//     finds how many are divided before the quota is reached, and rebuilds the list the same way.
//   * Final pick per node = highest response, FIRST in the reference's candidate order on ties
//     (:766-782).  Candidate order is cell-row-major, then row-major inside the cell (:811-874), so
//     the tie break is an atomicMax on (response, ~order_key(x, y)).
#include "introsort.cuh"
#include "vsg_internal.cuh"

namespace vsg {

struct ONode {
    short x0, y0, x1, y1;  // UL.x, UL.y, UR.x (== BR.x), BR.y (== BL.y), relative to minBorder
    int count;             // vKeys.size(); bNoMore <=> count == 1
};

__device__ __forceinline__ int quadrant(const ONode &n, int kx, int ky) {
    const int mx = n.x0 + ((n.x1 - n.x0 + 1) >> 1);  // UL.x + ceil((UR.x-UL.x)/2)
    const int my = n.y0 + ((n.y1 - n.y0 + 1) >> 1);
    // n1: left/top, n2: right/top, n3: left/bottom, n4: right/bottom   (:514-527)
    return (kx < mx ? 0 : 1) + (ky < my ? 0 : 2);
}

__device__ __forceinline__ ONode child_of(const ONode &n, int q, int count) {
    const int mx = n.x0 + ((n.x1 - n.x0 + 1) >> 1);
    const int my = n.y0 + ((n.y1 - n.y0 + 1) >> 1);
    ONode c;
    c.x0 = (q & 1) ? mx : n.x0;
    c.x1 = (q & 1) ? n.x1 : mx;
    c.y0 = (q & 2) ? my : n.y0;
    c.y1 = (q & 2) ? n.y1 : my;
    c.count = count;
    return c;
}

// Exclusive prefix sums of three per-thread values over the 256 threads of the CTA (thread order), in place;
// the block totals are returned in ta, tb, tc.  Contains two barriers.
__device__ __forceinline__ void block_scan3(int &a, int &b, int &c, int (*s_warp)[8], int &ta, int &tb, int &tc) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int ia = a, ib = b, ic = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int xa = __shfl_up_sync(0xffffffffu, ia, d), xb = __shfl_up_sync(0xffffffffu, ib, d),
                  xc = __shfl_up_sync(0xffffffffu, ic, d);
        if (lane >= d) { ia += xa; ib += xb; ic += xc; }
    }
    if (lane == 31) { s_warp[0][warp] = ia; s_warp[1][warp] = ib; s_warp[2][warp] = ic; }
    __syncthreads();
    int wa = 0, wb = 0, wc = 0;
    ta = tb = tc = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        const int va = s_warp[0][w], vb = s_warp[1][w], vc = s_warp[2][w];
        if (w < warp) { wa += va; wb += vb; wc += vc; }
        ta += va; tb += vb; tc += vc;
    }
    __syncthreads();
    a = wa + ia - a; b = wb + ib - b; c = wc + ic - c;
}

__global__ void __launch_bounds__(256) octree_kernel(FrameGeom g, const Cand *__restrict__ cand,
                                                     const int *__restrict__ cand_count,
                                                     unsigned short *__restrict__ node_of,
                                                     LevelKp *__restrict__ level_kps,
                                                     int *__restrict__ level_kp_count, int cap) {
    extern __shared__ __align__(16) uint8_t smem[];
    ONode *list_a = reinterpret_cast<ONode *>(smem);
    ONode *list_b = list_a + cap;
    int *cc = reinterpret_cast<int *>(list_b + cap);                       // [cap][4] child key counts
    unsigned long long *best = reinterpret_cast<unsigned long long *>(cc); // reused after the tree is final
    unsigned short *child_pos = reinterpret_cast<unsigned short *>(cc + 4 * cap);  // [cap][4]
    unsigned short *stay_pos = child_pos + 4 * cap;                        // [cap]
    short *expanded = reinterpret_cast<short *>(stay_pos + cap);           // [cap] 1 if divided this round
    unsigned short *exp_list = reinterpret_cast<unsigned short *>(expanded + cap);  // [cap] expandable nodes, push order
    SortItem *sort_buf = reinterpret_cast<SortItem *>(exp_list + cap + (cap & 1));  // [cap]
    __shared__ int s_size, s_nexp, s_state, s_E, s_T, s_X;
    __shared__ int s_scan[3][8];

    const int level = blockIdx.x, frame = blockIdx.y;
    const LevelGeom &L = g.lv[level];
    const int tid = threadIdx.x;
    const int n = min(cand_count[frame * g.nlevels + level], L.cand_cap);
    const Cand *keys = cand + L.cand_offset + (int64_t)frame * g.cand_total;
    const uint32_t *key_xy = reinterpret_cast<const uint32_t *>(keys);   // word 2k = x | y << 16 of candidate k
    unsigned short *nof = node_of + L.cand_offset + (int64_t)frame * g.cand_total;
    LevelKp *out = level_kps + (int64_t)frame * g.kp_total + L.kp_offset;
    int *out_count = level_kp_count + frame * g.nlevels + level;
    if (n == 0) {
        if (tid == 0) *out_count = 0;
        return;
    }
    const int N = L.quota;
    const int max_y = L.h - 2 * kBorderMin;  // maxBorderY - minBorderY

    // ---- roots (:566-593) ----
    ONode *cur = list_a, *nxt = list_b;
    for (int i = tid; i < L.n_ini; i += 256) {
        ONode r;
        r.x0 = (short)(int)__fmul_rn(L.h_x, (float)i);
        r.x1 = (short)(int)__fmul_rn(L.h_x, (float)(i + 1));
        r.y0 = 0;
        r.y1 = (short)max_y;
        r.count = 0;
        cur[i] = r;
    }
    __syncthreads();
    for (int k = tid; k < n; k += 256) {
        const int r = (int)__fdiv_rn((float)(keys[k].x - kBorderMin), L.h_x);
        nof[k] = (unsigned short)r;
        atomicAdd(&cur[r].count, 1);
    }
    __syncthreads();
    if (tid == 0) {  // drop empty roots, keep order (:597-608)
        int m = 0;
        for (int i = 0; i < L.n_ini; ++i) {
            stay_pos[i] = (unsigned short)m;
            if (cur[i].count > 0) nxt[m++] = cur[i];
        }
        s_size = m;
        s_state = 0;  // 0 = main passes, 1 = careful phase, 2 = finished
    }
    __syncthreads();
    for (int k = tid; k < n; k += 256) nof[k] = stay_pos[nof[k]];
    { ONode *t = cur; cur = nxt; nxt = t; }
    __syncthreads();

    // ---- subdivision rounds ----
    while (true) {
        const int size = s_size;
        const int state = s_state;
        if (state == 2) break;
        for (int i = tid; i < size * 4; i += 256) cc[i] = 0;
        __syncthreads();
        for (int k0 = tid; k0 < n; k0 += 4 * 256) {       // 4 keys per thread in flight (global loads are L2 latency)
            int nd[4];
            uint32_t xy[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int k = k0 + u * 256;
                nd[u] = k < n ? nof[k] : -1;
                xy[u] = k < n ? key_xy[2 * k] : 0u;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (nd[u] < 0) continue;
                const ONode &node = cur[nd[u]];
                if (node.count > 1)
                    atomicAdd(&cc[nd[u] * 4 + quadrant(node, (int)(xy[u] & 0xFFFF) - kBorderMin, (int)(xy[u] >> 16) - kBorderMin)], 1);
            }
        }
        __syncthreads();

        // every thread owns a contiguous chunk of the list so that prefix sums follow list order
        const int chunk = (size + 255) / 256;
        const int lo = min(tid * chunk, size), hi = min(lo + chunk, size);
        if (state == 0) {
            // main pass (:626-684): every multi-key node is divided, walking the list front to back; children
            // are pushed to the FRONT in n1..n4 order, single-key nodes keep their relative order behind them
            int kids = 0, stays = 0, exps = 0;
            for (int i = lo; i < hi; ++i) {
                if (cur[i].count == 1) { ++stays; continue; }
                for (int q = 0; q < 4; ++q) { kids += cc[i * 4 + q] > 0; exps += cc[i * 4 + q] > 1; }
            }
            int T, S, X;
            block_scan3(kids, stays, exps, s_scan, T, S, X);     // in: my sums, out: exclusive prefixes; totals in T,S,X
            int push = kids, stay = stays, nexp = exps;
            for (int i = lo; i < hi; ++i) {
                if (cur[i].count == 1) {
                    const int pos = T + stay++;
                    nxt[pos] = cur[i];
                    stay_pos[i] = (unsigned short)pos;
                    expanded[i] = 0;
                } else {
                    expanded[i] = 1;
                    for (int q = 0; q < 4; ++q) {
                        const int c = cc[i * 4 + q];
                        if (c == 0) continue;
                        const int pos = T - 1 - push++;
                        nxt[pos] = child_of(cur[i], q, c);
                        child_pos[i * 4 + q] = (unsigned short)pos;
                        if (c > 1) exp_list[nexp++] = (unsigned short)pos;
                    }
                }
            }
            if (tid == 0) {
                const int new_size = T + S;
                if (new_size >= N || new_size == size) s_state = 2;           // :690-694
                else if (new_size + X * 3 > N) s_state = 1;                   // :696
                s_size = new_size;
                s_nexp = X;
            }
        } else {
            // careful round (:698-759): sort the expandable nodes, divide from the back until size >= N
            const int m = s_nexp;
            for (int j = tid; j < m; j += 256) {
                const int pos = exp_list[j];
                sort_buf[j].count = cur[pos].count;
                sort_buf[j].ulx = cur[pos].x0;
                sort_buf[j].ref = pos;
            }
            for (int i = tid; i < size; i += 256) expanded[i] = 0;
            __syncthreads();
            if (tid == 0) {
                libstdcxx_sort(sort_buf, m);
                // processing order r = 0.. : sort_buf[m-1-r]; push_off[r] / exp_off[r] = children / expandable children
                // pushed before r.  They overwrite the (count, ulx) fields that are no longer needed.
                int running = size, E = 0, T = 0, X = 0;
                for (int j = m - 1; j >= 0; --j) {
                    const int pos = sort_buf[j].ref;
                    int c = 0, e = 0;
                    for (int q = 0; q < 4; ++q) { c += cc[pos * 4 + q] > 0; e += cc[pos * 4 + q] > 1; }
                    sort_buf[j].count = T;
                    sort_buf[j].ulx = X;
                    expanded[pos] = 1;
                    running += c - 1;
                    T += c; X += e;
                    ++E;
                    if (running >= N) break;
                }
                s_E = E; s_T = T; s_X = X;
            }
            __syncthreads();
            const int E = s_E, T = s_T;
            for (int r = tid; r < E; r += 256) {
                const SortItem it = sort_buf[m - 1 - r];
                const int pos = it.ref;
                int push = it.count, nexp = it.ulx;
                for (int q = 0; q < 4; ++q) {
                    const int c = cc[pos * 4 + q];
                    if (c == 0) continue;
                    const int np = T - 1 - push++;
                    nxt[np] = child_of(cur[pos], q, c);
                    child_pos[pos * 4 + q] = (unsigned short)np;
                    if (c > 1) exp_list[nexp++] = (unsigned short)np;      // old entries were copied to sort_buf
                }
            }
            int stays = 0;
            for (int i = lo; i < hi; ++i) stays += expanded[i] ? 0 : 1;
            int d0 = 0, d1 = 0, S, D0, D1;
            block_scan3(stays, d0, d1, s_scan, S, D0, D1);
            int stay = stays;
            for (int i = lo; i < hi; ++i) {
                if (expanded[i]) continue;
                const int np = T + stay++;
                nxt[np] = cur[i];
                stay_pos[i] = (unsigned short)np;
            }
            if (tid == 0) {
                const int new_size = T + S;
                if (new_size >= N || new_size == size) s_state = 2;           // :753-757
                s_size = new_size;
                s_nexp = s_X;
            }
        }
        __syncthreads();
        for (int k0 = tid; k0 < n; k0 += 4 * 256) {
            int nd[4];
            uint32_t xy[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int k = k0 + u * 256;
                nd[u] = k < n ? nof[k] : -1;
                xy[u] = k < n ? key_xy[2 * k] : 0u;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (nd[u] < 0) continue;
                const int k = k0 + u * 256;
                if (expanded[nd[u]]) {
                    const ONode &node = cur[nd[u]];
                    nof[k] = child_pos[nd[u] * 4 + quadrant(node, (int)(xy[u] & 0xFFFF) - kBorderMin, (int)(xy[u] >> 16) - kBorderMin)];
                } else {
                    nof[k] = stay_pos[nd[u]];
                }
            }
        }
        { ONode *t = cur; cur = nxt; nxt = t; }
        __syncthreads();
    }

    // ---- best key per node (:766-782) ----
    const int size = s_size;
    for (int i = tid; i < size; i += 256) best[i] = 0ull;
    __syncthreads();
    for (int k = tid; k < n; k += 256) {
        const Cand c = keys[k];                                // independent iterations: the loads pipeline
        const int rx = c.x - kEdge, ry = c.y - kEdge;          // offset inside the FAST-able area
        const int cx = rx / L.w_cell, cy = ry / L.h_cell;
        const unsigned order = ((unsigned)(cy * L.n_cols + cx) << 14) | ((unsigned)(ry - cy * L.h_cell) << 7) |
                               (unsigned)(rx - cx * L.w_cell);
        const unsigned long long key = ((unsigned long long)c.score << 32) | (unsigned long long)(0xFFFFFFFFu - order);
        atomicMax(&best[nof[k]], key);
    }
    __syncthreads();
    for (int i = tid; i < size && i < L.kp_cap; i += 256) {
        const unsigned long long b = best[i];
        const unsigned order = 0xFFFFFFFFu - (unsigned)(b & 0xFFFFFFFFull);
        const int cell = order >> 14, ly = (order >> 7) & 127, lx = order & 127;
        const int cy = cell / L.n_cols, cx = cell - cy * L.n_cols;
        LevelKp kp;
        kp.x = (unsigned short)(kEdge + cx * L.w_cell + lx);
        kp.y = (unsigned short)(kEdge + cy * L.h_cell + ly);
        kp.score = (unsigned short)(b >> 32);
        kp.pad = 0;
        out[i] = kp;
    }
    if (tid == 0) *out_count = min(size, L.kp_cap);
}

void launch_octree(const FrameGeom &g, const Cand *cand, const int *cand_count, unsigned short *node_of,
                   LevelKp *level_kps, int *level_kp_count, int max_nodes, int nframes, cudaStream_t s) {
    const int cap = (max_nodes + 3) & ~3;
    // list_a, list_b (12 B each), cc (16 B), child_pos (8 B), stay_pos, expanded, exp_list (2 B each), sort_buf (12 B)
    const size_t smem = (size_t)cap * (12 + 12 + 16 + 8 + 2 + 2 + 2 + 12) + 64;
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        cudaFuncSetAttribute(octree_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = smem;
    }
    octree_kernel<<<dim3(g.nlevels, nframes), 256, smem, s>>>(g, cand, cand_count, node_of, level_kps, level_kp_count,
                                                             cap);
    count_launch();
}

}  // namespace vsg
