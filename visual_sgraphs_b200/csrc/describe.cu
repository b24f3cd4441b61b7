// describe.cu — orientation, steered-BRIEF descriptors and the output ordering of operator().
//
// Reference (snt-arg/visual_sgraphs):
//   IC_Angle / computeOrientation        orb_slam3/src/ORBextractor.cc:73-100, :472-480   (un-blurred level)
//   computeOrbDescriptor / bit_pattern   :103-149, :151-409                                (blurred level)
//   ORBextractor::operator() tail        :1113-1168  (pt *= scale for level != 0; lapping-area partition:
//                                         x in [lap0, lap1] is written back-to-front, the rest front-to-back)
//   cv::fastAtan2                        SURVEY Appendix A4 (float32 polynomial, no FMA)
//
// slot_kernel:     one CTA per frame; prefix counts over the level-major keypoint order give every
//                  keypoint its output row, n and monoIndex.
// describe_kernel: 64 keypoints per CTA; a warp per keypoint for the moments (lanes 0..30 = the 31 columns of
//                  the radius-15 disc) and for the descriptor (lane i = byte i, 8 point pairs), one thread per
//                  keypoint for the scalar angle / cos / sin in between.
#include "vsg_internal.cuh"

namespace vsg {

__device__ const int8_t d_pattern[1024] = {
#include "orb_pattern.inc"
};
// umax[v]: half-width of the circular patch at row v (ORBextractor.cc:454-469 evaluates to this table)
__constant__ int c_umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};

__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * scale;
    const float p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale;
    const float p7 = -0.04432655554792128f * scale;
    const float eps = 2.2204460492503131e-16f;  // (float)DBL_EPSILON
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

__global__ void __launch_bounds__(256) slot_kernel(FrameGeom g, const LevelKp *__restrict__ level_kps,
                                                   const int *__restrict__ level_kp_count, int lap_x0, int lap_x1,
                                                   int out_cap, int *__restrict__ n_out, int *__restrict__ mono_out,
                                                   int *__restrict__ slot) {
    __shared__ int s_off[kMaxLevels + 1];
    __shared__ int s_part[256];
    pdl_launch_dependents();
    pdl_wait();
    const int frame = blockIdx.x, tid = threadIdx.x;
    if (tid == 0) {
        int acc = 0;
        for (int l = 0; l < g.nlevels; ++l) {
            s_off[l] = acc;
            acc += level_kp_count[frame * g.nlevels + l];
        }
        s_off[g.nlevels] = acc;
    }
    __syncthreads();
    const int total = s_off[g.nlevels];
    const int per = (total + 255) / 256;
    const int begin = min(tid * per, total), end = min(begin + per, total);
    const LevelKp *kps = level_kps + (int64_t)frame * g.kp_total;
    const float lo = (float)lap_x0, hi = (float)lap_x1;
    auto in_lap = [&](int i) {
        int l = 0;
        while (i >= s_off[l + 1]) ++l;
        const LevelKp k = kps[g.lv[l].kp_offset + (i - s_off[l])];
        float x = (float)k.x;
        if (l != 0) x = __fmul_rn(x, g.lv[l].scale);
        return x >= lo && x <= hi;
    };
    int mine = 0;
    for (int i = begin; i < end; ++i) mine += in_lap(i) ? 1 : 0;
    s_part[tid] = mine;
    __syncthreads();
    for (int d = 1; d < 256; d <<= 1) {  // inclusive Hillis-Steele scan
        const int v = tid >= d ? s_part[tid - d] : 0;
        __syncthreads();
        s_part[tid] += v;
        __syncthreads();
    }
    int lap_before = s_part[tid] - mine;
    const int lap_total = s_part[255];
    for (int i = begin; i < end; ++i) {
        const bool lap = in_lap(i);
        const int s = lap ? total - 1 - lap_before : i - lap_before;
        slot[(int64_t)frame * g.kp_total + i] = s < out_cap ? s : -1;
        lap_before += lap ? 1 : 0;
    }
    if (tid == 0) {
        n_out[frame] = total;
        mono_out[frame] = total - lap_total;
    }
}

// 64 keypoints per CTA (8 warps x 8 keypoints).  The warp-uniform scalar work — fastAtan2 and the
// double-precision cos/sin — is done once per keypoint by 64 threads in between the two warp-parallel
// phases instead of redundantly by all 32 lanes of a warp.
#ifndef VSG_DESC_MINB
#define VSG_DESC_MINB 4
#endif
// (few frames in flight: 16 keypoints per CTA — four times the CTAs, a quarter of the serial keypoints per warp)

template <int kDescKp>
__global__ void __launch_bounds__(256, VSG_DESC_MINB) describe_kernel(FrameGeom g, const uint8_t *__restrict__ lvl0_base,
                                                       int lvl0_pitch, int64_t lvl0_stride,
                                                       const uint8_t *__restrict__ pyr, const uint8_t *__restrict__ blur,
                                                       const LevelKp *__restrict__ level_kps,
                                                       const int *__restrict__ level_kp_count,
                                                       const int *__restrict__ slot, vsg_keypoint *__restrict__ kps_out,
                                                       uint8_t *__restrict__ desc_out, int out_cap, int *__restrict__ n_out,
                                                       int *__restrict__ mono_out) {
    __shared__ int s_x[kDescKp], s_y[kDescKp], s_level[kDescKp], s_row[kDescKp], s_score[kDescKp];
    __shared__ int s_m01[kDescKp], s_m10[kDescKp];
    __shared__ float s_angle[kDescKp], s_cos[kDescKp], s_sin[kDescKp];
    pdl_launch_dependents();
    pdl_wait();
    const int frame = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int first = blockIdx.x * kDescKp;             // keypoint index in level-major order

    if (tid < kDescKp) {
        const int i = first + tid;
        int level = 0, off = 0;
        for (; level < g.nlevels; ++level) {
            const int c = level_kp_count[frame * g.nlevels + level];
            if (i < off + c) break;
            off += c;
        }
        int row = -1;
        if (level < g.nlevels) {
            const LevelKp kp = level_kps[(int64_t)frame * g.kp_total + g.lv[level].kp_offset + (i - off)];
            // no lapping area (monocular / rectified stereo): output order = level-major order (:1113-1168), no slot kernel
            row = slot ? slot[(int64_t)frame * g.kp_total + i] : (i < out_cap ? i : -1);
            s_x[tid] = kp.x; s_y[tid] = kp.y; s_score[tid] = kp.score;
        }
        s_level[tid] = level;
        s_row[tid] = row;
        if (!slot && tid == 0 && blockIdx.x == 0) {          // ... and the counts the slot kernel would have written
            int total = 0;
            for (int l = 0; l < g.nlevels; ++l) total += level_kp_count[frame * g.nlevels + l];
            n_out[frame] = total;
            mono_out[frame] = total;
        }
    }
    __syncthreads();
    if (s_row[0] < 0 && s_level[0] >= g.nlevels) return;   // whole CTA is past the last keypoint

    // ---- IC_Angle moments (:73-100).  The disc is 31 rows of 31 bytes (u = -15..15); a row is eight 4-byte words once it is
    // shifted to start at u = -15.  Lane l owns word j = l % 8 of the rows l / 8 + 4 i (i = 0..7): a load instruction of the
    // warp reads four rows x 36 contiguous bytes.  Bytes outside the disc (|u| > umax[|v|]) are masked (table in shared
    // memory), and m10 += sum u * I, m01 += v * sum I are two DP4A per word — exact integer sums, any order.  Every plane's
    // pitch is a multiple of 4, so the misalignment of a row start is the same for all rows of a keypoint. ----
    __shared__ uint32_t s_umask[32 * 8];
    {
        const int row = tid >> 3, j = tid & 7;               // 256 threads = 32 rows x 8 words
        const int um = row < 31 ? c_umax[abs(row - kHalfPatch)] : -1;
        uint32_t mk = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int u = 4 * j + b - kHalfPatch;
            if (u <= kHalfPatch && abs(u) <= um) mk |= 0xFFu << (8 * b);
        }
        s_umask[tid] = mk;
    }
    __syncthreads();
    const int wj = lane & 7, wr = lane >> 3;
    const int u0 = 4 * wj - kHalfPatch;                      // weights u0 .. u0 + 3 as signed bytes
    const uint32_t wt = (uint32_t)(uint8_t)u0 | ((uint32_t)(uint8_t)(u0 + 1) << 8) | ((uint32_t)(uint8_t)(u0 + 2) << 16) |
                        ((uint32_t)(uint8_t)(u0 + 3) << 24);
    for (int k = warp; k < kDescKp; k += 8) {
        if (s_row[k] < 0) continue;
        const int level = s_level[k];
        const LevelGeom &L = g.lv[level];
        const uint8_t *img;
        int ipitch;
        if (level == 0) { img = lvl0_base + (int64_t)frame * lvl0_stride; ipitch = lvl0_pitch; }
        else { img = pyr + L.plane_offset + (int64_t)frame * L.plane_stride; ipitch = L.pitch; }
        // keypoints keep 16 pixels to every border (EDGE_THRESHOLD - 3): all 31 rows exist, and the up to four bytes a word
        // reaches beyond u = 15 still belong to the plane (or to the row below)
        const uint8_t *p0 = img + (int64_t)(s_y[k] - kHalfPatch + wr) * ipitch + (s_x[k] - kHalfPatch);
        const uint32_t off = (uint32_t)((uintptr_t)p0 & 3);
        const uint32_t *wp = reinterpret_cast<const uint32_t *>(p0 - off) + wj;
        const int step = ipitch;                             // four rows down, in words
        uint32_t lo[8], hi[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const bool ok = wr + 4 * i < 31;
            lo[i] = ok ? __ldg(wp + (int64_t)i * step) : 0u;
            hi[i] = ok && off ? __ldg(wp + (int64_t)i * step + 1) : 0u;
        }
        int m10 = 0, m01 = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint32_t px = __funnelshift_r(lo[i], hi[i], 8 * off) & s_umask[(wr + 4 * i) * 8 + wj];
            asm("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(m10) : "r"(px), "r"(wt));
            m01 += (wr + 4 * i - kHalfPatch) * (int)__dp4a(px, 0x01010101u, 0u);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            m10 += __shfl_xor_sync(0xffffffffu, m10, d);
            m01 += __shfl_xor_sync(0xffffffffu, m01, d);
        }
        if (lane == 0) { s_m10[k] = m10; s_m01[k] = m01; }
    }
    __syncthreads();

    // ---- one thread per keypoint: angle, steering coefficients, output record ----
    if (tid < kDescKp && s_row[tid] >= 0) {
        const float angle = fast_atan2_deg((float)s_m01[tid], (float)s_m10[tid]);
        const float factor_pi = (float)(3.14159265358979323846 / 180.0);  // (float)(CV_PI/180.f), :102
        const float arad = __fmul_rn(angle, factor_pi);
        // glibc's cosf/sinf are (all but) correctly rounded; double-precision cos/sin rounded to float
        // reproduces that where CUDA's cosf (2 ulp) would not.  SURVEY A8 allows <=0.1 % differing bits.
        s_cos[tid] = (float)cos((double)arad);
        s_sin[tid] = (float)sin((double)arad);
        s_angle[tid] = angle;
        const int level = s_level[tid];
        const LevelGeom &L = g.lv[level];
        vsg_keypoint o;
        o.x = (float)s_x[tid];
        o.y = (float)s_y[tid];
        if (level != 0) {  // keypoint->pt *= scale (:1147-1150)
            o.x = __fmul_rn(o.x, L.scale);
            o.y = __fmul_rn(o.y, L.scale);
        }
        o.size = L.kp_size;
        o.angle = angle;
        o.response = (float)s_score[tid];
        o.octave = level;
        o.class_id = -1;
        kps_out[(int64_t)frame * out_cap + s_row[tid]] = o;
    }
    __syncthreads();

    // ---- computeOrbDescriptor (:103-149) on the blurred level: lane i -> descriptor byte i ----
    // the lane's 16 pattern points as floats, converted once (they are the same for every keypoint of the CTA)
    const char4 *pat = reinterpret_cast<const char4 *>(d_pattern) + lane * 8;
    float4 p[8];
#pragma unroll
    for (int bit = 0; bit < 8; ++bit) {
        const char4 c = pat[bit];
        p[bit] = make_float4((float)c.x, (float)c.y, (float)c.z, (float)c.w);
    }
    for (int k = warp; k < kDescKp; k += 8) {
        const int out_row = s_row[k];
        if (out_row < 0) continue;
        const LevelGeom &L = g.lv[s_level[k]];
        const float a = s_cos[k], b = s_sin[k];
        const int bpitch = L.pitch;
        const uint8_t *bc = blur + L.plane_offset + (int64_t)frame * L.plane_stride + (int64_t)s_y[k] * L.pitch + s_x[k];
        int val = 0;
#pragma unroll
        for (int bit = 0; bit < 8; ++bit) {
            const float x0 = p[bit].x, y0 = p[bit].y, x1 = p[bit].z, y1 = p[bit].w;
            const int ry0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
            const int rx0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
            const int ry1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
            const int rx1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
            const int t0 = __ldg(bc + (ry0 * bpitch + rx0));       // one 32-bit offset, one widening add
            const int t1 = __ldg(bc + (ry1 * bpitch + rx1));
            val |= (t0 < t1) << bit;
        }
        desc_out[((int64_t)frame * out_cap + out_row) * 32 + lane] = (uint8_t)val;
    }
}

void launch_describe(const FrameGeom &g, const uint8_t *lvl0_base, int lvl0_pitch, int64_t lvl0_stride,
                     const uint8_t *pyr, const uint8_t *blur, const LevelKp *level_kps, const int *level_kp_count,
                     int lap_x0, int lap_x1, vsg_keypoint *kps_out, uint8_t *desc_out, int out_cap, int *n_out,
                     int *mono_out, int *slot_scratch, int nframes, cudaStream_t s) {
    // keypoints keep 16 pixels to the left border, so an empty lapping area [0, 0] holds none of them
    const bool lapping = !(lap_x0 == 0 && lap_x1 == 0);
    if (lapping)
        launch_kernel(slot_kernel, dim3(nframes), dim3(256), 0, s, true, g, level_kps, level_kp_count, lap_x0, lap_x1, out_cap,
                      n_out, mono_out, slot_scratch);
    const int *slot = lapping ? slot_scratch : nullptr;
    if (nframes <= 8)
        launch_kernel(describe_kernel<16>, dim3((g.kp_total + 15) / 16, nframes), dim3(256), 0, s, true, g, lvl0_base, lvl0_pitch,
                      lvl0_stride, pyr, blur, level_kps, level_kp_count, slot, kps_out, desc_out, out_cap, n_out, mono_out);
    else
        launch_kernel(describe_kernel<64>, dim3((g.kp_total + 63) / 64, nframes), dim3(256), 0, s, true, g, lvl0_base, lvl0_pitch,
                      lvl0_stride, pyr, blur, level_kps, level_kp_count, slot, kps_out, desc_out, out_cap, n_out, mono_out);
    count_launch(lapping ? 2 : 1);
}

}  // namespace vsg
