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
// describe_kernel: one warp per keypoint; lanes 0..30 are the 31 columns of the radius-15 disc for the
//                  moments, then lane i produces descriptor byte i (8 point pairs).
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

__global__ void __launch_bounds__(256) describe_kernel(FrameGeom g, const uint8_t *__restrict__ lvl0_base,
                                                       int lvl0_pitch, int64_t lvl0_stride,
                                                       const uint8_t *__restrict__ pyr, const uint8_t *__restrict__ blur,
                                                       const LevelKp *__restrict__ level_kps,
                                                       const int *__restrict__ level_kp_count,
                                                       const int *__restrict__ slot, vsg_keypoint *__restrict__ kps_out,
                                                       uint8_t *__restrict__ desc_out, int out_cap) {
    const int frame = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);  // keypoint index in level-major order
    int level = 0, off = 0;
    for (; level < g.nlevels; ++level) {
        const int c = level_kp_count[frame * g.nlevels + level];
        if (i < off + c) break;
        off += c;
    }
    if (level == g.nlevels) return;
    const LevelGeom &L = g.lv[level];
    const LevelKp kp = level_kps[(int64_t)frame * g.kp_total + L.kp_offset + (i - off)];
    const int out_row = slot[(int64_t)frame * g.kp_total + i];
    if (out_row < 0) return;

    const uint8_t *img;
    int ipitch;
    if (level == 0) { img = lvl0_base + (int64_t)frame * lvl0_stride; ipitch = lvl0_pitch; }
    else { img = pyr + L.plane_offset + (int64_t)frame * L.plane_stride; ipitch = L.pitch; }

    // ---- IC_Angle (:73-100): lanes = columns u = lane-15 of the disc ----
    int m10 = 0, m01 = 0;
    if (lane < 31) {
        const int u = lane - kHalfPatch;
        const int au = abs(u);
        const uint8_t *c = img + (int64_t)kp.y * ipitch + kp.x + u;
#pragma unroll 1
        for (int v = -kHalfPatch; v <= kHalfPatch; ++v) {
            if (au <= c_umax[abs(v)]) {
                const int val = __ldg(c + v * ipitch);
                m10 += u * val;
                m01 += v * val;
            }
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        m10 += __shfl_xor_sync(0xffffffffu, m10, d);
        m01 += __shfl_xor_sync(0xffffffffu, m01, d);
    }
    const float angle = fast_atan2_deg((float)m01, (float)m10);

    // ---- computeOrbDescriptor (:103-149) on the blurred level ----
    const float factor_pi = (float)(3.14159265358979323846 / 180.0);  // (float)(CV_PI/180.f), :102
    const float arad = __fmul_rn(angle, factor_pi);
    // glibc's cosf/sinf are (all but) correctly rounded; double-precision cos/sin rounded to float
    // reproduces that where CUDA's cosf (2 ulp) would not.  SURVEY A8 allows <=0.1 % differing bits.
    const float a = (float)cos((double)arad), b = (float)sin((double)arad);
    const uint8_t *bc = blur + L.plane_offset + (int64_t)frame * L.plane_stride + (int64_t)kp.y * L.pitch + kp.x;
    const char4 *pat = reinterpret_cast<const char4 *>(d_pattern) + lane * 8;
    int val = 0;
#pragma unroll
    for (int bit = 0; bit < 8; ++bit) {
        const char4 p = pat[bit];
        const float x0 = (float)p.x, y0 = (float)p.y, x1 = (float)p.z, y1 = (float)p.w;
        const int ry0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
        const int rx0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
        const int ry1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
        const int rx1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
        const int t0 = __ldg(bc + ry0 * L.pitch + rx0);
        const int t1 = __ldg(bc + ry1 * L.pitch + rx1);
        val |= (t0 < t1) << bit;
    }
    desc_out[((int64_t)frame * out_cap + out_row) * 32 + lane] = (uint8_t)val;

    if (lane == 0) {
        vsg_keypoint o;
        o.x = (float)kp.x;
        o.y = (float)kp.y;
        if (level != 0) {  // keypoint->pt *= scale (:1147-1150)
            o.x = __fmul_rn(o.x, L.scale);
            o.y = __fmul_rn(o.y, L.scale);
        }
        o.size = L.kp_size;
        o.angle = angle;
        o.response = (float)kp.score;
        o.octave = level;
        o.class_id = -1;
        kps_out[(int64_t)frame * out_cap + out_row] = o;
    }
}

void launch_describe(const FrameGeom &g, const uint8_t *lvl0_base, int lvl0_pitch, int64_t lvl0_stride,
                     const uint8_t *pyr, const uint8_t *blur, const LevelKp *level_kps, const int *level_kp_count,
                     int lap_x0, int lap_x1, vsg_keypoint *kps_out, uint8_t *desc_out, int out_cap, int *n_out,
                     int *mono_out, int *slot_scratch, int nframes, cudaStream_t s) {
    slot_kernel<<<nframes, 256, 0, s>>>(g, level_kps, level_kp_count, lap_x0, lap_x1, out_cap, n_out, mono_out,
                                       slot_scratch);
    describe_kernel<<<dim3((g.kp_total + 7) / 8, nframes), 256, 0, s>>>(g, lvl0_base, lvl0_pitch, lvl0_stride, pyr,
                                                                       blur, level_kps, level_kp_count, slot_scratch,
                                                                       kps_out, desc_out, out_cap);
    count_launch(2);
}

}  // namespace vsg
