// N3: the reference's per-frame evaluation on the GPU, without a host synchronisation per frame.
//
// Replaces the torch op sequences of model/codd.py:462-474 (EPE / 3-px error of the frame's disparity) and :476-515
// (temporal EPE: the current frame's ground truth, prediction and validity mask are pulled back to the previous
// frame with the previous frame's ground-truth flow — utils/warp.py:69-92, grid_sample nearest / zeros /
// align_corners — and compared with the previous frame; plus the flow magnitude meter), with the validity masks of
// utils/misc.py:12-36.  Each call is ONE pass over the frame that adds into a row of float64 accumulators; the means
// (utils/metric.py) are taken by the host once per sequence from those rows.
//
// Arithmetic: every per-pixel quantity is computed in fp32 with the reference's operation order (explicit
// __f*_rn intrinsics: no FMA contraction), so the masks, the nearest-neighbour indices and the thresholded counts
// are exact; only the sums differ from torch.mean (float64 atomics here, fp32 tree sums there).
#include "common.cuh"

namespace {

constexpr int MT_THREADS = 256;
constexpr float MT_BF = 1050 * 0.2f;   // utils/misc.py:7

// block-wide sum of NV doubles per thread -> atomicAdd into acc[0..NV)
template <int NV>
__device__ __forceinline__ void block_accumulate(double (&v)[NV], double* __restrict__ acc) {
    __shared__ double s_part[MT_THREADS / 32][NV];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double x = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
        if (lane == 0) s_part[warp][k] = x;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double x = 0.0;
#pragma unroll
        for (int wv = 0; wv < MT_THREADS / 32; ++wv) x += s_part[wv][threadIdx.x];
        if (x != 0.0) atomicAdd(acc + threadIdx.x, x);
    }
}

__device__ __forceinline__ bool in_range(float d, float lo, float hi) { return d > lo && d < hi; }

// acc[0] += #valid, acc[1] += sum |pred - gt|, acc[2] += #(|pred - gt| > 3), acc[3] += #(gt > 0); mask_out = validity
// (codd.py:462-474; acc[3] is the device-side form of `torch.any(gt_disp > 0.0)`, codd.py:482)
__global__ void __launch_bounds__(MT_THREADS) disp_metrics_kernel(const float* __restrict__ pred, size_t pred_ss,
                                                                  int pred_rs, const float* __restrict__ gt,
                                                                  const float* __restrict__ seg, int h, int w,
                                                                  size_t total, float lo, float hi,
                                                                  unsigned char* __restrict__ mask_out,
                                                                  double* __restrict__ acc) {
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    for (size_t i = (size_t)blockIdx.x * MT_THREADS + threadIdx.x; i < total; i += (size_t)gridDim.x * MT_THREADS) {
        const int x = (int)(i % w);
        const size_t t = i / w;
        const int y = (int)(t % h);
        const size_t s = t / h;
        const float g = __ldg(gt + i);
        v[3] += g > 0.f ? 1.0 : 0.0;
        bool m = in_range(g, lo, hi);
        if (seg) m = m && (__ldg(seg + i) > 0.f);
        if (mask_out) mask_out[i] = m ? 1 : 0;
        if (m) {
            const float e = fabsf(__fsub_rn(__ldg(pred + s * pred_ss + (size_t)y * pred_rs + x), g));
            v[0] += 1.0;
            v[1] += (double)e;
            v[2] += e > 3.0f ? 1.0 : 0.0;
        }
    }
    block_accumulate<4>(v, acc);
}

// grid + flow -> normalize_coords (warp.py:14-15) -> grid_sample unnormalise, align_corners -> nearbyint
__device__ __forceinline__ float sample_index(float base, float flow, int size) {
    const float sm1 = (float)(size - 1);
    const float s = __fadd_rn(base, flow);
    const float n = __fsub_rn(__fmul_rn(2.f, __fdiv_rn(s, sm1)), 1.f);
    const float u = __fmul_rn(__fdiv_rn(__fadd_rn(n, 1.f), 2.f), sm1);
    return nearbyintf(u);   // round half to even, as std::nearbyint in grid_sample
}

// codd.py:476-515.  acc[0] += #(mask_prev & mask_curr), [1] sum abs_err, [2] sum rel_err, [3] #(rel > 1),
// [4] #(abs > 3), [5] #mask_prev, [6] #mask_curr, [7] sum |flow|, [8] #pixels
__global__ void __launch_bounds__(MT_THREADS) temporal_metrics_kernel(
    const float* __restrict__ flow, const float* __restrict__ gt, const float* __restrict__ pred, size_t pred_ss,
    int pred_rs, const float* __restrict__ seg, const float* __restrict__ gt_prev, const float* __restrict__ pred_prev,
    size_t pprev_ss, int pprev_rs, const unsigned char* __restrict__ mask_prev, const float* __restrict__ gt_disp2_prev,
    const double* __restrict__ gt_pos_count, int h, int w, size_t total, float lo, float hi, double* __restrict__ acc) {
    double v[9] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    const size_t hw = (size_t)h * w;
    const float dummy = __fdiv_rn(MT_BF, 2.f);
    // KITTI provides disparity for one frame only: with no positive ground truth the mask is built from a dummy
    // disparity of BF/2 (codd.py:482-490); the count comes from codd_disp_metrics of the same frame
    const bool dummy_gt = gt_pos_count && (*gt_pos_count == 0.0);
    // validity of the CURRENT frame's pixel q under the flow stored at q (misc.py:26-31)
    auto cur_mask = [&](size_t s, size_t q) {
        const float g = dummy_gt ? dummy : __ldg(gt + s * hw + q);
        bool m = in_range(g, lo, hi);
        if (seg) m = m && (__ldg(seg + s * hw + q) > 0.f);
        const float fx = __ldg(flow + (s * 2) * hw + q), fy = __ldg(flow + (s * 2 + 1) * hw + q);
        const float mag = __fsqrt_rn(__fadd_rn(__fmul_rn(fx, fx), __fmul_rn(fy, fy)));
        return m && (mag < MT_BF);
    };
    for (size_t i = (size_t)blockIdx.x * MT_THREADS + threadIdx.x; i < total; i += (size_t)gridDim.x * MT_THREADS) {
        const size_t s = i / hw, q = i - s * hw;
        const int y = (int)(q / w), x = (int)(q - (size_t)y * w);
        const float fx = __ldg(flow + (s * 2) * hw + q), fy = __ldg(flow + (s * 2 + 1) * hw + q);
        v[7] += (double)__fsqrt_rn(__fadd_rn(__fmul_rn(fx, fx), __fmul_rn(fy, fy)));
        v[8] += 1.0;
        const bool mp = mask_prev[i] != 0;
        v[5] += mp ? 1.0 : 0.0;
        const float ix = sample_index((float)x, fx, w), iy = sample_index((float)y, fy, h);
        const bool inb = ix >= 0.f && ix <= (float)(w - 1) && iy >= 0.f && iy <= (float)(h - 1);
        if (!inb) continue;                                   // valid = 0 -> mask_curr = 0
        const int xs = (int)ix, ys = (int)iy;
        const size_t qs = (size_t)ys * w + xs;
        bool mc = cur_mask(s, qs) && cur_mask(s, q);          // warped mask & the unwarped one (codd.py:499)
        float w_gt = __ldg(gt + s * hw + qs);
        if (gt_disp2_prev) {                                  // dataset provides disp2 (codd.py:500-502)
            w_gt = __ldg(gt_disp2_prev + i);
            mc = mc && (w_gt > 0.f);
        }
        v[6] += mc ? 1.0 : 0.0;
        if (!(mc && mp)) continue;
        const float w_pred = __ldg(pred + s * pred_ss + (size_t)ys * pred_rs + xs);
        const float d_est = __fsub_rn(w_pred, __ldg(pred_prev + s * pprev_ss + (size_t)y * pprev_rs + x));
        const float d_gt = __fsub_rn(w_gt, __ldg(gt_prev + i));
        const float ae = fabsf(__fsub_rn(d_est, d_gt));
        const float re = __fdiv_rn(ae, __fadd_rn(fabsf(d_gt), 1e-3f));
        v[0] += 1.0;
        v[1] += (double)ae;
        v[2] += (double)re;
        v[3] += re > 1.0f ? 1.0 : 0.0;
        v[4] += ae > 3.0f ? 1.0 : 0.0;
    }
    block_accumulate<9>(v, acc);
}

// codd.py:519-575.  acc[0] += #valid, [1] += sum scene-flow EPE, [2] += sum optical-flow EPE, [3] += #(sf < 1), [4] += #(of < 1)
// Ts: dense SE3 field (tx, ty, tz, qx, qy, qz, qw) with pixel strides (a [:h,:w] crop of the padded field is fine);
// induced_flow of model/motion/raft3d/projective_ops.py:11-68 at depth = clip(BF / pred_prev, 0, BF).
__global__ void __launch_bounds__(MT_THREADS) sceneflow_metrics_kernel(
    const float* __restrict__ Ts, size_t ts_ss, size_t ts_rs, const float* __restrict__ pred_prev, size_t pp_ss, int pp_rs,
    const float* __restrict__ intr, const float* __restrict__ flow, const float* __restrict__ disp_change,
    const float* __restrict__ gt_prev, const float* __restrict__ seg, const unsigned char* __restrict__ flow_occ, int h,
    int w, size_t total, float lo, float hi, double* __restrict__ acc) {
    double v[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    const size_t hw = (size_t)h * w;
    for (size_t i = (size_t)blockIdx.x * MT_THREADS + threadIdx.x; i < total; i += (size_t)gridDim.x * MT_THREADS) {
        const size_t s = i / hw, q = i - s * hw;
        const int y = (int)(q / w), x = (int)(q - (size_t)y * w);
        const float fxg = __ldg(flow + (s * 2) * hw + q), fyg = __ldg(flow + (s * 2 + 1) * hw + q);
        const float dc = __ldg(disp_change + i);
        bool m = in_range(__ldg(gt_prev + i), lo, hi);
        if (seg) m = m && (__ldg(seg + i) > 0.f);
        m = m && (__fsqrt_rn(__fadd_rn(__fmul_rn(fxg, fxg), __fmul_rn(fyg, fyg))) < MT_BF) && (fabsf(dc) < MT_BF);
        if (flow_occ) m = m && (flow_occ[i] == 0);
        if (!m) continue;
        const float fx = __ldg(intr + s * 4), fy = __ldg(intr + s * 4 + 1), cx = __ldg(intr + s * 4 + 2), cy = __ldg(intr + s * 4 + 3);
        const float d = fminf(fmaxf(__fdiv_rn(MT_BF, __ldg(pred_prev + s * pp_ss + (size_t)y * pp_rs + x)), 0.f), MT_BF);
        const float X0x = __fmul_rn(d, __fdiv_rn(__fsub_rn((float)x, cx), fx));
        const float X0y = __fmul_rn(d, __fdiv_rn(__fsub_rn((float)y, cy), fy));
        const float X0z = d;
        const float* tp = Ts + s * ts_ss + (size_t)y * ts_rs + (size_t)x * 7;
        const float tx = __ldg(tp), ty = __ldg(tp + 1), tz = __ldg(tp + 2);
        const float qx = __ldg(tp + 3), qy = __ldg(tp + 4), qz = __ldg(tp + 5), qw = __ldg(tp + 6);
        // X1 = X0 + qw * uv + qv x uv + t,  uv = 2 (qv x X0)
        const float ux = 2.f * (qy * X0z - qz * X0y), uy = 2.f * (qz * X0x - qx * X0z), uz = 2.f * (qx * X0y - qy * X0x);
        const float X1x = X0x + qw * ux + (qy * uz - qz * uy) + tx;
        const float X1y = X0y + qw * uy + (qz * ux - qx * uz) + ty;
        const float X1z = X0z + qw * uz + (qx * uy - qy * ux) + tz;
        const float Z0 = X0z + 1e-5f, Z1 = X1z + 1e-5f;
        const float ex = (fx * (X1x / Z1) + cx) - (fx * (X0x / Z0) + cx);
        const float ey = (fy * (X1y / Z1) + cy) - (fy * (X0y / Z0) + cy);
        const float ez = (1.f / Z1 - 1.f / Z0) * MT_BF;
        const float dx = ex - fxg, dy = ey - fyg, dz = ez - dc;
        const float of2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        const float sf = __fsqrt_rn(__fadd_rn(of2, __fmul_rn(dz, dz))), of = __fsqrt_rn(of2);
        v[0] += 1.0;
        v[1] += (double)sf;
        v[2] += (double)of;
        v[3] += sf < 1.0f ? 1.0 : 0.0;
        v[4] += of < 1.0f ? 1.0 : 0.0;
    }
    block_accumulate<5>(v, acc);
}

// utils/misc.py:39-59 compute_gt_disp_change: change = flow_warp(gt_curr, flow, nearest, zeros) - gt_prev, BF where the
// sample falls outside the image or the previous frame's pixel is flow-occluded; warped_out (optional) = the warped map.
__global__ void __launch_bounds__(MT_THREADS) disp_change_kernel(const float* __restrict__ flow, const float* __restrict__ gt_curr,
                                                                 const float* __restrict__ gt_prev,
                                                                 const unsigned char* __restrict__ occ_prev, int h, int w,
                                                                 size_t total, float* __restrict__ change,
                                                                 float* __restrict__ warped_out) {
    const size_t hw = (size_t)h * w;
    for (size_t i = (size_t)blockIdx.x * MT_THREADS + threadIdx.x; i < total; i += (size_t)gridDim.x * MT_THREADS) {
        const size_t s = i / hw, q = i - s * hw;
        const int y = (int)(q / w), x = (int)(q - (size_t)y * w);
        const float ix = sample_index((float)x, __ldg(flow + (s * 2) * hw + q), w);
        const float iy = sample_index((float)y, __ldg(flow + (s * 2 + 1) * hw + q), h);
        const bool inb = ix >= 0.f && ix <= (float)(w - 1) && iy >= 0.f && iy <= (float)(h - 1);
        const float wv = inb ? __ldg(gt_curr + s * hw + (size_t)((int)iy) * w + (int)ix) : 0.f;
        float c = __fsub_rn(wv, __ldg(gt_prev + i));
        if (!inb || (occ_prev && occ_prev[i] != 0)) c = MT_BF;
        change[i] = c;
        if (warped_out) warped_out[i] = wv;
    }
}

int metrics_grid(size_t total) {
    const size_t blocks = (total + MT_THREADS - 1) / MT_THREADS;
    return (int)(blocks < 148 * 8 ? (blocks ? blocks : 1) : 148 * 8);
}

}  // namespace

extern "C" int codd_disp_metrics(const float* pred, long long pred_sample_stride, int pred_row_stride, const float* gt,
                                 const float* seg, int n, int h, int w, float disp_lo, float disp_hi,
                                 unsigned char* mask_out, double* acc, void* stream) {
    if (!pred || !gt || !acc || n <= 0 || h <= 0 || w <= 0) return CODD_E_BADARG;
    if (pred_row_stride < w || pred_sample_stride < (long long)h * pred_row_stride) return CODD_E_SHAPE;
    const size_t total = (size_t)n * h * w;
    disp_metrics_kernel<<<metrics_grid(total), MT_THREADS, 0, (cudaStream_t)stream>>>(
        pred, (size_t)pred_sample_stride, pred_row_stride, gt, seg, h, w, total, disp_lo, disp_hi, mask_out, acc);
    CODD_RETURN_IF_CUDA_ERROR();
    return 0;
}

extern "C" int codd_temporal_metrics(const float* flow_prev, const float* gt, const float* pred,
                                     long long pred_sample_stride, int pred_row_stride, const float* seg,
                                     const float* gt_prev, const float* pred_prev, long long pprev_sample_stride,
                                     int pprev_row_stride, const unsigned char* mask_prev, const float* gt_disp2_prev,
                                     const double* gt_pos_count, int n, int h, int w, float disp_lo, float disp_hi, double* acc,
                                     void* stream) {
    if (!flow_prev || !gt || !pred || !gt_prev || !pred_prev || !mask_prev || !acc || n <= 0 || h <= 1 || w <= 1)
        return CODD_E_BADARG;
    if (pred_row_stride < w || pprev_row_stride < w || pred_sample_stride < (long long)h * pred_row_stride ||
        pprev_sample_stride < (long long)h * pprev_row_stride)
        return CODD_E_SHAPE;
    const size_t total = (size_t)n * h * w;
    temporal_metrics_kernel<<<metrics_grid(total), MT_THREADS, 0, (cudaStream_t)stream>>>(
        flow_prev, gt, pred, (size_t)pred_sample_stride, pred_row_stride, seg, gt_prev, pred_prev,
        (size_t)pprev_sample_stride, pprev_row_stride, mask_prev, gt_disp2_prev, gt_pos_count, h, w, total, disp_lo, disp_hi,
        acc);
    CODD_RETURN_IF_CUDA_ERROR();
    return 0;
}

extern "C" int codd_sceneflow_metrics(const float* Ts, long long ts_sample_stride, long long ts_row_stride,
                                      const float* pred_prev, long long pprev_sample_stride, int pprev_row_stride,
                                      const float* intrinsics, const float* flow_prev, const float* gt_disp_change,
                                      const float* gt_prev, const float* seg, const unsigned char* flow_occ, int n, int h,
                                      int w, float disp_lo, float disp_hi, double* acc, void* stream) {
    if (!Ts || !pred_prev || !intrinsics || !flow_prev || !gt_disp_change || !gt_prev || !acc || n <= 0 || h <= 0 || w <= 0)
        return CODD_E_BADARG;
    if (ts_row_stride < (long long)w * 7 || ts_sample_stride < (long long)h * ts_row_stride || pprev_row_stride < w ||
        pprev_sample_stride < (long long)h * pprev_row_stride)
        return CODD_E_SHAPE;
    const size_t total = (size_t)n * h * w;
    sceneflow_metrics_kernel<<<metrics_grid(total), MT_THREADS, 0, (cudaStream_t)stream>>>(
        Ts, (size_t)ts_sample_stride, (size_t)ts_row_stride, pred_prev, (size_t)pprev_sample_stride, pprev_row_stride,
        intrinsics, flow_prev, gt_disp_change, gt_prev, seg, flow_occ, h, w, total, disp_lo, disp_hi, acc);
    CODD_RETURN_IF_CUDA_ERROR();
    return 0;
}

extern "C" int codd_gt_disp_change(const float* flow_prev, const float* gt_curr, const float* gt_prev,
                                   const unsigned char* flow_occ_prev, int n, int h, int w, float* change, float* warped,
                                   void* stream) {
    if (!flow_prev || !gt_curr || !gt_prev || !change || n <= 0 || h <= 1 || w <= 1) return CODD_E_BADARG;
    const size_t total = (size_t)n * h * w;
    disp_change_kernel<<<metrics_grid(total), MT_THREADS, 0, (cudaStream_t)stream>>>(flow_prev, gt_curr, gt_prev, flow_occ_prev,
                                                                                    h, w, total, change, warped);
    CODD_RETURN_IF_CUDA_ERROR();
    return 0;
}
