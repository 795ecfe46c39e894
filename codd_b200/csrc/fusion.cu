// K13 — Fusion input cues and blend (model/fusion/fusion.py:168-318, 383-394; utils/warp.py:43-66).
//
//   fusion_cues_lowres   1/4 resolution: pixel-to-patch feature correlations (3x3, dilation 2, zero
//                        pad; cross [curr x warp] 9 + self 8 + 8, / sqrt(C)) and the +-1 local stereo
//                        costs of the current and the warped disparity (6 disparity warps of the right
//                        feature in the reference, fusion.py:200-241) -> corr_feat [N,31,h,w]
//   fusion_forget_in     full resolution: |disparity - patch| cues (9 + 8 + 8), warped flow (3),
//                        validity (1), warped confidence (3) -> 32 cues, folded straight into the first
//                        1x1 conv of forget_head (32 -> 16): the 32-channel full-resolution cue tensor
//                        of the reference never exists
//   fusion_blend         reset weight = sigmoid(1x1 conv 8 -> 1), fusion weight nearest x4, validity
//                        masks, disp = curr*(1 - wf*wr) + warp*wf*wr
//
// The disparity warps reuse K4's sampling arithmetic (warp_sample.cuh): bit-identical local costs.
#include "common.cuh"
#include "warp_sample.cuh"

namespace {

constexpr int FC = 32;   // fusion channels
constexpr float INV_SQRT_DIV = 5.656854249492381f;   // (float)sqrt(32): the reference divides by it

struct CuesP {
    const float* feat_curr; int ldc;   // [N,h,w,32] NHWC
    const float* feat_warp; int ldw;
    const float* fea_l; int ldl;       // [N,h,w,cs] NHWC stereo feature (left)
    const float* fea_r;                // [N,cs,h,w] PLANAR stereo feature (right)
    int cs;
    const float* pred_curr;            // [N,H,W]
    const float* pred_warp;
    int N, h, w, ds;
    float* corr; int ldo;              // [N,h,w,>=31]
    float* disp2; int ld2;             // [N,h,w,2] sub-sampled (curr, warp)
    float* extra; int ldx;             // optional second copy of the two disparities (inp64[..., 62:64])
};

__global__ void __launch_bounds__(128) fusion_cues_lowres_kernel(CuesP p) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y, n = blockIdx.z;
    if (x >= p.w) return;
    const int H = p.h * p.ds, W = p.w * p.ds, o = p.ds / 2 - 1;
    const size_t pix = ((size_t)n * p.h + y) * p.w + x;
    float out[32];

    // ---- feature correlations
    float qc[FC], qw[FC];
    {
        const float* a = p.feat_curr + pix * p.ldc;
        const float* b = p.feat_warp + pix * p.ldw;
#pragma unroll
        for (int c = 0; c < FC; c += 4) {
            const float4 u = ldg4(a + c), v = ldg4(b + c);
            qc[c] = u.x; qc[c + 1] = u.y; qc[c + 2] = u.z; qc[c + 3] = u.w;
            qw[c] = v.x; qw[c + 1] = v.y; qw[c + 2] = v.z; qw[c + 3] = v.w;
        }
    }
    int self_i = 0;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        const int yy = y + 2 * (t / 3 - 1), xx = x + 2 * (t % 3 - 1);
        const bool in = yy >= 0 && yy < p.h && xx >= 0 && xx < p.w;
        float cross = 0.f, sc = 0.f, sw = 0.f;
        if (in) {
            const size_t np = ((size_t)n * p.h + yy) * p.w + xx;
            const float* a = p.feat_curr + np * p.ldc;
            const float* b = p.feat_warp + np * p.ldw;
#pragma unroll
            for (int c = 0; c < FC; c += 4) {
                const float4 u = ldg4(a + c), v = ldg4(b + c);
                cross = fmaf(qc[c], v.x, cross); cross = fmaf(qc[c + 1], v.y, cross);
                cross = fmaf(qc[c + 2], v.z, cross); cross = fmaf(qc[c + 3], v.w, cross);
                sw = fmaf(qw[c], v.x, sw); sw = fmaf(qw[c + 1], v.y, sw);
                sw = fmaf(qw[c + 2], v.z, sw); sw = fmaf(qw[c + 3], v.w, sw);
                sc = fmaf(qc[c], u.x, sc); sc = fmaf(qc[c + 1], u.y, sc);
                sc = fmaf(qc[c + 2], u.z, sc); sc = fmaf(qc[c + 3], u.w, sc);
            }
        }
        out[t] = __fdiv_rn(cross, INV_SQRT_DIV);
        if (t != 4) {
            out[9 + self_i] = __fdiv_rn(sc, INV_SQRT_DIV);
            out[17 + self_i] = __fdiv_rn(sw, INV_SQRT_DIV);
            ++self_i;
        }
    }

    // ---- local stereo costs of (pred_curr, pred_warp) / ds + k
    const float pc = __ldg(p.pred_curr + ((size_t)n * H + (size_t)y * p.ds + o) * W + (size_t)x * p.ds + o);
    const float pw = __ldg(p.pred_warp + ((size_t)n * H + (size_t)y * p.ds + o) * W + (size_t)x * p.ds + o);
    {
        const float wm1 = (float)(p.w - 1), hm1 = (float)(p.h - 1);
        const float wdiv = (float)max(p.w - 1, 1), hdiv = (float)max(p.h - 1, 1);
        const float wrcp = __frcp_rn(wdiv);
        const float gy = __fadd_rn(__fdiv_rn(__fmul_rn(2.f, (float)y), hdiv), -1.f);
        const float iy = __fmul_rn(__fmul_rn(__fadd_rn(gy, 1.f), 0.5f), hm1);
        const float fy = floorf(iy);
        const float fn = __fsub_rn(iy, fy), fs = __fsub_rn(1.f, fn);
        const int y0 = min(max((int)fy, 0), p.h - 1);
        const bool two_rows = fn != 0.f;
        const bool row1_ok = (y0 + 1 < p.h);
        const int rowstep = row1_ok ? p.w : 0;
        const float fds = (float)p.ds;
        Taps tp[2];
        sample_setup(__fdiv_rn(pc, fds), 0.f, 0.f, 0.f, 0.f, x, wm1, wdiv, wrcp, tp[0]);
        sample_setup(__fdiv_rn(pw, fds), 0.f, 0.f, 0.f, 0.f, x, wm1, wdiv, wrcp, tp[1]);
        float cost[2][3], wA[2][3], wB[2][3], wC[2][3], wD[2][3];
        int offA[2][3], offB[2][3];
        const int rowoff = y0 * p.w;
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int ki = 0; ki < 3; ++ki) {
                cost[s][ki] = 0.f;
                const int xa = tp[s].x0[ki];
                const bool va = (xa >= 0 && xa < p.w), vb = (xa + 1 >= 0 && xa + 1 < p.w);
                offA[s][ki] = rowoff + min(max(xa, 0), p.w - 1);
                offB[s][ki] = rowoff + min(max(xa + 1, 0), p.w - 1);
                wA[s][ki] = va ? __fmul_rn(fs, tp[s].fe[ki]) : 0.f;
                wB[s][ki] = vb ? __fmul_rn(fs, tp[s].fw[ki]) : 0.f;
                wC[s][ki] = (va && row1_ok) ? __fmul_rn(fn, tp[s].fe[ki]) : 0.f;
                wD[s][ki] = (vb && row1_ok) ? __fmul_rn(fn, tp[s].fw[ki]) : 0.f;
            }
        float lnorm = 0.f;
        const size_t cstride = (size_t)p.h * p.w;
        const float* flp = p.fea_l + pix * p.ldl;
        const float* frn = p.fea_r + (size_t)n * p.cs * cstride;
        if (two_rows) k4_channels<2, true, false>(flp, frn, cstride, p.cs, rowstep, offA, offB, wA, wB, wC, wD, cost, lnorm);
        else k4_channels<2, false, false>(flp, frn, cstride, p.cs, rowstep, offA, offB, wA, wB, wC, wD, cost, lnorm);
        const float cdiv = (float)p.cs / 24.0f;   // cost / (in_channels / 24)
#pragma unroll
        for (int ki = 0; ki < 3; ++ki) {
            out[25 + ki] = __fdiv_rn(cost[0][ki], cdiv);
            out[28 + ki] = __fdiv_rn(cost[1][ki], cdiv);
        }
    }
    out[31] = 0.f;
    float4* op = reinterpret_cast<float4*>(p.corr + pix * p.ldo);
#pragma unroll
    for (int c = 0; c < 8; ++c) op[c] = make_float4(out[4 * c], out[4 * c + 1], out[4 * c + 2], out[4 * c + 3]);
    p.disp2[pix * p.ld2] = pc;
    p.disp2[pix * p.ld2 + 1] = pw;
    if (p.extra) {
        p.extra[pix * p.ldx] = pc;
        p.extra[pix * p.ldx + 1] = pw;
    }
}

// ---------------------------------------------------------------------------------------------
struct ForgetP {
    const float* pred_curr;   // [N,H,W]
    const float* pred_warp;
    const float* flow;        // [N,3,H,W] planar
    const float* conf;        // [N,3,H,W] planar
    const float* w;           // torch [16][32]
    const float* b;
    int N, H, W;
    float* out; int ldo;      // [N,H,W,16] NHWC
    float* cues;              // optional [N,32,H,W] planar debug/test output
};

__device__ __forceinline__ float pad_ld(const float* img, int y, int x, int H, int W) {
    return (y >= 0 && y < H && x >= 0 && x < W) ? __ldg(img + (size_t)y * W + x) : 0.f;
}

__global__ void __launch_bounds__(128) fusion_forget_in_kernel(ForgetP p) {
    __shared__ __align__(16) float s_w[32][16];
    __shared__ float s_b[16];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) s_w[i & 31][i >> 5] = __ldg(p.w + i);   // [out][in] -> [in][out]
    if (threadIdx.x < 16) s_b[threadIdx.x] = __ldg(p.b + threadIdx.x);
    __syncthreads();
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y, n = blockIdx.z;
    if (x >= p.W) return;
    const size_t hw = (size_t)p.H * p.W;
    const float* pcur = p.pred_curr + (size_t)n * hw;
    const float* pwar = p.pred_warp + (size_t)n * hw;
    const float c0 = __ldg(pcur + (size_t)y * p.W + x), w0 = __ldg(pwar + (size_t)y * p.W + x);
    float cue[32];
    int si = 0;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        const int yy = y + 2 * (t / 3 - 1), xx = x + 2 * (t % 3 - 1);
        const float wn = pad_ld(pwar, yy, xx, p.H, p.W);
        cue[t] = fabsf(__fsub_rn(c0, wn));                       // |curr - warp patch|   (C == 1: difference, / sqrt(1))
        if (t != 4) {
            cue[9 + si] = fabsf(__fsub_rn(c0, pad_ld(pcur, yy, xx, p.H, p.W)));
            cue[17 + si] = fabsf(__fsub_rn(w0, wn));
            ++si;
        }
    }
    const size_t pp = (size_t)y * p.W + x;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        cue[25 + c] = __ldg(p.flow + ((size_t)n * 3 + c) * hw + pp);
        cue[29 + c] = __ldg(p.conf + ((size_t)n * 3 + c) * hw + pp);
    }
    cue[28] = w0 > 0.f ? 1.f : 0.f;
    if (p.cues)
        for (int c = 0; c < 32; ++c) p.cues[((size_t)n * 32 + c) * hw + pp] = cue[c];
    float acc[16];
#pragma unroll
    for (int o = 0; o < 16; ++o) acc[o] = s_b[o];
#pragma unroll
    for (int c = 0; c < 32; ++c) {
#pragma unroll
        for (int o4 = 0; o4 < 4; ++o4) {
            const float4 wv = *reinterpret_cast<const float4*>(&s_w[c][o4 * 4]);
            acc[o4 * 4 + 0] = fmaf(cue[c], wv.x, acc[o4 * 4 + 0]);
            acc[o4 * 4 + 1] = fmaf(cue[c], wv.y, acc[o4 * 4 + 1]);
            acc[o4 * 4 + 2] = fmaf(cue[c], wv.z, acc[o4 * 4 + 2]);
            acc[o4 * 4 + 3] = fmaf(cue[c], wv.w, acc[o4 * 4 + 3]);
        }
    }
    float4* op = reinterpret_cast<float4*>(p.out + ((size_t)n * hw + pp) * p.ldo);
#pragma unroll
    for (int o4 = 0; o4 < 4; ++o4) op[o4] = make_float4(acc[o4 * 4], acc[o4 * 4 + 1], acc[o4 * 4 + 2], acc[o4 * 4 + 3]);
}

// ---------------------------------------------------------------------------------------------
struct BlendP {
    const float* pred_curr;   // [N,H,W]
    const float* pred_warp;
    const float* r8; int ldr; // [N,H,W,8] NHWC: forget_head.1 output
    const float* w;           // [8]  forget_head.2 weight
    const float* b;           // [1]
    const float* wf_lr;       // [N,h,w] fusion weight at 1/ds resolution (after sigmoid)
    int N, H, W, ds;
    float* fused;             // [N,H,W]
    float* wf;                // [N,H,W]
    float* wr;                // [N,H,W]
};

__global__ void __launch_bounds__(256) fusion_blend_kernel(BlendP p) {
    const size_t total = (size_t)p.N * p.H * p.W;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int x = (int)(i % p.W);
    const size_t t = i / p.W;
    const int y = (int)(t % p.H);
    const int n = (int)(t / p.H);
    const float4 a = ldg4(p.r8 + i * p.ldr), c = ldg4(p.r8 + i * p.ldr + 4);
    float z = __ldg(p.b);
    z = fmaf(a.x, __ldg(p.w + 0), z); z = fmaf(a.y, __ldg(p.w + 1), z);
    z = fmaf(a.z, __ldg(p.w + 2), z); z = fmaf(a.w, __ldg(p.w + 3), z);
    z = fmaf(c.x, __ldg(p.w + 4), z); z = fmaf(c.y, __ldg(p.w + 5), z);
    z = fmaf(c.z, __ldg(p.w + 6), z); z = fmaf(c.w, __ldg(p.w + 7), z);
    const float pc = __ldg(p.pred_curr + i), pw = __ldg(p.pred_warp + i);
    const float mask = pw > 0.f ? 1.f : 0.f;
    const float wr = __fmul_rn(1.f / (1.f + expf(-z)), mask);
    const int hl = p.H / p.ds, wl = p.W / p.ds;
    const float wf = __fmul_rn(__ldg(p.wf_lr + ((size_t)n * hl + y / p.ds) * wl + x / p.ds), mask);
    // pred_curr * (1 - wf*wr) + pred_warp * wf * wr, every product / sum rounded separately
    const float g = __fmul_rn(wf, wr);
    const float fused = __fadd_rn(__fmul_rn(pc, __fsub_rn(1.f, g)), __fmul_rn(__fmul_rn(pw, wf), wr));
    p.fused[i] = fused;
    p.wf[i] = wf;
    p.wr[i] = wr;
}

}  // namespace

extern "C" int codd_fusion_cues_lowres(const float* feat_curr, int ldc, const float* feat_warp, int ldw,
                                       const float* fea_l, int ldl, const float* fea_r_planar, int cs,
                                       const float* pred_curr, const float* pred_warp, int n, int h, int w, int ds,
                                       float* corr, int ldo, float* disp2, int ld2, float* extra, int ldx,
                                       void* stream) {
    if (!feat_curr || !feat_warp || !fea_l || !fea_r_planar || !pred_curr || !pred_warp || !corr || !disp2)
        return CODD_E_BADARG;
    if (n <= 0 || h <= 0 || w <= 0 || ds < 2 || cs <= 0) return CODD_E_BADARG;
    if (ldc < FC || ldw < FC || ldc % 4 || ldw % 4 || ldl < cs || ldl % 4 || cs % 4 || ldo < 32 || ldo % 4 || ld2 < 2)
        return CODD_E_SHAPE;
    if (!codd_aligned16(feat_curr) || !codd_aligned16(feat_warp) || !codd_aligned16(fea_l) || !codd_aligned16(corr))
        return CODD_E_ALIGN;
    CuesP p;
    p.feat_curr = feat_curr; p.ldc = ldc; p.feat_warp = feat_warp; p.ldw = ldw; p.fea_l = fea_l; p.ldl = ldl;
    p.fea_r = fea_r_planar; p.cs = cs; p.pred_curr = pred_curr; p.pred_warp = pred_warp;
    p.N = n; p.h = h; p.w = w; p.ds = ds; p.corr = corr; p.ldo = ldo; p.disp2 = disp2; p.ld2 = ld2;
    p.extra = extra; p.ldx = ldx;
    dim3 grid(codd_ceil_div(w, 128), h, n);
    fusion_cues_lowres_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(p);
    CODD_RETURN_IF_CUDA_ERROR();
    return 0;
}

extern "C" int codd_fusion_forget_in(const float* pred_curr, const float* pred_warp, const float* flow_warp,
                                     const float* conf_warp, const float* weight, const float* bias, int n, int h,
                                     int w, float* out, int ldo, float* cues_debug, void* stream) {
    if (!pred_curr || !pred_warp || !flow_warp || !conf_warp || !weight || !bias || !out) return CODD_E_BADARG;
    if (n <= 0 || h <= 0 || w <= 0) return CODD_E_BADARG;
    if (ldo < 16 || ldo % 4) return CODD_E_SHAPE;
    if (!codd_aligned16(out)) return CODD_E_ALIGN;
    ForgetP p;
    p.pred_curr = pred_curr; p.pred_warp = pred_warp; p.flow = flow_warp; p.conf = conf_warp; p.w = weight; p.b = bias;
    p.N = n; p.H = h; p.W = w; p.out = out; p.ldo = ldo; p.cues = cues_debug;
    dim3 grid(codd_ceil_div(w, 128), h, n);
    fusion_forget_in_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(p);
    CODD_RETURN_IF_CUDA_ERROR();
    return 0;
}

extern "C" int codd_fusion_blend(const float* pred_curr, const float* pred_warp, const float* r8, int ldr,
                                 const float* weight, const float* bias, const float* wf_lowres, int n, int h, int w,
                                 int ds, float* fused, float* wf, float* wr, void* stream) {
    if (!pred_curr || !pred_warp || !r8 || !weight || !bias || !wf_lowres || !fused || !wf || !wr) return CODD_E_BADARG;
    if (n <= 0 || h <= 0 || w <= 0 || ds <= 0 || h % ds || w % ds) return CODD_E_SHAPE;
    if (ldr < 8 || ldr % 4) return CODD_E_SHAPE;
    if (!codd_aligned16(r8)) return CODD_E_ALIGN;
    BlendP p;
    p.pred_curr = pred_curr; p.pred_warp = pred_warp; p.r8 = r8; p.ldr = ldr; p.w = weight; p.b = bias;
    p.wf_lr = wf_lowres; p.N = n; p.H = h; p.W = w; p.ds = ds; p.fused = fused; p.wf = wf; p.wr = wr;
    const size_t total = (size_t)n * h * w;
    fusion_blend_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p);
    CODD_RETURN_IF_CUDA_ERROR();
    return 0;
}
