// Non-convolutional layers of the RAFT3D networks on NHWC fp32 activations:
//   * InstanceNorm2d (+ ReLU, + residual add) of BasicEncoder — model/motion/raft3d/blocks/extractor.py:28-55,124-190
//   * bilinear resize with optional accumulate (HRNet fuse layers, align_corners=False; ResizeConcatConv,
//     align_corners=True) — model/motion/raft3d/raft3d.py:125-137 and mmseg's HRModule.forward
//   * element-wise glue of the update block: tanh/relu split of the context features (raft3d.py:183-186),
//     ConvGRU gating (blocks/gru.py:30-34), target = coords + delta (raft3d.py:242)
//   * disparity <-> depth conversion and strided sub-sampling of Motion.forward (motion.py:154-165,196-197;
//     raft3d.py:217-219)
// All of them are HBM-bound streaming kernels: one pass over the data, 128-bit accesses where the
// strides allow it.
#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------
// instance norm: pass 1 accumulates per (sample, channel) sum / sum of squares in double
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) inorm_stats_kernel(const float* __restrict__ in, int ldi, int hw, int c,
                                                          int pix_per_block, double* __restrict__ ws) {
    extern __shared__ float4 smem4[];
    float* red = reinterpret_cast<float*>(smem4);   // [2][blockDim.x]
    const int s = blockIdx.y;
    const int lanes = blockDim.x / c;               // pixel lanes per channel
    const int ch = threadIdx.x % c, pl = threadIdx.x / c;
    const int p0 = blockIdx.x * pix_per_block;
    const int p1 = min(p0 + pix_per_block, hw);
    float sum = 0.f, sq = 0.f;
    if (pl < lanes) {
        for (int p = p0 + pl; p < p1; p += lanes) {
            const float v = __ldg(in + ((size_t)s * hw + p) * ldi + ch);
            sum += v;
            sq = fmaf(v, v, sq);
        }
    }
    red[threadIdx.x] = sum;
    red[blockDim.x + threadIdx.x] = sq;
    __syncthreads();
    if (threadIdx.x < c) {
        double a = 0.0, b = 0.0;
        for (int l = 0; l < lanes; ++l) {
            a += (double)red[l * c + threadIdx.x];
            b += (double)red[blockDim.x + l * c + threadIdx.x];
        }
        atomicAdd(ws + ((size_t)s * c + threadIdx.x) * 2, a);
        atomicAdd(ws + ((size_t)s * c + threadIdx.x) * 2 + 1, b);
    }
}

// pass 2: y = (x - mean) * rsqrt(var + eps); optional ReLU; optional out = relu(residual + y)
__global__ void __launch_bounds__(256) inorm_apply_kernel(const float* __restrict__ in, int ldi, int hw, int c,
                                                          const double* __restrict__ ws, float eps, int relu,
                                                          const float* __restrict__ res, int ldr,
                                                          float* __restrict__ out, int ldo, size_t total) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // over n*hw*c, c fastest
    if (i >= total) return;
    const int ch = (int)(i % c);
    const size_t pix = i / c;
    const size_t s = pix / hw;
    const double m = ws[(s * c + ch) * 2] / hw;
    const double var = fmax(ws[(s * c + ch) * 2 + 1] / hw - m * m, 0.0);
    const float mean = (float)m;
    const float inv = (float)(1.0 / sqrt(var + (double)eps));
    float v = (__ldg(in + pix * ldi + ch) - mean) * inv;
    if (relu) v = fmaxf(v, 0.f);
    if (res) v = fmaxf(v + __ldg(res + pix * ldr + ch), 0.f);
    out[pix * ldo + ch] = v;
}

// ---------------------------------------------------------------------------------------------
// bilinear resize (torch F.interpolate semantics), out = act(base + interp(in))
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) resize_bilinear_kernel(const float* __restrict__ in, int ldi, int h, int w,
                                                              int c, const float* __restrict__ base, int ldb,
                                                              float* __restrict__ out, int ldo, int ho, int wo,
                                                              float sy, float sx, int align, int relu, size_t total) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // over n*ho*wo*c
    if (i >= total) return;
    const int ch = (int)(i % c);
    size_t t = i / c;
    const int ox = (int)(t % wo);
    t /= wo;
    const int oy = (int)(t % ho);
    const size_t s = t / ho;
    float fy, fx;
    if (align) {
        fy = sy * oy;
        fx = sx * ox;
    } else {
        fy = fmaxf(sy * (oy + 0.5f) - 0.5f, 0.f);
        fx = fmaxf(sx * (ox + 0.5f) - 0.5f, 0.f);
    }
    const int y0 = min((int)fy, h - 1), x0 = min((int)fx, w - 1);
    const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
    const float ly = fy - y0, lx = fx - x0;
    const float hy = 1.f - ly, hx = 1.f - lx;
    const float* b = in + (s * h) * (size_t)w * ldi + ch;
    const float v00 = __ldg(b + ((size_t)y0 * w + x0) * ldi), v01 = __ldg(b + ((size_t)y0 * w + x1) * ldi);
    const float v10 = __ldg(b + ((size_t)y1 * w + x0) * ldi), v11 = __ldg(b + ((size_t)y1 * w + x1) * ldi);
    float v = hy * (hx * v00 + lx * v01) + ly * (hx * v10 + lx * v11);   // torch's upsample_bilinear2d form
    const size_t opix = (s * ho + oy) * (size_t)wo + ox;
    if (base) v += __ldg(base + opix * ldb + ch);
    if (relu) v = fmaxf(v, 0.f);
    out[opix * ldo + ch] = v;
}

// ---------------------------------------------------------------------------------------------
// element-wise glue
// ---------------------------------------------------------------------------------------------
enum { EW_ACT = 0, EW_MUL = 1, EW_GRU = 2, EW_ADD_ACT = 3, EW_RECIP = 4 };

__global__ void __launch_bounds__(256) eltwise_kernel(int op, int act, const float* __restrict__ a, int lda,
                                                      const float* __restrict__ b, int ldb,
                                                      const float* __restrict__ cc, int ldc, float* __restrict__ out,
                                                      int ldo, int c, size_t total) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int ch = (int)(i % c);
    const size_t pix = i / c;
    const float av = __ldg(a + pix * lda + ch);
    float v;
    switch (op) {
        case EW_MUL: v = av * __ldg(b + pix * ldb + ch); break;
        case EW_GRU: {   // h' = (1 - z) * h + z * q   (a = z, b = h, cc = q)
            const float hv = __ldg(b + pix * ldb + ch), qv = __ldg(cc + pix * ldc + ch);
            v = (1.f - av) * hv + av * qv;
            break;
        }
        case EW_ADD_ACT: v = codd_act(av + __ldg(b + pix * ldb + ch), act, ch); break;
        case EW_RECIP: v = 1.f / av; break;
        default: v = codd_act(av, act, ch); break;
    }
    out[pix * ldo + ch] = v;
}

__global__ void __launch_bounds__(256) disp_to_depth_kernel(const float* __restrict__ disp, float bf, size_t total,
                                                            float* __restrict__ depth) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const float d = bf / (__ldg(disp + i) + 1e-5f);
    depth[i] = fminf(fmaxf(d, 0.f), bf);
}

__global__ void __launch_bounds__(256) subsample_kernel(const float* __restrict__ in, int ldi, int h, int w, int c,
                                                        int off, int stride, float* __restrict__ out, int ldo, int ho,
                                                        int wo, int recip, size_t total) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int ch = (int)(i % c);
    size_t t = i / c;
    const int ox = (int)(t % wo);
    t /= wo;
    const int oy = (int)(t % ho);
    const size_t s = t / ho;
    const float v = __ldg(in + ((s * h + off + (size_t)oy * stride) * w + off + (size_t)ox * stride) * ldi + ch);
    out[((s * ho + oy) * (size_t)wo + ox) * ldo + ch] = recip ? 1.f / v : v;
}

inline unsigned blocks_for(size_t total) { return (unsigned)((total + 255) / 256); }

}  // namespace

extern "C" size_t codd_instance_norm_workspace_bytes(int n, int c) { return (size_t)n * c * 2 * sizeof(double); }

extern "C" int codd_instance_norm_nhwc(const float* in, int ldi, int n, int h, int w, int c, float eps, int relu,
                                       const float* residual, int ldr, float* out, int ldo, void* workspace,
                                       size_t ws_bytes, void* stream) {
    if (!in || !out || !workspace || n <= 0 || h <= 0 || w <= 0 || c <= 0) return CODD_E_BADARG;
    if (ldi < c || ldo < c || (residual && ldr < c) || c > 256) return CODD_E_SHAPE;
    if (ws_bytes < codd_instance_norm_workspace_bytes(n, c)) return CODD_E_SHAPE;
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(workspace, 0, codd_instance_norm_workspace_bytes(n, c), s);
    if (e != cudaSuccess) return (int)e;
    const int hw = h * w;
    const int lanes = 256 / c > 0 ? 256 / c : 1;
    const int threads = lanes * c;
    const int ppb = 512;    // pixels per block
    dim3 grid((unsigned)codd_ceil_div(hw, ppb), (unsigned)n);
    inorm_stats_kernel<<<grid, threads, 2 * threads * sizeof(float), s>>>(in, ldi, hw, c, ppb, (double*)workspace);
    CODD_RETURN_IF_CUDA_ERROR();
    const size_t total = (size_t)n * hw * c;
    inorm_apply_kernel<<<blocks_for(total), 256, 0, s>>>(in, ldi, hw, c, (const double*)workspace, eps, relu, residual,
                                                         ldr, out, ldo, total);
    CODD_RETURN_IF_CUDA_ERROR();
    return 0;
}

extern "C" int codd_resize_bilinear_nhwc(const float* in, int ldi, int n, int h, int w, int c, const float* base,
                                         int ldb, float* out, int ldo, int ho, int wo, int align_corners, int relu,
                                         void* stream) {
    if (!in || !out || n <= 0 || h <= 0 || w <= 0 || c <= 0 || ho <= 0 || wo <= 0) return CODD_E_BADARG;
    if (ldi < c || ldo < c || (base && ldb < c)) return CODD_E_SHAPE;
    float sy, sx;
    if (align_corners) {
        sy = ho > 1 ? (float)(h - 1) / (float)(ho - 1) : 0.f;
        sx = wo > 1 ? (float)(w - 1) / (float)(wo - 1) : 0.f;
    } else {
        sy = (float)h / (float)ho;
        sx = (float)w / (float)wo;
    }
    const size_t total = (size_t)n * ho * wo * c;
    resize_bilinear_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(in, ldi, h, w, c, base, ldb, out, ldo,
                                                                                 ho, wo, sy, sx, align_corners, relu,
                                                                                 total);
    CODD_RETURN_IF_CUDA_ERROR();
    return 0;
}

extern "C" int codd_eltwise_nhwc(int op, int act, const float* a, int lda, const float* b, int ldb, const float* c,
                                 int ldc, float* out, int ldo, size_t npix, int channels, void* stream) {
    if (!a || !out || npix == 0 || channels <= 0) return CODD_E_BADARG;
    if (op < 0 || op > EW_RECIP) return CODD_E_UNSUPPORTED;
    if ((op == EW_MUL || op == EW_GRU || op == EW_ADD_ACT) && !b) return CODD_E_BADARG;
    if (op == EW_GRU && !c) return CODD_E_BADARG;
    if (lda < channels || ldo < channels) return CODD_E_SHAPE;
    const size_t total = npix * channels;
    eltwise_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(op, act, a, lda, b, ldb, c, ldc, out, ldo,
                                                                         channels, total);
    CODD_RETURN_IF_CUDA_ERROR();
    return 0;
}

extern "C" int codd_disp_to_depth(const float* disp, size_t count, float bf, float* depth, void* stream) {
    if (!disp || !depth || count == 0) return CODD_E_BADARG;
    disp_to_depth_kernel<<<blocks_for(count), 256, 0, (cudaStream_t)stream>>>(disp, bf, count, depth);
    CODD_RETURN_IF_CUDA_ERROR();
    return 0;
}

extern "C" int codd_subsample_nhwc(const float* in, int ldi, int n, int h, int w, int c, int offset, int stride,
                                   int recip, float* out, int ldo, void* stream) {
    if (!in || !out || n <= 0 || h <= 0 || w <= 0 || c <= 0 || stride <= 0 || offset < 0) return CODD_E_BADARG;
    if (offset >= h || offset >= w || ldi < c || ldo < c) return CODD_E_SHAPE;
    const int ho = (h - offset + stride - 1) / stride, wo = (w - offset + stride - 1) / stride;
    const size_t total = (size_t)n * ho * wo * c;
    subsample_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(in, ldi, h, w, c, offset, stride, out, ldo, ho,
                                                                          wo, recip, total);
    CODD_RETURN_IF_CUDA_ERROR();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// N1 — input staging: the reference normalises and pads every frame on the CPU
// (datasets/transforms.py:391-421 Normalize -> mmcv.imnormalize; :147-176 Pad(size_divisor=64) -> mmcv.impad
// 'reflect'; datasets/formating.py:77-85 HWC -> CHW float tensor).  One pass on the GPU instead: uint8 HWC in,
// normalised fp32 NCHW out, reflect-padded (no edge repeat) on the bottom / right to (hp, wp).
// ---------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) stage_images_u8_kernel(const uint8_t* __restrict__ img, int n, int h, int w,
                                                              float m0, float m1, float m2, float s0, float s1, float s2,
                                                              int to_rgb, int hp, int wp, float* __restrict__ out,
                                                              size_t total) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // over n * hp * wp
    if (i >= total) return;
    const int x = (int)(i % wp);
    const int y = (int)((i / wp) % hp);
    const size_t s = i / ((size_t)wp * hp);
    const int ys = y < h ? y : 2 * (h - 1) - y;       // numpy / cv2 BORDER_REFLECT_101
    const int xs = x < w ? x : 2 * (w - 1) - x;
    const uint8_t* px = img + ((s * h + ys) * (size_t)w + xs) * 3;
    const float c0 = (float)px[to_rgb ? 2 : 0], c1 = (float)px[1], c2 = (float)px[to_rgb ? 0 : 2];
    const size_t plane = (size_t)hp * wp;
    float* o = out + s * 3 * plane + (size_t)y * wp + x;
    o[0] = __fmul_rn(__fsub_rn(c0, m0), s0);          // (v - mean) * (1 / std), as mmcv.imnormalize
    o[plane] = __fmul_rn(__fsub_rn(c1, m1), s1);
    o[2 * plane] = __fmul_rn(__fsub_rn(c2, m2), s2);
}
}  // namespace

extern "C" int codd_stage_images_u8(const uint8_t* img_hwc, int n, int h, int w, const float* mean, const float* std_,
                                    int to_rgb, int hp, int wp, float* out, void* stream) {
    if (!img_hwc || !mean || !std_ || !out || n <= 0 || h <= 0 || w <= 0) return CODD_E_BADARG;
    if (hp < h || wp < w || hp - h >= h || wp - w >= w) return CODD_E_SHAPE;   // reflect needs pad < size
    const size_t total = (size_t)n * hp * wp;
    stage_images_u8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        img_hwc, n, h, w, mean[0], mean[1], mean[2], 1.f / std_[0], 1.f / std_[1], 1.f / std_[2], to_rgb, hp, wp, out, total);
    CODD_RETURN_IF_CUDA_ERROR();
    return 0;
}
