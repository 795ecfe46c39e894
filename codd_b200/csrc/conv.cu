// Dense convolutions of the HITNetMF stereo path on NHWC fp32 activations.
//
// Replaces every nn.Conv2d / nn.ConvTranspose2d launch of
//   model/stereo/hitnet/backbone.py:8-39,69-88
//   model/stereo/hitnet/initialization.py:62-117,119-156
//   model/stereo/hitnet/propagation.py:89-121,131-150,181-199,258-280,300-323
// with one register-tiled direct-convolution kernel family:
//   * CTA tile = 32 output columns x (4*PW) output rows x all output channels,
//   * each thread owns 4 vertically adjacent pixels x CO_T output channels in registers,
//   * input channels are streamed through shared memory 8 at a time (pixel-major, padded to
//     12 floats so the 128-bit activation loads are bank-conflict free), weights for the same
//     8 channels sit beside them as [tap][ci][co] and are read as warp-uniform broadcasts,
//   * bias, residual add (full or single-channel broadcast) and the activation are fused into
//     the epilogue; the "torch.cat" inputs of the reference are two source pointers.
// fp32 FMA accumulation (tolerance-level parity with the reference, see tests/).
#include <cstdlib>

#include <cuda.h>

#include "common.cuh"
#include "ring_util.cuh"

namespace {

struct ConvP {
    const float* in0;
    const float* in1;
    const float* w;
    const float* bias;
    const float* res;
    float* out;
    int N, H, W, C0, ld0, C1, ld1, Cout, ldo, ph, pw, Ho, Wo, act, ldr, res_bcast, res_after;
    int wld;         // row stride of the packed weight (= total Cout; Cout above may be one chunk of it)
    int tilesX, tilesY;
    int vec0, vec1;  // 128-bit loads allowed on in0 / in1
};

constexpr int CK = 8;    // input channels per shared-memory stage
constexpr int CP = 12;   // padded per-pixel stride of the stage (floats)
constexpr int PX = 4;    // output rows per thread
constexpr int TW = 32;   // output columns per CTA (one per lane)

__device__ __forceinline__ float conv_load1(const ConvP& p, size_t pix, int c) {
    if (c < p.C0) return __ldg(p.in0 + pix * p.ld0 + c);
    c -= p.C0;
    if (c < p.C1) return __ldg(p.in1 + pix * p.ld1 + c);
    return 0.f;
}

__device__ __forceinline__ float4 conv_load4(const ConvP& p, size_t pix, int c) {
    if (c + 3 < p.C0 && p.vec0) return ldg4(p.in0 + pix * p.ld0 + c);
    if (c >= p.C0 && c - p.C0 + 3 < p.C1 && p.vec1 && ((c - p.C0) & 3) == 0)
        return ldg4(p.in1 + pix * p.ld1 + (c - p.C0));
    return make_float4(conv_load1(p, pix, c), conv_load1(p, pix, c + 1), conv_load1(p, pix, c + 2),
                       conv_load1(p, pix, c + 3));
}

// Fast epilogue for one pixel's group of NC accumulators (the common case: full channel group, 16-byte aligned
// output / residual rows, branch-free activation).  ~5 instructions per element; the general path below
// (ragged Cout, unaligned rows, transcendental activations, residual after the activation) costs ~45 and, fully
// unrolled over 64 accumulators, made these kernels instruction-fetch bound.
struct EpiFast {
    bool ok, res_vec, res_b;
    bool v8;     // 32-byte sectors per thread (stg8 / ldg8): rows, channel base and pointers 32-byte aligned
    float slope, slope0;
};
__device__ __forceinline__ EpiFast epi_fast_setup(const ConvP& p, int cbase, int nc) {
    EpiFast e;
    e.res_b = p.res && p.res_bcast;
    e.res_vec = p.res && !p.res_bcast;
    e.ok = (cbase + nc <= p.Cout) && ((p.ldo & 3) == 0) && ((((uintptr_t)p.out) & 15u) == 0) &&
           (p.act <= CODD_ACT_RELU_CH0) && !p.res_after && (!p.bias || (((uintptr_t)p.bias) & 15u) == 0) &&
           (!e.res_vec || (((p.ldr & 3) == 0) && ((((uintptr_t)p.res) & 15u) == 0)));
    e.slope = p.act == CODD_ACT_LEAKY ? CODD_LEAKY_SLOPE : (p.act == CODD_ACT_RELU ? 0.f : 1.f);
    e.slope0 = (p.act == CODD_ACT_RELU || p.act == CODD_ACT_RELU_CH0) ? 0.f : e.slope;
    e.v8 = e.ok && (nc % 8 == 0) && ((cbase & 7) == 0) && ((p.ldo & 7) == 0) && ((((uintptr_t)p.out) & 31u) == 0) &&
           (!e.res_vec || (((p.ldr & 7) == 0) && ((((uintptr_t)p.res) & 31u) == 0)));
    return e;
}
template <int NC>
__device__ __forceinline__ void epi_fast_store(const ConvP& p, const EpiFast& e, const float* acc, size_t opix, int cbase) {
    float* op = p.out + opix * p.ldo + cbase;
    const float rb = e.res_b ? __ldg(p.res + opix * p.ldr) : 0.f;
    if (NC % 8 == 0 && e.v8) {
#pragma unroll
        for (int o8 = 0; o8 < NC / 8; ++o8) {
            float v[8], r[8];
            if (e.res_vec) ldg8(p.res + opix * p.ldr + cbase + o8 * 8, r);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
                if (p.bias) b = ldg4(p.bias + cbase + o8 * 8 + h * 4);
                v[h * 4 + 0] = acc[o8 * 8 + h * 4 + 0] + b.x;
                v[h * 4 + 1] = acc[o8 * 8 + h * 4 + 1] + b.y;
                v[h * 4 + 2] = acc[o8 * 8 + h * 4 + 2] + b.z;
                v[h * 4 + 3] = acc[o8 * 8 + h * 4 + 3] + b.w;
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                v[k] += e.res_vec ? r[k] : rb;
                const float sl = (cbase + o8 * 8 + k == 0) ? e.slope0 : e.slope;
                v[k] = fmaxf(v[k], 0.f) + sl * fminf(v[k], 0.f);
            }
            stg8(op + o8 * 8, v);
        }
        return;
    }
#pragma unroll
    for (int o4 = 0; o4 < NC / 4; ++o4) {
        float4 v = make_float4(acc[o4 * 4], acc[o4 * 4 + 1], acc[o4 * 4 + 2], acc[o4 * 4 + 3]);
        if (p.bias) {
            const float4 b = ldg4(p.bias + cbase + o4 * 4);
            v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
        }
        if (e.res_vec) {
            const float4 r = ldg4(p.res + opix * p.ldr + cbase + o4 * 4);
            v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
        } else {
            v.x += rb; v.y += rb; v.z += rb; v.w += rb;
        }
        const float s0 = (cbase + o4 == 0) ? e.slope0 : e.slope;
        v.x = fmaxf(v.x, 0.f) + s0 * fminf(v.x, 0.f);
        v.y = fmaxf(v.y, 0.f) + e.slope * fminf(v.y, 0.f);
        v.z = fmaxf(v.z, 0.f) + e.slope * fminf(v.z, 0.f);
        v.w = fmaxf(v.w, 0.f) + e.slope * fminf(v.w, 0.f);
        *reinterpret_cast<float4*>(op + o4 * 4) = v;
    }
}

template <int KH, int KW, int SH, int SW, int DIL, int CO_T>
__global__ void __launch_bounds__(256, 2) conv_nhwc_kernel(ConvP p) {
    constexpr int IW = (TW - 1) * SW + (KW - 1) * DIL + 1;
    const int lane = threadIdx.x, pwi = threadIdx.y, g = threadIdx.z;
    const int PWn = blockDim.y, G = blockDim.z;
    const int TH = PWn * PX;
    const int IH = (TH - 1) * SH + (KH - 1) * DIL + 1;
    const int COP = G * CO_T;
    const int nthreads = 32 * PWn * G;
    const int tid = (g * PWn + pwi) * 32 + lane;

    extern __shared__ float4 smem4[];
    float* s_in = reinterpret_cast<float*>(smem4);
    float* s_w = s_in + IH * IW * CP;

    int tile = blockIdx.x;
    const int tx = tile % p.tilesX;
    tile /= p.tilesX;
    const int ty = tile % p.tilesY;
    const int n = tile / p.tilesY;
    const int oy0 = ty * TH, ox0 = tx * TW;
    const int iy0 = oy0 * SH - p.ph, ix0 = ox0 * SW - p.pw;
    const int Cin = p.C0 + p.C1;

    __align__(8) float acc[PX][CO_T];
#pragma unroll
    for (int i = 0; i < PX; ++i)
#pragma unroll
        for (int j = 0; j < CO_T; ++j) acc[i][j] = 0.f;

    // Stride-2 layers store even and odd input columns in separate halves of a staged row: the lanes of a warp
    // (consecutive output columns) then read consecutive slots for every tap instead of every second one, which
    // with 48-byte pixel slots was an 8-way bank conflict on the 128-bit activation loads.
    constexpr int IWH = (IW + 1) / 2;
    auto col_slot = [&](int cc) { return SW == 2 ? (cc & 1) * IWH + (cc >> 1) : cc; };
    const float* sa_base = s_in + ((pwi * PX * SH) * IW) * CP;

    for (int c0 = 0; c0 < Cin; c0 += CK) {
        __syncthreads();
        // ---- stage activations: IH x IW pixels x 8 channels
        for (int idx = tid; idx < IH * IW * 2; idx += nthreads) {
            const int c4 = idx & 1;
            const int pix = idx >> 1;
            const int r = pix / IW, cc = pix - r * IW;
            const int gy = iy0 + r, gx = ix0 + cc;
            const int c = c0 + c4 * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gy >= 0 && gy < p.H && gx >= 0 && gx < p.W && c < Cin)
                v = conv_load4(p, ((size_t)n * p.H + gy) * p.W + gx, c);
            *reinterpret_cast<float4*>(s_in + (r * IW + col_slot(cc)) * CP + c4 * 4) = v;
        }
        // ---- stage weights: [tap][ci][COP]
        for (int idx = tid; idx < KH * KW * CK * COP; idx += nthreads) {
            const int co = idx % COP;
            const int t = idx / COP;
            const int ci = t % CK, tap = t / CK;
            float v = 0.f;
            if (co < p.Cout && c0 + ci < Cin) v = __ldg(p.w + ((size_t)tap * Cin + c0 + ci) * p.wld + co);
            s_w[idx] = v;
        }
        __syncthreads();

        // taps are NOT unrolled: one tap body is 8 ci x 4 px x CO_T FMAs, plenty of ILP, and
        // keeping the loop rolled bounds live registers (no spills at 128 regs) and code size.
#pragma unroll 1
        for (int ky = 0; ky < KH; ++ky) {
#pragma unroll 1
            for (int kx = 0; kx < KW; ++kx) {
                const float* sa = sa_base + ((ky * DIL) * IW + col_slot(lane * SW + kx * DIL)) * CP;
                const float* sw = s_w + ((ky * KW + kx) * CK) * COP + g * CO_T;
#pragma unroll
                for (int c4 = 0; c4 < 2; ++c4) {
                    float4 a[PX];
#pragma unroll
                    for (int i = 0; i < PX; ++i)
                        a[i] = *reinterpret_cast<const float4*>(sa + i * SH * IW * CP + c4 * 4);
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
#pragma unroll
                        for (int o4 = 0; o4 < CO_T / 4; ++o4) {
                            const float4 wv = *reinterpret_cast<const float4*>(sw + (c4 * 4 + cc) * COP + o4 * 4);
#pragma unroll
                            for (int i = 0; i < PX; ++i) {
                                const float av = cc == 0 ? a[i].x : cc == 1 ? a[i].y : cc == 2 ? a[i].z : a[i].w;
                                fma4(&acc[i][o4 * 4], av, wv);
                            }
                        }
                    }
                }
            }
        }
    }

    // ---- epilogue: bias + residual + activation, NHWC store
    const int ox = ox0 + lane;
    if (ox >= p.Wo) return;
    const int cbase = g * CO_T;
    const bool vec_out = ((p.ldo & 3) == 0) && ((((uintptr_t)p.out) & 15u) == 0);
    const ActSel asel = codd_act_sel(p.act);
    const EpiFast fast = epi_fast_setup(p, cbase, CO_T);
#pragma unroll
    for (int i = 0; i < PX; ++i) {
        const int oy = oy0 + pwi * PX + i;
        if (oy >= p.Ho) continue;
        const size_t opix = ((size_t)n * p.Ho + oy) * p.Wo + ox;
        if (fast.ok) {
            epi_fast_store<CO_T>(p, fast, acc[i], opix, cbase);
            continue;
        }
        float* op = p.out + opix * p.ldo;
        float rb = 0.f;
        if (p.res && p.res_bcast) rb = __ldg(p.res + opix * p.ldr);
#pragma unroll
        for (int o4 = 0; o4 < CO_T / 4; ++o4) {
            const int co = cbase + o4 * 4;
            if (co >= p.Cout) break;
            float v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float t = acc[i][o4 * 4 + e];
                const int ce = co + e;
                if (ce < p.Cout) {
                    if (p.bias) t += __ldg(p.bias + ce);
                    const float rv = p.res ? (p.res_bcast ? rb : __ldg(p.res + opix * p.ldr + ce)) : 0.f;
                    t = p.res_after ? codd_act_apply(asel, t, ce) + rv : codd_act_apply(asel, t + rv, ce);
                }
                v[e] = t;
            }
            if (co + 3 < p.Cout && vec_out) {
                *reinterpret_cast<float4*>(op + co) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (co + e < p.Cout) op[co + e] = v[e];
            }
        }
    }
}

template <int KH, int KW, int SH, int SW, int DIL, int CO_T>
int launch_conv(ConvP p, int PWn, int G, cudaStream_t stream) {
    constexpr int IW = (TW - 1) * SW + (KW - 1) * DIL + 1;
    auto smem_for = [&](int pw) {
        const int IH = (pw * PX - 1) * SH + (KH - 1) * DIL + 1;
        return (size_t)(IH * IW * CP + KH * KW * CK * G * CO_T) * sizeof(float);
    };
    // keep the stage under ~100 KB so two CTAs fit on an SM; shrink the row count if needed
    while (PWn > 1 && smem_for(PWn) > 110 * 1024) PWn >>= 1;
    // do not launch row-warps that would only see padding
    while (PWn > 1 && (PWn / 2) * PX >= p.Ho) PWn >>= 1;
    const size_t smem = smem_for(PWn);
    if (smem > 227 * 1024) return CODD_E_UNSUPPORTED;
    auto kern = conv_nhwc_kernel<KH, KW, SH, SW, DIL, CO_T>;
    static CoddDeviceOnce once;   // per instantiation and device: graph capture sees no attribute calls afterwards
    if (int rc = codd_once_per_device(once, [&] {
            return codd_max_dynamic_smem(kern);
        }))
        return rc;
    p.tilesX = codd_ceil_div(p.Wo, TW);
    p.tilesY = codd_ceil_div(p.Ho, PWn * PX);
    dim3 block(32, PWn, G);
    dim3 grid((unsigned)(p.tilesX * p.tilesY * p.N));
    kern<<<grid, block, smem, stream>>>(p);
    CODD_RETURN_IF_CUDA_ERROR();
    return 0;
}

// FULL: instantiate the narrow per-thread channel tiles too (the geometries that see Cout in
// {1,3,13,24,34}); the other geometries only ever run with Cout in {16,24,32} and pad to 16s.
template <int KH, int KW, int SH, int SW, int DIL, bool FULL>
int dispatch_cout(const ConvP& p, cudaStream_t s) {
    const int co = p.Cout;
    if constexpr (FULL) {
        if (co <= 4) return launch_conv<KH, KW, SH, SW, DIL, 4>(p, 8, 1, s);
        if (co <= 8) return launch_conv<KH, KW, SH, SW, DIL, 8>(p, 8, 1, s);
        if (co > 16 && co <= 24) return launch_conv<KH, KW, SH, SW, DIL, 12>(p, 4, 2, s);
        if (co > 32 && co <= 36) return launch_conv<KH, KW, SH, SW, DIL, 12>(p, 2, 3, s);
    }
    if (co <= 16) return launch_conv<KH, KW, SH, SW, DIL, 16>(p, 8, 1, s);
    if (co <= 32) return launch_conv<KH, KW, SH, SW, DIL, 16>(p, 4, 2, s);
    if (co <= 48) return launch_conv<KH, KW, SH, SW, DIL, 16>(p, 2, 3, s);
    if (co <= 64) return launch_conv<KH, KW, SH, SW, DIL, 16>(p, 2, 4, s);
    return CODD_E_UNSUPPORTED;
}

// ---------------------------------------------------------------------------------------------
// 1x1 convolutions (merge*.0, conv0, conv1.0): a per-pixel GEMV.  No activation staging: every
// thread streams its own PX pixels (128-bit loads straight from the two concatenated sources),
// the whole [Cin][Cout] weight matrix sits in shared memory and is read as warp-uniform
// broadcasts.  HBM-bound (AI = 2*Cin*Cout / (4*(Cin+Cout)) < 12 flop/B for every layer here).
// ---------------------------------------------------------------------------------------------
template <int CO, int PX_T, bool V8IN>
__global__ void __launch_bounds__(256, 2) pointwise_kernel(ConvP p, size_t npix, int wbulk) {
    extern __shared__ float4 smem4[];
    float* s_w = reinterpret_cast<float*>(smem4);  // [Cin][CO]
    const int Cin = p.C0 + p.C1;
    if (wbulk) {
        // the packed weight IS the shared image (Cout == CO == row stride, 16-byte aligned): one 1-D bulk copy instead of
        // 2-8 scalar loads per thread and CTA (a CTA only covers 256 * PX_T pixels)
        __shared__ __align__(8) unsigned long long wbar;
        const uint32_t b = (uint32_t)__cvta_generic_to_shared(&wbar);
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            const uint32_t bytes = (uint32_t)(Cin * CO * 4);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"((uint32_t)__cvta_generic_to_shared(s_w)), "l"(p.w), "r"(bytes), "r"(b) : "memory");
        }
        __syncthreads();
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}"
                         : "=r"(done) : "r"(b), "r"(0) : "memory");
    } else {
        for (int i = threadIdx.x; i < Cin * CO; i += blockDim.x) {
            const int co = i % CO, ci = i / CO;
            s_w[i] = co < p.Cout ? __ldg(p.w + (size_t)ci * p.wld + co) : 0.f;
        }
        __syncthreads();
    }
    const size_t base = (size_t)blockIdx.x * (blockDim.x * PX_T) + threadIdx.x;
    __align__(8) float acc[PX_T][CO];
#pragma unroll
    for (int q = 0; q < PX_T; ++q)
#pragma unroll
        for (int c = 0; c < CO; ++c) acc[q][c] = 0.f;
    size_t pix[PX_T];
#pragma unroll
    for (int q = 0; q < PX_T; ++q) pix[q] = min(base + (size_t)q * blockDim.x, npix - 1);

    if (V8IN) {
        // 8 channels per step as one 256-bit load per pixel (whole 32-byte sectors; see ldg8)
        for (int c = 0; c < Cin; c += 8) {
            float a[PX_T][8];
            const bool first = c < p.C0;
#pragma unroll
            for (int q = 0; q < PX_T; ++q)
                ldg8(first ? p.in0 + pix[q] * p.ld0 + c : p.in1 + pix[q] * p.ld1 + (c - p.C0), a[q]);
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) {
                const float* wp = s_w + (c + cc) * CO;
#pragma unroll
                for (int o4 = 0; o4 < CO / 4; ++o4) {
                    const float4 wv = *reinterpret_cast<const float4*>(wp + o4 * 4);
#pragma unroll
                    for (int q = 0; q < PX_T; ++q) fma4(&acc[q][o4 * 4], a[q][cc], wv);
                }
            }
        }
    } else {
        for (int c = 0; c < Cin; c += 4) {
            float4 a[PX_T];
            const bool first = c < p.C0;
#pragma unroll
            for (int q = 0; q < PX_T; ++q)
                a[q] = first ? ldg4(p.in0 + pix[q] * p.ld0 + c) : ldg4(p.in1 + pix[q] * p.ld1 + (c - p.C0));
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                const float* wp = s_w + (c + cc) * CO;
#pragma unroll
                for (int o4 = 0; o4 < CO / 4; ++o4) {
                    const float4 wv = *reinterpret_cast<const float4*>(wp + o4 * 4);
#pragma unroll
                    for (int q = 0; q < PX_T; ++q) {
                        const float av = cc == 0 ? a[q].x : cc == 1 ? a[q].y : cc == 2 ? a[q].z : a[q].w;
                        fma4(&acc[q][o4 * 4], av, wv);
                    }
                }
            }
        }
    }
    const ActSel asel = codd_act_sel(p.act);
    const EpiFast fast = epi_fast_setup(p, 0, CO);
#pragma unroll
    for (int q = 0; q < PX_T; ++q) {
        const size_t px = base + (size_t)q * blockDim.x;
        if (px >= npix) continue;
        if (fast.ok) {
            epi_fast_store<CO>(p, fast, acc[q], px, 0);
            continue;
        }
        float* op = p.out + px * p.ldo;
        float rb = 0.f;
        if (p.res && p.res_bcast) rb = __ldg(p.res + px * p.ldr);
#pragma unroll
        for (int o4 = 0; o4 < CO / 4; ++o4) {
            float v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int ce = o4 * 4 + e;
                float t = acc[q][ce];
                if (ce < p.Cout) {
                    if (p.bias) t += __ldg(p.bias + ce);
                    const float rv = p.res ? (p.res_bcast ? rb : __ldg(p.res + px * p.ldr + ce)) : 0.f;
                    t = p.res_after ? codd_act_apply(asel, t, ce) + rv : codd_act_apply(asel, t + rv, ce);
                }
                v[e] = t;
            }
            if (o4 * 4 + 3 < p.Cout) {
                *reinterpret_cast<float4*>(op + o4 * 4) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (o4 * 4 + e < p.Cout) op[o4 * 4 + e] = v[e];
            }
        }
    }
}

// (A shared-memory-staged 1x1 variant was tried in round 2 and measured SLOWER on the GPU: 1.29 ms vs 0.81 ms per step
// for the 32->16 full-resolution layers — the extra shared-memory round trip costs more than the uncoalesced
// 64-byte-stride loads it removes.  Removed; see DESIGN.md.)

template <int CO, int PX_T>
int launch_pointwise(const ConvP& p, cudaStream_t s) {
    const size_t npix = (size_t)p.N * p.H * p.W;
    const size_t smem = (size_t)(p.C0 + p.C1) * CO * sizeof(float);
    const unsigned grid = (unsigned)((npix + 256 * PX_T - 1) / (256 * PX_T));
    // 256-bit input loads when both sources are made of whole, 32-byte aligned 8-channel groups
    const bool v8 = (p.C0 % 8 == 0) && (p.C1 % 8 == 0) && (p.ld0 % 8 == 0) && codd_aligned32(p.in0) &&
                    (p.C1 == 0 || ((p.ld1 % 8 == 0) && codd_aligned32(p.in1)));
    const int wbulk = (p.Cout == CO) && (p.wld == CO) && codd_aligned16(p.w) && (((p.C0 + p.C1) * CO) % 4 == 0);
    if (v8) pointwise_kernel<CO, PX_T, true><<<grid, 256, smem, s>>>(p, npix, wbulk);
    else pointwise_kernel<CO, PX_T, false><<<grid, 256, smem, s>>>(p, npix, wbulk);
    CODD_RETURN_IF_CUDA_ERROR();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// first layer: NCHW image (3 ch) -> NHWC, 3x3 pad 1, LeakyReLU
// ---------------------------------------------------------------------------------------------
constexpr int IMG_ROWS = 8;   // output rows per thread (the row loop re-uses the ~6 KB tap body from the I-cache)
template <int CO>
__global__ void __launch_bounds__(128) conv3x3_image_kernel(const float* __restrict__ left,
                                                            const float* __restrict__ right, int n, int h,
                                                            int w, const float* __restrict__ wgt,
                                                            const float* __restrict__ bias, int cout,
                                                            float* __restrict__ out, int ldo) {
    // each thread computes 4 horizontally adjacent pixels of IMG_ROWS consecutive rows: one weight broadcast
    // feeds 4 packed FMAs; the ky / row loops stay ROLLED so that the instruction footprint is one tap row
    constexpr int PXI = 4;
    __shared__ __align__(16) float s_w[27 * CO];
    __shared__ __align__(16) float s_b[CO];
    for (int i = threadIdx.x; i < 27 * CO; i += blockDim.x) {
        const int co = i % CO, t = i / CO;
        s_w[i] = co < cout ? wgt[t * cout + co] : 0.f;
    }
    for (int i = threadIdx.x; i < CO; i += blockDim.x) s_b[i] = i < cout ? bias[i] : 0.f;
    __syncthreads();
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * PXI;
    const int s = blockIdx.z;  // sample in [0, 2n)
    if (x0 >= w) return;
    const float* img = (s < n ? left + (size_t)s * 3 * h * w : right + (size_t)(s - n) * 3 * h * w);
    const bool vec = (ldo & 3) == 0 && cout == CO && ((((uintptr_t)out) & 15u) == 0);
    const bool vec8 = vec && (CO % 8 == 0) && (ldo & 7) == 0 && ((((uintptr_t)out) & 31u) == 0);
    const int y_end = min(h, (int)(blockIdx.y + 1) * IMG_ROWS);
#pragma unroll 1
    for (int y = blockIdx.y * IMG_ROWS; y < y_end; ++y) {
        __align__(8) float acc[PXI][CO];
#pragma unroll
        for (int q = 0; q < PXI; ++q)
#pragma unroll
            for (int i = 0; i < CO; ++i) acc[q][i] = s_b[i];
#pragma unroll 1
        for (int ky = 0; ky < 3; ++ky) {
            const int yy = y + ky - 1;
            if (yy < 0 || yy >= h) continue;
#pragma unroll
            for (int ci = 0; ci < 3; ++ci) {
                const float* rp = img + ((size_t)ci * h + yy) * w;
                float a[PXI + 2];
#pragma unroll
                for (int e = 0; e < PXI + 2; ++e) {
                    const int xx = x0 + e - 1;
                    a[e] = (xx >= 0 && xx < w) ? __ldg(rp + xx) : 0.f;
                }
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const float* wp = s_w + ((ky * 3 + kx) * 3 + ci) * CO;
#pragma unroll
                    for (int o4 = 0; o4 < CO / 4; ++o4) {
                        const float4 wv = *reinterpret_cast<const float4*>(wp + o4 * 4);
#pragma unroll
                        for (int q = 0; q < PXI; ++q) fma4(&acc[q][o4 * 4], a[q + kx], wv);
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < PXI; ++q) {
            if (x0 + q >= w) break;
            float* op = out + (((size_t)s * h + y) * w + x0 + q) * ldo;
            if (vec8) {
#pragma unroll
                for (int i = 0; i < CO; ++i) acc[q][i] = fmaxf(acc[q][i], 0.f) + CODD_LEAKY_SLOPE * fminf(acc[q][i], 0.f);
#pragma unroll
                for (int o8 = 0; o8 < CO / 8; ++o8) stg8(op + o8 * 8, &acc[q][o8 * 8]);
            } else if (vec) {
#pragma unroll
                for (int o4 = 0; o4 < CO / 4; ++o4) {
                    float4 v = make_float4(acc[q][o4 * 4], acc[q][o4 * 4 + 1], acc[q][o4 * 4 + 2], acc[q][o4 * 4 + 3]);
                    v.x = fmaxf(v.x, 0.f) + CODD_LEAKY_SLOPE * fminf(v.x, 0.f);
                    v.y = fmaxf(v.y, 0.f) + CODD_LEAKY_SLOPE * fminf(v.y, 0.f);
                    v.z = fmaxf(v.z, 0.f) + CODD_LEAKY_SLOPE * fminf(v.z, 0.f);
                    v.w = fmaxf(v.w, 0.f) + CODD_LEAKY_SLOPE * fminf(v.w, 0.f);
                    *reinterpret_cast<float4*>(op + o4 * 4) = v;
                }
            } else {
#pragma unroll
                for (int i = 0; i < CO; ++i)
                    if (i < cout) op[i] = codd_act(acc[q][i], CODD_ACT_LEAKY, 0);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// single-output 3x3 head: FinalTileUpdate's last layer in eval mode needs the disparity channel only
// (propagation.py:325-333, 16 -> 3 channels of which channel 0 is returned).  16 -> 1 is a memory-bound reduction
// (144 MACs per 64 bytes read), not a GEMM: a (32+2) x (8+2) NHWC tile is staged in shared memory with 128-bit copies
// (pixel pitch CIN+4 floats: conflict-free 128-bit reads for adjacent lanes), one thread per output pixel.
// ---------------------------------------------------------------------------------------------
constexpr int H1_TW = 32, H1_TH = 8;
template <int CIN>
__global__ void __launch_bounds__(H1_TW * H1_TH) conv3x3_head1_kernel(const float* __restrict__ in, int ldi, int n, int h,
                                                                      int w, const float* __restrict__ wgt,
                                                                      const float* __restrict__ bias,
                                                                      const float* __restrict__ res, int ldr, int act,
                                                                      float* __restrict__ out, int ldo) {
    constexpr int PITCH = CIN + 4;
    __shared__ __align__(16) float s_t[(H1_TH + 2) * (H1_TW + 2) * PITCH];
    __shared__ __align__(16) float s_w[9 * CIN];
    const int tid = threadIdx.y * H1_TW + threadIdx.x;
    const int x0 = blockIdx.x * H1_TW, y0 = blockIdx.y * H1_TH, s = blockIdx.z;
    for (int i = tid; i < 9 * CIN; i += H1_TW * H1_TH) s_w[i] = __ldg(wgt + i);       // packed [tap][cin][1]
    for (int i = tid; i < (H1_TH + 2) * (H1_TW + 2) * (CIN / 4); i += H1_TW * H1_TH) {
        const int c4 = i % (CIN / 4), pix = i / (CIN / 4);
        const int px = pix % (H1_TW + 2), py = pix / (H1_TW + 2);
        const int gx = x0 + px - 1, gy = y0 + py - 1;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gx >= 0 && gx < w && gy >= 0 && gy < h) v = ldg4(in + (((size_t)s * h + gy) * w + gx) * ldi + c4 * 4);
        *reinterpret_cast<float4*>(s_t + pix * PITCH + c4 * 4) = v;
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x >= w || y >= h) return;
    float acc = bias ? __ldg(bias) : 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const float* tp = s_t + ((threadIdx.y + ky) * (H1_TW + 2) + threadIdx.x + kx) * PITCH;
            const float* wp = s_w + (ky * 3 + kx) * CIN;
#pragma unroll
            for (int c4 = 0; c4 < CIN / 4; ++c4) {
                const float4 a = *reinterpret_cast<const float4*>(tp + c4 * 4);
                const float4 b = *reinterpret_cast<const float4*>(wp + c4 * 4);
                acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
            }
        }
    const size_t opix = ((size_t)s * h + y) * w + x;
    if (res) acc += __ldg(res + opix * ldr);
    out[opix * ldo] = codd_act(acc, act, 0);
}

// Two-output 3x3 head, 32 input channels: the two confidence filters of TileUpdate.lastconv (propagation.py:190-199,
// channels 32 and 33 of a 34-channel layer whose first 32 filters run on the tensor cores).  Same scheme as the
// single-output head: a (32+2) x (4+2) NHWC tile in shared memory (pixel pitch CIN+4), one thread per output pixel, both
// filters from one activation read.  As a general direct convolution with two live filters this took 48 us per launch.
constexpr int H2_TW = 32, H2_TH = 4;
template <int CIN>
__global__ void __launch_bounds__(H2_TW * H2_TH) conv3x3_head2_kernel(const float* __restrict__ in, int ldi, int n, int h,
                                                                      int w, const float* __restrict__ wgt, int wld,
                                                                      const float* __restrict__ bias,
                                                                      const float* __restrict__ res, int ldr, int res_bcast,
                                                                      int act, float* __restrict__ out, int ldo) {
    constexpr int PITCH = CIN + 4;
    __shared__ __align__(16) float s_t[(H2_TH + 2) * (H2_TW + 2) * PITCH];
    __shared__ __align__(16) float s_w[9 * 2 * CIN];                                  // [tap][filter][cin]
    const int tid = threadIdx.y * H2_TW + threadIdx.x;
    const int x0 = blockIdx.x * H2_TW, y0 = blockIdx.y * H2_TH, s = blockIdx.z;
    for (int i = tid; i < 9 * 2 * CIN; i += H2_TW * H2_TH) {
        const int ci = i % CIN, co = (i / CIN) & 1, tap = i / (2 * CIN);
        s_w[i] = __ldg(wgt + ((size_t)tap * CIN + ci) * wld + co);                    // packed [tap][cin][wld]
    }
    for (int i = tid; i < (H2_TH + 2) * (H2_TW + 2) * (CIN / 4); i += H2_TW * H2_TH) {
        const int c4 = i % (CIN / 4), pix = i / (CIN / 4);
        const int px = pix % (H2_TW + 2), py = pix / (H2_TW + 2);
        const int gx = x0 + px - 1, gy = y0 + py - 1;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gx >= 0 && gx < w && gy >= 0 && gy < h) v = ldg4(in + (((size_t)s * h + gy) * w + gx) * ldi + c4 * 4);
        *reinterpret_cast<float4*>(s_t + pix * PITCH + c4 * 4) = v;
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x >= w || y >= h) return;
    float a0 = bias ? __ldg(bias) : 0.f, a1 = bias ? __ldg(bias + 1) : 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const float* tp = s_t + ((threadIdx.y + ky) * (H2_TW + 2) + threadIdx.x + kx) * PITCH;
            const float* wp = s_w + (ky * 3 + kx) * 2 * CIN;
#pragma unroll
            for (int c4 = 0; c4 < CIN / 4; ++c4) {
                const float4 a = *reinterpret_cast<const float4*>(tp + c4 * 4);
                const float4 u = *reinterpret_cast<const float4*>(wp + c4 * 4);
                const float4 v = *reinterpret_cast<const float4*>(wp + CIN + c4 * 4);
                a0 = fmaf(a.x, u.x, a0); a0 = fmaf(a.y, u.y, a0); a0 = fmaf(a.z, u.z, a0); a0 = fmaf(a.w, u.w, a0);
                a1 = fmaf(a.x, v.x, a1); a1 = fmaf(a.y, v.y, a1); a1 = fmaf(a.z, v.z, a1); a1 = fmaf(a.w, v.w, a1);
            }
        }
    const size_t opix = ((size_t)s * h + y) * w + x;
    if (res) {
        a0 += __ldg(res + opix * ldr);
        a1 += __ldg(res + opix * ldr + (res_bcast ? 0 : 1));
    }
    out[opix * ldo] = codd_act(a0, act, 0);
    out[opix * ldo + 1] = codd_act(a1, act, 1);
}

// ---------------------------------------------------------------------------------------------
// ConvTranspose2d k=2 s=2: every output pixel sees exactly one input pixel and one of 4 taps
// ---------------------------------------------------------------------------------------------
constexpr int DC_PX = 2;   // input pixels per thread: every weight broadcast (2 x LDS.128) feeds 2 x 2 x 2 packed FMAs
template <int CO>
__global__ void __launch_bounds__(128) deconv2x2_kernel(const float* __restrict__ in, int ldi, int n, int h,
                                                        int w, int cin, const float* __restrict__ wgt,
                                                        const float* __restrict__ bias, int cout,
                                                        float* __restrict__ out, int ldo, int act) {
    // one thread per DC_PX horizontally adjacent INPUT pixels (x, x + 128) and output-row parity dy: per input pixel
    // it produces the two output pixels (2y+dy, 2x) and (2y+dy, 2x+1); every weight broadcast (warp-uniform:
    // dy is the block's) feeds both pixels' output channels.
    extern __shared__ float4 smem4[];
    float* s_w = reinterpret_cast<float*>(smem4);  // [4][cin][CO]
    for (int i = threadIdx.x; i < 4 * cin * CO; i += blockDim.x) {
        const int co = i % CO, t = i / CO;
        s_w[i] = co < cout ? wgt[(size_t)t * cout + co] : 0.f;
    }
    __syncthreads();
    const int x0 = blockIdx.x * (blockDim.x * DC_PX) + threadIdx.x;
    const int oy = blockIdx.y;          // output row
    const int s = blockIdx.z;
    if (x0 >= w) return;
    const int dy = oy & 1;
    const float* ip[DC_PX];
#pragma unroll
    for (int q = 0; q < DC_PX; ++q)
        ip[q] = in + (((size_t)s * h + (oy >> 1)) * w + min(x0 + q * (int)blockDim.x, w - 1)) * ldi;
    __align__(8) float acc[DC_PX][2][CO];
#pragma unroll
    for (int q = 0; q < DC_PX; ++q)
#pragma unroll
        for (int d = 0; d < 2; ++d)
#pragma unroll
            for (int i = 0; i < CO; ++i) acc[q][d][i] = 0.f;
    const float* w0 = s_w + (dy * 2 + 0) * cin * CO;
    const float* w1 = s_w + (dy * 2 + 1) * cin * CO;
    for (int ci = 0; ci < cin; ci += 4) {
        float4 a4[DC_PX];
#pragma unroll
        for (int q = 0; q < DC_PX; ++q) a4[q] = ldg4(ip[q] + ci);
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
#pragma unroll
            for (int o4 = 0; o4 < CO / 4; ++o4) {
                const float4 u = *reinterpret_cast<const float4*>(w0 + (ci + cc) * CO + o4 * 4);
                const float4 v = *reinterpret_cast<const float4*>(w1 + (ci + cc) * CO + o4 * 4);
#pragma unroll
                for (int q = 0; q < DC_PX; ++q) {
                    const float a = cc == 0 ? a4[q].x : cc == 1 ? a4[q].y : cc == 2 ? a4[q].z : a4[q].w;
                    fma4(&acc[q][0][o4 * 4], a, u);
                    fma4(&acc[q][1][o4 * 4], a, v);
                }
            }
        }
    }
    const bool vec = (ldo & 3) == 0 && (cout & 3) == 0 && ((((uintptr_t)out) & 15u) == 0) && ((((uintptr_t)bias) & 15u) == 0);
    const bool vec8 = vec && (CO % 8 == 0) && (ldo & 7) == 0 && (cout & 7) == 0 && ((((uintptr_t)out) & 31u) == 0);
    const ActSel asel = codd_act_sel(act);
#pragma unroll
    for (int q = 0; q < DC_PX; ++q) {
        const int x = x0 + q * (int)blockDim.x;
        if (x >= w) break;
        float* op = out + (((size_t)s * 2 * h + oy) * 2 * w + 2 * x) * ldo;
#pragma unroll
        for (int d = 0; d < 2; ++d) {
            if (vec8 && asel.simple) {
#pragma unroll
                for (int o8 = 0; o8 < CO / 8; ++o8) {
                    if (o8 * 8 >= cout) break;
                    float r[8];
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        const float4 b4 = ldg4(bias + o8 * 8 + hh * 4);
                        r[hh * 4 + 0] = acc[q][d][o8 * 8 + hh * 4 + 0] + b4.x;
                        r[hh * 4 + 1] = acc[q][d][o8 * 8 + hh * 4 + 1] + b4.y;
                        r[hh * 4 + 2] = acc[q][d][o8 * 8 + hh * 4 + 2] + b4.z;
                        r[hh * 4 + 3] = acc[q][d][o8 * 8 + hh * 4 + 3] + b4.w;
                    }
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        r[k] = fmaxf(r[k], 0.f) + ((o8 == 0 && k == 0) ? asel.slope0 : asel.slope) * fminf(r[k], 0.f);
                    stg8(op + d * ldo + o8 * 8, r);
                }
            } else if (vec && asel.simple) {
#pragma unroll
                for (int o4 = 0; o4 < CO / 4; ++o4) {
                    if (o4 * 4 >= cout) break;
                    const float4 b4 = ldg4(bias + o4 * 4);
                    float4 r = make_float4(acc[q][d][o4 * 4] + b4.x, acc[q][d][o4 * 4 + 1] + b4.y,
                                           acc[q][d][o4 * 4 + 2] + b4.z, acc[q][d][o4 * 4 + 3] + b4.w);
                    r.x = fmaxf(r.x, 0.f) + (o4 == 0 ? asel.slope0 : asel.slope) * fminf(r.x, 0.f);
                    r.y = fmaxf(r.y, 0.f) + asel.slope * fminf(r.y, 0.f);
                    r.z = fmaxf(r.z, 0.f) + asel.slope * fminf(r.z, 0.f);
                    r.w = fmaxf(r.w, 0.f) + asel.slope * fminf(r.w, 0.f);
                    *reinterpret_cast<float4*>(op + d * ldo + o4 * 4) = r;
                }
            } else {
#pragma unroll
                for (int i = 0; i < CO; ++i)
                    if (i < cout) op[d * ldo + i] = codd_act_apply(asel, acc[q][d][i] + __ldg(bias + i), i);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// layout helpers
// ---------------------------------------------------------------------------------------------
// NHWC -> planar through a shared-memory tile of 128 pixels x C channels: coalesced 128-bit
// reads along the channel-contiguous side, coalesced 128-byte row writes on the planar side.
constexpr int TR_PIX = 128;
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const float* __restrict__ in, int ldi, int hw, int c,
                                                           int ctot, int c0, float* __restrict__ out) {
    in += c0;   // this launch moves channels [c0, c0 + c) of ctot
    extern __shared__ float4 smem4[];
    float* tile = reinterpret_cast<float*>(smem4);   // [c][TR_PIX + 1]
    const int s = blockIdx.y;
    const int p0 = blockIdx.x * TR_PIX;
    const int np = min(TR_PIX, hw - p0);
    const bool vec = ((ldi & 3) == 0) && ((c & 3) == 0) && ((((uintptr_t)in) & 15u) == 0);
    if (vec) {
        const int c4n = c >> 2;
        for (int idx = threadIdx.x; idx < np * c4n; idx += blockDim.x) {
            const int pp = idx / c4n, c4 = idx - pp * c4n;
            const float4 v = ldg4(in + ((size_t)s * hw + p0 + pp) * ldi + c4 * 4);
            tile[(c4 * 4 + 0) * (TR_PIX + 1) + pp] = v.x;
            tile[(c4 * 4 + 1) * (TR_PIX + 1) + pp] = v.y;
            tile[(c4 * 4 + 2) * (TR_PIX + 1) + pp] = v.z;
            tile[(c4 * 4 + 3) * (TR_PIX + 1) + pp] = v.w;
        }
    } else {
        for (int idx = threadIdx.x; idx < np * c; idx += blockDim.x) {
            const int pp = idx / c, ch = idx - pp * c;
            tile[ch * (TR_PIX + 1) + pp] = __ldg(in + ((size_t)s * hw + p0 + pp) * ldi + ch);
        }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < c * TR_PIX; idx += blockDim.x) {
        const int ch = idx / TR_PIX, pp = idx - ch * TR_PIX;
        if (pp < np) out[((size_t)s * ctot + c0 + ch) * hw + p0 + pp] = tile[ch * (TR_PIX + 1) + pp];
    }
}

__global__ void nchw_to_nhwc_kernel(const float* __restrict__ in, int c, int hw, float* __restrict__ out, int ldo,
                                    size_t total) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // over n*hw*c, c fastest
    if (i >= total) return;
    const int ch = (int)(i % c);
    const size_t t = i / c;
    const size_t pix = t % hw;
    const size_t s = t / hw;
    out[(s * hw + pix) * ldo + ch] = __ldg(in + (s * c + ch) * hw + pix);
}

}  // namespace

namespace {
// geometry dispatch for one launch (p.Cout <= 64)
int conv_dispatch(const ConvP& p, const codd_conv_desc* d, cudaStream_t s) {
    const int kh = d->kh, kw = d->kw, sh = d->sh, sw = d->sw, dil = d->dil;
    if (kh == 1 && kw == 1 && sh == 1 && sw == 1 && d->ph == 0 && d->pw == 0 && d->ho == d->h && d->wo == d->w) {
        const bool vec_ok = p.vec0 && (d->c0 % 4 == 0) && (p.C1 == 0 || (p.vec1 && d->c1 % 4 == 0)) &&
                            codd_aligned16(p.out) && (d->ldo % 4 == 0) && (d->c0 + p.C1) * 32 * 4 <= 48 * 1024;
        if (vec_ok && p.Cout <= 16) return launch_pointwise<16, 4>(p, s);
        if (vec_ok && p.Cout <= 24) return launch_pointwise<24, 2>(p, s);
        if (vec_ok && p.Cout <= 32) return launch_pointwise<32, 2>(p, s);
    }
    if (kh == 1 && kw == 1 && sh == 1 && sw == 1) return dispatch_cout<1, 1, 1, 1, 1, true>(p, s);
    if (kh == 1 && kw == 1 && sh == 2 && sw == 2) return dispatch_cout<1, 1, 2, 2, 1, false>(p, s);
    if (kh == 3 && kw == 3 && sh == 1 && sw == 1 && dil == 1 && p.Cout == 1 && p.C0 == 16 && p.C1 == 0 && p.vec0 &&
        d->ph == 1 && d->pw == 1 && d->ho == d->h && d->wo == d->w && !p.res_after) {
        dim3 grid(codd_ceil_div(p.W, H1_TW), codd_ceil_div(p.H, H1_TH), p.N), block(H1_TW, H1_TH);
        conv3x3_head1_kernel<16><<<grid, block, 0, s>>>(p.in0, p.ld0, p.N, p.H, p.W, p.w, p.bias, p.res, p.ldr, p.act,
                                                        p.out, p.ldo);
        CODD_RETURN_IF_CUDA_ERROR();
        return 0;
    }
    if (kh == 3 && kw == 3 && sh == 1 && sw == 1 && dil == 1 && p.Cout == 2 && p.C0 == 32 && p.C1 == 0 && p.vec0 &&
        d->ph == 1 && d->pw == 1 && d->ho == d->h && d->wo == d->w && !p.res_after) {
        dim3 grid(codd_ceil_div(p.W, H2_TW), codd_ceil_div(p.H, H2_TH), p.N), block(H2_TW, H2_TH);
        conv3x3_head2_kernel<32><<<grid, block, 0, s>>>(p.in0, p.ld0, p.N, p.H, p.W, p.w, p.wld, p.bias, p.res, p.ldr,
                                                        p.res_bcast, p.act, p.out, p.ldo);
        CODD_RETURN_IF_CUDA_ERROR();
        return 0;
    }
    if (kh == 3 && kw == 3 && sh == 1 && sw == 1 && dil == 1) return dispatch_cout<3, 3, 1, 1, 1, true>(p, s);
    if (kh == 3 && kw == 3 && sh == 1 && sw == 1 && dil == 3) return dispatch_cout<3, 3, 1, 1, 3, false>(p, s);
    if (kh == 3 && kw == 3 && sh == 1 && sw == 1 && dil == 4) return dispatch_cout<3, 3, 1, 1, 4, false>(p, s);
    if (kh == 3 && kw == 3 && sh == 2 && sw == 2 && dil == 1) return dispatch_cout<3, 3, 2, 2, 1, false>(p, s);
    if (kh == 4 && kw == 4 && sh == 2 && sw == 2 && dil == 1) {
        // the stride-2 stage is wide (66 input columns); split the output channels over warps so that
        // a full 256-thread CTA shares it
        if (p.Cout <= 16) return launch_conv<4, 4, 2, 2, 1, 8>(p, 4, 2, s);
        if (p.Cout <= 24) return launch_conv<4, 4, 2, 2, 1, 8>(p, 2, 3, s);
        return dispatch_cout<4, 4, 2, 2, 1, false>(p, s);
    }
    if (kh == 4 && kw == 4 && sh == 4 && sw == 4 && dil == 1) return dispatch_cout<4, 4, 4, 4, 1, false>(p, s);
    if (kh == 4 && kw == 4 && sh == 4 && sw == 1 && dil == 1) return dispatch_cout<4, 4, 4, 1, 1, false>(p, s);
    if (kh == 7 && kw == 7 && sh == 1 && sw == 1 && dil == 1) return dispatch_cout<7, 7, 1, 1, 1, false>(p, s);
    if (kh == 7 && kw == 7 && sh == 2 && sw == 2 && dil == 1) return dispatch_cout<7, 7, 2, 2, 1, false>(p, s);
    return CODD_E_UNSUPPORTED;
}
}  // namespace

extern "C" int codd_conv2d_nhwc(const codd_conv_desc* d, const float* in0, const float* in1, const float* weight,
                                const float* bias, const float* residual, float* out, void* stream) {
    if (!d || !in0 || !weight || !out) return CODD_E_BADARG;
    if (d->n <= 0 || d->h <= 0 || d->w <= 0 || d->c0 <= 0 || d->cout <= 0 || d->ho <= 0 || d->wo <= 0)
        return CODD_E_BADARG;
    if (d->c1 > 0 && !in1) return CODD_E_BADARG;
    if (d->ld0 < d->c0 || (d->c1 > 0 && d->ld1 < d->c1) || d->ldo < d->cout) return CODD_E_SHAPE;
    if (residual && d->ldr < (d->res_bcast ? 1 : d->cout)) return CODD_E_SHAPE;
    // implied bottom/right extent must be consistent with a zero-padded convolution
    if ((d->ho - 1) * d->sh - d->ph + (d->kh - 1) * d->dil < 0) return CODD_E_SHAPE;
    ConvP p;
    p.in0 = in0;
    p.in1 = d->c1 > 0 ? in1 : nullptr;
    p.N = d->n; p.H = d->h; p.W = d->w;
    p.C0 = d->c0; p.ld0 = d->ld0;
    p.C1 = d->c1 > 0 ? d->c1 : 0; p.ld1 = d->ld1;
    p.ldo = d->ldo; p.wld = d->cout;
    p.ph = d->ph; p.pw = d->pw; p.Ho = d->ho; p.Wo = d->wo;
    p.ldr = d->ldr; p.res_bcast = d->res_bcast; p.res_after = d->res_after_act;
    p.tilesX = p.tilesY = 0;
    p.vec0 = codd_aligned16(in0) && (d->ld0 % 4 == 0);
    p.vec1 = p.in1 && codd_aligned16(in1) && (d->ld1 % 4 == 0) && (d->c0 % 4 == 0);
    cudaStream_t s = (cudaStream_t)stream;
    // wide layers (the RAFT3D encoders / update block, Cout up to 1024) run as 64-filter chunks:
    // each launch sees a column slice of the packed weight (row stride wld) and of bias / residual / out
    constexpr int CHUNK = 64;
    for (int co0 = 0; co0 < d->cout; co0 += CHUNK) {
        p.Cout = d->cout - co0 < CHUNK ? d->cout - co0 : CHUNK;
        p.w = weight + co0;
        p.bias = bias ? bias + co0 : nullptr;
        p.res = residual ? (d->res_bcast ? residual : residual + co0) : nullptr;
        p.out = out + co0;
        p.act = (co0 > 0 && d->act == CODD_ACT_RELU_CH0) ? CODD_ACT_NONE : d->act;
        const int rc = conv_dispatch(p, d, s);
        if (rc != 0) return rc;
    }
    return 0;
}

namespace {
// TMA-staged variant of conv3x3_image_kernel: ncu of the kernel above shows the L1 data pipe 77 % busy with 61 M
// sectors of scalar input loads for a 106 MB image pair (six 4-byte loads per row and channel, lanes 16 bytes apart).
// Here ONE bulk tensor load brings the CTA's (8 + 2) x (256 + 8) x 3 input tile (zero-filled outside the image = the
// padding) into shared memory; a thread reads its four pixels as one 128-bit shared load plus the two neighbours.
// Weights and bias arrive by bulk copies on the same mbarrier.  Arithmetic and its order are unchanged.
// (a TMA box is at most 256 elements wide: each of the CTA's two warps gets its own 128 + 8 column tile)
constexpr int IT_THREADS = 64, IT_PX = 4, IT_TW = IT_THREADS * IT_PX, IT_BOXW = 32 * IT_PX + 8, IT_BOXH = IMG_ROWS + 2;
constexpr int IT_TILE = (3 * IT_BOXH * IT_BOXW + 31) & ~31;      // floats per warp tile, padded to 128 bytes (TMA destination)
__global__ void __launch_bounds__(IT_THREADS) conv3x3_image_tma_kernel(const __grid_constant__ CUtensorMap lmap,
                                                                      const __grid_constant__ CUtensorMap rmap, int n, int h,
                                                                      int w, const float* __restrict__ wgt,
                                                                      const float* __restrict__ bias, float* __restrict__ out,
                                                                      int ldo) {
    constexpr int CO = 16, PXI = IT_PX;
    extern __shared__ uint8_t it_raw[];
    __shared__ __align__(8) unsigned long long bar;
    const uint32_t sbase = (s_u32(it_raw) + 127u) & ~127u;
    float* tile = reinterpret_cast<float*>(it_raw + (sbase - s_u32(it_raw)));     // [2 warps][3][IT_BOXH][IT_BOXW]
    float* s_w = tile + 2 * IT_TILE;                                               // [27][CO]
    float* s_b = s_w + 27 * CO;
    const int tid = threadIdx.x;
    const int xc0 = blockIdx.x * IT_TW, y0 = blockIdx.y * IMG_ROWS, s = blockIdx.z;
    const uint32_t b = s_u32(&bar);
    if (tid == 0) {
        mbar_init(b, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(b, (uint32_t)((2 * 3 * IT_BOXH * IT_BOXW + 27 * CO + CO) * 4));
#pragma unroll
        for (int wv = 0; wv < 2; ++wv)
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                         ::"r"(sbase + (uint32_t)wv * (IT_TILE * 4)), "l"(s < n ? &lmap : &rmap), "r"(b),
                           "r"(xc0 + wv * 32 * PXI - 4), "r"(y0 - 1), "r"(0), "r"(s < n ? s : s - n)
                         : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(s_u32(s_w)), "l"(wgt), "r"(27 * CO * 4), "r"(b) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(s_u32(s_b)), "l"(bias), "r"(CO * 4), "r"(b) : "memory");
    }
    __syncthreads();
    mbar_wait(b, 0);
    const int x0 = xc0 + tid * PXI;
    if (x0 >= w) return;
    const int y_end = min(h, y0 + IMG_ROWS);
#pragma unroll 1
    for (int y = y0; y < y_end; ++y) {
        __align__(8) float acc[PXI][CO];
#pragma unroll
        for (int q = 0; q < PXI; ++q)
#pragma unroll
            for (int i = 0; i < CO; ++i) acc[q][i] = s_b[i];
#pragma unroll 1
        for (int ky = 0; ky < 3; ++ky) {
            const int yy = y + ky - 1;
            if (yy < 0 || yy >= h) continue;       // (zero rows: skipped as in the kernel above, same FMA sequence)
#pragma unroll
            for (int ci = 0; ci < 3; ++ci) {
                const float* rp = tile + (tid >> 5) * IT_TILE + (ci * IT_BOXH + (y - y0 + ky)) * IT_BOXW + 4 + (tid & 31) * PXI;   // -> pixel x0
                const float4 m = *reinterpret_cast<const float4*>(rp);
                const float a[PXI + 2] = {rp[-1], m.x, m.y, m.z, m.w, rp[4]};
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const float* wp = s_w + ((ky * 3 + kx) * 3 + ci) * CO;
#pragma unroll
                    for (int o4 = 0; o4 < CO / 4; ++o4) {
                        const float4 wv = *reinterpret_cast<const float4*>(wp + o4 * 4);
#pragma unroll
                        for (int q = 0; q < PXI; ++q) fma4(&acc[q][o4 * 4], a[q + kx], wv);
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < PXI; ++q) {
            if (x0 + q >= w) break;
            float* op = out + (((size_t)s * h + y) * w + x0 + q) * ldo;
#pragma unroll
            for (int i = 0; i < CO; ++i) acc[q][i] = fmaxf(acc[q][i], 0.f) + CODD_LEAKY_SLOPE * fminf(acc[q][i], 0.f);
            stg8(op, &acc[q][0]);
            stg8(op + 8, &acc[q][8]);
        }
    }
}

int conv3x3_image_tma(const float* left, const float* right, int n, int h, int w, const float* weight, const float* bias,
                      float* out, int ldo, cudaStream_t s) {
    PFN_tmapEncodeTiled enc = rg_get_encode();
    if (!enc) return CODD_E_UNSUPPORTED;
    auto make = [&](CUtensorMap* m, const float* base) {
        const cuuint64_t dim[4] = {(cuuint64_t)w, (cuuint64_t)h, 3u, (cuuint64_t)n};
        const cuuint64_t str[3] = {(cuuint64_t)w * 4, (cuuint64_t)h * w * 4, (cuuint64_t)3 * h * w * 4};
        const cuuint32_t box[4] = {(cuuint32_t)IT_BOXW, (cuuint32_t)IT_BOXH, 3u, 1u}, es[4] = {1u, 1u, 1u, 1u};
        return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)base, dim, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    };
    CUtensorMap lmap, rmap;
    if (!make(&lmap, left) || !make(&rmap, right ? right : left)) return CODD_E_UNSUPPORTED;
    const size_t smem = (size_t)(2 * IT_TILE + 27 * 16 + 16) * 4 + 128;
    dim3 grid(codd_ceil_div(w, IT_TW), codd_ceil_div(h, IMG_ROWS), right ? 2 * n : n);
    conv3x3_image_tma_kernel<<<grid, IT_THREADS, smem, s>>>(lmap, rmap, n, h, w, weight, bias, out, ldo);
    CODD_RETURN_IF_CUDA_ERROR();
    return 0;
}
}  // namespace

extern "C" int codd_conv3x3_image(const float* left, const float* right, int n, int h, int w, const float* weight,
                                  const float* bias, int cout, float* out, int ldo, void* stream) {
    if (!left || !weight || !bias || !out || n <= 0 || h <= 0 || w <= 0) return CODD_E_BADARG;
    if (cout <= 0 || cout > 16 || ldo < cout) return CODD_E_SHAPE;
    if (cout == 16 && w % 4 == 0 && w >= 64 && ldo % 8 == 0 && codd_aligned32(out) && codd_aligned16(left) &&
        (!right || codd_aligned16(right)) && codd_aligned16(weight) && codd_aligned16(bias)) {
        const int rc = conv3x3_image_tma(left, right, n, h, w, weight, bias, out, ldo, (cudaStream_t)stream);
        if (rc != CODD_E_UNSUPPORTED) return rc;
    }
    dim3 block(64);
    dim3 grid(codd_ceil_div(w, 64 * 4), codd_ceil_div(h, IMG_ROWS), right ? 2 * n : n);
    conv3x3_image_kernel<16><<<grid, block, 0, (cudaStream_t)stream>>>(left, right ? right : left, n, h, w, weight,
                                                                        bias, cout, out, ldo);
    CODD_RETURN_IF_CUDA_ERROR();
    return 0;
}

extern "C" int codd_deconv2x2_nhwc(const float* in, int ldi, int n, int h, int w, int cin, const float* weight,
                                   const float* bias, int cout, float* out, int ldo, int act, void* stream) {
    if (!in || !weight || !bias || !out || n <= 0 || h <= 0 || w <= 0 || cin <= 0 || cout <= 0) return CODD_E_BADARG;
    if (cin % 4 != 0 || ldi % 4 != 0 || ldi < cin || ldo < cout || cout > 32) return CODD_E_SHAPE;
    if (!codd_aligned16(in)) return CODD_E_ALIGN;
    dim3 block(128);
    dim3 grid(codd_ceil_div(w, 128 * DC_PX), 2 * h, n);
    cudaStream_t s = (cudaStream_t)stream;
    if (cout <= 16) {
        deconv2x2_kernel<16><<<grid, block, 4 * cin * 16 * sizeof(float), s>>>(in, ldi, n, h, w, cin, weight, bias,
                                                                                 cout, out, ldo, act);
    } else if (cout <= 24) {
        deconv2x2_kernel<24><<<grid, block, 4 * cin * 24 * sizeof(float), s>>>(in, ldi, n, h, w, cin, weight, bias,
                                                                                 cout, out, ldo, act);
    } else {
        deconv2x2_kernel<32><<<grid, block, 4 * cin * 32 * sizeof(float), s>>>(in, ldi, n, h, w, cin, weight, bias,
                                                                                 cout, out, ldo, act);
    }
    CODD_RETURN_IF_CUDA_ERROR();
    return 0;
}

extern "C" int codd_nhwc_to_nchw(const float* in, int ldi, int n, int h, int w, int c, float* out, void* stream) {
    if (!in || !out || n <= 0 || h <= 0 || w <= 0 || c <= 0 || ldi < c) return CODD_E_BADARG;
    const int hw = h * w;
    constexpr int CCH = 256;   // channels per launch (shared tile = CCH x 129 floats)
    dim3 grid((unsigned)codd_ceil_div(hw, TR_PIX), (unsigned)n);
    for (int c0 = 0; c0 < c; c0 += CCH) {
        const int cc = c - c0 < CCH ? c - c0 : CCH;
        const size_t smem = (size_t)cc * (TR_PIX + 1) * sizeof(float);
        static CoddDeviceOnce once;
        if (int rc = codd_once_per_device(once, [&] {
                return codd_max_dynamic_smem(nhwc_to_nchw_kernel);
            }))
            return rc;
        nhwc_to_nchw_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(in, ldi, hw, cc, c, c0, out);
        CODD_RETURN_IF_CUDA_ERROR();
    }
    return 0;
}

extern "C" int codd_nchw_to_nhwc(const float* in, int n, int c, int h, int w, float* out, int ldo, void* stream) {
    if (!in || !out || n <= 0 || h <= 0 || w <= 0 || c <= 0 || ldo < c) return CODD_E_BADARG;
    const size_t total = (size_t)n * c * h * w;
    nchw_to_nhwc_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, c, h * w, out, ldo,
                                                                                            total);
    CODD_RETURN_IF_CUDA_ERROR();
    return 0;
}
