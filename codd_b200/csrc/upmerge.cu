// Fused conv_up + first layer of conv_merge of HITUNet (model/stereo/hitnet/backbone.py:17-32,75-88):
//
//   up  = LeakyReLU(ConvTranspose2d(k=2, s=2)(coarse) + b_up)                 (conv_up,  backbone.py:17-21)
//   out = LeakyReLU(Conv2d 1x1(cat(skip, up)) + b_m)                          (conv_merge[0], backbone.py:24-32,76)
//
// A 2x2 / stride-2 transposed convolution has no overlap between taps — output pixel (y, x) is a Cc x Cu GEMV of its
// parent pixel (y/2, x/2) with the weight slice of its position (y&1, x&1) — and the 1x1 conv that follows is another
// per-pixel GEMV, so the up-sampled tensor never has to exist: per output pixel this kernel reads the skip pixel and the
// parent pixel, keeps `up` in registers, and writes the merged pixel.  As two launches the pair moved the up-sampled
// tensor through HBM twice (1.1 GB per step at the finest level: 0.31 ms deconvolution + 0.27 ms 1x1).
// One thread = NPAR parent pixels and their 2 x 2 x NPAR output pixels (see upmerge_kernel).  Packed FFMA2, weights in
// shared memory.
#include <cuda.h>

#include "common.cuh"
#include "ring_util.cuh"

namespace {

struct UmP {
    const float* coarse; int ldc, Cc;     // [N, H/2, W/2, Cc] NHWC
    const float* skip;   int lds, Cs;     // [N, H, W, Cs] NHWC
    const float* w_up;                    // packed [4 positions dy*2+dx][Cc][Cu]
    const float* b_up;                    // [Cu]
    const float* w_m;                     // packed [Cs + Cu][Co]   (1x1: cat order = skip first)
    const float* b_m;                     // [Co]
    float* out; int ldo;                  // [N, H, W, Co] NHWC
    int N, H, W;
    int v8;                               // 256-bit output stores allowed
    int w_bulk;                           // weights / biases are 16-byte aligned: staged by 1-D bulk copies (TMA variant)
};

constexpr int UM_THREADS = 128;

// 8 channels of one pixel: one 256-bit load when rows and base are 32-byte aligned (V8), else two 128-bit loads
template <bool V8>
__device__ __forceinline__ void um_load8(const float* p, float* v) {
    if (V8) {
        ldg8(p, v);
    } else {
        const float4 a = ldg4(p), b = ldg4(p + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
}

// Second decomposition (round 2): PARENT-centric.  The first one gave a thread two output pixels of one row, 256 columns
// apart; adjacent lanes then had different column parities, i.e. two weight slices per warp-wide shared load, and only
// two pixels per weight broadcast: ncu showed the L1 / shared data pipe 85 % busy (81 M wavefronts, 20 M of them bank
// conflicts) with the fp32 pipe at 21 %.  Now a thread owns NPAR parent pixels (128 columns apart) and produces their
// 2 x NPAR output pixels of row 2*yp + dy, then the same for the other dy: both column parities live in one thread, so
// every deconvolution weight load (two slices, warp-uniform addresses) feeds 2*NPAR pixels and every merge weight load
// 2*NPAR pixels — 256 instead of ~450 shared wavefront pairs per four pixels — and a parent is fetched once per CTA row
// pair instead of four times.
template <int CU, int CO, int NPAR, bool V8>
__global__ void __launch_bounds__(UM_THREADS) upmerge_kernel(UmP p) {
    constexpr int NPX = 2 * NPAR;
    extern __shared__ float4 um_smem4[];
    float* s_wu = reinterpret_cast<float*>(um_smem4);        // [4][Cc][CU]
    float* s_wm = s_wu + 4 * p.Cc * CU;                      // [Cs + CU][CO]
    float* s_bu = s_wm + (p.Cs + CU) * CO;                   // [CU]
    float* s_bm = s_bu + CU;                                 // [CO]
    if (p.w_bulk) {
        // weights and biases by four 1-D bulk copies: staged by the threads (12-24 scalar loads each, per CTA) they were a
        // fifth of this kernel's time
        __shared__ __align__(8) unsigned long long wbar;
        const uint32_t b = s_u32(&wbar);
        if (threadIdx.x == 0) {
            mbar_init(b, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            const uint32_t nwu = 4 * p.Cc * CU * 4, nwm = (p.Cs + CU) * CO * 4;
            mbar_expect_tx(b, nwu + nwm + (CU + CO) * 4);
            auto bulk = [&](const float* dst, const float* src, uint32_t bytes) {
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(s_u32(dst)), "l"(src), "r"(bytes), "r"(b) : "memory");
            };
            bulk(s_wu, p.w_up, nwu);
            bulk(s_wm, p.w_m, nwm);
            bulk(s_bu, p.b_up, CU * 4);
            bulk(s_bm, p.b_m, CO * 4);
        }
        __syncthreads();
        mbar_wait(b, 0);
    } else {
        for (int i = threadIdx.x; i < 4 * p.Cc * CU; i += UM_THREADS) s_wu[i] = __ldg(p.w_up + i);
        for (int i = threadIdx.x; i < (p.Cs + CU) * CO; i += UM_THREADS) s_wm[i] = __ldg(p.w_m + i);
        for (int i = threadIdx.x; i < CU; i += UM_THREADS) s_bu[i] = __ldg(p.b_up + i);
        for (int i = threadIdx.x; i < CO; i += UM_THREADS) s_bm[i] = __ldg(p.b_m + i);
        __syncthreads();
    }

    const int Hc = p.H >> 1, Wc = p.W >> 1;
    const int yp = blockIdx.y, n = blockIdx.z;
    const int xp0 = blockIdx.x * (UM_THREADS * NPAR) + threadIdx.x;
    if (xp0 >= Wc) return;
    int xp[NPAR];
    bool ok[NPAR];
    const float* cp[NPAR];
#pragma unroll
    for (int q = 0; q < NPAR; ++q) {
        ok[q] = xp0 + q * UM_THREADS < Wc;
        xp[q] = ok[q] ? xp0 + q * UM_THREADS : xp0;          // (a clamped duplicate: computed, not stored)
        cp[q] = p.coarse + (((size_t)n * Hc + yp) * Wc + xp[q]) * p.ldc;
    }

#pragma unroll 1
    for (int dy = 0; dy < 2; ++dy) {
        const int y = 2 * yp + dy;
        // ---- up[2q + dx] = LeakyReLU(W_up[dy][dx]^T * parent_q + b_up)
        __align__(8) float up[NPX][CU];
#pragma unroll
        for (int j = 0; j < NPX; ++j)
#pragma unroll
            for (int c = 0; c < CU; ++c) up[j][c] = s_bu[c];
        {
            const float* w0 = s_wu + (dy * 2 + 0) * p.Cc * CU;
            const float* w1 = s_wu + (dy * 2 + 1) * p.Cc * CU;
            for (int c = 0; c < p.Cc; c += 8) {             // channel counts are multiples of 8 (checked by the entry point)
                float a[NPAR][8];
#pragma unroll
                for (int q = 0; q < NPAR; ++q) um_load8<V8>(cp[q] + c, a[q]);
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) {
#pragma unroll
                    for (int o4 = 0; o4 < CU / 4; ++o4) {
                        const float4 u = *reinterpret_cast<const float4*>(w0 + (c + cc) * CU + o4 * 4);
                        const float4 v = *reinterpret_cast<const float4*>(w1 + (c + cc) * CU + o4 * 4);
#pragma unroll
                        for (int q = 0; q < NPAR; ++q) {
                            fma4(&up[2 * q][o4 * 4], a[q][cc], u);
                            fma4(&up[2 * q + 1][o4 * 4], a[q][cc], v);
                        }
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < NPX; ++j)
#pragma unroll
                for (int c = 0; c < CU; ++c) up[j][c] = fmaxf(up[j][c], 0.f) + CODD_LEAKY_SLOPE * fminf(up[j][c], 0.f);
        }

        // ---- out = LeakyReLU(W_m^T * cat(skip, up) + b_m)
        __align__(8) float acc[NPX][CO];
#pragma unroll
        for (int j = 0; j < NPX; ++j)
#pragma unroll
            for (int c = 0; c < CO; ++c) acc[j][c] = s_bm[c];
        const float* sp[NPAR];                               // skip pixels 2*xp, 2*xp + 1 are adjacent: sp[q] + dx * lds
#pragma unroll
        for (int q = 0; q < NPAR; ++q) sp[q] = p.skip + (((size_t)n * p.H + y) * p.W + 2 * xp[q]) * p.lds;
        for (int c = 0; c < p.Cs; c += 8) {
            float a[NPX][8];
#pragma unroll
            for (int j = 0; j < NPX; ++j) um_load8<V8>(sp[j >> 1] + (j & 1) * p.lds + c, a[j]);
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) {
                const float* wp = s_wm + (c + cc) * CO;
#pragma unroll
                for (int o4 = 0; o4 < CO / 4; ++o4) {
                    const float4 wv = *reinterpret_cast<const float4*>(wp + o4 * 4);
#pragma unroll
                    for (int j = 0; j < NPX; ++j) fma4(&acc[j][o4 * 4], a[j][cc], wv);
                }
            }
        }
        const float* wmu = s_wm + p.Cs * CO;
#pragma unroll
        for (int cu = 0; cu < CU; ++cu) {
            const float* wp = wmu + cu * CO;
#pragma unroll
            for (int o4 = 0; o4 < CO / 4; ++o4) {
                const float4 wv = *reinterpret_cast<const float4*>(wp + o4 * 4);
#pragma unroll
                for (int j = 0; j < NPX; ++j) fma4(&acc[j][o4 * 4], up[j][cu], wv);
            }
        }
#pragma unroll
        for (int j = 0; j < NPX; ++j) {
            if (!ok[j >> 1]) continue;
            float* op = p.out + (((size_t)n * p.H + y) * p.W + 2 * xp[j >> 1] + (j & 1)) * p.ldo;
#pragma unroll
            for (int c = 0; c < CO; ++c) acc[j][c] = fmaxf(acc[j][c], 0.f) + CODD_LEAKY_SLOPE * fminf(acc[j][c], 0.f);
            if (p.v8) {   // whole 32-byte sectors per thread (see stg8)
#pragma unroll
                for (int o8 = 0; o8 < CO / 8; ++o8) stg8(op + o8 * 8, &acc[j][o8 * 8]);
            } else {
#pragma unroll
                for (int o4 = 0; o4 < CO / 4; ++o4)
                    *reinterpret_cast<float4*>(op + o4 * 4) =
                        make_float4(acc[j][o4 * 4], acc[j][o4 * 4 + 1], acc[j][o4 * 4 + 2], acc[j][o4 * 4 + 3]);
            }
        }
    }
}

// Third decomposition, for the all-16-channel case (the finest level, 60 % of this op's time): activations staged by TMA.
// ncu of the kernel above: L1 data pipe 74 % busy, 56 M of its 96 M wavefronts from the lane-strided global loads and
// stores (~1 wavefront per 32-byte sector).  Here a CTA of 64 threads owns 128 parents of one parent row: three
// cp.async.bulk.tensor loads bring the parent pixels (8 KB) and the two skip rows (2 x 256 pixels, 32 KB) into SWIZZLE_64B
// tiles while the threads stage the weights; every thread computes its two parents' 2 x 2 outputs for both rows as above,
// reading activations with 128-bit shared loads, overwrites its own skip pixels with the results, and two bulk tensor
// stores write the rows back (columns >= W clipped by the hardware).  Same FMA order as above: bit-identical results.
constexpr int UT_THREADS = 64, UT_PAR = 128, UT_PX = 2 * UT_PAR;
// CC = coarse channels: 16 (SWIZZLE_64B tile, as the skip rows) or 24 (96-byte pixel rows, unswizzled: the 128-bit reads of
// neighbouring lanes then conflict two ways, on 6 of ~280 shared loads per thread and row)
template <int CC>
__global__ void __launch_bounds__(UT_THREADS) upmerge_tma_kernel(const __grid_constant__ CUtensorMap cmap,
                                                                const __grid_constant__ CUtensorMap smap,
                                                                const __grid_constant__ CUtensorMap omap, UmP p) {
    constexpr int C = 16;
    extern __shared__ uint8_t ut_raw[];
    __shared__ __align__(8) unsigned long long bar;
    const uint32_t sbase = (s_u32(ut_raw) + 1023u) & ~1023u;
    uint8_t* gbase = ut_raw + (sbase - s_u32(ut_raw));
    // [coarse 128 px x CC*4 B | skip row 0: 256 px x 64 B | skip row 1 | weights]
    constexpr uint32_t OFF_S0 = (UT_PAR * CC * 4 + 1023u) & ~1023u, OFF_S1 = OFF_S0 + UT_PX * 64, OFF_W = OFF_S1 + UT_PX * 64;
    float* s_wu = reinterpret_cast<float*>(gbase + OFF_W);   // [4][CC][C]
    float* s_wm = s_wu + 4 * CC * C;                         // [2C][C]
    float* s_bu = s_wm + 2 * C * C;
    float* s_bm = s_bu + C;
    const int tid = threadIdx.x;
    const int xp0 = blockIdx.x * UT_PAR, yp = blockIdx.y, n = blockIdx.z;
    const uint32_t b = s_u32(&bar);
    if (tid == 0) {
        mbar_init(b, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(b, (uint32_t)(UT_PAR * CC * 4 + 2 * UT_PX * 64 + (p.w_bulk ? (4 * CC * C + 2 * C * C + 2 * C) * 4 : 0)));
        if (p.w_bulk) {       // weights and biases as four 1-D bulk copies (16-byte aligned sources): no thread touches them
            auto bulk = [&](const float* dst, const float* src, uint32_t bytes) {
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(s_u32(dst)), "l"(src), "r"(bytes), "r"(b) : "memory");
            };
            bulk(s_wu, p.w_up, 4 * CC * C * 4);
            bulk(s_wm, p.w_m, 2 * C * C * 4);
            bulk(s_bu, p.b_up, C * 4);
            bulk(s_bm, p.b_m, C * 4);
        }
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                     ::"r"(sbase), "l"(&cmap), "r"(b), "r"(0), "r"(xp0), "r"(yp), "r"(n) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                     ::"r"(sbase + OFF_S0), "l"(&smap), "r"(b), "r"(0), "r"(2 * xp0), "r"(2 * yp), "r"(n) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                     ::"r"(sbase + OFF_S1), "l"(&smap), "r"(b), "r"(0), "r"(2 * xp0), "r"(2 * yp + 1), "r"(n) : "memory");
    }
    if (!p.w_bulk) {
        for (int i = tid; i < 4 * CC * C; i += UT_THREADS) s_wu[i] = __ldg(p.w_up + i);
        for (int i = tid; i < 2 * C * C; i += UT_THREADS) s_wm[i] = __ldg(p.w_m + i);
        if (tid < C) { s_bu[tid] = __ldg(p.b_up + tid); s_bm[tid] = __ldg(p.b_m + tid); }
    }
    __syncthreads();
    mbar_wait(b, 0);

    constexpr int NPAR = 2, NPX = 4;
    int lp[NPAR];                                             // local parent index
#pragma unroll
    for (int q = 0; q < NPAR; ++q) lp[q] = tid + q * UT_THREADS;
#pragma unroll 1
    for (int dy = 0; dy < 2; ++dy) {
        __align__(8) float up[NPX][C];
#pragma unroll
        for (int j = 0; j < NPX; ++j)
#pragma unroll
            for (int c = 0; c < C; ++c) up[j][c] = s_bu[c];
        {
            const float* w0 = s_wu + (dy * 2 + 0) * CC * C;
            const float* w1 = s_wu + (dy * 2 + 1) * CC * C;
#pragma unroll
            for (int c4 = 0; c4 < CC / 4; ++c4) {
                float4 a[NPAR];
#pragma unroll
                for (int q = 0; q < NPAR; ++q)
                    a[q] = *reinterpret_cast<const float4*>(gbase + (CC == 16 ? swz_off<16>(lp[q], c4)
                                                                             : (uint32_t)(lp[q] * CC * 4 + c4 * 16)));
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
#pragma unroll
                    for (int o4 = 0; o4 < C / 4; ++o4) {
                        const float4 u = *reinterpret_cast<const float4*>(w0 + (c4 * 4 + cc) * C + o4 * 4);
                        const float4 v = *reinterpret_cast<const float4*>(w1 + (c4 * 4 + cc) * C + o4 * 4);
#pragma unroll
                        for (int q = 0; q < NPAR; ++q) {
                            const float av = cc == 0 ? a[q].x : cc == 1 ? a[q].y : cc == 2 ? a[q].z : a[q].w;
                            fma4(&up[2 * q][o4 * 4], av, u);
                            fma4(&up[2 * q + 1][o4 * 4], av, v);
                        }
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < NPX; ++j)
#pragma unroll
                for (int c = 0; c < C; ++c) up[j][c] = fmaxf(up[j][c], 0.f) + CODD_LEAKY_SLOPE * fminf(up[j][c], 0.f);
        }
        __align__(8) float acc[NPX][C];
#pragma unroll
        for (int j = 0; j < NPX; ++j)
#pragma unroll
            for (int c = 0; c < C; ++c) acc[j][c] = s_bm[c];
        uint8_t* srow = gbase + (dy ? OFF_S1 : OFF_S0);
#pragma unroll
        for (int c4 = 0; c4 < C / 4; ++c4) {
            float4 a[NPX];
#pragma unroll
            for (int j = 0; j < NPX; ++j)
                a[j] = *reinterpret_cast<const float4*>(srow + swz_off<16>(2 * lp[j >> 1] + (j & 1), c4));
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                const float* wp = s_wm + (c4 * 4 + cc) * C;
#pragma unroll
                for (int o4 = 0; o4 < C / 4; ++o4) {
                    const float4 wv = *reinterpret_cast<const float4*>(wp + o4 * 4);
#pragma unroll
                    for (int j = 0; j < NPX; ++j) {
                        const float av = cc == 0 ? a[j].x : cc == 1 ? a[j].y : cc == 2 ? a[j].z : a[j].w;
                        fma4(&acc[j][o4 * 4], av, wv);
                    }
                }
            }
        }
        const float* wmu = s_wm + C * C;
#pragma unroll
        for (int cu = 0; cu < C; ++cu) {
            const float* wp = wmu + cu * C;
#pragma unroll
            for (int o4 = 0; o4 < C / 4; ++o4) {
                const float4 wv = *reinterpret_cast<const float4*>(wp + o4 * 4);
#pragma unroll
                for (int j = 0; j < NPX; ++j) fma4(&acc[j][o4 * 4], up[j][cu], wv);
            }
        }
        // results overwrite this thread's own skip pixels (read above, by this thread only)
#pragma unroll
        for (int j = 0; j < NPX; ++j) {
#pragma unroll
            for (int c = 0; c < C; ++c) acc[j][c] = fmaxf(acc[j][c], 0.f) + CODD_LEAKY_SLOPE * fminf(acc[j][c], 0.f);
#pragma unroll
            for (int c4 = 0; c4 < C / 4; ++c4)
                *reinterpret_cast<float4*>(srow + swz_off<16>(2 * lp[j >> 1] + (j & 1), c4)) =
                    make_float4(acc[j][c4 * 4], acc[j][c4 * 4 + 1], acc[j][c4 * 4 + 2], acc[j][c4 * 4 + 3]);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                     ::"l"(&omap), "r"(sbase + OFF_S0), "r"(0), "r"(2 * xp0), "r"(2 * yp), "r"(n) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                     ::"l"(&omap), "r"(sbase + OFF_S1), "r"(0), "r"(2 * xp0), "r"(2 * yp + 1), "r"(n) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

template <int CC>
int um_launch_tma(const UmP& p, cudaStream_t s) {
    PFN_tmapEncodeTiled enc = rg_get_encode();
    if (!enc) return CODD_E_UNSUPPORTED;
    const int Hc = p.H / 2, Wc = p.W / 2;
    auto make = [&](CUtensorMap* m, const float* base, int c, int ld, int w, int h, int boxw) {
        const cuuint64_t dim[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)p.N};
        const cuuint64_t str[3] = {(cuuint64_t)ld * 4, (cuuint64_t)w * ld * 4, (cuuint64_t)h * w * ld * 4};
        const cuuint32_t box[4] = {(cuuint32_t)c, (cuuint32_t)boxw, 1u, 1u}, es[4] = {1u, 1u, 1u, 1u};
        return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)base, dim, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   c == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    };
    CUtensorMap cmap, smap, omap;
    if (!make(&cmap, p.coarse, CC, p.ldc, Wc, Hc, UT_PAR) || !make(&smap, p.skip, 16, p.lds, p.W, p.H, UT_PX) ||
        !make(&omap, p.out, 16, p.ldo, p.W, p.H, UT_PX))
        return CODD_E_UNSUPPORTED;
    const size_t smem = (((size_t)UT_PAR * CC * 4 + 1023) & ~(size_t)1023) + 2 * UT_PX * 64 +
                        (4 * CC * 16 + 2 * 256 + 32) * sizeof(float) + 1024;
    static CoddDeviceOnce once;
    if (int rc = codd_once_per_device(once, [&] {
            return cudaFuncSetAttribute(upmerge_tma_kernel<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        }))
        return rc;
    dim3 grid((unsigned)codd_ceil_div(Wc, UT_PAR), (unsigned)Hc, (unsigned)p.N);
    upmerge_tma_kernel<CC><<<grid, UT_THREADS, smem, s>>>(cmap, smap, omap, p);
    CODD_RETURN_IF_CUDA_ERROR();
    return 0;
}

template <int CU, int CO, int NPAR>
int um_launch(const UmP& p, cudaStream_t s) {
    const size_t smem = ((size_t)4 * p.Cc * CU + (size_t)(p.Cs + CU) * CO + CU + CO) * sizeof(float);
    if (smem > 48 * 1024) return CODD_E_UNSUPPORTED;
    dim3 grid((unsigned)codd_ceil_div(p.W / 2, UM_THREADS * NPAR), (unsigned)(p.H / 2), (unsigned)p.N);
    const bool v8in = (p.ldc % 8 == 0) && (p.lds % 8 == 0) && codd_aligned32(p.coarse) && codd_aligned32(p.skip);
    if (v8in) upmerge_kernel<CU, CO, NPAR, true><<<grid, UM_THREADS, smem, s>>>(p);
    else upmerge_kernel<CU, CO, NPAR, false><<<grid, UM_THREADS, smem, s>>>(p);
    CODD_RETURN_IF_CUDA_ERROR();
    return 0;
}

}  // namespace

extern "C" int codd_upmerge_nhwc(const float* coarse, int ldc, int cc, const float* skip, int lds, int cs,
                                 const float* w_up, const float* b_up, int cu, const float* w_merge, const float* b_merge,
                                 int co, int n, int h, int w, float* out, int ldo, void* stream) {
    if (!coarse || !skip || !w_up || !b_up || !w_merge || !b_merge || !out || n <= 0 || h <= 0 || w <= 0) return CODD_E_BADARG;
    if ((h & 1) || (w & 1) || cc <= 0 || cs <= 0 || cc % 8 || cs % 8 || ldc < cc || lds < cs || ldc % 4 || lds % 4 ||
        ldo < co || ldo % 4 || h > 65535 || n > 65535)
        return CODD_E_SHAPE;
    if (!codd_aligned16(coarse) || !codd_aligned16(skip) || !codd_aligned16(out)) return CODD_E_ALIGN;
    UmP p;
    p.coarse = coarse; p.ldc = ldc; p.Cc = cc; p.skip = skip; p.lds = lds; p.Cs = cs;
    p.w_up = w_up; p.b_up = b_up; p.w_m = w_merge; p.b_m = b_merge; p.out = out; p.ldo = ldo;
    p.N = n; p.H = h; p.W = w;
    p.v8 = (ldo % 8 == 0) && codd_aligned32(out);
    p.w_bulk = codd_aligned16(w_up) && codd_aligned16(w_merge) && codd_aligned16(b_up) && codd_aligned16(b_merge);
    cudaStream_t s = (cudaStream_t)stream;
    // two parents per thread when a row of parents fills the 128-thread blocks that way (>= 3/4 of the slots used)
    const int wc = w / 2;
    const bool two = (wc % 256 == 0) || (wc % 256 > 192) || (wc > 256 && wc % 256 > 128);
    // the all-16-channel case at sizes that fill 128-parent CTAs: TMA-staged variant
    if (cu == 16 && co == 16 && (cc == 16 || cc == 24) && cs == 16 && wc >= 96) {
        const int rc = cc == 16 ? um_launch_tma<16>(p, s) : um_launch_tma<24>(p, s);
        if (rc != CODD_E_UNSUPPORTED) return rc;
    }
    if (cu == 16 && co == 16) return two ? um_launch<16, 16, 2>(p, s) : um_launch<16, 16, 1>(p, s);
    if (cu == 24 && co == 24) return um_launch<24, 24, 1>(p, s);
    return CODD_E_UNSUPPORTED;
}
