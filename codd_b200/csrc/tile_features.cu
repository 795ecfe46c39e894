// K2 — tile features: the first stage of TileInitialization
// (model/stereo/hitnet/initialization.py:62-95, 119-156).
//
//   hidden = LeakyReLU(conv4x4(in) + b0)      stride (4,4) on the left features,
//                                              stride (4,1) on the right features, whose input is
//                                              zero-padded by 3 columns on the right (so right
//                                              column x sees input columns x .. x+3)
//   out    = LeakyReLU(conv1x1(hidden) + b1)   16 -> 16
//
// Both convolutions run in one kernel (the 16 hidden channels never leave registers) and the
// result is written PLANAR ([N,16,h,Wo]): K1 stages whole channel rows with bulk-TMA copies and
// reads them at unit stride.  The reference mutates conv.stride between the two calls
// (initialization.py:122-124); here the stride is a template parameter.
//
// CTA = 128 threads = one tile row i x (128*NOUT) output columns.  Input rows 4i..4i+3 are
// streamed through shared memory one row and 8 channels at a time (pixel-major, padded to 12
// floats -> conflict-free 128-bit reads for unit lane stride); the matching weight slice sits
// beside it as [kx][ci][co] and is read as warp-uniform broadcasts.  Each thread owns NOUT
// output columns (t, t+128, ...) x 16 hidden channels.
#include "common.cuh"

namespace {

struct TfP {
    const float* in;   // [N,H,W,Cin] NHWC
    int ldi, Cin;
    const float* w0;   // packed [16 taps = ky*4+kx][Cin][16]
    const float* b0;
    const float* w1;   // torch [16 out][16 in]
    const float* b1;
    float* out;        // planar [N,16,h,Wo]
    int N, H, W, h, Wo;
};

constexpr int TF_THREADS = 128;
constexpr int TF_CK = 8, TF_CP = 12;

template <int S, int NOUT>
__global__ void __launch_bounds__(TF_THREADS) tile_features_kernel(TfP p) {
    constexpr int OCOLS = TF_THREADS * NOUT;            // output columns per CTA
    constexpr int ICOLS = (OCOLS - 1) * S + 4;          // input columns per CTA
    extern __shared__ float4 smem4[];
    float* s_in = reinterpret_cast<float*>(smem4);      // [ICOLS][TF_CP]
    float* s_w = s_in + ICOLS * TF_CP;                  // [4 kx][TF_CK][16]
    __shared__ __align__(16) float s_w1[16][16];        // [hidden][out]
    __shared__ float s_b0[16], s_b1[16];

    const int t = threadIdx.x;
    const int nblk = (p.Wo + OCOLS - 1) / OCOLS;
    int b = blockIdx.x;
    const int xblk = (b % nblk) * OCOLS;
    b /= nblk;
    const int i = b % p.h;
    const int n = b / p.h;
    const int ix0 = xblk * S;

    for (int k = t; k < 256; k += TF_THREADS) s_w1[k & 15][k >> 4] = __ldg(p.w1 + k);   // w1[out][in] -> [in][out]
    if (t < 16) {
        s_b0[t] = __ldg(p.b0 + t);
        s_b1[t] = __ldg(p.b1 + t);
    }

    __align__(8) float acc[NOUT][16];
#pragma unroll
    for (int o = 0; o < NOUT; ++o)
#pragma unroll
        for (int c = 0; c < 16; ++c) acc[o][c] = 0.f;

    for (int ky = 0; ky < 4; ++ky) {
        const int gy = 4 * i + ky;
        const float* rowp = p.in + ((size_t)n * p.H + gy) * p.W * p.ldi;
        for (int c0 = 0; c0 < p.Cin; c0 += TF_CK) {
            __syncthreads();
#pragma unroll 2
            for (int idx = t; idx < ICOLS * 2; idx += TF_THREADS) {
                const int c4 = idx & 1, col = idx >> 1;
                const int gx = ix0 + col;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (gx < p.W && c0 + c4 * 4 < p.Cin) v = ldg4(rowp + (size_t)gx * p.ldi + c0 + c4 * 4);
                *reinterpret_cast<float4*>(s_in + col * TF_CP + c4 * 4) = v;
            }
#pragma unroll 2
            for (int idx = t; idx < 4 * TF_CK * 16; idx += TF_THREADS) {
                const int co = idx & 15, ci = (idx >> 4) % TF_CK, kx = idx / (16 * TF_CK);
                float v = 0.f;
                if (c0 + ci < p.Cin) v = __ldg(p.w0 + ((size_t)(ky * 4 + kx) * p.Cin + c0 + ci) * 16 + co);
                s_w[idx] = v;
            }
            __syncthreads();
#pragma unroll 1
            for (int kx = 0; kx < 4; ++kx) {
#pragma unroll 1
                for (int c4 = 0; c4 < 2; ++c4) {
                    float4 a[NOUT];
#pragma unroll
                    for (int o = 0; o < NOUT; ++o)
                        a[o] = *reinterpret_cast<const float4*>(s_in + ((t + TF_THREADS * o) * S + kx) * TF_CP + c4 * 4);
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        const float* wp = s_w + (kx * TF_CK + c4 * 4 + cc) * 16;
#pragma unroll
                        for (int o4 = 0; o4 < 4; ++o4) {
                            const float4 wv = *reinterpret_cast<const float4*>(wp + o4 * 4);
#pragma unroll
                            for (int o = 0; o < NOUT; ++o) {
                                const float av = cc == 0 ? a[o].x : cc == 1 ? a[o].y : cc == 2 ? a[o].z : a[o].w;
                                fma4(&acc[o][o4 * 4], av, wv);
                            }
                        }
                    }
                }
            }
        }
    }

    // ---- LeakyReLU, 1x1 conv, LeakyReLU, planar store.  The weight loop is OUTERMOST: one broadcast of a w1 row
    // feeds all NOUT columns (with the column loop outside, the compiler kept all 256 weights live in registers
    // to share them between columns: 255 registers + spills).
    const size_t plane = (size_t)p.h * p.Wo;
#pragma unroll
    for (int o = 0; o < NOUT; ++o)
#pragma unroll
        for (int c = 0; c < 16; ++c) acc[o][c] = codd_act(acc[o][c] + s_b0[c], CODD_ACT_LEAKY, 0);   // hidden
    __align__(8) float outv[NOUT][16];
#pragma unroll
    for (int o = 0; o < NOUT; ++o)
#pragma unroll
        for (int c = 0; c < 16; ++c) outv[o][c] = s_b1[c];
#pragma unroll
    for (int hh = 0; hh < 16; ++hh) {
#pragma unroll
        for (int o4 = 0; o4 < 4; ++o4) {
            const float4 wv = *reinterpret_cast<const float4*>(&s_w1[hh][o4 * 4]);
#pragma unroll
            for (int o = 0; o < NOUT; ++o) fma4(&outv[o][o4 * 4], acc[o][hh], wv);
        }
    }
#pragma unroll
    for (int o = 0; o < NOUT; ++o) {
        const int xo = xblk + t + TF_THREADS * o;
        if (xo < p.Wo) {
            float* op = p.out + ((size_t)n * 16 * p.h + i) * p.Wo + xo;
#pragma unroll
            for (int c = 0; c < 16; ++c) op[(size_t)c * plane] = codd_act(outv[o][c], CODD_ACT_LEAKY, 0);
        }
    }
}

template <int S, int NOUT>
int launch_tf(const TfP& p, cudaStream_t s) {
    constexpr int OCOLS = TF_THREADS * NOUT;
    constexpr int ICOLS = (OCOLS - 1) * S + 4;
    const size_t smem = (size_t)(ICOLS * TF_CP + 4 * TF_CK * 16) * sizeof(float);
    const int nblk = codd_ceil_div(p.Wo, OCOLS);
    tile_features_kernel<S, NOUT><<<(unsigned)(p.N * p.h * nblk), TF_THREADS, smem, s>>>(p);
    CODD_RETURN_IF_CUDA_ERROR();
    return 0;
}

}  // namespace

extern "C" int codd_tile_features(const float* in, int ldi, int cin, int n, int h_in, int w_in, const float* w0,
                                  const float* b0, const float* w1, const float* b1, int right, float* out,
                                  void* stream) {
    if (!in || !w0 || !b0 || !w1 || !b1 || !out || n <= 0 || h_in <= 0 || w_in <= 0 || cin <= 0) return CODD_E_BADARG;
    if (cin % 4 != 0 || ldi < cin || ldi % 4 != 0 || h_in % 4 != 0 || w_in % 4 != 0) return CODD_E_SHAPE;
    if (!codd_aligned16(in)) return CODD_E_ALIGN;
    TfP p;
    p.in = in; p.ldi = ldi; p.Cin = cin; p.w0 = w0; p.b0 = b0; p.w1 = w1; p.b1 = b1; p.out = out;
    p.N = n; p.H = h_in; p.W = w_in; p.h = h_in / 4; p.Wo = right ? w_in : w_in / 4;
    if (right) return launch_tf<1, 4>(p, (cudaStream_t)stream);
    return launch_tf<4, 1>(p, (cudaStream_t)stream);
}
