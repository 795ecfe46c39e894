// 4x4 / stride 2 / pad 1 convolution on the 5th-generation tensor cores (tcgen05, sm_100a), Cin = 16.
//
// Replaces the first layer of HITUNet's conv_down blocks at the two finest levels
// (model/stereo/hitnet/backbone.py:8-14: Conv2d(inp, oup, 4, stride=2, padding=1) + LeakyReLU; down1 16->16 at
// 576x960 -> 288x480, down2 16->24 at 288x480 -> 144x240).  On the fp32 CUDA cores (conv.cu, packed FFMA2) these two
// launches took 1.03 ms of a 9.4 ms step at 10-15 % of the HBM roofline: 256 MACs per output and channel are
// FMA-issue bound there, while the tensor pipe does them in a few hundred cycles per 128-pixel tile.
//
// Implicit GEMM, NHWC fp32 activations, M = 128 output pixels of one output row, N = Cout (padded to 16 / 32),
// K = 16 input channels per filter tap, 16 taps:
//   * input column of output xo and tap kx is xi = 2 xo + kx - 1: taps kx = 1, 3 read EVEN columns 2(xo + {0,1}),
//     taps kx = 0, 2 read ODD columns 2(xo + {-1,0}) + 1.  The feature map is presented to TMA as a 5-D tensor
//     {C, 2 (column parity), W/2, H, N}, so one box {16, 1, 136, 1, 1} delivers 136 same-parity pixels of one input row
//     K-major into a 64-byte-swizzled slot; the two taps that share a parity are the SAME slot seen from a start address
//     shifted by one pixel (the swizzle is a function of the address) — every input element is loaded once per
//     output row, no im2col, and out-of-range coordinates (left / right / top / bottom padding) arrive as zeros;
//   * a stage = the two parity slots of ONE input row (one ky); a tile consumes 4 stages; NBUF stages in flight;
//   * precision: three fp16 products with fp32 accumulation, as the rolling-ring kernel (conv_tc_ring.cu): the split
//     warps rewrite every staged pixel row IN PLACE from KC fp32 values to [x_hi = fp16(x) (KC halves) | fp16(2^10 (x -
//     x_hi)) (KC halves)] — the same bytes — so that the row's first half is the pass-A operand and its second half the
//     pass-B operand of kind::f16 MMAs (K = 16 per MMA: half the MMAs and half the shared-memory operand reads of the
//     3xTF32 version this replaced, which was bound by exactly those reads).  Pass A: x_hi x [w_hi | 2^10 w_lo] into
//     columns [0, 2 NP); pass B: 2^10 x_lo x w_hi into columns [NP, 2 NP); the epilogue adds hi + 2^-10 lo.  |x|, |w| are
//     clamped to the fp16 range (65504);
//   * accumulators in TMEM (NACC tiles x 2 NP columns), epilogue warps read them with tcgen05.ld (lane = pixel), add
//     the bias, apply the activation and store NHWC.
//
// The same kernel, re-parametrised, computes TileInitialization.tile_features (initialization.py:119-156) at the two
// finest levels (Cin = 16): the 4x4 conv with stride (4,4) on the left features (columns split into FOUR residue classes
// mod 4, one tap each, no shift) and with stride (4,1) over the right features zero-padded by 3 columns (ONE class, taps
// shifted by kx pixels; the padding is TMA's out-of-bounds fill), followed in the epilogue by LeakyReLU, the 16x16 1x1
// conv, LeakyReLU and a PLANAR store (what the cost-volume kernel reads).  General rule: input column
// xi = SW*xo + kx - PW = SW*(xo + s) + r with r = (kx - PW) mod SW, s = floor((kx - PW) / SW): slot r, pixel shift s.
#include <cuda_fp16.h>

#include "tc_util.cuh"

namespace {

constexpr int S2_TW = 128;                 // output columns per tile (= MMA M)
// KC = input channels per tap as staged (16, or 32 for Cin = 24 / 32: channels beyond Cin are zero-filled by TMA);
// a pixel row of the K-major tile is KC * 4 bytes = the swizzle span (SWIZZLE_64B / SWIZZLE_128B)
constexpr int S2_EPI_THREADS = 128;
constexpr int S2_SPLIT_THREADS = 256;
constexpr int S2_THREADS = 64 + S2_EPI_THREADS + S2_SPLIT_THREADS;   // warps 0-7 split, 8-11 epilogue, 12 TMA, 13 MMA

// geometry of a 4-wide kernel row at horizontal stride SW with left padding PW
template <int SW, int PW, int KC>
struct K4Geo {
    static constexpr uint32_t S2_ROWB = KC * 4;
    static constexpr int NSLOT = SW;
    __host__ __device__ static constexpr int cls(int kx) { return (((kx - PW) % SW) + SW) % SW; }
    __host__ __device__ static constexpr int fdiv(int kx) { return (kx - PW - cls(kx)) / SW; }            // floor((kx - PW) / SW)
    __host__ __device__ static constexpr int smin(int r) {
        int m = 1 << 20;
        for (int kx = 0; kx < 4; ++kx)
            if (cls(kx) == r && fdiv(kx) < m) m = fdiv(kx);
        return m == (1 << 20) ? 0 : m;
    }
    __host__ __device__ static constexpr int shift(int kx) { return fdiv(kx) - smin(cls(kx)); }             // pixel offset of tap kx inside its slot
    __host__ __device__ static constexpr int maxshift() {
        int m = 0;
        for (int kx = 0; kx < 4; ++kx)
            if (shift(kx) > m) m = shift(kx);
        return m;
    }
    static constexpr int BOXP = (S2_TW + maxshift() + 7) & ~7;                          // pixels per slot
    static constexpr uint32_t SLOT = ((uint32_t)BOXP * S2_ROWB + 1023u) & ~1023u;       // 1024-aligned for the swizzle atom
    static constexpr uint32_t STAGE = (uint32_t)NSLOT * SLOT;
};
static_assert(K4Geo<2, 1, 16>::cls(0) == 1 && K4Geo<2, 1, 16>::shift(0) == 0 && K4Geo<2, 1, 16>::shift(2) == 1 && K4Geo<2, 1, 16>::cls(3) == 0 &&
              K4Geo<2, 1, 16>::shift(3) == 1 && K4Geo<2, 1, 16>::smin(1) == -1 && K4Geo<2, 1, 16>::BOXP == 136, "stride-2 geometry");
static_assert(K4Geo<4, 0, 16>::cls(3) == 3 && K4Geo<4, 0, 16>::shift(3) == 0 && K4Geo<4, 0, 16>::BOXP == 128, "stride-4 geometry");
static_assert(K4Geo<1, 0, 16>::cls(3) == 0 && K4Geo<1, 0, 16>::shift(3) == 3 && K4Geo<1, 0, 16>::BOXP == 136, "stride-1 geometry");
struct S2P {
    const float* wpk;   // fp16 data: [16 taps][2 NP rows: w_hi (NP) | 2^10 w_lo (NP)][KC halves] (ops.pack_conv_weight_tc4)
    const float* bias;
    float* out;
    int N, Ho, Wo, Cout, ldo, act;
    int tilesX, ntiles;
    const float* w1;    // tile-feature epilogue: 1x1 conv [16][16] (torch layout) and its bias; out is PLANAR then
    const float* b1;
};

// NBUF stage buffers (as many as shared memory allows: a buffer's cycle — TMA latency, conversion, MMAs, commit — is
// ~3 us, so the stages in flight set the throughput), NACC accumulator buffers; LAG is unused since the fp16 conversion.
// SW / SH: strides, PW / PH: left / top padding, TILEF: tile-feature epilogue (LeakyReLU, 1x1, LeakyReLU, planar store)
template <int SW, int SH, int PW, int PH, int KC, int NP, int NBUF, int NACC, int LAG, bool TILEF>
__global__ void __launch_bounds__(S2_THREADS, 1) conv4x4s2_tc_kernel(const __grid_constant__ CUtensorMap tmap, S2P p) {
    static_assert(LAG >= 1 && LAG < NBUF && NACC >= 2, "pipeline depths");
    // the two conversion groups take alternate stages: with an even ring depth a stage barrier is always waited on by
    // the same group, which then sees every phase of it; with an odd depth a group would see every other phase and the
    // one-bit parity wait could pass a phase early (the hazard documented in conv_tc_ring.cu; found here as a hang)
    static_assert(NBUF % 2 == 0, "stage ring depth must be even");
    using G = K4Geo<SW, PW, KC>;
    constexpr int S2_KC = KC;
    constexpr uint32_t S2_ROWB = KC * 4;
    constexpr int S2_BOXP = G::BOXP;
    constexpr uint32_t S2_SLOT = G::SLOT, S2_STAGE = G::STAGE;
    constexpr uint32_t ACC_COLS = 2 * NP;          // per tile: [0,NP) = x_hi*w_hi, [NP,2NP) = 2^10 (x_hi*w_lo + x_lo*w_hi)
    constexpr uint32_t TMEM_COLS = (NACC * ACC_COLS <= 128) ? 128u : (NACC * ACC_COLS <= 256) ? 256u : 512u;
    constexpr uint32_t B_TAP = 2 * NP * S2_ROWB;   // per tap: NP rows of w_hi followed by NP rows of 2^10 w_lo; a row holds
                                                   // KC halves in its first S2_ROWB / 2 bytes (same pitch and swizzle as A)
    constexpr uint32_t IDESC_BASE = (1u << 4) | ((128u >> 4) << 24);                            // f32 acc, f16 x f16, M = 128
    constexpr uint32_t LO_OFF = S2_ROWB / 2;       // byte offset of the x_lo halves inside a staged pixel row
    constexpr uint32_t IDESC2 = IDESC_BASE | ((uint32_t)((2 * NP) >> 3) << 17);
    constexpr uint32_t IDESC1 = IDESC_BASE | ((uint32_t)(NP >> 3) << 17);
    constexpr uint32_t BOX_BYTES = S2_BOXP * S2_ROWB;

    extern __shared__ uint8_t s2_smem_raw[];
    __shared__ __align__(8) unsigned long long bars[4 * NBUF + 2 * NACC];
    __shared__ uint32_t tmem_base_slot;
    __shared__ __align__(16) float s_w1[TILEF ? 16 * 16 : 4];     // 1x1 weights, TRANSPOSED [ci][co]

    const uint32_t sbase = (s2_u32(s2_smem_raw) + 1023u) & ~1023u;
    uint8_t* gbase = s2_smem_raw + (sbase - s2_u32(s2_smem_raw));
    const uint32_t sB = sbase + NBUF * S2_STAGE;
    uint8_t* gB = gbase + NBUF * S2_STAGE;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar0 = s2_u32(&bars[0]);
    auto SBAR = [&](int kind, int b) { return bar0 + (uint32_t)(kind * NBUF + b) * 8u; };
    auto ABAR = [&](int kind, int b) { return bar0 + (uint32_t)(4 * NBUF + kind * NACC + b) * 8u; };
    enum { FULL = 0, EMPTY = 1, P12 = 2, LO = 3 };
    enum { ACCF = 0, ACCE = 1 };

    if (tid == 0) {
        for (int b = 0; b < NBUF; ++b) {
            s2_mbar_init(SBAR(FULL, b), 1);
            s2_mbar_init(SBAR(EMPTY, b), 1);
            s2_mbar_init(SBAR(P12, b), 1);
            s2_mbar_init(SBAR(LO, b), S2_SPLIT_THREADS / 2);
        }
        for (int b = 0; b < NACC; ++b) {
            s2_mbar_init(ABAR(ACCF, b), 1);
            s2_mbar_init(ABAR(ACCE, b), S2_EPI_THREADS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 13) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s2_u32(&tmem_base_slot)),
                     "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    // weights -> swizzled shared image (once per CTA): [tap][w_hi rows | 2^10 w_lo rows], KC halves per row
    for (int idx = tid; idx < 16 * 2 * NP * (S2_KC / 8); idx += S2_THREADS) {
        const int j = idx % (S2_KC / 8);
        const int r = (idx / (S2_KC / 8)) % (2 * NP);
        const int tap = idx / ((S2_KC / 8) * 2 * NP);
        const float4 v = ldg4(p.wpk + ((size_t)tap * 2 * NP + r) * (S2_KC / 2) + j * 4);
        *reinterpret_cast<float4*>(gB + tap * B_TAP + s2_swz<KC>(r, j)) = v;
    }
    if (TILEF)
        for (int idx = tid; idx < 256; idx += S2_THREADS) s_w1[(idx & 15) * 16 + (idx >> 4)] = __ldg(p.w1 + idx);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    s2_fence_before();
    __syncthreads();
    s2_fence_after();
    const uint32_t tmem = tmem_base_slot;
    codd_pdl_trigger();      // prologue above reads weights / bias only (programmatic dependent launch, common.cuh)
    codd_pdl_wait();

    // stage counter `it` = 4 * (tile index of this CTA) + ky
    if (warp == 12) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int it = 0;
            for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
                int q = t;
                const int tx = q % p.tilesX;
                q /= p.tilesX;
                const int yo = q % p.Ho;
                const int n = q / p.Ho;
                const int xo0 = tx * S2_TW;
                for (int ky = 0; ky < 4; ++ky, ++it) {
                    const int sb = it % NBUF;
                    s2_mbar_wait<false>(SBAR(EMPTY, sb), (((uint32_t)(it / NBUF)) & 1u) ^ 1u);
                    s2_mbar_expect_tx(SBAR(FULL, sb), G::NSLOT * BOX_BYTES);
                    const int y = SH * yo + ky - PH;
                    // slot r: columns SW*p + r, p from xo0 + smin(r) (stride 2, pad 1: slot 0 = even columns from xo0 for
                    // taps kx = 1, 3; slot 1 = odd columns from xo0 - 1 for kx = 0, 2)
#pragma unroll
                    for (int r = 0; r < G::NSLOT; ++r) {
                        asm volatile(
                            "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, "
                            "%5, %6, %7}], [%2];" ::"r"(sbase + sb * S2_STAGE + (uint32_t)r * S2_SLOT),
                            "l"(&tmap), "r"(SBAR(FULL, sb)), "r"(0), "r"(r), "r"(xo0 + G::smin(r)), "r"(y), "r"(n)
                            : "memory");
                    }
                }
            }
        }
    } else if (warp == 13) {
        // ===================== MMA issuer =====================
        if (codd_elect_one()) {
            const uint64_t b_desc = s2_desc<KC>(sB);
            int it = 0;
            for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
                for (int ky = 0; ky < 4; ++ky, ++it) {
                    const int sb = it % NBUF, ti = it >> 2, ab = ti % NACC;
                    s2_mbar_wait<false>(SBAR(LO, sb), ((uint32_t)(it / NBUF)) & 1u);          // operands written by the split warps
                    if (ky == 0) s2_mbar_wait<false>(ABAR(ACCE, ab), (((uint32_t)(ti / NACC)) & 1u) ^ 1u);
                    s2_fence_after();
                    const uint64_t a_desc = s2_desc<KC>(sbase + sb * S2_STAGE);
                    const uint32_t d_tmem = tmem + (uint32_t)ab * ACC_COLS;
                    // pass A: x_hi * [w_hi | 2^10 w_lo]
#pragma unroll
                    for (int kx = 0; kx < 4; ++kx)
#pragma unroll
                        for (int k = 0; k < S2_KC / 16; ++k) {
                            const uint32_t aoff = (uint32_t)G::cls(kx) * S2_SLOT + (uint32_t)G::shift(kx) * S2_ROWB + k * 32;
                            const uint32_t boff = (uint32_t)(ky * 4 + kx) * B_TAP + k * 32;
                            if (ky == 0 && kx == 0 && k == 0)
                                s2_mma_f16<false>(d_tmem, a_desc + (aoff >> 4), b_desc + (boff >> 4), IDESC2);
                            else
                                s2_mma_f16<true>(d_tmem, a_desc + (aoff >> 4), b_desc + (boff >> 4), IDESC2);
                        }
                    // pass B: 2^10 x_lo * w_hi into the scaled half
#pragma unroll
                    for (int kx = 0; kx < 4; ++kx)
#pragma unroll
                        for (int k = 0; k < S2_KC / 16; ++k) {
                            const uint32_t aoff = (uint32_t)G::cls(kx) * S2_SLOT + (uint32_t)G::shift(kx) * S2_ROWB + LO_OFF + k * 32;
                            const uint32_t boff = (uint32_t)(ky * 4 + kx) * B_TAP + k * 32;
                            s2_mma_f16<true>(d_tmem + NP, a_desc + (aoff >> 4), b_desc + (boff >> 4), IDESC1);
                        }
                    s2_commit(SBAR(EMPTY, sb));                  // stage buffer free -> producer
                    if (ky == 3) s2_commit(ABAR(ACCF, ab));      // the tile's accumulators are complete -> epilogue
                }
            }
        }
    } else if (warp >= 8) {
        // ===================== epilogue (warps 8-11) =====================
        const int quarter = warp & 3;                 // TMEM lanes 32*quarter .. +31
        const float slope = p.act == CODD_ACT_LEAKY ? CODD_LEAKY_SLOPE : (p.act == CODD_ACT_RELU ? 0.f : 1.f);
        const float slope0 = (p.act == CODD_ACT_RELU || p.act == CODD_ACT_RELU_CH0) ? 0.f : slope;
        const bool full_vec = ((p.ldo & 3) == 0) && ((((uintptr_t)p.out) & 15u) == 0) && (p.Cout % 4 == 0);
        const bool full_vec8 = full_vec && ((p.ldo & 7) == 0) && ((((uintptr_t)p.out) & 31u) == 0) && (p.Cout % 8 == 0);
        float biasr[NP];
#pragma unroll
        for (int c = 0; c < NP; ++c) biasr[c] = (p.bias && c < p.Cout) ? __ldg(p.bias + c) : 0.f;
        int ti = 0;
        for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++ti) {
            const int ab = ti % NACC;
            int q = t;
            const int tx = q % p.tilesX;
            q /= p.tilesX;
            const int yo = q % p.Ho;
            const int n = q / p.Ho;
            s2_mbar_wait<true, 128>(ABAR(ACCF, ab), ((uint32_t)(ti / NACC)) & 1u);
            s2_fence_after();
            float acc[ACC_COLS];
#pragma unroll
            for (int c = 0; c < (int)ACC_COLS; c += 16)
                s2_ld16(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)ab * ACC_COLS + c, &acc[c]);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            s2_fence_before();
            s2_mbar_arrive(ABAR(ACCE, ab));
            const int xo = tx * S2_TW + quarter * 32 + lane;
            if (xo >= p.Wo) continue;
            float v[NP];
#pragma unroll
            for (int c = 0; c < NP; ++c) v[c] = fmaf(acc[NP + c], 1.f / 1024.f, acc[c]) + biasr[c];
            if (TILEF) {
                // initialization.py:119-124: LeakyReLU -> 1x1 conv (16 -> 16) -> LeakyReLU, planar [n,16,Ho,Wo]
#pragma unroll
                for (int c = 0; c < 16; ++c) v[c] = v[c] > 0.f ? v[c] : v[c] * CODD_LEAKY_SLOPE;
                float o[16];
#pragma unroll
                for (int c = 0; c < 16; ++c) o[c] = __ldg(p.b1 + c);
#pragma unroll
                for (int ci = 0; ci < 16; ++ci) {
#pragma unroll
                    for (int c4 = 0; c4 < 16; c4 += 4) {
                        const float4 w4 = *reinterpret_cast<const float4*>(&s_w1[ci * 16 + c4]);
                        fma4(&o[c4], v[ci], w4);
                    }
                }
                float* pp = p.out + (((size_t)n * 16) * p.Ho + yo) * p.Wo + xo;
                const size_t plane = (size_t)p.Ho * p.Wo;
#pragma unroll
                for (int c = 0; c < 16; ++c) pp[c * plane] = o[c] > 0.f ? o[c] : o[c] * CODD_LEAKY_SLOPE;
                continue;
            }
            float* op = p.out + (((size_t)n * p.Ho + yo) * p.Wo + xo) * p.ldo;
            if (p.act <= CODD_ACT_RELU_CH0) {
#pragma unroll
                for (int c = 0; c < NP; ++c) {
                    const float sl = c == 0 ? slope0 : slope;
                    v[c] = fmaxf(v[c], 0.f) + sl * fminf(v[c], 0.f);
                }
            } else {
#pragma unroll
                for (int c = 0; c < NP; ++c) v[c] = codd_act(v[c], p.act, c);
            }
            if (full_vec8) {
#pragma unroll
                for (int c8 = 0; c8 < NP; c8 += 8)
                    if (c8 < p.Cout) stg8(op + c8, &v[c8]);
            } else if (full_vec) {
#pragma unroll
                for (int c4 = 0; c4 < NP; c4 += 4)
                    if (c4 < p.Cout) *reinterpret_cast<float4*>(op + c4) = make_float4(v[c4], v[c4 + 1], v[c4 + 2], v[c4 + 3]);
            } else {
#pragma unroll
                for (int c = 0; c < NP; ++c)
                    if (c < p.Cout) op[c] = v[c];
            }
        }
    } else {
        // ===================== in-place fp32 -> [fp16 hi | fp16 2^10 lo] conversion of the stage (warps 0-7) =====================
        // one thread owns whole pixel rows (the conversion permutes bytes inside a row, never across rows)
        constexpr int NPIX = G::NSLOT * S2_BOXP;                                      // pixel rows per stage
        // two groups of four warps take alternate stages (a stage's conversion is a latency chain: wait, loads,
        // conversions, stores, proxy fence, arrive — two of them in flight)
        constexpr int GT = S2_SPLIT_THREADS / 2;
        constexpr int ITERS = (NPIX + GT - 1) / GT;
        constexpr int NCH = S2_KC / 4;                                                // 16-byte chunks per row
        const int grp = warp >> 2, gtid = tid & (GT - 1);
        int it = 0;
        for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
            for (int ky = 0; ky < 4; ++ky, ++it) {
                if ((it & 1) != grp) continue;
                const int sb = it % NBUF;
                s2_mbar_wait<true, 64>(SBAR(FULL, sb), ((uint32_t)(it / NBUF)) & 1u);     // the raw rows have landed
                // rows in batches of BR (all loads of a batch first, then its conversions and stores): the whole stage at
                // once for 64-byte rows, two rows at a time for 128-byte rows (register budget)
                constexpr int BR = (NCH == 8) ? 2 : ITERS;
#pragma unroll
                for (int i0 = 0; i0 < ITERS; i0 += BR) {
                    float4 v[BR][NCH];
#pragma unroll
                    for (int ib = 0; ib < BR; ++ib) {
                        const int idx = gtid + (i0 + ib) * GT;
                        if (i0 + ib < ITERS && idx < NPIX) {
                            const int r = idx / S2_BOXP, px = idx - r * S2_BOXP;
                            const uint8_t* row = gbase + sb * S2_STAGE + r * S2_SLOT;
#pragma unroll
                            for (int j = 0; j < NCH; ++j) v[ib][j] = *reinterpret_cast<const float4*>(row + s2_swz<KC>(px, j));
                        }
                    }
#pragma unroll
                    for (int ib = 0; ib < BR; ++ib) {
                        const int idx = gtid + (i0 + ib) * GT;
                        if (i0 + ib < ITERS && idx < NPIX) {
                            const int r = idx / S2_BOXP, px = idx - r * S2_BOXP;
                            uint8_t* row = gbase + sb * S2_STAGE + r * S2_SLOT;
                            uint32_t hi2[NCH * 2], lo2[NCH * 2];
#pragma unroll
                            for (int j = 0; j < NCH; ++j) {
                                const float xs[4] = {v[ib][j].x, v[ib][j].y, v[ib][j].z, v[ib][j].w};
#pragma unroll
                                for (int e = 0; e < 2; ++e) {
                                    const float a = fminf(fmaxf(xs[2 * e], -65504.f), 65504.f);
                                    const float b = fminf(fmaxf(xs[2 * e + 1], -65504.f), 65504.f);
                                    const __half2 h = __floats2half2_rn(a, b);
                                    const float2 hf = __half22float2(h);
                                    const __half2 l = __floats2half2_rn((a - hf.x) * 1024.f, (b - hf.y) * 1024.f);
                                    hi2[2 * j + e] = *reinterpret_cast<const uint32_t*>(&h);
                                    lo2[2 * j + e] = *reinterpret_cast<const uint32_t*>(&l);
                                }
                            }
#pragma unroll
                            for (int jj = 0; jj < NCH / 2; ++jj) {
                                *reinterpret_cast<uint4*>(row + s2_swz<KC>(px, jj)) =
                                    make_uint4(hi2[4 * jj], hi2[4 * jj + 1], hi2[4 * jj + 2], hi2[4 * jj + 3]);
                                *reinterpret_cast<uint4*>(row + s2_swz<KC>(px, NCH / 2 + jj)) =
                                    make_uint4(lo2[4 * jj], lo2[4 * jj + 1], lo2[4 * jj + 2], lo2[4 * jj + 3]);
                            }
                        }
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                s2_mbar_arrive(SBAR(LO, sb));
            }
        }
    }
    s2_fence_before();
    __syncthreads();
    if (warp == 13) {
        s2_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS));
    }
}

template <int SW, int SH, int PW, int PH, int KC, int NP, int NBUF, int NACC, int LAG, bool TILEF>
int s2_launch(const CUtensorMap& tmap, S2P p, cudaStream_t s) {
    constexpr size_t SMEM = (size_t)NBUF * K4Geo<SW, PW, KC>::STAGE + 16 * 2 * NP * (KC * 4) + 1024;
    static_assert(SMEM + 2048 <= 232448, "shared memory budget");
    static_assert(NACC * 2 * NP <= 512, "TMEM columns");
    const size_t smem = SMEM;
    auto kern = conv4x4s2_tc_kernel<SW, SH, PW, PH, KC, NP, NBUF, NACC, LAG, TILEF>;
    static CoddDeviceOnce once;
    if (int rc = codd_once_per_device(once, [&] {
            return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        }))
        return rc;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = p.ntiles < sms ? p.ntiles : sms;
    if (cudaError_t e = codd_launch_pdl(kern, dim3(grid), dim3(S2_THREADS), smem, s, tmap, p)) return (int)e;
    CODD_RETURN_IF_CUDA_ERROR();
    return 0;
}

}  // namespace

namespace {
// 5-D view {channel, column class (x mod SW), x / SW, row, sample} of an NHWC map whose width is a multiple of SW
int s2_make_tmap(CUtensorMap* tmap, const float* in, int ldi, int cin, int kc, int n, int h, int w, int sw, int boxp) {
    static PFN_s2EncodeTiled enc = s2_get_encode();      // C++11 magic static
    if (!enc) return CODD_E_UNSUPPORTED;
    const cuuint64_t gdim[5] = {(cuuint64_t)cin, (cuuint64_t)sw, (cuuint64_t)(w / sw), (cuuint64_t)h, (cuuint64_t)n};
    const cuuint64_t gstr[4] = {(cuuint64_t)ldi * 4, (cuuint64_t)ldi * 4 * sw, (cuuint64_t)w * ldi * 4,
                                (cuuint64_t)h * w * ldi * 4};
    const cuuint32_t box[5] = {(cuuint32_t)kc, 1u, (cuuint32_t)boxp, 1u, 1u};
    const cuuint32_t estr[5] = {1u, 1u, 1u, 1u, 1u};
    const CUresult r = enc(tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)in, gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, kc == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : CODD_E_UNSUPPORTED;
}
}  // namespace

extern "C" int codd_conv4x4s2_tc(const float* in, int ldi, int cin, int n, int h, int w, const float* weight_split,
                                 const float* bias, int cout, int act, float* out, int ldo, void* stream) {
    if (!in || !weight_split || !out || n <= 0 || h <= 0 || w <= 0 || cout <= 0) return CODD_E_BADARG;
    if ((cin != 16 && cin != 24 && cin != 32) || cout > 32 || (w & 1) || (h & 1) || ldi % 4 != 0 || ldi < cin || ldo < cout)
        return CODD_E_UNSUPPORTED;
    if (!codd_aligned16(in)) return CODD_E_ALIGN;
    const int kc = cin <= 16 ? 16 : 32;
    CUtensorMap tmap;
    if (int rc = s2_make_tmap(&tmap, in, ldi, cin, kc, n, h, w, 2, K4Geo<2, 1, 16>::BOXP)) return rc;
    S2P p;
    p.wpk = weight_split; p.bias = bias; p.out = out; p.w1 = nullptr; p.b1 = nullptr;
    p.N = n; p.Ho = h / 2; p.Wo = w / 2; p.Cout = cout; p.ldo = ldo; p.act = act;
    p.tilesX = codd_ceil_div(p.Wo, S2_TW);
    const long long nt = (long long)p.tilesX * p.Ho * n;
    if (nt > 0x7fffffffLL) return CODD_E_SHAPE;
    p.ntiles = (int)nt;
    cudaStream_t s = (cudaStream_t)stream;
    if (kc == 16) {
        if (cout <= 16) return s2_launch<2, 2, 1, 1, 16, 16, 10, 4, 2, false>(tmap, p, s);
        return s2_launch<2, 2, 1, 1, 16, 32, 8, 4, 2, false>(tmap, p, s);
    }
    // Cin = 24 / 32 (down3, down4): 128 KB of weights leave room for two stages
    if (cout <= 16) return s2_launch<2, 2, 1, 1, 32, 16, 4, 4, 1, false>(tmap, p, s);
    return s2_launch<2, 2, 1, 1, 32, 32, 2, 4, 1, false>(tmap, p, s);
}

// Tile features on the tensor cores: see the header comment.  w0_split as for codd_conv4x4s2_tc with NP = 16.
extern "C" int codd_tile_features_tc(const float* in, int ldi, int cin, int n, int h_in, int w_in, const float* w0_split,
                                     const float* b0, const float* w1, const float* b1, int right, float* out,
                                     void* stream) {
    if (!in || !w0_split || !b0 || !w1 || !b1 || !out || n <= 0 || h_in <= 0 || w_in <= 0) return CODD_E_BADARG;
    if ((cin != 16 && cin != 24 && cin != 32) || ldi < cin || ldi % 4 != 0 || h_in % 4 != 0 || w_in % 4 != 0)
        return CODD_E_UNSUPPORTED;
    if (!codd_aligned16(in)) return CODD_E_ALIGN;
    const int kc = cin <= 16 ? 16 : 32;
    CUtensorMap tmap;
    const int sw = right ? 1 : 4;
    if (int rc = s2_make_tmap(&tmap, in, ldi, cin, kc, n, h_in, w_in, sw,
                              right ? K4Geo<1, 0, 16>::BOXP : K4Geo<4, 0, 16>::BOXP))
        return rc;
    S2P p;
    p.wpk = w0_split; p.bias = b0; p.out = out; p.w1 = w1; p.b1 = b1;
    p.N = n; p.Ho = h_in / 4; p.Wo = right ? w_in : w_in / 4; p.Cout = 16; p.ldo = 16; p.act = CODD_ACT_LEAKY;
    p.tilesX = codd_ceil_div(p.Wo, S2_TW);
    const long long nt = (long long)p.tilesX * p.Ho * n;
    if (nt > 0x7fffffffLL) return CODD_E_SHAPE;
    p.ntiles = (int)nt;
    cudaStream_t s = (cudaStream_t)stream;
    if (kc == 16) {
        if (right) return s2_launch<1, 4, 0, 0, 16, 16, 12, 4, 2, true>(tmap, p, s);
        return s2_launch<4, 4, 0, 0, 16, 16, 4, 4, 2, true>(tmap, p, s);
    }
    if (right) return s2_launch<1, 4, 0, 0, 32, 16, 8, 4, 2, true>(tmap, p, s);
    return s2_launch<4, 4, 0, 0, 32, 16, 2, 4, 1, true>(tmap, p, s);
}
