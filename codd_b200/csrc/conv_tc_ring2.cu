// Two stacked 3x3 / stride 1 / pad 1 convolutions (16 -> 16 -> 16 channels) in ONE rolling-ring tcgen05 kernel:
//
//     t   = act_a(conv_a(x) + bias_a)                      (never written to global memory)
//     out = act_b(conv_b(t) + bias_b [+ residual])
//
// = HITUNet's conv_merge tail (backbone.py:17-32: ... 3x3, LeakyReLU, 3x3, LeakyReLU) and the 16-channel ResBlocks of
// FinalTileUpdate (propagation.py:103-121: conv1, LeakyReLU, conv2, + x, LeakyReLU).  At 16 channels the single-conv
// ring kernel (conv_tc_ring.cu) is HBM-bound (read x, write y = 128 bytes per pixel at ~70 % of the measured copy
// bandwidth): a chain of two costs 4 tensor passes, this kernel 2.
//
// Dataflow (all of conv_tc_ring.cu's machinery, twice, chained through shared memory and TMEM):
//   TMA (input row, 130 pixels) -> split warps (fp16 x_hi | 2^10 x_lo operand tiles) -> MMA-a threads -> TMEM ring A
//   (8 slots: one intermediate row each) -> epilogue-a warps: bias, activation, ZERO outside the image (conv_b's
//   padding), the same fp16 hi/lo split the single-conv kernel applies to a loaded row, written as conv_b's operand
//   tile into a T stage -> MMA-b thread -> TMEM ring B (8 slots) -> epilogue-b warps: bias, residual, activation,
//   256-bit stores.
// A strip is 126 output columns: the intermediate tile has 128 columns (x0-1 .. x0+126, one MMA M), conv_b's kx
// shifts read its rows p .. p+2, so outputs p = 126, 127 would need columns the tile does not have and are discarded
// (rows 128, 129 of a T stage are never written; whatever they hold only reaches those two discarded outputs).
// Row bookkeeping per item (sample, strip, segment of `rows` output rows starting at y0): input rows y0-2 .. y0+rows+1
// are staged (rows + 4), intermediate rows y0-1 .. y0+rows (rows + 2) are produced and consumed, so conv_a runs the
// single-conv schedule with rows + 2 "output" rows and conv_b the one with rows.
// Results are BIT-IDENTICAL to two codd_conv3x3_tc_ring launches: same operand split, same per-row MMA order (one
// issuing thread per conv and row, pass A then pass B), same epilogue arithmetic.
#include <cuda.h>
#include <cuda_fp16.h>

#include <type_traits>

#include "common.cuh"
#include "ring_util.cuh"

namespace {

constexpr int R2_KC = 16, R2_NP = 16;
constexpr int R2_TW = 126;              // valid output columns per strip
constexpr int R2_BOXW = 130;            // staged input pixels per row (x0-2 .. x0+127)
constexpr int R2_SLOT = 2 * R2_NP;      // TMEM columns of one accumulator row: [hi | lo]
constexpr int R2_RING = 8;              // accumulator rows in flight, per convolution
constexpr int R2_NBUF = 6;              // fp32 input row stages
constexpr int R2_NH = 6;                // fp16 operand stages of the input rows
constexpr int R2_NT = 4;                // fp16 operand stages of the intermediate rows
constexpr int R2_SPLIT_GROUP = 128;
// warps 0-7 split | 8-11 epilogue-a | 12 TMA | 13, 15 MMA-a (even / odd rows; 14 idle) | 16-19 epilogue-b | 20 MMA-b
constexpr int R2_THREADS = 21 * 32;

struct R2P {
    const float* wa;    // ring-packed weights of conv_a / conv_b (ops.pack_conv_weight_ring)
    const float* wb;
    const float* bias_a;
    const float* bias_b;
    const float* res;
    float* out;
    int N, H, W, ldo, ldr, act_a, act_b;
    int tilesX, nseg, seg, nitems;
};

// Cursor over the rows of one pipeline stage: items in grid-stride order; inside an item `cnt` = rows + EXTRA rows.
// EXTRA = 4: staged input rows, EXTRA = 2: intermediate rows (then `rows_out` = rows + 2 is conv_a's output count).
template <int EXTRA>
struct Cur2 {
    int item, t, rows, y0, x0, n;
    int g;         // running row counter of this CTA over all items (stage = g % depth)
    int orow0;     // running OUTPUT-row counter of the convolution this cursor feeds, at the item's first output row
    __device__ __forceinline__ void load(const R2P& p) {
        if (item >= p.nitems) return;
        int q = item;
        const int sg = q % p.nseg;
        q /= p.nseg;
        const int tx = q % p.tilesX;
        n = q / p.tilesX;
        x0 = tx * R2_TW;
        y0 = sg * p.seg;
        rows = min(p.seg, p.H - y0);
    }
    __device__ __forceinline__ void init(const R2P& p) {
        item = blockIdx.x; t = 0; g = 0; orow0 = 0;
        load(p);
    }
    __device__ __forceinline__ bool valid(const R2P& p) const { return item < p.nitems; }
    __device__ __forceinline__ int outs() const { return rows + EXTRA - 2; }   // output rows of the conv it feeds
    __device__ __forceinline__ void next(const R2P& p) {
        ++g;
        if (++t == rows + EXTRA) {
            t = 0;
            orow0 += rows + EXTRA - 2;
            item += gridDim.x;
            load(p);
        }
    }
};

// One pass (PASS 0: x_hi row x [w_hi | 2^10 w_lo], PASS 1: 2^10 x_lo row x [0 | w_hi]) of one staged row `t` of a
// convolution with `outs` output rows in this item: the row contributes tap ky to output row t - ky.  Same schedule
// as conv_tc_ring.cu (adjacent ring slots share one MMA with N = 2 or 3 slots; interior rows fully unrolled).
template <int PASS>
__device__ __forceinline__ void r2_issue(uint32_t tmem_ring, uint32_t a_desc, uint32_t b_pass, int t, int outs, int orow0,
                                         uint32_t acce_bar0, uint32_t a_issued, int g) {
    constexpr uint32_t RB = R2_KC * 2;                              // fp16 operand rows: 32 bytes
    constexpr uint32_t WB = 6 * R2_NP * RB;                         // one kx weight block
    constexpr uint32_t IDB = (1u << 4) | ((128u >> 4) << 24);       // D = f32, A = B = f16, M = 128
    constexpr int SLOT = R2_SLOT, RING = R2_RING;
    auto mma = [&](uint32_t d, uint32_t da, uint32_t db, uint32_t idesc, uint32_t acc) {
        tc_mma_lo<desc_hi<(int)RB>(), desc_hi<(int)RB>(), true>(d, da, db, idesc, acc);
    };
    const int kylo = max(0, t - outs + 1), kyhi = min(2, t);
    int slot[3], runn[3];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) slot[ky] = RING - 1 - ((orow0 + t - ky + 3 * RING) % RING);
    const bool v0 = kylo <= 0 && 0 <= kyhi, v1 = kylo <= 1 && 1 <= kyhi, v2 = kylo <= 2 && 2 <= kyhi;
    const bool j01 = v0 && v1 && slot[1] == slot[0] + 1;
    const bool j12 = v1 && v2 && slot[2] == slot[1] + 1;
    runn[0] = v0 ? 1 + (j01 ? 1 + (j12 ? 1 : 0) : 0) : 0;
    runn[1] = (v1 && !j01) ? 1 + (j12 ? 1 : 0) : 0;
    runn[2] = (v2 && !j12) ? 1 : 0;
    const bool has_fresh = (PASS == 0 && v0);
    if (has_fresh) {
        const int orow = orow0 + t;
        mbar_wait(acce_bar0 + (uint32_t)slot[0] * 8u, (((uint32_t)(orow / RING)) & 1u) ^ 1u);   // slot drained
        tc_fence_after();
    }
    if (PASS == 0 && a_issued != 0) {
        while (ld_acquire_s32(a_issued) < g) {}      // the two MMA-a threads hand the rows over in order
    }
    constexpr uint32_t IDESC3 = IDB | ((uint32_t)((3 * SLOT) >> 3) << 17);
    constexpr uint32_t IDESC2 = IDB | ((uint32_t)((2 * SLOT) >> 3) << 17);
    constexpr uint32_t IDESC1 = IDB | ((uint32_t)(SLOT >> 3) << 17);
    constexpr uint32_t KYB = (SLOT * RB) >> 4;
    auto interior = [&](auto split_tag) {
        constexpr int SPLIT = decltype(split_tag)::value;
        const uint32_t d0 = tmem_ring + (uint32_t)(slot[0] * SLOT);
        const uint32_t d1 = tmem_ring + (uint32_t)(slot[1] * SLOT);
        const uint32_t d2 = tmem_ring + (uint32_t)(slot[2] * SLOT);
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const uint32_t a_k = a_desc + (((uint32_t)kx * RB) >> 4);
            const uint32_t b_k = b_pass + (((uint32_t)kx * WB) >> 4);
            const bool first = (kx == 0);
            if (first && PASS == 0) {
                mma(d0, a_k, b_k, IDESC1, 0u);
                if (SPLIT == 1 || SPLIT == 0) {
                    mma(d1, a_k, b_k + KYB, IDESC2, 1u);
                } else {
                    mma(d1, a_k, b_k + KYB, IDESC1, 1u);
                    mma(d2, a_k, b_k + 2 * KYB, IDESC1, 1u);
                }
            } else if (SPLIT == 0) {
                mma(d0, a_k, b_k, IDESC3, 1u);
            } else if (SPLIT == 1) {
                mma(d0, a_k, b_k, IDESC1, 1u);
                mma(d1, a_k, b_k + KYB, IDESC2, 1u);
            } else {
                mma(d0, a_k, b_k, IDESC2, 1u);
                mma(d2, a_k, b_k + 2 * KYB, IDESC1, 1u);
            }
        }
    };
    if (v0 && v2) {
        if (j01 && j12) interior(std::integral_constant<int, 0>{});
        else if (j12) interior(std::integral_constant<int, 1>{});
        else interior(std::integral_constant<int, 2>{});
        return;
    }
    uint32_t d_run[3], i_run[3];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        d_run[ky] = tmem_ring + (uint32_t)(slot[ky] * SLOT);
        i_run[ky] = IDB | ((uint32_t)((runn[ky] * SLOT) >> 3) << 17);
    }
    if (has_fresh) {
        mma(d_run[0], a_desc, b_pass, IDESC1, 0u);
        if (runn[0] > 1) mma(d_run[0] + SLOT, a_desc, b_pass + KYB, IDB | ((uint32_t)(((runn[0] - 1) * SLOT) >> 3) << 17), 1u);
    }
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
        const uint32_t a_k = a_desc + (((uint32_t)kx * RB) >> 4);
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            if (runn[ky] == 0) continue;
            if (ky == 0 && kx == 0 && has_fresh) continue;
            const uint32_t boff = (uint32_t)kx * WB + (uint32_t)ky * (SLOT * RB);
            mma(d_run[ky], a_k, b_pass + (boff >> 4), i_run[ky], 1u);
        }
    }
}

// fp32 -> the operand pair of the ring kernels: x_hi = fp16(x), fp16(2^10 (x - x_hi)); |x| clamped to the fp16 range
__device__ __forceinline__ void r2_split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    a = fminf(fmaxf(a, -65504.f), 65504.f);
    b = fminf(fmaxf(b, -65504.f), 65504.f);
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn((a - hf.x) * RG_LO_SCALE, (b - hf.y) * RG_LO_SCALE);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

__global__ void __launch_bounds__(R2_THREADS, 1) conv3x3x2_tc_ring_kernel(const __grid_constant__ CUtensorMap tmap, R2P p) {
    constexpr int KC = R2_KC, NP = R2_NP, SLOT = R2_SLOT, RING = R2_RING, NBUF = R2_NBUF, NH = R2_NH, NT = R2_NT;
    constexpr uint32_t ROWB = KC * 4;
    constexpr uint32_t A_BYTES = R2_BOXW * ROWB;
    constexpr uint32_t A_STRIDE = (A_BYTES + 1023u) & ~1023u;
    constexpr uint32_t ROWH = KC * 2;
    constexpr uint32_t H_STRIDE = ((R2_BOXW * ROWH) + 1023u) & ~1023u;
    constexpr uint32_t HL_STRIDE = 2 * H_STRIDE;
    constexpr uint32_t WBLKH = 6 * NP * ROWH;
    constexpr uint32_t W_BYTES = 2 * 3 * WBLKH;      // one convolution's weights: pass A blocks, then pass B blocks

    extern __shared__ uint8_t smem_raw[];
    // barriers: raw FULL/EMPTY [NBUF] | x-operand LO/HEMPTY [NH] | t-operand TFULL/TEMPTY [NT] | ACCF/ACCE a [RING] | b [RING]
    __shared__ __align__(8) unsigned long long bars[2 * NBUF + 2 * NH + 2 * NT + 4 * RING];
    __shared__ uint32_t tmem_base_slot;
    __shared__ int a_issued_s;
    __shared__ __align__(16) float s_bias[2][NP];

    const uint32_t sbase = (s_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gbase = smem_raw + (sbase - s_u32(smem_raw));
    const uint32_t sH = sbase + NBUF * A_STRIDE;
    uint8_t* gH = gbase + NBUF * A_STRIDE;
    const uint32_t sT = sH + NH * HL_STRIDE;
    uint8_t* gT = gH + NH * HL_STRIDE;
    const uint32_t sW = sT + NT * HL_STRIDE;         // conv_a weights, then conv_b weights
    uint8_t* gW = gT + NT * HL_STRIDE;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar0 = s_u32(&bars[0]);
    auto SBAR = [&](int kind, int b) { return bar0 + (uint32_t)(kind * NBUF + b) * 8u; };
    auto HBAR = [&](int kind, int b) { return bar0 + (uint32_t)(2 * NBUF + kind * NH + b) * 8u; };
    auto TBAR = [&](int kind, int b) { return bar0 + (uint32_t)(2 * NBUF + 2 * NH + kind * NT + b) * 8u; };
    auto ABAR = [&](int conv, int kind, int b) {
        return bar0 + (uint32_t)(2 * NBUF + 2 * NH + 2 * NT + (conv * 2 + kind) * RING + b) * 8u;
    };
    enum { FULL = 0, EMPTY = 1 };
    enum { LO = 0, HEMPTY = 1 };
    enum { TFULL = 0, TEMPTY = 1 };
    enum { ACCF = 0, ACCE = 1 };
    const uint32_t a_issued = s_u32(&a_issued_s);

    if (tid == 0) {
        for (int b = 0; b < NBUF; ++b) {
            mbar_init(SBAR(FULL, b), 1);
            mbar_init(SBAR(EMPTY, b), R2_SPLIT_GROUP);
        }
        for (int b = 0; b < NH; ++b) {
            mbar_init(HBAR(LO, b), R2_SPLIT_GROUP);
            mbar_init(HBAR(HEMPTY, b), 1);
        }
        for (int b = 0; b < NT; ++b) {
            mbar_init(TBAR(TFULL, b), 128);
            mbar_init(TBAR(TEMPTY, b), 1);
        }
        for (int cv = 0; cv < 2; ++cv)
            for (int b = 0; b < RING; ++b) {
                mbar_init(ABAR(cv, ACCF, b), 1);
                mbar_init(ABAR(cv, ACCE, b), 128);
            }
        a_issued_s = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 2 * NP) {
        const float* bp = tid < NP ? p.bias_a : p.bias_b;
        s_bias[tid / NP][tid % NP] = bp ? __ldg(bp + (tid % NP)) : 0.f;
    }
    if (warp == 13) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(&tmem_base_slot)),
                     "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    // weights of both convolutions -> swizzled shared images: per conv 2 passes x 3 kx blocks of 6*NP rows x KC halves
    for (int idx = tid; idx < 2 * 2 * 3 * 6 * NP * (KC / 8); idx += R2_THREADS) {
        const int j = idx % (KC / 8);
        const int r = (idx / (KC / 8)) % (6 * NP);
        const int blk = (idx / ((KC / 8) * 6 * NP)) % 6;     // pass * 3 + kx
        const int cv = idx / ((KC / 8) * 6 * NP * 6);
        const float4 v = ldg4((cv ? p.wb : p.wa) + ((size_t)blk * 6 * NP + r) * (KC / 2) + j * 4);
        *reinterpret_cast<float4*>(gW + cv * W_BYTES + blk * WBLKH + swz_rb<ROWH>(r, j)) = v;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_slot;
    const uint32_t tmem_a = tmem, tmem_b = tmem + RING * SLOT;
    codd_pdl_trigger();      // prologue above reads weights / bias only (see common.cuh)
    codd_pdl_wait();

    if (warp == 12) {
        // ===================== TMA producer: one staged input row per step =====================
        if (codd_elect_one()) {
            Cur2<4> c;
            for (c.init(p); c.valid(p); c.next(p)) {
                const int sb = c.g % NBUF;
                mbar_wait(SBAR(EMPTY, sb), (((uint32_t)(c.g / NBUF)) & 1u) ^ 1u);
                mbar_expect_tx(SBAR(FULL, sb), A_BYTES);
                const int cx = c.x0 - 2, cy = c.y0 - 2 + c.t;
                asm volatile(
                    "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
                    "%6}], [%2];" ::"r"(sbase + sb * A_STRIDE),
                    "l"(&tmap), "r"(SBAR(FULL, sb)), "r"(0), "r"(cx), "r"(cy), "r"(c.n)
                    : "memory");
            }
        }
    } else if (warp == 13 || warp == 15) {
        // ===================== MMA-a issuers: warp 13 = even staged rows, warp 15 = odd ones =====================
        if (codd_elect_one()) {
            const uint32_t wA = desc_lo(sW), wAB = desc_lo(sW + 3 * WBLKH);
            Cur2<4> c;
            c.init(p);
            if (warp == 15 && c.valid(p)) c.next(p);
            while (c.valid(p)) {
                const int hb = c.g % NH;
                mbar_wait(HBAR(LO, hb), ((uint32_t)(c.g / NH)) & 1u);
                tc_fence_after();
                const uint32_t a0 = desc_lo(sH + hb * HL_STRIDE);
                r2_issue<0>(tmem_a, a0, wA, c.t, c.outs(), c.orow0, ABAR(0, ACCE, 0), a_issued, c.g);
                r2_issue<1>(tmem_a, a0 + (H_STRIDE >> 4), wAB, c.t, c.outs(), c.orow0, ABAR(0, ACCE, 0), 0u, c.g);
                st_release_s32(a_issued, c.g + 1);
                tc_commit(HBAR(HEMPTY, hb));
                if (c.t >= 2) {                               // intermediate row t - 2 of this item is complete
                    const int orow = c.orow0 + c.t - 2;
                    tc_commit(ABAR(0, ACCF, RING - 1 - (orow % RING)));
                }
                c.next(p);
                if (c.valid(p)) c.next(p);
            }
        }
    } else if (warp == 20) {
        // ===================== MMA-b issuer: one intermediate row per step =====================
        if (codd_elect_one()) {
            const uint32_t wB = desc_lo(sW + W_BYTES), wBB = desc_lo(sW + W_BYTES + 3 * WBLKH);
            Cur2<2> c;
            for (c.init(p); c.valid(p); c.next(p)) {
                const int tb = c.g % NT;
                mbar_wait(TBAR(TFULL, tb), ((uint32_t)(c.g / NT)) & 1u);
                tc_fence_after();
                const uint32_t a0 = desc_lo(sT + tb * HL_STRIDE);
                r2_issue<0>(tmem_b, a0, wB, c.t, c.outs(), c.orow0, ABAR(1, ACCE, 0), 0u, c.g);
                r2_issue<1>(tmem_b, a0 + (H_STRIDE >> 4), wBB, c.t, c.outs(), c.orow0, ABAR(1, ACCE, 0), 0u, c.g);
                tc_commit(TBAR(TEMPTY, tb));
                if (c.t >= 2) {
                    const int orow = c.orow0 + c.t - 2;
                    tc_commit(ABAR(1, ACCF, RING - 1 - (orow % RING)));
                }
            }
        }
    } else if (warp >= 8 && warp < 12) {
        // ===================== epilogue-a (warps 8-11): intermediate row -> conv_b operand tile =====================
        const int quarter = warp & 3;
        const int px = quarter * 32 + lane;                     // tile row = intermediate column x0 - 1 + px
        const float sa = p.act_a == CODD_ACT_LEAKY ? CODD_LEAKY_SLOPE : (p.act_a == CODD_ACT_RELU ? 0.f : 1.f);
        const float sa0 = (p.act_a == CODD_ACT_RELU || p.act_a == CODD_ACT_RELU_CH0) ? 0.f : sa;
        const float2 sl2 = make_float2(sa, sa), sl20 = make_float2(sa0, sa);
        const float2 un = make_float2(1.f / RG_LO_SCALE, 1.f / RG_LO_SCALE);
        Cur2<2> c;
        for (c.init(p); c.valid(p); c.next(p)) {
            const int orow = c.g;                               // intermediate rows are conv_a's output rows, in order
            const int slot = RING - 1 - (orow % RING);
            const int tb = c.g % NT;
            const int y = c.y0 - 1 + c.t, x = c.x0 - 1 + px;
            const bool inside = (y >= 0) && (y < p.H) && (x >= 0) && (x < p.W);
            float hi[16], lo[16];
            mbar_wait(ABAR(0, ACCF, slot), ((uint32_t)(orow / RING)) & 1u);
            tc_fence_after();
            const uint32_t tbase = tmem_a + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(slot * SLOT);
            tc_ld16(tbase, hi);
            tc_ld16(tbase + NP, lo);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            tc_fence_before();
            mbar_arrive(ABAR(0, ACCE, slot));
            uint32_t h2[8], l2[8];
#pragma unroll
            for (int ch = 0; ch < 16; ch += 2) {
                float2 t = __ffma2_rn(make_float2(lo[ch], lo[ch + 1]), un, make_float2(hi[ch], hi[ch + 1]));
                t = __fadd2_rn(t, *reinterpret_cast<const float2*>(&s_bias[0][ch]));
                const float2 m = __fmul2_rn(t, ch == 0 ? sl20 : sl2);
                const float va = inside ? fmaxf(t.x, m.x) : 0.f, vb = inside ? fmaxf(t.y, m.y) : 0.f;
                r2_split2(va, vb, h2[ch / 2], l2[ch / 2]);
            }
            mbar_wait(TBAR(TEMPTY, tb), (((uint32_t)(c.g / NT)) & 1u) ^ 1u);      // conv_b has read this stage's last row
            uint8_t* t8 = gT + tb * HL_STRIDE;
            *reinterpret_cast<uint4*>(t8 + swz_rb<ROWH>(px, 0)) = make_uint4(h2[0], h2[1], h2[2], h2[3]);
            *reinterpret_cast<uint4*>(t8 + swz_rb<ROWH>(px, 1)) = make_uint4(h2[4], h2[5], h2[6], h2[7]);
            *reinterpret_cast<uint4*>(t8 + H_STRIDE + swz_rb<ROWH>(px, 0)) = make_uint4(l2[0], l2[1], l2[2], l2[3]);
            *reinterpret_cast<uint4*>(t8 + H_STRIDE + swz_rb<ROWH>(px, 1)) = make_uint4(l2[4], l2[5], l2[6], l2[7]);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(TBAR(TFULL, tb));
        }
    } else if (warp >= 16 && warp < 20) {
        // ===================== epilogue-b (warps 16-19): one output row per step =====================
        const int quarter = warp & 3;
        const int px = quarter * 32 + lane;
        const float sb_ = p.act_b == CODD_ACT_LEAKY ? CODD_LEAKY_SLOPE : (p.act_b == CODD_ACT_RELU ? 0.f : 1.f);
        const float sb0 = (p.act_b == CODD_ACT_RELU || p.act_b == CODD_ACT_RELU_CH0) ? 0.f : sb_;
        const float2 sl2 = make_float2(sb_, sb_), sl20 = make_float2(sb0, sb_);
        const float2 un = make_float2(1.f / RG_LO_SCALE, 1.f / RG_LO_SCALE);
        int orow = 0;
        for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
            int q = item;
            const int sg = q % p.nseg;
            q /= p.nseg;
            const int tx = q % p.tilesX;
            const int n = q / p.tilesX;
            const int y0 = sg * p.seg;
            const int rows = min(p.seg, p.H - y0);
            const int x = tx * R2_TW + px;
            const bool xin = (px < R2_TW) && (x < p.W);
            const int xc = min(x, p.W - 1);
            for (int r = 0; r < rows; ++r, ++orow) {
                const int slot = RING - 1 - (orow % RING);
                const size_t opix = ((size_t)n * p.H + (y0 + r)) * p.W + xc;
                float* op = p.out + opix * p.ldo;
                float hi[16], lo[16], rr[16];
                mbar_wait(ABAR(1, ACCF, slot), ((uint32_t)(orow / RING)) & 1u);
                tc_fence_after();
                const uint32_t tbase = tmem_b + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(slot * SLOT);
                tc_ld16(tbase, hi);
                tc_ld16(tbase + NP, lo);
                if (p.res) {
                    const float* rp = p.res + opix * p.ldr;
                    ldg8(rp, &rr[0]);
                    ldg8(rp + 8, &rr[8]);
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                tc_fence_before();
                mbar_arrive(ABAR(1, ACCE, slot));
                float v[16];
#pragma unroll
                for (int ch = 0; ch < 16; ch += 2) {
                    float2 t = __ffma2_rn(make_float2(lo[ch], lo[ch + 1]), un, make_float2(hi[ch], hi[ch + 1]));
                    t = __fadd2_rn(t, *reinterpret_cast<const float2*>(&s_bias[1][ch]));
                    if (p.res) t = __fadd2_rn(t, make_float2(rr[ch], rr[ch + 1]));
                    const float2 m = __fmul2_rn(t, ch == 0 ? sl20 : sl2);
                    v[ch] = fmaxf(t.x, m.x);
                    v[ch + 1] = fmaxf(t.y, m.y);
                }
                if (xin) {
                    stg8(op, &v[0]);
                    stg8(op + 8, &v[8]);
                }
            }
        }
    } else if (warp < 8) {
        // ===================== split warps (0-3: even staged rows, 4-7: odd ones) =====================
        constexpr int UNITS = R2_BOXW * (KC / 8);
        constexpr int UMAX = (UNITS + R2_SPLIT_GROUP - 1) / R2_SPLIT_GROUP;
        const int gt = tid & (R2_SPLIT_GROUP - 1);
        Cur2<4> c;
        c.init(p);
        if (warp >= 4 && c.valid(p)) c.next(p);
        while (c.valid(p)) {
            const int sb = c.g % NBUF, hb = c.g % NH;
            mbar_wait(HBAR(HEMPTY, hb), (((uint32_t)(c.g / NH)) & 1u) ^ 1u);
            mbar_wait(SBAR(FULL, sb), ((uint32_t)(c.g / NBUF)) & 1u);
            const uint8_t* a8 = gbase + sb * A_STRIDE;
            uint8_t* h8 = gH + hb * HL_STRIDE;
            float4 v0[UMAX], v1[UMAX];
#pragma unroll
            for (int k = 0; k < UMAX; ++k) {
                const int idx = gt + k * R2_SPLIT_GROUP;
                if (idx < UNITS) {
                    const int ipx = idx / (KC / 8), u = idx - ipx * (KC / 8);
                    v0[k] = *reinterpret_cast<const float4*>(a8 + swz_off<KC>(ipx, 2 * u));
                    v1[k] = *reinterpret_cast<const float4*>(a8 + swz_off<KC>(ipx, 2 * u + 1));
                }
            }
#pragma unroll
            for (int k = 0; k < UMAX; ++k) {
                const int idx = gt + k * R2_SPLIT_GROUP;
                if (idx < UNITS) {
                    const int ipx = idx / (KC / 8), u = idx - ipx * (KC / 8);
                    uint32_t hi2[4], lo2[4];
                    r2_split2(v0[k].x, v0[k].y, hi2[0], lo2[0]);
                    r2_split2(v0[k].z, v0[k].w, hi2[1], lo2[1]);
                    r2_split2(v1[k].x, v1[k].y, hi2[2], lo2[2]);
                    r2_split2(v1[k].z, v1[k].w, hi2[3], lo2[3]);
                    *reinterpret_cast<uint4*>(h8 + swz_rb<ROWH>(ipx, u)) = make_uint4(hi2[0], hi2[1], hi2[2], hi2[3]);
                    *reinterpret_cast<uint4*>(h8 + H_STRIDE + swz_rb<ROWH>(ipx, u)) = make_uint4(lo2[0], lo2[1], lo2[2], lo2[3]);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(HBAR(LO, hb));
            mbar_arrive(SBAR(EMPTY, sb));
            c.next(p);
            if (c.valid(p)) c.next(p);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 13) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
    }
}

}  // namespace

/* see include/codd_b200.h */
extern "C" int codd_conv3x3x2_tc_ring(const float* in, int ldi, int n, int h, int w, const float* weight_ring_a,
                                      const float* bias_a, int act_a, const float* weight_ring_b, const float* bias_b,
                                      const float* residual, int ldr, int act_b, float* out, int ldo, void* stream) {
    if (!in || !weight_ring_a || !weight_ring_b || !out || n <= 0 || h <= 0 || w <= 0) return CODD_E_BADARG;
    if (ldi % 4 != 0 || ldi < 16 || ldo < 16 || ldo % 8 != 0 || (residual && (ldr < 16 || ldr % 8 != 0))) return CODD_E_SHAPE;
    if (act_a > CODD_ACT_RELU_CH0 || act_b > CODD_ACT_RELU_CH0 || act_a < 0 || act_b < 0) return CODD_E_UNSUPPORTED;
    if (!codd_aligned16(in) || !codd_aligned32(out) || (residual && !codd_aligned32(residual))) return CODD_E_ALIGN;
    PFN_tmapEncodeTiled enc = rg_get_encode();
    if (!enc) return CODD_E_UNSUPPORTED;
    CUtensorMap tmap;
    const cuuint64_t gdim[4] = {16u, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
    const cuuint64_t gstr[3] = {(cuuint64_t)ldi * 4, (cuuint64_t)w * ldi * 4, (cuuint64_t)h * w * ldi * 4};
    const cuuint32_t box[4] = {16u, (cuuint32_t)R2_BOXW, 1u, 1u};
    const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)in, gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return CODD_E_SHAPE;
    R2P p;
    p.wa = weight_ring_a; p.wb = weight_ring_b; p.bias_a = bias_a; p.bias_b = bias_b; p.res = residual; p.out = out;
    p.N = n; p.H = h; p.W = w; p.ldo = ldo; p.ldr = ldr; p.act_a = act_a; p.act_b = act_b;
    p.tilesX = codd_ceil_div(w, R2_TW);

    constexpr uint32_t A_STRIDE = ((R2_BOXW * R2_KC * 4) + 1023u) & ~1023u;
    constexpr uint32_t H_STRIDE = ((R2_BOXW * R2_KC * 2) + 1023u) & ~1023u;
    constexpr uint32_t W_BYTES = 2 * 3 * 6 * R2_NP * R2_KC * 2;
    constexpr size_t smem = R2_NBUF * A_STRIDE + (R2_NH + R2_NT) * 2 * H_STRIDE + 2 * W_BYTES + 1024;
    static_assert(smem + 2048 <= 232448, "shared memory budget");
    static CoddDeviceOnce once;
    if (int rc = codd_once_per_device(once, [&] {
            return cudaFuncSetAttribute(conv3x3x2_tc_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        }))
        return rc;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // row segments as in conv_tc_ring.cu; a segment's overhead is 4 halo rows here
    const int strips = p.N * p.tilesX;
    int best_seg = p.H;
    double best = -1.0;
    for (int nseg = 1; nseg <= p.H && nseg <= 64; ++nseg) {
        const int seg = codd_ceil_div(p.H, nseg);
        if (seg < 8 && nseg > 1) break;
        const int items = strips * codd_ceil_div(p.H, seg);
        const int rounds = codd_ceil_div(items, sms);
        const double eff = ((double)strips * p.H) / ((double)rounds * sms * (seg + 4));
        if (eff > best + 1e-9) { best = eff; best_seg = seg; }
    }
    p.seg = best_seg;
    p.nseg = codd_ceil_div(p.H, p.seg);
    p.nitems = strips * p.nseg;
    const int grid = p.nitems < sms ? p.nitems : sms;
    if (cudaError_t e = codd_launch_pdl(conv3x3x2_tc_ring_kernel, dim3(grid), dim3(R2_THREADS), smem, (cudaStream_t)stream,
                                        tmap, p))
        return (int)e;
    CODD_RETURN_IF_CUDA_ERROR();
    return 0;
}
