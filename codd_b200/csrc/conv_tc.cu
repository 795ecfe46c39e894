// 3x3 / stride 1 / pad 1 convolution on the 5th-generation tensor cores (tcgen05, sm_100a).
//
// Replaces the dense 3x3 layers of HITUNet and of the tile-update networks
// (model/stereo/hitnet/backbone.py:8-39; propagation.py:89-121,258-323) — 55 % of the stereo step
// when run on the fp32 CUDA cores (conv.cu), where they are FMA-bound.
//
// Implicit GEMM, NHWC fp32 activations:
//   M = 128 output pixels of one image row, N = Cout (padded to 16/32), K = Cin per filter tap.
//   * a persistent CTA (one per SM) keeps ALL weights resident in shared memory
//     ([pass][tap][cout][cin], K-major, 128B/64B-swizzled) and loops over output tiles of
//     2 rows x 128 columns;
//   * the (2+2) x (128+2) x Cin input halo tile of a step arrives by ONE 4-D TMA load
//     (cp.async.bulk.tensor, out-of-bounds = zero fill = the convolution's zero padding;
//     channels 24..31 of a 24-channel tensor are zero-filled the same way) into a
//     double-buffered, hardware-swizzled stage: one pixel = one swizzle row, so the A operand
//     of tap (ky,kx) is the same stage viewed from a start address shifted by
//     (ky*130 + kx) pixels — no im2col, no per-tap copies;
//   * tcgen05.mma.kind::tf32 (M128 x N x K8) issued by one thread, accumulators in TMEM
//     (double-buffered: 2 tiles x 2 rows x N columns), completion tracked with tcgen05.commit
//     on mbarriers; the epilogue warps read TMEM with tcgen05.ld (lane = pixel), add bias /
//     residual, apply the activation and store NHWC.
//
// Precision: 3xTF32.  A single TF32 pass (10-bit mantissa) costs ~7e-4 relative error per layer
// and flips the network's discrete tile selections; splitting both operands (x = hi + lo) and
// accumulating hi*hi + hi*lo + lo*hi in fp32 recovers fp32-class accuracy (~2^-21).  Weights are
// split on the host.  Activations are split IN PLACE: passes 1-2 read the raw fp32 stage (the
// tensor core uses the top 19 bits = hi), then the epilogue warps overwrite the stage with
// lo = x - hi(x) and pass 3 runs lo * w_hi.  One stage buffer per tile instead of two.
#include <cuda.h>

#include "common.cuh"

namespace {

constexpr int TC_TW = 128;          // output columns per tile (= MMA M)
// Tile geometry per dilation.  DIL = 1: 2 output rows per tile, one 4-D TMA box of (2+2) x (128+2)
// pixels.  DIL = 3 (the dilated resblocks of tile_update4_1 / tile_update5, propagation.py:258-280):
// 1 output row per tile, the three input rows y-3, y, y+3 arrive as three one-row TMA boxes of 128+6
// pixels; the staged row pitch is padded to a multiple of 8 pixels so that every row starts on a
// 1024-byte swizzle-atom boundary.
template <int DIL>
struct TcGeo {
    static constexpr int R = (DIL == 1) ? 2 : 1;              // output rows per tile
    static constexpr int TROWS = (DIL == 1) ? 4 : 3;          // staged rows
    static constexpr int BOXW = TC_TW + 2 * DIL;              // staged columns delivered by TMA
    static constexpr int TWP = (DIL == 1) ? BOXW : ((BOXW + 7) & ~7);   // staged row pitch (pixels)
    static constexpr int NLOADS = (DIL == 1) ? 1 : 3;         // TMA boxes per tile
    static constexpr int BOXROWS = (DIL == 1) ? 4 : 1;
};
// warp roles (the SM's issue arbiter favours HIGH warp ids, and waiting warps poll their mbarrier,
// so the latency-critical single-thread roles get the highest ids and the bulk workers the lowest):
//   0-7  in-place hi/lo split of the stage,  8-11 epilogue (TMEM lane quarter = warp % 4),
//   12   TMA producer,  13  MMA issuer (+ TMEM alloc)
constexpr int TC_EPI_THREADS = 128;
constexpr int TC_SPLIT_THREADS = 256;
constexpr int TC_THREADS = 64 + TC_EPI_THREADS + TC_SPLIT_THREADS;

struct TcP {
    const float* wpk;   // [2][9][NP][KC] fp32: pass 0 = hi, pass 1 = lo (low 13 mantissa bits zero)
    const float* bias;
    const float* res;
    float* out;
    int N, H, W, Cout, ldo, ldr, res_bcast, act;
    int tilesX, tilesY, ntiles;
    long long* dbg;       // optional [grid][8] cycle counters (role wait times), NULL in production
    int split_rna;        // 1: hi(x) = round-to-nearest tf32, 0: truncation (what the MMA datapath does)
    int use_base_offset;  // 1: set the descriptor base-offset field from the start address (diagnostic)
    int diag;             // bit0: skip epilogue stores, bit1: skip the stage split (diagnostics only)
};

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
template <bool BACKOFF = false>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (BACKOFF && !done) __nanosleep(32);   // a polling warp must not starve the working warps of its SM sub-partition
    }
}
template <bool BACKOFF = false>
__device__ __forceinline__ void mbar_wait_t(uint32_t bar, uint32_t parity, long long& acc) {
    const long long t0 = clock64();
    mbar_wait<BACKOFF>(bar, parity);
    acc += clock64() - t0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
template <bool ACC>
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc) {
    if (ACC) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.eq.u32 p, 1, 1;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.eq.u32 p, 1, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc)
            : "memory");
    }
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
        "[%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// shared-memory matrix descriptor, K-major, swizzled (cute::UMMA::SmemDescriptor bit layout).
// The swizzle is a function of the shared-memory ADDRESS, so a tap's operand is simply the stage
// seen from a shifted start address; the base-offset field stays 0 (verified on B200: setting it
// from the address corrupts the result).
template <int KC>
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, int use_base_offset) {
    constexpr uint32_t ROWB = KC * 4;
    constexpr uint64_t LAYOUT = (KC == 32) ? 2ull : 4ull;   // SWIZZLE_128B : SWIZZLE_64B
    constexpr uint32_t SBO = 8 * ROWB;                      // 8-row group pitch
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);               // start address
    d |= (uint64_t)1 << 16;                                 // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(SBO >> 4) << 32;                        // stride byte offset
    d |= (uint64_t)1 << 46;                                 // descriptor version (Blackwell)
    if (use_base_offset) d |= (uint64_t)((saddr >> 7) & 7u) << 49;
    d |= LAYOUT << 61;
    return d;
}

// byte offset of 16-byte chunk j of row r inside a K-major swizzled tile whose base is 1024-aligned
template <int KC>
__device__ __forceinline__ uint32_t swz_off(int r, int j) {
    constexpr uint32_t ROWB = KC * 4;
    const uint32_t off = (uint32_t)r * ROWB + (uint32_t)j * 16u;
    constexpr uint32_t MASK = (KC == 32) ? 7u : 3u;
    return off ^ (((off >> 7) & MASK) << 4);
}

__device__ __forceinline__ float tf32_lo(float x, int rna) {
    uint32_t hi;
    if (rna) {
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
    } else {
        hi = __float_as_uint(x) & 0xFFFFE000u;
    }
    return __fsub_rn(x, __uint_as_float(hi));
}

// One pass (9 taps x KC/8 k-steps) of one 128-pixel row: fully unrolled, every operand descriptor is
// the stage / weight base descriptor plus a compile-time constant (>> 4) — a single thread issues
// an MMA every few instructions.  FIRST: the very first MMA overwrites the accumulator.
template <int KC, int NP, bool FIRST, int DIL>
__device__ __forceinline__ void tc_issue_pass(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc) {
    constexpr int TC_TWP = TcGeo<DIL>::TWP;
    constexpr uint32_t ROWB = KC * 4;
    constexpr uint32_t B_TAP = 2 * NP * ROWB;   // per tap: NP rows of w_hi followed by NP rows of w_lo
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
        for (int k = 0; k < KC / 8; ++k) {
            const uint32_t aoff = (uint32_t)((tap / 3) * TC_TWP + (tap % 3) * DIL) * ROWB + k * 32;
            const uint32_t boff = (uint32_t)tap * B_TAP + k * 32;
            if (FIRST && tap == 0 && k == 0)
                tc_mma_tf32<false>(d_tmem, a_desc + (aoff >> 4), b_desc + (boff >> 4), idesc);
            else
                tc_mma_tf32<true>(d_tmem, a_desc + (aoff >> 4), b_desc + (boff >> 4), idesc);
        }
    }
}

// NBUF stage buffers, NACC accumulator buffers (x TC_R rows x NP TMEM columns), LAG = how many tiles
// pass 3 trails passes 1-2 (LAG < NBUF, NACC >= LAG + 2).
template <int KC, int NP, int NBUF, int NACC, int LAG, int DIL>
__global__ void __launch_bounds__(TC_THREADS, 1) conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmap, TcP p) {
    static_assert(LAG >= 1 && LAG < NBUF && NACC >= LAG + 2, "pipeline depths");
    constexpr int TC_R = TcGeo<DIL>::R, TC_TWP = TcGeo<DIL>::TWP, TC_TROWS = TcGeo<DIL>::TROWS;
    constexpr uint32_t BOX_BYTES = TcGeo<DIL>::BOXROWS * TcGeo<DIL>::BOXW * KC * 4;   // bytes of one TMA box
    constexpr uint32_t TMEM_COLS = (NACC * TC_R * 2 * NP <= 128) ? 128u : (NACC * TC_R * 2 * NP <= 256) ? 256u : 512u;
    constexpr uint32_t ROWB = KC * 4;
    constexpr uint32_t A_BYTES = TC_TROWS * TC_TWP * ROWB;                 // bytes of one stage (incl. row padding)
    constexpr uint32_t A_STRIDE = (A_BYTES + 1023u) & ~1023u;
    constexpr uint32_t B_TAP = 2 * NP * ROWB;
    // instruction descriptors: N = 2*NP (x * [w_hi | w_lo] in ONE MMA: the stage is read once for both
    // weight halves — the MMA rate here is bound by the shared-memory read of A) and N = NP (x_lo * w_hi)
    constexpr uint32_t IDESC_BASE = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 4) << 24);
    constexpr uint32_t IDESC2 = IDESC_BASE | ((uint32_t)((2 * NP) >> 3) << 17);
    constexpr uint32_t IDESC1 = IDESC_BASE | ((uint32_t)(NP >> 3) << 17);
    constexpr int ACC_COLS = 2 * NP;   // per output row: [0,NP) = x*w_hi + x_lo*w_hi, [NP,2NP) = x*w_lo

    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) unsigned long long bars[4 * NBUF + 2 * NACC];
    __shared__ uint32_t tmem_base_slot;

    const uint32_t sbase = (s_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gbase = smem_raw + (sbase - s_u32(smem_raw));
    const uint32_t sB = sbase + NBUF * A_STRIDE;
    uint8_t* gB = gbase + NBUF * A_STRIDE;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar0 = s_u32(&bars[0]);
    // stage barriers (per stage buffer): FULL, EMPTY, P12, LO; accumulator barriers (2): ACCF, ACCE
    auto SBAR = [&](int kind, int b) { return bar0 + (uint32_t)(kind * NBUF + b) * 8u; };
    auto ABAR = [&](int kind, int b) { return bar0 + (uint32_t)(4 * NBUF + kind * NACC + b) * 8u; };
    enum { FULL = 0, EMPTY = 1, P12 = 2, LO = 3 };
    enum { ACCF = 0, ACCE = 1 };

    if (tid == 0) {
        for (int b = 0; b < NBUF; ++b) {
            mbar_init(SBAR(FULL, b), 1);
            mbar_init(SBAR(EMPTY, b), 1);
            mbar_init(SBAR(P12, b), 1);
            mbar_init(SBAR(LO, b), TC_SPLIT_THREADS);
        }
        for (int b = 0; b < NACC; ++b) {
            mbar_init(ABAR(ACCF, b), 1);
            mbar_init(ABAR(ACCE, b), TC_EPI_THREADS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 13) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(&tmem_base_slot)),
                     "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    // weights -> swizzled shared image (once per CTA)
    for (int idx = tid; idx < 18 * NP * (KC / 4); idx += TC_THREADS) {
        const int j = idx % (KC / 4);
        const int r = (idx / (KC / 4)) % NP;
        const int pt = idx / ((KC / 4) * NP);
        const float4 v = ldg4(p.wpk + ((size_t)pt * NP + r) * KC + j * 4);
        const int pass = pt / 9, tap = pt - 9 * pass;
        *reinterpret_cast<float4*>(gB + tap * B_TAP + swz_off<KC>(r + pass * NP, j)) = v;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_slot;
    codd_pdl_trigger();      // prologue above reads weights / bias only (programmatic dependent launch, common.cuh)
    codd_pdl_wait();

    if (warp == 12) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            long long w0 = 0;
            const long long tstart = clock64();
            int it = 0;
            for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++it) {
                const int sb = it % NBUF;
                const uint32_t ph = (uint32_t)(it / NBUF) & 1u;
                int q = t;
                const int tx = q % p.tilesX;
                q /= p.tilesX;
                const int ty = q % p.tilesY;
                const int n = q / p.tilesY;
                mbar_wait_t<true>(SBAR(EMPTY, sb), ph ^ 1u, w0);
                mbar_expect_tx(SBAR(FULL, sb), BOX_BYTES * TcGeo<DIL>::NLOADS);
                const int cx = tx * TC_TW - DIL;
#pragma unroll
                for (int ld = 0; ld < TcGeo<DIL>::NLOADS; ++ld) {
                    const int cy = (DIL == 1) ? ty * TC_R - 1 : ty + (ld - 1) * DIL;
                    asm volatile(
                        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, "
                        "%5, %6}], [%2];" ::"r"(sbase + sb * A_STRIDE + (uint32_t)ld * TC_TWP * ROWB),
                        "l"(&tmap), "r"(SBAR(FULL, sb)), "r"(0), "r"(cx), "r"(cy), "r"(n)
                        : "memory");
                }
            }
            (void)tstart;
        }
    } else if (warp == 13) {
        // ===================== MMA issuer =====================
        // Software-pipelined over tiles: passes 1-2 of tile i are issued BEFORE pass 3 of tile i-1, so
        // the tensor pipe works on the next tile while the split warps rewrite the previous stage.
        if (codd_elect_one()) {
            const uint64_t b_desc = make_desc<KC>(sB, 0);
            long long w1 = 0, w2 = 0, w3 = 0;
            auto pass3 = [&](int it) {
                const int sb = it % NBUF, ab = it % NACC;
                mbar_wait_t<true>(SBAR(LO, sb), (uint32_t)(it / NBUF) & 1u, w3);
                tc_fence_after();
                const uint64_t a_desc = make_desc<KC>(sbase + sb * A_STRIDE, p.use_base_offset);
#pragma unroll
                for (int mt = 0; mt < TC_R; ++mt)
                    tc_issue_pass<KC, NP, false, DIL>(tmem + (uint32_t)((ab * TC_R + mt) * ACC_COLS),
                                                 a_desc + ((uint32_t)(mt * TC_TWP) * ROWB >> 4), b_desc, IDESC1);
                tc_commit(ABAR(ACCF, ab));     // accumulators complete -> epilogue
                tc_commit(SBAR(EMPTY, sb));    // stage buffer free -> producer
            };
            int it = 0;
            for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++it) {
                const int sb = it % NBUF, ab = it % NACC;
                mbar_wait_t<true>(SBAR(FULL, sb), (uint32_t)(it / NBUF) & 1u, w1);
                mbar_wait_t<true>(ABAR(ACCE, ab), ((uint32_t)(it / NACC) & 1u) ^ 1u, w2);
                tc_fence_after();
                const uint64_t a_desc = make_desc<KC>(sbase + sb * A_STRIDE, p.use_base_offset);
#pragma unroll
                for (int mt = 0; mt < TC_R; ++mt) {
                    const uint32_t d_tmem = tmem + (uint32_t)((ab * TC_R + mt) * ACC_COLS);
                    const uint64_t a_mt = a_desc + ((uint32_t)(mt * TC_TWP) * ROWB >> 4);
                    tc_issue_pass<KC, NP, true, DIL>(d_tmem, a_mt, b_desc, IDESC2);   // x_hi * [w_hi | w_lo] (raw stage: MMA reads hi)
                }
                tc_commit(SBAR(P12, sb));
                if (it >= LAG) pass3(it - LAG);                                // x_lo * w_hi of an earlier tile
            }
            for (int j = (it > LAG ? it - LAG : 0); j < it; ++j) pass3(j);
            if (p.dbg) { p.dbg[blockIdx.x * 8 + 1] = w1; p.dbg[blockIdx.x * 8 + 2] = w2; p.dbg[blockIdx.x * 8 + 3] = w3; }
        }
    } else if (warp >= 8) {
        // ===================== epilogue (warps 8-11) =====================
        const int quarter = warp & 3;                 // TMEM lanes 32*quarter .. +31
        long long w4 = 0, wld = 0, warr = 0, wrest = 0;
        // run-time epilogue switches, evaluated once
        const float slope = p.act == CODD_ACT_LEAKY ? CODD_LEAKY_SLOPE : (p.act == CODD_ACT_RELU ? 0.f : 1.f);
        const float slope0 = (p.act == CODD_ACT_RELU || p.act == CODD_ACT_RELU_CH0) ? 0.f : slope;
        const bool full_vec = (p.Cout == NP) && ((p.ldo & 3) == 0) && ((((uintptr_t)p.out) & 15u) == 0);
        const bool res_vec = p.res && !p.res_bcast && (p.Cout == NP) && ((p.ldr & 3) == 0) &&
                             ((((uintptr_t)p.res) & 15u) == 0);
        const bool full_vec8 = full_vec && ((p.ldo & 7) == 0) && ((((uintptr_t)p.out) & 31u) == 0);   // see stg8
        const bool res_vec8 = res_vec && ((p.ldr & 7) == 0) && ((((uintptr_t)p.res) & 31u) == 0);
        float biasr[NP];                              // bias lives in registers for the whole kernel
#pragma unroll
        for (int c = 0; c < NP; ++c) biasr[c] = (p.bias && c < p.Cout) ? __ldg(p.bias + c) : 0.f;
        int it = 0;
        for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++it) {
            const int ab = it % NACC;
            int q = t;
            const int tx = q % p.tilesX;
            q /= p.tilesX;
            const int ty = q % p.tilesY;
            const int n = q / p.tilesY;
            mbar_wait_t<true>(ABAR(ACCF, ab), (uint32_t)(it / NACC) & 1u, w4);
            tc_fence_after();
            const int x = tx * TC_TW + quarter * 32 + lane;
#pragma unroll
            for (int mt = 0; mt < TC_R; ++mt) {
                float acc[ACC_COLS];
                const long long tl0 = clock64();
#pragma unroll
                for (int c = 0; c < ACC_COLS; c += 16)
                    tc_ld16(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((ab * TC_R + mt) * ACC_COLS + c), &acc[c]);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                wld += clock64() - tl0;
                const long long ta0 = clock64();
                if (mt == TC_R - 1) {
                    tc_fence_before();
                    mbar_arrive(ABAR(ACCE, ab));
                }
                warr += clock64() - ta0;
                const long long tr0 = clock64();
                const int y = ty * TC_R + mt;
                if (x >= p.W || y >= p.H) continue;
                const size_t opix = ((size_t)n * p.H + y) * p.W + x;
                float* op = p.out + opix * p.ldo;
                // straight-line epilogue: every run-time switch is hoisted out of the per-element code
                float v[NP];
#pragma unroll
                for (int c = 0; c < NP; ++c) v[c] = (acc[c] + acc[NP + c]) + biasr[c];
                if (p.res) {
                    const float* rp = p.res + opix * p.ldr;
                    if (p.res_bcast) {
                        const float rb = __ldg(rp);
#pragma unroll
                        for (int c = 0; c < NP; ++c) v[c] += rb;
                    } else if (res_vec8) {
#pragma unroll
                        for (int c8 = 0; c8 < NP; c8 += 8) {
                            float r8[8];
                            ldg8(rp + c8, r8);
#pragma unroll
                            for (int e = 0; e < 8; ++e) v[c8 + e] += r8[e];
                        }
                    } else if (res_vec) {
#pragma unroll
                        for (int c4 = 0; c4 < NP; c4 += 4) {
                            const float4 r4 = ldg4(rp + c4);
                            v[c4] += r4.x; v[c4 + 1] += r4.y; v[c4 + 2] += r4.z; v[c4 + 3] += r4.w;
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < NP; ++c)
                            if (c < p.Cout) v[c] += __ldg(rp + c);
                    }
                }
                if (p.act <= CODD_ACT_RELU_CH0) {
                    // none / leaky / relu / relu(ch0): max(v,0) + slope*min(v,0), slope in {1, 0.2, 0}
#pragma unroll
                    for (int c = 0; c < NP; ++c) {
                        const float sl = c == 0 ? slope0 : slope;
                        v[c] = fmaxf(v[c], 0.f) + sl * fminf(v[c], 0.f);
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < NP; ++c) v[c] = codd_act(v[c], p.act, c);
                }
                if (!(p.diag & 1)) {
                    if (full_vec8) {
#pragma unroll
                        for (int c8 = 0; c8 < NP; c8 += 8) stg8(op + c8, &v[c8]);
                    } else if (full_vec) {
#pragma unroll
                        for (int c4 = 0; c4 < NP; c4 += 4)
                            *reinterpret_cast<float4*>(op + c4) = make_float4(v[c4], v[c4 + 1], v[c4 + 2], v[c4 + 3]);
                    } else {
#pragma unroll
                        for (int c = 0; c < NP; ++c)
                            if (c < p.Cout) op[c] = v[c];
                    }
                }
                wrest += clock64() - tr0;
            }
        }
        if (p.dbg && tid == 256) { p.dbg[blockIdx.x * 8 + 4] = w4; p.dbg[blockIdx.x * 8 + 0] = wld; p.dbg[blockIdx.x * 8 + 7] = wrest; }
    } else {
        // ===================== in-place hi/lo split of the stage (warps 0-7) =====================
        const int stid = tid;
        long long w5 = 0, w6 = 0;
        int it = 0;
        for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++it) {
            const int sb = it % NBUF;
            mbar_wait_t<true>(SBAR(P12, sb), (uint32_t)(it / NBUF) & 1u, w5);   // passes 1-2 have consumed the raw stage
            tc_fence_after();
            const long long ts = clock64();
            float4* a4 = reinterpret_cast<float4*>(gbase + sb * A_STRIDE);
#pragma unroll 4
            for (int idx = stid; idx < ((p.diag & 2) ? 0 : (int)(A_BYTES / 16)); idx += TC_SPLIT_THREADS) {
                float4 v = a4[idx];
                v.x = tf32_lo(v.x, p.split_rna);
                v.y = tf32_lo(v.y, p.split_rna);
                v.z = tf32_lo(v.z, p.split_rna);
                v.w = tf32_lo(v.w, p.split_rna);
                a4[idx] = v;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(SBAR(LO, sb));
            w6 += clock64() - ts;
        }
        if (p.dbg && stid == 0) { p.dbg[blockIdx.x * 8 + 5] = w5; p.dbg[blockIdx.x * 8 + 6] = w6; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 13) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS));
    }
}

#ifdef CODD_DIAG
long long* g_tc_dbg = nullptr;   // diagnostic builds only (make DIAG=1): cycle-counter buffer
#endif

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_tmapEncodeTiled get_encode() {
    static PFN_tmapEncodeTiled fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (PFN_tmapEncodeTiled)ptr;
    }
    return fn;
}

template <int KC, int NP, int NBUF, int NACC, int LAG, int DIL>
int launch_tc(const CUtensorMap& tmap, TcP p, cudaStream_t s) {
    constexpr uint32_t ROWB = KC * 4;
    constexpr uint32_t A_STRIDE = ((TcGeo<DIL>::TROWS * TcGeo<DIL>::TWP * ROWB) + 1023u) & ~1023u;
    constexpr uint32_t B_BYTES = 2 * 9 * NP * ROWB;
    const size_t smem = NBUF * A_STRIDE + B_BYTES + 1024;
    auto kern = conv3x3_tc_kernel<KC, NP, NBUF, NACC, LAG, DIL>;
    static CoddDeviceOnce once;   // one per template instantiation
    if (int rc = codd_once_per_device(once, [&] {
            return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        }))
        return rc;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = p.ntiles < sms ? p.ntiles : sms;
    if (cudaError_t e = codd_launch_pdl(kern, dim3(grid), dim3(TC_THREADS), smem, s, tmap, p)) return (int)e;
    CODD_RETURN_IF_CUDA_ERROR();
    return 0;
}

}  // namespace

extern "C" int codd_conv3x3_tc_dil(const float* in, int ldi, int cin, int n, int h, int w, const float* weight_split,
                                   const float* bias, const float* residual, int ldr, int res_bcast, int cout, int act,
                                   float* out, int ldo, int dil, int flags, void* stream) {
    if (!in || !weight_split || !out || n <= 0 || h <= 0 || w <= 0 || cin <= 0 || cout <= 0) return CODD_E_BADARG;
    if (cin > 32 || cout > 32 || cin % 4 != 0 || ldi % 4 != 0 || ldi < cin || ldo < cout) return CODD_E_SHAPE;
    if (!codd_aligned16(in)) return CODD_E_ALIGN;
    PFN_tmapEncodeTiled enc = get_encode();
    if (!enc) return CODD_E_UNSUPPORTED;
    const int KC = cin <= 16 ? 16 : 32;
    const int NP = cout <= 16 ? 16 : 32;
    if (KC == 16 && NP == 32) return CODD_E_UNSUPPORTED;
    if (dil != 1 && !(dil == 3 && KC == 32 && NP == 32)) return CODD_E_UNSUPPORTED;
    const int boxw = dil == 1 ? TcGeo<1>::BOXW : TcGeo<3>::BOXW, boxrows = dil == 1 ? TcGeo<1>::BOXROWS : TcGeo<3>::BOXROWS;
    const int rows_per_tile = dil == 1 ? TcGeo<1>::R : TcGeo<3>::R;
    CUtensorMap tmap;
    const cuuint64_t gdim[4] = {(cuuint64_t)cin, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
    const cuuint64_t gstr[3] = {(cuuint64_t)ldi * 4, (cuuint64_t)w * ldi * 4, (cuuint64_t)h * w * ldi * 4};
    const cuuint32_t box[4] = {(cuuint32_t)KC, (cuuint32_t)boxw, (cuuint32_t)boxrows, 1u};
    const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)in, gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, KC == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return CODD_E_SHAPE;
    TcP p;
    p.wpk = weight_split; p.bias = bias; p.res = residual; p.out = out;
    p.N = n; p.H = h; p.W = w; p.Cout = cout; p.ldo = ldo; p.ldr = ldr; p.res_bcast = res_bcast; p.act = act;
    p.tilesX = codd_ceil_div(w, TC_TW);
    p.tilesY = codd_ceil_div(h, rows_per_tile);
    p.ntiles = p.tilesX * p.tilesY * n;
#ifdef CODD_DIAG
    p.dbg = g_tc_dbg;
#else
    p.dbg = nullptr;
#endif
    p.split_rna = (flags & 1) ? 1 : 0;
    p.use_base_offset = (flags & 2) ? 1 : 0;
    p.diag = (flags >> 2) & 3;
    cudaStream_t s = (cudaStream_t)stream;
    if (dil == 3) return launch_tc<32, 32, 2, 3, 1, 3>(tmap, p, s);
    if (KC == 32 && NP == 32) return launch_tc<32, 32, 2, 3, 1, 1>(tmap, p, s);
    if (KC == 32 && NP == 16) return launch_tc<32, 16, 2, 3, 1, 1>(tmap, p, s);
    return launch_tc<16, 16, 4, 4, 2, 1>(tmap, p, s);
}

extern "C" int codd_conv3x3_tc(const float* in, int ldi, int cin, int n, int h, int w, const float* weight_split,
                               const float* bias, const float* residual, int ldr, int res_bcast, int cout, int act,
                               float* out, int ldo, int flags, void* stream) {
    return codd_conv3x3_tc_dil(in, ldi, cin, n, h, w, weight_split, bias, residual, ldr, res_bcast, cout, act, out, ldo,
                               1, flags, stream);
}

// diagnostic: device buffer of [grid][8] int64 cycle counters filled by the next codd_conv3x3_tc launches
// (0 producer wait-empty, 1 mma wait-full, 2 mma wait-acc-empty, 3 mma wait-lo, 4 epilogue wait-acc-full,
//  5 split wait-p12, 6 split work, 7 producer total); NULL disables.
#ifdef CODD_DIAG
extern "C" CODD_API int codd_conv3x3_tc_debug(long long* dbg) {
    g_tc_dbg = dbg;
    return 0;
}
#endif
