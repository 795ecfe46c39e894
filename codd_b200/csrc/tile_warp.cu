// K4 (second decomposition) — local cost volume of HITNet's tile update + fused `decrease` 1x1 conv.
//
//   propagation.py:35-86 (warp, TileWarping), :156-160, :206-219 (decrease, augmented hypothesis concat)
//
// Same arithmetic as the first K4 (tile.cu, kept for planar right features): the reference's normalise /
// un-normalise coordinate round trip, torch's bilinear FMA order, sequential channel sum — bit-identical raw costs.
// What changed is where the bytes go (ncu of the first kernel: 836 L1 data-pipe wavefronts per 32 pixels, the L1 /
// shared-memory data pipe 75 % busy, 67 of them register spills, 256 the `decrease` GEMV out of shared memory):
//
//   * the right-feature window a CTA touches (rows x columns, data dependent) is staged by TMA:
//     `cp.async.bulk.tensor` boxes of {C+4 channels, 32 columns, 1 row} from the NHWC map — the 4 extra floats of a
//     box lie beyond the tensor's channel extent, so TMA zero-fills them and the window lands in shared memory with a
//     pixel pitch of C+4 floats (odd number of 16-byte chunks: the 128-bit reads of 8 neighbouring lanes hit 8 different
//     bank groups) without any thread copying a byte;
//   * the three planes k = -1, 0, +1 of a hypothesis set sample one pixel apart: when all lanes of a warp see the
//     nominal alignment (floor(ix_k) = floor(ix_0) - k, which holds unless ix sits within an ulp of an integer) the
//     planes share a 4-column window held in registers (4 LDS.128 per 4 channels instead of 6); otherwise the warp
//     takes the per-plane path (6 loads), so the result never depends on the alignment;
//   * `decrease` (64 -> 16, LeakyReLU) runs on the tensor cores straight out of the registers that hold the costs:
//     lane = 4*tile + xo of a warp IS the A-fragment layout of mma.m16n8k8 (row = tile / set, column = xo), the K
//     dimension is permuted to match and split over the four pixel rows (= four warps) of a tile; 3xTF32
//     (a_lo*b_hi + a_hi*b_lo + a_hi*b_hi) keeps fp32-class accuracy.  12 MMAs per warp replace 128 FFMA + 256
//     shared-memory reads per thread;
//   * 128 registers, no spills (2 CTAs of 256 threads per SM).
//
// A window that does not fit the shared-memory budget (hypotheses of one CTA more than ~150 columns apart) is read in
// place from global memory with the same code.
#include <cuda.h>

#include "common.cuh"
#include "warp_sample.cuh"

namespace {

constexpr int W2_TILES = 16;                 // tile columns per CTA
constexpr int W2_PXW = W2_TILES * 4;         // pixel columns per CTA
constexpr int W2_THREADS = W2_PXW * 4;       // one thread per pixel of the 4-row strip
constexpr int W2_BOXW = 32;                  // window columns per TMA box
constexpr int W2_WIN_BYTES = 88 * 1024;      // staged window budget (with ~19 KB static: two CTAs per SM)

struct W2P {
    const float* fl;
    const float* fr;      // NHWC [n][H][W][ldfr]
    int ldfl, ldfr;
    int fl_v8;            // left-feature rows allow 256-bit loads (pitch and base 32-byte aligned)
    const float* cur;
    int ldc;
    const float* prev;
    int ldp;
    const float* dec_w;   // [16][64]
    const float* dec_b;
    int N, h, w;
    float* aug;
    int ldaug;
    float* raw;
};

__device__ __forceinline__ uint32_t w2_s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t tf32_hi(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}

// D += A(16x8, row) * B(8x8, col), tf32 in, fp32 accumulate
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// ---------------------------------------------------------------------------------------------------------------
// Channel loops.  Bit-exactness: per channel  v = fma(se,wD, fma(sw,wC, fma(ne,wB, nw*wA)))  (torch's bilinear order),
// cost += |l - v| sequentially over channels.  The staged path evaluates -v with NEGATED weights (round-to-nearest is
// sign-symmetric, so fma(b,-wB, a*(-wA)) == -fma(b,wB, a*wA) bit for bit) and forms l - v as one packed add: per two
// channels 3 packed instructions (FMUL2, FFMA2, FADD2; sm_100 issues two independent fp32 operations per slot) plus
// the two sequential |.| accumulations, instead of 8 scalar instructions.
// ---------------------------------------------------------------------------------------------------------------
struct W2W {              // negated tap weights of one plane, duplicated for the packed instructions
    float2 a, b, c, d;    // -nw, -ne, -sw, -se
};

template <bool TWO_ROWS>
__device__ __forceinline__ void w2_acc4(float& acc, const float4& a4, const float4& b4, const float4& c4, const float4& d4,
                                        const W2W& w, const float2& l01, const float2& l23) {
    float2 v01 = __ffma2_rn(make_float2(b4.x, b4.y), w.b, __fmul2_rn(make_float2(a4.x, a4.y), w.a));
    float2 v23 = __ffma2_rn(make_float2(b4.z, b4.w), w.b, __fmul2_rn(make_float2(a4.z, a4.w), w.a));
    if (TWO_ROWS) {
        v01 = __ffma2_rn(make_float2(d4.x, d4.y), w.d, __ffma2_rn(make_float2(c4.x, c4.y), w.c, v01));
        v23 = __ffma2_rn(make_float2(d4.z, d4.w), w.d, __ffma2_rn(make_float2(c4.z, c4.w), w.c, v23));
    }
    const float2 e01 = __fadd2_rn(l01, v01), e23 = __fadd2_rn(l23, v23);     // l - v
    acc = __fadd_rn(acc, fabsf(e01.x));
    acc = __fadd_rn(acc, fabsf(e01.y));
    acc = __fadd_rn(acc, fabsf(e23.x));
    acc = __fadd_rn(acc, fabsf(e23.y));
}

// One hypothesis set from the STAGED window.  `base` points at window element (row y0, column xlo, channel 0) minus
// xlo columns, i.e. column x lives at base + x * PITCH; taps outside the image read the zeros TMA filled in, so no
// clamping and no validity predicates.  NOMINAL: the three planes share the 4-column register window that starts at
// x0[2] (k = +1, the left-most plane): plane ki uses window columns (2 - ki, 3 - ki).
template <int C, bool TWO_ROWS, bool NOMINAL>
__device__ __forceinline__ void w2_set_costs_staged(const float* __restrict__ base, int rowstep, const int (&x0)[3],
                                                    const W2W (&w)[3], const float2 (&lp)[C / 2], float (&cost)[3]) {
    constexpr int PITCH = C + 4;
    cost[0] = cost[1] = cost[2] = 0.f;
    if (NOMINAL) {
        const float* p0 = base + x0[2] * PITCH;
#pragma unroll
        for (int c = 0; c < C; c += 4) {
            float4 win[4], win2[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                win[j] = *reinterpret_cast<const float4*>(p0 + j * PITCH + c);
                if (TWO_ROWS) win2[j] = *reinterpret_cast<const float4*>(p0 + j * PITCH + rowstep + c);
                else win2[j] = win[j];
            }
#pragma unroll
            for (int ki = 0; ki < 3; ++ki)
                w2_acc4<TWO_ROWS>(cost[ki], win[2 - ki], win[3 - ki], win2[2 - ki], win2[3 - ki], w[ki], lp[c / 2], lp[c / 2 + 1]);
        }
    } else {
        const float* pa[3];
#pragma unroll
        for (int ki = 0; ki < 3; ++ki) pa[ki] = base + x0[ki] * PITCH;
#pragma unroll
        for (int c = 0; c < C; c += 4) {
#pragma unroll
            for (int ki = 0; ki < 3; ++ki) {
                const float4 a4 = *reinterpret_cast<const float4*>(pa[ki] + c);
                const float4 b4 = *reinterpret_cast<const float4*>(pa[ki] + PITCH + c);
                float4 c4 = a4, d4 = b4;
                if (TWO_ROWS) {
                    c4 = *reinterpret_cast<const float4*>(pa[ki] + rowstep + c);
                    d4 = *reinterpret_cast<const float4*>(pa[ki] + PITCH + rowstep + c);
                }
                w2_acc4<TWO_ROWS>(cost[ki], a4, b4, c4, d4, w[ki], lp[c / 2], lp[c / 2 + 1]);
            }
        }
    }
}

// The same set read IN PLACE from global memory (window not staged): clamped columns, zero WEIGHT for taps outside the
// image (identical cost: torch's blend then only adds +-0 for them).
template <int C, bool TWO_ROWS>
__device__ __forceinline__ void w2_set_costs_global(const float* __restrict__ rowp, int pitch, int rowstep, int W, bool row1_ok,
                                                    float fs, float fn, const int (&x0)[3], const float (&fw)[3],
                                                    const float2 (&lp)[C / 2], float (&cost)[3]) {
    cost[0] = cost[1] = cost[2] = 0.f;
#pragma unroll
    for (int ki = 0; ki < 3; ++ki) {
        const int xa = x0[ki];
        const bool va = (xa >= 0 && xa < W), vb = (xa + 1 >= 0 && xa + 1 < W);
        const float fe = __fsub_rn(1.f, fw[ki]);
        const float* pa = rowp + min(max(xa, 0), W - 1) * pitch;
        const float* pb = rowp + min(max(xa + 1, 0), W - 1) * pitch;
        W2W w;
        const float wa = va ? -__fmul_rn(fs, fe) : 0.f, wb = vb ? -__fmul_rn(fs, fw[ki]) : 0.f;
        const float wc = (va && row1_ok) ? -__fmul_rn(fn, fe) : 0.f, wd = (vb && row1_ok) ? -__fmul_rn(fn, fw[ki]) : 0.f;
        w.a = make_float2(wa, wa); w.b = make_float2(wb, wb); w.c = make_float2(wc, wc); w.d = make_float2(wd, wd);
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < C; c += 4) {
            const float4 a4 = ldg4(pa + c), b4 = ldg4(pb + c);
            float4 c4 = a4, d4 = b4;
            if (TWO_ROWS) { c4 = ldg4(pa + rowstep + c); d4 = ldg4(pb + rowstep + c); }
            w2_acc4<TWO_ROWS>(acc, a4, b4, c4, d4, w, lp[c / 2], lp[c / 2 + 1]);
        }
        cost[ki] = acc;
    }
}

// ---- window of one work item (a strip of 16 tiles x 4 pixel rows), computed by ONE warp from the 16 x NSETS plane
// hypotheses alone so that it can be issued ahead of the per-pixel work.  Conservative in x: a plane samples
// ld = ((d + k) + a dx) + b dy with |a|, |b| <= 1.5, |k| <= 1, so x - ld over the tile lies in
// [4j - d - 1 - 1.5(|dx|+|dy|), 4j + 3 - d + 1 + 1.5(|dx|+|dy|)]; one more column on each side covers the rounding of the
// coordinate chain and the floor() jitter, one more on the right the second tap.  The window is NOT clamped to the
// image: columns -2 .. W+1 and row H are legal TMA coordinates and arrive as zeros, which is exactly what a
// zeros-padded bilinear tap contributes.  Exact in y (the row chain depends on y only).  Every warp still checks its own
// taps against the window and reads in place from global memory if one falls outside, so the estimate can only cost
// time, never correctness.
struct W2Meta {
    int xlo, rlo, wc, rwin;
    int staged;
};

// Window producer, step 1: ISSUE the load of one plane hypothesis per lane (lane < 16: current hypothesis of tile
// j0 + lane; lane >= 16: previous-level hypothesis under tile j0 + lane - 16).  Nothing here consumes the loaded value,
// so the warp does not wait on memory before the CTA's early barrier.
template <int NSETS>
__device__ __forceinline__ float4 w2_item_hyps_load(const W2P& p, int n, int i, int j0, int lane) {
    const int j = j0 + (lane & 15);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j < p.w) {
        if (lane < 16) {
            v = ldg4(p.cur + (((size_t)n * p.h + i) * p.w + j) * p.ldc);
        } else if (NSETS == 2) {
            const int hp = p.h >> 1, wp = p.w >> 1;
            v = ldg4(p.prev + (((size_t)n * hp + (i >> 1)) * wp + (j >> 1)) * p.ldp);
        }
    }
    return v;
}
// step 2: disparity at the tile centre and |dx| + |dy| of that plane (sl < 0: no plane on this lane)
template <int NSETS>
__device__ __forceinline__ void w2_item_hyps(const W2P& p, const float4& v, int i, int j0, int lane, float& d, float& sl) {
    const int j = j0 + (lane & 15);
    d = 0.f;
    sl = -1.f;
    if (j >= p.w) return;
    if (lane < 16) {
        d = v.x;
        sl = fabsf(v.y) + fabsf(v.z);
    } else if (NSETS == 2) {
        const float cx = (float)(j & 1) - 0.5f, cy = (float)(i & 1) - 0.5f;
        d = __fmul_rn(__fadd_rn(__fadd_rn(v.x, __fmul_rn(cx, v.y)), __fmul_rn(cy, v.z)), 2.f);
        sl = fabsf(v.y) + fabsf(v.z);
    }
}

template <int C>
__device__ __forceinline__ void w2_produce(const CUtensorMap* tmap, const W2P& p, int n, int i, int j0, float d, float sl,
                                           W2Meta* meta, float* s_win, uint32_t bar, int lane) {
    constexpr int PITCH = C + 4;
    const int H = 4 * p.h, W = 4 * p.w;
    float lo = 3.0e9f, hi = -3.0e9f;
    int rlo = 0x7fffffff, rhi = -1;
    if (sl >= 0.f || !(sl == sl)) {
        const int j = j0 + (lane & 15);
        const float m = 1.5f * sl + 1.f;
        lo = floorf((float)(4 * j) - d - m - 2.f);          // floor, minus the jitter column
        hi = floorf((float)(4 * j + 3) - d + m + 3.f);      // plus jitter and second tap
        // (x0 is clamped to [-2, W] by the sampling set-up; NaN hypotheses: the whole row, i.e. read in place)
        lo = fminf(fmaxf(lo, -2.f), (float)W);
        hi = fmaxf(fminf(hi, (float)(W + 1)), -1.f);
        if (!(d == d) || !(sl == sl)) { lo = -2.f; hi = (float)(W + 1); }
    }
    if (lane < 4) {
        // rows of the strip: the same normalise / un-normalise round trip as the pixels
        const int y = 4 * i + lane;
        const float hm1 = (float)(H - 1), hdiv = (float)max(H - 1, 1);
        const float gy = __fadd_rn(__fdiv_rn(__fmul_rn(2.f, (float)y), hdiv), -1.f);
        const float iy = __fmul_rn(__fmul_rn(__fadd_rn(gy, 1.f), 0.5f), hm1);
        const float fy = floorf(iy);
        const int y0 = min(max((int)fy, 0), H - 1);
        rlo = y0;
        rhi = (__fsub_rn(iy, fy) != 0.f) ? y0 + 1 : y0;     // row H (below the image) is zero-filled by TMA
    }
    const int xlo = __reduce_min_sync(0xffffffffu, (int)lo);
    const int xhi = __reduce_max_sync(0xffffffffu, (int)hi);
    rlo = __reduce_min_sync(0xffffffffu, rlo);
    rhi = __reduce_max_sync(0xffffffffu, rhi);
    if (lane == 0) {
        const int nbox = (xhi - xlo + W2_BOXW) / W2_BOXW;
        const int wc = nbox * W2_BOXW, rwin = rhi - rlo + 1;
        const bool staged = xhi >= xlo && (size_t)rwin * wc * PITCH * sizeof(float) <= (size_t)W2_WIN_BYTES;
        meta->xlo = xlo; meta->rlo = rlo; meta->wc = wc; meta->rwin = rwin; meta->staged = staged ? 1 : 0;
        if (staged) {
            const uint32_t box_bytes = W2_BOXW * PITCH * (uint32_t)sizeof(float);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(box_bytes * rwin * nbox)
                         : "memory");
            for (int r = 0; r < rwin; ++r)
                for (int bx = 0; bx < nbox; ++bx) {
                    const uint32_t dst = w2_s32(s_win + (size_t)(r * wc + bx * W2_BOXW) * PITCH);
                    asm volatile(
                        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, "
                        "%5, %6}], [%2];" ::"r"(dst),
                        "l"(tmap), "r"(bar), "r"(0), "r"(xlo + bx * W2_BOXW), "r"(rlo + r), "r"(n)
                        : "memory");
                }
        } else {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");   // nothing to wait for
        }
    }
}

// One CTA per strip of 16 tiles x 4 pixel rows (grid = column blocks x tile rows x samples).  Every global load a thread
// needs is issued first; warp 0 then derives the window from the 32 plane hypotheses of the strip and issues its TMA
// copies, which travel while all warps run their per-pixel sampling set-up.
template <int NSETS, int C>
__global__ void __launch_bounds__(W2_THREADS, 2) tile_warp_cost2_kernel(const __grid_constant__ CUtensorMap tmap, W2P p) {
    constexpr int PITCH = C + 4;
    __shared__ __align__(16) float s_part[4][NSETS][W2_TILES][20];   // per pixel-row partial sums of `decrease` (padded)
    __shared__ __align__(16) uint32_t s_wh[16][68], s_wl[16][68];   // `decrease` weights [co][ci] as tf32 hi / lo parts
    __shared__ W2Meta s_meta;
    __shared__ __align__(8) unsigned long long s_bar, s_bar_w;     // window landed / `decrease` weights staged
    extern __shared__ __align__(128) float4 w2_dyn[];
    float* s_win = reinterpret_cast<float*>(w2_dyn);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t bar = w2_s32(&s_bar), bar_w = w2_s32(&s_bar_w);
    const int j0 = blockIdx.x * W2_TILES, i = blockIdx.y, n = blockIdx.z;
    // warp = (pixel row yo, half of the strip): lane = 4 * (tile within the half) + xo
    const int yo = warp >> 1;
    const int tl = ((warp & 1) << 3) + (lane >> 2), xo = lane & 3;
    const int H = 4 * p.h, W = 4 * p.w;
    const int j = j0 + tl;
    const int y = 4 * i + yo, x = 4 * j + xo;
    const bool on = j < p.w;

    // ---- every independent global load first: `decrease` weights (4 per thread), left features, hypotheses
    const float4 w4 = ldg4(p.dec_w + tid * 4);
    float2 lp[C / 2];
    float4 c4 = make_float4(0.f, 0.f, 0.f, 0.f), q4 = c4;
    if (on) {
        const float* flp = p.fl + (((size_t)n * H + y) * W + x) * p.ldfl;
        if (p.fl_v8) {           // whole 32-byte sectors per thread (see ldg8)
#pragma unroll
            for (int c = 0; c < C; c += 8) {
                float l8[8];
                ldg8(flp + c, l8);
#pragma unroll
                for (int e = 0; e < 4; ++e) lp[c / 2 + e] = make_float2(l8[2 * e], l8[2 * e + 1]);
            }
        } else {
#pragma unroll
            for (int c = 0; c < C; c += 4) {
                const float4 l4 = ldg4(flp + c);
                lp[c / 2] = make_float2(l4.x, l4.y);
                lp[c / 2 + 1] = make_float2(l4.z, l4.w);
            }
        }
        c4 = ldg4(p.cur + (((size_t)n * p.h + i) * p.w + j) * p.ldc);
        if (NSETS == 2) {
            const int hp = p.h >> 1, wp = p.w >> 1;
            q4 = ldg4(p.prev + (((size_t)n * hp + (i >> 1)) * wp + (j >> 1)) * p.ldp);
        }
    } else {
#pragma unroll
        for (int c = 0; c < C / 2; ++c) lp[c] = make_float2(0.f, 0.f);
    }
    float4 ph = make_float4(0.f, 0.f, 0.f, 0.f);
    if (warp == 0) {
        if (lane == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_w), "n"(W2_THREADS));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        ph = w2_item_hyps_load<NSETS>(p, n, i, j0, lane);
    }
    // The only CTA-wide barrier before the output phase, placed where nobody waits on memory yet: it publishes the
    // mbarrier.  From here on every warp runs on its own (set-up, window wait, channel loops).
    __syncthreads();
    if (warp == 0) {
        float pd, psl;
        w2_item_hyps<NSETS>(p, ph, i, j0, lane, pd, psl);
        w2_produce<C>(&tmap, p, n, i, j0, pd, psl, &s_meta, s_win, bar, lane);
    }

    // ---- per-pixel sampling state of every set (compact: floor column and fraction per plane)
    const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
    const float wdiv = (float)max(W - 1, 1), hdiv = (float)max(H - 1, 1);
    const float wrcp = __frcp_rn(wdiv);
    int x0[NSETS][3] = {};
    float fw[NSETS][3] = {};
    bool two_rows = false;
    int y0 = 0;
    float fn = 0.f, fs = 1.f;
    if (on) {
        // row coordinate through the same normalise / un-normalise round trip
        const float gy = __fadd_rn(__fdiv_rn(__fmul_rn(2.f, (float)y), hdiv), -1.f);
        const float iy = __fmul_rn(__fmul_rn(__fadd_rn(gy, 1.f), 0.5f), hm1);
        const float fy = floorf(iy);
        fn = __fsub_rn(iy, fy);
        fs = __fsub_rn(1.f, fn);
        y0 = min(max((int)fy, 0), H - 1);
        two_rows = fn != 0.f;                 // uniform per image row => per warp
        const float a = (float)xo - 1.5f, bb = (float)yo - 1.5f;
        Taps tp;
        sample_setup(c4.x, c4.y, c4.z, a, bb, x, wm1, wdiv, wrcp, tp);
#pragma unroll
        for (int ki = 0; ki < 3; ++ki) { x0[0][ki] = tp.x0[ki]; fw[0][ki] = tp.fw[ki]; }
        if (NSETS == 2) {
            const float cx = (float)(j & 1) - 0.5f, cy = (float)(i & 1) - 0.5f;
            const float du = __fmul_rn(__fadd_rn(__fadd_rn(q4.x, __fmul_rn(cx, q4.y)), __fmul_rn(cy, q4.z)), 2.f);
            sample_setup(du, q4.y, q4.z, a, bb, x, wm1, wdiv, wrcp, tp);
#pragma unroll
            for (int ki = 0; ki < 3; ++ki) { x0[NSETS - 1][ki] = tp.x0[ki]; fw[NSETS - 1][ki] = tp.fw[ki]; }
        }
    }
    float lnorm = 0.f;
#pragma unroll
    for (int c = 0; c < C / 2; ++c) {
        lnorm = __fadd_rn(lnorm, fabsf(lp[c].x));
        lnorm = __fadd_rn(lnorm, fabsf(lp[c].y));
    }
    {
        // this thread's four `decrease` weights as tf32 hi / lo parts (consumed after the channel loops)
        const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            hi[e] = tf32_hi(wv[e]);
            lo[e] = tf32_hi(wv[e] - __uint_as_float(hi[e]));
        }
        const int co = tid >> 4, ci = (tid & 15) * 4;
        *reinterpret_cast<uint4*>(&s_wh[co][ci]) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(&s_wl[co][ci]) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        // split barrier (an mbarrier counting all threads): arrive now, wait right before the fragments are read, after
        // the channel loops — by then every thread has arrived long ago and nobody waits
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_w) : "memory");
    }

    // ---- wait for the window
    {
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(bar), "r"(0)
                : "memory");
        }
    }
    // (the window description was written by the thread that armed the mbarrier, before it did: visible after the wait)
    const int xlo = s_meta.xlo, rlo = s_meta.rlo, wc = s_meta.wc;
    const bool win_staged = s_meta.staged != 0;
    const bool rows_in = y0 >= rlo && y0 + (two_rows ? 1 : 0) < rlo + s_meta.rwin;

    // ---- costs per set
    float cost[NSETS][3];
#pragma unroll
    for (int s = 0; s < NSETS; ++s) {
        bool inwin = rows_in;
#pragma unroll
        for (int ki = 0; ki < 3; ++ki) inwin = inwin && x0[s][ki] >= xlo && x0[s][ki] + 1 < xlo + wc;
        // nominal alignment: plane k+1 samples exactly one column left of plane k
        const bool nominal = __all_sync(0xffffffffu, ((x0[s][0] == x0[s][1] + 1) && (x0[s][1] == x0[s][2] + 1)) || !on);
        const bool smem_ok = __all_sync(0xffffffffu, (win_staged && inwin) || !on);
        cost[s][0] = cost[s][1] = cost[s][2] = 0.f;
        if (on) {
            if (smem_ok) {
                W2W w[3];
#pragma unroll
                for (int ki = 0; ki < 3; ++ki) {
                    const float fe = __fsub_rn(1.f, fw[s][ki]);
                    const float wa = -__fmul_rn(fs, fe), wb = -__fmul_rn(fs, fw[s][ki]);
                    const float wc_ = -__fmul_rn(fn, fe), wd = -__fmul_rn(fn, fw[s][ki]);
                    w[ki].a = make_float2(wa, wa); w[ki].b = make_float2(wb, wb);
                    w[ki].c = make_float2(wc_, wc_); w[ki].d = make_float2(wd, wd);
                }
                const float* base = s_win + (ptrdiff_t)((y0 - rlo) * wc - xlo) * PITCH;
                const int rowstep = wc * PITCH;
                if (nominal) {
                    if (two_rows) w2_set_costs_staged<C, true, true>(base, rowstep, x0[s], w, lp, cost[s]);
                    else w2_set_costs_staged<C, false, true>(base, rowstep, x0[s], w, lp, cost[s]);
                } else {
                    if (two_rows) w2_set_costs_staged<C, true, false>(base, rowstep, x0[s], w, lp, cost[s]);
                    else w2_set_costs_staged<C, false, false>(base, rowstep, x0[s], w, lp, cost[s]);
                }
            } else {
                const bool row1_ok = y0 + 1 < H;
                const float* rowp = p.fr + ((size_t)n * H + y0) * (size_t)W * p.ldfr;
                const int rowstep = row1_ok ? W * p.ldfr : 0;
                if (two_rows) w2_set_costs_global<C, true>(rowp, p.ldfr, rowstep, W, row1_ok, fs, fn, x0[s], fw[s], lp, cost[s]);
                else w2_set_costs_global<C, false>(rowp, p.ldfr, rowstep, W, row1_ok, fs, fn, x0[s], fw[s], lp, cost[s]);
            }
        }
    }

    // ---- raw costs for the tests / training API: [tile][set][q * 16 + yo * 4 + xo], q = 0 (|fL|_1), 1..3 (k = -1, 0, +1)
    if (p.raw && on) {
        float* rp = p.raw + (((size_t)n * p.h + i) * p.w + j) * (NSETS * 64) + yo * 4 + xo;
#pragma unroll
        for (int s = 0; s < NSETS; ++s) {
            rp[s * 64] = lnorm;
#pragma unroll
            for (int ki = 0; ki < 3; ++ki) rp[s * 64 + 16 + ki * 16] = cost[s][ki];
        }
    }

    // ---- output phase, part 1 (its global loads travel behind the tensor-core tail): one float4 granule per thread of the
    // augmented hypothesis tensor [cur | cur_cv | up_prev | prev_cv] of the strip's 16 tiles
    constexpr int nf4 = NSETS * 8;            // float4 granules per tile
    static_assert(W2_TILES * nf4 <= W2_THREADS, "one output granule per thread");
    const int ot = tid / nf4, of = tid - ot * nf4;
    const int ojj = j0 + ot;
    const bool oon = tid < W2_TILES * nf4 && ojj < p.w;
    const size_t otpix = ((size_t)n * p.h + i) * p.w + ojj;
    float4 ov = make_float4(0.f, 0.f, 0.f, 0.f);
    if (oon) {
        if (of < 4) {
            ov = ldg4(p.cur + otpix * p.ldc + of * 4);
        } else if (of < 8 || of >= 12) {
            ov = ldg4(p.dec_b + (of & 3) * 4);
        } else {
            const int hp = p.h >> 1, wp = p.w >> 1;
            ov = ldg4(p.prev + (((size_t)n * hp + (i >> 1)) * wp + (ojj >> 1)) * p.ldp + (of - 8) * 4);
        }
    }

    // ---- `decrease` on the tensor cores.  A rows 0..7: set 0 of the warp's 8 tiles, rows 8..15: the last set;
    // this warp contributes the 16 features of its pixel row yo (two k-steps of 8).  B fragments: k-step s covers
    // feature blocks q = 2s (k index t) and 2s+1 (t + 4), feature index ci = 16 q + 4 yo + xo; n-tile nt covers
    // output channels 8 nt + g.
    {
        uint32_t done = 0;      // `decrease` weights of every thread are in shared memory
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(bar_w), "r"(0)
                : "memory");
        }
    }
    {
        const int g = lane >> 2, t = lane & 3;
        float acc[2][4];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            // feature blocks q = 2s, 2s+1:  q = 0 -> |fL|_1, q = 1..3 -> cost[.][q - 1]
            float v[4];
            v[0] = on ? (s == 0 ? lnorm : cost[0][1]) : 0.f;                   // set 0, q = 2s      (row g,     col t)
            v[1] = on ? (s == 0 ? lnorm : cost[NSETS - 1][1]) : 0.f;           // last set, q = 2s   (row g + 8, col t)
            v[2] = on ? (s == 0 ? cost[0][0] : cost[0][2]) : 0.f;              // set 0, q = 2s + 1  (row g,     col t + 4)
            v[3] = on ? (s == 0 ? cost[NSETS - 1][0] : cost[NSETS - 1][2]) : 0.f;
            uint32_t ah[4], al[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                ah[r] = tf32_hi(v[r]);
                al[r] = tf32_hi(v[r] - __uint_as_float(ah[r]));
            }
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                const int ci = 32 * s + 4 * yo + t;
                const uint32_t bh0 = s_wh[8 * nt + g][ci], bh1 = s_wh[8 * nt + g][ci + 16];
                const uint32_t bl0 = s_wl[8 * nt + g][ci], bl1 = s_wl[8 * nt + g][ci + 16];
                mma_tf32(acc[nt], al, bh0, bh1);
                mma_tf32(acc[nt], ah, bl0, bl1);
                mma_tf32(acc[nt], ah, bh0, bh1);
            }
        }
        // C fragment: c0,c1 -> (row g, cols 2t, 2t+1), c2,c3 -> (row g + 8, ...)
        const int tg = ((warp & 1) << 3) + g;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            *reinterpret_cast<float2*>(&s_part[yo][0][tg][8 * nt + 2 * t]) = make_float2(acc[nt][0], acc[nt][1]);
            if (NSETS == 2)
                *reinterpret_cast<float2*>(&s_part[yo][NSETS - 1][tg][8 * nt + 2 * t]) = make_float2(acc[nt][2], acc[nt][3]);
        }
    }
    __syncthreads();

    // ---- output phase, part 2
    if (oon) {
        if (of >= 4 && (of < 8 || of >= 12)) {
            const int s = of < 8 ? 0 : NSETS - 1, co = (of & 3) * 4;
            float r[4] = {ov.x, ov.y, ov.z, ov.w};       // bias
#pragma unroll
            for (int yy = 0; yy < 4; ++yy) {          // fixed order: deterministic
                const float4 q = *reinterpret_cast<const float4*>(&s_part[yy][s][ot][co]);
                r[0] += q.x; r[1] += q.y; r[2] += q.z; r[3] += q.w;
            }
            ov = make_float4(codd_act(r[0], CODD_ACT_LEAKY, 0), codd_act(r[1], CODD_ACT_LEAKY, 0),
                             codd_act(r[2], CODD_ACT_LEAKY, 0), codd_act(r[3], CODD_ACT_LEAKY, 0));
        } else if (of == 8) {
            const float cx = (float)(ojj & 1) - 0.5f, cy = (float)(i & 1) - 0.5f;
            ov.x = __fmul_rn(__fadd_rn(__fadd_rn(ov.x, __fmul_rn(cx, ov.y)), __fmul_rn(cy, ov.z)), 2.f);
        }
        *reinterpret_cast<float4*>(p.aug + otpix * p.ldaug + of * 4) = ov;
    }
}

typedef CUresult (*PFN_w2EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_w2EncodeTiled w2_get_encode() {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
        return (PFN_w2EncodeTiled)ptr;
    return nullptr;
}

template <int NSETS, int C>
int w2_launch(const CUtensorMap& tmap, const W2P& p, cudaStream_t s) {
    auto kern = tile_warp_cost2_kernel<NSETS, C>;
    static CoddDeviceOnce once;
    if (int rc = codd_once_per_device(once, [&] {
            return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, W2_WIN_BYTES);
        }))
        return rc;
    const int nblk = codd_ceil_div(p.w, W2_TILES);
    if (p.h > 65535 || p.N > 65535) return CODD_E_SHAPE;
    kern<<<dim3((unsigned)nblk, (unsigned)p.h, (unsigned)p.N), W2_THREADS, W2_WIN_BYTES, s>>>(tmap, p);
    CODD_RETURN_IF_CUDA_ERROR();
    return 0;
}

}  // namespace

// Entry point behind codd_tile_warp_cost_nhwc for C in {16, 24, 32} (tile.cu dispatches here); returns
// CODD_E_UNSUPPORTED for other channel counts so that the caller can fall back to the first kernel.
int codd_tile_warp_cost2(const float* fea_l, int ldfl, const float* fea_r, int ldfr, int c, const float* cur, int ldc,
                         const float* prev, int ldp, const float* dec_w, const float* dec_b, int n, int h, int w,
                         float* aug, int ldaug, float* raw_cv, void* stream) {
    if (c != 16 && c != 24 && c != 32) return CODD_E_UNSUPPORTED;
    static PFN_w2EncodeTiled enc = w2_get_encode();      // C++11 magic static: initialised once, thread-safe
    if (!enc) return CODD_E_UNSUPPORTED;
    const int H = 4 * h, W = 4 * w;
    CUtensorMap tmap;
    const cuuint64_t gdim[4] = {(cuuint64_t)c, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n};
    const cuuint64_t gstr[3] = {(cuuint64_t)ldfr * 4, (cuuint64_t)W * ldfr * 4, (cuuint64_t)H * W * ldfr * 4};
    const cuuint32_t box[4] = {(cuuint32_t)(c + 4), (cuuint32_t)W2_BOXW, 1u, 1u};
    const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)fea_r, gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return CODD_E_UNSUPPORTED;
    W2P p;
    p.fl = fea_l; p.ldfl = ldfl; p.fr = fea_r; p.ldfr = ldfr;
    p.fl_v8 = (c % 8 == 0) && (ldfl % 8 == 0) && codd_aligned32(fea_l);
    p.cur = cur; p.ldc = ldc; p.prev = prev; p.ldp = ldp;
    p.dec_w = dec_w; p.dec_b = dec_b; p.N = n; p.h = h; p.w = w;
    p.aug = aug; p.ldaug = ldaug; p.raw = raw_cv;
    cudaStream_t s = (cudaStream_t)stream;
    if (prev) {
        if (c == 16) return w2_launch<2, 16>(tmap, p, s);
        if (c == 24) return w2_launch<2, 24>(tmap, p, s);
        return w2_launch<2, 32>(tmap, p, s);
    }
    if (c == 16) return w2_launch<1, 16>(tmap, p, s);
    if (c == 24) return w2_launch<1, 24>(tmap, p, s);
    return w2_launch<1, 32>(tmap, p, s);
}
