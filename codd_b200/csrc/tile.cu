// Tile-hypothesis kernels of HITNet's propagation stage (NHWC fp32).
//
//   K3  plane_upsample     propagation.py:10-32   (to_plane / upsample)
//   K4  tile_warp_cost     propagation.py:35-86 (warp, TileWarping), :156-160, :206-219
//   K5  hyp_select         propagation.py:225-248
//   --  tile_hyp_init      initialization.py:186-208 (descriptor 1x1 conv + hypothesis concat)
//
// The discrete decisions downstream (arg-max of the two confidences) amplify ulp noise, so the
// sampling arithmetic of K4 restates the reference + torch grid_sample rounding sequence exactly
// (see oracle/hitnet_oracle.py: warp_coords / warp_right_direct, which is bit-identical to
// F.grid_sample on CPU): explicit __f*_rn intrinsics, no FMA contraction except where torch's own
// kernel uses fused multiply-adds (the 4-tap blend).
#include <cstdlib>

#include "common.cuh"
#include "warp_sample.cuh"

namespace {

// ------------------------------------------------------------------------------------------
// tile_hyp_init
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) tile_hyp_init_kernel(const float* __restrict__ min_cost,
                                                            const float* __restrict__ min_disp,
                                                            const float* __restrict__ feat, int ldf, int cf,
                                                            const float* __restrict__ wgt,
                                                            const float* __restrict__ bias, size_t npix,
                                                            size_t planar_hw, float* __restrict__ hyp, int ldh) {
    extern __shared__ float4 smem4[];
    float* s_w = reinterpret_cast<float*>(smem4);  // [1+cf][16] transposed, 13 used
    const int cin = 1 + cf;
    for (int i = threadIdx.x; i < cin * 16; i += blockDim.x) {
        const int co = i & 15, ci = i >> 4;
        s_w[i] = co < 13 ? wgt[co * cin + ci] : 0.f;
    }
    __syncthreads();
    const size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= npix) return;
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = i < 13 ? __ldg(bias + i) : 0.f;
    const float c = __ldg(min_cost + pix);
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fmaf(c, s_w[i], acc[i]);
    const bool planar = (ldf == 0);
    const size_t hw = planar_hw;
    const float* fp = planar ? feat + (pix / hw) * (size_t)cf * hw + (pix % hw) : feat + pix * ldf;
    for (int ci = 0; ci < cf; ci += 4) {
        const float4 a4 = planar ? make_float4(__ldg(fp + (size_t)ci * hw), __ldg(fp + (size_t)(ci + 1) * hw),
                                               __ldg(fp + (size_t)(ci + 2) * hw), __ldg(fp + (size_t)(ci + 3) * hw))
                                 : ldg4(fp + ci);
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            const float a = cc == 0 ? a4.x : cc == 1 ? a4.y : cc == 2 ? a4.z : a4.w;
            const float* wp = s_w + (1 + ci + cc) * 16;
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = fmaf(a, wp[i], acc[i]);
        }
    }
    float o[16];
    o[0] = __ldg(min_disp + pix);
    o[1] = 0.f;
    o[2] = 0.f;
#pragma unroll
    for (int i = 0; i < 13; ++i) o[3 + i] = codd_act(acc[i], CODD_ACT_LEAKY, 0);
    float4* op = reinterpret_cast<float4*>(hyp + pix * ldh);
#pragma unroll
    for (int i = 0; i < 4; ++i) op[i] = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
}

// ------------------------------------------------------------------------------------------
// K3 plane_upsample
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) plane_upsample_kernel(const float* __restrict__ in, int ldi, int n, int h,
                                                             int w, int size, float scale,
                                                             float* __restrict__ out, int ldo) {
    const size_t total = (size_t)n * h * size * w * size * 4;  // float4 granules
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c4 = (int)(idx & 3);
    size_t pix = idx >> 2;
    const int ow = w * size, oh = h * size;
    const int x = (int)(pix % ow);
    const size_t t = pix / ow;
    const int y = (int)(t % oh);
    const int s = (int)(t / oh);
    const float* ip = in + (((size_t)s * h + y / size) * w + x / size) * ldi;
    float4 v = ldg4(ip + c4 * 4);
    if (c4 == 0) {
        const float half = 0.5f * (float)(size - 1);
        const float cx = (float)(x % size) - half;
        const float cy = (float)(y % size) - half;
        // (d + cx*dx) + cy*dy, then * scale  — multiply and add rounded separately
        const float d = __fadd_rn(__fadd_rn(v.x, __fmul_rn(cx, v.y)), __fmul_rn(cy, v.z));
        v.x = __fmul_rn(d, scale);
    }
    *reinterpret_cast<float4*>(out + pix * ldo + c4 * 4) = v;
}

// ------------------------------------------------------------------------------------------
// K4 tile_warp_cost
// ------------------------------------------------------------------------------------------
struct WarpP {
    const float* fl;
    const float* fr;      // PLANAR [n][C][H][W] (staged / per-channel gather paths), or NULL
    const float* fr_nhwc; // NHWC [n][H][W][ldfr] (direct 128-bit gather path), or NULL
    int ldfr;
    int ldfl, C;
    const float* cur;
    int ldc;
    const float* prev;
    int ldp;
    const float* dec_w;
    const float* dec_b;
    int N, h, w;
    float* aug;
    int ldaug;
    float* raw;
    int max_win;   // widest right-feature window (columns) worth staging in shared memory
};

constexpr int K4_TILES = 16;              // tile columns per CTA
constexpr int K4_PXW = K4_TILES * 4;      // pixel columns per CTA
constexpr int K4_THREADS = K4_PXW * 4;    // one thread per pixel of the 4-row strip

// The right features are read PLANAR ([n][C][H][W]): lanes of a warp are horizontally adjacent
// pixels, so every per-channel tap load is a (nearly) contiguous 128-byte request — one L1
// wavefront — where the NHWC layout costs 16 (lanes 64 B apart).  The left features stay NHWC:
// each lane reads its own pixel's channels once.
constexpr int K4_STAGE_BYTES = 58 * 1024;   // shared-memory budget of the staged right-feature window (2 CTAs / SM)

// Channel loop of K4 on a STAGED window: the CTA's right-feature window (all taps of its 64 x 4 pixels) sits in shared
// memory as [row][x][C + 4] (channel innermost, pitch C+4 floats => the 128-bit reads of 8 horizontally adjacent
// lanes fall on 32 distinct banks), so one LDS.128 fetches 4 channels of a tap and the address is one 32-bit add —
// versus one LDG + 64-bit address arithmetic per channel and tap on the gather path (k4_channels).  Same arithmetic,
// same order: bit-identical results.
template <int NSETS, bool TWO_ROWS>
__device__ __forceinline__ void k4_channels_staged(const float* __restrict__ flp, const float* s_R, int C, int rowstep,
                                                   const int (&offA)[NSETS][3], const int (&offB)[NSETS][3],
                                                   const float (&wA)[NSETS][3], const float (&wB)[NSETS][3],
                                                   const float (&wC)[NSETS][3], const float (&wD)[NSETS][3],
                                                   float (&cost)[NSETS][3], float& lnorm) {
    for (int c = 0; c < C; c += 4) {
        const float4 l4 = ldg4(flp + c);
        const float lv[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) lnorm = __fadd_rn(lnorm, fabsf(lv[cc]));
#pragma unroll
        for (int s = 0; s < NSETS; ++s)
#pragma unroll
            for (int ki = 0; ki < 3; ++ki) {
                const float4 a4 = *reinterpret_cast<const float4*>(s_R + offA[s][ki] + c);
                const float4 b4 = *reinterpret_cast<const float4*>(s_R + offB[s][ki] + c);
                const float ta[4] = {a4.x, a4.y, a4.z, a4.w}, tb[4] = {b4.x, b4.y, b4.z, b4.w};
                float ua[4], ub[4];
                if (TWO_ROWS) {
                    const float4 c4 = *reinterpret_cast<const float4*>(s_R + offA[s][ki] + rowstep + c);
                    const float4 d4 = *reinterpret_cast<const float4*>(s_R + offB[s][ki] + rowstep + c);
                    ua[0] = c4.x; ua[1] = c4.y; ua[2] = c4.z; ua[3] = c4.w;
                    ub[0] = d4.x; ub[1] = d4.y; ub[2] = d4.z; ub[3] = d4.w;
                }
                float acc = cost[s][ki];
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    float v = __fmaf_rn(tb[cc], wB[s][ki], __fmul_rn(ta[cc], wA[s][ki]));
                    if (TWO_ROWS) v = __fmaf_rn(ub[cc], wD[s][ki], __fmaf_rn(ua[cc], wC[s][ki], v));
                    acc = __fadd_rn(acc, fabsf(__fsub_rn(lv[cc], v)));
                }
                cost[s][ki] = acc;
            }
    }
}

template <int NSETS>
__global__ void __launch_bounds__(K4_THREADS, 3) tile_warp_cost_kernel(WarpP p) {
    __shared__ __align__(16) float s_raw[NSETS][K4_TILES][64];
    __shared__ __align__(16) float s_dec[NSETS][K4_TILES][16];
    __shared__ __align__(16) float s_wt[64][16];
    __shared__ float s_b[16];
    __shared__ int s_rng[4];   // window of right-feature columns / rows touched by this CTA: xlo, xhi, rlo, rhi
    extern __shared__ float4 k4_dyn[];
    float* s_R = reinterpret_cast<float*>(k4_dyn);

    const int tid = threadIdx.x;
    if (tid == 0) {
        s_rng[0] = 0x7fffffff; s_rng[1] = -1; s_rng[2] = 0x7fffffff; s_rng[3] = -1;
    }
    for (int i = tid; i < 64 * 16; i += K4_THREADS) {
        const int co = i & 15, ci = i >> 4;
        s_wt[ci][co] = __ldg(p.dec_w + co * 64 + ci);
    }
    if (tid < 16) s_b[tid] = __ldg(p.dec_b + tid);

    const int nblk = (p.w + K4_TILES - 1) / K4_TILES;
    int b = blockIdx.x;
    const int jblk = b % nblk;
    b /= nblk;
    const int i = b % p.h;
    const int n = b / p.h;
    const int j0 = jblk * K4_TILES;

    const int yo = tid / K4_PXW;
    const int tx = tid - yo * K4_PXW;
    const int tl = tx >> 2, xo = tx & 3;
    const int j = j0 + tl;
    const int H = 4 * p.h, W = 4 * p.w;
    const int y = 4 * i + yo, x = 4 * j + xo;
    const bool on = j < p.w;

    __syncthreads();   // s_rng initialised
    // per-pixel sampling state (registers; computed before the CTA-wide window reduction)
    float cost[NSETS][3];
    float wA[NSETS][3], wB[NSETS][3], wC[NSETS][3], wD[NSETS][3];
    int colA[NSETS][3], colB[NSETS][3];
    bool paired = true, two_rows = false, row1_ok = false;
    int y0 = 0;
    int lxlo = 0x7fffffff, lxhi = -1, lrlo = 0x7fffffff, lrhi = -1;   // this pixel's window (empty when the pixel is off)
    if (on) {
        const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
        const float wdiv = (float)max(W - 1, 1), hdiv = (float)max(H - 1, 1);
        const float wrcp = __frcp_rn(wdiv);
        // row coordinate through the same normalise / un-normalise round trip
        const float gy = __fadd_rn(__fdiv_rn(__fmul_rn(2.f, (float)y), hdiv), -1.f);
        const float iy = __fmul_rn(__fmul_rn(__fadd_rn(gy, 1.f), 0.5f), hm1);
        const float fy = floorf(iy);
        const float fn = __fsub_rn(iy, fy);
        const float fs = __fsub_rn(1.f, fn);
        y0 = min(max((int)fy, 0), H - 1);
        two_rows = fn != 0.f;                 // uniform per image row => per warp
        row1_ok = (y0 + 1 < H);

        const float a = (float)xo - 1.5f, bb = (float)yo - 1.5f;
        Taps tp[NSETS];
        {
            const float4 c4 = ldg4(p.cur + (((size_t)n * p.h + i) * p.w + j) * p.ldc);
            sample_setup(c4.x, c4.y, c4.z, a, bb, x, wm1, wdiv, wrcp, tp[0]);
        }
        if (NSETS == 2) {
            const int hp = p.h >> 1, wp = p.w >> 1;
            const float4 q4 = ldg4(p.prev + (((size_t)n * hp + (i >> 1)) * wp + (j >> 1)) * p.ldp);
            const float cx = (float)(j & 1) - 0.5f, cy = (float)(i & 1) - 0.5f;
            const float du = __fmul_rn(__fadd_rn(__fadd_rn(q4.x, __fmul_rn(cx, q4.y)), __fmul_rn(cy, q4.z)), 2.f);
            sample_setup(du, q4.y, q4.z, a, bb, x, wm1, wdiv, wrcp, tp[NSETS - 1]);
        }
        int xlo = 0x7fffffff, xhi = -1;
#pragma unroll
        for (int s = 0; s < NSETS; ++s) {
#pragma unroll
            for (int ki = 0; ki < 3; ++ki) {
                cost[s][ki] = 0.f;
                const int xa = tp[s].x0[ki];
                const bool va = (xa >= 0 && xa < W), vb = (xa + 1 >= 0 && xa + 1 < W);
                paired = paired && va && vb;
                colA[s][ki] = min(max(xa, 0), W - 1);
                colB[s][ki] = min(max(xa + 1, 0), W - 1);
                xlo = min(xlo, colA[s][ki]);
                xhi = max(xhi, colB[s][ki]);
                wA[s][ki] = va ? __fmul_rn(fs, tp[s].fe[ki]) : 0.f;
                wB[s][ki] = vb ? __fmul_rn(fs, tp[s].fw[ki]) : 0.f;
                wC[s][ki] = (va && row1_ok) ? __fmul_rn(fn, tp[s].fe[ki]) : 0.f;
                wD[s][ki] = (vb && row1_ok) ? __fmul_rn(fn, tp[s].fw[ki]) : 0.f;
            }
        }
        lxlo = xlo; lxhi = xhi; lrlo = y0; lrhi = (two_rows && row1_ok) ? y0 + 1 : y0;
    }
    // CTA-wide window: warp reductions first (256 same-address shared atomics serialise), one atomic per warp
    lxlo = __reduce_min_sync(0xffffffffu, lxlo);
    lxhi = __reduce_max_sync(0xffffffffu, lxhi);
    lrlo = __reduce_min_sync(0xffffffffu, lrlo);
    lrhi = __reduce_max_sync(0xffffffffu, lrhi);
    if ((tid & 31) == 0 && lxhi >= 0) {
        atomicMin(&s_rng[0], lxlo);
        atomicMax(&s_rng[1], lxhi);
        atomicMin(&s_rng[2], lrlo);
        atomicMax(&s_rng[3], lrhi);
    }
    __syncthreads();
    const int xlo = s_rng[0], rlo = s_rng[2];
    const int wwin = s_rng[1] - xlo + 1, rwin = s_rng[3] - rlo + 1;
    const int pitch = p.C + 4;
    const bool direct = p.fr_nhwc != nullptr;
    const bool staged = (s_rng[1] >= 0) && wwin <= p.max_win &&
                        ((size_t)wwin * rwin * pitch * sizeof(float) <= (size_t)K4_STAGE_BYTES);
    const size_t cstride = (size_t)H * W;
    const float* frn = direct ? p.fr_nhwc + (size_t)n * cstride * p.ldfr : p.fr + (size_t)n * p.C * cstride;
    if (staged && direct) {
        // NHWC source: the window is already channel-innermost in global memory — one 128-bit load + one 128-bit store
        // per (pixel, 4 channels), consecutive threads on consecutive 16-byte pieces of a pixel
        const int c4n = p.C >> 2;
        const int per_row = wwin * c4n;
        for (int idx = tid; idx < rwin * per_row; idx += K4_THREADS) {
            const int r = idx / per_row, rem = idx - r * per_row;
            const int xi = rem / c4n, c4 = rem - xi * c4n;
            const float4 v = ldg4(frn + ((size_t)(rlo + r) * W + xlo + xi) * p.ldfr + c4 * 4);
            *reinterpret_cast<float4*>(s_R + (size_t)(r * wwin + xi) * pitch + c4 * 4) = v;
        }
        __syncthreads();
    } else if (staged) {
        // planar global rows (coalesced along x) -> [row][x][C+4]: a warp takes one (4-channel group, row) segment at
        // a time; every lane loads 4 channel planes at its column and writes them as one 128-bit store
        const int nseg = (p.C >> 2) * rwin;
        const int warp = tid >> 5, lane = tid & 31;
        for (int seg = warp; seg < nseg; seg += K4_THREADS / 32) {
            const int c4 = seg / rwin, r = seg - c4 * rwin;
            const float* src = frn + (size_t)(c4 * 4) * cstride + (size_t)(rlo + r) * W + xlo;
            float* dst = s_R + (size_t)(r * wwin) * pitch + c4 * 4;
            for (int xi = lane; xi < wwin; xi += 32) {
                const float4 v = make_float4(__ldg(src + xi), __ldg(src + cstride + xi), __ldg(src + 2 * cstride + xi),
                                             __ldg(src + 3 * cstride + xi));
                *reinterpret_cast<float4*>(dst + (size_t)xi * pitch) = v;
            }
        }
        __syncthreads();
    }

    if (on) {
        const float* flp = p.fl + (((size_t)n * H + y) * W + x) * p.ldfl;
        int offA[NSETS][3], offB[NSETS][3];
        float lnorm = 0.f;
        if (direct && !staged) {
            // window too large for shared memory: right features read in place, NHWC: one 128-bit load fetches 4 channels of a tap (horizontally adjacent
            // lanes read adjacent pixels: contiguous 64..128-byte runs, L1-resident across the planes of a pixel)
            const int rbase = y0 * W;
#pragma unroll
            for (int s = 0; s < NSETS; ++s)
#pragma unroll
                for (int ki = 0; ki < 3; ++ki) {
                    offA[s][ki] = (rbase + colA[s][ki]) * p.ldfr;
                    offB[s][ki] = (rbase + colB[s][ki]) * p.ldfr;
                }
            const int rowstep = row1_ok ? W * p.ldfr : 0;
            if (two_rows) k4_channels_staged<NSETS, true>(flp, frn, p.C, rowstep, offA, offB, wA, wB, wC, wD, cost, lnorm);
            else k4_channels_staged<NSETS, false>(flp, frn, p.C, rowstep, offA, offB, wA, wB, wC, wD, cost, lnorm);
        } else if (staged) {
            const int rbase = (y0 - rlo) * wwin - xlo;
#pragma unroll
            for (int s = 0; s < NSETS; ++s)
#pragma unroll
                for (int ki = 0; ki < 3; ++ki) {
                    offA[s][ki] = (rbase + colA[s][ki]) * pitch;
                    offB[s][ki] = (rbase + colB[s][ki]) * pitch;
                }
            const int rowstep = row1_ok ? wwin * pitch : 0;
            if (two_rows) k4_channels_staged<NSETS, true>(flp, s_R, p.C, rowstep, offA, offB, wA, wB, wC, wD, cost, lnorm);
            else k4_channels_staged<NSETS, false>(flp, s_R, p.C, rowstep, offA, offB, wA, wB, wC, wD, cost, lnorm);
        } else {
            const int rowoff = y0 * W;
            const int rowstep = row1_ok ? W : 0;             // invalid second row: any address, zero weight
#pragma unroll
            for (int s = 0; s < NSETS; ++s)
#pragma unroll
                for (int ki = 0; ki < 3; ++ki) {
                    offA[s][ki] = rowoff + colA[s][ki];
                    offB[s][ki] = rowoff + colB[s][ki];
                }
            if (paired) {
                if (two_rows) k4_channels<NSETS, true, true>(flp, frn, cstride, p.C, rowstep, offA, offB, wA, wB, wC, wD, cost, lnorm);
                else k4_channels<NSETS, false, true>(flp, frn, cstride, p.C, rowstep, offA, offB, wA, wB, wC, wD, cost, lnorm);
            } else {
                if (two_rows) k4_channels<NSETS, true, false>(flp, frn, cstride, p.C, rowstep, offA, offB, wA, wB, wC, wD, cost, lnorm);
                else k4_channels<NSETS, false, false>(flp, frn, cstride, p.C, rowstep, offA, offB, wA, wB, wC, wD, cost, lnorm);
            }
        }
        const int po = yo * 4 + xo;
#pragma unroll
        for (int s = 0; s < NSETS; ++s) {
            s_raw[s][tl][po] = lnorm;
#pragma unroll
            for (int ki = 0; ki < 3; ++ki) s_raw[s][tl][16 + ki * 16 + po] = cost[s][ki];
        }
    }
    __syncthreads();

    // ---- `decrease`: 1x1 conv 64 -> 16 + LeakyReLU per tile and set
    for (int o = tid; o < NSETS * K4_TILES * 16; o += K4_THREADS) {
        const int co = o & 15;
        const int t = (o >> 4) % K4_TILES;
        const int s = o / (16 * K4_TILES);
        if (j0 + t < p.w) {
            float acc = s_b[co];
            const float* rp = s_raw[s][t];
#pragma unroll 16
            for (int ci = 0; ci < 64; ++ci) acc = fmaf(rp[ci], s_wt[ci][co], acc);
            s_dec[s][t][co] = codd_act(acc, CODD_ACT_LEAKY, 0);
        }
    }
    __syncthreads();

    // ---- write the augmented hypothesis tensor: [cur | cur_cv | up_prev | prev_cv]
    const int nf4 = NSETS * 8;  // float4 granules per tile
    for (int o = tid; o < K4_TILES * nf4; o += K4_THREADS) {
        const int t = o / nf4, f = o - t * nf4;
        const int jj = j0 + t;
        if (jj >= p.w) continue;
        const size_t tpix = ((size_t)n * p.h + i) * p.w + jj;
        float4 v;
        if (f < 4) {
            v = ldg4(p.cur + tpix * p.ldc + f * 4);
        } else if (f < 8) {
            v = *reinterpret_cast<const float4*>(&s_dec[0][t][(f - 4) * 4]);
        } else if (f < 12) {
            const int hp = p.h >> 1, wp = p.w >> 1;
            v = ldg4(p.prev + (((size_t)n * hp + (i >> 1)) * wp + (jj >> 1)) * p.ldp + (f - 8) * 4);
            if (f == 8) {
                const float cx = (float)(jj & 1) - 0.5f, cy = (float)(i & 1) - 0.5f;
                v.x = __fmul_rn(__fadd_rn(__fadd_rn(v.x, __fmul_rn(cx, v.y)), __fmul_rn(cy, v.z)), 2.f);
            }
        } else {
            v = *reinterpret_cast<const float4*>(&s_dec[NSETS - 1][t][(f - 12) * 4]);
        }
        *reinterpret_cast<float4*>(p.aug + tpix * p.ldaug + f * 4) = v;
    }
    if (p.raw) {
        for (int o = tid; o < K4_TILES * NSETS * 16; o += K4_THREADS) {
            const int t = o / (NSETS * 16), f = o - t * (NSETS * 16);
            const int jj = j0 + t;
            if (jj >= p.w) continue;
            const size_t tpix = ((size_t)n * p.h + i) * p.w + jj;
            const int s = f >> 4, e = f & 15;
            *reinterpret_cast<float4*>(p.raw + tpix * (NSETS * 64) + f * 4) =
                *reinterpret_cast<const float4*>(&s_raw[s][t][e * 4]);
        }
    }
}

// ------------------------------------------------------------------------------------------
// K5 hyp_select
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) hyp_select_kernel(const float* __restrict__ upd, int ldu,
                                                         const float* __restrict__ aug, int ldaug, size_t npix,
                                                         float* __restrict__ out, int ldr) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= npix * 4) return;
    const int c4 = (int)(idx & 3);
    const size_t pix = idx >> 2;
    const float* u = upd + pix * ldu;
    const float conf_prev = __ldg(u), conf_cur = __ldg(u + 1);
    const bool take_cur = conf_cur > conf_prev;  // arg-max, first index (previous) on ties
    const float* hyp = aug + pix * ldaug + (take_cur ? 0 : 32) + c4 * 4;
    const float* del = u + (take_cur ? 18 : 2) + c4 * 4;
    const float4 hv = ldg4(hyp);
    float4 r;
    r.x = __fadd_rn(hv.x, __ldg(del));
    r.y = __fadd_rn(hv.y, __ldg(del + 1));
    r.z = __fadd_rn(hv.z, __ldg(del + 2));
    r.w = __fadd_rn(hv.w, __ldg(del + 3));
    if (c4 == 0) r.x = fmaxf(r.x, 0.f);
    *reinterpret_cast<float4*>(out + pix * ldr + c4 * 4) = r;
}

}  // namespace

extern "C" int codd_tile_hyp_init(const float* min_cost, const float* min_disp, const float* feat, int ldf, int cf,
                                  const float* weight, const float* bias, int n, int h, int w, float* hyp, int ldh,
                                  void* stream) {
    if (!min_cost || !min_disp || !feat || !weight || !bias || !hyp || n <= 0 || h <= 0 || w <= 0) return CODD_E_BADARG;
    if (cf <= 0 || cf % 4 != 0 || cf > 64 || ldh < 16 || ldh % 4 != 0) return CODD_E_SHAPE;
    if (ldf != 0 && (ldf < cf || ldf % 4 != 0)) return CODD_E_SHAPE;
    if ((ldf != 0 && !codd_aligned16(feat)) || !codd_aligned16(hyp)) return CODD_E_ALIGN;
    const size_t npix = (size_t)n * h * w;
    const size_t smem = (size_t)(1 + cf) * 16 * sizeof(float);
    tile_hyp_init_kernel<<<(unsigned)((npix + 127) / 128), 128, smem, (cudaStream_t)stream>>>(
        min_cost, min_disp, feat, ldf, cf, weight, bias, npix, (size_t)h * w, hyp, ldh);
    CODD_RETURN_IF_CUDA_ERROR();
    return 0;
}

extern "C" int codd_plane_upsample(const float* in, int ldi, int n, int h, int w, int size, float scale, float* out,
                                   int ldo, void* stream) {
    if (!in || !out || n <= 0 || h <= 0 || w <= 0 || size <= 0) return CODD_E_BADARG;
    if (ldi < 16 || ldo < 16 || ldi % 4 != 0 || ldo % 4 != 0) return CODD_E_SHAPE;
    if (!codd_aligned16(in) || !codd_aligned16(out)) return CODD_E_ALIGN;
    const size_t total = (size_t)n * h * size * w * size * 4;
    plane_upsample_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, ldi, n, h, w, size,
                                                                                              scale, out, ldo);
    CODD_RETURN_IF_CUDA_ERROR();
    return 0;
}

namespace {
int tile_warp_cost_impl(const float* fea_l, int ldfl, const float* fea_r, int ldfr, int c, const float* cur, int ldc,
                        const float* prev, int ldp, const float* dec_w, const float* dec_b, int n, int h, int w, float* aug,
                        int ldaug, float* raw_cv, void* stream) {
    if (!fea_l || !fea_r || !cur || !dec_w || !dec_b || !aug || n <= 0 || h <= 0 || w <= 0) return CODD_E_BADARG;
    if (c <= 0 || c % 4 != 0 || ldfl < c || ldfl % 4 || ldc < 16 || ldc % 4) return CODD_E_SHAPE;
    if (prev && (ldp < 16 || ldp % 4 || (h & 1) || (w & 1))) return CODD_E_SHAPE;
    if (ldaug < (prev ? 64 : 32) || ldaug % 4) return CODD_E_SHAPE;
    if (!codd_aligned16(fea_l) || !codd_aligned16(cur) || !codd_aligned16(aug) ||
        (prev && !codd_aligned16(prev)) || (raw_cv && !codd_aligned16(raw_cv)))
        return CODD_E_ALIGN;
    WarpP p;
    if (ldfr != 0 && (ldfr < c || ldfr % 4 || !codd_aligned16(fea_r))) return CODD_E_SHAPE;
    p.fl = fea_l; p.ldfl = ldfl; p.C = c;
    p.fr = ldfr == 0 ? fea_r : nullptr;
    p.fr_nhwc = ldfr != 0 ? fea_r : nullptr;
    p.ldfr = ldfr;
    p.cur = cur; p.ldc = ldc; p.prev = prev; p.ldp = ldp;
    p.dec_w = dec_w; p.dec_b = dec_b; p.N = n; p.h = h; p.w = w;
    p.aug = aug; p.ldaug = ldaug; p.raw = raw_cv;
#ifdef CODD_DIAG
    static const int max_win = getenv("CODD_K4_MAXWIN") ? atoi(getenv("CODD_K4_MAXWIN")) : 4096;
    p.max_win = max_win;
#else
    p.max_win = 4096;
#endif
    const int nblk = codd_ceil_div(w, K4_TILES);
    dim3 grid((unsigned)(n * h * nblk)), block(K4_THREADS);
    static CoddDeviceOnce once;   // once per device, so graph capture sees no attribute calls
    if (int rc = codd_once_per_device(once, [&] {
            cudaError_t e = cudaFuncSetAttribute(tile_warp_cost_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, K4_STAGE_BYTES);
            if (e == cudaSuccess)
                e = cudaFuncSetAttribute(tile_warp_cost_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, K4_STAGE_BYTES);
            return e;
        }))
        return rc;
    if (prev) tile_warp_cost_kernel<2><<<grid, block, K4_STAGE_BYTES, (cudaStream_t)stream>>>(p);
    else tile_warp_cost_kernel<1><<<grid, block, K4_STAGE_BYTES, (cudaStream_t)stream>>>(p);
    CODD_RETURN_IF_CUDA_ERROR();
    return 0;
}
}  // namespace

extern "C" int codd_tile_warp_cost(const float* fea_l, int ldfl, const float* fea_r_planar, int c,
                                   const float* cur, int ldc, const float* prev, int ldp, const float* dec_w,
                                   const float* dec_b, int n, int h, int w, float* aug, int ldaug, float* raw_cv,
                                   void* stream) {
    return tile_warp_cost_impl(fea_l, ldfl, fea_r_planar, 0, c, cur, ldc, prev, ldp, dec_w, dec_b, n, h, w, aug, ldaug, raw_cv,
                               stream);
}

// second decomposition (tile_warp.cu): TMA-staged window, planes sharing a register window, tensor-core `decrease`
int codd_tile_warp_cost2(const float* fea_l, int ldfl, const float* fea_r, int ldfr, int c, const float* cur, int ldc,
                         const float* prev, int ldp, const float* dec_w, const float* dec_b, int n, int h, int w,
                         float* aug, int ldaug, float* raw_cv, void* stream);

extern "C" int codd_tile_warp_cost_nhwc(const float* fea_l, int ldfl, const float* fea_r, int ldfr, int c,
                                        const float* cur, int ldc, const float* prev, int ldp, const float* dec_w,
                                        const float* dec_b, int n, int h, int w, float* aug, int ldaug, float* raw_cv,
                                        void* stream) {
    if (ldfr <= 0) return CODD_E_BADARG;
    bool v2 = true;
#ifdef CODD_DIAG
    static const bool force_v1 = getenv("CODD_K4_V1") && atoi(getenv("CODD_K4_V1")) != 0;   // A/B against the first kernel
    v2 = !force_v1;
#endif
    if (v2) {
        // same argument checks as the first kernel, then the TMA / tensor-core kernel for C in {16, 24, 32}
        if (!fea_l || !fea_r || !cur || !dec_w || !dec_b || !aug || n <= 0 || h <= 0 || w <= 0) return CODD_E_BADARG;
        if (c <= 0 || c % 4 != 0 || ldfl < c || ldfl % 4 || ldc < 16 || ldc % 4 || ldfr < c || ldfr % 4) return CODD_E_SHAPE;
        if (prev && (ldp < 16 || ldp % 4 || (h & 1) || (w & 1))) return CODD_E_SHAPE;
        if (ldaug < (prev ? 64 : 32) || ldaug % 4) return CODD_E_SHAPE;
        if (!codd_aligned16(fea_l) || !codd_aligned16(fea_r) || !codd_aligned16(cur) || !codd_aligned16(aug) ||
            !codd_aligned16(dec_b) || !codd_aligned16(dec_w) || (prev && !codd_aligned16(prev)))
            return CODD_E_ALIGN;
        const int rc = codd_tile_warp_cost2(fea_l, ldfl, fea_r, ldfr, c, cur, ldc, prev, ldp, dec_w, dec_b, n, h, w, aug,
                                            ldaug, raw_cv, stream);
        if (rc != CODD_E_UNSUPPORTED) return rc;
    }
    return tile_warp_cost_impl(fea_l, ldfl, fea_r, ldfr, c, cur, ldc, prev, ldp, dec_w, dec_b, n, h, w, aug, ldaug, raw_cv,
                               stream);
}

extern "C" int codd_hyp_select(const float* update, int ldu, const float* aug, int ldaug, int n, int h, int w,
                               float* refined, int ldr, void* stream) {
    if (!update || !aug || !refined || n <= 0 || h <= 0 || w <= 0) return CODD_E_BADARG;
    if (ldu < 34 || ldaug < 64 || ldaug % 4 || ldr < 16 || ldr % 4) return CODD_E_SHAPE;
    if (!codd_aligned16(aug) || !codd_aligned16(refined)) return CODD_E_ALIGN;
    const size_t npix = (size_t)n * h * w;
    hyp_select_kernel<<<(unsigned)((npix * 4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(update, ldu, aug, ldaug,
                                                                                           npix, refined, ldr);
    CODD_RETURN_IF_CUDA_ERROR();
    return 0;
}
