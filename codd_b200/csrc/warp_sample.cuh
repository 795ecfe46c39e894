// Disparity-warp sampling shared by K4 (tile_warp_cost) and the Fusion cue kernels: the exact
// coordinate / blend arithmetic of the reference's warp + torch.grid_sample (see tile.cu header).
#pragma once
#include "common.cuh"

namespace {

struct Taps {
    // sampling state of one hypothesis plane at one pixel (k = -1, 0, +1)
    int x0[3];
    float fw[3], fe[3];
};

// a / b rounded to nearest for a fixed, normal divisor b with r = RN(1/b) (Markstein): q0 = RN(a*r),
// rem = a - q0*b exactly (one FMA), q = RN(q0 + rem*r).  Operands here are pixel coordinates
// (|a| < 2^20, 1 <= b < 2^20): no overflow / denormal cases, so the two-FMA correction yields the
// correctly rounded quotient; tests/test_gpu_ops.py compares it bit-for-bit with the oracle.
__device__ __forceinline__ float div_rn_const(float a, float b, float r) {
    const float q0 = __fmul_rn(a, r);
    const float rem = __fmaf_rn(-q0, b, a);
    return __fmaf_rn(rem, r, q0);
}

__device__ __forceinline__ void sample_setup(float d, float dx, float dy, float a, float b, int x, float wm1,
                                             float wdiv, float wrcp, Taps& t) {
#pragma unroll
    for (int ki = 0; ki < 3; ++ki) {
        const float k = (float)(ki - 1);
        // Eq.(5): ((d + k) + a*dx) + b*dy
        const float ld = __fadd_rn(__fadd_rn(__fadd_rn(d, k), __fmul_rn(a, dx)), __fmul_rn(b, dy));
        // reference warp(): 2*(x - d)/max(W-1,1) - 1 ; grid_sample: ((g+1)/2)*(W-1)
        const float g = __fadd_rn(div_rn_const(__fmul_rn(2.f, __fsub_rn((float)x, ld)), wdiv, wrcp), -1.f);
        const float ix = __fmul_rn(__fmul_rn(__fadd_rn(g, 1.f), 0.5f), wm1);
        const float fx = floorf(ix);
        t.fw[ki] = __fsub_rn(ix, fx);
        t.fe[ki] = __fsub_rn(1.f, t.fw[ki]);
        // clamp before the int conversion: anything outside [-2, W] samples only zeros anyway
        t.x0[ki] = (int)fminf(fmaxf(fx, -2.f), wm1 + 1.f);
    }
}

// Channel loops of K4.  Every (set, k) plane gathers its own two columns x0, x0+1: the planes of a
// set are nominally one pixel apart, but for integer disparities (the arg-min initialisation) ix
// sits within an ulp of an integer and floor() jitters by one independently per plane, so no
// shared window is assumed.  Out-of-range taps get a zero WEIGHT instead of a zero value —
// identical cost: torch's blend then only adds (+-0) for them.
//
// PAIRED: both columns of every plane are inside the row (0 <= x0 < W-1), so the second tap is the
// first one's neighbour (one address computation per pair).  Otherwise (a plane touching the
// image border) the columns are clamped individually.
template <int NSETS, bool TWO_ROWS, bool PAIRED>
__device__ __forceinline__ void k4_channels(const float* __restrict__ flp, const float* __restrict__ frn,
                                            size_t cstride, int C, int rowstep, const int (&offA)[NSETS][3],
                                            const int (&offB)[NSETS][3], const float (&wA)[NSETS][3],
                                            const float (&wB)[NSETS][3], const float (&wC)[NSETS][3],
                                            const float (&wD)[NSETS][3], float (&cost)[NSETS][3], float& lnorm) {
    for (int c = 0; c < C; c += 4) {
        const float4 l4 = ldg4(flp + c);
        const float lv[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            const float l = lv[cc];
            lnorm = __fadd_rn(lnorm, fabsf(l));
            const float* base = frn + (size_t)(c + cc) * cstride;
            float tA[NSETS][3], tB[NSETS][3], uA[NSETS][3], uB[NSETS][3];
#pragma unroll
            for (int s = 0; s < NSETS; ++s)
#pragma unroll
                for (int ki = 0; ki < 3; ++ki) {
                    const float* pa = base + offA[s][ki];
                    const float* pb = PAIRED ? pa + 1 : base + offB[s][ki];
                    tA[s][ki] = __ldg(pa);
                    tB[s][ki] = __ldg(pb);
                    if (TWO_ROWS) {
                        uA[s][ki] = __ldg(pa + rowstep);
                        uB[s][ki] = __ldg(pb + rowstep);
                    }
                }
#pragma unroll
            for (int s = 0; s < NSETS; ++s)
#pragma unroll
                for (int ki = 0; ki < 3; ++ki) {
                    // torch's bilinear: fma(se_v, se, fma(sw_v, sw, fma(ne_v, ne, nw_v*nw)))
                    float v = __fmaf_rn(tB[s][ki], wB[s][ki], __fmul_rn(tA[s][ki], wA[s][ki]));
                    if (TWO_ROWS) v = __fmaf_rn(uB[s][ki], wD[s][ki], __fmaf_rn(uA[s][ki], wC[s][ki], v));
                    cost[s][ki] = __fadd_rn(cost[s][ki], fabsf(__fsub_rn(l, v)));
                }
        }
    }
}


}  // namespace
