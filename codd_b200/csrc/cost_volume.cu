// K1 — fused L1 cost volume + arg-min tile initialisation.
//
// Replaces calc_init_disp (model/stereo/hitnet/initialization.py:18-45: repeat a [N,D,h,w,3]
// sampling grid, 5-D nearest grid_sample into an [N,16,D,h,w] tensor, subtract, L1-norm over
// channels) and the torch.min over D that follows it (initialization.py:167-171).
//
//   cv[n,d,i,j] = sum_{c=0..15} | L[n,i,j,c] - R[n,i,4j-d,c] |      (R := 0 for 4j-d < 0)
//
// The channel sum is sequential in fp32 (two FADD per term, no FMA) so every cost is
// bit-identical to the reference's CPU result, and ties in the minimum resolve to the first d
// as torch.min does — the reference has 2.5-15 % exact ties (all zero-filled shifts give
// |L|_1), so both properties are needed for bit-exact arg-min indices.
//
// Work decomposition.  Write d = 4q + r.  Then 4j - d = 4(j-q) - r: the four right-feature
// columns a (j,q) pair needs are one aligned quad that depends only on m = j - q.
//   * one CTA per (sample n, tile row i, block of tile columns);
//   * the right row R[n,i,:,:] and left row L[n,i,:,:] are staged once in shared memory,
//     transposed to channel-planar so that lanes touching consecutive columns are conflict free;
//   * lane <-> one value of m.  It keeps the 16-channel x 4-column quad of R for its m in 64
//     registers for the whole kernel and walks q = 0..D/4-1, i.e. j = m + q.  At a given step all
//     lanes of a warp share q (same four disparities) and hold consecutive j, so
//       - L is read from shared memory with unit stride (16 LDS.32 per step),
//       - the four cost rows cv[n,4q+r,i,j..j+31] are written as full 128-byte lines,
//       - every R value is read from shared memory exactly once per CTA.
//   * arithmetic runs as packed FADD2 (two disparities per instruction, |.| folded into the
//     accumulate), which is what bounds the fused arg-min variant; the materialising variant is
//     bound by the cv write (4*D bytes per tile).
//   * arg-min: the running (min, argmin) of column j travels one lane down per step together
//     with j (j is served by lane j-m0-q of the warp at step q, so d ascends as the lane index
//     descends); per-warp partial results are merged in shared memory in ascending-d order with
//     a strict '<', preserving torch's first-index tie rule.
//   * columns with 4j-d < -3 (m < 0) need no work: their cost is |L|_1, which is also the cost
//     of d = 4j+1 (computed in the main loop), so they can never win the strict arg-min; the
//     materialising variant fills them from a per-column |L|_1 table.
#include "common.cuh"

namespace {

constexpr int CV_C = 16;  // tile-feature channels (TileInitialization always emits 16)

struct CvP {
    const float* L;
    const float* R;
    int ldl, ldr;
    int N, h, w, D;
    int JB;        // tile columns per CTA
    int nblk;      // column blocks per row
    float* cv;
    float* min_cost;
    float* min_disp;
};

template <bool WRITE_CV, bool ARGMIN>
__global__ void __launch_bounds__(384) cost_volume_kernel(CvP p) {
    extern __shared__ float4 smem4[];
    const int Dq = (p.D + 3) >> 2;   // disparity quads; the last one may be partial
    int b = blockIdx.x;
    const int jblk = b % p.nblk;
    b /= p.nblk;
    const int i = b % p.h;
    const int n = b / p.h;
    const int jb = jblk * p.JB;
    const int je = min(jb + p.JB, p.w);          // columns [jb, je)
    const int mlo = max(jb - Dq + 1, 0);         // m range [mlo, je)
    const int nm = je - mlo;
    const int nj = je - jb;
    const int W = 4 * p.w;

    // shared layout: S[c][RW] (right row, planar, index x - 4*mlo + 3), Lp[c][LW], l1[LW],
    // partial cost/disp [nwarps][LW]
    const int RW = 4 * nm + 4;
    const int LW = (nj + 3) & ~3;
    float* S = reinterpret_cast<float*>(smem4);
    float* Lp = S + CV_C * RW;
    float* l1 = Lp + CV_C * LW;
    float* pc = l1 + LW;
    const int nwarps = blockDim.x >> 5;
    float* pd = pc + nwarps * LW;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t rowL = ((size_t)n * p.h + i) * p.w;
    const size_t rowR = ((size_t)n * p.h + i) * W;

    // ---- stage R: columns x in [4*mlo - 3, 4*(je-1)]  -> S[c][x - 4*mlo + 3]
    for (int idx = tid; idx < RW * 4; idx += blockDim.x) {
        const int c4 = idx & 3;
        const int sx = idx >> 2;
        const int x = sx + 4 * mlo - 3;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (x >= 0 && x < W) v = ldg4(p.R + (rowR + x) * p.ldr + c4 * 4);
        S[(c4 * 4 + 0) * RW + sx] = v.x;
        S[(c4 * 4 + 1) * RW + sx] = v.y;
        S[(c4 * 4 + 2) * RW + sx] = v.z;
        S[(c4 * 4 + 3) * RW + sx] = v.w;
    }
    // ---- stage L
    for (int idx = tid; idx < LW * 4; idx += blockDim.x) {
        const int c4 = idx & 3;
        const int jj = idx >> 2;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (jj < nj) v = ldg4(p.L + (rowL + jb + jj) * p.ldl + c4 * 4);
        Lp[(c4 * 4 + 0) * LW + jj] = v.x;
        Lp[(c4 * 4 + 1) * LW + jj] = v.y;
        Lp[(c4 * 4 + 2) * LW + jj] = v.z;
        Lp[(c4 * 4 + 3) * LW + jj] = v.w;
    }
    if (ARGMIN)
        for (int idx = tid; idx < nwarps * LW; idx += blockDim.x) {
            pc[idx] = INFINITY;
            pd[idx] = 0.f;
        }
    __syncthreads();

    if (WRITE_CV) {
        // |L|_1 per column, channel-sequential: the cost of every zero-filled shift
        for (int jj = tid; jj < nj; jj += blockDim.x) {
            float a = fabsf(Lp[jj]);
#pragma unroll
            for (int c = 1; c < CV_C; ++c) a = __fadd_rn(a, fabsf(Lp[c * LW + jj]));
            l1[jj] = a;
        }
    }

    // ---- main loop: lane <-> m
    const int m0 = mlo + warp * 32;         // first m of this warp
    const int m = m0 + lane;
    const bool lane_on = (m < je);
    float2 rq[CV_C][2];                     // R quad per channel: (x=4m-3, 4m-2), (4m-1, 4m)
    if (lane_on) {
#pragma unroll
        for (int c = 0; c < CV_C; ++c) {
            const float4 v = *reinterpret_cast<const float4*>(S + c * RW + 4 * (m - mlo));
            rq[c][0] = make_float2(v.x, v.y);
            rq[c][1] = make_float2(v.z, v.w);
        }
    } else {
#pragma unroll
        for (int c = 0; c < CV_C; ++c) rq[c][0] = rq[c][1] = make_float2(0.f, 0.f);
    }

    float bc = INFINITY;  // travelling best (cost, disp) of the column currently at this lane
    float bd = 0.f;
    const size_t plane = (size_t)p.h * p.w;
    float* cvrow = WRITE_CV ? p.cv + (size_t)n * p.D * plane + (size_t)i * p.w : nullptr;

    if (m0 < je) {  // warp-uniform
        for (int q = 0; q < Dq; ++q) {
            const int j = m + q;
            const bool on = lane_on && j >= jb && j < je;
            if (on) {
                const float* lp = Lp + (j - jb);
                // quad element e <-> x = 4m-3+e <-> r = 3-e:  hi = (r=3, r=2), lo = (r=1, r=0)
                float2 hi = make_float2(0.f, 0.f), lo = make_float2(0.f, 0.f);
#pragma unroll
                for (int c = 0; c < CV_C; ++c) {
                    const float l = lp[c * LW];
                    const float2 ll = make_float2(l, l);
                    float2 d0 = __fadd2_rn(ll, make_float2(-rq[c][0].x, -rq[c][0].y));
                    float2 d1 = __fadd2_rn(ll, make_float2(-rq[c][1].x, -rq[c][1].y));
                    hi = __fadd2_rn(hi, make_float2(fabsf(d0.x), fabsf(d0.y)));
                    lo = __fadd2_rn(lo, make_float2(fabsf(d1.x), fabsf(d1.y)));
                }
                const float c0 = lo.y, c1 = lo.x, c2 = hi.y, c3 = hi.x;  // r = 0,1,2,3
                const int nd = p.D - 4 * q;   // valid disparities in this quad (>= 1; < 4 only in the last)
                if (WRITE_CV) {
                    float* o = cvrow + (size_t)(4 * q) * plane + j;
                    __stcs(o, c0);
                    if (nd > 1) __stcs(o + plane, c1);
                    if (nd > 2) __stcs(o + 2 * plane, c2);
                    if (nd > 3) __stcs(o + 3 * plane, c3);
                }
                if (ARGMIN) {
                    const float dq = (float)(4 * q);
                    if (c0 < bc) { bc = c0; bd = dq; }
                    if (nd > 1 && c1 < bc) { bc = c1; bd = dq + 1.f; }
                    if (nd > 2 && c2 < bc) { bc = c2; bd = dq + 2.f; }
                    if (nd > 3 && c3 < bc) { bc = c3; bd = dq + 3.f; }
                }
            }
            if (ARGMIN) {
                // column j leaves the warp through lane 0; everything else moves one lane down
                if (lane == 0 && on) {
                    pc[warp * LW + (j - jb)] = bc;
                    pd[warp * LW + (j - jb)] = bd;
                }
                bc = __shfl_down_sync(0xffffffffu, bc, 1);
                bd = __shfl_down_sync(0xffffffffu, bd, 1);
                if (lane == 31) { bc = INFINITY; bd = 0.f; }
            }
        }
        if (ARGMIN) {
            // columns still in flight: lane now holds the state of column m + Dq
            const int j = m + Dq;
            if (lane < 31 && j >= jb && j < je) {
                pc[warp * LW + (j - jb)] = bc;
                pd[warp * LW + (j - jb)] = bd;
            }
        }
    }
    __syncthreads();

    if (ARGMIN) {
        // merge per-warp partials: higher warp (larger m) <-> smaller d, so walk warps downwards
        for (int jj = tid; jj < nj; jj += blockDim.x) {
            float c = INFINITY, d = 0.f;
            for (int wv = nwarps - 1; wv >= 0; --wv) {
                const float cc = pc[wv * LW + jj];
                if (cc < c) { c = cc; d = pd[wv * LW + jj]; }
            }
            const size_t o = rowL + jb + jj;
            if (p.min_cost) p.min_cost[o] = c;
            if (p.min_disp) p.min_disp[o] = d;
        }
    }
    if (WRITE_CV) {
        // zero-filled region: 4j+4 <= d < D  (only columns j < (D-1)/4 have one)
        const int jz = min(nj, max(0, (p.D - 1) / 4 - jb));   // local columns [0, jz)
        if (jz > 0) {
            for (int idx = tid; idx < p.D * jz; idx += blockDim.x) {
                const int d = idx / jz, jj = idx - d * jz;
                if (d >= 4 * (jb + jj) + 4) __stcs(cvrow + (size_t)d * plane + jb + jj, l1[jj]);
            }
        }
    }
}

}  // namespace

extern "C" int codd_cost_volume(const float* tile_l, int ldl, const float* tile_r, int ldr, int n, int h, int w,
                                int max_disp, float* cv, float* min_cost, float* min_disp, void* stream) {
    if (!tile_l || !tile_r || n <= 0 || h <= 0 || w <= 0 || max_disp <= 0) return CODD_E_BADARG;
    if (ldl < CV_C || ldr < CV_C || ldl % 4 != 0 || ldr % 4 != 0) return CODD_E_SHAPE;
    if (!codd_aligned16(tile_l) || !codd_aligned16(tile_r)) return CODD_E_ALIGN;
    const bool argmin = (min_cost != nullptr) || (min_disp != nullptr);
    if (!cv && !argmin) return CODD_E_BADARG;
    const int Dq = (max_disp + 3) / 4;
    // columns per CTA: the m-range (JB + Dq - 1 values, one lane each) must fit 12 warps
    const int max_m = 12 * 32;
    int nblk = 1;
    while (codd_ceil_div(w, nblk) + Dq - 1 > max_m) ++nblk;
    if (Dq - 1 >= max_m) return CODD_E_SHAPE;
    const int JB = codd_ceil_div(w, nblk);
    nblk = codd_ceil_div(w, JB);
    const int nm_max = JB + ((nblk > 1) ? (Dq - 1) : 0);
    const int nm = nblk > 1 ? nm_max : w;   // single block: m in [0, w)
    const int nwarps = codd_ceil_div(nm, 32);
    const int RW = 4 * nm + 4;
    const int LW = (JB + 3) & ~3;
    const size_t smem = (size_t)(CV_C * RW + CV_C * LW + LW + 2 * nwarps * LW) * sizeof(float);
    if (smem > 227 * 1024) return CODD_E_SHAPE;

    CvP p;
    p.L = tile_l; p.R = tile_r; p.ldl = ldl; p.ldr = ldr;
    p.N = n; p.h = h; p.w = w; p.D = max_disp; p.JB = JB; p.nblk = nblk;
    p.cv = cv; p.min_cost = min_cost; p.min_disp = min_disp;
    dim3 grid((unsigned)(n * h * nblk)), block(32 * nwarps);
    cudaStream_t s = (cudaStream_t)stream;
    void (*kern)(CvP) = nullptr;
    static size_t configured[3] = {48 * 1024, 48 * 1024, 48 * 1024};
    int slot;
    if (cv && argmin) { kern = cost_volume_kernel<true, true>; slot = 0; }
    else if (cv) { kern = cost_volume_kernel<true, false>; slot = 1; }
    else { kern = cost_volume_kernel<false, true>; slot = 2; }
    if (smem > configured[slot]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        configured[slot] = smem;
    }
    kern<<<grid, block, smem, s>>>(p);
    CODD_RETURN_IF_CUDA_ERROR();
    return 0;
}
