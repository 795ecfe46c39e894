// K1 — fused L1 cost volume + arg-min tile initialisation.
//
// Replaces calc_init_disp (model/stereo/hitnet/initialization.py:18-45: repeat a [N,D,h,w,3]
// sampling grid, 5-D nearest grid_sample into an [N,16,D,h,w] tensor, subtract, L1-norm over
// channels) and the torch.min over D that follows it (initialization.py:167-171).
//
//   cv[n,d,i,j] = sum_{c=0..15} | L[n,c,i,j] - R[n,c,i,4j-d] |      (R := 0 for 4j-d < 0)
//
// The channel sum is sequential in fp32 (two FADD per term, no FMA) so every cost is
// bit-identical to the reference's CPU result, and ties in the minimum resolve to the first d
// as torch.min does — the reference has 2.5-15 % exact ties (all zero-filled shifts give
// |L|_1), so both properties are needed for bit-exact arg-min indices.
//
// Inputs are PLANAR tile features ([N,16,h,w] and [N,16,h,4w], what K2 emits), so a CTA's
// right-feature row is 16 contiguous global rows.
//
// Work decomposition.  Write d = 4q - r, r in 0..3.  Then 4j - d = 4(j-q) + r: the four
// right-feature columns a (j,q) pair needs are the ALIGNED quad R[4m .. 4m+3] with m = j - q.
//   * one CTA per (sample n, tile row i, block of tile columns);
//   * the 16 channel rows of R[n,:,i,4*mlo ..] are staged in shared memory by bulk-TMA copies
//     (cp.async.bulk, one per channel, completion on an mbarrier) — no thread touches them;
//   * lane <-> one value of m.  It keeps the 16-channel x 4-column quad of R for its m in 64
//     registers for the whole kernel and walks q = 0..ceil(D/4), i.e. j = m + q.  At a given step
//     all lanes of a warp share q (same four disparities) and hold consecutive j, so
//       - L is read with unit stride (16 coalesced 128-byte loads per step, L1 resident),
//       - the four cost rows cv[n,4q-r,i,j..j+31] are written as full 128-byte lines,
//       - every R value is read from shared memory exactly once per CTA.
//   * arithmetic runs as packed FADD2 (two disparities per instruction, |.| folded into the
//     accumulate), which is what bounds the fused arg-min variant; the materialising variant is
//     bound by the cv write (4*D bytes per tile).
//   * arg-min: the running (min, argmin) of column j travels one lane down per step together
//     with j (j is served by lane j-m0-q of the warp at step q, so d ascends as the lane index
//     descends); per-warp partial results are merged in shared memory in ascending-d order with
//     a strict '<', preserving torch's first-index tie rule.
//   * shifts that fall off the left edge (m < 0, i.e. d >= 4j+1) need no work: their cost is
//     |L|_1 for every d, computed once per column; it enters the arg-min as the single candidate
//     d = 4j+1 (merged last) and the materialising variant fills it in from a per-column table.
#include <type_traits>

#include "common.cuh"

namespace {

constexpr int CV_C = 16;        // tile-feature channels (TileInitialization always emits 16)
constexpr int CV_MAXW = 8;      // warps per CTA (one lane per m)
constexpr int CV_LW = CV_MAXW * 32 + 8;   // shared row stride of the staged left features (compile-time:
                                          // the 16 per-step channel reads become immediate offsets)

struct CvP {
    const float* L;   // [N,16,h,w]
    const float* R;   // [N,16,h,4w]
    int N, h, w, D;
    int JB;        // tile columns per CTA
    int nblk;      // column blocks per row
    float* cv;
    float* min_cost;
    float* min_disp;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <bool WRITE_CV, bool ARGMIN>
__device__ __forceinline__ void cost_volume_body(const CvP& p, int b) {
    extern __shared__ float4 smem4[];
    __shared__ __align__(8) unsigned long long mbar;
    const int qmax = (p.D + 2) >> 2;             // q in [0, qmax];  d = 4q - r
    const int jblk = b % p.nblk;
    b /= p.nblk;
    const int i = b % p.h;
    const int n = b / p.h;
    const int jb = jblk * p.JB;
    const int je = min(jb + p.JB, p.w);          // columns [jb, je)
    const int mlo = max(jb - qmax, 0);           // m range [mlo, je)
    const int nm = je - mlo;
    const int nj = je - jb;
    const int W = 4 * p.w;
    const size_t plane = (size_t)p.h * p.w;

    // shared layout: S[c][RW] (right row, x - 4*mlo), Lp[c][CV_LW], l1[LW], partial cost/disp [nwarps][LW]
    const int RW = 4 * nm + 4;                   // +4: keeps rows 16-byte aligned and de-phased
    const int LW = (nj + 3) & ~3;
    float* S = reinterpret_cast<float*>(smem4);
    float* Lp = S + CV_C * RW;
    float* l1 = Lp + CV_C * CV_LW;
    float* pc = l1 + LW;
    const int nwarps = blockDim.x >> 5;
    float* pd = pc + nwarps * LW;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* Lrow = p.L + ((size_t)n * CV_C * p.h + i) * p.w;             // + c*plane + j
    const float* Rrow = p.R + ((size_t)n * CV_C * p.h + i) * W + 4 * mlo;     // + c*h*W + x

    // ---- stage R with bulk-TMA copies: 16 rows of 16*nm bytes each
    const uint32_t bar = smem_u32(&mbar);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        const uint32_t bytes = (uint32_t)(16 * nm);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes * CV_C) : "memory");
        for (int c = 0; c < CV_C; ++c) {
            const float* src = Rrow + (size_t)c * p.h * W;
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                    smem_u32(S + c * RW)),
                "l"(src), "r"(bytes), "r"(bar)
                : "memory");
        }
    }
    // meanwhile: stage L (planar -> planar, coalesced), per-column |L|_1 (channel-sequential)
    for (int jj = tid; jj < nj; jj += blockDim.x) {
        float lv[CV_C];
#pragma unroll
        for (int c = 0; c < CV_C; ++c) lv[c] = __ldg(Lrow + (size_t)c * plane + jb + jj);
        float a = fabsf(lv[0]);
        Lp[jj] = lv[0];
#pragma unroll
        for (int c = 1; c < CV_C; ++c) {
            a = __fadd_rn(a, fabsf(lv[c]));
            Lp[c * CV_LW + jj] = lv[c];
        }
        l1[jj] = a;
    }
    if (ARGMIN)
        for (int idx = tid; idx < nwarps * LW; idx += blockDim.x) {
            pc[idx] = INFINITY;
            pd[idx] = 0.f;
        }
    // wait for the bulk copies (phase 0)
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
    __syncthreads();

    // ---- main loop: lane <-> m
    const int m0 = mlo + warp * 32;         // first m of this warp
    const int m = m0 + lane;
    const bool lane_on = (m < je);
    float2 rq[CV_C][2];                     // R quad per channel: (x=4m, 4m+1), (4m+2, 4m+3)
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
    float* cvrow = WRITE_CV ? p.cv + (size_t)n * p.D * plane + (size_t)i * p.w : nullptr;

    // one step of the walk; FULL: all four disparities 4q-3..4q lie in [0, D)
    auto step = [&](const int q, auto full_tag) {
        constexpr bool FULL = decltype(full_tag)::value;
        const int j = m + q;
        const bool on = lane_on && j >= jb && j < je;
        if (on) {
            const float* lp = Lp + (j - jb);
            // lo = (r=0, r=1) <-> d = (4q, 4q-1);  hi = (r=2, r=3) <-> d = (4q-2, 4q-3)
            float2 lo = make_float2(0.f, 0.f), hi = make_float2(0.f, 0.f);
#pragma unroll
            for (int c = 0; c < CV_C; ++c) {
                const float l = lp[c * CV_LW];
                const float2 ll = make_float2(l, l);
                float2 d0 = __fadd2_rn(ll, make_float2(-rq[c][0].x, -rq[c][0].y));
                float2 d1 = __fadd2_rn(ll, make_float2(-rq[c][1].x, -rq[c][1].y));
                lo = __fadd2_rn(lo, make_float2(fabsf(d0.x), fabsf(d0.y)));
                hi = __fadd2_rn(hi, make_float2(fabsf(d1.x), fabsf(d1.y)));
            }
            const int d0i = 4 * q;   // disparity of r = 0; the r-th cost belongs to d0i - r
            const float c0 = lo.x, c1 = lo.y, c2 = hi.x, c3 = hi.y;
            const bool v0 = FULL || d0i < p.D;
            const bool v1 = FULL || ((d0i >= 1) && (d0i - 1 < p.D));
            const bool v2 = FULL || ((d0i >= 2) && (d0i - 2 < p.D));
            const bool v3 = FULL || ((d0i >= 3) && (d0i - 3 < p.D));
            if (WRITE_CV) {
                float* o = cvrow + (ptrdiff_t)d0i * (ptrdiff_t)plane + j;
                if (v3) __stcs(o - 3 * plane, c3);
                if (v2) __stcs(o - 2 * plane, c2);
                if (v1) __stcs(o - plane, c1);
                if (v0) __stcs(o, c0);
            }
            if (ARGMIN) {
                const float dq = (float)d0i;
                if (v3 && c3 < bc) { bc = c3; bd = dq - 3.f; }
                if (v2 && c2 < bc) { bc = c2; bd = dq - 2.f; }
                if (v1 && c1 < bc) { bc = c1; bd = dq - 1.f; }
                if (v0 && c0 < bc) { bc = c0; bd = dq; }
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
    };

    // Two consecutive FULL steps of a warp whose 32 lanes are all inside the row for both: the eight cost chains
    // (four 16-term FADD2 chains instead of two) are computed in ONE straight-line block — twice the independent
    // work per warp for the FP32 pipe, no per-step branches — and then consumed in step order exactly as `step`
    // does (stores, strict-< arg-min updates, hand-off shuffle), so results are bit-identical.
    auto step2 = [&](const int q) {
        const int j = m + q;
        const float* lp = Lp + (j - jb);
        float2 lo0 = make_float2(0.f, 0.f), hi0 = lo0, lo1 = lo0, hi1 = lo0;
#pragma unroll
        for (int c = 0; c < CV_C; ++c) {
            const float la = lp[c * CV_LW], lb = lp[c * CV_LW + 1];
            const float2 n0 = make_float2(-rq[c][0].x, -rq[c][0].y), n1 = make_float2(-rq[c][1].x, -rq[c][1].y);
            const float2 a0 = __fadd2_rn(make_float2(la, la), n0), a1 = __fadd2_rn(make_float2(la, la), n1);
            const float2 b0 = __fadd2_rn(make_float2(lb, lb), n0), b1 = __fadd2_rn(make_float2(lb, lb), n1);
            lo0 = __fadd2_rn(lo0, make_float2(fabsf(a0.x), fabsf(a0.y)));
            hi0 = __fadd2_rn(hi0, make_float2(fabsf(a1.x), fabsf(a1.y)));
            lo1 = __fadd2_rn(lo1, make_float2(fabsf(b0.x), fabsf(b0.y)));
            hi1 = __fadd2_rn(hi1, make_float2(fabsf(b1.x), fabsf(b1.y)));
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const float2 lo = h ? lo1 : lo0, hi = h ? hi1 : hi0;
            const int jj = j + h, d0i = 4 * (q + h);
            if (WRITE_CV) {
                float* o = cvrow + (ptrdiff_t)d0i * (ptrdiff_t)plane + jj;
                __stcs(o - 3 * plane, hi.y);
                __stcs(o - 2 * plane, hi.x);
                __stcs(o - plane, lo.y);
                __stcs(o, lo.x);
            }
            if (ARGMIN) {
                const float dq = (float)d0i;
                if (hi.y < bc) { bc = hi.y; bd = dq - 3.f; }
                if (hi.x < bc) { bc = hi.x; bd = dq - 2.f; }
                if (lo.y < bc) { bc = lo.y; bd = dq - 1.f; }
                if (lo.x < bc) { bc = lo.x; bd = dq; }
                if (lane == 0) {
                    pc[warp * LW + (jj - jb)] = bc;
                    pd[warp * LW + (jj - jb)] = bd;
                }
                bc = __shfl_down_sync(0xffffffffu, bc, 1);
                bd = __shfl_down_sync(0xffffffffu, bd, 1);
                if (lane == 31) { bc = INFINITY; bd = 0.f; }
            }
        }
    };

    if (m0 < je) {  // warp-uniform
        const int qfull = min((p.D - 1) >> 2, qmax);     // q in [1, qfull]: all four disparities valid
        step(0, std::false_type{});
        int q = 1;
        // all 32 lanes on for steps q and q+1  <=>  m0 + q >= jb  and  m0 + 31 + q + 1 < je  (and every lane has an m)
        if (m0 + 31 < je) {
            for (; q + 1 <= qfull && m0 + q < jb; ++q) step(q, std::true_type{});
            for (; q + 1 <= qfull && m0 + 32 + q < je; q += 2) step2(q);
        }
#pragma unroll 2
        for (; q <= qfull; ++q) step(q, std::true_type{});
        for (int q = max(qfull, 0) + 1; q <= qmax; ++q) step(q, std::false_type{});
        if (ARGMIN) {
            // columns still in flight: lane now holds the state of column m + qmax + 1
            const int j = m + qmax + 1;
            if (lane < 31 && j >= jb && j < je) {
                pc[warp * LW + (j - jb)] = bc;
                pd[warp * LW + (j - jb)] = bd;
            }
        }
    }
    __syncthreads();

    if (ARGMIN) {
        // merge per-warp partials: higher warp (larger m) <-> smaller d, so walk warps downwards;
        // the zero-filled shifts (d >= 4j+1, all |L|_1) enter last as the candidate d = 4j+1
        for (int jj = tid; jj < nj; jj += blockDim.x) {
            float c = INFINITY, d = 0.f;
            for (int wv = nwarps - 1; wv >= 0; --wv) {
                const float cc = pc[wv * LW + jj];
                if (cc < c) { c = cc; d = pd[wv * LW + jj]; }
            }
            const int dz = 4 * (jb + jj) + 1;
            if (dz < p.D && l1[jj] < c) { c = l1[jj]; d = (float)dz; }
            const size_t o = ((size_t)n * p.h + i) * p.w + jb + jj;
            if (p.min_cost) p.min_cost[o] = c;
            if (p.min_disp) p.min_disp[o] = d;
        }
    }
    if (WRITE_CV) {
        // zero-filled region: 4j+1 <= d < D
        const int jz = min(nj, max(0, (p.D + 2) / 4 - jb));   // local columns [0, jz) that may have one
        if (jz > 0) {
            for (int idx = tid; idx < p.D * jz; idx += blockDim.x) {
                const int d = idx / jz, jj = idx - d * jz;
                if (d >= 4 * (jb + jj) + 1) __stcs(cvrow + (size_t)d * plane + jb + jj, l1[jj]);
            }
        }
    }
}

template <bool WRITE_CV, bool ARGMIN>
__global__ void __launch_bounds__(CV_MAXW * 32, 2) cost_volume_kernel(CvP p) {
    cost_volume_body<WRITE_CV, ARGMIN>(p, (int)blockIdx.x);
}

// All levels of the tile-initialisation pyramid in ONE launch: the four coarse levels (6-30 us each, latency bound when
// launched alone) fill the tail of the finest level's last wave.  Blocks are ordered finest level first.
constexpr int CV_MAXLEV = 8;
struct CvPyr {
    CvP lv[CV_MAXLEV];
    int first[CV_MAXLEV + 1];   // first[l] = first block of level l (in launch order)
    int nlev;
};
template <bool WRITE_CV, bool ARGMIN>
__global__ void __launch_bounds__(CV_MAXW * 32, 2) cost_volume_pyramid_kernel(const __grid_constant__ CvPyr P) {
    int l = 0;
    while (l + 1 < P.nlev && (int)blockIdx.x >= P.first[l + 1]) ++l;
    cost_volume_body<WRITE_CV, ARGMIN>(P.lv[l], (int)blockIdx.x - P.first[l]);
}

}  // namespace

namespace {
// column blocking of one level; returns the dynamic shared-memory bytes for `nwarps_launch` warps, or 0 if unsupported
size_t cv_setup(CvP& p, const float* tile_l, const float* tile_r, int n, int h, int w, int max_disp, float* cv,
                float* min_cost, float* min_disp, int* nwarps_needed, int nwarps_launch) {
    const int qmax = (max_disp + 2) / 4;
    const int max_m = CV_MAXW * 32;
    if (qmax + 1 > max_m) return 0;
    int nblk = 1;
    while (nblk < w && codd_ceil_div(w, nblk) + (nblk > 1 ? qmax : 0) > max_m) ++nblk;
    const int JB = codd_ceil_div(w, nblk);
    nblk = codd_ceil_div(w, JB);
    const int nm = (nblk > 1) ? JB + qmax : w;   // single block: m in [0, w)
    if (nm > max_m) return 0;
    const int nwarps = codd_ceil_div(nm, 32);
    *nwarps_needed = nwarps;
    const int nw = nwarps_launch > 0 ? nwarps_launch : nwarps;
    const int RW = 4 * nm + 4;
    const int LW = (JB + 3) & ~3;
    p.L = tile_l; p.R = tile_r;
    p.N = n; p.h = h; p.w = w; p.D = max_disp; p.JB = JB; p.nblk = nblk;
    p.cv = cv; p.min_cost = min_cost; p.min_disp = min_disp;
    return (size_t)(CV_C * RW + CV_C * CV_LW + LW + 2 * nw * LW) * sizeof(float);
}
}  // namespace

extern "C" int codd_cost_volume(const float* tile_l, const float* tile_r, int n, int h, int w, int max_disp,
                                float* cv, float* min_cost, float* min_disp, void* stream) {
    if (!tile_l || !tile_r || n <= 0 || h <= 0 || w <= 0 || max_disp <= 0) return CODD_E_BADARG;
    if (!codd_aligned16(tile_r)) return CODD_E_ALIGN;   // bulk-TMA source rows (4w floats) must be 16-byte aligned
    const bool argmin = (min_cost != nullptr) || (min_disp != nullptr);
    if (!cv && !argmin) return CODD_E_BADARG;
    CvP p;
    int nwarps = 0;
    const size_t smem = cv_setup(p, tile_l, tile_r, n, h, w, max_disp, cv, min_cost, min_disp, &nwarps, 0);
    if (smem == 0 || smem > 227 * 1024) return CODD_E_SHAPE;
    dim3 grid((unsigned)(n * h * p.nblk)), block(32 * nwarps);
    cudaStream_t s = (cudaStream_t)stream;
    void (*kern)(CvP) = nullptr;
    static CoddDeviceOnce once[3];
    int slot;
    if (cv && argmin) { kern = cost_volume_kernel<true, true>; slot = 0; }
    else if (cv) { kern = cost_volume_kernel<true, false>; slot = 1; }
    else { kern = cost_volume_kernel<false, true>; slot = 2; }
    // opt in to the full 227 KB once per device (the per-launch request stays `smem`)
    if (int rc = codd_once_per_device(once[slot], [&] {
            return codd_max_dynamic_smem(kern);
        }))
        return rc;
    kern<<<grid, block, smem, s>>>(p);
    CODD_RETURN_IF_CUDA_ERROR();
    return 0;
}

extern "C" int codd_cost_volume_pyramid(int levels, const float* const* tile_l, const float* const* tile_r, int n,
                                        const int* h, const int* w, const int* max_disp, float* const* cv,
                                        float* const* min_cost, float* const* min_disp, void* stream) {
    if (levels <= 0 || levels > CV_MAXLEV || !tile_l || !tile_r || !h || !w || !max_disp || n <= 0) return CODD_E_BADARG;
    const bool want_cv = cv && cv[0];
    const bool argmin = (min_cost && min_cost[0]) || (min_disp && min_disp[0]);
    if (!want_cv && !argmin) return CODD_E_BADARG;
    CvPyr P;
    P.nlev = levels;
    // launch order: largest level first
    int order[CV_MAXLEV];
    for (int i = 0; i < levels; ++i) order[i] = i;
    for (int i = 0; i < levels; ++i)
        for (int j = i + 1; j < levels; ++j)
            if ((long long)h[order[j]] * w[order[j]] * max_disp[order[j]] > (long long)h[order[i]] * w[order[i]] * max_disp[order[i]]) {
                const int t = order[i]; order[i] = order[j]; order[j] = t;
            }
    size_t smem = 0;
    int nblocks = 0;
    for (int i = 0; i < levels; ++i) {
        const int l = order[i];
        if (!tile_l[l] || !tile_r[l] || h[l] <= 0 || w[l] <= 0 || max_disp[l] <= 0) return CODD_E_BADARG;
        if (!codd_aligned16(tile_r[l])) return CODD_E_ALIGN;
        if ((want_cv && !cv[l]) || (min_cost && min_cost[0] && !min_cost[l]) || (min_disp && min_disp[0] && !min_disp[l]))
            return CODD_E_BADARG;
        int nwarps = 0;
        const size_t sm = cv_setup(P.lv[i], tile_l[l], tile_r[l], n, h[l], w[l], max_disp[l], want_cv ? cv[l] : nullptr,
                                   (min_cost && min_cost[0]) ? min_cost[l] : nullptr,
                                   (min_disp && min_disp[0]) ? min_disp[l] : nullptr, &nwarps, CV_MAXW);
        if (sm == 0 || sm > 227 * 1024) return CODD_E_SHAPE;
        smem = sm > smem ? sm : smem;
        P.first[i] = nblocks;
        nblocks += n * h[l] * P.lv[i].nblk;
    }
    P.first[levels] = nblocks;
    void (*kern)(const CvPyr) = nullptr;
    static CoddDeviceOnce once[3];
    int slot;
    if (want_cv && argmin) { kern = cost_volume_pyramid_kernel<true, true>; slot = 0; }
    else if (want_cv) { kern = cost_volume_pyramid_kernel<true, false>; slot = 1; }
    else { kern = cost_volume_pyramid_kernel<false, true>; slot = 2; }
    if (int rc = codd_once_per_device(once[slot], [&] {
            return codd_max_dynamic_smem(kern);
        }))
        return rc;
    kern<<<(unsigned)nblocks, CV_MAXW * 32, smem, (cudaStream_t)stream>>>(P);
    CODD_RETURN_IF_CUDA_ERROR();
    return 0;
}
