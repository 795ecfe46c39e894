// Dense convolutions of the RAFT3D update block as a tcgen05 GEMM (model/motion/raft3d/raft3d.py:43-106,
// blocks/gru.py:10-35: 128..384-channel 3x3 / 7x7 / 1x1 layers at 1/8 resolution, ~60 GFLOP per iteration and sample —
// the "motion-head conv contractions").
//
//   out[m, n] = act( sum_k A[m, k] * B[n, k] + bias[n] + residual[m, n] )
//   m = output pixel (NHWC order), n = output channel, k = (tap, input channel)
//
// A is produced by codd_im2col_split below (one pass over the NHWC activation: the k-contiguous patch matrix plus its
// tf32 remainder A_lo = A - tf32(A)); B is the [Cout][taps*Cin] weight matrix, split into tf32 halves on the host.
// Precision: 3xTF32 as three K-segments accumulated in the same TMEM tile — (A, B_hi), (A, B_lo), (A_lo, B_hi); the tensor
// core reads the raw fp32 A and uses its top 19 bits.  With B_lo == NULL the kernel runs the single (A, B_hi) segment
// (plain TF32, what the reference's cuDNN path does by default on GPUs).
//
// Kernel: persistent CTAs over 128 x BN output tiles; warp 0 = TMA producer (A box 128 x 32 floats, B box BN x 32,
// SWIZZLE_128B, STAGES-deep mbarrier ring), warp 1 = MMA issuer (tcgen05.mma kind::tf32, M128 x BN x K8, accumulators
// double-buffered in TMEM), warps 2-5 = epilogue (tcgen05.ld, bias / residual / activation, NHWC store).  Out-of-range
// rows / columns / k come back as zeros from TMA, so M, N, K need no padding (row pitches must be 16-byte multiples).
#include <cuda.h>

#include "common.cuh"

namespace {

constexpr int GM_BM = 128;
constexpr int GM_BK = 32;     // floats per k-step stage row (128 bytes = one swizzle row)
constexpr int GM_THREADS = 192;

struct GmP {
    const float* bias;
    const float* res;
    float* out;
    int M, N, K, ldo, ldr, act;
    int tilesM, tilesN, nseg;   // nseg = 1 (TF32) or 3 (3xTF32)
};

__device__ __forceinline__ uint32_t gs_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void gbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void gbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void gbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void gbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!done) __nanosleep(20);
    }
}
__device__ __forceinline__ void gtma_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
                 "l"(map), "r"(bar), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ uint64_t gmake_desc(uint32_t saddr) {   // K-major, SWIZZLE_128B, 128-byte rows
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024u >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void gmma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.u32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void gcommit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void gld16(uint32_t taddr, float* v) {
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

template <int BN, int STAGES>
__global__ void __launch_bounds__(GM_THREADS, 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                const __grid_constant__ CUtensorMap mapAlo,
                                                                const __grid_constant__ CUtensorMap mapBhi,
                                                                const __grid_constant__ CUtensorMap mapBlo, GmP p) {
    constexpr uint32_t A_BYTES = GM_BM * GM_BK * 4;      // 16 KB
    constexpr uint32_t B_BYTES = BN * GM_BK * 4;
    constexpr uint32_t STAGE = A_BYTES + B_BYTES;
    constexpr uint32_t TMEM_COLS = (2 * BN <= 128) ? 128u : (2 * BN <= 256) ? 256u : 512u;
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((128u >> 4) << 24);

    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) unsigned long long bars[2 * STAGES + 4];
    __shared__ uint32_t tmem_base_slot;
    const uint32_t sbase = (gs_u32(smem_raw) + 1023u) & ~1023u;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar0 = gs_u32(&bars[0]);
    auto FULL = [&](int s) { return bar0 + (uint32_t)s * 8u; };
    auto EMPTY = [&](int s) { return bar0 + (uint32_t)(STAGES + s) * 8u; };
    auto ACCF = [&](int b) { return bar0 + (uint32_t)(2 * STAGES + b) * 8u; };
    auto ACCE = [&](int b) { return bar0 + (uint32_t)(2 * STAGES + 2 + b) * 8u; };

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            gbar_init(FULL(s), 1);
            gbar_init(EMPTY(s), 1);
        }
        for (int b = 0; b < 2; ++b) {
            gbar_init(ACCF(b), 1);
            gbar_init(ACCE(b), 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(gs_u32(&tmem_base_slot)),
                     "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_slot;
    const int ksteps = (p.K + GM_BK - 1) / GM_BK;
    const int ntiles = p.tilesM * p.tilesN;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int it = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
                const int m0 = (t / p.tilesN) * GM_BM, n0 = (t % p.tilesN) * BN;
                for (int seg = 0; seg < p.nseg; ++seg) {
                    const CUtensorMap* ma = seg == 2 ? &mapAlo : &mapA;
                    const CUtensorMap* mb = seg == 1 ? &mapBlo : &mapBhi;
                    for (int k = 0; k < ksteps; ++k, ++it) {
                        const int s = it % STAGES;
                        gbar_wait(EMPTY(s), (((uint32_t)(it / STAGES)) & 1u) ^ 1u);
                        gbar_expect_tx(FULL(s), STAGE);
                        gtma_2d(sbase + s * STAGE, ma, FULL(s), k * GM_BK, m0);
                        gtma_2d(sbase + s * STAGE + A_BYTES, mb, FULL(s), k * GM_BK, n0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (codd_elect_one()) {
            int it = 0, tl = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++tl) {
                const int ab = tl & 1;
                gbar_wait(ACCE(ab), (((uint32_t)(tl >> 1)) & 1u) ^ 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d = tmem + (uint32_t)(ab * BN);
                const int total = p.nseg * ksteps;
                for (int kk = 0; kk < total; ++kk, ++it) {
                    const int s = it % STAGES;
                    gbar_wait(FULL(s), ((uint32_t)(it / STAGES)) & 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint64_t da = gmake_desc(sbase + s * STAGE), db = gmake_desc(sbase + s * STAGE + A_BYTES);
#pragma unroll
                    for (int j = 0; j < GM_BK / 8; ++j)
                        gmma_tf32(d, da + (uint64_t)(j * 2), db + (uint64_t)(j * 2), IDESC, (kk > 0 || j > 0) ? 1u : 0u);
                    gcommit(EMPTY(s));
                }
                gcommit(ACCF(ab));
            }
        }
    } else {
        // ===================== epilogue (warps 2-5) =====================
        const int quarter = warp & 3;
        const ActSel asel = codd_act_sel(p.act);
        int tl = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++tl) {
            const int ab = tl & 1;
            const int m0 = (t / p.tilesN) * GM_BM, n0 = (t % p.tilesN) * BN;
            gbar_wait(ACCF(ab), ((uint32_t)(tl >> 1)) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int m = m0 + quarter * 32 + lane;
            const bool vec = ((p.ldo & 3) == 0) && ((((uintptr_t)p.out) & 15u) == 0);
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 16) {
                float acc[16];
                gld16(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(ab * BN + c0), acc);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (m < p.M && n0 + c0 < p.N) {
                    float* op = p.out + (size_t)m * p.ldo + n0 + c0;
                    const float* rp = p.res ? p.res + (size_t)m * p.ldr + n0 + c0 : nullptr;
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const int n = n0 + c0 + e;
                        if (n < p.N) {
                            float v = acc[e];
                            if (p.bias) v += __ldg(p.bias + n);
                            if (rp) v += __ldg(rp + e);
                            acc[e] = codd_act_apply(asel, v, n);
                        }
                    }
                    if (vec && n0 + c0 + 16 <= p.N) {
#pragma unroll
                        for (int e = 0; e < 16; e += 4)
                            *reinterpret_cast<float4*>(op + e) = make_float4(acc[e], acc[e + 1], acc[e + 2], acc[e + 3]);
                    } else {
#pragma unroll
                        for (int e = 0; e < 16; ++e)
                            if (n0 + c0 + e < p.N) op[e] = acc[e];
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            gbar_arrive(ACCE(ab));
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS));
    }
}

typedef CUresult (*PFN_gmEncode)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_gmEncode gm_get_encode() {
    static PFN_gmEncode fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (PFN_gmEncode)ptr;
    }
    return fn;
}

// 2-D map of a row-major [rows][cols] fp32 matrix with row pitch ld (floats); box = box_rows x 32 columns
int gm_make_map(PFN_gmEncode enc, CUtensorMap* map, const float* base, int rows, int cols, int ld, int box_rows) {
    const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
    const cuuint32_t box[2] = {(cuuint32_t)GM_BK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1u, 1u};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : CODD_E_SHAPE;
}

template <int BN, int STAGES>
int launch_gemm(const CUtensorMap& a, const CUtensorMap& alo, const CUtensorMap& bhi, const CUtensorMap& blo, GmP p,
                cudaStream_t s) {
    const size_t smem = (size_t)STAGES * (GM_BM * GM_BK * 4 + BN * GM_BK * 4) + 1024;
    auto kern = gemm_tc_kernel<BN, STAGES>;
    static CoddDeviceOnce once;   // one per template instantiation
    if (int rc = codd_once_per_device(once, [&] {
            return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        }))
        return rc;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    p.tilesM = codd_ceil_div(p.M, GM_BM);
    p.tilesN = codd_ceil_div(p.N, BN);
    const int ntiles = p.tilesM * p.tilesN;
    kern<<<ntiles < sms ? ntiles : sms, GM_THREADS, smem, s>>>(a, alo, bhi, blo, p);
    CODD_RETURN_IF_CUDA_ERROR();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// im2col + tf32 split: NHWC activation -> patch matrix A [M][taps*C] (and A_lo = A - tf32(A)), stride 1
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) im2col_split_kernel(const float* __restrict__ in, int ldi, int n, int h, int w, int c,
                                                           int kh, int kw, int ph, int pw, int dil, float* __restrict__ A,
                                                           float* __restrict__ Alo, int lda, size_t total4) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // over M * taps * (c/4)
    if (i >= total4) return;
    const int c4n = c >> 2;
    const int c4 = (int)(i % c4n);
    size_t t = i / c4n;
    const int tap = (int)(t % (kh * kw));
    const size_t m = t / (kh * kw);
    const int x = (int)(m % w);
    const int y = (int)((m / w) % h);
    const size_t s = m / ((size_t)w * h);
    const int ky = tap / kw, kx = tap - ky * kw;
    const int yy = y - ph + ky * dil, xx = x - pw + kx * dil;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (yy >= 0 && yy < h && xx >= 0 && xx < w) v = ldg4(in + ((s * h + yy) * (size_t)w + xx) * ldi + c4 * 4);
    const size_t o = m * lda + (size_t)tap * c + c4 * 4;
    *reinterpret_cast<float4*>(A + o) = v;
    if (Alo) {
        float4 l;
        l.x = __fsub_rn(v.x, __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u));
        l.y = __fsub_rn(v.y, __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u));
        l.z = __fsub_rn(v.z, __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u));
        l.w = __fsub_rn(v.w, __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u));
        *reinterpret_cast<float4*>(Alo + o) = l;
    }
}

}  // namespace

extern "C" int codd_im2col_split(const float* in, int ldi, int n, int h, int w, int c, int kh, int kw, int ph, int pw,
                                 int dil, float* A, float* A_lo, int lda, void* stream) {
    if (!in || !A || n <= 0 || h <= 0 || w <= 0 || c <= 0 || kh <= 0 || kw <= 0 || dil <= 0) return CODD_E_BADARG;
    if (c % 4 != 0 || ldi % 4 != 0 || ldi < c || lda % 4 != 0 || lda < kh * kw * c) return CODD_E_SHAPE;
    if (!codd_aligned16(in) || !codd_aligned16(A) || (A_lo && !codd_aligned16(A_lo))) return CODD_E_ALIGN;
    const size_t total4 = (size_t)n * h * w * kh * kw * (c / 4);
    im2col_split_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, ldi, n, h, w, c, kh, kw, ph,
                                                                                            pw, dil, A, A_lo, lda, total4);
    CODD_RETURN_IF_CUDA_ERROR();
    return 0;
}

extern "C" int codd_gemm_tc(const float* A, const float* A_lo, int lda, const float* B_hi, const float* B_lo, int ldb, int m,
                            int n, int k, const float* bias, const float* residual, int ldr, int act, float* out, int ldo,
                            void* stream) {
    if (!A || !B_hi || !out || m <= 0 || n <= 0 || k <= 0) return CODD_E_BADARG;
    if ((B_lo != nullptr) != (A_lo != nullptr)) return CODD_E_BADARG;     // 3xTF32 needs both remainders
    if (lda % 4 != 0 || ldb % 4 != 0 || lda < k || ldb < k || ldo < n || (residual && ldr < n)) return CODD_E_SHAPE;
    if (!codd_aligned16(A) || !codd_aligned16(B_hi) || (A_lo && !codd_aligned16(A_lo)) || (B_lo && !codd_aligned16(B_lo)))
        return CODD_E_ALIGN;
    PFN_gmEncode enc = gm_get_encode();
    if (!enc) return CODD_E_UNSUPPORTED;
    const int BN = n <= 64 ? 64 : (n <= 128 ? 128 : 256);
    CUtensorMap ma, malo, mbhi, mblo;
    int rc = gm_make_map(enc, &ma, A, m, k, lda, GM_BM);
    if (!rc) rc = gm_make_map(enc, &malo, A_lo ? A_lo : A, m, k, lda, GM_BM);
    if (!rc) rc = gm_make_map(enc, &mbhi, B_hi, n, k, ldb, BN);
    if (!rc) rc = gm_make_map(enc, &mblo, B_lo ? B_lo : B_hi, n, k, ldb, BN);
    if (rc) return rc;
    GmP p;
    p.bias = bias; p.res = residual; p.out = out;
    p.M = m; p.N = n; p.K = k; p.ldo = ldo; p.ldr = ldr; p.act = act;
    p.tilesM = p.tilesN = 0;
    p.nseg = A_lo ? 3 : 1;
    cudaStream_t s = (cudaStream_t)stream;
    if (BN == 64) return launch_gemm<64, 6>(ma, malo, mbhi, mblo, p, s);
    if (BN == 128) return launch_gemm<128, 6>(ma, malo, mbhi, mblo, p, s);
    return launch_gemm<256, 4>(ma, malo, mbhi, mblo, p, s);
}
