// tcgen05 / TMA / mbarrier helpers shared by the rolling-ring convolution kernels (conv_tc_ring.cu, conv_tc_ring2.cu).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace {

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
        if (!done) __nanosleep(32);
    }
}
__device__ __forceinline__ void mbar_wait_t(uint32_t bar, uint32_t parity, long long& acc, bool timing) {
    if (!timing) {
        mbar_wait(bar, parity);
        return;
    }
    const long long t0 = clock64();
    mbar_wait(bar, parity);
    acc += clock64() - t0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
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

template <int KC>
__device__ __forceinline__ uint32_t swz_off(int r, int j) {
    constexpr uint32_t ROWB = KC * 4;
    const uint32_t off = (uint32_t)r * ROWB + (uint32_t)j * 16u;
    constexpr uint32_t MASK = (KC == 32) ? 7u : 3u;
    return off ^ (((off >> 7) & MASK) << 4);
}
// byte offset of 16-byte chunk j of row r in a tile of RB-byte rows whose base is 1024-aligned
template <int RB>
__device__ __forceinline__ uint32_t swz_rb(int r, int j) {
    const uint32_t off = (uint32_t)r * RB + (uint32_t)j * 16u;
    constexpr uint32_t MASK = RB / 16 - 1;     // 1, 3, 7
    return off ^ (((off >> 7) & MASK) << 4);
}
// MMA with descriptors given as LOW 32-bit words (start address | LBO) plus compile-time HIGH words (SBO | version |
// swizzle mode): per-MMA operand arithmetic is one 32-bit add per descriptor and the issuing thread moves three
// instead of five values to uniform registers.
template <uint32_t HI_A, uint32_t HI_B, bool F16>
__device__ __forceinline__ void tc_mma_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
    if (F16) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
            "setp.ne.u32 p, %4, 0;\n\t"
            "mov.b64 da, {%1, %5};\n\t"
            "mov.b64 db, {%2, %6};\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}"
            ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "n"(HI_A), "n"(HI_B)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
            "setp.ne.u32 p, %4, 0;\n\t"
            "mov.b64 da, {%1, %5};\n\t"
            "mov.b64 db, {%2, %6};\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %3, p;\n\t}"
            ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "n"(HI_A), "n"(HI_B)
            : "memory");
    }
}
template <int RB>
__host__ __device__ constexpr uint32_t desc_hi() {   // high word of a K-major swizzled descriptor for RB-byte rows:
                                                     // 8-row group pitch (SBO) | descriptor version | SWIZZLE_32B/64B/128B
    return ((8u * RB) >> 4) | (1u << 14) | (((RB == 128) ? 2u : (RB == 64) ? 4u : 6u) << 29);
}
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | (1u << 16); }
constexpr float RG_LO_SCALE = 1024.f;   // x_lo and w_lo travel scaled by 2^10 (fp16 range); the epilogue undoes it

__device__ __forceinline__ void st_release_s32(uint32_t addr, int v) {
    asm volatile("st.release.cta.shared::cta.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_s32(uint32_t addr) {
    int v;
    asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline PFN_tmapEncodeTiled rg_get_encode() {
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

}  // namespace
