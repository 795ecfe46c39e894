// tcgen05 / TMA / mbarrier helpers of the strided tensor-core kernel (conv_tc_s2.cu).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace {

__device__ __forceinline__ uint32_t s2_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void s2_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void s2_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void s2_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
template <bool BACKOFF = true, int NS = 32>
__device__ __forceinline__ void s2_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (BACKOFF && !done) __nanosleep(NS);   // polling WARPS must not starve the working warps of their SM
                                                 // sub-partition; the single-thread roles (TMA, MMA issue) poll hot
    }
}
__device__ __forceinline__ void s2_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void s2_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void s2_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
template <bool ACC>
__device__ __forceinline__ void s2_mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc) {
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
template <bool ACC>
__device__ __forceinline__ void s2_mma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc) {
    if (ACC) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.eq.u32 p, 1, 1;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.eq.u32 p, 1, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc)
            : "memory");
    }
}
__device__ __forceinline__ void s2_ld16(uint32_t taddr, float* v) {
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

// shared-memory matrix descriptor, K-major, swizzled (64-byte rows: SWIZZLE_64B, 128-byte rows: SWIZZLE_128B), as
// conv_tc.cu make_desc
template <int KC>
__device__ __forceinline__ uint64_t s2_desc(uint32_t saddr) {
    constexpr uint32_t ROWB = KC * 4;
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);               // start address
    d |= (uint64_t)1 << 16;                                 // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)((8 * ROWB) >> 4) << 32;                 // stride byte offset: 8-row group pitch
    d |= (uint64_t)1 << 46;                                 // descriptor version (Blackwell)
    d |= ((KC == 32) ? 2ull : 4ull) << 61;                  // SWIZZLE_128B : SWIZZLE_64B
    return d;
}
// byte offset of 16-byte chunk j of row r inside a K-major swizzled tile whose base is 1024-aligned
template <int KC>
__device__ __forceinline__ uint32_t s2_swz(int r, int j) {
    const uint32_t off = (uint32_t)r * (KC * 4) + (uint32_t)j * 16u;
    return off ^ (((off >> 7) & ((KC == 32) ? 7u : 3u)) << 4);
}


typedef CUresult (*PFN_s2EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline PFN_s2EncodeTiled s2_get_encode() {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
        return (PFN_s2EncodeTiled)ptr;
    return nullptr;
}

}  // namespace
