// Shared device/host helpers for the codd_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/codd_b200.h"

#define CODD_LEAKY_SLOPE 0.2f

#define CODD_RETURN_IF_CUDA_ERROR()                 \
    do {                                            \
        cudaError_t e__ = cudaGetLastError();       \
        if (e__ != cudaSuccess) return (int)e__;    \
    } while (0)

static inline bool codd_aligned16(const void* p) { return (((uintptr_t)p) & 15u) == 0; }

// One-time, PER-DEVICE function-attribute set-up (cudaFuncSetAttribute applies to the current device only), safe
// under concurrent first calls from several host threads.  Each call site owns one CoddDeviceOnce; `done` holds one bit
// per device ordinal, published with release/acquire so a thread that sees the bit also sees the attribute set.  The
// fast path is one relaxed-cost atomic load, no API call except cudaGetDevice — legal during stream capture.
#include <atomic>
#include <mutex>
struct CoddDeviceOnce {
    std::atomic<unsigned long long> done[4];   // 256 device ordinals
    std::mutex mu;
};
// Opt a kernel in to the largest dynamic shared-memory size its static allocation leaves room for (the opt-in limit
// counts static + dynamic bytes; asking for the full 227 KB fails for a kernel with any __shared__ variable).
template <typename K>
static inline cudaError_t codd_max_dynamic_smem(K kern) {
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, kern);
    if (e != cudaSuccess) return e;
    int dev = 0, optin = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev)) != cudaSuccess) return e;
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes);
}

template <typename F>
static inline int codd_once_per_device(CoddDeviceOnce& o, F&& setup) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    const unsigned long long bit = 1ull << (dev & 63);
    std::atomic<unsigned long long>& word = o.done[(dev >> 6) & 3];
    if (word.load(std::memory_order_acquire) & bit) return 0;
    std::lock_guard<std::mutex> lock(o.mu);
    if (word.load(std::memory_order_relaxed) & bit) return 0;
    e = setup();
    if (e != cudaSuccess) return (int)e;
    word.fetch_or(bit, std::memory_order_release);
    return 0;
}

// Transcendental activations (fusion heads, GRU gates): OUT OF LINE on purpose.  Inlined into a fully unrolled
// 64-element epilogue they blow a kernel up to several hundred KB of straight-line code that every warp
// executes once — the direct convolutions were instruction-fetch bound on it (ncu: stall_no_inst > 50 %).
static __device__ __noinline__ float codd_act_slow(float v, int act) {
    if (act == CODD_ACT_SIGMOID) return 1.f / (1.f + __expf(-v));
    if (act == CODD_ACT_TANH) return tanhf(v);
    const float sp = v > 20.f ? v : log1pf(expf(v));   // mish: x * tanh(softplus(x)), torch threshold 20
    return v * tanhf(sp);
}

__device__ __forceinline__ float codd_act(float v, int act, int ch) {
    // the activations of the stereo path (none / leaky / relu / relu on channel 0) are branch-free
    // selects; only the rare transcendental ones take a (call) branch.
    if (act >= CODD_ACT_SIGMOID) return codd_act_slow(v, act);
    const bool clamp = (act == CODD_ACT_RELU) || (act == CODD_ACT_RELU_CH0 && ch == 0);
    const float neg = (act == CODD_ACT_LEAKY) ? v * CODD_LEAKY_SLOPE : (clamp ? 0.f : v);
    return v > 0.f ? v : neg;
}

// Hoisted form for epilogues: evaluate the run-time activation code ONCE (ActSel), then apply it
// per element as max(v,0) + slope*min(v,0) with slope in {1 (none), 0.2 (leaky), 0 (relu)}; only
// the transcendental activations fall back to codd_act.
struct ActSel {
    float slope, slope0;
    bool simple;
    int act;
};
__device__ __forceinline__ ActSel codd_act_sel(int act) {
    ActSel a;
    a.act = act;
    a.simple = act <= CODD_ACT_RELU_CH0;
    a.slope = act == CODD_ACT_LEAKY ? CODD_LEAKY_SLOPE : (act == CODD_ACT_RELU ? 0.f : 1.f);
    a.slope0 = (act == CODD_ACT_RELU || act == CODD_ACT_RELU_CH0) ? 0.f : a.slope;
    return a;
}
__device__ __forceinline__ float codd_act_apply(const ActSel& a, float v, int ch) {
    if (a.simple) return fmaxf(v, 0.f) + (ch == 0 ? a.slope0 : a.slope) * fminf(v, 0.f);
    return codd_act(v, a.act, ch);
}

// one thread of a converged warp; unlike `lane == 0` ptxas knows the branch body runs in exactly one thread, so
// tcgen05 operands move to uniform registers without a per-instruction "elect / issue / loop" waterfall
__device__ __forceinline__ bool codd_elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
// acc[0..3] += a * w.{x,y,z,w} as two packed FFMA2 (sm_100a: one instruction = two independent fp32 FMAs, so
// bit-identical to four FFMA at half the issue slots and half the code size — the direct convolutions are
// instruction-fetch / issue bound, not FMA-pipe bound).
__device__ __forceinline__ void fma4(float* acc, float a, const float4& w) {
    const float2 aa = make_float2(a, a);
    float2* a2 = reinterpret_cast<float2*>(acc);
    a2[0] = __ffma2_rn(aa, make_float2(w.x, w.y), a2[0]);
    a2[1] = __ffma2_rn(aa, make_float2(w.z, w.w), a2[1]);
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256).  A thread that owns >= 32 contiguous bytes of an NHWC pixel moves
// them as whole 32-byte sectors: half the memory instructions of the float4 form and no half-written sectors — the
// lane-strided float4 epilogues cost ~8 L1 wavefronts per instruction.  `p` must be 32-byte aligned.
__device__ __forceinline__ void stg8(float* p, const float* v) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
                 "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}
__device__ __forceinline__ void ldg8(const float* p, float* v) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p));
}
static inline bool codd_aligned32(const void* p) { return (((uintptr_t)p) & 31u) == 0; }

static inline int codd_ceil_div(int a, int b) { return (a + b - 1) / b; }

// Programmatic dependent launch (PDL).  A kernel launched through codd_launch_pdl may start while the previous kernel
// of the stream is still running: its CTAs become resident as SMs free up and run their prologue (barrier set-up, TMEM
// allocation, weight staging — nothing the previous kernel writes) before they block in codd_pdl_wait(), which
// returns once every earlier kernel has completed and flushed.  Such a kernel MUST call codd_pdl_wait() before its
// first read of activations and before any global write.  codd_pdl_trigger() lets the NEXT kernel do the same with
// respect to this one.  Works under stream capture (the graph gets programmatic edges).
__device__ __forceinline__ void codd_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void codd_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
template <typename... KArgs, typename... Args>
static inline cudaError_t codd_launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                          Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
