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

__device__ __forceinline__ float codd_act(float v, int act, int ch) {
    switch (act) {
        case CODD_ACT_LEAKY: return v > 0.f ? v : v * CODD_LEAKY_SLOPE;
        case CODD_ACT_RELU: return fmaxf(v, 0.f);
        case CODD_ACT_RELU_CH0: return ch == 0 ? fmaxf(v, 0.f) : v;
        case CODD_ACT_SIGMOID: return 1.f / (1.f + __expf(-v));
        case CODD_ACT_MISH: {
            // x * tanh(softplus(x)); softplus with torch's threshold 20
            float sp = v > 20.f ? v : log1pf(expf(v));
            return v * tanhf(sp);
        }
        default: return v;
    }
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

static inline int codd_ceil_div(int a, int b) { return (a + b - 1) / b; }
