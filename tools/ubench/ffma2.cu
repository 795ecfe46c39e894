// Microbenchmark: issue rate of FADD vs packed FFMA2 on sm_100a (decides K1's FP32-pipe ceiling).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fadd2 fadd2.cu && ./fadd2
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, float seed) {
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(seed + i + threadIdx.x, seed - i);
    const float2 b = make_float2(seed * 0.5f, seed * 0.25f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { a[i].x = __fmaf_rn(a[i].x, b.x, b.y); a[i].y = __fmaf_rn(a[i].y, b.y, b.x); }
            else a[i] = __ffma2_rn(a[i], b, make_float2(b.y, b.x));
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int warps = 4; warps <= 32; warps *= 2)
        for (int mode = 0; mode < 2; ++mode) {
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                if (mode == 0) k<0><<<148, warps * 32>>>(out, iters, 1.5f); else k<1><<<148, warps * 32>>>(out, iters, 1.5f);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
            }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double lane_ops = 148.0 * warps * 32 * iters * 16.0;
            printf("warps/SM %2d  %s  %.3f ms  %.1f lane-FMA/clk/SM (at 1.965 GHz)\n", warps, mode ? "FFMA2" : "FFMA ", ms,
                   lane_ops / (ms * 1e-3) / 148 / 1.965e9);
        }
    return 0;
}
