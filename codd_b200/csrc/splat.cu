// K9 — SE(3) transform + forward point-splat warp (model/motion/motion.py:82-130 with
// PointsRendererWithDepth :28-42; replaces pytorch3d's PointsRasterizer + AlphaCompositor).
//
//   X' = T * pi^-1(depth);  (u,v) = pinhole(X');  every pixel whose centre (px+.5, py+.5) lies within
//   `radius` (NDC radius = radius/h, NDC scale 2/min(h,w)) of (u,v) receives the point;
//   per pixel the K=8 nearest-in-z points are alpha-composited front to back with weights
//   1 - d^2/r^2:  out = sum_k w_k prod_{j<k}(1-w_j) f_k;  zbuf = nearest z (0 when empty).
//
// Three launches: fill the key table with "empty"; scatter (one thread per source point inserts the 64-bit key
// (z bits << 32 | point index) into the K-slot sorted list of every pixel it covers, lock-free: atomicMin on slot k keeps
// the smaller key there and carries the larger one on to slot k+1, so when all insertions are done the slots hold the
// K smallest keys in ascending (z, index) order whatever the arrival order was — exact and deterministic, no candidate
// cap: pytorch3d's rasteriser likewise keeps the exact points_per_pixel nearest); composite (one thread per pixel
// re-projects its <= K points for the squared distances and blends front to back).
// Parity with pytorch3d is UNPINNED (extension not available here; checked against the brute-force restatement).
#include "common.cuh"

namespace {

constexpr int SPLAT_K = 8;
constexpr unsigned long long SPLAT_EMPTY = 0xffffffffffffffffull;

struct SplatP {
    const float* Ts;     // [N,h,w,7]  (tx,ty,tz,qx,qy,qz,qw)
    const float* depth;  // [N,h,w]
    const float* intr;   // [N,4]
    int N, h, w;
    float radius;        // in the reference's units (NDC radius = radius / h)
    unsigned long long* keys;   // [N,h,w,K] ascending (z bits << 32 | source pixel index), SPLAT_EMPTY = none
};

__device__ __forceinline__ void rot(const float* q, float X, float Y, float Z, float& ox, float& oy, float& oz) {
    const float ux = 2.f * (q[1] * Z - q[2] * Y), uy = 2.f * (q[2] * X - q[0] * Z), uz = 2.f * (q[0] * Y - q[1] * X);
    ox = X + q[3] * ux + (q[1] * uz - q[2] * uy);
    oy = Y + q[3] * uy + (q[2] * ux - q[0] * uz);
    oz = Z + q[3] * uz + (q[0] * uy - q[1] * ux);
}

// transformed point of source pixel `rem` of sample n: screen position (u, v) and depth Z.  One out-of-line body for both
// kernels: the composite pass must reproduce the scatter pass's arithmetic bit for bit.
__device__ __noinline__ void splat_project(const SplatP& p, int n, int rem, float& u, float& v, float& Z) {
    const size_t i = (size_t)n * p.h * p.w + rem;
    const int y = rem / p.w, x = rem - y * p.w;
    const float fx = __ldg(p.intr + n * 4), fy = __ldg(p.intr + n * 4 + 1), cx = __ldg(p.intr + n * 4 + 2), cy = __ldg(p.intr + n * 4 + 3);
    const float d = __ldg(p.depth + i);
    const float X0 = d * (((float)x - cx) / fx), Y0 = d * (((float)y - cy) / fy);
    const float* T = p.Ts + i * 7;
    float q[4] = {__ldg(T + 3), __ldg(T + 4), __ldg(T + 5), __ldg(T + 6)};
    float X, Y;
    rot(q, X0, Y0, d, X, Y, Z);
    X += __ldg(T); Y += __ldg(T + 1); Z += __ldg(T + 2);
    u = fx * X / Z + cx;
    v = fy * Y / Z + cy;
}

__global__ void __launch_bounds__(256) splat_scatter_kernel(SplatP p) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t hw = (size_t)p.h * p.w;
    if (i >= (size_t)p.N * hw) return;
    const int n = (int)(i / hw);
    const int rem = (int)(i - (size_t)n * hw);
    float u, v, Z;
    splat_project(p, n, rem, u, v, Z);
    if (!(Z > 0.f)) return;
    const float s = 2.f / (float)min(p.h, p.w);
    const float rn = p.radius / (float)p.h, r2 = rn * rn;
    const float rpx = rn / s;
    const int px0 = max(0, (int)floorf(u - 0.5f - rpx)), px1 = min(p.w - 1, (int)ceilf(u - 0.5f + rpx));
    const int py0 = max(0, (int)floorf(v - 0.5f - rpx)), py1 = min(p.h - 1, (int)ceilf(v - 0.5f + rpx));
    // Z > 0: the fp32 bit pattern orders like the value
    const unsigned long long mine = ((unsigned long long)__float_as_uint(Z) << 32) | (unsigned int)rem;
    for (int py = py0; py <= py1; ++py)
        for (int px = px0; px <= px1; ++px) {
            const float dx = (u - ((float)px + 0.5f)) * s, dy = (v - ((float)py + 0.5f)) * s;
            const float d2 = dx * dx + dy * dy;
            if (d2 < r2) {
                unsigned long long* slot = p.keys + ((size_t)n * hw + (size_t)py * p.w + px) * SPLAT_K;
                unsigned long long key = mine;
                for (int k = 0; k < SPLAT_K; ++k) {
                    const unsigned long long old = atomicMin(slot + k, key);
                    key = old > key ? old : key;            // the larger key moves on to the next slot
                    if (key == SPLAT_EMPTY) break;          // displaced an empty slot: done
                }
            }
        }
}

__global__ void __launch_bounds__(128) splat_composite_kernel(SplatP p, const float* __restrict__ feat, int ldf, int c,
                                                              float* __restrict__ out, int ldo, float* __restrict__ zbuf,
                                                              float* __restrict__ disp, float bf) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t hw = (size_t)p.h * p.w;
    if (i >= (size_t)p.N * hw) return;
    const int n = (int)(i / hw);
    const int rem = (int)(i - (size_t)n * hw);
    const int py = rem / p.w, px = rem - py * p.w;
    const float s = 2.f / (float)min(p.h, p.w);
    const float rn = p.radius / (float)p.h, r2 = rn * rn;
    float kz0 = 0.f, kw[SPLAT_K];
    int ki[SPLAT_K];
    int nk = 0;
    for (int k = 0; k < SPLAT_K; ++k) {
        const unsigned long long key = p.keys[i * SPLAT_K + k];
        if (key == SPLAT_EMPTY) break;
        const int src = (int)(unsigned int)(key & 0xffffffffull);
        float u, v, Z;
        splat_project(p, n, src, u, v, Z);
        const float dx = (u - ((float)px + 0.5f)) * s, dy = (v - ((float)py + 0.5f)) * s;
        if (k == 0) kz0 = __uint_as_float((unsigned int)(key >> 32));
        ki[nk] = src;
        kw[nk] = 1.f - (dx * dx + dy * dy) / r2;
        ++nk;
    }
    float* op = out + i * ldo;
    for (int ch = 0; ch < c; ++ch) {
        float acc = 0.f, trans = 1.f;
        for (int k = 0; k < nk; ++k) {
            acc += kw[k] * trans * __ldg(feat + ((size_t)n * hw + ki[k]) * ldf + ch);
            trans *= (1.f - kw[k]);
        }
        op[ch] = acc;
    }
    if (zbuf) zbuf[i] = kz0;
    if (disp) {
        // motion.py:190-193: disp = bf / (depth_warp + 1e-5); disp[disp > w] = 0
        const float dv = bf / (kz0 + 1e-5f);
        disp[i] = dv > (float)p.w ? 0.f : dv;
    }
}

}  // namespace

extern "C" size_t codd_splat_workspace_bytes(int n, int h, int w) {
    return (size_t)n * h * w * SPLAT_K * sizeof(unsigned long long);
}

extern "C" int codd_splat_warp(const float* Ts, const float* depth, const float* intr, const float* feat, int ldf, int c,
                               int n, int h, int w, float radius, float bf, float* out, int ldo, float* zbuf, float* disp,
                               void* workspace, size_t ws_bytes, void* stream) {
    if (!Ts || !depth || !intr || !feat || !out || !workspace || n <= 0 || h <= 0 || w <= 0 || c <= 0) return CODD_E_BADARG;
    if (ldf < c || ldo < c || !(radius > 0.f)) return CODD_E_SHAPE;
    if (ws_bytes < codd_splat_workspace_bytes(n, h, w)) return CODD_E_SHAPE;
    if ((((uintptr_t)workspace) & 7u) != 0) return CODD_E_ALIGN;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t npix = (size_t)n * h * w;
    SplatP p;
    p.Ts = Ts; p.depth = depth; p.intr = intr; p.N = n; p.h = h; p.w = w; p.radius = radius;
    p.keys = (unsigned long long*)workspace;
    cudaError_t e = cudaMemsetAsync(p.keys, 0xff, npix * SPLAT_K * sizeof(unsigned long long), s);
    if (e != cudaSuccess) return (int)e;
    splat_scatter_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, s>>>(p);
    splat_composite_kernel<<<(unsigned)((npix + 127) / 128), 128, 0, s>>>(p, feat, ldf, c, out, ldo, zbuf, disp, bf);
    CODD_RETURN_IF_CUDA_ERROR();
    return 0;
}
