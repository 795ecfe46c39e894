// RAFT3D / Motion non-convolutional kernels.
//
//   raft_motion_info   projective_ops.py:44-52 + sampler_ops.py:26-28 + raft3d.py:227-240
//   avgpool2 / corr_lookup   blocks/corr.py:28-62 (all-pairs volume never materialised: a bilinear
//                        lookup in the pooled correlation volume equals the correlation with the
//                        pooled, bilinearly interpolated feature — everything is linear in fmap2)
//   se3_gn_step        se3_field.py:150-170 (lietorch_extras.se3_build_inplace + cholesky6x6 + exp)
//   cvx_upsample       se3_field.py:173-186
//   se3_upsample_flow  se3_field.py:189-192 + projective_ops.py:55-68 (raft3d.py:267-270)
//
// SE3 elements are (tx,ty,tz,qx,qy,qz,qw); the group math restates lietorch's published formulas
// (parity unpinned: the extension is not available, see oracle/motion_oracle.py).
#include "common.cuh"

namespace {

constexpr float P_EPS = 1e-5f;      // projective_ops.py:8
constexpr float SE3_EPS = 1e-6f;

struct V3 { float x, y, z; };
__device__ __forceinline__ V3 v3(float x, float y, float z) { V3 r = {x, y, z}; return r; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ V3 add(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 scale(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }

struct SE3 { V3 t; V3 qv; float qw; };

__device__ __forceinline__ SE3 se3_load(const float* p) {
    SE3 T;
    T.t = v3(p[0], p[1], p[2]);
    T.qv = v3(p[3], p[4], p[5]);
    T.qw = p[6];
    return T;
}
__device__ __forceinline__ void se3_store(float* p, const SE3& T) {
    p[0] = T.t.x; p[1] = T.t.y; p[2] = T.t.z; p[3] = T.qv.x; p[4] = T.qv.y; p[5] = T.qv.z; p[6] = T.qw;
}
__device__ __forceinline__ V3 quat_rot(V3 qv, float qw, V3 X) {
    const V3 uv = scale(cross(qv, X), 2.f);
    return add(add(X, scale(uv, qw)), cross(qv, uv));
}
__device__ __forceinline__ V3 se3_act(const SE3& T, V3 X) { return add(quat_rot(T.qv, T.qw, X), T.t); }

__device__ __forceinline__ SE3 se3_exp(const float* xi) {
    const V3 tau = v3(xi[0], xi[1], xi[2]), phi = v3(xi[3], xi[4], xi[5]);
    const float th2 = phi.x * phi.x + phi.y * phi.y + phi.z * phi.z;
    const float th = sqrtf(th2);
    float imag, real, a, b;
    if (th2 < SE3_EPS) {
        imag = 0.5f - th2 / 48.f + th2 * th2 / 3840.f;
        real = 1.f - th2 / 8.f + th2 * th2 / 384.f;
        a = 0.5f - th2 / 24.f;
        b = 1.f / 6.f - th2 / 120.f;
    } else {
        imag = sinf(0.5f * th) / th;
        real = cosf(0.5f * th);
        a = (1.f - cosf(th)) / th2;
        b = (th - sinf(th)) / (th2 * th);
    }
    SE3 T;
    const V3 pxt = cross(phi, tau);
    T.t = add(add(tau, scale(pxt, a)), scale(cross(phi, pxt), b));
    T.qv = scale(phi, imag);
    T.qw = real;
    return T;
}

__device__ __forceinline__ void se3_log(const SE3& T, float* xi) {
    const float n2 = T.qv.x * T.qv.x + T.qv.y * T.qv.y + T.qv.z * T.qv.z;
    const float n = sqrtf(n2);
    float k;
    if (n < SE3_EPS) {
        const float w = fabsf(T.qw) < SE3_EPS ? SE3_EPS : T.qw;
        k = 2.f / w - (2.f / 3.f) * n2 / (w * w * w);
    } else if (fabsf(T.qw) < SE3_EPS) {
        k = (T.qw >= 0.f ? 3.14159265358979f : -3.14159265358979f) / n;
    } else {
        k = 2.f * atanf(n / T.qw) / n;
    }
    const V3 phi = scale(T.qv, k);
    const float th2 = phi.x * phi.x + phi.y * phi.y + phi.z * phi.z;
    const float th = sqrtf(th2);
    const float c = th2 < SE3_EPS ? 1.f / 12.f : (1.f - th * sinf(th) / (2.f * (1.f - cosf(th)))) / th2;
    const V3 pxt = cross(phi, T.t);
    const V3 tau = add(add(T.t, scale(pxt, -0.5f)), scale(cross(phi, pxt), c));
    xi[0] = tau.x; xi[1] = tau.y; xi[2] = tau.z; xi[3] = phi.x; xi[4] = phi.y; xi[5] = phi.z;
}

__device__ __forceinline__ SE3 se3_mul(const SE3& A, const SE3& B) {
    SE3 C;
    C.t = add(quat_rot(A.qv, A.qw, B.t), A.t);
    const V3 cr = cross(A.qv, B.qv);
    C.qv = add(add(scale(B.qv, A.qw), scale(A.qv, B.qw)), cr);
    C.qw = A.qw * B.qw - (A.qv.x * B.qv.x + A.qv.y * B.qv.y + A.qv.z * B.qv.z);
    return C;
}

// torch-style bilinear grid_sample (zeros padding, align_corners=True) of a single-channel map at pixel (px,py)
__device__ __forceinline__ float bilerp_zero(const float* img, int H, int W, float px, float py) {
    // the reference normalises 2*c/(W-1) - 1 and torch un-normalises ((g+1)/2)*(W-1)
    const float gx = __fadd_rn(__fdiv_rn(__fmul_rn(2.f, px), (float)(W - 1)), -1.f);
    const float gy = __fadd_rn(__fdiv_rn(__fmul_rn(2.f, py), (float)(H - 1)), -1.f);
    const float ix = __fmul_rn(__fmul_rn(__fadd_rn(gx, 1.f), 0.5f), (float)(W - 1));
    const float iy = __fmul_rn(__fmul_rn(__fadd_rn(gy, 1.f), 0.5f), (float)(H - 1));
    const float fx = floorf(ix), fy = floorf(iy);
    const float tw = ix - fx, te = 1.f - tw, tn = iy - fy, ts = 1.f - tn;
    const int x0 = (int)fmaxf(fminf(fx, (float)W), -2.f), y0 = (int)fmaxf(fminf(fy, (float)H), -2.f);
    auto at = [&](int yy, int xx) { return (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(img + (size_t)yy * W + xx) : 0.f; };
    float v = __fmul_rn(at(y0, x0), __fmul_rn(ts, te));
    v = __fmaf_rn(at(y0, x0 + 1), __fmul_rn(ts, tw), v);
    v = __fmaf_rn(at(y0 + 1, x0), __fmul_rn(tn, te), v);
    v = __fmaf_rn(at(y0 + 1, x0 + 1), __fmul_rn(tn, tw), v);
    return v;
}

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) raft_motion_info_kernel(const float* __restrict__ Ts,
                                                               const float* __restrict__ depth1,
                                                               const float* __restrict__ depth2_inv,
                                                               const float* __restrict__ intr, int N, int h, int w,
                                                               float* __restrict__ xyz, float* __restrict__ info,
                                                               int ldi) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, n = blockIdx.z;
    if (x >= w) return;
    const size_t pix = ((size_t)n * h + y) * w + x;
    const float fx = __ldg(intr + n * 4), fy = __ldg(intr + n * 4 + 1), cx = __ldg(intr + n * 4 + 2), cy = __ldg(intr + n * 4 + 3);
    const float d = __ldg(depth1 + pix);
    const V3 X0 = v3(d * (((float)x - cx) / fx), d * (((float)y - cy) / fy), d);
    const SE3 T = se3_load(Ts + pix * 7);
    const V3 X1 = se3_act(T, X0);
    const float Z = X1.z + P_EPS;
    const float u = fx * (X1.x / Z) + cx, v = fy * (X1.y / Z) + cy, zi = 1.f / Z;
    xyz[pix * 3] = u; xyz[pix * 3 + 1] = v; xyz[pix * 3 + 2] = zi;
    const float zinv = bilerp_zero(depth2_inv + (size_t)n * h * w, h, w, u, v);
    float tw[6];
    se3_log(T, tw);
    float o[9];
    o[0] = u - (float)x; o[1] = v - (float)y;
#pragma unroll
    for (int i = 0; i < 6; ++i) o[2 + i] = 10.f * tw[i];
    o[8] = 10.f * (zinv - zi);
    float* ip = info + pix * ldi;
#pragma unroll
    for (int i = 0; i < 9; ++i) ip[i] = fminf(fmaxf(o[i], -50.f), 50.f);
}

// ------------------------------------------------------------------------------------------
__global__ void avgpool2_nhwc_kernel(const float* __restrict__ in, int ldi, int N, int h, int w, int c,
                                     float* __restrict__ out, int ldo) {
    const size_t total = (size_t)N * (h / 2) * (w / 2) * (c / 4);
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c4 = (int)(i % (c / 4));
    size_t t = i / (c / 4);
    const int x = (int)(t % (w / 2));
    t /= (w / 2);
    const int y = (int)(t % (h / 2));
    const int n = (int)(t / (h / 2));
    const float* b = in + (((size_t)n * h + 2 * y) * w + 2 * x) * ldi + c4 * 4;
    const float4 p = ldg4(b), q = ldg4(b + ldi), r = ldg4(b + (size_t)w * ldi), s = ldg4(b + (size_t)w * ldi + ldi);
    float4 o;
    o.x = (p.x + q.x + r.x + s.x) * 0.25f; o.y = (p.y + q.y + r.y + s.y) * 0.25f;
    o.z = (p.z + q.z + r.z + s.z) * 0.25f; o.w = (p.w + q.w + r.w + s.w) * 0.25f;
    *reinterpret_cast<float4*>(out + (((size_t)n * (h / 2) + y) * (w / 2) + x) * ldo + c4 * 4) = o;
}

struct CorrP {
    const float* f1; int ld1;          // [N,h,w,C]
    const float* f2[4]; int ld2[4];    // pooled fmap2 pyramid, level l is [N,h>>l,w>>l,C]
    const float* coords; int ldc;      // [N,h,w,>=2] (x,y) at level 0
    int N, h, w, C, levels, radius;
    float* out; int ldo;               // [N,h,w,levels*(2r+1)^2]
};

constexpr int CORR_MAXC = 256;

__global__ void __launch_bounds__(64) corr_lookup_kernel(CorrP p) {
    __shared__ __align__(16) float s_f1[CORR_MAXC];
    __shared__ float s_g[8][8];   // [x index][y index]
    const int pix = blockIdx.x;
    const int n = pix / (p.h * p.w);
    const int t = threadIdx.x;
    for (int c = t; c < p.C; c += 64) s_f1[c] = __ldg(p.f1 + (size_t)pix * p.ld1 + c) * (1.f / 16.f);   // (f1/4).(f2/4)
    const float cx = __ldg(p.coords + (size_t)pix * p.ldc), cy = __ldg(p.coords + (size_t)pix * p.ldc + 1);
    const int rd = 2 * p.radius + 1;
    __syncthreads();
    for (int l = 0; l < p.levels; ++l) {
        const int hl = p.h >> l, wl = p.w >> l;
        const float sc = 1.f / (float)(1 << l);
        const float x = cx * sc, y = cy * sc;
        const float fx = floorf(x), fy = floorf(y);
        const float dx = x - fx, dy = y - fy;
        const int gi = t & 7, gj = t >> 3;
        float dot = 0.f;
        if (gi <= rd && gj <= rd) {
            const float xf = fx - (float)p.radius + (float)gi, yf = fy - (float)p.radius + (float)gj;
            if (xf >= 0.f && xf < (float)wl && yf >= 0.f && yf < (float)hl) {
                const float* q = p.f2[l] + (((size_t)n * hl + (int)yf) * wl + (int)xf) * p.ld2[l];
                for (int c = 0; c < p.C; c += 4) {
                    const float4 v = ldg4(q + c);
                    const float4 a = *reinterpret_cast<const float4*>(&s_f1[c]);
                    dot = fmaf(a.x, v.x, dot); dot = fmaf(a.y, v.y, dot); dot = fmaf(a.z, v.z, dot); dot = fmaf(a.w, v.w, dot);
                }
            }
        }
        s_g[gi][gj] = dot;
        __syncthreads();
        if (t < rd * rd) {
            const int i = t / rd, j = t - i * rd;   // i: x offset, j: y offset
            const float v = (1.f - dx) * (1.f - dy) * s_g[i][j] + dx * (1.f - dy) * s_g[i + 1][j] +
                            (1.f - dx) * dy * s_g[i][j + 1] + dx * dy * s_g[i + 1][j + 1];
            p.out[(size_t)pix * p.ldo + l * rd * rd + t] = v;
        }
        __syncthreads();
    }
}

// Cooperative variant (C a multiple of 32): in the kernel above every lane walks its own neighbour's C floats, so a warp-wide
// 128-bit load touches 32 different lines (one sector each) — the kernel is bound by L1 tag / data-pipe wavefronts (0.93 ms
// per call, 16 calls per frame).  Here EIGHT lanes share one window position and read a contiguous 128-byte segment of it
// per step (a warp instruction covers 4 positions = 4 full lines), accumulate their part of the dot product over the C / 32
// segments and combine with three shuffles.  Same bilinear blend; the dot products are summed in a different order.
__global__ void __launch_bounds__(64) corr_lookup_coop_kernel(CorrP p) {
    __shared__ __align__(16) float s_f1[CORR_MAXC];
    __shared__ float s_g[8][8];   // [x index][y index]
    const int pix = blockIdx.x;
    const int n = pix / (p.h * p.w);
    const int t = threadIdx.x, lane = t & 31, wv = t >> 5;
    for (int c = t; c < p.C; c += 64) s_f1[c] = __ldg(p.f1 + (size_t)pix * p.ld1 + c) * (1.f / 16.f);   // (f1/4).(f2/4)
    const float cx = __ldg(p.coords + (size_t)pix * p.ldc), cy = __ldg(p.coords + (size_t)pix * p.ldc + 1);
    const int rd = 2 * p.radius + 1;
    const int sub = lane & 7;                 // 16-byte chunk of the 128-byte segment
    const int nseg = p.C >> 5;
    __syncthreads();
    for (int l = 0; l < p.levels; ++l) {
        const int hl = p.h >> l, wl = p.w >> l;
        const float sc = 1.f / (float)(1 << l);
        const float x = cx * sc, y = cy * sc;
        const float fx = floorf(x), fy = floorf(y);
        const float dx = x - fx, dy = y - fy;
#pragma unroll 2
        for (int g = 0; g < 8; ++g) {
            const int pos = wv * 32 + g * 4 + (lane >> 3);     // window position of this lane's group of eight
            const int gi = pos & 7, gj = pos >> 3;
            float dot = 0.f;
            if (gi <= rd && gj <= rd) {
                const float xf = fx - (float)p.radius + (float)gi, yf = fy - (float)p.radius + (float)gj;
                if (xf >= 0.f && xf < (float)wl && yf >= 0.f && yf < (float)hl) {
                    const float* q = p.f2[l] + (((size_t)n * hl + (int)yf) * wl + (int)xf) * p.ld2[l] + sub * 4;
                    const float* f = s_f1 + sub * 4;
                    for (int sg = 0; sg < nseg; ++sg) {
                        const float4 v = ldg4(q + sg * 32);
                        const float4 a = *reinterpret_cast<const float4*>(f + sg * 32);
                        dot = fmaf(a.x, v.x, dot); dot = fmaf(a.y, v.y, dot); dot = fmaf(a.z, v.z, dot); dot = fmaf(a.w, v.w, dot);
                    }
                }
            }
            dot += __shfl_xor_sync(0xffffffffu, dot, 1);
            dot += __shfl_xor_sync(0xffffffffu, dot, 2);
            dot += __shfl_xor_sync(0xffffffffu, dot, 4);
            if (sub == 0) s_g[gi][gj] = dot;
        }
        __syncthreads();
        if (t < rd * rd) {
            const int i = t / rd, j = t - i * rd;   // i: x offset, j: y offset
            const float v = (1.f - dx) * (1.f - dy) * s_g[i][j] + dx * (1.f - dy) * s_g[i + 1][j] +
                            (1.f - dx) * dy * s_g[i][j + 1] + dx * dy * s_g[i + 1][j + 1];
            p.out[(size_t)pix * p.ldo + l * rd * rd + t] = v;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// K12: one warp per centre pixel.  Lanes stride over the (2R+1)^2 window, each accumulating the 21
// unique entries of J^T W J and the 6 of J^T W r; warp reduction; damping (lm*H + ep) on the
// diagonal; in-register 6x6 Cholesky solve; T <- exp(dx) * T.
struct GnP {
    const float* Ts;       // [N,h,w,7]
    const float* ae; int lda;   // [N,h,w,32] NHWC (un-scaled; /8 applied here)
    const float* target; int ldt;   // [N,h,w,3] NHWC: coords1_xyz + delta
    const float* weight; int ldw;   // [N,h,w,3] NHWC
    const float* depth;    // [N,h,w]
    const float* intr;     // [N,4]
    int N, h, w, radius;
    float lm, ep;
    float* out;            // [N,h,w,7]
};

// Second version (round 2).  The first one gave every centre its own warp that read its 65 x 65 neighbours straight
// from global memory: lane k fetched neighbour k's 128-byte embedding with eight 128-bit loads, every one of which touched
// 32 different cache lines — ~256 L1 tag wavefronts per 32 (centre, neighbour) pairs, 1.2 G wavefronts per call, which
// IS the 3.3 ms it took (the arithmetic is ~0.8 ms).  Now a CTA owns GN_CX = 8 horizontally adjacent centres (one warp
// each), whose windows overlap by 57 of 65 columns; the union window is streamed through shared memory GN_ROWS rows at a
// time (cp.async, double buffered): embeddings with a pitch of 36 floats (so the 128-bit reads of 8 neighbouring lanes hit
// 8 different bank groups), the 7 per-neighbour scalars as separate rows.  Each neighbour record is fetched from L2 once
// per CTA (8 centres) instead of once per centre, and the L1 data pipe sees 4 wavefronts per 128-bit load instead of 32.
constexpr int GN_CX = 8;            // centres (= warps) per CTA
constexpr int GN_ROWS = 4;          // window rows per stage
constexpr int GN_MAXR = 32;         // largest radius the staged kernel handles (reference: 32)
constexpr int GN_WMAX = 2 * GN_MAXR + GN_CX;      // 72 window columns per CTA
constexpr int GN_AEP = 36;          // embedding pitch (floats)
constexpr int GN_REC = GN_ROWS * GN_WMAX;         // records per stage
constexpr int GN_STAGE_FLOATS = GN_REC * (GN_AEP + 8);      // embeddings + 7 scalar rows (+1 spare)

__device__ __forceinline__ void gn_cp16(float* dst, const float* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void gn_cp4(float* dst, const float* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

__global__ void __launch_bounds__(GN_CX * 32, 2) se3_gn_step_kernel(GnP p) {
    extern __shared__ __align__(16) float gn_smem[];      // 2 stages
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n = blockIdx.z, y = blockIdx.y, xc0 = blockIdx.x * GN_CX;
    const int x = xc0 + warp;
    const bool on = x < p.w;
    const int R = p.radius;
    const int y0 = max(0, y - R), y1 = min(p.h - 1, y + R);
    const int cx0 = max(0, xc0 - R), cx1 = min(p.w - 1, xc0 + GN_CX - 1 + R);     // CTA window columns
    const int cw = cx1 - cx0 + 1;
    const int nchunks = (y1 - y0 + GN_ROWS) / GN_ROWS;

    auto stage_load = [&](int chunk, int buf) {
        float* sb = gn_smem + buf * GN_STAGE_FLOATS;
        float* s_ae = sb;
        float* s_sc = sb + GN_REC * GN_AEP;
        const int r0 = y0 + chunk * GN_ROWS;
        const int rows = min(GN_ROWS, y1 - r0 + 1);
        // embeddings: 8 x 16 bytes per record, consecutive threads on consecutive pieces (coalesced 128-byte rows)
        for (int e = tid; e < rows * cw * 8; e += GN_CX * 32) {
            const int rec = e >> 3, piece = e & 7;
            const int rr = rec / cw, cc = rec - rr * cw;
            const size_t q = ((size_t)n * p.h + r0 + rr) * p.w + cx0 + cc;
            gn_cp16(s_ae + (rr * GN_WMAX + cc) * GN_AEP + piece * 4, p.ae + q * p.lda + piece * 4);
        }
        for (int e = tid; e < rows * cw; e += GN_CX * 32) {
            const int rr = e / cw, cc = e - rr * cw;
            const size_t q = ((size_t)n * p.h + r0 + rr) * p.w + cx0 + cc;
            float* d = s_sc + rr * GN_WMAX + cc;
            gn_cp4(d, p.depth + q);
            gn_cp4(d + GN_REC, p.target + q * p.ldt);
            gn_cp4(d + 2 * GN_REC, p.target + q * p.ldt + 1);
            gn_cp4(d + 3 * GN_REC, p.target + q * p.ldt + 2);
            gn_cp4(d + 4 * GN_REC, p.weight + q * p.ldw);
            gn_cp4(d + 5 * GN_REC, p.weight + q * p.ldw + 1);
            gn_cp4(d + 6 * GN_REC, p.weight + q * p.ldw + 2);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    stage_load(0, 0);

    const int pix = (n * p.h + y) * p.w + min(x, p.w - 1);
    const float fx = __ldg(p.intr + n * 4), fy = __ldg(p.intr + n * 4 + 1), cx = __ldg(p.intr + n * 4 + 2), cy = __ldg(p.intr + n * 4 + 3);
    const SE3 T = se3_load(p.Ts + (size_t)pix * 7);
    float aei[32];
#pragma unroll
    for (int c = 0; c < 32; c += 4) {
        const float4 v = ldg4(p.ae + (size_t)pix * p.lda + c);
        aei[c] = v.x * 0.125f; aei[c + 1] = v.y * 0.125f; aei[c + 2] = v.z * 0.125f; aei[c + 3] = v.w * 0.125f;
    }
    float Hm[21], g[6];
#pragma unroll
    for (int i = 0; i < 21; ++i) Hm[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 6; ++i) g[i] = 0.f;
    const int x0 = max(0, x - R), x1 = min(p.w - 1, x + R);      // this centre's columns
    const int ww = x1 - x0 + 1, lc0 = x0 - cx0;

    for (int chunk = 0; chunk < nchunks; ++chunk) {
        const int buf = chunk & 1;
        if (chunk + 1 < nchunks) {
            stage_load(chunk + 1, buf ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const float* s_ae = gn_smem + buf * GN_STAGE_FLOATS;
        const float* s_sc = s_ae + GN_REC * GN_AEP;
        const int r0 = y0 + chunk * GN_ROWS;
        const int rows = min(GN_ROWS, y1 - r0 + 1);
        const int cnt = on ? rows * ww : 0;
        for (int k = lane; k < cnt; k += 32) {
            const int rr = k / ww, cc = k - rr * ww;
            const int yy = r0 + rr, xx = x0 + cc;
            const int rec = rr * GN_WMAX + lc0 + cc;
            const float* ap = s_ae + rec * GN_AEP;
            float d2 = 0.f;
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
                const float4 v = *reinterpret_cast<const float4*>(ap + c);
                const float e0 = v.x * 0.125f - aei[c], e1 = v.y * 0.125f - aei[c + 1], e2 = v.z * 0.125f - aei[c + 2],
                            e3 = v.w * 0.125f - aei[c + 3];
                d2 += e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3;
            }
            const float aff = 1.f / (1.f + expf(d2));                      // sigmoid(-d2)
            const float dj = s_sc[rec];
            const V3 Xj = v3(dj * (((float)xx - cx) / fx), dj * (((float)yy - cy) / fy), dj);
            const V3 Y = se3_act(T, Xj);
            const float iz = 1.f / Y.z;
            const float r0_ = s_sc[rec + GN_REC] - (fx * Y.x * iz + cx);
            const float r1_ = s_sc[rec + 2 * GN_REC] - (fy * Y.y * iz + cy);
            const float r2_ = s_sc[rec + 3 * GN_REC] - iz;
            const float w0 = aff * s_sc[rec + 4 * GN_REC], w1 = aff * s_sc[rec + 5 * GN_REC], w2 = aff * s_sc[rec + 6 * GN_REC];
            // rows of J = J_pi * [I | -[Y]x]
            const float a = fx * iz, b = -fx * Y.x * iz * iz, c = fy * iz, e = -fy * Y.y * iz * iz, f = -iz * iz;
            float J0[6] = {a, 0.f, b, b * Y.y, a * Y.z - b * Y.x, -a * Y.y};
            float J1[6] = {0.f, c, e, -c * Y.z + e * Y.y, -e * Y.x, c * Y.x};
            float J2[6] = {0.f, 0.f, f, f * Y.y, -f * Y.x, 0.f};
            int idx = 0;
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                g[i] += w0 * J0[i] * r0_ + w1 * J1[i] * r1_ + w2 * J2[i] * r2_;
#pragma unroll
                for (int j = i; j < 6; ++j) Hm[idx++] += w0 * J0[i] * J0[j] + w1 * J1[i] * J1[j] + w2 * J2[i] * J2[j];
            }
        }
        __syncthreads();       // the buffer is refilled two iterations later
    }
    if (!on) return;
#pragma unroll
    for (int i = 0; i < 21; ++i)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) Hm[i] += __shfl_xor_sync(0xffffffffu, Hm[i], o);
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) g[i] += __shfl_xor_sync(0xffffffffu, g[i], o);
    if (lane != 0) return;
    // unpack, damp the diagonal, Cholesky H = L L^T, solve
    float A[6][6];
    {
        int idx = 0;
        for (int i = 0; i < 6; ++i)
            for (int j = i; j < 6; ++j) { A[i][j] = Hm[idx]; A[j][i] = Hm[idx]; ++idx; }
        for (int i = 0; i < 6; ++i) A[i][i] += p.lm * A[i][i] + p.ep;
    }
    float L[6][6];
    bool ok = true;
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j <= i; ++j) {
            float s = A[i][j];
            for (int k = 0; k < j; ++k) s -= L[i][k] * L[j][k];
            if (i == j) {
                if (!(s > 0.f)) { ok = false; s = 1.f; }
                L[i][i] = sqrtf(s);
            } else {
                L[i][j] = s / L[j][j];
            }
        }
    float z[6], dxv[6];
    for (int i = 0; i < 6; ++i) {
        float s = g[i];
        for (int k = 0; k < i; ++k) s -= L[i][k] * z[k];
        z[i] = s / L[i][i];
    }
    for (int i = 5; i >= 0; --i) {
        float s = z[i];
        for (int k = i + 1; k < 6; ++k) s -= L[k][i] * dxv[k];
        dxv[i] = s / L[i][i];
    }
    if (!ok)
        for (int i = 0; i < 6; ++i) dxv[i] = 0.f;       // non-PD system: leave the transform unchanged
    se3_store(p.out + (size_t)pix * 7, se3_mul(se3_exp(dxv), T));
}

// ------------------------------------------------------------------------------------------
// convex up-sampling x8: out[n,8y+i,8x+j,:] = sum_k softmax_k(mask[n,y,x,k*64+i*8+j]) * data[n,y+ky-1,x+kx-1,:]
template <int MODE>   // 0: plain data -> out;  1: data = twist, out = (SE3 exp, induced flow)
__global__ void __launch_bounds__(64) cvx_upsample_kernel(const float* __restrict__ data, int ldd, int dim,
                                                          const float* __restrict__ mask, int ldm, int N, int h, int w,
                                                          float* __restrict__ out, int ldo,
                                                          const float* __restrict__ depth, const float* __restrict__ intr,
                                                          float* __restrict__ flow) {
    const int pix = blockIdx.x;
    const int n = pix / (h * w);
    const int rem = pix - n * h * w;
    const int y = rem / w, x = rem - y * w;
    const int t = threadIdx.x, i = t >> 3, j = t & 7;
    float m[9], mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        m[k] = __ldg(mask + (size_t)pix * ldm + k * 64 + t);
        mx = fmaxf(mx, m[k]);
    }
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        m[k] = expf(m[k] - mx);
        sum += m[k];
    }
    float acc[8];
#pragma unroll
    for (int d = 0; d < 8; ++d) acc[d] = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const int yy = y + k / 3 - 1, xx = x + k % 3 - 1;
        if (yy < 0 || yy >= h || xx < 0 || xx >= w) continue;
        const float wk = m[k] / sum;
        const float* dp = data + (((size_t)n * h + yy) * w + xx) * ldd;
        for (int d = 0; d < dim; ++d) acc[d] = fmaf(wk, __ldg(dp + d), acc[d]);
    }
    const int H = 8 * h, W = 8 * w, oy = 8 * y + i, ox = 8 * x + j;
    const size_t op = ((size_t)n * H + oy) * W + ox;
    if (MODE == 0) {
        for (int d = 0; d < dim; ++d) out[op * ldo + d] = acc[d];
    } else {
        const SE3 T = se3_exp(acc);
        se3_store(out + op * 7, T);
        const float fx = __ldg(intr + n * 4), fy = __ldg(intr + n * 4 + 1), cx = __ldg(intr + n * 4 + 2), cy = __ldg(intr + n * 4 + 3);
        const float dd = __ldg(depth + op);
        const V3 X0 = v3(dd * (((float)ox - cx) / fx), dd * (((float)oy - cy) / fy), dd);
        const V3 X1 = se3_act(T, X0);
        const float Z0 = X0.z + P_EPS, Z1 = X1.z + P_EPS;
        flow[op * 3] = (fx * (X1.x / Z1) + cx) - (fx * (X0.x / Z0) + cx);
        flow[op * 3 + 1] = (fy * (X1.y / Z1) + cy) - (fy * (X0.y / Z0) + cy);
        flow[op * 3 + 2] = 1.f / Z1 - 1.f / Z0;
    }
}

}  // namespace

extern "C" int codd_raft_motion_info(const float* Ts, const float* depth1, const float* depth2_inv, const float* intr,
                                     int n, int h, int w, float* xyz, float* info, int ldi, void* stream) {
    if (!Ts || !depth1 || !depth2_inv || !intr || !xyz || !info || n <= 0 || h <= 1 || w <= 1) return CODD_E_BADARG;
    if (ldi < 9) return CODD_E_SHAPE;
    dim3 grid(codd_ceil_div(w, 128), h, n);
    raft_motion_info_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(Ts, depth1, depth2_inv, intr, n, h, w, xyz, info, ldi);
    CODD_RETURN_IF_CUDA_ERROR();
    return 0;
}

extern "C" int codd_avgpool2_nhwc(const float* in, int ldi, int n, int h, int w, int c, float* out, int ldo, void* stream) {
    if (!in || !out || n <= 0 || h < 2 || w < 2 || c <= 0) return CODD_E_BADARG;
    if (c % 4 || ldi % 4 || ldo % 4 || ldi < c || ldo < c) return CODD_E_SHAPE;
    if (!codd_aligned16(in) || !codd_aligned16(out)) return CODD_E_ALIGN;
    const size_t total = (size_t)n * (h / 2) * (w / 2) * (c / 4);
    avgpool2_nhwc_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, ldi, n, h, w, c, out, ldo);
    CODD_RETURN_IF_CUDA_ERROR();
    return 0;
}

extern "C" int codd_corr_lookup(const float* fmap1, int ld1, const float* const* fmap2_pyramid, const int* ld2, int levels,
                                const float* coords, int ldc, int n, int h, int w, int c, int radius, float* out, int ldo,
                                void* stream) {
    if (!fmap1 || !fmap2_pyramid || !ld2 || !coords || !out || n <= 0 || h <= 0 || w <= 0) return CODD_E_BADARG;
    if (levels < 1 || levels > 4 || radius < 1 || radius > 3 || c % 4 || c > CORR_MAXC || ldc < 2) return CODD_E_SHAPE;
    if (ldo < levels * (2 * radius + 1) * (2 * radius + 1)) return CODD_E_SHAPE;
    CorrP p;
    p.f1 = fmap1; p.ld1 = ld1;
    for (int l = 0; l < 4; ++l) { p.f2[l] = l < levels ? fmap2_pyramid[l] : nullptr; p.ld2[l] = l < levels ? ld2[l] : 0; }
    p.coords = coords; p.ldc = ldc; p.N = n; p.h = h; p.w = w; p.C = c; p.levels = levels; p.radius = radius;
    p.out = out; p.ldo = ldo;
    bool coop = (c % 32 == 0) && (ld1 % 4 == 0) && codd_aligned16(fmap1);
    for (int l = 0; l < levels; ++l) coop = coop && (ld2[l] % 4 == 0) && codd_aligned16(fmap2_pyramid[l]);
    if (coop) corr_lookup_coop_kernel<<<(unsigned)(n * h * w), 64, 0, (cudaStream_t)stream>>>(p);
    else corr_lookup_kernel<<<(unsigned)(n * h * w), 64, 0, (cudaStream_t)stream>>>(p);
    CODD_RETURN_IF_CUDA_ERROR();
    return 0;
}

extern "C" int codd_se3_gn_step(const float* Ts, const float* ae, int lda, const float* target, int ldt,
                                const float* weight, int ldw, const float* depth, const float* intr, int n, int h, int w,
                                int radius, float lm, float ep, float* Ts_out, void* stream) {
    if (!Ts || !ae || !target || !weight || !depth || !intr || !Ts_out || n <= 0 || h <= 0 || w <= 0) return CODD_E_BADARG;
    if (lda < 32 || lda % 4 || ldt < 3 || ldw < 3 || radius < 0) return CODD_E_SHAPE;
    if (radius > GN_MAXR || h > 65535 || n > 65535) return CODD_E_UNSUPPORTED;
    if (!codd_aligned16(ae)) return CODD_E_ALIGN;
    GnP p;
    p.Ts = Ts; p.ae = ae; p.lda = lda; p.target = target; p.ldt = ldt; p.weight = weight; p.ldw = ldw; p.depth = depth;
    p.intr = intr; p.N = n; p.h = h; p.w = w; p.radius = radius; p.lm = lm; p.ep = ep; p.out = Ts_out;
    const size_t smem = 2 * (size_t)GN_STAGE_FLOATS * sizeof(float);
    static CoddDeviceOnce once;
    if (int rc = codd_once_per_device(once, [&] {
            return cudaFuncSetAttribute(se3_gn_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        }))
        return rc;
    se3_gn_step_kernel<<<dim3((unsigned)codd_ceil_div(w, GN_CX), (unsigned)h, (unsigned)n), GN_CX * 32, smem,
                         (cudaStream_t)stream>>>(p);
    CODD_RETURN_IF_CUDA_ERROR();
    return 0;
}

extern "C" int codd_cvx_upsample(const float* data, int ldd, int dim, const float* mask, int ldm, int n, int h, int w,
                                 float* out, int ldo, void* stream) {
    if (!data || !mask || !out || n <= 0 || h <= 0 || w <= 0) return CODD_E_BADARG;
    if (dim < 1 || dim > 8 || ldd < dim || ldm < 576 || ldo < dim) return CODD_E_SHAPE;
    cvx_upsample_kernel<0><<<(unsigned)(n * h * w), 64, 0, (cudaStream_t)stream>>>(data, ldd, dim, mask, ldm, n, h, w, out,
                                                                                   ldo, nullptr, nullptr, nullptr);
    CODD_RETURN_IF_CUDA_ERROR();
    return 0;
}

extern "C" int codd_se3_upsample_flow(const float* Ts, const float* mask, int ldm, const float* depth, const float* intr,
                                      int n, int h, int w, float* twist_ws, float* Ts_up, float* flow, void* stream);

namespace {
__global__ void se3_log_kernel(const float* __restrict__ Ts, size_t count, float* __restrict__ tw) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    float xi[6];
    se3_log(se3_load(Ts + i * 7), xi);
#pragma unroll
    for (int k = 0; k < 6; ++k) tw[i * 6 + k] = xi[k];
}
}  // namespace

// Ts [n,h,w,7] at 1/8 resolution -> twist_ws [n,h,w,6] (workspace) -> Ts_up [n,8h,8w,7] = exp(cvx_upsample(log Ts))
// and flow [n,8h,8w,3] = induced_flow(Ts_up, depth [n,8h,8w], intr [n,4])
extern "C" int codd_se3_upsample_flow(const float* Ts, const float* mask, int ldm, const float* depth, const float* intr,
                                      int n, int h, int w, float* twist_ws, float* Ts_up, float* flow, void* stream) {
    if (!Ts || !mask || !depth || !intr || !twist_ws || !Ts_up || !flow || n <= 0 || h <= 0 || w <= 0) return CODD_E_BADARG;
    if (ldm < 576) return CODD_E_SHAPE;
    const size_t count = (size_t)n * h * w;
    se3_log_kernel<<<(unsigned)((count + 127) / 128), 128, 0, (cudaStream_t)stream>>>(Ts, count, twist_ws);
    cvx_upsample_kernel<1><<<(unsigned)count, 64, 0, (cudaStream_t)stream>>>(twist_ws, 6, 6, mask, ldm, n, h, w, Ts_up, 7,
                                                                            depth, intr, flow);
    CODD_RETURN_IF_CUDA_ERROR();
    return 0;
}
