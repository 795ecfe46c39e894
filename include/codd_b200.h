/*
 * codd_b200 — C ABI of the B200-native (sm_100a) CODD stereo hot path.
 *
 * This header is the drop-in boundary.  The reference has no native code and
 * no FFI: its hot path is a chain of PyTorch op launches inside three registry-built
 * nn.Modules.  Each entry point below replaces one such op sequence; the reference lines it
 * replaces are cited per function (paths relative to the CODD repository root).
 *
 * Conventions
 *   - plain C, no torch / C++ types in any signature; loaded with ctypes.
 *   - every pointer is a DEVICE pointer to fp32 unless stated; the caller owns and allocates
 *     every input, output and workspace buffer (torch's caching allocator on the Python side).
 *   - activations are NHWC ("channels_last"): element (n,y,x,c) of a tensor with pixel stride
 *     `ld` floats lives at  base[((n*H + y)*W + x)*ld + c].  ld >= C lets a tensor be a channel
 *     slice of a wider buffer (how the reference's torch.cat inputs are avoided).
 *     Images enter as NCHW (the reference's layout); cost volumes are [N,D,h,w] planar.
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); nothing synchronises.
 *     The tcgen05 convolutions (codd_conv3x3_tc*, codd_conv3x3x2_tc_ring, codd_conv4x4s2_tc, codd_tile_features_tc) are
 *     launched with programmatic stream serialization: their prologue (barrier set-up, TMEM allocation, weight / bias
 *     staging) may overlap the tail of the previous kernel of the same stream, but they read activations and write
 *     outputs only after that kernel has completed (griddepcontrol.wait) — stream order is what a caller observes.
 *     Consequence for callers: weight / bias buffers passed to these entry points must not be written by the kernel
 *     that immediately precedes the call on the same stream unless that kernel is an ordinary (non-triggering) launch.
 *   - return value: 0 success; < 0 argument error (CODD_E_*); > 0 a cudaError_t.
 *   - no mutable state behind the ABI except the one-time, per-device kernel attribute set-up (an atomic bit per
 *     device ordinal under a mutex, csrc/common.cuh: CoddDeviceOnce), and no environment switches: concurrent calls
 *     from several host threads / on several devices of one process are safe (tests/test_gpu_boundary.py).
 *     Diagnostic hooks (cycle counters, CODD_* environment variables) exist only in `make DIAG=1` builds.
 *   - there is no CPU fallback: on a machine without an sm_100 device every call fails.
 */
#ifndef CODD_B200_H
#define CODD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define CODD_API __attribute__((visibility("default")))
#else
#define CODD_API
#endif

#define CODD_E_BADARG (-1)   /* null pointer / non-positive dimension */
#define CODD_E_SHAPE (-2)    /* dimensions violate a documented constraint */
#define CODD_E_UNSUPPORTED (-3) /* kernel-geometry not instantiated */
#define CODD_E_ALIGN (-4)    /* pointer / stride not 16-byte aligned where required */

/* activation codes for the conv epilogues */
#define CODD_ACT_NONE 0
#define CODD_ACT_LEAKY 1      /* LeakyReLU(0.2): every activation in HITNetMF */
#define CODD_ACT_RELU 2
#define CODD_ACT_RELU_CH0 3   /* ReLU on output channel 0 only (disparity >= 0) */
#define CODD_ACT_SIGMOID 4
#define CODD_ACT_MISH 5
#define CODD_ACT_TANH 6       /* ConvGRU candidate state (raft3d/blocks/gru.py:31) */

CODD_API int codd_version(void);
CODD_API const char* codd_error_string(int code);

/* ------------------------------------------------------------------------------------------
 * Dense convolutions (reference: every nn.Conv2d / ConvTranspose2d on the stereo path —
 * model/stereo/hitnet/backbone.py:8-39,69-88; initialization.py:62-117; propagation.py:89-333).
 * ------------------------------------------------------------------------------------------ */
typedef struct codd_conv_desc {
    int n, h, w;          /* input batch / spatial size */
    int c0, ld0;          /* first input: channels, pixel stride */
    int c1, ld1;          /* optional second input, concatenated after the first (0 = none) */
    int cout, ldo;        /* output channels, output pixel stride */
    int kh, kw;           /* kernel size */
    int sh, sw;           /* stride */
    int ph, pw;           /* zero padding top / left (bottom / right are implied by ho, wo) */
    int dil;              /* dilation (both axes) */
    int ho, wo;           /* output spatial size */
    int act;              /* CODD_ACT_* applied after bias (+ residual) */
    int ldr;              /* residual pixel stride (ignored when residual == NULL) */
    int res_bcast;        /* 1: residual has one channel, broadcast over cout */
    int res_after_act;    /* 1: out = act(conv + bias) + residual  (fusion.py:349 long skip) instead of act(.. + residual) */
} codd_conv_desc;

/* weight is PACKED [kh*kw][c0+c1][cout] (host side: w.permute(2,3,1,0).contiguous()).
 * out = act(conv(cat(in0,in1)) + bias + residual).  residual may be NULL. */
CODD_API int codd_conv2d_nhwc(const codd_conv_desc* d, const float* in0, const float* in1,
                     const float* weight, const float* bias, const float* residual,
                     float* out, void* stream);

/* 3x3 / stride 1 / pad 1 convolution on the tensor cores (tcgen05.mma kind::tf32, 3xTF32 split,
 * accumulators in TMEM, input halo tiles by 4-D TMA) — same math and epilogue as codd_conv2d_nhwc
 * for the layers it covers: cin in {16,24,32} (single source), cout <= 32 (not 16 -> 32).
 * weight_split is [2][9][NP][KC] fp32 (NP = 16|32 >= cout, KC = 16|32 >= cin, zero padded):
 * [0] = tf32-rounded weights, [1] = tf32-rounded remainder  (host: ops.pack_conv_weight_tc).
 * flags: bit0 = split activations by round-to-nearest instead of truncation, bit1 = set the
 * descriptor base-offset field (diagnostic switches; default 0). */
CODD_API int codd_conv3x3_tc(const float* in, int ldi, int cin, int n, int h, int w,
                             const float* weight_split, const float* bias, const float* residual, int ldr,
                             int res_bcast, int cout, int act, float* out, int ldo, int flags, void* stream);

/* Same kernel family with a dilation: dil = 1 (as above) or dil = 3 with pad 3, cin = cout = 32 — the dilated
 * resblocks of tile_update4_1 / tile_update5 (propagation.py:258-280).  One output row per tile; the input rows
 * y-3, y, y+3 arrive as three one-row TMA boxes. */
CODD_API int codd_conv3x3_tc_dil(const float* in, int ldi, int cin, int n, int h, int w,
                                 const float* weight_split, const float* bias, const float* residual, int ldr,
                                 int res_bcast, int cout, int act, float* out, int ldo, int dil, int flags,
                                 void* stream);

/* Rolling-ring formulation of the same 3x3 / stride 1 / pad 1 tensor-core convolution (csrc/conv_tc_ring.cu): a CTA
 * walks a 128-column strip one staged input row at a time and ONE MMA per (kx, k-step) accumulates into the three
 * output rows a staged row contributes to (N = 3 * 2*Cout; accumulators = a ring of TMEM slots).  Same shapes as
 * codd_conv3x3_tc (dilation 1).  weight_ring (host: ops.pack_conv_weight_ring) holds fp16 data: the pass-A block
 * [3 kx][6*NP rows][KC] — rows per ky = [w_hi (NP) | 2^10 * w_lo (NP)], w_hi = fp16(w), w_lo = w - w_hi — followed by the
 * pass-B block [3 kx][6*NP rows][KC] — rows per ky = [0 | w_hi].  The kernel splits the activations the same way
 * (x_hi = fp16(x), fp16(2^10 * x_lo)); both passes run as kind::f16 with fp32 accumulation, the lo half of each TMEM slot
 * collects 2^10 (x_hi w_lo + x_lo w_hi) and the epilogue adds hi + 2^-10 * lo: 3xTF32-class accuracy (11-bit halves) at
 * half the MMAs and operand traffic of a tf32 pass.  |x|, |w| are clamped to the fp16 range (65504). */
CODD_API int codd_conv3x3_tc_ring(const float* in, int ldi, int cin, int n, int h, int w, const float* weight_ring,
                                  const float* bias, const float* residual, int ldr, int res_bcast, int cout, int act,
                                  float* out, int ldo, void* stream);

/* The same with dilation `dil` (pad = dil): dil = 1 is codd_conv3x3_tc_ring; dil = 3 (the dilated ResBlocks of
 * tile_update4_1 / tile_update5, propagation.py:258-280) runs for 32 -> 32 channels and h % 3 == 0 — in y the three row
 * phases of an image are walked as independent dilation-1 sub-images (tensor-map strides), in x the taps are 3 pixels apart
 * in the staged row.  CODD_E_UNSUPPORTED otherwise (callers then use codd_conv3x3_tc_dil). */
CODD_API int codd_conv3x3_tc_ring_dil(const float* in, int ldi, int cin, int n, int h, int w, const float* weight_ring,
                                      const float* bias, const float* residual, int ldr, int res_bcast, int cout, int act,
                                      float* out, int ldo, int dil, void* stream);

/* Two stacked 3x3 / stride 1 / pad 1 convolutions, 16 -> 16 -> 16 channels, in one rolling-ring launch
 * (csrc/conv_tc_ring2.cu):  out = act_b(conv_b(act_a(conv_a(in) + bias_a)) + bias_b [+ residual]).  The intermediate
 * tensor never leaves the SM (TMEM -> fp16 operand tile in shared memory).  Replaces HITUNet conv_merge's 3x3 pair
 * (backbone.py:17-32) and the 16-channel ResBlocks (propagation.py:103-121, residual = in).  Bit-identical to two
 * codd_conv3x3_tc_ring launches.  weight_ring_a / _b as codd_conv3x3_tc_ring (cin = cout = 16); act_a, act_b in
 * {NONE, LEAKY, RELU, RELU_CH0}; ldo (and ldr) multiples of 8 floats, out / residual 32-byte aligned;
 * CODD_E_UNSUPPORTED / CODD_E_SHAPE / CODD_E_ALIGN otherwise (callers then issue the two launches). */
CODD_API int codd_conv3x3x2_tc_ring(const float* in, int ldi, int n, int h, int w, const float* weight_ring_a,
                                    const float* bias_a, int act_a, const float* weight_ring_b, const float* bias_b,
                                    const float* residual, int ldr, int act_b, float* out, int ldo, void* stream);

/* 4x4 / stride 2 / pad 1 convolution (HITUNet conv_down first layer, backbone.py:8-14) on the tensor cores, Cin in
 * {16, 24, 32}, Cout <= 32, even H and W: implicit GEMM over TMA boxes of same-parity columns (csrc/conv_tc_s2.cu), three
 * fp16 products with fp32 accumulation (x = x_hi + x_lo, w = w_hi + w_lo as in codd_conv3x3_tc_ring: fp32-class accuracy).
 * weight_split (host: ops.pack_conv_weight_tc4) holds fp16 data [16 taps][2 NP rows][KC]: per tap NP rows of w_hi = fp16(w)
 * followed by NP rows of fp16(2^10 (w - w_hi)), NP = 16 | 32 >= cout, KC = 16 | 32 >= cin, zero padded.
 * out [n, h/2, w/2, cout] NHWC (ldo).  CODD_E_UNSUPPORTED for other geometries (callers fall back to codd_conv2d_nhwc). */
CODD_API int codd_conv4x4s2_tc(const float* in, int ldi, int cin, int n, int h, int w, const float* weight_split,
                               const float* bias, int cout, int act, float* out, int ldo, void* stream);

/* First backbone layer (backbone.py:35-39,70): 3x3, pad 1, 3 -> cout (<=16) channels,
 * LeakyReLU, reading NCHW images and writing NHWC.  `left` and `right` are two [n,3,h,w]
 * images batches; the output holds 2n samples: left batch first, then right (right may be
 * NULL -> n samples).  weight PACKED [9][3][cout]. */
CODD_API int codd_conv3x3_image(const float* left, const float* right, int n, int h, int w,
                       const float* weight, const float* bias, int cout, float* out, int ldo,
                       void* stream);

/* ConvTranspose2d(k=2, s=2) + LeakyReLU (backbone.py:17-21).  weight PACKED [4][cin][cout]
 * with tap = dy*2+dx (host: w.permute(2,3,0,1).contiguous()).  in [n,h,w,cin], out [n,2h,2w,cout]. */
CODD_API int codd_deconv2x2_nhwc(const float* in, int ldi, int n, int h, int w, int cin,
                        const float* weight, const float* bias, int cout, float* out, int ldo,
                        int act, void* stream);

/* Fused conv_up + conv_merge[0] of HITUNet (backbone.py:17-32,75-88):
 *   out = LeakyReLU(conv1x1(cat(skip, LeakyReLU(deconv2x2(coarse) + b_up))) + b_merge)
 * without materialising the up-sampled tensor (csrc/upmerge.cu).  coarse [n,h/2,w/2,cc] (ldc), skip [n,h,w,cs] (lds),
 * w_up PACKED [4][cc][cu] (as codd_deconv2x2_nhwc), w_merge PACKED [cs+cu][co] (as codd_conv2d_nhwc for a 1x1), out
 * [n,h,w,co] (ldo).  cc and cs multiples of 8; (cu, co) in {(16,16), (24,24)}; CODD_E_UNSUPPORTED otherwise. */
CODD_API int codd_upmerge_nhwc(const float* coarse, int ldc, int cc, const float* skip, int lds, int cs,
                               const float* w_up, const float* b_up, int cu, const float* w_merge,
                               const float* b_merge, int co, int n, int h, int w, float* out, int ldo, void* stream);

/* ------------------------------------------------------------------------------------------
 * K2  tile features (reference: TileInitialization.tile_features, initialization.py:62-95,119-156)
 *   out = LeakyReLU(conv1x1(LeakyReLU(conv4x4(in) + b0)) + b1), 16 channels, written PLANAR.
 * in [n,h_in,w_in,cin] NHWC (ldi); right == 0: 4x4 stride (4,4)  -> out [n,16,h_in/4,w_in/4]
 *                                  right != 0: stride (4,1) over the input zero-padded by 3 columns
 *                                              on the right      -> out [n,16,h_in/4,w_in]
 * w0 PACKED [16 taps][cin][16] (w.permute(2,3,1,0)), w1 torch layout [16][16].
 * ------------------------------------------------------------------------------------------ */
CODD_API int codd_tile_features(const float* in, int ldi, int cin, int n, int h_in, int w_in,
                                const float* w0, const float* b0, const float* w1, const float* b1,
                                int right, float* out, void* stream);

/* The same on the tensor cores (csrc/conv_tc_s2.cu: the 4x4 conv as a tcgen05 implicit GEMM over TMA boxes of column
 * residue classes, three fp16 products as codd_conv4x4s2_tc; LeakyReLU, the 1x1 conv, LeakyReLU and the planar store in the
 * epilogue).  w0_split from ops.pack_conv_weight_tc4 (fp16 [16 taps][32 rows: w_hi | 2^10 w_lo][KC]), cin in {16, 24, 32}. */
CODD_API int codd_tile_features_tc(const float* in, int ldi, int cin, int n, int h_in, int w_in,
                                   const float* w0_split, const float* b0, const float* w1, const float* b1,
                                   int right, float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * K1  L1 cost volume + arg-min tile initialisation
 * (reference: calc_init_disp, initialization.py:18-45; torch.min, :167-171).
 *   cv[n,d,i,j] = sum_c | L[n,c,i,j] - R[n,c,i,4j-d] |   (R := 0 when 4j-d < 0), c sequential.
 * tile_l [n,16,h,w], tile_r [n,16,h,4w]: PLANAR (contiguous NCHW) tile features as written by
 * codd_tile_features; tile_r must be 16-byte aligned (rows are staged with bulk-TMA copies).
 * Any of cv / min_cost / min_disp may be NULL (not produced).
 *   cv        [n,max_disp,h,w] planar   (the reference's init_cv_pyramid entry)
 *   min_cost  [n,h,w]
 *   min_disp  [n,h,w]  arg-min as float (first index on ties), as the reference casts it.
 * Any max_disp >= 1 (the reference uses max_disp // {16,8,4,2,1}); channels fixed at 16
 * (TileInitialization always emits 16).
 * ------------------------------------------------------------------------------------------ */
CODD_API int codd_cost_volume(const float* tile_l, const float* tile_r,
                     int n, int h, int w, int max_disp,
                     float* cv, float* min_cost, float* min_disp, void* stream);

/* The whole tile-initialisation pyramid (initialization.py:158-183, one calc_init_disp + torch.min per level) in ONE
 * launch: `levels` entries (<= 8) of host arrays — device pointers and sizes per level, same meaning as the arguments
 * of codd_cost_volume; cv / min_cost / min_disp must be given for all levels or for none (array or its first entry
 * NULL).  Results are bit-identical to per-level codd_cost_volume calls; the coarse levels, latency-bound when launched
 * alone, fill the tail of the finest level's last wave. */
CODD_API int codd_cost_volume_pyramid(int levels, const float* const* tile_l, const float* const* tile_r, int n,
                                      const int* h, const int* w, const int* max_disp, float* const* cv,
                                      float* const* min_cost, float* const* min_disp, void* stream);

/* Tile descriptor + hypothesis assembly (initialization.py:186-208):
 *   hyp[n,i,j,0:16] = [ min_disp, 0, 0, LeakyReLU(W . cat[min_cost, feat] + b) (13 ch) ]
 * feat [n,h,w,cf] NHWC (ldf > 0) or PLANAR [n,cf,h,w] (ldf == 0, the K2 tile features);
 * weight is the torch layout [13][1+cf]; hyp pixel stride ldh. */
CODD_API int codd_tile_hyp_init(const float* min_cost, const float* min_disp, const float* feat, int ldf,
                       int cf, const float* weight, const float* bias,
                       int n, int h, int w, float* hyp, int ldh, void* stream);

/* ------------------------------------------------------------------------------------------
 * K3  slanted-plane up-sampling (reference: to_plane / upsample, propagation.py:10-32).
 * in [n,h,w,16] -> out [n,h*size,w*size,16]; ch0 = ((d + cx*dx) + cy*dy) * scale, others nearest.
 * ------------------------------------------------------------------------------------------ */
CODD_API int codd_plane_upsample(const float* in, int ldi, int n, int h, int w, int size, float scale,
                        float* out, int ldo, void* stream);

/* ------------------------------------------------------------------------------------------
 * K4  tile warping + local cost volume + `decrease` conv, for the current hypothesis set and
 * (optionally) the up-sampled previous-level set
 * (reference: TileWarping.forward, propagation.py:61-86; warp :35-58; TileUpdate0.forward
 *  :156-160; TileUpdate.forward :206-219).
 * fea_l          [n,H,W,c]  NHWC left features of this level (H = 4h, W = 4w), c in {16,24,32}
 * fea_r_planar   [n,c,H,W]  right features, PLANAR (contiguous NCHW): the warp gathers along x per
 *                channel, which is a contiguous 128-byte request per warp only in this layout
 *                (codd_nhwc_to_nchw produces it from the backbone's NHWC map)
 * cur            [n,h,w,16] current hypotheses
 * prev           [n,h/2,w/2,16] refined hypotheses of the coarser level, or NULL (TileUpdate0)
 * dec_w [16][64] (torch layout), dec_b [16]: the `decrease` 1x1 conv (input = cat[|fea_l|_1
 *                unshuffled (16), local cost volume (48)]), LeakyReLU.
 * aug            [n,h,w,ldaug] receives  [cur(16) | cur_cv(16)]            when prev == NULL
 *                                        [cur | cur_cv | up_prev(16) | prev_cv(16)] otherwise
 *                i.e. exactly the tensor the reference feeds to conv0.
 * raw_cv         optional debug/test output [n,h,w,sets*64]: per set cat[|fea_l|_1 (16), cv48],
 *                the un-reduced `decrease` input (NULL in production).
 * ------------------------------------------------------------------------------------------ */
CODD_API int codd_tile_warp_cost(const float* fea_l, int ldfl, const float* fea_r_planar, int c,
                        const float* cur, int ldc, const float* prev, int ldp,
                        const float* dec_w, const float* dec_b,
                        int n, int h, int w, float* aug, int ldaug, float* raw_cv,
                        void* stream);

/* Same operation with the right features given in NHWC ([n,H,W,c], pixel stride ldfr) — the backbone's own layout, no
 * planar copy: every tap is gathered in place with 128-bit loads (4 channels each). */
CODD_API int codd_tile_warp_cost_nhwc(const float* fea_l, int ldfl, const float* fea_r, int ldfr, int c,
                                      const float* cur, int ldc, const float* prev, int ldp,
                                      const float* dec_w, const float* dec_b,
                                      int n, int h, int w, float* aug, int ldaug, float* raw_cv,
                                      void* stream);

/* ------------------------------------------------------------------------------------------
 * K5  hypothesis selection (reference: TileUpdate.forward, propagation.py:225-248).
 * update [n,h,w,34] = [conf_prev, conf_cur, dprev(16), dcur(16)];  aug as written by K4.
 * refined [n,h,w,16] = conf_cur > conf_prev ? relu0(cur + dcur) : relu0(up_prev + dprev).
 * ------------------------------------------------------------------------------------------ */
CODD_API int codd_hyp_select(const float* update, int ldu, const float* aug, int ldaug,
                    int n, int h, int w, float* refined, int ldr, void* stream);

/* ------------------------------------------------------------------------------------------
 * K13  Fusion input cues and blend (reference: model/fusion/fusion.py:168-318,383-394;
 *      disp_warp utils/warp.py:43-66).
 * ------------------------------------------------------------------------------------------ */
/* 1/ds-resolution cues: corr[n,y,x,0:31] = [feat cross-corr (9), self-corr curr (8), self-corr warp (8),
 * local stereo cost of pred_curr/ds + {-1,0,1} (3), of pred_warp/ds + {-1,0,1} (3)]; disp2 = the two
 * disparities sub-sampled at [ds/2-1::ds] (conv_disp input); extra (optional) = the same pair again.
 * feat_* [n,h,w,32] NHWC; fea_l [n,h,w,cs] NHWC, fea_r_planar [n,cs,h,w]; pred_* [n,h*ds,w*ds]. */
CODD_API int codd_fusion_cues_lowres(const float* feat_curr, int ldc, const float* feat_warp, int ldw,
                                     const float* fea_l, int ldl, const float* fea_r_planar, int cs,
                                     const float* pred_curr, const float* pred_warp, int n, int h, int w, int ds,
                                     float* corr, int ldo, float* disp2, int ld2, float* extra, int ldx,
                                     void* stream);
/* Full-resolution cues [|curr - warp patch| (9), |self| (8+8), flow_warp (3), warp>0 (1), conf_warp (3)]
 * folded into forget_head's first 1x1 conv: out[n,y,x,0:16] = W[16][32] . cues + b.
 * flow_warp / conf_warp [n,3,h,w] planar.  cues_debug (optional) receives the 32 cues [n,32,h,w]. */
CODD_API int codd_fusion_forget_in(const float* pred_curr, const float* pred_warp, const float* flow_warp,
                                   const float* conf_warp, const float* weight, const float* bias, int n, int h,
                                   int w, float* out, int ldo, float* cues_debug, void* stream);
/* reset weight = sigmoid(w[8] . r8 + b) * (warp>0); fusion weight = wf_lowres nearest-upsampled x ds * (warp>0);
 * fused = curr*(1 - wf*wr) + warp*wf*wr.  r8 [n,h,w,8] NHWC; outputs [n,h,w]. */
CODD_API int codd_fusion_blend(const float* pred_curr, const float* pred_warp, const float* r8, int ldr,
                               const float* weight, const float* bias, const float* wf_lowres, int n, int h, int w,
                               int ds, float* fused, float* wf, float* wr, void* stream);

/* ------------------------------------------------------------------------------------------
 * Motion / RAFT3D non-convolutional ops (reference: model/motion/motion.py:82-130,154-207;
 * raft3d/projective_ops.py:11-68; sampler_ops.py:9-28; se3_field.py:150-192; blocks/corr.py:28-62).
 * SE3 elements are 7 floats (tx,ty,tz,qx,qy,qz,qw).  The lietorch / lietorch_extras / pytorch3d
 * semantics are restated from their published algorithms — parity unpinned.
 * ------------------------------------------------------------------------------------------ */
/* One iteration's geometry: xyz[n,h,w,3] = project(Ts * inv_project(depth1)); info[n,h,w,0:9] =
 * clamp([flow(2), 10*log(Ts)(6), 10*(sample(depth2_inv, xy) - 1/Z)(1)], +-50).  intr [n,4] device. */
CODD_API int codd_raft_motion_info(const float* Ts, const float* depth1, const float* depth2_inv, const float* intr,
                                   int n, int h, int w, float* xyz, float* info, int ldi, void* stream);
/* 2x2 average pooling of an NHWC map (correlation pyramid = correlation with pooled fmap2). */
CODD_API int codd_avgpool2_nhwc(const float* in, int ldi, int n, int h, int w, int c, float* out, int ldo, void* stream);
/* Windowed bilinear lookup into the (never materialised) all-pairs correlation pyramid:
 * out[n,y,x, l*(2r+1)^2 + i*(2r+1) + j] = bilinear_{(coords/2^l) + (i-r, j-r)} <f1(y,x)/4, pool_l(f2)/4>. */
CODD_API int codd_corr_lookup(const float* fmap1, int ld1, const float* const* fmap2_pyramid, const int* ld2, int levels,
                              const float* coords, int ldc, int n, int h, int w, int c, int radius, float* out, int ldo,
                              void* stream);
/* Dense Gauss-Newton step: per pixel, affinity-weighted 6x6 normal equations over a (2*radius+1)^2
 * window, damping (lm*H + ep) on the diagonal, Cholesky solve, Ts_out = exp(dx) * Ts. */
CODD_API int codd_se3_gn_step(const float* Ts, const float* ae, int lda, const float* target, int ldt,
                              const float* weight, int ldw, const float* depth, const float* intr, int n, int h, int w,
                              int radius, float lm, float ep, float* Ts_out, void* stream);
/* Convex (softmax-9) x8 up-sampling: data [n,h,w,dim<=8], mask [n,h,w,576] NHWC -> out [n,8h,8w,dim]. */
CODD_API int codd_cvx_upsample(const float* data, int ldd, int dim, const float* mask, int ldm, int n, int h, int w,
                               float* out, int ldo, void* stream);
/* Ts_up [n,8h,8w,7] = exp(cvx_upsample(log Ts)); flow [n,8h,8w,3] = induced_flow(Ts_up, depth, intr).
 * twist_ws: workspace [n,h,w,6]. */
CODD_API int codd_se3_upsample_flow(const float* Ts, const float* mask, int ldm, const float* depth, const float* intr,
                                    int n, int h, int w, float* twist_ws, float* Ts_up, float* flow, void* stream);
/* K9 splat warp: out [n,h,w,c] (NHWC, ldo) = z-sorted top-8 alpha compositing of feat [n,h,w,c] (ldf) moved
 * by Ts; zbuf / disp optional [n,h,w] (disp = bf/(z+1e-5), 0 where > w). workspace: codd_splat_workspace_bytes. */
CODD_API size_t codd_splat_workspace_bytes(int n, int h, int w);
CODD_API int codd_splat_warp(const float* Ts, const float* depth, const float* intr, const float* feat, int ldf, int c,
                             int n, int h, int w, float radius, float bf, float* out, int ldo, float* zbuf, float* disp,
                             void* workspace, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Wide convolutions of the RAFT3D update block as a tcgen05 GEMM (reference: BasicUpdateBlock / ConvGRU,
 * model/motion/raft3d/raft3d.py:43-106, blocks/gru.py:10-35; 128..384-channel layers at 1/8 resolution).
 * ------------------------------------------------------------------------------------------ */
/* Patch matrix of a stride-1 convolution: A[m][tap*c + ch] = in[pixel m shifted by tap] (zero outside), m in NHWC pixel
 * order, row pitch lda >= kh*kw*c; A_lo (optional) = A - tf32(A), the remainder the 3xTF32 GEMM consumes. */
CODD_API int codd_im2col_split(const float* in, int ldi, int n, int h, int w, int c, int kh, int kw, int ph, int pw,
                               int dil, float* A, float* A_lo, int lda, void* stream);
/* out[m][n] = act(sum_k A[m][k] * B[n][k] + bias[n] + residual[m][n]) on tcgen05 (kind::tf32, TMA-fed, TMEM
 * accumulators).  B_hi / B_lo: tf32-rounded weights and remainder, [n][k] row-major (pitch ldb).  A_lo and B_lo both
 * given: 3xTF32 (fp32-class accuracy); both NULL: single-pass TF32. */
CODD_API int codd_gemm_tc(const float* A, const float* A_lo, int lda, const float* B_hi, const float* B_lo, int ldb, int m,
                          int n, int k, const float* bias, const float* residual, int ldr, int act, float* out, int ldo,
                          void* stream);

/* ------------------------------------------------------------------------------------------
 * RAFT3D network glue (reference: model/motion/raft3d/blocks/extractor.py:28-55,124-190 instance
 * norm; raft3d.py:125-137 ResizeConcatConv + mmseg HRModule fuse layers; blocks/gru.py:30-34;
 * raft3d.py:183-186,242; motion.py:154-165,196-197).  The convolutions themselves go through
 * codd_conv2d_nhwc (any Cout: wide layers run as 64-filter chunks).
 * ------------------------------------------------------------------------------------------ */
/* InstanceNorm2d (affine=False, biased variance): y = (x - mean) / sqrt(var + eps) per (sample, channel);
 * relu != 0: y = max(y, 0); residual != NULL: out = max(residual + y, 0) (ResidualBlock tail).
 * workspace: codd_instance_norm_workspace_bytes(n, c) bytes (zeroed by the call).  c <= 256. */
CODD_API size_t codd_instance_norm_workspace_bytes(int n, int c);
CODD_API int codd_instance_norm_nhwc(const float* in, int ldi, int n, int h, int w, int c, float eps, int relu,
                                     const float* residual, int ldr, float* out, int ldo, void* workspace,
                                     size_t ws_bytes, void* stream);
/* out[n,ho,wo,c] = relu?(base + bilinear(in[n,h,w,c]))  (torch F.interpolate semantics; base may be NULL) */
CODD_API int codd_resize_bilinear_nhwc(const float* in, int ldi, int n, int h, int w, int c, const float* base,
                                       int ldb, float* out, int ldo, int ho, int wo, int align_corners, int relu,
                                       void* stream);
/* element-wise over npix pixels x channels:  op 0: act(a)   1: a*b   2: (1-a)*b + a*c (GRU state update)
 *                                            3: act(a+b)    4: 1/a */
CODD_API int codd_eltwise_nhwc(int op, int act, const float* a, int lda, const float* b, int ldb, const float* c,
                               int ldc, float* out, int ldo, size_t npix, int channels, void* stream);
/* depth = clip(bf / (disp + 1e-5), 0, bf)   (motion.py:157-165) */
CODD_API int codd_disp_to_depth(const float* disp, size_t count, float bf, float* depth, void* stream);
/* out[n, i, j, :] = in[n, offset + i*stride, offset + j*stride, :]  (recip != 0: reciprocal), e.g. the 1/8 and
 * 1/4 sampling of depth maps and SE3 fields (raft3d.py:218-219; motion.py:196-197) */
CODD_API int codd_subsample_nhwc(const float* in, int ldi, int n, int h, int w, int c, int offset, int stride,
                                 int recip, float* out, int ldo, void* stream);

/* ------------------------------------------------------------------------------------------
 * N1  input staging (reference: datasets/transforms.py:391-421 Normalize, :147-176 Pad(size_divisor=64,
 * 'reflect'), datasets/formating.py:77-85): uint8 HWC frames [n,h,w,3] (device) -> normalised fp32 NCHW [n,3,hp,wp],
 * out = (v - mean[c]) * (1/std[c]) with the channel order reversed first when to_rgb != 0, reflect-padded (no edge
 * repeat) on the bottom / right.  mean / std are HOST arrays of 3 floats (the config's img_norm_cfg).
 * ------------------------------------------------------------------------------------------ */
CODD_API int codd_stage_images_u8(const uint8_t* img_hwc, int n, int h, int w, const float* mean, const float* std_,
                                  int to_rgb, int hp, int wp, float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * layout helpers at the module boundary
 * ------------------------------------------------------------------------------------------ */
/* NHWC [n,h,w,c] (pixel stride ldi) -> NCHW contiguous */
CODD_API int codd_nhwc_to_nchw(const float* in, int ldi, int n, int h, int w, int c, float* out,
                      void* stream);
/* NCHW contiguous -> NHWC (pixel stride ldo) */
CODD_API int codd_nchw_to_nhwc(const float* in, int n, int c, int h, int w, float* out, int ldo,
                      void* stream);

/* ------------------------------------------------------------------------------------------
 * N3  on-GPU evaluation (reference: model/codd.py:435-517 calc_metric, utils/misc.py:12-36
 * compute_valid_mask, utils/warp.py:69-92 flow_warp(mode="nearest", padding_mode="zeros"), utils/metric.py:9-54).
 * Each call makes one pass over a frame and ADDS into a row of float64 accumulators on the device; the host takes the
 * means once per sequence (no .item() per frame).  All tensors are dense [n,1,h,w] / [n,2,h,w] fp32 except pred /
 * pred_prev, which may be views of the padded network output (sample / row strides in elements).  seg, mask_out,
 * gt_disp2_prev may be NULL.
 *   codd_disp_metrics:     acc[0] += #valid, [1] += sum |pred-gt|, [2] += #(|pred-gt| > 3), [3] += #(gt > 0);
 *                          mask_out[n,h,w] (uint8) = (disp_lo < gt < disp_hi) & (seg > 0)            (codd.py:462-474)
 *   codd_temporal_metrics: flow_prev = ground-truth flow of the previous frame; the current gt / pred / mask are
 *                          sampled at p + flow(p) (nearest) and compared with the previous frame:
 *                          acc[0] += #(mask_prev & mask_curr), [1] += sum abs_err, [2] += sum rel_err, [3] += #(rel > 1),
 *                          [4] += #(abs > 3), [5] += #mask_prev, [6] += #mask_curr, [7] += sum |flow|, [8] += #pixels.
 *                          gt_pos_count: device pointer to acc[3] of codd_disp_metrics for the same frame (0 -> the
 *                          KITTI dummy-disparity mask of codd.py:486-490), or NULL.                  (codd.py:476-517)
 *   codd_sceneflow_metrics: the motion block (codd.py:519-575): 2-D / 3-D flow induced by the dense SE3 field Ts
 *                          [n,h,w,7] = (t, q) at depth clip(BF / pred_prev, 0, BF) (projective_ops.py:11-68) against
 *                          (gt flow, gt disparity change), under compute_valid_mask(gt_prev, flow, disp_change, seg) minus
 *                          flow_occ: acc[0] += #valid, [1] += sum scene-flow EPE, [2] += sum optical-flow EPE,
 *                          [3] += #(sf < 1 px), [4] += #(of < 1 px).  Ts strides in floats (a [:h,:w] crop is fine).
 * ------------------------------------------------------------------------------------------ */
CODD_API int codd_disp_metrics(const float* pred, long long pred_sample_stride, int pred_row_stride, const float* gt,
                               const float* seg, int n, int h, int w, float disp_lo, float disp_hi,
                               unsigned char* mask_out, double* acc, void* stream);
CODD_API int codd_temporal_metrics(const float* flow_prev, const float* gt, const float* pred,
                                   long long pred_sample_stride, int pred_row_stride, const float* seg,
                                   const float* gt_prev, const float* pred_prev, long long pprev_sample_stride,
                                   int pprev_row_stride, const unsigned char* mask_prev, const float* gt_disp2_prev,
                                   const double* gt_pos_count, int n, int h, int w, float disp_lo, float disp_hi, double* acc,
                                   void* stream);
/* utils/misc.py:39-59 compute_gt_disp_change (ground-truth preparation of the motion meters, codd.py:331-340):
 * change[n,1,h,w] = flow_warp(gt_curr, flow_prev, nearest, zeros) - gt_prev, BF_DEFAULT (210) where the sample leaves the
 * image or flow_occ_prev (uint8, may be NULL) marks the pixel occluded; warped (may be NULL) receives the warped map. */
CODD_API int codd_gt_disp_change(const float* flow_prev, const float* gt_curr, const float* gt_prev,
                                 const unsigned char* flow_occ_prev, int n, int h, int w, float* change, float* warped,
                                 void* stream);
CODD_API int codd_sceneflow_metrics(const float* Ts, long long ts_sample_stride, long long ts_row_stride,
                                    const float* pred_prev, long long pprev_sample_stride, int pprev_row_stride,
                                    const float* intrinsics, const float* flow_prev, const float* gt_disp_change,
                                    const float* gt_prev, const float* seg, const unsigned char* flow_occ, int n, int h,
                                    int w, float disp_lo, float disp_hi, double* acc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CODD_B200_H */
