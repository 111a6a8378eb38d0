/*
 * adamvs_b200 — C ABI of the B200-native (sm_100a) Ada-MVS cascade cost-volume hot path.
 *
 * The reference (gpcv-liujin/Ada-MVS, pure Python/PyTorch) has no FFI or plugin layer; its seam
 * for this path is the Python module `models.adamvs` (train_whu.py:100, predict_whu.py:74).  The
 * entry points below are what a binding for that seam calls: each one replaces a group of ATen
 * op sequences in the reference, cited per function (paths relative to the reference root).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to contiguous fp32 unless named `host_*`;
 *   - tensors are row-major with the dimension order written in the comment (x / width fastest);
 *   - nothing allocates, nothing synchronises, nothing touches the legacy default stream: work is
 *     enqueued on `stream` (a cudaStream_t passed as void*) of the CURRENT device;
 *   - re-entrant and free of global mutable state (safe under nn.DataParallel's thread-per-GPU);
 *   - return value: 0 on success, >0 a cudaError_t from a launch, <0 an ADAMVS_E* argument error.
 */
#ifndef ADAMVS_B200_H
#define ADAMVS_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADAMVS_ABI_VERSION 1

#define ADAMVS_EINVAL   (-1)   /* bad shape / null pointer / unsupported size               */
#define ADAMVS_ENOSPACE (-2)   /* workspace smaller than adamvs_*_workspace_floats() says   */

/* depth-hypothesis source (models/module.py:646-663 get_depth_range_samples) */
#define ADAMVS_HYP_PLANES    0 /* hyp_src = depth_values [B,ncol]: d_k = dv[b,0] + k*(dv[b,1]-dv[b,0])/(D-1)      */
#define ADAMVS_HYP_PER_PIXEL 1 /* hyp_src = cur_depth [B,h,w]: lo = cur - *half_range, hi = cur + *half_range,
                                  d_k = lo + k*((hi-lo)/(D-1))                                                     */

/* where the 1e-5 goes in the view-weighted mean */
#define ADAMVS_EPS_NUMERATOR   0 /* (1e-5 + sum_v w_v ref warp_v) / sum_v w_v      DepthNet0, adamvs.py:261-262,290,301 */
#define ADAMVS_EPS_DENOMINATOR 1 /* sum_v w_v ref warp_v / (1e-5 + sum_v w_v)      InferDepthNet0, adamvs.py:496-512    */

/* probability convention of the regression */
#define ADAMVS_PROB_SOFTMAX 0 /* softmax over D, sum p d, max p                     adamvs.py:306-310, module.py:617-625 */
#define ADAMVS_PROB_EXP_EPS 1 /* e=exp(logit) unshifted, sum d e/(sum e+1e-10), max e/(sum e+1e-10)  adamvs.py:516-531   */

/* interval convention of the cascade */
#define ADAMVS_INTERVAL_LAST_COLUMN 0 /* interval = depth_values[0,ncol-1]                  adamvs.py:346 */
#define ADAMVS_INTERVAL_FROM_RANGE  1 /* interval = (dv[0,ncol-1] - dv[0,0]) / num_depth    adamvs.py:569-571 */

int adamvs_abi_version(void);

/* Once per forward, replaces torch.inverse+matmul of models/module.py:539 (544 host-syncing calls per
 * depth map in the reference) and the Python-double scalar arithmetic of adamvs.py:344-347,379-380 /
 * 569-571,603-604.  For each stage s and batch item b and source view v (1..V-1):
 *   relproj[s][b][v-1] = { rot(3x3 row-major), trans(3) } of  proj_s[b,v] * inverse(proj_s[b,0]),
 * evaluated in fp64 and rounded once to fp32;
 *   half_range[s] = (float)( (ndepths[s] / 2.0) * (ratios[s] * interval) ), interval in fp64 from batch item 0.
 * proj_s*: [B,V,4,4]; depth_values: [B,ncol]; relproj: [3,B,V-1,12]; half_range: [3]. */
int adamvs_cascade_prepare(const float* proj_s1, const float* proj_s2, const float* proj_s3,
                           const float* depth_values, int ncol, int B, int V,
                           int interval_mode, int num_depth,
                           const int* host_ndepths, const double* host_ratios,
                           float* relproj, float* half_range, void* stream);

/* K1 — stage-1 pairwise matching score; replaces homo_warping_float + product + mean(dim=1)
 * (module.py:527-568, adamvs.py:269-272 / 472-478).
 *   score[b,v,k,y,x] = (1/C) sum_c feat[b,0,c,y,x] * bilinear(feat[b,v+1,c], u_v(x,y,d_k))
 * feat: [B,V,C,h,w]; relproj: [B,V-1,12] (this stage's slice); score: [B,V-1,D,h,w]. */
int adamvs_pair_score_f32(const float* feat, const float* relproj,
                          int hyp_mode, const float* hyp_src, int hyp_ncol, const float* half_range,
                          float* score, int B, int V, int C, int D, int h, int w, void* stream);

/* Bilinear resize with align_corners=False of N planes (F.interpolate at adamvs.py:296,505):
 * in [N,hi,wi] -> out [N,ho,wo]. */
int adamvs_resize_bilinear_f32(const float* in, float* out, int N, int hi, int wi, int ho, int wo, void* stream);

/* K2 — fused homography warp + product + per-view weighted aggregation; replaces the whole of
 * adamvs.py:285-301 (DepthNet0) / :495-512 (InferDepthNet0) including every per-view warped volume.
 *   volume[b,c,k,y,x] per ADAMVS_EPS_* with w_v = weights[b,v,y,x] (already at h x w).
 * feat: [B,V,C,h,w]; weights: [B,V-1,h,w]; volume: [B,C,D,h,w]. */
int adamvs_fused_volume_f32(const float* feat, const float* relproj,
                            int hyp_mode, const float* hyp_src, int hyp_ncol, const float* half_range,
                            const float* weights, int eps_mode,
                            float* volume, int B, int V, int C, int D, int h, int w, void* stream);

/* K3 (+K4 fused) — recurrent conv-GRU encoder-decoder over the D planes with the softmax regression
 * folded into the last layer; replaces CostRegNetRED.forward / SliceCostRegNetRED.forward
 * (adamvs.py:172-195 / 415-424), ConvGRUCell.forward (module.py:24-52) and the regression
 * (adamvs.py:306-310 / 516-531, module.py:617-625).
 * Weights are device pointers in the reference's own layouts:
 *   conv1_w [8,C,3,3]; gates1_w [16,16,3,3] b[16]; cand1_w [8,16,3,3] b[8]; conv2_w [16,8,3,3];
 *   gates2_w [32,32,3,3] b[32]; cand2_w [16,32,3,3] b[16]; up1_w [16,8,3,3] (ConvTranspose2d) b[8];
 *   out_w [8,1,3,3] (ConvTranspose2d, out_up=1) or [1,8,3,3] (Conv2d, out_up=0), b[1]. */
typedef struct adamvs_regnet_weights {
    const float *conv1_w;
    const float *gates1_w, *gates1_b;
    const float *cand1_w, *cand1_b;
    const float *conv2_w;
    const float *gates2_w, *gates2_b;
    const float *cand2_w, *cand2_b;
    const float *up1_w, *up1_b;
    const float *out_w, *out_b;
} adamvs_regnet_weights;

size_t adamvs_regnet_red_workspace_floats(int B, int C, int D, int h, int w, int out_up);

/* volume: [B,C,D,h,w]; hypotheses as for K1/K2 (defined on the h x w grid; when out_up they are
 * bilinearly upsampled x2, align_corners=False, per module.py:622 / adamvs.py:522);
 * depth, conf: [B,Ho,Wo] with Ho,Wo = 2h,2w if out_up else h,w; logits_out: optional [B,D,Ho,Wo] or NULL.
 * h and w must be even. */
int adamvs_regnet_red_f32(const float* volume, const adamvs_regnet_weights* host_weights,
                          int hyp_mode, const float* hyp_src, int hyp_ncol, const float* half_range,
                          int out_up, int prob_mode,
                          float* workspace, size_t workspace_floats,
                          float* depth, float* conf, float* logits_out,
                          int B, int C, int D, int h, int w, void* stream);

/* Arithmetic of the regulariser's 3x3 convolutions.
 *   FFMA     fp32 FFMA direct convolutions.
 *   TC_FP32  tcgen05.mma kind::tf32 with hi/lo operand splits (A_hi W_hi + A_hi W_lo + A_lo W_hi; the dropped A_lo W_lo
 *            is below the rounding of the fp32 accumulation in TMEM): fp32 accuracy, same tolerance as FFMA.
 *   TC_TF32  activations rounded to tf32 (one pass), weights still split: reported separately with its own tolerance.
 *   AUTO     per layer, whichever of FFMA / TC_FP32 is faster for the plane size (measured on B200):
 *            both are fp32-accurate, so the choice does not change the tolerance.
 * Planes whose width is not a multiple of 8 fall back to FFMA in every mode. */
#define ADAMVS_MATH_FFMA    0
#define ADAMVS_MATH_TC_FP32 1
#define ADAMVS_MATH_TC_TF32 2
#define ADAMVS_MATH_AUTO    3
#define ADAMVS_MATH_DEFAULT ADAMVS_MATH_AUTO

/* adamvs_regnet_red_f32 with an explicit arithmetic mode (adamvs_regnet_red_f32 uses ADAMVS_MATH_DEFAULT). */
int adamvs_regnet_red_ex_f32(const float* volume, const adamvs_regnet_weights* host_weights,
                             int hyp_mode, const float* hyp_src, int hyp_ncol, const float* half_range,
                             int out_up, int prob_mode, int math_mode,
                             float* workspace, size_t workspace_floats,
                             float* depth, float* conf, float* logits_out,
                             int B, int C, int D, int h, int w, void* stream);

/* K4 standalone — softmax over D + expectation + max for a materialised logit volume (the stage-1
 * pair branch: adamvs.py:274-283 / 481-489).  logits: [N,D,h,w]; hypothesis batch index = n / n_per_batch;
 * depth, conf: [N,h,w]. Hypotheses are taken at the logits' own resolution. */
int adamvs_softmax_regress_f32(const float* logits,
                               int hyp_mode, const float* hyp_src, int hyp_ncol, const float* half_range,
                               int prob_mode, float* depth, float* conf,
                               int N, int n_per_batch, int D, int h, int w, void* stream);

/* ---- MS-REDNet (BASELINE config 5; reference models/msrednet.py) ------------------------------------ */

/* K5 - variance cost volume over the reference view and the V-1 warped source views; replaces the per-plane
 * loop of InferDepthNet.forward (msrednet.py:402-420) / the whole-volume form of DepthNet.forward (:214-231):
 *   volume[b,c,k,y,x] = (ref^2 + sum_v warp_v^2)/V - ((ref + sum_v warp_v)/V)^2
 * feat: [B,V,C,h,w]; relproj: [B,V-1,12]; volume: [B,C,D,h,w]. */
int adamvs_variance_volume_f32(const float* feat, const float* relproj,
                               int hyp_mode, const float* hyp_src, int hyp_ncol, const float* half_range,
                               float* volume, int B, int V, int C, int D, int h, int w, void* stream);

/* K6 - four-level GroupNorm conv-GRU recurrent regulariser on -volume with the regression folded in; replaces
 * slice_RED_Regularization.forward / RED_Regularization.forward (msrednet.py:355-372 / 150-181), ConvGRUCell2
 * (module.py:54-106) and the regression (msrednet.py:422-436 / 233-240).  Index l = 0..3 is conv_gru1..conv_gru4
 * (8/16/32/64 hidden channels at h, h/2, h/4, h/8).  Weights in the reference's layouts:
 *   conv{1,2,3}_w [16,C,3,3] [32,16,3,3] [64,32,3,3] (stride 2, no bias);
 *   gate_w[l] [2HC, X+HC, 3,3] gate_b[l] [2HC]; out_w[l] [HC, X+HC, 3,3] out_b[l] [HC]  (X = C,16,32,64);
 *   rnorm/unorm/onorm _w/_b [l] [HC] (GroupNorm(1,HC) affine);
 *   up{3,2,1}_w [64,32,3,3] [32,16,3,3] [16,8,3,3] (ConvTranspose2d stride 2, no bias);
 *   prob_w [8,1,3,3] prob_b [1] (ConvTranspose2d stride 1).
 * volume: [B,C,D,h,w] (h, w multiples of 8); depth, conf: [B,h,w]; logits_out optional [B,D,h,w]. */
typedef struct adamvs_msred_weights {
    const float *conv1_w, *conv2_w, *conv3_w;
    const float *gate_w[4], *gate_b[4], *rnorm_w[4], *rnorm_b[4], *unorm_w[4], *unorm_b[4];
    const float *out_w[4], *out_b[4], *onorm_w[4], *onorm_b[4];
    const float *up3_w, *up2_w, *up1_w;
    const float *prob_w, *prob_b;
} adamvs_msred_weights;

size_t adamvs_regnet_msred_workspace_floats(int B, int C, int D, int h, int w);

int adamvs_regnet_msred_f32(const float* volume, const adamvs_msred_weights* host_weights,
                            int hyp_mode, const float* hyp_src, int hyp_ncol, const float* half_range,
                            int prob_mode, float* workspace, size_t workspace_floats,
                            float* depth, float* conf, float* logits_out,
                            int B, int C, int D, int h, int w, void* stream);

/* ---- 3x3 convolutions of the 2-D networks around the path --------------------------- */

/* out = act(conv3x3(cat(inA, inB), w) + bias), padding 1, stride 1 or 2; replaces Conv2d/ConvBnReLU with eval-mode
 * BatchNorm folded by the caller (module.py:164-199, 254-261) in FeatureNet0 (adamvs.py:57-104) and CostRegNet2D
 * (adamvs.py:198-238), and DeConv2dFuse's cat + conv (module.py:519-523).
 * inA [N,CA,hin,win], inB [N,CB,hin,win] or NULL (CB = 0); wpk: weights re-laid as [CA+CB][9][COUT]; bias [COUT];
 * out [N,COUT,hin/stride,win/stride].  Channel combinations: adamvs_conv3x3_supported(). */
int adamvs_conv3x3_supported(int CA, int CB, int COUT, int stride);
int adamvs_conv3x3_f32(const float* inA, int CA, const float* inB, int CB, const float* wpk, const float* bias,
                       int relu, int stride, float* out, int N, int COUT, int hin, int win, void* stream);

/* FeatureNet0's output heads (adamvs.py:112-149): out = conv1x1(cat(up(ctx_a), up(ctx_c), x), weight), where up() is the
 * bilinear resize (align_corners = False, F.upsample / F.interpolate) of the pooled-context maps to x's size; replaces the
 * two interpolations, the concatenation and the bias-free 1x1 Conv2d out1/out2/out3.
 * x [N,CX,h,w]; ctx_a [N,CCTX,ha,wa]; ctx_c [N,CCTX,hc,wc]; weight [COUT][2*CCTX+CX] (the Conv2d weight, cat order a, c, x);
 * out [N,COUT,h,w].  w % 4 == 0.  Channel combinations: adamvs_context_head_supported(). */
int adamvs_context_head_supported(int CX, int CCTX, int COUT);
int adamvs_context_head_f32(const float* x, const float* ctx_a, const float* ctx_c, const float* weight, float* out,
                            int N, int CX, int CCTX, int COUT, int h, int w, int ha, int wa, int hc, int wc, void* stream);

/* out = act(convT3x3 stride 2, padding 1, output_padding 1 (in; CIN -> COUT) + bias); replaces Deconv2d + BatchNorm (eval,
 * folded by the caller) + ReLU (module.py:202-245) and CostRegNet2D's transposed blocks (adamvs.py:212-225).
 * in [N,CIN,hin,win]; wpk: ConvTranspose2d weight [CIN,COUT,3,3] re-laid as [CIN][9][COUT]; out [N,COUT,2hin,2win]. */
int adamvs_deconv3x3_supported(int CIN, int COUT);
int adamvs_deconv3x3_f32(const float* in, const float* wpk, const float* bias, int relu, float* out,
                         int N, int CIN, int COUT, int hin, int win, void* stream);
/* The same with a skip tensor added AFTER the activation (CostRegNet2D's `e + up(y)`, adamvs.py:233-235):
 * out = act(convT(in) + bias) + residual; residual [N,COUT,2h,2w] or NULL. */
int adamvs_deconv3x3_res_f32(const float* in, const float* wpk, const float* bias, int relu, const float* residual,
                             float* out, int N, int CIN, int COUT, int hin, int win, void* stream);

/* The two pooled-context branches in front of a FeatureNet0 output head (adamvs.py:112-147): AvgPool2d(4) / AvgPool2d(8)
 * -> 1x1 conv -> eval-mode BatchNorm (folded into wa/ba, wc/bc by the caller) -> ReLU, one pass over x.
 * x [N,C,h,w] (h, w multiples of 8); wa, wc [CO,C]; a [N,CO,h/4,w/4]; c [N,CO,h/8,w/8]. */
int adamvs_context_pool_supported(int C, int CO);
int adamvs_context_pool_f32(const float* x, const float* wa, const float* ba, const float* wc, const float* bc,
                            float* a, float* c, int N, int C, int CO, int h, int w, void* stream);


/* ---- Training path: backward kernels, csrc/train.cu ------------------------------------------- */

/* 3x3 convolution with run-time channel counts, the building block of the regulariser's training forward and of every
 * data gradient (reference layers: models/adamvs.py:157-195, models/module.py:24-52):
 *   transposed = 0: y = act(conv2d(x, w [Cout,Cin,3,3], bias, stride 1|2, padding 1))
 *   transposed = 1: y = act(conv_transpose2d(x, w [Cin,Cout,3,3], bias, stride 2, padding 1, output_padding 1))
 * Data gradients: of a stride-1 conv = this op on the output gradient with the flipped, transposed weight; of a stride-2
 * conv = the transposed form with the SAME weight tensor; of the transposed conv = the stride-2 conv form with the same
 * weight tensor.  bias may be NULL.  x: [N,Cin,hin,win]; y: [N,Cout,hout,wout]. */
int adamvs_conv2d_f32(const float* x, const float* w, const float* bias, float* y, int N, int Cin, int Cout,
                      int hin, int win, int stride, int transposed, int relu, void* stream);

/* Weight gradient of conv2d(x, w [Cout,Cin,3,3], stride 1|2, padding 1): gw += sum_{n,pixels} gy * x (gw zeroed by the
 * caller; atomics).  The transposed conv's weight gradient is the same call with x = its output gradient and gy = its
 * input (stride 2), which yields its [Cin,Cout,3,3] layout.  x: [N,Cin,hin,win]; gy: [N,Cout,hin/stride,win/stride]. */
int adamvs_conv2d_wgrad_f32(const float* x, const float* gy, float* gw, int N, int Cin, int Cout,
                            int hin, int win, int stride, void* stream);

/* Backward of adamvs_pair_score_f32 (K1): g_feat [B,V,C,h,w] += d score / d feat (zeroed by the caller).  The sampling
 * grid carries no gradient (models/module.py:538). */
int adamvs_pair_score_bwd_f32(const float* feat, const float* relproj,
                              int hyp_mode, const float* hyp_src, int hyp_ncol, const float* half_range,
                              const float* g_score, float* g_feat, int B, int V, int C, int D, int h, int w, void* stream);

/* Backward of adamvs_fused_volume_f32 (K2): g_feat [B,V,C,h,w] and g_weights [B,V-1,h,w], both += (zeroed by the caller). */
int adamvs_fused_volume_bwd_f32(const float* feat, const float* relproj,
                                int hyp_mode, const float* hyp_src, int hyp_ncol, const float* half_range,
                                const float* weights, int eps_mode, const float* g_volume,
                                float* g_feat, float* g_weights, int B, int V, int C, int D, int h, int w, void* stream);

/* K4 in its training form: softmax over D, expectation over a materialised hypothesis tensor, max probability
 * (adamvs.py:306-310, module.py:617-625), and its backward (g_depth / g_conf may be NULL = zero; g_hyp may be NULL).
 * logits, hyp, g_logits, g_hyp: [N,D,h,w]; depth, conf, g_depth, g_conf: [N,h,w]. */
int adamvs_softmax_expect_f32(const float* logits, const float* hyp, float* depth, float* conf,
                              int N, int D, int h, int w, void* stream);
int adamvs_softmax_expect_bwd_f32(const float* logits, const float* hyp, const float* depth,
                                  const float* g_depth, const float* g_conf, float* g_logits, float* g_hyp,
                                  int N, int D, int h, int w, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ADAMVS_B200_H */
