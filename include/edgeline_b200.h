/*
 * edgeline_b200.h -- C ABI of libedgeline_b200.so: hand-written sm_100a kernels for the
 * custom-operator path of EdgeLine-YOLO (OneWalkman/EDGE-YOLO, a fork of ultralytics 8.3.63).
 *
 * The reference is pure Python; its "plugin API" for this path is a set of nn.Module classes
 * and functions looked up by name.  Each entry point below replaces the
 * body of one of them; the citation names the reference code it stands in for, paths relative
 * to /root/reference/ultralytics/.  A reference-side binding is a ctypes call, see
 * INTEGRATION.md.
 *
 * Conventions
 *   - every pointer named `x`, `y`, `out`, ... is a DEVICE pointer unless the comment says
 *     "host"; stride arrays, level tables and pointer tables are HOST arrays read during the call;
 *   - strides are in ELEMENTS, logical order (n, c, h, w); any memory format (NCHW, NHWC,
 *     channel-slice views) is expressed through them; 16-byte aligned, channel-contiguous views
 *     take the vectorised kernels, everything else a generic strided kernel;
 *   - `dtype` is an el_dtype and is the storage type of activations; arithmetic is fp32;
 *   - `stream` is a cudaStream_t passed as void*; calls only enqueue work: they never
 *     allocate, never synchronise, and may be captured in a CUDA graph;
 *   - return value: EL_OK (0) or an el_status error; there is no CPU fallback.
 */
#ifndef EDGELINE_B200_H
#define EDGELINE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum { EL_F32 = 0, EL_F16 = 1, EL_BF16 = 2 } el_dtype;

typedef enum {
    EL_OK = 0,
    EL_ERR_ARG = 1,         /* null pointer, non-positive size, bad enum */
    EL_ERR_UNSUPPORTED = 2, /* shape outside the kernel's compiled envelope */
    EL_ERR_WORKSPACE = 3,   /* workspace smaller than el_*_workspace_bytes says */
    EL_ERR_CUDA = 4         /* launch failed; see el_last_cuda_error() */
} el_status;

const char* el_version(void);
const char* el_status_string(int status);
/* cudaError_t of the most recent failed launch on this thread (0 if none). */
int el_last_cuda_error(void);
/* Number of kernels this library has enqueued so far in the process (measurement aid: bench.py
 * reports the per-step delta as `gpu_launches`). */
unsigned long long el_launch_count(void);
/* ---- a1. Haar analysis: _PywtDWT2D.forward, nn/modules/block.py:3619-3642 -------------------
 * x (B,C,H,W) -> bands (4,B,C,H/2,W/2) in the order LL,LH,HL,HH; xs = {sn,sc,sh,sw},
 * bs = {sband,sn,sc,sh,sw}.  Odd trailing row/column is dropped (floor), taps are
 * float32(2^-1/2)^2 = 0.49999997 (block.py:3597-3609). */
int el_dwt_haar_fwd(const void* x, const int64_t xs[4], void* bands, const int64_t bs[5],
                    int B, int C, int H, int W, int dtype, void* stream);
/* Adjoint (== the exact inverse up to the tap rounding): gradient of the above, and the Haar
 * synthesis that inverse_2d_wavelet_transform (nn/modules/conv.py:438-443) performs.
 * gbands (4,B,C,H/2,W/2) -> gx (B,C,H,W), every element of gx written. */
int el_dwt_haar_bwd(const void* gbands, const int64_t gs[5], void* gx, const int64_t gxs[4],
                    int B, int C, int H, int W, int dtype, void* stream);

/* ---- a2. sub-band merge: _WaveletEnhancer.forward, block.py:3696-3708 -----------------------
 * out (B,3c,H,W) = cat[b, up(LLp)*w0, up(LHp)*w1, up(HLp)*w2, up(HHp)*w3], up = bilinear
 * align_corners=False (h,w)->(H,W), w = softplus(alpha)/(sum+1e-6) computed on the device from
 * `alpha` (4 floats).  band[i] (B,c/2,h,w) with strides band_s[4*i..4*i+3].
 * b == NULL (inference engine, NHWC 16-byte-vector views only): bands-only output (B,2c,H,W) without
 * the pass-through copy of b -- the consumer (`fuse`, el_pwconv_fwd) reads b in place. */
int el_wave_merge_fwd(const void* b, const int64_t bs[4], const void* const band[4],
                      const int64_t band_s[16], const float* alpha, void* out,
                      const int64_t os[4], int B, int c, int H, int W, int h, int w, int dtype,
                      void* stream);
/* Gradient of the merge: gout (B,3c,H,W) -> gb (B,c,H,W) (= gout[:, :c], written, not
 * accumulated), gband[i] (B,c/2,h,w) (adjoint of the scaled upsample) and galpha_w (4 floats,
 * d loss / d w[i] = <gout_i, up(band_i)>, accumulated with atomics -- zero it first). */
int el_wave_merge_bwd(const void* gout, const int64_t gos[4], const void* const band[4],
                      const int64_t band_s[16], const float* alpha, void* gb,
                      const int64_t gbs[4], void* const gband[4], const int64_t gband_s[16],
                      float* galpha_w, int B, int c, int H, int W, int h, int w, int dtype,
                      void* stream);
/* Gated residual, block.py:3710: out = b + tanh(gamma) * y; gamma is one device float.
 * out may alias b.  out2 (optional, NHWC engine path only) receives a second copy of the result,
 * e.g. the block's slice of the concat buffer. */
int el_gated_residual_fwd(const void* b, const int64_t bs[4], const void* y, const int64_t ys[4],
                          const float* gamma, void* out, const int64_t os[4], void* out2,
                          const int64_t os2[4], int B, int C, int H, int W, int dtype, void* stream);

/* Backward of the gated residual: gy = tanh(gamma) * g, *ggamma += (1 - tanh^2(gamma)) * <g, y> (zero it first);
 * the gradient with respect to b is g itself. */
int el_gated_residual_bwd(const void* g, const int64_t gs[4], const void* y, const int64_t ys[4],
                          const float* gamma, void* gy, const int64_t os[4], float* ggamma, int B, int C,
                          int H, int W, int dtype, void* stream);

/* ---- a4. linear-attention core: LinearAttention.forward, block.py:3364-3372 -----------------
 * qkv (B,3C,N) with strides qs = {sb, sc, sn} (channel = t*C + head*64 + j, t in q,k,v);
 * y (B,C,N), ys likewise.  softmax_d(K), softmax_N(Q), ctx = K^T V, y = Q ctx.  head_dim is 64
 * (heads = C/64, block.py:3474).  EL_BF16/EL_F16 inputs run the contraction on tcgen05 tensor
 * cores with the accumulator in TMEM; EL_F32 runs an fp32 CUDA-core kernel (1e-5 contract). */
int el_linattn_fwd(const void* qkv, const int64_t qs[3], void* y, const int64_t ys[3], int B,
                   int heads, int N, int dtype, void* stream);
/* Backward: gy (B,C,N) -> gqkv (B,3C,N) through both softmaxes and both contractions (fp32 arithmetic). */
int el_linattn_bwd(const void* qkv, const int64_t qs[3], const void* gy, const int64_t gs[3], void* gqkv,
                   const int64_t os[3], int B, int heads, int N, int dtype, void* stream);

/* ---- a6+a7. GFLv2 x UniHead decode: head.py:227-243 + 301-345, block.py:87-90, tal.py:333-357 --
 * Per level l < nl: box[l] (B,64,Hl,Wl) DFL logits, cls[l] (B,nc,Hl,Wl) class logits (host
 * tables of device pointers, strides box_s/cls_s[4*l..]), hw[2*l..] = {Hl,Wl}, stride[l].
 * DGQP head weights (fp32, device): w1[l] (64,20), b1[l] (64), w2[l] (64), b2[l] (1).
 * box_bias / cls_bias: NULL, or per-level fp32 device vectors (64) / (nc) added to the logits first
 * (the engine strips the bias of the towers' last 1x1 convs and folds it in here); entries may be NULL.
 * Writes y (B,4+nc,A) fp32 contiguous: rows 0-3 = (cx,cy,w,h)*stride, rows 4.. =
 * sigmoid(cls)*clamp(q,1e-6,1-1e-6); optional q_out (B,A) fp32 (may be NULL). */
int el_gfl_decode_fwd(int nl, const void* const* box, const int64_t* box_s, const void* const* cls,
                      const int64_t* cls_s, const int32_t* hw, const float* stride,
                      const float* const* w1, const float* const* b1, const float* const* w2,
                      const float* const* b2, const float* const* box_bias,
                      const float* const* cls_bias, float* y, float* q_out, int B, int nc, int dtype,
                      void* stream);

/* ---- a6+a7+a9 fused (engine path): decode -> NMS candidates -> sort -> sweep, no dense score tensor --
 * Same inputs as el_gfl_decode_fwd plus the NMS arguments of el_nms_batched.  Persistent CTAs stage
 * 64-anchor tiles of the NHWC head outputs with bulk TMA copies, decode them in shared memory and write
 * only the xywh boxes (B,A,4) and the candidate keys; results are bit-identical to
 * el_gfl_decode_fwd + el_nms_batched.  Returns EL_ERR_UNSUPPORTED when the maps are not dense NHWC with
 * 16-byte-aligned tile rows (callers then use the two-call path).
 * `stages` selects which part of the chain this call enqueues on the caller's workspace (bit 0 decode + candidate emit,
 * bit 1 max_nms cut + sort, bit 2 sweep; 7 = everything): the inference engine captures stage 1 and stages 6 as two CUDA
 * graphs so that the NMS tail of batch i overlaps the forward of batch i+1, and bench.py times the stages separately.
 * The argument is per call -- the library keeps no mutable state between calls (re-entrant across threads / streams). */
int el_gfl_detect_workspace_bytes(int B, int nc, int A, int multi_label, int max_nms, size_t* bytes);
int el_gfl_detect_fwd(int nl, const void* const* box, const int64_t* box_s, const void* const* cls,
                      const int64_t* cls_s, const int32_t* hw, const float* stride,
                      const float* const* w1, const float* const* b1, const float* const* w2,
                      const float* const* b2, const float* const* box_bias,
                      const float* const* cls_bias, int B, int nc, int dtype, float conf_thres,
                      double iou_thres, int multi_label, int agnostic, const int32_t* class_keep,
                      int max_det, int max_nms, float max_wh, int stages, void* workspace,
                      size_t workspace_bytes, float* out, int32_t* out_count, int64_t* out_index, void* stream);

/* ---- a9. batched NMS: non_max_suppression, utils/ops.py:167-316 (detection, nm=0) ------------
 * pred (B,4+nc,A) fp32 contiguous, xywh + scores (the tensor el_gfl_decode_fwd writes).
 * Candidate build (conf filter, multi-label expansion or first-max class, optional class
 * filter), top-max_nms cut, class-offset (cls*max_wh), stable descending sort and greedy IoU
 * suppression (torchvision.ops.nms semantics: strict `>`, IoU in fp32 compared with the
 * threshold as a double, no epsilon) all run on the device with no host synchronisation.
 * out (B,max_det,6) rows [x1,y1,x2,y2,score,cls]; out_count (B); out_index (B,max_det) int64 =
 * anchor*nc+cls of each kept row (may be NULL).  class_keep: nc int32 flags or NULL. */
int el_nms_workspace_bytes(int B, int nc, int A, int multi_label, int max_nms, size_t* bytes);
int el_nms_batched(const float* pred, int B, int nc, int A, float conf_thres, double iou_thres,
                   int multi_label, int agnostic, const int32_t* class_keep, int max_det,
                   int max_nms, float max_wh, void* workspace, size_t workspace_bytes, float* out,
                   int32_t* out_count, int64_t* out_index, void* stream);
/* torchvision.ops.nms(boxes, scores, iou) itself (call site utils/ops.py:296): boxes (n,4) xyxy
 * fp32, scores (n) fp32 -> keep (n) int64 indices in descending-score order, keep_count (1). */
int el_nms_boxes_workspace_bytes(int n, size_t* bytes);
int el_nms_boxes(const float* boxes, const float* scores, int n, double iou_thres, void* workspace,
                 size_t workspace_bytes, int64_t* keep, int32_t* keep_count, void* stream);

/* ---- a10. Quality Focal Loss: quality_focal_loss, utils/loss.py:22-70 ------------------------
 * pred (n) logits in `dtype`, target (n) fp32.  fwd: loss (n) fp32 elementwise (may be NULL)
 * and/or loss_sum (1 float, deterministic two-stage reduction; may be NULL; needs `partials`
 * of el_qfl_partials() floats).  bwd: gpred = g * dloss/dpred with g = gout[i] (elementwise,
 * fp32) or *gscalar (one device float) -- exactly one of the two non-NULL. */
int el_qfl_partials(int64_t n);
int el_qfl_fwd(const void* pred, const float* target, int64_t n, float beta, int dtype, float* loss,
               float* loss_sum, float* partials, void* stream);
int el_qfl_bwd(const void* pred, const float* target, int64_t n, float beta, int dtype,
               const float* gout, const float* gscalar, void* gpred, void* stream);

/* ---- a11. Distribution Focal Loss: DFLoss.__call__, utils/loss.py:209-224 -------------------
 * pred (rows*4, 16) logits in `dtype` contiguous, target (rows,4) fp32 (NOT modified; the
 * reference clamps in place, the clamp is applied on the fly here).  fwd: loss (rows) fp32 =
 * mean over the 4 sides.  bwd: gpred (rows*4,16) = gout[row] * (softmax - wl*d_tl - wr*d_tr)/4. */
int el_dfl_fwd(const void* pred, const float* target, int64_t rows, int dtype, float* loss,
               void* stream);
int el_dfl_bwd(const void* pred, const float* target, int64_t rows, int dtype, const float* gout,
               void* gpred, void* stream);
/* Per-side form: distribution_focal_loss(reduction="none"), utils/loss.py:88-137.  pred (sides, 16), target (sides) fp32 -> loss (sides) fp32;
 * bwd: gpred (sides, 16) = gout[side] * (softmax - wl*d_tl - wr*d_tr).  Same kernel as above without the mean over the 4 sides of a row. */
int el_dfl_side_fwd(const void* pred, const float* target, int64_t sides, int dtype, float* loss,
                    void* stream);
int el_dfl_side_bwd(const void* pred, const float* target, int64_t sides, int dtype, const float* gout,
                    void* gpred, void* stream);

/* ---- ingest: uint8 HWC images -> normalised NHWC activations (engine/predictor.py:117-135) ----
 * src (B,H,W,3) uint8 -> dst (B,3,H,W) logical with strides ds, value/255 in `dtype`. */
int el_ingest_u8(const uint8_t* src, void* dst, const int64_t ds[4], int B, int H, int W, int dtype,
                 void* stream);
/* el_stem_conv_u8: the same preprocess fused with layer 0 of the yaml (cfg/models/11/yolo11-test.yaml:21,
 * Conv(3, C0, k=3, s=2, p=1) + BatchNorm + SiLU, nn/modules/conv.py:41-60): src (B,H,W,3) uint8 ->
 * dst (B,C0,H/2,W/2) logical with strides ds (channel-contiguous).  w (C0,3,3,3) fp32 = BN-folded weights
 * already divided by 255, bias (C0) fp32 = folded BN bias.  C0 in {16,32,64}, H even, W a multiple of 4. */
int el_stem_conv_u8(const uint8_t* src, const float* w, const float* bias, void* dst,
                    const int64_t ds[4], int B, int C0, int H, int W, int dtype, void* stream);

/* ---- conv epilogues of the inference engine (the convolutions themselves stay on cuDNN) ---------
 * el_bias_act_fwd: out = act(x + bias[c]) (+ residual): the BatchNorm-folded bias and SiLU of Conv /
 * DSConv (nn/modules/conv.py:41-60, 87-104) and the shortcut add of DSBottleneck (block.py:1500-1503)
 * in one pass.  act: 0 none, 1 SiLU, 2 ReLU; bias (C) fp32 or NULL; residual (same shape) or NULL;
 * out may alias x and may be a channel slice of a wider (concat) buffer.  With out2 != NULL channels
 * [0, split) go to out and [split, C) to out2 (a C2f block's pass-through half into the concat buffer,
 * the processed half into a dense tensor).
 * el_upsample2x_cat_fwd: out (B,C1+C2,H,W) = cat[nearest2x(x (B,C1,H/2,W/2)), skip (B,C2,H,W)], the
 * nn.Upsample + Concat pairs of the neck (cfg/models/11/yolo11-test.yaml:34-39); NHWC views only. */
int el_bias_act_fwd(const void* x, const int64_t xs[4], const float* bias, const void* residual,
                    const int64_t rs[4], void* out, const int64_t os[4], void* out2,
                    const int64_t os2[4], int split, int B, int C, int H, int W, int act, int dtype,
                    void* stream);
int el_upsample2x_cat_fwd(const void* x, const int64_t xs[4], const void* skip, const int64_t ss[4],
                          void* out, const int64_t os[4], int B, int C1, int C2, int H, int W,
                          int dtype, void* stream);
/* el_dwconv_fwd: depthwise k x k convolution, stride 1, padding k/2 (k in {3,5,7}), NHWC views, optional fused
 * bias + activation: DSConv.dw (nn/modules/conv.py:87-104, no epilogue) and DWConv + folded BatchNorm + SiLU
 * (conv.py:124-130; Detect cls tower head.py:66-71).  w: fp32 (k*k, C) tap-major; bias fp32 (C) or NULL;
 * act as el_bias_act_fwd.  C must be a multiple of the 16-byte channel vector. */
int el_dwconv_fwd(const void* x, const int64_t xs[4], const float* w, const float* bias, void* out,
                  const int64_t os[4], int B, int C, int H, int W, int k, int act, int dtype, void* stream);
/* el_dsconv3_fwd: depthwise 3 x 3 (stride 1, padding 1; optional bias + activation) -> pointwise 1 x 1 + bias + activation in ONE kernel:
 * DSConv.forward with k = 3 (nn/modules/conv.py:100-104: dw -> pw -> BatchNorm -> SiLU; cv1 of DSBottleneck, nn/modules/block.py:1494-1503) and
 * the DWConv(x, x, 3) -> Conv(x, c, 1) stages of the class towers (nn/modules/head.py:66-71).  The depthwise result is rounded to the
 * activation type (as el_dwconv_fwd would store it) and written straight into the swizzled A tile of the tcgen05 GEMM; it never reaches HBM.
 * x / out NHWC 16-bit views (channels contiguous, other strides in 8s); dw_w fp32 (9, C) tap-major, dw_bias fp32 (C) or NULL, dw_act / act
 * as el_bias_act_fwd; wpk = the el_pwconv_fwd weight packing for ONE source of C channels with a single output-channel tile of
 * ceil16(N) rows; bias fp32 (N) or NULL.  el_dsconv3_ok: C = 16 / 32 or >= 64 (multiple of 8), N a multiple of 8 up to 256, weights resident. */
int el_dsconv3_ok(int C, int N);
int el_dsconv3_fwd(const void* x, const int64_t xs[4], int C, const float* dw_w, const float* dw_bias, int dw_act, const void* wpk,
                   const float* bias, int act, void* out, const int64_t os[4], int B, int H, int W, int N, int dtype, void* stream);
/* el_pwconv_fwd: pointwise (1x1) convolution + folded-BatchNorm bias + activation (+ shortcut) as one streaming
 * tcgen05 GEMM over NHWC pixels (Conv(k=1).forward_fuse nn/modules/conv.py:58-60; DSConv.pw + bn + act conv.py:100-104;
 * LinearAttention.qkv / proj block.py:3353-3373).  out[p, n] = res_scale * act(sum_k X[p,k] W[n,k] + bias[n]) (+ res[p,n]);
 * res_scale = tanh(gamma) with res = b is the gated residual of _WaveletEnhancer (block.py:3708-3710), else 1.
 * up_H > 0 switches `res` to a PRE-activation addend at half resolution: out[b,y,x] = act(conv + bias + res[b, y/2, x/2]) with
 * (up_H, up_W) the output map size -- nn.Upsample(2, nearest) + Concat + 1x1 conv of the neck (yolo11-test.yaml:34-39) without
 * ever materialising the upsampled / concatenated tensor: res = W[:, :C1] . x_low computed at low resolution.
 * The K dimension is the concatenation of `nsrc` (<= 4) source tensors (src[i]: 16-bit, src_c[i] channels, multiple
 * of 8, pixel pitch src_pitch[i] elements) -- the torch.cat of the C2f-style blocks (block.py:3783-3788) is never
 * materialised.  Channels >= split go to out2 when out2 != NULL (chunk(2,1) of cv1).  M = B*H*W pixels; every view
 * must be pixel-linear (address = base + p * pitch).  wpk: weights packed by the host into UMMA tiles
 * [n_tiles][k-groups (padded per 64-channel chunk to even)][n_tile = el_pwconv_tile(N, weight row bytes, M)][8]; see ops.pack_pw_weight.
 * bf16 / fp16 only (EL_ERR_UNSUPPORTED otherwise: the fp32 API path keeps cuDNN). */
int el_pwconv_tile(int N, int w_row_bytes, int64_t M);
int el_pwconv_fwd(int nsrc, const void* const src[], const int64_t src_pitch[], const int32_t src_c[],
                  const void* wpk, const float* bias, const void* res, int64_t res_pitch, float res_scale,
                  int up_H, int up_W, void* out, int64_t out_pitch, void* out2, int64_t out2_pitch, int split,
                  int64_t M, int N, int act, int dtype, void* stream);
/* el_conv3x3_fwd: dense 3x3 convolution, padding 1, stride 1 or 2, + folded-BatchNorm bias + activation as an implicit GEMM on
 * the el_pwconv_fwd kernel (Conv(k=3).forward_fuse nn/modules/conv.py:58-60: the stride-2 downsampling convs of the yaml,
 * _WaveletEnhancer.f_h block.py:3657-3679, the Detect box tower head.py:59-63).  x (B,C,H,W) / out (B,N,Ho,Wo) NHWC views with
 * element strides xs / os = {n,c,h,w}; wpk = ops.pack_pw_weight of the (N, 9*C) matrix in (ky, kx, c) order with nine
 * "sources" of C channels and n_tile = el_conv3x3_tile(N, C, B*Ho*Wo) (C <= 32: weight tiles resident in shared memory;
 * wider: streamed through the ring with the activation boxes).  bf16 / fp16 only. */
int el_conv3x3_tile(int N, int C, int64_t M);
int el_conv3x3_fwd(const void* x, const int64_t xs[4], int C, const void* wpk, const float* bias, void* out,
                   const int64_t os[4], int B, int H, int W, int N, int stride, int act, int dtype, void* stream);
/* Wide dense 3x3 conv (stride 1, padding 1, C_in a multiple of 64): Conv.forward_fuse, nn/modules/conv.py:58-60, at the sites where the
 * tap-shifted form above loses to cuDNN (f_h of _WaveletEnhancer block.py:3668-3673 from c = 64, the Detect box towers head.py:59-63).
 * One haloed (10 x 16 pixel) TMA box per 64-channel K chunk serves all nine taps through row-shifted UMMA descriptors (conv3x3_halo.cu).
 * wpk: the (N, 9*C) matrix in (ky, kx, c) order packed as [tap][chunk] K-major SW128 tiles of n_pad = ceil16(N) rows (ops.pack_conv3x3_halo_weight);
 * x / out NHWC 16-bit with element strides {n, c = 1, h, w}.  el_conv3x3_halo_ok: 1 if the nine weight tiles of a (C, N) site fit shared memory. */
int el_conv3x3_halo_ok(int C, int N);
int el_conv3x3_halo_fwd(const void* x, const int64_t xs[4], int C, const void* wpk, const float* bias, void* out,
                        const int64_t os[4], int B, int H, int W, int N, int act, int dtype, void* stream);
/* el_conv3x3_mma_fwd: dense 3 x 3 convolution (stride 1 / 2, padding 1) + bias + activation for NARROW channel counts (C_in = 16 / 32,
 * N a multiple of 8 up to 64; el_conv3x3_mma_ok): the shared high-band conv f_h of _WaveletEnhancer (nn/modules/block.py:3668-3673,
 * `Conv.forward_fuse` nn/modules/conv.py:58-60) on the large early maps and the stride-2 layer 1 of the yaml (cfg/models/11/yolo11-test.yaml:22).
 * One haloed 18 x 18 (stride 2: 33 x 33) tile per 16 x 16 output pixels staged by
 * cp.async, mma.sync.m16n8k16 from ldmatrix fragments of that tile.  x / out NHWC 16-bit views (channels contiguous, x strides in 8s,
 * out strides even); w is the plain fp32 (N, C, 3, 3) weight (fragments are built in the kernel), bias fp32 (N) or NULL;
 * act as el_bias_act_fwd. */
int el_conv3x3_mma_ok(int C, int N, int stride);
int el_conv3x3_mma_fwd(const void* x, const int64_t xs[4], int C, const float* w, const float* bias, void* out, const int64_t os[4], int B,
                       int H, int W, int N, int stride, int act, int dtype, void* stream);
/* el_sppf_pool_fwd: out (B,4C,H,W) = cat[x, m(x), m(m(x)), m(m(m(x)))], m = MaxPool2d(5,1,2): the pooling
 * pyramid of SPPF (nn/modules/block.py:204-223) as separable 5/9/13 window maxima in shared memory.
 * NHWC views, H*W*64 B of shared memory (maps up to ~56x56), else EL_ERR_UNSUPPORTED. */
int el_sppf_pool_fwd(const void* x, const int64_t xs[4], void* out, const int64_t os[4], int B, int C,
                     int H, int W, int dtype, void* stream);

/* ---- 8f-4. validator metrics: box_iou (utils/metrics.py:55-71) and DetectionValidator.match_predictions, non-scipy branch
 * (engine/validator.py:222-262) --------------------------------------------------------------------------------------------
 * el_box_iou: box1 (N rows of 4 floats, row pitch stride1 elements, xyxy), box2 (M rows) -> out (N, M) dense fp32,
 *   inter / (area1 + area2 - inter + eps) in fp32, evaluated left to right like the reference.
 * el_match_predictions: iou (L labels x D detections, dense row-major, as the validator passes it), pred_cls (D), true_cls (L),
 *   iouv (T <= 16 thresholds, device) -> correct (D, T) bytes 0/1.  One image per call; no host synchronisation. */
int el_box_iou(const float* box1, int64_t stride1, const float* box2, int64_t stride2, float* out, int N, int M, float eps, void* stream);
int el_match_predictions(const float* iou, const float* pred_cls, const float* true_cls, const float* iouv, int L, int D, int T,
                         uint8_t* correct, void* stream);
/* el_ap_per_class: ap_per_class + compute_ap, utils/metrics.py:505-623, up to the per-class curves (float64 like numpy).
 *   tp (N, T) bytes 0/1 row-major (T <= 16 IoU thresholds), conf (N), pred_cls (N) class ids as floats -- the concatenated stats of the
 *   validator (models/yolo/detect/val.py:170-195), device pointers; classes (nc) = np.unique(target_cls) ascending as floats and
 *   n_labels (nc) its counts, device pointers.  Outputs (device, float64): ap (nc, T), p_curve / r_curve / prec_values (nc, 1000)
 *   and n_pred (nc) int32 predictions per class; rows of classes without predictions or labels stay zero (metrics.py:571).
 *   Detections are ordered like np.argsort(-conf, kind="stable").  The O(nc x 1000) tail (F1, smoothing, arg-max, tp / fp
 *   rounding, metrics.py:602-622) is host work on these arrays.  Workspace from el_ap_per_class_workspace_bytes; N < 2^31.
 * el_scale_boxes: scale_boxes + clip_boxes, utils/ops.py:92-127 and :319-338, in place on n rows of >= 4 fp32 values (row pitch in
 *   elements): (x - pad) / gain in IEEE fp32, then clamp x to [0, clip_w] and y to [0, clip_h]; xywh != 0 leaves columns 2, 3
 *   un-padded as the reference does.  gain / pad are computed by the caller exactly as the reference computes them (host scalars). */
int el_ap_per_class_workspace_bytes(int64_t N, int T, int nc, size_t* bytes);
int el_ap_per_class(const uint8_t* tp, const float* conf, const float* pred_cls, int64_t N, int T, const float* classes, const int64_t* n_labels,
                    int nc, double eps, void* workspace, size_t workspace_bytes, double* ap, double* p_curve, double* r_curve, double* prec_values,
                    int32_t* n_pred, void* stream);
int el_scale_boxes(float* boxes, int64_t n, int64_t row_stride, float pad_x, float pad_y, float gain, int padding, int xywh, float clip_w,
                   float clip_h, void* stream);

/* ---- 8f-1. task-aligned target assignment: TaskAlignedAssigner.forward, utils/tal.py:14-295 (CIoU: utils/metrics.py:74-134) ----
 * scores (B,A,nc) class probabilities, boxes (B,A,4) predicted xyxy in pixels, anchors (A,2) cell centres in pixels,
 * gt_labels (B,M) class ids as floats (what the reference's target tensor holds), gt_boxes (B,M,4) xyxy pixels, gt_valid (B,M)
 * bytes (mask_gt, loss.py:374); all dense fp32 / uint8.  metric = score[gt class]^alpha * clamp(CIoU, 0)^beta over the anchors whose
 * centre lies strictly inside a valid ground truth, the `topk` best anchors per ground truth, multi-claims to the larger overlap.
 * Outputs (dense): labels (B,A) int64, tboxes (B,A,4) fp32 (16-byte aligned), tscores (B,A,nc) fp32 normalised soft one-hot,
 * fg (B,A) bytes, gt_idx (B,A) int64 -- the reference's return tuple (tal.py:118).  Background anchors carry ground truth 0's
 * label / box like the reference.  Three kernels over a (B,M,A) workspace; no host synchronisation.  M >= 1 (no targets: the
 * caller returns the reference's constant tuple, tal.py:64-71). */
int el_tal_workspace_bytes(int B, int M, int A, size_t* bytes);
int el_tal_assign(const float* scores, const float* boxes, const float* anchors, const float* gt_labels, const float* gt_boxes,
                  const uint8_t* gt_valid, int B, int A, int nc, int M, int topk, float alpha, float beta, float eps, void* workspace,
                  size_t workspace_bytes, int64_t* labels, float* tboxes, float* tscores, uint8_t* fg, int64_t* gt_idx, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EDGELINE_B200_H */
