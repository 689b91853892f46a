/*
 * ay2.h — C-ABI of libay2.so, the B200 (sm_100a) hot path of ayolov2_b200.
 *
 * The reference (j-marple-dev/AYolov2) has no FFI of its own for this path: the boundary is the Python
 * operator API (kindle modules named in res/configs/model/*.yaml, ComputeLoss, non_max_suppression,
 * batched_nms). This header declares what the replacement exports UNDER that Python boundary; every entry
 * point cites the reference interface it replaces (file:line relative to the reference tree). The Python
 * host (ayolov2_b200/_lib.py) binds these with ctypes; INTEGRATION.md shows the reference-side stub.
 *
 * Conventions
 *  - Plain pointers + sizes only. All data pointers are DEVICE pointers owned by the caller (PyTorch);
 *    the library never allocates, frees or retains tensor memory past a call, except that a *plan* keeps
 *    the raw addresses it was created with (the caller keeps those buffers alive and at the same address;
 *    this is what makes CUDA-graph capture of a step possible).
 *  - Every function returns 0 on success, a negative ay2 error code otherwise, and never throws.
 *    ay2_last_error_string() returns a thread-local description of the last failure.
 *  - `stream` is a cudaStream_t passed as void*. All launches are asynchronous on that stream.
 *  - Activations are NHWC bf16 with an explicit channel stride (so a tensor may be a channel slice of a
 *    wider concat buffer); weights are bf16 [Cout_pad][KH*KW*Cin] (K-major) with the BatchNorm folded in.
 */
#ifndef AY2_H_
#define AY2_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AY2_OK 0
#define AY2_ERR_INVALID -1   /* bad argument / unsupported shape */
#define AY2_ERR_CUDA -2      /* a CUDA runtime / driver call failed */
#define AY2_ERR_NO_DEVICE -3 /* no sm_100 device */

#define AY2_ACT_NONE 0
#define AY2_ACT_SILU 1

int ay2_version(void);
const char* ay2_last_error_string(void);
/* sha256 (hex) of the sources + compile flags this binary was built from; the Python loader compares it with the tree. */
const char* ay2_source_hash(void);
/* Number of kernels this library has launched since load (bench.py's gpu_launches). */
int64_t ay2_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Fused Conv2d + folded BatchNorm + SiLU (+ residual add), NHWC bf16, implicit GEMM on tcgen05.
 * Replaces kindle.modules.conv.Conv.forward (conv -> batch_norm -> activation; reference call sites
 * train.py:137, scripts/train/yolo_trainer.py:322-323, scripts/utils/train_utils.py:436-444), the
 * Bottleneck shortcut add, and — because input and output carry channel strides/offsets — the
 * kindle Concat module (res/configs/model/yolov5s.yaml:37).
 * Also runs each link of the Tucker-2 chain built by scripts/tensor_decomposition/decomposition.py:363-424.
 * ---------------------------------------------------------------------------------------------- */
typedef struct ay2_conv_desc {
  int32_t batch;
  int32_t in_h, in_w;     /* input spatial size */
  int32_t cin;            /* input channels consumed (multiple of 16) */
  int32_t in_cstride;     /* channel stride (elements) of the input buffer, multiple of 8 */
  int32_t out_h, out_w;   /* output spatial size */
  int32_t cout;           /* output channels produced */
  int32_t out_cstride;    /* channel stride (elements) of the output buffer, multiple of 8 */
  int32_t kh, kw;         /* kernel size (1x1, 3x3; KxK stride 1 or 2) */
  int32_t stride;         /* 1 or 2 */
  int32_t pad;            /* symmetric zero padding */
  int32_t act;            /* AY2_ACT_* */
  int32_t res_cstride;    /* channel stride of the residual buffer (0 = no residual) */
  int32_t cout_pad;       /* rows in the packed weight matrix (>= cout, multiple of the N tile) */
  int32_t pad_w;          /* horizontal padding if different from `pad` (-1 = same) */
  int32_t in_pix_stride;  /* elements between horizontally adjacent input pixels (0 = in_cstride). A value below
                             `cin` makes each input "pixel" an overlapping window of neighbouring pixels: the
                             16-channel space-to-depth stem is run as kh x 1 taps over 4-pixel windows (cin 64) */
  int32_t in_row_pixels;  /* physical pixels per input row (0 = in_w); > in_w for a horizontally padded buffer */
  int32_t out_pix_stride; /* elements between horizontally adjacent OUTPUT pixels (0 = out_cstride) ... */
  int32_t out_row_pixels; /* ... and output pixels per physical row (0 = out_w). With pix stride 2*cstride and row pixels
                             = full width, the output (and the residual) is one (row, column) parity sub-grid of a
                             twice-as-large tensor: how the data gradient of a stride-2 conv is written (4 launches) */
  int32_t cin_split;      /* > 0 (stride-1 convs): input channels [0, cin_split) come from `in`, channels [cin_split, cin)
                             from `in2` -- torch.cat([a, b], 1) resolved by the consumer when a and b live in different
                             buffers (kindle C3 conv3 after out-of-place fused bottlenecks) */
  int32_t in2_cstride;    /* channel stride of the second input buffer */
  int32_t x3;             /* 1: split-precision ("bf16x3") verification mode. Every activation tensor is stored as the planes
                             [hi | lo | hi] (hi = bf16(v), lo = bf16(v - hi)) of each producer's output segment and weights
                             as [w_hi | w_hi | w_lo] along K, so `cin` counts 3x the logical input channels and the main
                             loop is unchanged; the epilogue applies the exact SiLU, splits, and writes the three planes of
                             `cout` channels each at out, out + cout, out + 2 cout (the residual is read as hi + lo from
                             residual, residual + cout). fp32-equivalent accuracy on the tcgen05 path; ~4x the work. */
  int32_t stride_w;       /* horizontal stride if different from `stride` (0 = same). Only (stride 2, stride_w 1) is built: a
                             3x3 / stride-2 conv over 32 channels reads PAIRS of pixels as 64-channel rows -- the tensor viewed
                             as [B, H, W/2, 64] -- as a 3x2-tap conv (left pad 1, no right pad: out_w = in_w) with stride 2 over
                             rows only; 128-byte operand rows and 6 instead of 9 taps */
  const void* in2;        /* second input (device pointer, same batch / height / width), NULL when cin_split == 0 */
} ay2_conv_desc;

typedef struct ay2_conv_plan ay2_conv_plan;

/* N-tile (and therefore the cout_pad granularity) the library will use for `cout`. */
int ay2_conv_block_n(int32_t cout);

/* in/out/residual: bf16 NHWC (already offset to the first channel of the slice).
 * weight: bf16 [cout_pad][kh*kw*cin]; bias: fp32 [cout_pad]. */
int ay2_conv_plan_create(const ay2_conv_desc* desc, const void* in, const void* weight, const float* bias,
                         const void* residual, void* out, ay2_conv_plan** plan);
int ay2_conv_plan_run(const ay2_conv_plan* plan, void* stream);
int ay2_conv_plan_destroy(ay2_conv_plan* plan);
/* How the plan will run: out8 = {grid, CTAs per SM, N tile, K chunk, halo kernel (0/1), CTA-pair form -- tcgen05
 * cta_group::2, two CTAs of a cluster sharing every weight tile -- (0/1), cluster size, dynamic shared memory bytes}. */
int ay2_conv_plan_info(const ay2_conv_plan* plan, int32_t* out8);
/* FLOPs (2*MAC, unpadded) one run of the plan performs — used for roofline accounting. */
double ay2_conv_plan_flops(const ay2_conv_plan* plan);

/* Reference SIMT direct convolution (same math, same layouts, no tensor cores). Test infrastructure
 * for full-size parity checks on the GPU where the CPU oracle is too slow; never on the product path. */
int ay2_conv_reference_simt(const ay2_conv_desc* desc, const void* in, const void* weight, const float* bias,
                            const void* residual, void* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused convolution chain: 1x1 (cin -> c1) -> 3x3/s1/p1 (c1 -> c2) [-> 1x1 (c2 -> c3)] in ONE kernel; the c1- and
 * c2-channel intermediates live in shared memory / TMEM only. Replaces
 *   - the nn.Sequential(Conv2d 1x1, Conv2d kxk, Conv2d 1x1) that tucker_decomposition_conv_layer builds
 *     (scripts/tensor_decomposition/decomposition.py:363-424) inside a kindle Conv (its BN + SiLU fold into stage 3), and
 *   - kindle.modules.bottleneck.Bottleneck.forward, x + conv2(conv1(x)) (c3 = 0, residual = x).
 * Every stage is y = act(bias + W * x); the optional residual is added after the last activation.
 * Weights: bf16 K-major w1 [c1][cin], w2 [c2][9*c1] (tap-major: (ky*3+kx)*c1 + c), w3 [c3][c2];
 * bias: fp32 [c1 + c2 + c3] (zeros where the reference has no bias). `out` must not alias `in`.
 * ---------------------------------------------------------------------------------------------- */
typedef struct ay2_chain_desc {
  int32_t batch, in_h, in_w;  /* stride 1: output size == input size */
  int32_t cin, in_cstride;
  int32_t c1, act1;
  int32_t kh, kw, stride, pad; /* stage 2: 3, 3, 1, 1 */
  int32_t c2, act2;
  int32_t c3, act3;            /* c3 == 0: no third stage */
  int32_t out_cstride;
  int32_t res_cstride;         /* 0 = no residual */
} ay2_chain_desc;
typedef struct ay2_chain_plan ay2_chain_plan;
/* 1 if the fused kernel covers this chain (otherwise run the links as separate ay2_conv_plan launches). */
int ay2_chain_supported(const ay2_chain_desc* desc);
int ay2_chain_plan_create(const ay2_chain_desc* desc, const void* in, const void* w1, const void* w2, const void* w3,
                          const float* bias, const void* residual, void* out, ay2_chain_plan** plan);
int ay2_chain_plan_run(const ay2_chain_plan* plan, void* stream);
int ay2_chain_plan_destroy(ay2_chain_plan* plan);
double ay2_chain_plan_flops(const ay2_chain_plan* plan);
/* Diagnostics: chosen launch configuration {CTAs/SM, grid, smem bytes, X slots, W slots, staging aliased, TMEM columns,
 * stage-2 K chunk}, and an optional device buffer (16 x 16 uint64) into which CTA 0 records %globaltimer per phase. */
int ay2_chain_plan_info(const ay2_chain_plan* plan, int32_t* out8);
int ay2_chain_plan_set_debug(ay2_chain_plan* plan, unsigned long long* dbg);
/* Diagnostics for the single-conv plan: info4 = {grid, CTAs/SM, N tile, tiles}; dbg = device buffer of grid x 16 uint64
 * into which every CTA records %globaltimer per phase (tools/conv_timeline.py), or NULL to switch it off. */
int ay2_conv_plan_set_debug(ay2_conv_plan* plan, unsigned long long* dbg, int32_t* info4);

/* ------------------------------------------------------------------------------------------------
 * Input side: NCHW image (uint8 or fp32) -> 2x2 space-to-depth NHWC bf16 with 16 channels
 * (12 used: channel = (dy*2+dx)*3 + c for input pixel (2y+dy, 2x+dx); 4 zero), scaled by `scale`.
 * Replaces YoloValidator.prepare_img / AbstractTrainer.prepare_img (scripts/utils/train_utils.py:255-260,
 * scripts/train/abstract_trainer.py:252-261) and turns the 6x6/s2/p2 stem Conv
 * (res/configs/model/yolov5s.yaml:21) and kindle Focus (res/configs/model/yolov5_v5.yaml:21) into a 3x3/s1 conv.
 * ---------------------------------------------------------------------------------------------- */
#define AY2_DT_U8 0
#define AY2_DT_F32 1
/* out: [batch, h/2, out_row_pixels, 16]; logical pixel x is written at physical column x + out_x_offset
 * (columns outside are never touched: the caller zero-fills them once; they are the conv's horizontal padding). */
int ay2_space_to_depth(const void* img, int32_t dtype, int32_t batch, int32_t h, int32_t w, float scale,
                       void* out, int32_t out_row_pixels, int32_t out_x_offset, void* stream);

/* SPPF / SPP pooling: x1 = in slice; writes maxpool windows k1,k2,k3 (stride 1, same padding) into three
 * channel slices. Replaces kindle.modules.poolings.SPPF / SPP (yolov5s.yaml:33, yolov5_v5.yaml:32). */
int ay2_sppf_pool(const void* in, int32_t batch, int32_t h, int32_t w, int32_t c, int32_t cstride, int32_t k1,
                  int32_t k2, int32_t k3, void* out1, void* out2, void* out3, void* stream);

/* Nearest 2x upsample into a channel slice: replaces nn.Upsample(scale_factor=2) + Concat
 * (res/configs/model/yolov5s.yaml:36-37). */
int ay2_upsample2x(const void* in, int32_t batch, int32_t h, int32_t w, int32_t c, int32_t in_cstride, void* out,
                   int32_t out_cstride, void* stream);

/* ------------------------------------------------------------------------------------------------
 * YOLOHead decode (eval mode): per level, logits NHWC bf16 [B,ny,nx,cstride] (channel = a*no + o) ->
 *   pred  fp32 [B, total_rows, no] rows [row_offset + a*ny*nx + y*nx + x]  (sigmoid, xy/wh decode)
 *   raw   fp32 [B, na, ny, nx, no] (optional, may be NULL: the train-layout tensor)
 * Replaces kindle.modules.yolo_head.YOLOHead.forward (consumers: scripts/utils/train_utils.py:441-444,
 * scripts/loss/losses.py:245-255).
 * ---------------------------------------------------------------------------------------------- */
int ay2_head_decode(const void* logits, int32_t batch, int32_t ny, int32_t nx, int32_t cstride, int32_t na,
                    int32_t no, float stride_px, const float* anchor_wh_px /* [na*2] device */, float* pred,
                    int64_t total_rows, int64_t row_offset, float* raw, void* stream);

/* Split-precision ("bf16x3", ay2_conv_desc::x3) forms of the data-movement layers, and the general head decode
 * (csrc/precise.cu). A tensor of C channels is three planes [hi | lo | hi] of C channels; pointers address plane 0.
 *   ay2_space_to_depth_x3: out [batch, h/2, out_row_pixels, 48] = planes of the 16-channel pixel; value = pixel / divisor.
 *   ay2_sppf_pool_x3:      as ay2_sppf_pool on split-precision segments of `c` channels (max of hi + lo, re-split).
 *   ay2_head_decode2:      ay2_head_decode with lo_offset > 0: logits = plane 0 + the plane lo_offset channels further;
 *                          flags bit 0: boxes as x1 y1 x2 y2 (YOLOHead.out_xyxy); bit 1: exact fp32 sigmoid.
 * Upsample needs no special form (a split-precision segment is copied like 3 c ordinary channels). */
int ay2_space_to_depth_x3(const void* img, int32_t dtype, int32_t batch, int32_t h, int32_t w, float divisor, void* out,
                          int32_t out_row_pixels, int32_t out_x_offset, void* stream);
int ay2_sppf_pool_x3(const void* in, int32_t batch, int32_t h, int32_t w, int32_t c, int32_t cstride, int32_t k1, int32_t k2,
                     int32_t k3, void* out1, void* out2, void* out3, void* stream);
int ay2_head_decode2(const void* logits, int32_t lo_offset, int32_t batch, int32_t ny, int32_t nx, int32_t cstride, int32_t na,
                     int32_t no, float stride_px, const float* anchor_wh_px, int32_t flags, float* pred, int64_t total_rows,
                     int64_t row_offset, float* raw, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Batched NMS: replaces scripts/utils/metrics.py:285-443 non_max_suppression(nms_type="nms") and the
 * torchvision.ops.nms call at :385 for the whole batch in a fixed number of launches.
 *   pred: fp32 [B, n, no] (xywh, obj, cls...) ; out_det: fp32 [B, max_det, 6]; out_count: int32 [B].
 * workspace: ay2_nms_workspace_bytes(...) bytes. Selection is bit-identical to the reference on identical
 * inputs (stable descending score order, fp32 IoU, strict > iou_thres).
 * ---------------------------------------------------------------------------------------------- */
typedef struct ay2_nms_params {
  double iou_thres;     /* compared as (double)iou > iou_thres, like torchvision's CPU kernel */
  float conf_thres;     /* compared in fp32, like `prediction[..., 4] > conf_thres` */
  float max_wh;         /* metrics.py:326 (4096) */
  int32_t batch, n, no;
  int32_t multi_label;  /* metrics.py:330,359-364 */
  int32_t agnostic;     /* metrics.py:383 */
  int32_t max_det;      /* metrics.py:293 (300) */
  int32_t max_nms;      /* metrics.py:327 (30000) */
  int32_t max_candidates; /* per-image capacity of the candidate list (workspace sizing) */
} ay2_nms_params;
size_t ay2_nms_workspace_bytes(const ay2_nms_params* p);
/* class_mask: optional device uint8[nc] (metrics.py:367-368 `classes` filter), NULL = keep all.
 * overflow_flag: optional device int32, set to 1 if some image had more than max_candidates candidates. */
int ay2_nms_batched(const float* pred, const ay2_nms_params* p, const uint8_t* class_mask, void* workspace,
                    size_t workspace_bytes, float* out_det, int32_t* out_count, int32_t* overflow_flag, void* stream);

/* Fused head: the same NMS fed straight from the bf16 YOLOHead logits (NHWC [B, ny, nx, cstride], channel = a*no + o)
 * of every pyramid level. Candidate scores and boxes are decoded on the fly with the arithmetic of ay2_head_decode,
 * so the result is bit-identical to ay2_head_decode followed by ay2_nms_batched, without materialising the
 * (B, sum na*ny*nx, no) tensor. params.n must equal sum_l na*ny_l*nx_l (row order = the reference's cat order). */
#define AY2_NMS_MAX_LEVELS 5
#define AY2_NMS_MAX_ANCHORS 8
typedef struct ay2_head_levels {
  int32_t nl, na;
  const void* logits[AY2_NMS_MAX_LEVELS];
  int32_t ny[AY2_NMS_MAX_LEVELS], nx[AY2_NMS_MAX_LEVELS], cstride[AY2_NMS_MAX_LEVELS];
  float stride_px[AY2_NMS_MAX_LEVELS];
  float anchor_px[AY2_NMS_MAX_LEVELS][AY2_NMS_MAX_ANCHORS][2]; /* YOLOHead.anchor_grid */
} ay2_head_levels;
int ay2_nms_from_logits(const ay2_head_levels* levels, const ay2_nms_params* p, const uint8_t* class_mask,
                        void* workspace, size_t workspace_bytes, float* out_det, int32_t* out_count,
                        int32_t* overflow_flag, void* stream);

/* Candidate generation fused into the detect-head convolutions (the producing kernel): after
 * ay2_conv_plan_set_head_candidates, every run of that plan also scores its output tile for NMS candidates
 * (metrics.py:313-364: obj > conf, conf_c = cls_c * obj, best class or multi_label) and appends the 64-bit keys to
 * the NMS workspace, so no kernel re-reads the logits to find candidates. One step is then
 *   ay2_nms_candidates_begin (zeroes the per-image counters)  ->  the head convolutions of every level, any order
 *   ->  ay2_nms_from_candidates (sort + greedy suppression + output; reads boxes of the survivors from the logits).
 * `p` carries conf_thres / multi_label / max_candidates / batch / n / no (the values the NMS call will use);
 * `row_off` is the level's first row in the reference's cat order (sum over earlier levels of na*ny*nx). The plan
 * must produce all na*no channels in one N tile (cout <= 256). Pass p = NULL to switch the fusion off.
 * Results are bit-identical to ay2_nms_from_logits on the same logits. */
int ay2_conv_plan_set_head_candidates(ay2_conv_plan* plan, const ay2_nms_params* p, int32_t na, int32_t row_off,
                                      const uint8_t* class_mask, void* nms_workspace, size_t workspace_bytes);
int ay2_nms_candidates_begin(const ay2_nms_params* p, void* workspace, size_t workspace_bytes, void* stream);
int ay2_nms_from_candidates(const ay2_head_levels* levels, const ay2_nms_params* p, void* workspace, size_t workspace_bytes,
                            float* out_det, int32_t* out_count, int32_t* overflow_flag, void* stream);

/* Pairwise IoU matrix: replaces scripts/utils/metrics.py:138-164 box_iou (fast/matrix/merge NMS in scripts/utils/nms.py:74-110,
 * mAP matching in scripts/utils/train_utils.py:294-401). box1: fp32 (n, 4) xyxy, box2: fp32 (m, 4) xyxy, out: fp32 (n, m);
 * bit-identical to the reference expression on identical inputs. */
int ay2_box_iou(const float* box1, int32_t n, const float* box2, int32_t m, float* out, void* stream);

/* torchvision.ops.nms on a plain box list -- the call inside non_max_suppression (metrics.py:385) and its "batched_nms" /
 * "merge_nms" nms_type branches (:391-431). boxes: fp32 [n][4] xyxy; order: indices by descending score (stable);
 * mask_ws: n * ceil(n / 64) uint64 of scratch; keep: [n] kept ORIGINAL indices in score order; count: their number.
 * Greedy suppression IoU > iou_thres with the reference's fp32 IoU and its comparison against the double threshold. */
int ay2_nms_boxes(const float* boxes, const int32_t* order, int32_t n, double iou_thres, unsigned long long* mask_ws,
                  int32_t* keep, int32_t* count, void* stream);

/* The non-default suppression rules, batched over images (csrc/nms_variants.cu): replace the per-image host loops and n x n
 * IoU matrices of scripts/utils/metrics.py:388-431 (nms_type "batched_nms" / "fast_nms" / "matrix_nms" / "merge_nms") and of
 * scripts/utils/nms.py:63-110 (the same rules on the val2 path).
 *   ay2_nms_candidate_table: pred fp32 [B, n, no] -> table fp32 [B][cap][8] = {x1, y1, x2, y2, conf, cls, 0, 0} in the
 *     reference's candidate order (metrics.py:337-368; p->conf_thres / multi_label / max_nms are used), counts int32 [B],
 *     max_coord fp32 [B] (largest box coordinate among an image's candidates), flags int32 [1]: bit 0 = some image had
 *     more than `cap` candidates (table truncated), bit 1 = some image had more than p->max_nms (metrics.py:378-379 applies).
 *   ay2_nms_fast:   survivors = rows no EARLIER row overlaps with IoU >= iou_thres (fp32 compare), table order, first out_cap.
 *   ay2_nms_matrix: every row, conf scaled by the gaussian (sigma 0.5) matrix-NMS decay; first out_cap rows.
 *     class_offset: boxes are shifted by cls * class_offset before the IoU (0 = class-agnostic / no separation);
 *     colmax_ws: fp32 [B][cap] scratch. out_det fp32 [B][out_cap][6], out_count int32 [B].
 *   ay2_nms_merge: det / det_count hold greedy-NMS survivors in kept order (ay2_nms_batched output, det_cap <= 1024); every
 *     survivor's box becomes the conf-weighted mean of the candidates overlapping it (IoU > iou_thres) and survivors with
 *     no second supporter are dropped, for images with n_min_excl < candidates < n_max_excl (metrics.py:420: 1 < n < 3000).
 *   ay2_nms_batched_scaled: ay2_nms_batched with the class offset of torchvision.ops.boxes.batched_nms: max_coord[b] + 1
 *     per image (metrics.py:391-394) instead of p->max_wh. */
int ay2_nms_candidate_table(const float* pred, const ay2_nms_params* p, const uint8_t* class_mask, float* table, int32_t cap,
                            int32_t* counts, float* max_coord, int32_t* flags, void* stream);
int ay2_nms_fast(const float* table, const int32_t* counts, int32_t batch, int32_t cap, float class_offset, float iou_thres,
                 float* colmax_ws, float* out_det, int32_t out_cap, int32_t* out_count, void* stream);
int ay2_nms_matrix(const float* table, const int32_t* counts, int32_t batch, int32_t cap, float class_offset, float* colmax_ws,
                   float* out_det, int32_t out_cap, int32_t* out_count, void* stream);
int ay2_nms_merge(const float* table, const int32_t* counts, int32_t batch, int32_t cap, float class_offset, float iou_thres,
                  int32_t n_min_excl, int32_t n_max_excl, float* det, int32_t det_cap, int32_t* det_count, void* stream);
int ay2_nms_batched_scaled(const float* pred, const ay2_nms_params* p, const uint8_t* class_mask, const float* max_coord,
                           void* workspace, size_t workspace_bytes, float* out_det, int32_t* out_count, int32_t* overflow_flag,
                           void* stream);

/* ------------------------------------------------------------------------------------------------
 * Input side: letterbox + BGR->RGB + HWC->CHW + collate of a whole batch in one launch. Replaces
 * LoadImages._letterbox (scripts/data_loader/data_loader.py:395-459: cv2.resize INTER_LINEAR to the unpadded size +
 * cv2.copyMakeBorder), the transpose / channel flip of :388-389 and the torch.stack of collate_fn (:461-477, :905-909);
 * with AY2_LB_S2D_BF16 also prepare_img + ay2_space_to_depth (the uint8 NCHW batch is never written).
 *   arena : DEVICE bytes holding the loaded images, HWC BGR uint8 (ragged sizes; rows src_row_bytes apart)
 *   table : DEVICE array of `batch` records; the geometry is the host's (data_loader.py:428-455 is scalar arithmetic):
 *           dst_w / dst_h = new_unpad (the size after the resize; == src size: plain copy), top / left = the border
 *   out   : AY2_LB_NCHW_U8  -> uint8 [batch][3][out_h][out_w] RGB (the reference's collated tensor), bit-exact with cv2
 *           AY2_LB_S2D_BF16 -> bf16 [batch][out_h/2][out_row_pixels][16] at column offset out_x_offset, value * scale
 *                              (identical to ay2_space_to_depth of the uint8 tensor)
 *   color_bgr : border colour, b | g << 8 | r << 16 (the reference: 114, 114, 114). out_h even, out_w % 4 == 0.
 *   kinds : which kinds of image the table holds (AY2_LB_HAS_*; 0 = unknown). Images that enter at their final size and
 *           images that are resized run in separate launches over the same grid (different register budgets). */
#define AY2_LB_NCHW_U8 0
#define AY2_LB_S2D_BF16 1
#define AY2_LB_HAS_COPY 1
#define AY2_LB_HAS_RESIZE 2
typedef struct ay2_letterbox_image {
  int64_t src_offset;     /* byte offset of the image inside `arena` */
  double scale_x;         /* 1.0 / ((double)dst_w / src_w): cv2's source step per destination pixel */
  double scale_y;         /* 1.0 / ((double)dst_h / src_h) */
  int32_t src_h, src_w;   /* the loaded image */
  int32_t src_row_bytes;  /* >= 3 * src_w */
  int32_t dst_h, dst_w;   /* new_unpad (h, w) */
  int32_t top, left;      /* border above / left of the resized image */
  int32_t reserved;
} ay2_letterbox_image;    /* 56 bytes */
int ay2_letterbox_collate(const uint8_t* arena, const ay2_letterbox_image* table, int32_t batch, int32_t kinds, int32_t out_h,
                          int32_t out_w, uint32_t color_bgr, int32_t out_kind, void* out, int32_t out_row_pixels,
                          int32_t out_x_offset, float scale, void* stream);
/* LoadImages._load_image after the decode (scripts/data_loader/data_loader.py:320-329): every image of the table is resized
 * inside `arena` (DEVICE bytes; src and dst are byte offsets, HWC BGR uint8) with cv2's INTER_AREA (AY2_LR_AREA: the image
 * shrinks in both directions) or INTER_LINEAR (AY2_LR_LINEAR) arithmetic, bit-exact. scale_x / scale_y as in
 * ay2_letterbox_image. max_dst_pixels = the largest dst_h * dst_w of the table (grid sizing). The resized images are then
 * ordinary inputs of ay2_letterbox_collate (its src_offset = this dst_offset). */
#define AY2_LR_LINEAR 0
#define AY2_LR_AREA 1
typedef struct ay2_load_resize_image {
  int64_t src_offset, dst_offset;
  double scale_x, scale_y;      /* 1.0 / ((double)dst / src) */
  int32_t src_h, src_w, src_row_bytes;
  int32_t dst_h, dst_w, dst_row_bytes;
  int32_t mode, reserved;
} ay2_load_resize_image;        /* 64 bytes */
int ay2_load_resize(uint8_t* arena, const ay2_load_resize_image* table, int32_t count, int32_t max_dst_pixels, void* stream);

/* YoloTrainer.multi_scale (scripts/train/yolo_trainer.py:223-248) with prepare_img (scripts/train/abstract_trainer.py:252-261)
 * fused in: out fp32 [batch][3][out_h][out_w] = F.interpolate(img * pre_scale, (out_h, out_w), mode="bilinear",
 * align_corners=False); img NCHW uint8 (AY2_DT_U8) or fp32 (AY2_DT_F32). Replaces the .float() / 255 pass, the ATen
 * interpolation and the copy into the training engine's static input (the engine passes its input buffer as `out`). */
int ay2_resize_bilinear(const void* img, int32_t dtype, int32_t batch, int32_t h, int32_t w, float pre_scale, float* out,
                        int32_t out_h, int32_t out_w, void* stream);
/* LoadImagesAndLabels.collate_fn (data_loader.py:905-909): labels fp32 [total][6] (already concatenated), offsets int32
 * [batch + 1] (row range of every image, DEVICE): writes the image index into column 0. */
int ay2_collate_labels(float* labels, const int32_t* offsets, int32_t batch, int32_t total, void* stream);

/* Knowledge-distillation pseudo-labels (the KD caller of the NMS kernels): replaces
 * SoftTeacherTrainer.prepare_labels_for_augmention / filter_invalid and the label assembly of get_pseudo_labeled_batch
 * (scripts/train/kd_trainer.py:385-397,436-487) + xyxy2xywh with its validity correction (scripts/utils/general.py:250-295).
 * det / counts: the NMS output ([batch][max_det][6] xyxy conf cls, [batch]). Keeps score > score_thr and (use_min_size)
 * width, height > min_size; labels: fp32 [batch * max_det][6] = image, class, x, y, w, h (normalised by width / height), rows in
 * image order then detection order; out_counts: int32 [batch + 1] = labels per image, then the total. */
int ay2_pseudo_labels(const float* det, const int32_t* counts, int32_t batch, int32_t max_det, float score_thr, float min_size,
                      int32_t use_min_size, float width, float height, float* labels, int32_t* out_counts, void* stream);

/* Validation statistics: replaces the per-image host loop of YoloValidator.statistics_per_image / process_batch
 * (scripts/utils/train_utils.py:294-401) for a whole batch. det / counts: the NMS output ([batch][max_det][6], [batch]);
 * labels: fp32 [nt][6] = image, class, box; meta == NULL: boxes are xyxy in the detections' coordinates (the plain
 * process_batch contract); meta = fp32 [batch][5] {gain, pad_x, pad_y, native_w, native_h}: label boxes are xywh pixels of
 * the network input and both sides are mapped to the native image first (xywh2xyxy + scale_coords + clip_coords,
 * scripts/utils/general.py:203-230,316-358). labels_cap: upper bound of labels per image (shared-memory sizing).
 * correct: uint8 [batch][max_det][niou], 1 where the detection is matched with IoU >= iouv[k]. */
int ay2_match_detections(const float* det, const int32_t* counts, int32_t batch, int32_t max_det, const float* labels,
                         int32_t nt, int32_t labels_cap, const float* meta, const float* iouv, int32_t niou,
                         uint8_t* correct, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Detection loss forward + analytic backward: replaces scripts/loss/losses.py:168-391 (ComputeLoss.__call__,
 * build_targets) and scripts/utils/metrics.py:60-135 (bbox_iou, CIoU) for the default configuration
 * (fl_gamma = 0, gr = 1, autobalance off).
 *   preds[i]  : fp32 (bs, na, ny_i, nx_i, 5+nc) logits, contiguous        (losses.py:245)
 *   grads[i]  : same shape, ZERO-INITIALISED by the caller, receives d(loss*bs)/d(preds[i]) * (*gscale);
 *               pass grads = NULL for a forward-only evaluation
 *   targets   : fp32 (nt, 6) [img, cls, x, y, w, h] normalised             (scripts/data_loader/data_loader.py:888-909)
 *   anchors   : fp32 (nl, na, 2) in grid units (YOLOHead.anchors)
 *   gscale    : optional device scalar multiplied into every gradient (autograd grad_output), NULL = 1
 *   out5      : device fp32 [loss*bs, lbox, lobj, lcls, loss]               (losses.py:294-300)
 * ---------------------------------------------------------------------------------------------- */
#define AY2_LOSS_MAX_LEVELS 5
typedef struct ay2_loss_params {
  int32_t nl, na, nc, bs, nt;
  int32_t ny[AY2_LOSS_MAX_LEVELS], nx[AY2_LOSS_MAX_LEVELS];
  float balance[AY2_LOSS_MAX_LEVELS]; /* losses.py:204-206 */
  float anchor_t, box, obj, cls;      /* hyp gains (res/configs/cfg/train_config.yaml:29-54) */
  float cls_pw, obj_pw;               /* BCE pos_weight */
  float cp, cn;                       /* smooth_BCE targets (losses.py:16-27) */
  float fl_gamma, fl_alpha;           /* > 0: FocalLoss wrapper around both BCE criteria (losses.py:64-114,193-196; alpha 0.25) */
} ay2_loss_params;
size_t ay2_yolo_loss_workspace_bytes(const ay2_loss_params* p);
int ay2_yolo_loss(const ay2_loss_params* p, const float* const* preds, float* const* grads, const float* targets,
                  const float* anchors, const float* gscale, void* workspace, size_t workspace_bytes, float* out5,
                  void* stream);

/* ------------------------------------------------------------------------------------------------
 * Training step pieces (scripts/train/yolo_trainer.py:289-358: forward under autocast, ComputeLoss, backward, SGD step,
 * EMA). In train mode a kindle Conv is conv -> BatchNorm2d(batch statistics, eps 1e-3, momentum 0.03) -> SiLU: the conv
 * runs through ay2_conv_plan_* with act = NONE and zero bias (raw output z), then the kernels below.
 * All activation tensors are NHWC bf16 channel slices (pointer to the first channel + channel stride).
 * ---------------------------------------------------------------------------------------------- */
/* per-channel sum / sum of squares of z over npix pixels, ACCUMULATED into double[c] buffers (zero them first) */
int ay2_bn_stats(const void* z, int64_t npix, int32_t c, int32_t cstride, double* sum, double* sumsq, void* stream);
/* mean / invstd (biased variance) + running-statistics update (unbiased variance), nn.BatchNorm2d semantics */
int ay2_bn_finalize(const double* sum, const double* sumsq, int64_t n, int32_t c, float eps, float momentum,
                    float* running_mean, float* running_var, float* mean, float* invstd, void* stream);
/* y = act(gamma * (z - mean) * invstd + beta) (+ residual) */
int ay2_bn_act_fwd(const void* z, int64_t npix, int32_t c, int32_t z_cstride, const float* mean, const float* invstd,
                   const float* gamma, const float* beta, int32_t act, void* y, int32_t y_cstride, const void* residual,
                   int32_t res_cstride, void* stream);
/* backward of the above w.r.t. z: s1 = sum dy*act'(u) (= d beta), s2 = sum dy*act'(u)*xhat (= d gamma) as double[c];
 * dz = gamma*invstd*(dy*act' - s1/N - xhat*s2/N). dz may alias dy. */
int ay2_bn_act_bwd(const void* dy, int32_t dy_cstride, const void* z, int32_t z_cstride, int64_t npix, int32_t c,
                   const float* mean, const float* invstd, const float* gamma, const float* beta, int32_t act, double* s1,
                   double* s2, void* dz, int32_t dz_cstride, void* stream);
/* ay2_bn_act_bwd that also ACCUMULATES the parameter gradients (d beta = s1, d gamma = s2) into two fp32 arrays of c
 * elements -- the slices of the trainer's flat gradient buffer -- from inside the apply pass. sums_zeroed != 0: s1 / s2
 * are already zero (the caller clears the sums of every layer with one fill per backward pass). */
int ay2_bn_act_bwd_grads(const void* dy, int32_t dy_cstride, const void* z, int32_t z_cstride, int64_t npix, int32_t c,
                         const float* mean, const float* invstd, const float* gamma, const float* beta, int32_t act,
                         double* s1, double* s2, void* dz, int32_t dz_cstride, float* dbeta_acc, float* dgamma_acc,
                         int32_t sums_zeroed, void* stream);
/* The same in two phases, for SyncBatchNorm (scripts/train/train_model_builder.py:86-91): phase 1 computes this rank's s1 / s2,
 * the caller sums them over the ranks (all-reduce of 2 c doubles), phase 2 applies with npix_total = pixels of the whole
 * cross-rank batch. (The forward is ay2_bn_stats -> all-reduce of sum / sumsq -> ay2_bn_finalize with the total count.) */
int ay2_bn_act_bwd_phase(const void* dy, int32_t dy_cstride, const void* z, int32_t z_cstride, int64_t npix, int32_t c,
                         const float* mean, const float* invstd, const float* gamma, const float* beta, int32_t act, double* s1,
                         double* s2, void* dz, int32_t dz_cstride, int32_t phase, int64_t npix_total, void* stream);
/* weight gradient of a conv: dw[cout][kh*kw][cin] (fp32, ACCUMULATED) += dz^T (*) x ; tcgen05, split-K */
int ay2_conv_wgrad(const ay2_conv_desc* desc, const void* x, const void* dz, float* dw, void* stream);
/* dst (+)= src over a channel slice: residual / concat / fan-out gradient accumulation */
int ay2_add_slices(const void* src, int32_t src_cstride, void* dst, int32_t dst_cstride, int64_t npix, int32_t c,
                   int32_t accumulate, void* stream);
/* backward of nearest-2x upsample: dx[B,h,w,c] (+)= sum of the 2x2 blocks of dy[B,2h,2w,c] */
int ay2_upsample2x_bwd(const void* dy, int32_t dy_cstride, int32_t batch, int32_t h, int32_t w, int32_t c, void* dx,
                       int32_t dx_cstride, int32_t accumulate, void* stream);
/* backward of MaxPool2d(k, 1, k/2) with nn.MaxPool2d's tie rule; x = the pool's input */
int ay2_maxpool_bwd(const void* x, int32_t x_cstride, const void* dy, int32_t dy_cstride, int32_t batch, int32_t h, int32_t w,
                    int32_t c, int32_t k, void* dx, int32_t dx_cstride, int32_t accumulate, void* stream);
/* YOLOHead train layout: (bs, na, ny, nx, no) fp32 <-> NHWC bf16 [B, ny, nx, cstride] (channel = a*no + o) */
int ay2_head_grad_to_nhwc(const float* grad, int32_t batch, int32_t na, int32_t ny, int32_t nx, int32_t no, void* out,
                          int32_t out_cstride, void* stream);
int ay2_head_logits_to_train(const void* logits, int32_t cstride, int32_t batch, int32_t na, int32_t ny, int32_t nx,
                             int32_t no, float* out, void* stream);
/* per-channel sum over pixels (bias gradients), ACCUMULATED into double[c] */
int ay2_channel_sum(const void* g, int64_t npix, int32_t c, int32_t cstride, double* sum, void* stream);
/* fused SGD (nesterov, weight decay) + EMA over one flat fp32 range (yolo_trainer.py:332-338, torch_utils.py:405-416);
 * inv_scale: optional device scalar multiplied into the gradient (GradScaler unscale), ema may be NULL */
int ay2_sgd_ema_step(float* param, const float* grad, float* momentum_buf, float* ema, int64_t n, float lr, float momentum,
                     float weight_decay, int32_t nesterov, float ema_decay, const float* inv_scale, void* stream);
/* Re-pack every bf16 convolution operand of the training engine from its fp32 parameter in ONE launch (what the reference
 * gets for free from cuDNN reading OIHW fp32 weights directly; here the tcgen05 kernels want K-major bf16, and the data
 * gradient wants flipped / transposed / parity-split copies). `segments`: DEVICE array of nseg records
 *   { const float* src; bf16* dst; const int32_t* idx; int64_t begin; }   (32 bytes each, sorted by begin)
 * dst[i] = bf16(src[idx[i]]) (idx[i] < 0: zero) for the `count` = next.begin - begin elements of each record; every begin and
 * `total` (the end of the last record) are multiples of 8, dst and idx 16-byte aligned. */
int ay2_repack_weights(const void* segments, int32_t nseg, int64_t total, void* stream);
/* The same fused update with per-element parameter groups: group[i] in 0..3 selects lr4[group] / wd4[group] (HOST arrays of
 * four floats; the reference's optimizer has three groups -- BatchNorm weights, decayed weights, biases -- whose learning
 * rates differ during warm-up, scripts/train/yolo_trainer.py:149-168,194-221). grad is multiplied by grad_scale first
 * (1 / world_size after a sum all-reduce, 1 / loss scale). One launch for the whole flat parameter buffer. */
int ay2_sgd_ema_step_groups(float* param, const float* grad, float* momentum_buf, float* ema, const uint8_t* group, int64_t n,
                            const float* lr4, const float* wd4, float momentum, int32_t nesterov, float ema_decay,
                            float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AY2_H_ */
