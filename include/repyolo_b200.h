/*
 * repyolo_b200 -- C ABI of the B200-native Rep-YOLO deploy path (fused conv stack -> Detect decode -> NMS).
 *
 * The reference (DrLSB/Rep-YOLO) is pure Python/PyTorch and has no plugin / FFI layer; the drop-in boundary is the
 * Python signatures listed below.  This header is what a host-language binding (ctypes in `rep-yolo_b200/`) binds to:
 * plain pointers and sizes, no torch types, no exceptions, `int` status (0 = ok, see ry_last_error()).
 * All data pointers are DEVICE pointers unless the name says `host`; `stream` is a cudaStream_t passed as void*.
 * No entry point allocates or synchronises inside the hot calls (ry_forward / ry_run_ops / ry_nms); the caller
 * provides the workspace (ry_plan_workspace_bytes / ry_nms_workspace_bytes).
 *
 * Reference interfaces replaced (file:line relative to the reference repo):
 *   ry_plan_create      <- Model.fuse()              models/yolo.py:681-704  (graph rewrite result: the fused op list;
 *                          the fold arithmetic itself -- common.py:597-657, 3436-3517, torch_utils.py:181-201,
 *                          yolo.py:170-182 -- runs on the host in rep-yolo_b200/fold.py and hands fp32 OIHW weights here)
 *   ry_forward          <- Model.forward / forward_once + IDetect.fuseforward   models/yolo.py:569-619, 135-168
 *   ry_run_ops          <- one top-level module of forward_once (teacher-forced module parity; yolo.py:613)
 *   ry_nms              <- non_max_suppression       utils/general.py:953-1045 (+ xywh2xyxy :265-272,
 *                          torchvision.ops.nms call site :1029)
 *   ry_decode_filter    <- IDetect.fuseforward (yolo.py:139-156) fused with the first step of non_max_suppression,
 *                          `xc = prediction[..., 4] > conf_thres` (general.py:961): same launches as ry_forward, the Detect
 *                          epilogue also leaves the filter result as warp-ballot words
 *   ry_nms_filtered     <- non_max_suppression consuming those words: ordered (prefix-sum) compaction of the survivors only
 */
#ifndef REPYOLO_B200_H
#define REPYOLO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RY_ABI_VERSION 5

typedef struct ry_plan ry_plan;

/* ---- plan IR ------------------------------------------------------------------------------------------------ */

enum ry_dtype { RY_BF16 = 0, RY_F32 = 1, RY_U8 = 2 };

enum ry_tensor_kind {
    RY_T_MAP = 0,      /* activation map, NHWC: [B, H>>level, W>>level, channels] */
    RY_T_VEC = 1,      /* per-image vector:     [B, channels]                      */
    RY_T_EXTERNAL = 2  /* bound per call (image / pred / raw heads), see slot      */
};

enum ry_external_slot { RY_X_IMAGE = 0, RY_X_PRED = 1, RY_X_RAW0 = 2, RY_X_RAW1 = 3, RY_X_RAW2 = 4 };

typedef struct ry_tensor_desc {
    int32_t kind;      /* ry_tensor_kind */
    int32_t dtype;     /* ry_dtype */
    int32_t channels;
    int32_t level;     /* log2 of the spatial down-scale w.r.t. the input image (MAP only) */
    int32_t slot;      /* ry_external_slot (EXTERNAL only) */
    int32_t pad_;
} ry_tensor_desc;

typedef struct ry_view {   /* a channel range of a tensor */
    int32_t tensor;        /* index into the tensor table, -1 = none */
    int32_t c_off;
    int32_t c_len;
} ry_view;

enum ry_op_kind {
    RY_OP_STEM = 1,       /* 3x3 s2 conv on the fp32 NCHW image + bias + SiLU -> NHWC bf16      (RepS_Block L0)   */
    RY_OP_CONV = 2,       /* 1x1 / 3x3 (s1,s2) implicit GEMM on tcgen05 + fused epilogue                            */
    RY_OP_DW5 = 3,        /* depthwise 5x5 s1 + bias + act                                       (GSConv.cv2)      */
    RY_OP_MAXPOOL2 = 4,   /* 2x2 s2 max-pool                                                     (MP)              */
    RY_OP_SPP = 5,        /* 5/9/13 s1 max-pools, written at three channel offsets               (SPPCSPC.m)       */
    RY_OP_UPSAMPLE2 = 6,  /* nearest x2                                                          (nn.Upsample)     */
    RY_OP_CA = 7,         /* global avg-pool + f1/ReLU + f2/sigmoid, out = p*s + p -> VEC         (CA)              */
    RY_OP_ATTN_QK = 8,    /* (stand-alone q/k projection -> fp32 q,k; the planner fuses it into ops 9/10 instead)          */
    RY_OP_CRISSCROSS = 9, /* CrissCrossAttention core: gamma*out + x                                                */
    RY_OP_VERTICAL = 10,  /* VerticalAttention core:   gamma*out + x                                                */
    RY_OP_DETECT = 11,    /* head 1x1 conv on tcgen05 + sigmoid/grid/anchor decode -> pred + raw  (IDetect)         */
    RY_OP_CONV_CHAIN = 12 /* 3x3 s1 conv -> 1x1 conv [-> 1x1 conv] fused in one kernel (small-channel DER_Block stages,    */
                          /* common.py:3646-3651): out0/out1/out2 = optional stores of stage 0/1/2 (tensor -1 = on-chip only) */
};

enum ry_act { RY_ACT_NONE = 0, RY_ACT_SILU = 1 };

typedef struct ry_op_desc {
    int32_t kind;          /* ry_op_kind */
    int32_t layer;         /* top-level reference layer index (models/yolo.py forward_once) this op belongs to */
    ry_view in0;           /* main input                                                           */
    ry_view in1;           /* CONV: residual added after act (same shape as out) | attention: q    */
    ry_view in2;           /* CONV: per-image broadcast vector added after act   | attention: k    */
                           /* CONV with n_src > 1 (1x1 only): in0..in{n_src-1} are the channel-concatenated inputs    */
                           /* (the torch.cat feeding DER_Block.cv1, common.py:3653, never materialises)               */
    ry_view out0;          /* main output (SPP: the 5x5 pool; ATTN_QK: q)                          */
    ry_view out1;          /* CONV/DW5 split store: second half of the channels | SPP: 9x9 | ATTN_QK: k */
    ry_view out2;          /* SPP: 13x13                                                           */
    int32_t ksize;         /* 1 or 3 (CONV), 3 (STEM), 5 (DW5) */
    int32_t stride;        /* 1 or 2 */
    int32_t act;           /* ry_act */
    int32_t cin, cout;
    int32_t level_idx;     /* DETECT: pyramid level (row offset / stride / anchors); CONV (1x1): 1 = MaxPool2d(2,2) of the    */
                           /* activated output fused in the epilogue, out0 is then the pooled map (MP after DER_Block.cv1)   */
    int32_t n_src;         /* CONV: number of concatenated input views (0 or 1 = just in0) */
    int32_t pad_;
    int32_t n_post;        /* CONV_CHAIN: number of fused 1x1 stages (1 or 2); their weights/biases are aux_off[0..3]     */
    int32_t post_cout[2];  /*   = {w1, b1, w2, b2} (fp32 [cout_s][cin_s], [cout_s]), cin_s = previous stage's cout          */
    int32_t post_act[2];   /*   ry_act of each fused stage */
    int32_t pad2_;
    int64_t w_off;         /* byte offsets into the host weight blob; -1 = none.                                    */
    int64_t b_off;         /*   CONV/STEM/DETECT: w = fp32 [cout][cin][k][k] (PyTorch OIHW), b = fp32 [cout]        */
    int64_t aux_off[6];    /*   DW5: w = fp32 [C][5][5]; CA: w = f1 [C/16][C], aux0 = f2 [C][C/16];                 */
                           /*   ATTN_QK: w = wq [Cq][8], b = bq, aux0 = wk, aux1 = bk, aux2/3 = shared BN scale/shift*/
                           /*   CRISSCROSS/VERTICAL: w = wv [C], b = bv [C], aux0/1 = BN1 scale/shift [C], aux2 = packed q/k   */
                           /*   projection (C/8 x 20 fp32): wq [Cq][8] | bq | wk [Cq][8] | bk | shared BN scale | shift      */
    float fparam[8];       /* CRISSCROSS/VERTICAL: [0] = gamma;  DETECT: [0] = stride, [1..6] = anchor (w,h) x 3    */
} ry_op_desc;

/* ---- plan lifetime ------------------------------------------------------------------------------------------- */

int ry_abi_version(void);
int ry_abi_sizeof(int which);                   /* 0: sizeof(ry_tensor_desc), 1: sizeof(ry_op_desc) -- binding self-check */
const char *ry_last_error(void);            /* thread-local, valid until the next failing call on this thread */

/* Copies the op list, packs the weights (bf16, UMMA K-major order) into plan-owned device memory on `device`. */
int ry_plan_create(const ry_tensor_desc *tensors, int n_tensors, const ry_op_desc *ops, int n_ops,
                   const void *weights_host, size_t weight_bytes, int nc, int device, ry_plan **out);
void ry_plan_destroy(ry_plan *plan);

/* Bytes of caller-provided workspace (activation arena) for a [B,3,H,W] input; H, W multiples of 32. */
int ry_plan_workspace_bytes(ry_plan *plan, int B, int H, int W, size_t *bytes);
/* Lays the tensors out in `workspace`, encodes the TMA descriptors for this shape.  Not a hot call. */
int ry_plan_bind(ry_plan *plan, int B, int H, int W, void *workspace, size_t workspace_bytes);
/* Byte offset inside the bound workspace and NHWC dims of tensor `t` (for teacher-forced tests / taps). */
int ry_plan_tensor_info(ry_plan *plan, int t, size_t *offset, int *h, int *w, int *c, int *dtype);
int ry_plan_num_candidates(ry_plan *plan, int *n);   /* rows of pred per image for the bound shape (25200 @ 640x640) */

/* ---- hot calls ----------------------------------------------------------------------------------------------- */

/* image: fp32 NCHW [B,3,H,W] in [0,1];  pred: fp32 [B,N,5+nc];  raw0..2: fp32 [B,na,ny,nx,5+nc] (may be NULL). */
int ry_forward(ry_plan *plan, const float *image, float *pred, float *raw0, float *raw1, float *raw2, void *stream);
/* Runs ops [first,last) of the bound plan (externals as given; NULL allowed when the range does not touch them). */
int ry_run_ops(ry_plan *plan, int first, int last, const float *image, float *pred, float *raw0, float *raw1,
               float *raw2, void *stream);
/* Input image element type of the following ry_forward / ry_run_ops calls: RY_F32 (default, [0,1] as Model.forward takes
 * it) or RY_U8 (uint8 NCHW 0..255 as detect.py:73-78 ships it to the device; the /255 of detect.py:76 is fused in the stem).
 * With RY_U8 the `image` argument is reinterpreted as const uint8_t*. */
int ry_plan_set_image_dtype(ry_plan *plan, int dtype);
/* Number of kernel launches ry_forward issues for the bound shape (for bench.py's gpu_launches). */
int ry_plan_launch_count(ry_plan *plan, int *n);

/* Optional per-op timing for roofline reports: when on, ry_forward / ry_run_ops bracket every op with CUDA events on the
 * caller's stream; ry_plan_op_times returns the last elapsed milliseconds per op (host array of n = number of ops) after
 * the caller has synchronised the stream. */
int ry_plan_set_profiling(ry_plan *plan, int on);
int ry_plan_op_times(ry_plan *plan, float *ms_host, int n);

/* non_max_suppression.  pred: fp32 [B,N,5+nc] (not modified).  out: fp32 [B,max_det,6] rows (x1,y1,x2,y2,conf,cls)
 * in score order, counts: int32 [B].  classes: HOST array or NULL.  iou_thres is compared in double like torchvision. */
int ry_nms_workspace_bytes(int B, int N, int nc, int multi_label, size_t *bytes);
int ry_nms_launch_count(int B, int N, int nc, int multi_label, int *n);   /* kernels one ry_nms call launches (bench.py's gpu_launches) */
int ry_nms(const float *pred, int B, int N, int nc, float conf_thres, double iou_thres, const int32_t *classes_host,
           int n_classes, int agnostic, int multi_label, int max_det, int max_nms, float *out, int32_t *counts,
           void *workspace, size_t workspace_bytes, void *stream);

/* Fused decode + confidence filter (the reference decodes every candidate, yolo.py:139-156, and only then filters them,
 * general.py:961-978).  ry_decode_filter == ry_forward (pred / raw heads are materialised as always) and additionally
 * cand_mask: uint32 [B][(N + 31) / 32], zeroed by the call, bit (i & 31) of word (i >> 5) of image b = pred[b, i, 4] > conf_thres,
 * written from warp ballots over the decoded tile in the Detect epilogue.  ry_nms_filtered == ry_nms, except that candidates are
 * found by a prefix sum over the mask words (only rows that passed are read); its conf_thres must be >= the mask's.  Outputs
 * are byte-identical to ry_nms(pred, ...). */
int ry_decode_filter(ry_plan *plan, const float *image, float conf_thres, float *pred, float *raw0, float *raw1, float *raw2,
                     uint32_t *cand_mask, void *stream);
int ry_nms_filtered(const float *pred, const uint32_t *cand_mask, int B, int N, int nc, float conf_thres, double iou_thres,
                    const int32_t *classes_host, int n_classes, int agnostic, int multi_label, int max_det, int max_nms, float *out,
                    int32_t *counts, void *workspace, size_t workspace_bytes, void *stream);

/* ---- the callers either side of the path ----
 * ry_letterbox_u8  <- letterbox  utils/datasets.py:984-1014 (cv2.resize INTER_LINEAR + copyMakeBorder; the shape arithmetic --
 *                     new_w/new_h/left/top -- stays on the host, rep-yolo_b200/preproc.py) fused with the BGR->RGB, HWC->CHW
 *                     packing of LoadImages.__next__ (datasets.py:191-195).  src: uint8 HWC (3 channels) image; dst: uint8
 *                     [3][H1][W1] planes with the channel order reversed (planar_rgb = 1: one image of the uint8 NCHW batch that
 *                     ry_forward takes after ry_plan_set_image_dtype(RY_U8)) or HWC in the source order (planar_rgb = 0: what
 *                     letterbox itself returns).  pad_value3_host: HOST int32[3], per source channel.  Bit-exact with cv2.
 * ry_scale_coords  <- scale_coords + clip_coords  utils/general.py:319-340 (+ the .round() of detect.py:114 when
 *                     round_result != 0), in place on n rows of >= 4 fp32 (x1,y1,x2,y2,...), row_stride in elements;
 *                     count_dev: optional DEVICE int32 row count (e.g. one entry of ry_nms's counts), capped by n_max. */
int ry_letterbox_u8(const uint8_t *src_hwc, int H0, int W0, int src_row_bytes, uint8_t *dst, int H1, int W1, int new_w, int new_h,
                    int left, int top, const int32_t *pad_value3_host, int planar_rgb, void *stream);
int ry_scale_coords(float *coords, const int32_t *count_dev, int n_max, int row_stride, float pad_x, float pad_y, float gain, int w0,
                    int h0, int round_result, void *stream);
/* ry_nchw_to_nhwc_bf16 <- the layout IDetect.fuseforward's callers hand over (models/yolo.py:135: a list of fp32 NCHW maps) -> channels
 *                     [dst_c_off, dst_c_off + C) of an NHWC bf16 tensor with dst_channels channels per pixel (a plan tensor located
 *                     with ry_plan_tensor_info).  C, dst_c_off, dst_channels: multiples of 8. */
int ry_nchw_to_nhwc_bf16(const float *src_nchw, int B, int C, int H, int W, void *dst_nhwc_bf16, int dst_channels, int dst_c_off,
                         void *stream);

#ifdef __cplusplus
}
#endif
#endif /* REPYOLO_B200_H */
