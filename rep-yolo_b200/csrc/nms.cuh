// GPU non_max_suppression (nms.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace ry {
size_t nms_workspace_bytes(int B, int N, int nc, int multi_label);
int nms_launch_count(int B, int N, int nc, int multi_label);
// cand_mask: NULL (every row of pred is tested) or the `obj > conf` ballot words written by the Detect epilogue
// (uint32 [B][(N + 31) / 32], ry_decode_filter): only the rows that passed are read
int nms_run(const float *pred, const uint32_t *cand_mask, int B, int N, int nc, float conf, double iou, const int32_t *classes_host,
            int n_classes, int agnostic, int multi_label, int max_det, int max_nms, float *out, int32_t *counts, void *workspace,
            size_t workspace_bytes, cudaStream_t st);
}  // namespace ry
