// GPU non_max_suppression (nms.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace ry {
size_t nms_workspace_bytes(int B, int N, int nc, int multi_label);
int nms_launch_count(int B, int N, int nc, int multi_label);
int nms_run(const float *pred, int B, int N, int nc, float conf, double iou, const int32_t *classes_host, int n_classes,
            int agnostic, int multi_label, int max_det, int max_nms, float *out, int32_t *counts, void *workspace,
            size_t workspace_bytes, cudaStream_t st);
}  // namespace ry
