/*
 * CPU oracle for the post-processing half of the Rep-YOLO deploy path -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of
 *   - utils/general.py:953-1045  non_max_suppression   (conf filter, xywh->xyxy, best-class / multi-label rows,
 *                                                      class filter, max_nms cut, class-offset boxes, max_det cut)
 *   - utils/general.py:265-272   xywh2xyxy
 *   - torchvision.ops.nms (CPU kernel, call site utils/general.py:1029).  torchvision is a third-party dependency that is
 *     NOT vendored in the reference (requirements.txt:12 "torchvision>=0.8.1,!=0.13.0"); the installed 0.26.0 CPU kernel
 *     is the pin.  Its published algorithm: stable sort by score descending; areas=(x2-x1)*(y2-y1); for each
 *     unsuppressed i (in order) keep it and suppress every later j with inter/(area_i+area_j-inter) > thr, the
 *     quotient in fp32, the comparison against the *double* threshold.
 *
 * Parity status: PINNED by tests/test_oracle_golden.py against (a) torchvision.ops.nms run live on CPU and
 * (b) the reference's own non_max_suppression outputs stored in tests/golden/nms_cases.npz.
 *
 * Deliberate deviations (documented in DESIGN.md):
 *   - the 10 s time_limit break (general.py:1041-1043) is not replicated;
 *   - when n > max_nms the reference uses a non-stable argsort (general.py:1024); here the cut is the stable one
 *     (score descending, original row order on ties).
 *
 * Build: gcc -O2 -fno-fast-math -ffp-contract=off -shared -fPIC  (no FMA contraction: the arithmetic is bit-exact).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* stable merge sort of idx[0..n) by score descending (ties keep ascending idx order) */
static void sort_desc_stable(const float *score, int32_t *idx, int32_t *tmp, int n) {
    for (int w = 1; w < n; w *= 2) {
        for (int lo = 0; lo < n; lo += 2 * w) {
            int mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
            int a = lo, b = mid, o = lo;
            while (a < mid && b < hi) tmp[o++] = (score[idx[b]] > score[idx[a]]) ? idx[b++] : idx[a++];
            while (a < mid) tmp[o++] = idx[a++];
            while (b < hi) tmp[o++] = idx[b++];
        }
        memcpy(idx, tmp, (size_t)n * sizeof(int32_t));
    }
}

/* Greedy NMS over boxes[n][4] (x1,y1,x2,y2) / scores[n].  Writes kept indices (score order) into keep[], stops after
 * max_keep keeps (exact: a keep decision only depends on higher-ranked boxes).  Returns the number of keeps. */
int ry_oracle_greedy_nms(const float *boxes, const float *scores, int n, double iou_thres, int max_keep, int32_t *keep) {
    if (n <= 0) return 0;
    int32_t *order = (int32_t *)malloc(sizeof(int32_t) * (size_t)n), *tmp = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    uint8_t *dead = (uint8_t *)calloc((size_t)n, 1);
    float *area = (float *)malloc(sizeof(float) * (size_t)n);
    for (int i = 0; i < n; ++i) {
        order[i] = i;
        area[i] = (boxes[4 * i + 2] - boxes[4 * i + 0]) * (boxes[4 * i + 3] - boxes[4 * i + 1]);
    }
    sort_desc_stable(scores, order, tmp, n);
    int nk = 0;
    for (int a = 0; a < n && nk < max_keep; ++a) {
        int i = order[a];
        if (dead[i]) continue;
        keep[nk++] = i;
        const float ix1 = boxes[4 * i], iy1 = boxes[4 * i + 1], ix2 = boxes[4 * i + 2], iy2 = boxes[4 * i + 3], ia = area[i];
        for (int b = a + 1; b < n; ++b) {
            int j = order[b];
            if (dead[j]) continue;
            float xx1 = ix1 > boxes[4 * j] ? ix1 : boxes[4 * j];
            float yy1 = iy1 > boxes[4 * j + 1] ? iy1 : boxes[4 * j + 1];
            float xx2 = ix2 < boxes[4 * j + 2] ? ix2 : boxes[4 * j + 2];
            float yy2 = iy2 < boxes[4 * j + 3] ? iy2 : boxes[4 * j + 3];
            float w = xx2 - xx1, h = yy2 - yy1;
            w = w > 0.f ? w : 0.f;
            h = h > 0.f ? h : 0.f;
            float inter = w * h;
            float ovr = inter / (ia + area[j] - inter);
            if ((double)ovr > iou_thres) dead[j] = 1;
        }
    }
    free(order); free(tmp); free(dead); free(area);
    return nk;
}

/* One image of non_max_suppression.  pred: [n][5+nc] rows (cx,cy,w,h,obj,cls...).  out: [max_det][6] rows
 * (x1,y1,x2,y2,conf,cls).  Returns the number of detections. */
int ry_oracle_nms_image(const float *pred, int n, int nc, float conf_thres, double iou_thres, const int *classes,
                        int n_classes, int agnostic, int multi_label, int max_det, int max_nms, float max_wh, float *out) {
    const int no = 5 + nc;
    multi_label = multi_label && nc > 1;                                       /* general.py:970 */
    size_t cap = (size_t)n * (size_t)(multi_label ? nc : 1);
    if (cap == 0) return 0;
    float *rows = (float *)malloc(sizeof(float) * 6 * cap);
    int m = 0;
    for (int i = 0; i < n; ++i) {
        const float *p = pred + (size_t)i * no;
        if (!(p[4] > conf_thres)) continue;                                    /* general.py:962, 978 */
        float x1 = p[0] - p[2] / 2, y1 = p[1] - p[3] / 2, x2 = p[0] + p[2] / 2, y2 = p[1] + p[3] / 2; /* :265-272 */
        if (multi_label) {                                                     /* general.py:1004-1006 */
            for (int j = 0; j < nc; ++j) {
                float c = p[5 + j] * p[4];                                     /* general.py:998 */
                if (c > conf_thres) {
                    float *r = rows + 6 * (size_t)m++;
                    r[0] = x1; r[1] = y1; r[2] = x2; r[3] = y2; r[4] = c; r[5] = (float)j;
                }
            }
        } else {                                                               /* general.py:1008-1009 */
            float best = 0.f; int bj = 0;
            for (int j = 0; j < nc; ++j) {
                float c = (nc == 1) ? p[4] : p[5 + j] * p[4];                  /* general.py:994-998 */
                if (j == 0 || c > best) { best = c; bj = j; }
            }
            if (best > conf_thres) {
                float *r = rows + 6 * (size_t)m++;
                r[0] = x1; r[1] = y1; r[2] = x2; r[3] = y2; r[4] = best; r[5] = (float)bj;
            }
        }
    }
    if (classes && n_classes > 0) {                                            /* general.py:1012-1013 */
        int k = 0;
        for (int i = 0; i < m; ++i) {
            int ok = 0;
            for (int c = 0; c < n_classes; ++c) ok |= (rows[6 * (size_t)i + 5] == (float)classes[c]);
            if (ok) { if (k != i) memcpy(rows + 6 * (size_t)k, rows + 6 * (size_t)i, 6 * sizeof(float)); ++k; }
        }
        m = k;
    }
    if (m == 0) { free(rows); return 0; }
    float *scores = (float *)malloc(sizeof(float) * (size_t)m);
    int32_t *order = (int32_t *)malloc(sizeof(int32_t) * (size_t)m), *tmp = (int32_t *)malloc(sizeof(int32_t) * (size_t)m);
    for (int i = 0; i < m; ++i) { scores[i] = rows[6 * (size_t)i + 4]; order[i] = i; }
    if (m > max_nms) {                                                         /* general.py:1023-1024 (stable variant) */
        sort_desc_stable(scores, order, tmp, m);
        float *cut = (float *)malloc(sizeof(float) * 6 * (size_t)max_nms);
        for (int i = 0; i < max_nms; ++i) memcpy(cut + 6 * (size_t)i, rows + 6 * (size_t)order[i], 6 * sizeof(float));
        free(rows); rows = cut; m = max_nms;
        for (int i = 0; i < m; ++i) scores[i] = rows[6 * (size_t)i + 4];
    }
    float *boxes = (float *)malloc(sizeof(float) * 4 * (size_t)m);
    for (int i = 0; i < m; ++i) {                                              /* general.py:1027-1028 */
        float c = rows[6 * (size_t)i + 5] * (agnostic ? 0.f : max_wh);
        for (int k = 0; k < 4; ++k) boxes[4 * (size_t)i + k] = rows[6 * (size_t)i + k] + c;
    }
    int nk = ry_oracle_greedy_nms(boxes, scores, m, iou_thres, max_det, order); /* general.py:1029-1031 */
    for (int i = 0; i < nk; ++i) memcpy(out + 6 * (size_t)i, rows + 6 * (size_t)order[i], 6 * sizeof(float));
    free(rows); free(scores); free(order); free(tmp); free(boxes);
    return nk;
}

/* Batch wrapper: pred [B][n][5+nc] -> out [B][max_det][6], counts[B]. */
void ry_oracle_nms_batch(const float *pred, int B, int n, int nc, float conf_thres, double iou_thres, const int *classes,
                         int n_classes, int agnostic, int multi_label, int max_det, int max_nms, float *out, int *counts) {
    for (int b = 0; b < B; ++b)
        counts[b] = ry_oracle_nms_image(pred + (size_t)b * n * (5 + nc), n, nc, conf_thres, iou_thres, classes, n_classes,
                                        agnostic, multi_label, max_det, max_nms, 4096.f, out + (size_t)b * max_det * 6);
}
