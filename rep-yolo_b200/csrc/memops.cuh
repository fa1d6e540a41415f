// Launchers of the memory-bound kernels (memops.cu) and the attention kernels (attention.cu).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace ry {

int stem_launch(const void *img, int img_u8, const float *w27, const float *bias, __nv_bfloat16 *out, int out_cs, int out_off,
                int cout, int B, int H, int W, cudaStream_t st);
// imap: optional tensor map of the input tensor {C, W, H, B} with box {8, dw5_tile_w(W) + 4, dw5_halo_rows(), 1}, no swizzle
// (the halo planes are then fetched by TMA); NULL = per-thread cp.async staging
void dw5_launch(const __nv_bfloat16 *in, int in_cs, int in_off0, int in_off1, __nv_bfloat16 *out, int out_cs, int out_off0,
                int out_off1, const float *w, const float *bias, int C, int half, int B, int H, int W, int act,
                const CUtensorMap *imap, cudaStream_t st);
int dw5_tile_w(int W);
int dw5_halo_rows();
void maxpool2_launch(const __nv_bfloat16 *in, int in_cs, int in_off, __nv_bfloat16 *out, int out_cs, int out_off, int C,
                     int B, int H, int W, cudaStream_t st);
void spp_launch(const __nv_bfloat16 *in, int in_cs, int in_off, __nv_bfloat16 *out, int out_cs, int off5, int off9,
                int off13, int C, int B, int H, int W, cudaStream_t st);
void upsample2_launch(const __nv_bfloat16 *in, int in_cs, int in_off, __nv_bfloat16 *out, int out_cs, int out_off, int C,
                      int B, int H, int W, cudaStream_t st);
void ca_launch(const __nv_bfloat16 *in, int in_cs, int in_off, float *out, int out_cs, int out_off, const float *f1,
               const float *f2, int C, int B, int HW, float *scratch, cudaStream_t st);
size_t ca_scratch_bytes(int B, int C);   // per-split partial sums

// ---- attention (attention.cu) ----
struct AttnParams {
    const __nv_bfloat16 *x;   // input map view
    int x_cs, x_off;
    int C, Cq;                // Cq = C / 8
    int B, H, W;
    const float *qk;          // packed q/k conv parameters: wq [Cq][8] | bq [Cq] | wk [Cq][8] | bk [Cq] | shared BN scale [Cq] | shift [Cq]
    const float *wv, *bv;     // value conv (depthwise 1x1 folded with its BN): v = relu6(s1 * silu(wv*x + bv) + t1)
    const float *s1, *t1;     // stand-alone BN1 as scale/shift
    float gamma;
    __nv_bfloat16 *out;       // output map view
    int out_cs, out_off;
    float *scratch;           // row partials [B*H*W, C] bf16 | (max, sum) fp32 [B*H*W, 2] | vertical energies [B,H,W,LP] bf16
};
void attn_qk_launch(const __nv_bfloat16 *x, int x_cs, int x_off, int C, int Cq, size_t npix, const float *wq,
                    const float *bq, const float *wk, const float *bk, const float *s, const float *t, float *q, float *k,
                    cudaStream_t st);
// next: the VerticalAttention fed by this module's output (CCVA: m1(m(x))) or NULL; *energies_done = 1 when its energy pass
// ran fused in the column pass (vertical_launch(..., energies_ready = 1) then runs the value pass only)
int crisscross_launch(const AttnParams &p, const AttnParams *next, int *energies_done, cudaStream_t st);
int vertical_launch(const AttnParams &p, int energies_ready, cudaStream_t st);
size_t attn_scratch_bytes(int B, int H, int W, int C);   // row partials + statistics + energies

}  // namespace ry
