// Implicit-GEMM convolution (1x1, 3x3 s1/s2) for NHWC bf16 activations on the 5th-gen tensor cores.
//   D[pixel, cout] = sum_{tap, cin} A[pixel + tap, cin] * W[cout, tap, cin]
// A tiles are fetched by TMA as shifted 4-D boxes of the activation tensor (zero fill outside the image = conv padding),
// W tiles by 2-D TMA; both land in 32/64/128-byte-swizzled K-major shared memory and are consumed by tcgen05.mma with the
// fp32 accumulator in TMEM (double buffered).  One persistent CTA per SM: warp 0 = TMA producer, warp 1 = MMA issuer,
// warps 2..5 = epilogue (TMEM -> registers -> bias/SiLU/residual/broadcast-add -> bf16 NHWC at a channel offset, or the
// Detect decode).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace ry {

constexpr int kConvThreads = 192;
constexpr int kConvMaxTaps = 9;

struct ConvArgs {
    const CUtensorMap *amap;   // up to 4 activation maps (stride-2 convs use the 4 parity phases of the input)
    const CUtensorMap *wmap;   // packed weights [Cout_pad][K_pad]
    int kb;                    // channels per K block (16/32/64) == swizzle span / 2
    int cblk;                  // K blocks per tap
    int ntaps;
    int kblocks;               // ntaps * cblk
    int8_t tap_map[kConvMaxTaps], tap_dh[kConvMaxTaps], tap_dw[kConvMaxTaps];
    int tw, th, tn;            // box (tile) extent in w, h, image
    int tiles_w, tiles_h, tiles_n;
    int n_ntiles, BN;          // output-channel tiling
    int Wo, Ho, Bo;            // logical output extent the tile grid covers (1x1: Wo = B*H*W, Ho = Bo = 1)
    int img_w, img_hw;         // true W and H*W of the output map
    int stages;
    int mode;                  // 0 = bf16 NHWC store, 1 = Detect decode
    const float *bias;         // [Cout_pad]
    __nv_bfloat16 *out;
    int out_cs, off0, off1, split_at, cout;
    int act;
    const __nv_bfloat16 *res;  // optional residual (same pixel grid)
    int res_cs, res_off;
    const float *bvec;         // optional per-image vector [B][bvec_cs]
    int bvec_cs, bvec_off;
    float *pred, *raw;         // Detect outputs
    int no, na, row_off, rows_total;
    float det_stride;
    float anchors[6];
};

size_t conv_smem_bytes(int BN, int stages);
int conv_pick_stages(int BN);
void conv_launch(const ConvArgs &a, int grid, cudaStream_t stream);

}  // namespace ry
