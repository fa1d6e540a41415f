// Implicit-GEMM convolution (1x1, 3x3 s1/s2) for NHWC bf16 activations on the 5th-gen tensor cores.
//   D[pixel, cout] = sum_{tap, cin} A[pixel + tap, cin] * W[cout, tap, cin]
// One persistent CTA per SM: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer, warps 2..9 (or 2..17) = epilogue.
//
// A operand (activations), two fetch modes:
//   A_BOX   one TMA box per (tap, K block): [tn][th][tw] pixels x kb channels, shifted by the tap, zero fill outside the
//           image = conv padding.  Used by 1x1 convs (one "tap"), stride-2 3x3 (four parity-phase maps) and 3x3 on small
//           maps.  Re-reads the input once per tap through L2.
//   A_HALO  3x3 s1 on maps with W % 8 == 0: ONE TMA box per K block brings the (8+2) x (16+2) pixel halo of an
//           8 wide x 16 high output tile; the nine taps are nine tcgen05.mma descriptor windows into that tile (start
//           address shifted by (dh*10 + dw) pixels, 8-row groups 10 pixels apart).  The input crosses L2 1.4x, not 9x.
// Channel counts that are not a multiple of the K block are padded by TMA out-of-bounds zero fill (the tensor map's
// channel extent is the view length), so every cin > 32 runs 128-byte-swizzled 64-channel K blocks.
//
// B operand (weights, packed [cout_pad][tap][cblk*kb] bf16): resident in shared memory for the whole kernel when it fits
// (small-channel layers: no per-tile weight re-fetch), otherwise streamed through its own ring.
//
// Epilogue: TMEM -> registers -> +bias, SiLU (0.5x(1+tanh(0.5x)), one MUFU), residual / per-image vector add -> bf16 ->
// swizzled shared-memory staging -> TMA store into a channel range of the NHWC output (concat slots, GSConv shuffle
// halves).  N <= 128 (4 x N TMEM columns fit): four 4-warp teams take alternate tiles; wider N tiles: two column groups of
// four warps on disjoint accumulator columns.  HBM-bound 256-wide 1x1 layers are therefore run as two 128-wide N tiles;
// with round-robin tiles and gridDim % n_ntiles == 0 a CTA only ever sees one N tile, whose weights then stay resident.
// Optional fused MaxPool2d(2,2) (1x1 conv + the MP that follows a DER_Block): 2-D even pixel tiles, the 2x2 max is taken
// from the staged tile and only the pooled tile is stored.  Residual / per-image vector operands are fetched one step
// ahead; the per-image vector is staged in shared memory per image (those convs walk contiguous tile ranges).
// The Detect head (mode 1): the eight epilogue warps split the head columns, decode in registers, stage
// [anchor][pixel][no] records in shared memory and write them as contiguous 16-byte vectors.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace ry {

constexpr int kConvMaxThreads = 576;   // 2 role warps + up to 4 epilogue groups of 4 warps
constexpr int kConvMaxTaps = 9;
constexpr int kConvMaxSrcBlocks = 24;  // K blocks of a multi-source 1x1 conv
constexpr int kConvMaxSegs = 8;      // store segments per epilogue column group
constexpr int kHaloTw = 8, kHaloTh = 16;

enum { A_BOX = 0, A_HALO = 1 };

struct ConvSeg {
    int16_t col0, ncol;      // accumulator columns [col0, col0+ncol) of the N tile; ncol <= 64, multiple of 8
    int32_t chan;            // absolute channel in the output tensor (n-tile offset added at run time)
    int16_t map;             // index into omap[]
    int16_t swz;             // XOR mask of the staging layout: 7 / 3 / 1 for 128 / 64 / 32-byte rows, 0 = dense rows
};

struct ConvArgs {
    const CUtensorMap *amap;   // activation maps (1, or the 4 parity phases of a stride-2 conv)
    const CUtensorMap *wmap;   // packed weights [Cout_pad][K_pad], box {kb, BN}
    const CUtensorMap *omap;   // output store maps (box widths 64/32/16/8 channels as needed)
    int a_mode;                // A_BOX / A_HALO
    int kb;                    // channels per K block (16/32/64) == swizzle span / 2
    int cblk;                  // K blocks per tap
    int ntaps;
    int kblocks;               // ntaps * cblk
    int ksteps_last;           // K=16 MMA steps holding real channels in the last K block of a tap
    int8_t tap_map[kConvMaxTaps], tap_dh[kConvMaxTaps], tap_dw[kConvMaxTaps];
    int n_src;                 // > 1: 1x1 conv over concatenated inputs; K block i comes from map kb_map[i], channel kb_coord[i]
    int8_t kb_map[kConvMaxSrcBlocks], kb_ks[kConvMaxSrcBlocks];
    int16_t kb_coord[kConvMaxSrcBlocks];
    int tw, th, tn;            // output tile extent in w, h, image
    int tiles_w, tiles_h, tiles_n;
    uint64_t div_hw;           // ceil(2^40 / img_hw): image index of a pixel = (pix * div_hw) >> 40
    uint32_t div_imgw;         // ceil(2^32 / img_w) (0 when img_w == 1): row of a pixel inside its image (Detect grid)
    uint32_t div_nt, div_tw, div_th;   // ceil(2^32 / d) for d = n_ntiles, tiles_w, tiles_h (0 when d == 1): exact q = umulhi(n, m)
    int n_ntiles, BN;          // output-channel tiling
    int Wo, Ho, Bo;            // logical output extent the tile grid covers (1x1: Wo = B*H*W, Ho = Bo = 1)
    int img_w, img_hw;         // true W and H*W of the output map
    int halo_w;                // pixels per halo row (A_HALO)
    int a_stage_bytes, a_stages, a_box_bytes;
    int b_stage_bytes, b_stages, b_resident;
    int n_acc;                 // TMEM accumulator stages: 2 or 4 (4 * BN <= 512)
    int stage_buf_bytes;       // epilogue staging: n_groups x (2 buffers, or 1 when n_groups == 4) of this size
    int pool;                  // 1: MaxPool2d(2,2) of the activated tile fused in the epilogue (MP after DER_Block.cv1); the
                               //    staged tile is reduced 2x2 in shared memory and only the pooled tile is stored
    int mode;                  // 0 = bf16 NHWC store, 1 = Detect decode
    int ep_teams;              // 1: the 4-warp epilogue groups take alternate TILES (small N); 0: disjoint COLUMNS of each tile
    int n_groups;              // epilogue groups: 2, or 4 (tile teams of the memory-bound small-N layers)
    const float *bias;         // [Cout_pad]
    int cout, cout_pad;
    int act;
    int nseg[2];               // tile teams share list 0
    ConvSeg seg[2][kConvMaxSegs];
    const __nv_bfloat16 *res;  // optional residual (same pixel grid), added after the activation
    int res_cs, res_off;
    const float *bvec;         // optional per-image vector [B][bvec_cs], added after the activation
    int bvec_cs, bvec_off;
    int bv_bytes;              // shared-memory copy of the per-image vectors of a tile's (up to two) images: 4 groups x 2 x cout_pad fp32
    int n_img;                 // images in the batch (bound of the staged second image)
    int tile_contig;           // 1: every CTA takes one contiguous range of tiles instead of every gridDim-th tile
    int n_pinned;              // 1: gridDim % n_ntiles == 0 with round-robin tiles, i.e. a CTA only ever sees N tile blockIdx % n_ntiles,
                               //    whose weights may then stay resident in shared memory
    float *pred, *raw;         // Detect outputs
    uint32_t *cand_mask;       // Detect, optional (ry_decode_filter): `obj > cand_conf` of every decoded candidate as ballot words
    float cand_conf;           //   [B][mask_words] (bit i & 31 of word i >> 5, i = row of pred inside the image), OR-ed in
    int mask_words;
    int no, na, row_off, rows_total;
    float det_stride;
    float anchors[6];
};

size_t conv_smem_bytes(const ConvArgs &a);
// Fills a_stage_bytes / a_stages / b_* / stage_buf_bytes from the geometry already in `a`; max_seg_cols = widest store
// segment.  Returns non-zero when the shape does not fit.
int conv_plan_smem(ConvArgs &a, int max_seg_cols);
void conv_launch(const ConvArgs &a, int grid, cudaStream_t stream);

}  // namespace ry
