// Fused convolution chain for the small-channel DER_Block stages (reference models/common.py:3644-3654):
//     3x3 s1 conv (+bias+SiLU)  ->  1x1 conv (+bias+act)  [->  1x1 conv (+bias+act)]
// e.g. x4_1 = cv0_2(S4(cv0_1-output)) followed by cv1_1, where the intermediate maps never leave the SM:
//   stage 0  halo-tile implicit GEMM exactly like conv_umma's A_HALO mode (resident weights, 4 TMEM accumulator slots);
//   stage s  the epilogue team turns the fp32 accumulator into bf16, writes it as the K-major swizzled A operand of the
//            next GEMM into shared memory, the MMA warp multiplies it with the resident 1x1 weights into the team's own
//            TMEM slot, and so on.  Any stage may also TMA-store its result (concat inputs, next layer's input).
// Layers this small are bound by HBM and by per-tile overheads, so removing the intermediate tensors' write + read and
// two of three kernel launches is the win; the extra MMAs are negligible.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace ry {

constexpr int kChainTeams = 5;
constexpr int kChainThreads = 64 + 128 * kChainTeams + 64;   // producer, stage-0 MMA issuer, kChainTeams x 4 warps, two 1x1-stage MMA issuers
constexpr int kChainMaxStages = 3;

struct ChainStage {
    int ncol;        // real output channels
    int N;           // MMA N: ncol padded to a multiple of 16
    int act;         // 1 = SiLU
    int store;       // 1 = TMA-store this stage's output (map index = stage index)
    int chan;        // absolute channel offset of the store in its tensor
    int swz;         // staging swizzle mask of the store (7/3/1, 0 = dense rows)
    int kb_next;     // channels per row of the next stage's A operand (32 or 64), 0 = last stage
    int tmem_col;    // accumulator base column: stage 0 = one slot of N per team; stages 1, 2 share one slot of max(N) per team
    int bias_off;    // float offset into the shared-memory bias array
    int ks;          // stages 1,2: K=16 steps (ceil(previous ncol / 16)); stage 0: K steps per tap
    int stg_off;     // byte offset of this stage's staging tile inside the team's staging area
    int w_off;       // stages 1,2: byte offset of the pre-swizzled weight image in shared memory
    int w_bytes;
    const void *w_img;   // stages 1,2: pre-swizzled [N][kb] bf16 image in global memory
    const float *bias;   // [N] fp32
};

struct ChainArgs {
    const CUtensorMap *amap;   // input halo boxes
    const CUtensorMap *wmap;   // stage-0 weights [BN][9*kb]
    const CUtensorMap *omap;   // omap[s] = store map of stage s (when stage[s].store)
    int n_stages;              // 2 or 3
    int kb;                    // stage-0 K block (32 or 64)
    int tiles_w, tiles_h, tiles_n;
    uint32_t div_tw, div_th;
    int halo_w;
    int a_stage_bytes, a_stages, a_box_bytes, b_stage_bytes;
    int anext_bytes;           // per-team A-operand buffer of the 1x1 stages
    int stage_buf_bytes;       // per-team store staging area (one tile per storing stage: back-to-back stages never wait)
    int n_store;               // number of storing stages
    int bias_floats;
    int tmem_cols;             // power of two >= all accumulator columns
    int post_stride;           // TMEM columns between the teams' (aliased) 1x1-stage accumulators
    int off_b, off_anext, off_stage, off_bias, off_bar;   // shared-memory byte offsets (from the 1 KiB-aligned base)
    ChainStage stage[kChainMaxStages];
};

size_t chain_smem_bytes(const ChainArgs &a);
int chain_plan_smem(ChainArgs &a);     // fills a_stages / a_stage_bytes / offsets from the geometry; non-zero = does not fit
void chain_launch(const ChainArgs &a, int grid, cudaStream_t stream);

}  // namespace ry
