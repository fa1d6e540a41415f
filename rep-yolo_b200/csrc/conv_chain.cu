// Fused 3x3 -> 1x1 [-> 1x1] convolution chain on tcgen05 / TMEM / TMA (see conv_chain.cuh).
#include "conv_chain.cuh"

#include <algorithm>

#include "common.cuh"
#include "ptx.cuh"

namespace ry {

namespace {

constexpr int kSmemLimit = 227 * 1024;
constexpr int kBarBytes = 512;

__device__ __forceinline__ uint32_t fdiv(uint32_t n, uint32_t m) { return m ? __umulhi(n, m) : n; }

struct Tile { int w0, h0, n0; };
__device__ __forceinline__ Tile tile_of(const ChainArgs &p, int t) {
    uint32_t q = fdiv((uint32_t)t, p.div_tw);
    const int wi = t - (int)q * p.tiles_w;
    const uint32_t m = q;
    q = fdiv(m, p.div_th);
    const int hi = (int)m - (int)q * p.tiles_h;
    return {wi * 8, hi * 16, (int)q};
}

__device__ __forceinline__ void sts_v4(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(ptx::smem_u32(bar))
                 : "memory");
}

// 8 accumulator columns -> t = acc*scale + bias' (bias' pre-multiplied by scale), SiLU as h + h*tanh(h) -> 8 bf16
__device__ __forceinline__ uint4 chunk8(const uint32_t *raw, uint32_t sb_addr, float scale, bool act) {
    float x[8];
    const float4 b0 = lds_f4(sb_addr), b1 = lds_f4(sb_addr + 16);
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = fmaf(__uint_as_float(raw[i]), scale, bb[i]);
    if (act) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fmaf(x[i], ptx::tanh_approx(x[i]), x[i]);
    }
    return make_uint4(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]), pack_bf16x2(x[4], x[5]), pack_bf16x2(x[6], x[7]));
}


// One epilogue stage of one tile for one thread (= one accumulator row), fully specialised:
//   N     accumulator columns (multiple of 16),  NCOL real channels (store width),
//   KBN   channels per row of the next GEMM's A operand (0 = last stage),  STORE: also write the TMA-store staging tile.
template <int N, int NCOL, int KBN, bool STORE>
__device__ __forceinline__ void stage_body(uint32_t taddr, uint32_t sb_addr, float scale, bool act, int row, uint32_t anx_u,
                                           uint32_t stg_u) {
    constexpr int RBN = KBN * 2;
    const uint32_t arow = anx_u + (uint32_t)row * RBN;
    const uint32_t axor = RBN == 128 ? (uint32_t)(row & 7) : (uint32_t)((row >> 1) & 3);   // swizzle phase of this row
    const uint32_t srow = stg_u + (uint32_t)row * (NCOL * 2);                               // dense staging rows (NCOL = 24 / 48)
#pragma unroll
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t raw[16];
        ptx::tmem_ld16_nowait(taddr + c0, raw);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int c = c0 + 8 * j;
            const uint4 o = chunk8(raw + 8 * j, sb_addr + (uint32_t)c * 4, scale, act);
            if (KBN) sts_v4(arow + (((uint32_t)(c / 8) ^ axor) << 4), o);
            if (STORE && c < NCOL) sts_v4(srow + (uint32_t)c * 2, o);
        }
    }
    if (KBN) {
#pragma unroll
        for (int c = N; c < KBN; c += 8) sts_v4(arow + (((uint32_t)(c / 8) ^ axor) << 4), make_uint4(0, 0, 0, 0));   // K padding
    }
}

// stage "shape code" -> specialisation (the four DER_Block chains use five shapes)
__device__ __forceinline__ bool run_stage(const ChainStage &sg, uint32_t taddr, uint32_t sb_addr, int row, uint32_t anx_u, uint32_t stg_u) {
    const float scale = sg.act == 1 ? 0.5f : 1.0f;
    const bool act = sg.act == 1;
    const int code = sg.N * 1000 + sg.kb_next * 10 + sg.store;
    if (sg.swz != 0 || (sg.store && sg.ncol != (sg.N == 32 ? 24 : 48))) return false;
    switch (code) {
        case 32 * 1000 + 32 * 10 + 0: stage_body<32, 24, 32, false>(taddr, sb_addr, scale, act, row, anx_u, stg_u); return true;
        case 32 * 1000 + 32 * 10 + 1: stage_body<32, 24, 32, true>(taddr, sb_addr, scale, act, row, anx_u, stg_u); return true;
        case 32 * 1000 + 0 * 10 + 1: stage_body<32, 24, 0, true>(taddr, sb_addr, scale, act, row, anx_u, stg_u); return true;
        case 48 * 1000 + 64 * 10 + 0: stage_body<48, 48, 64, false>(taddr, sb_addr, scale, act, row, anx_u, stg_u); return true;
        case 48 * 1000 + 64 * 10 + 1: stage_body<48, 48, 64, true>(taddr, sb_addr, scale, act, row, anx_u, stg_u); return true;
        case 48 * 1000 + 0 * 10 + 1: stage_body<48, 48, 0, true>(taddr, sb_addr, scale, act, row, anx_u, stg_u); return true;
        default: return false;
    }
}

__global__ void __launch_bounds__(kChainThreads, 1) conv_chain_kernel(const __grid_constant__ ChainArgs p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sA = smem, *sB = smem + p.off_b;
    float *sbias = reinterpret_cast<float *>(smem + p.off_bias);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + p.off_bar);
    uint64_t *fullA = bars, *emptyA = fullA + 8, *tfull = emptyA + 8, *tempty = tfull + 8, *bres = tempty + 8;
    uint64_t *afull = bres + 1;              // [2][kChainTeams]: A operand of stage s+1 written by team
    uint64_t *pfull = afull + 2 * kChainTeams;   // [2][kChainTeams]: accumulator of stage s+1 complete
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(pfull + 2 * kChainTeams);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rb = p.kb * 2;
    const int total_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
    const int n_post = p.n_stages - 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < 8; ++s) { ptx::mbar_init(fullA + s, 1); ptx::mbar_init(emptyA + s, 1); }
        for (int a = 0; a < kChainTeams; ++a) { ptx::mbar_init(tfull + a, 1); ptx::mbar_init(tempty + a, 4); }
        ptx::mbar_init(bres, 1);
        for (int i = 0; i < 2 * kChainTeams; ++i) { ptx::mbar_init(afull + i, 1); ptx::mbar_init(pfull + i, 1); }
        ptx::fence_mbar_init();
        ptx::prefetch_tmap(p.amap);
        ptx::prefetch_tmap(p.wmap);
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
        ptx::tmem_relinquish();
    }
    for (int s = 0; s < p.n_stages; ++s) {
        const float scale = p.stage[s].act == 1 ? 0.5f : 1.0f;
        for (int i = threadIdx.x; i < p.stage[s].N; i += blockDim.x) sbias[p.stage[s].bias_off + i] = __ldg(p.stage[s].bias + i) * scale;
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();

    if (warp == 0) {
        // ===================== TMA producer =====================
        const bool issuer = ptx::elect_one();
        if (issuer) {
            uint32_t bytes = 9u * (uint32_t)p.b_stage_bytes;
            for (int s = 1; s < p.n_stages; ++s) bytes += (uint32_t)p.stage[s].w_bytes;
            ptx::mbar_expect_tx(bres, bytes);
            for (int tap = 0; tap < 9; ++tap) ptx::tma_load_2d(sB + (size_t)tap * p.b_stage_bytes, p.wmap, bres, tap * p.kb, 0);
            for (int s = 1; s < p.n_stages; ++s)
                bulk_load_1d(ptx::smem_u32(smem + p.stage[s].w_off), p.stage[s].w_img, (uint32_t)p.stage[s].w_bytes, bres);
        }
        int sa = 0;
        uint32_t pha = 0;
        pdl_wait();
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            const Tile tc = tile_of(p, t);
            ptx::mbar_wait(emptyA + sa, pha ^ 1);
            if (issuer) {
                ptx::mbar_expect_tx(fullA + sa, (uint32_t)p.a_box_bytes);
                ptx::tma_load_4d(sA + (size_t)sa * p.a_stage_bytes, p.amap, fullA + sa, 0, tc.w0 - 1, tc.h0 - 1, tc.n0);
            }
            if (++sa == p.a_stages) { sa = 0; pha ^= 1; }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== stage-0 MMA issuer (halo-tile 3x3) =====================
        const uint32_t issue = ptx::elect_one() ? 1u : 0u;
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint32_t idesc0 = ptx::umma_idesc_bf16(128, p.stage[0].N);
        const uint64_t a_desc0 = ptx::umma_smem_desc(ptx::smem_u32(sA), rb, p.halo_w * rb);
        const uint64_t b_desc0 = ptx::umma_smem_desc(ptx::smem_u32(sB), rb, 8 * rb);
        const uint32_t a_inc = (uint32_t)p.a_stage_bytes >> 4, b_inc = (uint32_t)p.b_stage_bytes >> 4;
        const uint32_t pix_inc = (uint32_t)rb >> 4, row_inc = (uint32_t)(p.halo_w * rb) >> 4;
        const int ks0 = p.stage[0].ks;
        int n_my = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) ++n_my;
        ptx::mbar_wait(bres, 0);
        ptx::tc_fence_after();
        int sa = 0;
        uint32_t pha = 0;
        for (int it = 0; it < n_my; ++it) {
            {
                const int acc = it % kChainTeams;                  // one stage-0 accumulator slot per team
                const uint32_t aph = (uint32_t)(it / kChainTeams) & 1u;
                ptx::mbar_wait(tempty + acc, aph ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_u + p.stage[0].tmem_col + acc * p.stage[0].N;
                ptx::mbar_wait(fullA + sa, pha);
                ptx::tc_fence_after();
                uint64_t a_row = a_desc0 + (uint64_t)(sa * a_inc);
                uint64_t b_desc = b_desc0;
                uint32_t accumulate = 0;
#pragma unroll 1
                for (int th = 0; th < 3; ++th) {
#pragma unroll
                    for (int tw = 0; tw < 3; ++tw) {
                        const uint64_t a_t = a_row + (uint64_t)(tw * pix_inc);
                        ptx::umma_bf16_d64_first(d_tmem, a_t, b_desc, idesc0, issue, accumulate);
                        if (ks0 > 1) ptx::umma_bf16_d64<2, 1>(d_tmem, a_t, b_desc, idesc0, issue);
                        if (ks0 > 2) ptx::umma_bf16_d64<4, 1>(d_tmem, a_t, b_desc, idesc0, issue);
                        if (ks0 > 3) ptx::umma_bf16_d64<6, 1>(d_tmem, a_t, b_desc, idesc0, issue);
                        accumulate = 1;
                        b_desc += b_inc;
                    }
                    a_row += row_inc;
                }
                ptx::umma_commit_pred(emptyA + sa, issue);
                if (++sa == p.a_stages) { sa = 0; pha ^= 1; }
                ptx::umma_commit_pred(tfull + acc, issue);
            }
        }
        __syncwarp();
    } else if (warp >= 2 + 4 * kChainTeams) {
        // ===================== 1x1-stage MMA issuers: one warp per fused stage, tiles in order =====================
        // A separate issuer per stage means a stage's GEMM is launched the moment its operand is ready, independent of where
        // the stage-0 main loop is: the per-tile latency of the chain (what the four teams' throughput hangs on) stays short.
        const int s = warp - (2 + 4 * kChainTeams) + 1;
        if (s <= n_post) {
            const uint32_t issue = ptx::elect_one() ? 1u : 0u;
            const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
            int n_my = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) ++n_my;
            const int rbn = p.stage[s - 1].kb_next * 2;
            const uint64_t a_desc0 = ptx::umma_smem_desc(ptx::smem_u32(smem + p.off_anext), rbn, 8 * rbn);
            const uint64_t w_desc = ptx::umma_smem_desc(ptx::smem_u32(smem + p.stage[s].w_off), rbn, 8 * rbn);
            const uint32_t idesc = ptx::umma_idesc_bf16(128, p.stage[s].N);
            const uint32_t a_inc = (uint32_t)p.anext_bytes >> 4;
            const int ks = p.stage[s].ks;
            ptx::mbar_wait(bres, 0);
            ptx::tc_fence_after();
            for (int j = 0; j < n_my; ++j) {
                const int team = j % kChainTeams;
                const uint32_t ph = (uint32_t)(j / kChainTeams) & 1u;
                ptx::mbar_wait(afull + (s - 1) * kChainTeams + team, ph);
                ptx::tc_fence_after();
                const uint64_t a_desc = a_desc0 + (uint64_t)(team * a_inc);
                const uint32_t d_tmem = tmem_u + p.stage[s].tmem_col + team * p.post_stride;
                ptx::umma_bf16_d64_first(d_tmem, a_desc, w_desc, idesc, issue, 0u);
                if (ks > 1) ptx::umma_bf16_d64<2, 1>(d_tmem, a_desc, w_desc, idesc, issue);
                if (ks > 2) ptx::umma_bf16_d64<4, 1>(d_tmem, a_desc, w_desc, idesc, issue);
                if (ks > 3) ptx::umma_bf16_d64<6, 1>(d_tmem, a_desc, w_desc, idesc, issue);
                ptx::umma_commit_pred(pfull + (s - 1) * kChainTeams + team, issue);
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue teams: team g takes tiles g, g+4, ... and walks them through all stages ========
        const int e = warp - 2, team = e >> 2, quarter = warp & 3;
        const bool leader = (e & 3) == 0 && lane == 0;
        const int row = quarter * 32 + lane;
        const uint32_t sbias_u = ptx::smem_u32(sbias);
        const uint32_t stg_u = ptx::smem_u32(smem + p.off_stage + team * p.stage_buf_bytes);
        const uint32_t anx_u = ptx::smem_u32(smem + p.off_anext + team * p.anext_bytes);
        const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
        pdl_wait();
        int it = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
            if (it % kChainTeams != team) continue;
            const Tile tc = tile_of(p, t);
            const uint32_t ph = (uint32_t)(it / kChainTeams) & 1u;
            for (int s = 0; s < p.n_stages; ++s) {
                const ChainStage &sg = p.stage[s];
                uint32_t taddr;
                if (s == 0) {
                    const int acc = team;
                    ptx::mbar_wait(tfull + acc, ph);
                    taddr = tmem_base + sg.tmem_col + acc * sg.N + lane_sel;
                } else {
                    ptx::mbar_wait(pfull + (s - 1) * kChainTeams + team, ph);
                    taddr = tmem_base + sg.tmem_col + team * p.post_stride + lane_sel;
                }
                ptx::tc_fence_after();
                if (sg.store) {
                    if (leader) {                                  // the previous tile's store of this stage has left its staging tile
                        if (p.n_store >= 2) ptx::bulk_wait_read<1>(); else ptx::bulk_wait_read<0>();
                    }
                    ptx::bar_sync(1 + team, 128);
                }
                const uint32_t stg_s = stg_u + (uint32_t)sg.stg_off;
                if (!run_stage(sg, taddr, sbias_u + (uint32_t)sg.bias_off * 4, row, anx_u, stg_s)) {
                    // generic path (any channel counts <= 64)
                    const float scale = sg.act == 1 ? 0.5f : 1.0f;
                    const bool act = sg.act == 1;
                    const int rbn = sg.kb_next * 2;
                    const uint32_t nmask = rbn == 128 ? 7u : 3u;
                    const uint32_t srow = (uint32_t)row * (uint32_t)(sg.ncol * 2);
                    for (int c0 = 0; c0 < sg.N; c0 += 16) {
                        uint32_t raw[16];
                        ptx::tmem_ld16_nowait(taddr + c0, raw);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const int c = c0 + 8 * j;
                            const uint4 o = chunk8(raw + 8 * j, sbias_u + (uint32_t)(sg.bias_off + c) * 4, scale, act);
                            if (sg.kb_next) {                      // A operand of the next GEMM (K-major, hardware swizzle pattern)
                                uint32_t lin = (uint32_t)row * (uint32_t)rbn + (uint32_t)c * 2;
                                lin ^= ((lin >> 7) & nmask) << 4;
                                sts_v4(anx_u + lin, o);
                            }
                            if (sg.store && c < sg.ncol) {
                                uint32_t lin = srow + (uint32_t)c * 2;
                                lin ^= ((lin >> 7) & (uint32_t)sg.swz) << 4;
                                sts_v4(stg_s + lin, o);
                            }
                        }
                    }
                    if (sg.kb_next)                                // zero the K padding of the next operand (N < kb_next)
                        for (int c = sg.N; c < sg.kb_next; c += 8) {
                            uint32_t lin = (uint32_t)row * (uint32_t)rbn + (uint32_t)c * 2;
                            lin ^= ((lin >> 7) & nmask) << 4;
                            sts_v4(anx_u + lin, make_uint4(0, 0, 0, 0));
                        }
                }
                ptx::tc_fence_before();
                if (s == 0) {
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(tempty + team);
                }
                ptx::fence_proxy_async();
                ptx::bar_sync(1 + team, 128);
                if (leader) {
                    if (sg.store) {
                        ptx::tma_store_4d(p.omap + s, reinterpret_cast<const void *>(smem + p.off_stage + team * p.stage_buf_bytes + sg.stg_off), sg.chan,
                                          tc.w0, tc.h0, tc.n0);
                        ptx::bulk_commit();
                    }
                    if (sg.kb_next) ptx::mbar_arrive(afull + s * kChainTeams + team);
                }
            }
        }
        if (leader) ptx::bulk_wait_read<0>();
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

}  // namespace

size_t chain_smem_bytes(const ChainArgs &a) { return (size_t)a.off_bar + kBarBytes + 1024; }

int chain_plan_smem(ChainArgs &a) {
    const int rb = a.kb * 2;
    a.a_box_bytes = a.halo_w * 18 * rb;
    a.a_stage_bytes = (a.a_box_bytes + 1023) & ~1023;
    a.b_stage_bytes = a.stage[0].N * rb;
    int cols = kChainTeams * a.stage[0].N;
    a.stage[0].tmem_col = 0;
    a.anext_bytes = 0;
    a.stage_buf_bytes = 0;
    a.n_store = 0;
    a.bias_floats = 0;
    int post_n = 0;
    for (int s = 0; s < a.n_stages; ++s) {
        if (s > 0) { a.stage[s].tmem_col = cols; post_n = std::max(post_n, a.stage[s].N); }   // a team's 1x1 accumulators alias: stage s+1 is
                                                                                               // issued only after the team has read stage s out
        if (a.stage[s].kb_next) a.anext_bytes = std::max(a.anext_bytes, 128 * a.stage[s].kb_next * 2);
        if (a.stage[s].store) {
            a.stage[s].stg_off = a.stage_buf_bytes;
            a.stage_buf_bytes += (128 * a.stage[s].ncol * 2 + 1023) & ~1023;
            ++a.n_store;
        }
        a.stage[s].bias_off = a.bias_floats;
        a.bias_floats += a.stage[s].N;
    }
    a.post_stride = post_n;
    cols += kChainTeams * post_n;
    if (cols > 512) return 1;
    a.tmem_cols = 32;
    while (a.tmem_cols < cols) a.tmem_cols <<= 1;
    long fixed = 9L * a.b_stage_bytes;
    for (int s = 1; s < a.n_stages; ++s) fixed += (a.stage[s].w_bytes + 1023) & ~1023;
    fixed += (long)kChainTeams * (a.anext_bytes + a.stage_buf_bytes) + ((a.bias_floats * 4 + 127) & ~127) + kBarBytes + 1024;
    const long avail = kSmemLimit - fixed;
    a.a_stages = (int)std::min<long>(6, avail / a.a_stage_bytes);
    if (a.a_stages < 2) return 1;
    int off = a.a_stages * a.a_stage_bytes;
    a.off_b = off; off += 9 * a.b_stage_bytes;
    for (int s = 1; s < a.n_stages; ++s) { a.stage[s].w_off = off; off += (a.stage[s].w_bytes + 1023) & ~1023; }
    a.off_anext = off; off += kChainTeams * a.anext_bytes;
    a.off_stage = off; off += kChainTeams * a.stage_buf_bytes;
    a.off_bias = off; off += (a.bias_floats * 4 + 127) & ~127;
    a.off_bar = off;
    return 0;
}

void chain_launch(const ChainArgs &a, int grid, cudaStream_t stream) {
    RY_ENSURE_DYN_SMEM(conv_chain_kernel, kSmemLimit);
    launch_pdl(conv_chain_kernel, dim3(grid), dim3(kChainThreads), chain_smem_bytes(a), stream, a);
}

}  // namespace ry
