// tcgen05 / TMEM / TMA implicit-GEMM convolution kernel (see conv_umma.cuh for the design).
#include "conv_umma.cuh"

#include <algorithm>

#include "common.cuh"
#include "ptx.cuh"

namespace ry {

namespace {

constexpr int kSmemLimit = 227 * 1024;
constexpr int kBarrierBytes = 512;

struct TileCoord {
    int w0, h0, n0, nc0;
};

// n / d through the precomputed m = ceil(2^32 / d): exact while n * d < 2^32 (tile counts are < 2^20, d < 2^12)
__device__ __forceinline__ uint32_t fast_div(uint32_t n, uint32_t m) { return m ? __umulhi(n, m) : n; }

__device__ __forceinline__ TileCoord tile_coord(const ConvArgs &p, int t) {
    uint32_t m = fast_div((uint32_t)t, p.div_nt);
    const int nt = t - (int)m * p.n_ntiles;
    uint32_t q = fast_div(m, p.div_tw);
    const int wi = (int)m - (int)q * p.tiles_w;
    m = q;
    q = fast_div(m, p.div_th);
    const int hi = (int)m - (int)q * p.tiles_h;
    return {wi * p.tw, hi * p.th, (int)q * p.tn, nt * p.BN};
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ float ld_shared_f(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ float4 ld_shared_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

// Residual / per-image vector operands of one 16-column epilogue step, fetched one step AHEAD of their use (the loads are
// L1/L2-latency bound; issued back to back they overlap the TMEM load and the math of the previous step).
struct EpiExtra {
    uint4 r[2];      // residual: 2 x 8 bf16
    float4 v[4];     // per-image vector: 16 fp32
};

// EXTRA: bit 0 = residual, bit 1 = per-image vector (compile time, so the unused operand costs no registers)
// sbv_row: shared-memory address of the staged per-image vector of this row's image (0: not staged, read global memory)
template <int EXTRA>
__device__ __forceinline__ void epi_fetch(const ConvArgs &p, EpiExtra &x, bool valid, size_t pix, int img, int ch, int halves,
                                          uint32_t sbv_row) {
    if (!EXTRA) return;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        if (EXTRA & 1) x.r[j] = make_uint4(0, 0, 0, 0);
        if (EXTRA & 2) x.v[2 * j] = x.v[2 * j + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid && j < halves) {
            if (EXTRA & 1) x.r[j] = __ldg(reinterpret_cast<const uint4 *>(p.res + pix * p.res_cs + p.res_off + ch + 8 * j));
            if (EXTRA & 2) {
                if (sbv_row) {
                    x.v[2 * j] = ld_shared_f4(sbv_row + (uint32_t)(ch + 8 * j) * 4);
                    x.v[2 * j + 1] = ld_shared_f4(sbv_row + (uint32_t)(ch + 8 * j) * 4 + 16);
                } else {
                    const float4 *bv = reinterpret_cast<const float4 *>(p.bvec + (size_t)img * p.bvec_cs + p.bvec_off + ch + 8 * j);
                    x.v[2 * j] = __ldg(bv);
                    x.v[2 * j + 1] = __ldg(bv + 1);
                }
            }
        }
    }
}

// 8 accumulator columns -> bias, activation, optional residual / per-image vector -> 8 bf16 (one 16-byte chunk).
//   t = acc*scale + sbias (sbias is pre-multiplied by scale: 0.5 for SiLU, 1 for identity)
//   SiLU(x) = x*sigmoid(x) = h + h*tanh(h), h = x/2   (one MUFU.TANH instead of EX2 + RCP)
// EXTRA = false compiles the residual / broadcast-add paths out (predicated-off instructions still cost issue slots).
template <int EXTRA>
__device__ __forceinline__ uint4 epi_chunk8(const uint32_t *raw, uint32_t sb_addr, float scale, bool act, const EpiExtra &ex, int j) {
    float x[8];
    const float4 b0 = ld_shared_f4(sb_addr), b1 = ld_shared_f4(sb_addr + 16);
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = fmaf(__uint_as_float(raw[i]), scale, bb[i]);
    if (act) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fmaf(x[i], ptx::tanh_approx(x[i]), x[i]);
    }
    if (EXTRA & 1) {
        const uint32_t rw[4] = {ex.r[j].x, ex.r[j].y, ex.r[j].z, ex.r[j].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = unpack_bf16x2(rw[i]);
            x[2 * i] += f.x;
            x[2 * i + 1] += f.y;
        }
    }
    if (EXTRA & 2) {
        const float4 v0 = ex.v[2 * j], v1 = ex.v[2 * j + 1];
        x[0] += v0.x; x[1] += v0.y; x[2] += v0.z; x[3] += v0.w;
        x[4] += v1.x; x[5] += v1.y; x[6] += v1.z; x[7] += v1.w;
    }
    return make_uint4(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]), pack_bf16x2(x[4], x[5]), pack_bf16x2(x[6], x[7]));
}

// IDetect.fuseforward decode (reference models/yolo.py:139-156): same op order, fp32.
__device__ __forceinline__ float detect_decode(const ConvArgs &p, float t, int a, int o, int gx, int gy) {
    const float s = __fdividef(1.0f, 1.0f + __expf(-t));        // |error| ~1e-7, far inside the stated decode tolerance
    if (o == 0) return (s * 2.0f - 0.5f + (float)gx) * p.det_stride;
    if (o == 1) return (s * 2.0f - 0.5f + (float)gy) * p.det_stride;
    if (o == 2) { const float u = s * 2.0f; return u * u * p.anchors[2 * a]; }
    if (o == 3) { const float u = s * 2.0f; return u * u * p.anchors[2 * a + 1]; }
    return s;
}


struct MmaCtx {
    uint64_t *fullA, *emptyA, *fullB, *emptyB, *tfull, *tempty, *bres;
    uint32_t sA_u, sB_u, tmem_base;
    int t_begin, t_end, t_step, n_acc;
};

// One tap / K block: KS K=16 steps into the same accumulator; the first step takes the run-time accumulate flag.
template <int KS>
__device__ __forceinline__ void mma_block(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t issue,
                                          uint32_t first_acc, int ks) {
    ptx::umma_bf16_d64_first(d_tmem, a_desc, b_desc, idesc, issue, first_acc);
    if (KS ? KS > 1 : ks > 1) ptx::umma_bf16_d64<2, 1>(d_tmem, a_desc, b_desc, idesc, issue);
    if (KS ? KS > 2 : ks > 2) ptx::umma_bf16_d64<4, 1>(d_tmem, a_desc, b_desc, idesc, issue);
    if (KS ? KS > 3 : ks > 3) ptx::umma_bf16_d64<6, 1>(d_tmem, a_desc, b_desc, idesc, issue);
}

// MMA issuer role: the whole warp stays in uniform control flow, one elected lane issues (predicated statements).
// HALO / STREAM_B / KS are compile-time so that the per-tap code is a handful of uniform instructions (the issue rate of
// this single warp bounds small-N layers).  KS = 0: K steps per block decided at run time (multi-chunk layers).
template <bool HALO, bool STREAM_B, int KS>
__device__ __forceinline__ void mma_role(const ConvArgs &p, const MmaCtx &cx) {
    const uint32_t issue = ptx::elect_one() ? 1u : 0u;
    const int rb = p.kb * 2;
    const uint32_t idesc = ptx::umma_idesc_bf16(128, p.BN);
    const uint32_t a_sbo = (HALO ? p.halo_w : 8) * rb;
    const uint64_t a_desc0 = ptx::umma_smem_desc(cx.sA_u, rb, a_sbo);
    const uint64_t b_desc0 = ptx::umma_smem_desc(cx.sB_u, rb, 8 * rb);
    const uint32_t a_inc = (uint32_t)p.a_stage_bytes >> 4, b_inc = (uint32_t)p.b_stage_bytes >> 4;
    const uint32_t pix_inc = (uint32_t)rb >> 4;                     // one halo pixel
    const uint32_t row_inc = (uint32_t)(p.halo_w * rb) >> 4;        // one halo row
    const int n_a = HALO ? p.cblk : p.kblocks;
    const int ksteps_full = p.kb / 16;
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, cx.tmem_base, 0);
    if (!STREAM_B) {
        ptx::mbar_wait(cx.bres, 0);
        ptx::tc_fence_after();
    }
    int sa = 0, sb = 0, acc = 0;
    uint32_t pha = 0, phb = 0, aph = 0;
    for (int t = cx.t_begin; t < cx.t_end; t += cx.t_step) {
        ptx::mbar_wait(cx.tempty + acc, aph ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_u + acc * p.BN;
        uint32_t accumulate = 0;
        int c = 0;
        for (int ia = 0; ia < n_a; ++ia) {
            const int ks = KS ? KS : (p.n_src > 1 ? (int)p.kb_ks[ia] : ((c == p.cblk - 1) ? p.ksteps_last : ksteps_full));
            ptx::mbar_wait(cx.fullA + sa, pha);
            ptx::tc_fence_after();
            const uint64_t a_desc = a_desc0 + (uint64_t)(sa * a_inc);
            if (HALO) {
                uint64_t a_row = a_desc;
                uint32_t b_idx = c;
#pragma unroll 1
                for (int th = 0; th < 3; ++th) {
#pragma unroll
                    for (int tw = 0; tw < 3; ++tw) {
                        uint64_t b_desc;
                        if (STREAM_B) {
                            ptx::mbar_wait(cx.fullB + sb, phb);
                            ptx::tc_fence_after();
                            b_desc = b_desc0 + (uint64_t)(sb * b_inc);
                        } else {
                            b_desc = b_desc0 + (uint64_t)(b_idx * b_inc);
                        }
                        mma_block<KS>(d_tmem, a_row + (uint64_t)(tw * pix_inc), b_desc, idesc, issue, accumulate, ks);
                        accumulate = 1;
                        if (STREAM_B) {
                            ptx::umma_commit_pred(cx.emptyB + sb, issue);
                            if (++sb == p.b_stages) { sb = 0; phb ^= 1; }
                        }
                        b_idx += p.cblk;
                    }
                    a_row += row_inc;
                }
            } else {
                uint64_t b_desc;
                if (STREAM_B) {
                    ptx::mbar_wait(cx.fullB + sb, phb);
                    ptx::tc_fence_after();
                    b_desc = b_desc0 + (uint64_t)(sb * b_inc);
                } else {
                    b_desc = b_desc0 + (uint64_t)(ia * b_inc);
                }
                mma_block<KS>(d_tmem, a_desc, b_desc, idesc, issue, accumulate, ks);
                accumulate = 1;
                if (STREAM_B) {
                    ptx::umma_commit_pred(cx.emptyB + sb, issue);
                    if (++sb == p.b_stages) { sb = 0; phb ^= 1; }
                }
            }
            ptx::umma_commit_pred(cx.emptyA + sa, issue);    // frees the A stage once these MMAs retire
            if (++sa == p.a_stages) { sa = 0; pha ^= 1; }
            if (++c == p.cblk) c = 0;
        }
        ptx::umma_commit_pred(cx.tfull + acc, issue);        // accumulator ready for the epilogue
        if (++acc == cx.n_acc) { acc = 0; aph ^= 1; }
    }
}

template <bool HALO, bool STREAM_B>
__device__ __forceinline__ void mma_role_ks(const ConvArgs &p, const MmaCtx &cx) {
    int ks = p.cblk == 1 ? p.ksteps_last : (p.ksteps_last == p.kb / 16 ? p.kb / 16 : 0);   // uniform K steps?
    if (p.n_src > 1) ks = 0;
    switch (ks) {
        case 2: mma_role<HALO, STREAM_B, 2>(p, cx); break;
        case 3: mma_role<HALO, STREAM_B, 3>(p, cx); break;
        case 4: mma_role<HALO, STREAM_B, 4>(p, cx); break;
        default: mma_role<HALO, STREAM_B, 0>(p, cx); break;
    }
}


struct EpiCtx {
    uint64_t *tfull, *tempty;
    uint32_t tmem_base, stage_u, sbias_u;
    uint8_t *sStage;
    float *sbv;
    int t_begin, t_end, t_step, n_acc, warp, lane;
};

// Epilogue role (mode 0): 2 groups x 4 warps (TMEM lane quarter = warp % 4).  ep_teams: the groups take alternate tiles,
// otherwise disjoint column segments of every tile.  TMEM -> registers -> bias/act(/extras) -> bf16 -> swizzled staging ->
// TMA store (two staging buffers per group; the leader lane tracks the bulk groups).
template <int EXTRA>
__device__ __forceinline__ void epilogue_store_role(const ConvArgs &p, const EpiCtx &cx) {
    const int e = cx.warp - 2, lane = cx.lane;
    const int grp = e >> 2;
    const int quarter = cx.warp & 3;
    const bool leader = (e & 3) == 0 && lane == 0;
    const int row = quarter * 32 + lane;
    const float scale = p.act == 1 ? 0.5f : 1.0f;
    const bool act = p.act == 1;
    int w_in = 0, h_in = 0, n_in = 0;
    if (EXTRA) { w_in = row % p.tw; h_in = (row / p.tw) % p.th; n_in = row / (p.tw * p.th); }
    const int nbuf = p.n_groups == 4 ? 1 : 2;                 // staging buffers per group
    const int lst = p.ep_teams ? 0 : grp;                     // segment list: teams share list 0
    const uint32_t stage_u = cx.stage_u + grp * nbuf * p.stage_buf_bytes;
    const int n_acc = cx.n_acc;
    const int gmask = p.n_groups - 1;
    int bufsel = 0, it = 0;
    int bv_img = -1;                                          // image whose vector (and its successor's) is staged
    float *sbv = cx.sbv + (size_t)grp * 2 * p.cout_pad;
    const uint32_t sbv_u = ptx::smem_u32(sbv);
    for (int t = cx.t_begin; t < cx.t_end; t += cx.t_step, ++it) {
        if (p.ep_teams && (it & gmask) != grp) continue;      // tile teams
        const int acc = n_acc == 4 ? (it & 3) : (it & 1);
        const uint32_t aph = (uint32_t)(n_acc == 4 ? (it >> 2) : (it >> 1)) & 1u;
        const TileCoord tc = tile_coord(p, t);
        bool valid = false;
        size_t pix = 0;
        int img = 0;
        if (EXTRA) {                                          // pixel coordinates: residual / broadcast add only
            const int w = tc.w0 + w_in, h = tc.h0 + h_in, n = tc.n0 + n_in;
            valid = (n_in < p.tn) && (w < p.Wo) && (h < p.Ho) && (n < p.Bo);
            const uint32_t pix32 = ((uint32_t)n * (uint32_t)p.Ho + (uint32_t)h) * (uint32_t)p.Wo + (uint32_t)w;   // < 2^31 pixels per batch
            pix = pix32;
            img = (int)(((uint64_t)pix32 * p.div_hw) >> 40);
        }
        uint32_t sbv_row = 0;
        if (EXTRA & 2) {
            // per-image vectors of the tile's first image and the next one staged in shared memory (the operand is the same
            // for all rows of an image; with ~200 KB of shared memory in use there is hardly any L1 left to catch the re-reads)
            const uint32_t pix0 = ((uint32_t)tc.n0 * (uint32_t)p.Ho + (uint32_t)tc.h0) * (uint32_t)p.Wo + (uint32_t)tc.w0;
            const int img0 = (int)(((uint64_t)pix0 * p.div_hw) >> 40);
            if (img0 != bv_img) {
                ptx::bar_sync(1 + grp, 128);                      // nobody still reads the previous pair
                for (int i = (e & 3) * 32 + lane; i < 2 * p.cout_pad; i += 128) {
                    const int second = i >= p.cout_pad ? 1 : 0, c = i - second * p.cout_pad;
                    sbv[i] = (c < p.cout && img0 + second < p.n_img) ? __ldg(p.bvec + (size_t)(img0 + second) * p.bvec_cs + p.bvec_off + c) : 0.0f;
                }
                ptx::bar_sync(1 + grp, 128);
                bv_img = img0;
            }
            const int dimg = img - img0;
            if (dimg == 0 || dimg == 1) sbv_row = sbv_u + (uint32_t)(dimg * p.cout_pad) * 4;
        }
        ptx::mbar_wait(cx.tfull + acc, aph);
        ptx::tc_fence_after();
        const uint32_t taddr = cx.tmem_base + acc * p.BN + ((uint32_t)(quarter * 32) << 16);
        const uint32_t sb_tile = cx.sbias_u + (uint32_t)tc.nc0 * 4;
        for (int si = 0; si < p.nseg[lst]; ++si) {
            const ConvSeg sg = p.seg[lst][si];
            const uint32_t buf = stage_u + bufsel * p.stage_buf_bytes;
            if (leader) {                                     // the store that last read this buffer is done
                if (nbuf == 2) ptx::bulk_wait_read<1>(); else ptx::bulk_wait_read<0>();
            }
            ptx::bar_sync(1 + grp, 128);
            const uint32_t rowoff = (uint32_t)row * (uint32_t)(sg.ncol * 2);
            const uint32_t swz = (uint32_t)sg.swz;
            EpiExtra ex_cur, ex_nxt;
            epi_fetch<EXTRA>(p, ex_cur, valid, pix, img, tc.nc0 + sg.col0, sg.ncol >= 16 ? 2 : 1, sbv_row);
            for (int c0 = 0; c0 < sg.ncol; c0 += 16) {
                uint32_t raw[16];
                const int col = sg.col0 + c0;
                const int halves = sg.ncol - c0 >= 16 ? 2 : 1;
                if (halves == 2) ptx::tmem_ld16_nowait(taddr + col, raw); else ptx::tmem_ld8_nowait(taddr + col, raw);
                if (EXTRA && c0 + 16 < sg.ncol) epi_fetch<EXTRA>(p, ex_nxt, valid, pix, img, tc.nc0 + col + 16, sg.ncol - c0 - 16 >= 16 ? 2 : 1, sbv_row);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    if (j < halves) {
                        const uint4 o = epi_chunk8<EXTRA>(raw + 8 * j, sb_tile + (uint32_t)(col + 8 * j) * 4, scale, act, ex_cur, j);
                        uint32_t lin = rowoff + (uint32_t)(c0 * 2 + j * 16);
                        lin ^= ((lin >> 7) & swz) << 4;
                        st_shared_v4(buf + lin, o);
                    }
                }
                if (EXTRA) ex_cur = ex_nxt;
            }
            if (p.pool) {
                // fused MaxPool2d(2, 2): the 128-pixel tile is tw x th (both even); 2x2 windows reduced from the staged tile into
                // a 32-pixel pooled tile behind it (same row width / swizzle), which is what the TMA store sends out
                ptx::bar_sync(1 + grp, 128);
                const int cpr = sg.ncol >> 3, rbs = sg.ncol * 2, hw = p.tw >> 1, hh = p.th >> 1;
                const uint32_t pbuf = buf + 128u * (uint32_t)rbs;
                for (int idx = (e & 3) * 32 + lane; idx < hw * hh * p.tn * cpr; idx += 128) {
                    const int prow = idx / cpr, ch = idx - prow * cpr;
                    const int pyn = prow / hw, px = prow - pyn * hw;          // pyn = pn * hh + py: source row 2 * pyn of the tile
                    const int r0 = 2 * pyn * p.tw + 2 * px;
                    uint4 m = make_uint4(0, 0, 0, 0);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int r = r0 + (q & 1) + (q >> 1) * p.tw;
                        uint32_t lin = (uint32_t)(r * rbs + ch * 16);
                        lin ^= ((lin >> 7) & swz) << 4;
                        uint4 v;
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(buf + lin));
                        m = q == 0 ? v : bf16x8_max(m, v);
                    }
                    uint32_t lo = (uint32_t)(prow * rbs + ch * 16);
                    lo ^= ((lo >> 7) & swz) << 4;
                    st_shared_v4(pbuf + lo, m);
                }
            }
            ptx::fence_proxy_async();
            ptx::bar_sync(1 + grp, 128);
            if (leader) {
                const uint8_t *src = cx.sStage + (size_t)(grp * nbuf + bufsel) * p.stage_buf_bytes;
                if (p.pool)
                    ptx::tma_store_4d(p.omap + sg.map, reinterpret_cast<const void *>(src + 128 * sg.ncol * 2), sg.chan + tc.nc0, tc.w0 >> 1,
                                      tc.h0 >> 1, tc.n0);
                else
                    ptx::tma_store_4d(p.omap + sg.map, reinterpret_cast<const void *>(src), sg.chan + tc.nc0, tc.w0, tc.h0, tc.n0);
                ptx::bulk_commit();
            }
            if (nbuf == 2) bufsel ^= 1;
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(cx.tempty + acc);
    }
    if (leader) ptx::bulk_wait_read<0>();                     // staging must outlive the last TMA store's read
}

// Epilogue role (mode 1): Detect decode.  The 8 epilogue warps split the na*no (= 18) head columns evenly (group g: columns
// [9g, 9g+9)), decode in registers and write the tile as [anchor][pixel][no] fp32 records into shared memory (decoded and
// raw); then all 256 threads stream the staged tile out: the 128 pixels of one anchor are ONE contiguous 128*no*4-byte run
// of `pred` (and of the raw head tensor), written as 16-byte vectors (scalar but still contiguous when the tile straddles
// two images or the end of the batch).  Two staging buffers: one 256-thread barrier per tile.
// NO6 = true: the Rep-YOLO head (na = 3, no = 6) with every index computation on compile-time constants.
constexpr int kDetHalf = 9;
constexpr int kDetHalfMax = 16;      // generic heads: na*no <= 32 columns, ceil(nn / 2) per group
template <bool NO6>
__device__ __forceinline__ void epilogue_detect_role(const ConvArgs &p, const EpiCtx &cx) {
    const int e = cx.warp - 2, lane = cx.lane;
    const int grp = e >> 2;
    const int quarter = cx.warp & 3;
    const int row = quarter * 32 + lane;
    const int tid = e * 32 + lane;                               // 0..255
    const int n_acc = cx.n_acc;
    const int no = NO6 ? 6 : p.no, na = NO6 ? 3 : p.na, nn = na * no;
    const int rec = 128 * no;                                    // floats per anchor per tile
    const int tile_f = na * rec;                                 // floats per staged tile (one of decoded / raw)
    const int half = NO6 ? kDetHalf : (nn + 1) / 2;              // columns per group (generic head: any nn <= 32)
    const int col0 = grp * half;
    // NO6: one 16-column TMEM load window [ldcol, ldcol + 16) covers the group's 9 columns; generic: both 16-column halves
    const int ldcol = NO6 ? (grp ? kDetHalf - 1 : 0) : 0;
    float *stage = reinterpret_cast<float *>(cx.sStage);
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(p.pred) | reinterpret_cast<uintptr_t>(p.raw)) & 15) == 0 && (rec % 4) == 0 &&
                        ((p.img_hw * no) % 4) == 0 && (((size_t)p.rows_total * no) % 4) == 0 && (((size_t)p.row_off * no) % 4) == 0;
    int it = 0;
    for (int t = cx.t_begin; t < cx.t_end; t += cx.t_step, ++it) {
        const int acc = n_acc == 4 ? (it & 3) : (it & 1);
        const uint32_t aph = (uint32_t)(n_acc == 4 ? (it >> 2) : (it >> 1)) & 1u;
        const TileCoord tc = tile_coord(p, t);                   // 1x1 conv: flat pixel tiles, tc.w0 = first pixel
        float *sp = stage + (size_t)(it & 1) * 2 * tile_f, *sr = sp + tile_f;
        ptx::mbar_wait(cx.tfull + acc, aph);
        ptx::tc_fence_after();
        const uint32_t taddr = cx.tmem_base + acc * p.BN + ((uint32_t)(quarter * 32) << 16);
        uint32_t raw[NO6 ? 16 : 32];
        ptx::tmem_ld16_nowait(taddr + ldcol, raw);
        if constexpr (!NO6) ptx::tmem_ld16_nowait(taddr + 16, raw + 16);      // generic head: all 32 columns
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(cx.tempty + acc);        // accumulator is in registers: the MMA warp may reuse the stage
        {
            const uint32_t pix = (uint32_t)tc.w0 + (uint32_t)row;
            const int b = (int)(((uint64_t)pix * p.div_hw) >> 40);
            const int rem = (int)(pix - (uint32_t)b * (uint32_t)p.img_hw);
            const int gy = (int)fast_div((uint32_t)rem, p.div_imgw), gx = rem - gy * p.img_w;
#pragma unroll
            for (int i = 0; i < (NO6 ? kDetHalf : 2 * kDetHalfMax); ++i) {
                // NO6: i walks the group's 9 columns, raw[j - ldcol]; generic: i walks all 32 accumulator columns (register
                // indices stay compile-time), the group keeps its own [col0, col0 + half)
                const int j = NO6 ? col0 + i : i;
                const bool mine = NO6 ? (j < nn) : (j < nn && j >= col0 && j < col0 + half);
                if (mine) {
                    const int a = j / no, o = j - a * no;
                    uint32_t rv;
                    if constexpr (NO6) rv = grp ? raw[i + 1] : raw[i];
                    else rv = raw[i];
                    const float tv = __uint_as_float(rv) + ld_shared_f(cx.sbias_u + (uint32_t)j * 4);
                    sr[a * rec + row * no + o] = tv;
                    sp[a * rec + row * no + o] = detect_decode(p, tv, a, o, gx, gy);
                }
            }
        }
        ptx::bar_sync(1, 256);
        const uint32_t pix0 = (uint32_t)tc.w0;
        if (p.cand_mask != nullptr) {
            // fused confidence filter (general.py:961 `prediction[..., 4] > conf_thres`) on the decoded objectness of the staged
            // tile: a warp = 32 consecutive pixels of one anchor = 32 consecutive rows of pred (inside one image), so its ballot is
            // (part of) one or two words of the image's candidate mask; lanes of the same word OR their bits with one atomic
            for (int c = tid; c < na * 128; c += 256) {                  // warp-uniform trip count (na * 128 is a multiple of 32)
                const int a = c >> 7, r = c & 127;
                const uint32_t pix = pix0 + (uint32_t)r;
                bool pass = false;
                uint32_t widx = 0, bit = 0;
                if ((int)pix < p.Wo) {
                    const int b = (int)(((uint64_t)pix * p.div_hw) >> 40);
                    const int rem = (int)(pix - (uint32_t)b * (uint32_t)p.img_hw);
                    pass = sp[a * rec + r * no + 4] > p.cand_conf;
                    const uint32_t i = (uint32_t)(p.row_off + a * p.img_hw + rem);
                    widx = (uint32_t)b * (uint32_t)p.mask_words + (i >> 5);
                    bit = 1u << (i & 31);
                }
                const uint32_t pm = __ballot_sync(0xffffffffu, pass);
                if (pass) {
                    const uint32_t peers = __match_any_sync(pm, widx);
                    const uint32_t bits = __reduce_or_sync(peers, bit);
                    if ((peers & ((1u << lane) - 1u)) == 0) atomicOr(p.cand_mask + widx, bits);
                }
            }
        }
        const int b0 = (int)(((uint64_t)pix0 * p.div_hw) >> 40);
        const int rem0 = (int)(pix0 - (uint32_t)b0 * (uint32_t)p.img_hw);
        const bool whole = vec_ok && rem0 + 128 <= p.img_hw && (int)pix0 + 128 <= p.Wo;     // one image, no batch tail
        const int n_out = p.raw != nullptr ? 2 : 1;
        if (whole) {
            const int vpa = rec / 4;                             // 16-byte vectors per anchor
            for (int v = tid; v < n_out * na * vpa; v += 256) {
                const int which = v >= na * vpa ? 1 : 0, vv = v - which * na * vpa;
                const int a = vv / vpa, q = vv - a * vpa;
                const float4 val = *reinterpret_cast<const float4 *>((which ? sr : sp) + a * rec + q * 4);
                float *dst = which ? p.raw + (((size_t)b0 * na + a) * p.img_hw + rem0) * no
                                   : p.pred + ((size_t)b0 * p.rows_total + p.row_off + (size_t)a * p.img_hw + rem0) * no;
                *reinterpret_cast<float4 *>(dst + q * 4) = val;
            }
        } else {
            for (int v = tid; v < n_out * tile_f; v += 256) {
                const int which = v >= tile_f ? 1 : 0, vv = v - which * tile_f;
                const int a = vv / rec, q = vv - a * rec;
                const int r = q / no, o = q - r * no;
                const uint32_t pix = pix0 + (uint32_t)r;
                if ((int)pix >= p.Wo) continue;
                const int b = (int)(((uint64_t)pix * p.div_hw) >> 40);
                const int rem = (int)(pix - (uint32_t)b * (uint32_t)p.img_hw);
                const float val = (which ? sr : sp)[vv];
                if (which) p.raw[(((size_t)b * na + a) * p.img_hw + rem) * no + o] = val;
                else p.pred[((size_t)b * p.rows_total + p.row_off + (size_t)a * p.img_hw + rem) * no + o] = val;
            }
        }
    }
}

__global__ void __launch_bounds__(kConvMaxThreads, 1) conv_umma_kernel(const __grid_constant__ ConvArgs p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int nb_slots = p.b_resident ? p.kblocks : p.b_stages;
    uint8_t *sA = smem;
    uint8_t *sB = sA + (size_t)p.a_stages * p.a_stage_bytes;
    uint8_t *sStage = sB + (size_t)nb_slots * p.b_stage_bytes;
    float *sbias = reinterpret_cast<float *>(sStage + (size_t)(p.n_groups == 4 ? 4 : 4) * p.stage_buf_bytes);
    float *sbv = reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(sbias) + ((p.cout_pad * 4 + 127) & ~127));
    uint64_t *bars = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(sbv) + p.bv_bytes);
    uint64_t *fullA = bars, *emptyA = fullA + 8, *fullB = emptyA + 8, *emptyB = fullB + 8;
    uint64_t *tfull = emptyB + 8, *tempty = tfull + 4, *bres = tempty + 4;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bres + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rb = p.kb * 2;                                   // bytes per operand row == swizzle span
    const int total_tiles = p.tiles_w * p.tiles_h * p.tiles_n * p.n_ntiles;
    // tile -> CTA: round robin (neighbouring CTAs work on neighbouring tiles: halo / weight reuse in L2), or one contiguous
    // range per CTA (per-image epilogue operands staged in shared memory change once per image instead of once per tile)
    int t_begin = blockIdx.x, t_end = total_tiles, t_step = gridDim.x;
    if (p.tile_contig) {
        const int per = (total_tiles + gridDim.x - 1) / gridDim.x;
        t_begin = blockIdx.x * per;
        t_end = min(total_tiles, t_begin + per);
        t_step = 1;
    }
    const int n_acc = p.n_acc;                                  // TMEM accumulator stages (MMA runs ahead of the epilogue)
    uint32_t tmem_cols = 32;
    while (tmem_cols < (uint32_t)(n_acc * p.BN)) tmem_cols <<= 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < 8; ++s) {
            ptx::mbar_init(fullA + s, 1);
            ptx::mbar_init(emptyA + s, 1);
            ptx::mbar_init(fullB + s, 1);
            ptx::mbar_init(emptyB + s, 1);
        }
        for (int a = 0; a < 4; ++a) {
            ptx::mbar_init(tfull + a, 1);
            ptx::mbar_init(tempty + a, p.ep_teams ? 4 : 4 * p.n_groups);
        }
        ptx::mbar_init(bres, 1);
        ptx::fence_mbar_init();
        ptx::prefetch_tmap(p.amap);
        ptx::prefetch_tmap(p.wmap);
        if (p.mode == 0) ptx::prefetch_tmap(p.omap);
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, tmem_cols);
        ptx::tmem_relinquish();
    }
    {
        const float scale = p.act == 1 ? 0.5f : 1.0f;
        for (int i = threadIdx.x; i < p.cout_pad; i += blockDim.x) sbias[i] = __ldg(p.bias + i) * scale;
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();                  // the next kernel may start its prologue on SMs this grid has left

    if (warp == 0) {
        // ===================== TMA producer (whole warp in uniform control flow, one elected lane issues) ==========
        const bool issuer = ptx::elect_one();
        if (p.b_resident && issuer) {
            const int nc_res = p.n_ntiles > 1 ? (int)(blockIdx.x % (unsigned)p.n_ntiles) * p.BN : 0;   // the one N tile this CTA works on
            ptx::mbar_expect_tx(bres, (uint32_t)(p.kblocks * p.b_stage_bytes));
            for (int kbi = 0; kbi < p.kblocks; ++kbi)
                ptx::tma_load_2d(sB + (size_t)kbi * p.b_stage_bytes, p.wmap, bres, kbi * p.kb, nc_res);
        }
        int sa = 0, sb = 0;
        uint32_t pha = 0, phb = 0;
        const bool stream_b = !p.b_resident;
        pdl_wait();                 // activations are written by earlier kernels (weights above are constants)
        for (int t = t_begin; t < t_end; t += t_step) {
            const TileCoord tc = tile_coord(p, t);
            if (p.a_mode == A_HALO) {
                for (int c = 0; c < p.cblk; ++c) {
                    ptx::mbar_wait(emptyA + sa, pha ^ 1);
                    if (issuer) {
                        ptx::mbar_expect_tx(fullA + sa, (uint32_t)p.a_box_bytes);
                        ptx::tma_load_4d(sA + (size_t)sa * p.a_stage_bytes, p.amap, fullA + sa, c * p.kb, tc.w0 - 1, tc.h0 - 1, tc.n0);
                    }
                    if (++sa == p.a_stages) { sa = 0; pha ^= 1; }
                    if (stream_b) {
                        for (int tap = 0; tap < 9; ++tap) {
                            ptx::mbar_wait(emptyB + sb, phb ^ 1);
                            if (issuer) {
                                ptx::mbar_expect_tx(fullB + sb, (uint32_t)p.b_stage_bytes);
                                ptx::tma_load_2d(sB + (size_t)sb * p.b_stage_bytes, p.wmap, fullB + sb, (tap * p.cblk + c) * p.kb, tc.nc0);
                            }
                            if (++sb == p.b_stages) { sb = 0; phb ^= 1; }
                        }
                    }
                }
            } else {
                int tap = 0, cb = 0;
                for (int kbi = 0; kbi < p.kblocks; ++kbi) {
                    ptx::mbar_wait(emptyA + sa, pha ^ 1);
                    if (issuer) {
                        ptx::mbar_expect_tx(fullA + sa, (uint32_t)p.a_box_bytes);
                        if (p.n_src > 1)
                            ptx::tma_load_4d(sA + (size_t)sa * p.a_stage_bytes, p.amap + p.kb_map[kbi], fullA + sa, p.kb_coord[kbi], tc.w0,
                                             tc.h0, tc.n0);
                        else
                            ptx::tma_load_4d(sA + (size_t)sa * p.a_stage_bytes, p.amap + p.tap_map[tap], fullA + sa, cb * p.kb,
                                             tc.w0 + p.tap_dw[tap], tc.h0 + p.tap_dh[tap], tc.n0);
                    }
                    if (++sa == p.a_stages) { sa = 0; pha ^= 1; }
                    if (stream_b) {
                        ptx::mbar_wait(emptyB + sb, phb ^ 1);
                        if (issuer) {
                            ptx::mbar_expect_tx(fullB + sb, (uint32_t)p.b_stage_bytes);
                            ptx::tma_load_2d(sB + (size_t)sb * p.b_stage_bytes, p.wmap, fullB + sb, kbi * p.kb, tc.nc0);
                        }
                        if (++sb == p.b_stages) { sb = 0; phb ^= 1; }
                    }
                    if (++cb == p.cblk) { cb = 0; ++tap; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        MmaCtx cx;
        cx.fullA = fullA; cx.emptyA = emptyA; cx.fullB = fullB; cx.emptyB = emptyB; cx.tfull = tfull; cx.tempty = tempty; cx.bres = bres;
        cx.sA_u = ptx::smem_u32(sA); cx.sB_u = ptx::smem_u32(sB); cx.tmem_base = tmem_base; cx.t_begin = t_begin; cx.t_end = t_end; cx.t_step = t_step; cx.n_acc = n_acc;
        if (p.a_mode == A_HALO) {
            if (p.b_resident) mma_role_ks<true, false>(p, cx); else mma_role_ks<true, true>(p, cx);
        } else {
            if (p.b_resident) mma_role_ks<false, false>(p, cx); else mma_role_ks<false, true>(p, cx);
        }
        __syncwarp();
    } else {
        // ===================== epilogue =====================
        EpiCtx cx;
        cx.tfull = tfull; cx.tempty = tempty; cx.tmem_base = tmem_base; cx.stage_u = ptx::smem_u32(sStage);
        cx.sbias_u = ptx::smem_u32(sbias); cx.sStage = sStage; cx.sbv = sbv; cx.t_begin = t_begin; cx.t_end = t_end; cx.t_step = t_step; cx.n_acc = n_acc;
        cx.warp = warp; cx.lane = lane;
        pdl_wait();                 // residual / per-image vector reads and all output writes come after the prerequisites
        if (p.mode != 0) { if (p.no == 6 && p.na == 3) epilogue_detect_role<true>(p, cx); else epilogue_detect_role<false>(p, cx); }
        else if (p.res != nullptr) epilogue_store_role<1>(p, cx);
        else if (p.bvec != nullptr) epilogue_store_role<2>(p, cx);
        else epilogue_store_role<0>(p, cx);
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, tmem_cols);
    }
}

}  // namespace

size_t conv_smem_bytes(const ConvArgs &a) {
    const int nb = a.b_resident ? a.kblocks : a.b_stages;
    return (size_t)a.a_stages * a.a_stage_bytes + (size_t)nb * a.b_stage_bytes + 4 * (size_t)a.stage_buf_bytes +   // 2x2 or 4x1 buffers
           ((a.cout_pad * 4 + 127) & ~127) + a.bv_bytes + kBarrierBytes + 1024;
}

int conv_plan_smem(ConvArgs &a, int max_seg_cols) {
    const int rb = a.kb * 2;
    const int box_rows = a.a_mode == A_HALO ? a.halo_w * (a.th + 2) : a.tw * a.th * a.tn;
    a.a_box_bytes = box_rows * rb;
    a.a_stage_bytes = a.a_mode == A_HALO ? ((a.a_box_bytes + 1023) & ~1023) : 128 * rb;
    a.b_stage_bytes = a.BN * rb;
    a.stage_buf_bytes = a.mode == 0 ? (((a.pool ? 160 : 128) * max_seg_cols * 2 + 1023) & ~1023)   // + pooled tile behind it
                                    : ((a.na * 128 * a.no * 4 + 1023) & ~1023);                    // Detect: 2 buffers x (decoded, raw) tiles
    const long fixed = 4L * a.stage_buf_bytes + ((a.cout_pad * 4 + 127) & ~127) + a.bv_bytes + kBarrierBytes + 1024;
    const long avail = kSmemLimit - fixed;
    const long b_total = (long)a.kblocks * a.b_stage_bytes;
    const int a_per_tile = a.a_mode == A_HALO ? a.cblk : a.kblocks;
    a.b_resident = ((a.n_ntiles == 1 || a.n_pinned) && b_total <= 100 * 1024 && avail - b_total >= 2L * a.a_stage_bytes) ? 1 : 0;
    if (a.b_resident) {
        a.b_stages = 0;
        a.a_stages = (int)std::min<long>(8, (avail - b_total) / a.a_stage_bytes);
        a.a_stages = std::min(a.a_stages, std::max(2, 6 * a_per_tile));
    } else if (a.a_mode == A_HALO) {
        a.a_stages = a.cblk >= 2 ? 3 : 2;
        while (a.a_stages > 2 && avail - (long)a.a_stages * a.a_stage_bytes < 4L * a.b_stage_bytes) --a.a_stages;
        a.b_stages = (int)std::min<long>(8, (avail - (long)a.a_stages * a.a_stage_bytes) / a.b_stage_bytes);
        if (a.b_stages < 2) return 1;
    } else {
        const int n = (int)std::min<long>(8, avail / (a.a_stage_bytes + a.b_stage_bytes));
        if (n < 2) return 1;
        a.a_stages = a.b_stages = n;
    }
    return a.a_stages >= 2 ? 0 : 1;
}

void conv_launch(const ConvArgs &a, int grid, cudaStream_t stream) {
    RY_ENSURE_DYN_SMEM(conv_umma_kernel, kSmemLimit);
    launch_pdl(conv_umma_kernel, dim3(grid), dim3(64 + 128 * a.n_groups), conv_smem_bytes(a), stream, a);
}

}  // namespace ry
