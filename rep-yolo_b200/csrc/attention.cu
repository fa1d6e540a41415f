// Attention kernels of CCVA (reference models/common.py:3675-3786).
//
//  attn_qk     : q = ReLU6(bn(SiLU(gconv_q(x)))), k likewise with the SAME bn (common.py:3693-3701); the grouped 1x1 convs
//                have 8 input channels per output channel (DWConv(c, c/8): groups = c/8, common.py:154-156).
//  crisscross  : eH[h,w,g] = q[h,w].k[g,w], eW[h,w,g] = q[h,w].k[h,g]; softmax over the H+W energies (no -inf diagonal);
//                out = sum_g v[g,w] aH + sum_g v[h,g] aW; gamma*out + x                     (common.py:3704-3726)
//                Two "line attention" passes merged like a split soft-max:
//                  row pass  (CTA = one image row)    -> un-normalised partial O_W, running max m_W, sum s_W (scratch)
//                  col pass  (CTA = one image column) -> O_H, m_H, s_H, merge with the row partials, gamma*out + x
//  vertical    : out[y=i,x=n] = sum_j v[j,n] * E(m = n*H + i)[j],  E(h*W + w)[j] = q[h,w].k[j,w]   (raw energies; the flat
//                re-view of the reference mixes rows and columns, common.py:3763-3778; H == W: out[y,x] = sum_j v[j,x] q[x,y].k[j,y])
//                  energy pass (CTA = image column j): E_j[i][k] = q[i,j].k[k,j]  -> bf16 scratch [b][i][j][k]
//                  value pass  (CTA = output column i): out[j,i] = E[i][j][:] . v[:,i]; gamma*out + x
//  v = ReLU6(bn1(SiLU(wv*x + bv))) is recomputed from x where needed (depthwise 1x1: purely per element).
//
// Both matrix products of a line run on the tensor cores (mma.sync m16n8k16, bf16 in / fp32 accumulate; the lines are
// 20..160 long, far below a tcgen05 tile).  Energies feed a soft-max, so q and k are split into bf16 hi + lo parts and
// E = Qhi.Khi + Qhi.Klo + Qlo.Khi is ONE contraction over 3*Cq (error ~2^-16 relative, fp32-like); soft-max statistics
// stay in fp32 registers; P and V enter the second product as bf16.
#include "memops.cuh"

#include <algorithm>

#include "common.cuh"

namespace ry {

namespace {

enum { MODE_ROW = 0, MODE_COL = 1, MODE_VE = 2, MODE_VPV = 3 };
constexpr int kStagePitch = 40;                               // epilogue staging row: 32 channels + 8 pad (bf16)
constexpr int kStageElems = 16 * kStagePitch + 32;            // per warp: 16 rows + 16 fp32 row weights

__global__ void __launch_bounds__(256) attn_qk_kernel(const __nv_bfloat16 *__restrict__ x, int x_cs, int x_off, int Cq,
                                                      size_t npix, const float *__restrict__ wq, const float *__restrict__ bq,
                                                      const float *__restrict__ wk, const float *__restrict__ bk,
                                                      const float *__restrict__ s, const float *__restrict__ t,
                                                      float *__restrict__ q, float *__restrict__ k) {
    pdl_trigger();
    pdl_wait();
    const size_t total = npix * Cq;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int d = (int)(i % Cq);
        const size_t pix = i / Cq;
        const uint4 u = __ldg(reinterpret_cast<const uint4 *>(x + pix * x_cs + x_off + d * 8));
        const float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
        const float xv[8] = {f0.x, f0.y, f1.x, f1.y, f2.x, f2.y, f3.x, f3.y};
        float aq = __ldg(bq + d), ak = __ldg(bk + d);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            aq = fmaf(__ldg(wq + d * 8 + j), xv[j], aq);
            ak = fmaf(__ldg(wk + d * 8 + j), xv[j], ak);
        }
        const float sc = __ldg(s + d), sh = __ldg(t + d);
        q[i] = relu6_f(fmaf(sc, silu_f(aq), sh));
        k[i] = relu6_f(fmaf(sc, silu_f(ak), sh));
    }
}

__device__ __forceinline__ float v_of(float x, float wv, float bv, float s1, float t1) {
    return relu6_f(fmaf(s1, silu_f(fmaf(wv, x, bv)), t1));
}

// The line kernels are instruction-issue bound (ncu: 58 % issue slots busy at 22 % occupancy, DRAM < 15 %), and two thirds of
// their instructions were the range-checked expansions of __expf / __fdividef in the operand staging.  Raw MUFU forms:
constexpr float kLog2e = 1.4426950408889634f;
__device__ __forceinline__ float ex2_fast(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_fast(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// q / k feed a soft-max: full-precision SiLU (two MUFU, ~2^-22 each)
__device__ __forceinline__ float silu_mufu(float x) { return x * rcp_fast(1.0f + ex2_fast(-kLog2e * x)); }
// v is rounded to bf16 right away: SiLU(z) = h + h*tanh(h), h = z/2 (one MUFU, |error| <= |h| * 2^-11), the 1/2 folded into wv, bv
__device__ __forceinline__ float v_fast(float x, float wv_h, float bv_h, float s1, float t1) {
    const float h = fmaf(wv_h, x, bv_h);
    return relu6_f(fmaf(s1, fmaf(h, tanh_fast(h), h), t1));
}
__device__ __forceinline__ void ldg8(float (&r)[8], const float *p) {
    const float4 a = __ldg(reinterpret_cast<const float4 *>(p)), b = __ldg(reinterpret_cast<const float4 *>(p) + 1);
    r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; r[4] = b.x; r[5] = b.y; r[6] = b.z; r[7] = b.w;
}

// ---- operand computation, fused into the line kernels' staging ----
// For every line position and channel group d (the 8 input channels of q/k output channel d = the 8 value channels 8d..8d+7):
//   q = ReLU6(bn(SiLU(gconv_q(x)))), k likewise with the SAME bn (common.py:3693-3701), v = ReLU6(bn1(SiLU(wv*x + bv)))
// written straight into the shared-memory operand rows of the line's two matrix products (bf16):
//   Aq[pos] = [Qhi | Qhi | Qlo | 0],  Bk[pos] = [Khi | Klo | Khi | 0]   (KQ = 3*Cq padded to 16)   ->  Aq.Bk^T = Qhi.Khi + Qhi.Klo + Qlo.Khi
//   Vs[pos] = v[0..C)
// One 16-byte load of x per (position, group) is everything a line pass reads for its operands: the former preparation
// kernel and its QA / KB / V round trip through HBM (3.1x the bytes of x per module) are gone.  A thread keeps the
// constants of its group d in registers (threads per line % Cq == 0).  Rows >= L are zero.
struct AttnW {
    const float *qk, *wv, *bv, *s1, *t1;
};

struct Geom {
    int L, LP, KQ;              // line length, padded to 16, padded 3*Cq contraction length
    int sq, sv, sp;             // row strides (elements) of Aq/Bk, Vs, Ps
    size_t stats_off, e_off;    // byte offsets in the scratch region: fp32 (max, sum) of the row pass, bf16 vertical energies
};

template <bool QK, bool VV, bool RAW, typename Src>
__device__ __forceinline__ void stage_operands(const AttnW &w, int Cq, const Geom &gm, int LP, int tis, int nthr, Src src,
                                               __nv_bfloat16 *Aq, __nv_bfloat16 *Bk, __nv_bfloat16 *Vs, __nv_bfloat16 *Xr) {
    const int d = tis % Cq, r0 = tis / Cq, rstep = nthr / Cq;
    float rwq[8], rwk[8], rwv[8], rbv[8], rs1[8], rt1[8];
    float rbq = 0.0f, rbk = 0.0f, sc = 0.0f, sh = 0.0f;
    if (QK) {
        const float *wq = w.qk, *bq = wq + Cq * 8, *wk = bq + Cq, *bk = wk + Cq * 8, *qs = bk + Cq, *qt = qs + Cq;
        ldg8(rwq, wq + d * 8);                                    // (every array of the pack starts 16-byte aligned: Cq % 4 == 0)
        ldg8(rwk, wk + d * 8);
        rbq = __ldg(bq + d); rbk = __ldg(bk + d); sc = __ldg(qs + d); sh = __ldg(qt + d);
    }
    if (VV) {
        ldg8(rwv, w.wv + d * 8);
        ldg8(rbv, w.bv + d * 8);
        ldg8(rs1, w.s1 + d * 8);
        ldg8(rt1, w.t1 + d * 8);
#pragma unroll
        for (int j = 0; j < 8; ++j) { rwv[j] *= 0.5f; rbv[j] *= 0.5f; }
    }
    const int npad = gm.KQ - 3 * Cq;                               // 0, 4 or 8 zero columns behind the three Cq-wide parts
    constexpr int U = 4;
    for (int pb = r0; pb < LP; pb += U * rstep) {
        uint4 u[U];
#pragma unroll
        for (int r = 0; r < U; ++r) {                                // all loads in flight before the math
            const int pi = pb + r * rstep;
            u[r] = pi < gm.L ? src(pi, d) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int r = 0; r < U; ++r) {
            const int pi = pb + r * rstep;
            if (pi >= LP) break;
            const bool ok = pi < gm.L;
            if (RAW) *reinterpret_cast<uint4 *>(Xr + (size_t)pi * gm.sv + d * 8) = u[r];      // the line itself, for the "+ x" of the epilogue
            const float2 f0 = unpack_bf16x2(u[r].x), f1 = unpack_bf16x2(u[r].y), f2 = unpack_bf16x2(u[r].z), f3 = unpack_bf16x2(u[r].w);
            const float xv[8] = {f0.x, f0.y, f1.x, f1.y, f2.x, f2.y, f3.x, f3.y};
            if (QK) {
                float qa = 0.0f, kb = 0.0f;
                if (ok) {
                    float aq = rbq, ak = rbk;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        aq = fmaf(rwq[j], xv[j], aq);
                        ak = fmaf(rwk[j], xv[j], ak);
                    }
                    qa = relu6_f(fmaf(sc, silu_mufu(aq), sh));
                    kb = relu6_f(fmaf(sc, silu_mufu(ak), sh));
                }
                const __nv_bfloat16 qh = __float2bfloat16_rn(qa), kh = __float2bfloat16_rn(kb);
                const __nv_bfloat16 ql = __float2bfloat16_rn(qa - __bfloat162float(qh)), kl = __float2bfloat16_rn(kb - __bfloat162float(kh));
                __nv_bfloat16 *ar = Aq + (size_t)pi * gm.sq, *br = Bk + (size_t)pi * gm.sq;
                ar[d] = qh; ar[Cq + d] = qh; ar[2 * Cq + d] = ql;
                br[d] = kh; br[Cq + d] = kl; br[2 * Cq + d] = kh;
                if (d < npad) ar[3 * Cq + d] = br[3 * Cq + d] = __float2bfloat16_rn(0.0f);
            }
            if (VV) {
                uint32_t ow[4] = {0u, 0u, 0u, 0u};
                if (ok) {
#pragma unroll
                    for (int h = 0; h < 4; ++h)
                        ow[h] = pack_bf16x2(v_fast(xv[2 * h], rwv[2 * h], rbv[2 * h], rs1[2 * h], rt1[2 * h]),
                                            v_fast(xv[2 * h + 1], rwv[2 * h + 1], rbv[2 * h + 1], rs1[2 * h + 1], rt1[2 * h + 1]));
                }
                *reinterpret_cast<uint4 *>(Vs + (size_t)pi * gm.sv + d * 8) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
            }
        }
    }
}

// ---- warp-level tensor-core primitives ----
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float *c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, bool ok) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? 16 : 0) : "memory");   // !ok: zero fill
}

// E = Aq . Bk^T for the warp's 16 query rows (row0..row0+15), all LP columns
template <int NT>
__device__ __forceinline__ void line_energies(float (&e)[NT][4], const __nv_bfloat16 *Aq, const __nv_bfloat16 *Bk, int sq, int KQ, int row0, int lane) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) e[nt][0] = e[nt][1] = e[nt][2] = e[nt][3] = 0.0f;
    const uint32_t a_base = smem_addr(Aq + (size_t)(row0 + (lane & 15)) * sq + (lane >> 4) * 8);
    const uint32_t b_base = smem_addr(Bk + (size_t)((lane & 7) + ((lane >> 4) << 3)) * sq + ((lane >> 3) & 1) * 8);
    for (int kt = 0; kt < KQ / 16; ++kt) {
        uint32_t a0, a1, a2, a3;
        ldsm_x4(a_base + kt * 32, a0, a1, a2, a3);
#pragma unroll
        for (int np = 0; np < NT / 2; ++np) {
            uint32_t b0, b1, b2, b3;
            ldsm_x4(b_base + (uint32_t)(np * 16 * sq * 2) + kt * 32, b0, b1, b2, b3);
            mma_bf16(e[2 * np], a0, a1, a2, a3, b0, b1);
            mma_bf16(e[2 * np + 1], a0, a1, a2, a3, b2, b3);
        }
    }
}

// raw energies of image column `line` -> bf16 scratch [b][i][j = line][k]; padded k columns are exact zeros (zero Bk rows)
template <int NT>
__device__ __forceinline__ void store_energies(const float (&e)[NT][4], __nv_bfloat16 *E, int b, int H, int W, int line, int L, int row0,
                                               int g, int t4) {
    constexpr int LP = NT * 8;
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
        const int i = row0 + g + 8 * hh;
        if (i < L) {
            uint32_t *dst = reinterpret_cast<uint32_t *>(E + (((size_t)b * H + i) * W + line) * LP) + t4;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) dst[nt * 4] = pack_bf16x2(e[nt][2 * hh], e[nt][2 * hh + 1]);
        }
    }
}

// LPC lines per CTA (short lines share a CTA for occupancy), one 16-row query tile per warp.  NT = LP / 8 (compile time:
// register arrays).  Each line slot is an independent "virtual CTA" of NT*16 threads; only __syncthreads is shared.
// FUSE (column pass only): the criss-cross output column just computed is also the input column of the VerticalAttention
// that follows (CCVA: m1(m(x)), common.py:2654-2655), whose energy pass needs exactly one image column of q/k: it runs
// here on the column kept in shared memory (w2 = the vertical module's parameters) instead of a launch that re-reads it.
// (forcing 5 CTAs / SM through __launch_bounds__ -- 72 registers, ~80 B of spills -- was measured slower: 0.74 vs 0.68 ms)
template <int MODE, int NT, int LPC, bool FUSE>
__global__ void __launch_bounds__(NT * 16 * LPC, ((MODE == MODE_ROW || MODE == MODE_COL) && NT <= 12) ? 640 / (NT * 16 * LPC) : 1) attn_mma_kernel(const AttnParams p, const AttnW w1, const AttnW w2, const Geom gm,
                                                                 int total_lines, size_t slot_bytes) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(16) uint8_t sm_all[];
    constexpr int LP = NT * 8;
    constexpr int kThreads = NT * 16;                        // threads per line slot
    const int slot = threadIdx.x / kThreads, tis = threadIdx.x - slot * kThreads;
    const int vb = blockIdx.x * LPC + slot;                  // global line index
    const bool active = vb < total_lines;
    uint8_t *sm_raw = sm_all + (size_t)slot * slot_bytes;
    const int L = gm.L, C = p.C, KQ = gm.KQ;
    const int sq = gm.sq, sv = gm.sv, sp = gm.sp;
    __nv_bfloat16 *Aq = reinterpret_cast<__nv_bfloat16 *>(sm_raw);
    __nv_bfloat16 *Bk = Aq + (size_t)LP * sq;
    __nv_bfloat16 *Vs = (MODE == MODE_VPV) ? Aq : Bk + (size_t)LP * sq;      // value pass has no q/k operands
    __nv_bfloat16 *Ps = (MODE == MODE_VE) ? Vs : Vs + (size_t)LP * sv;        // value pass only: P = the energies read back from HBM
    __nv_bfloat16 *Xr = Ps;                                                   // COL: the raw line (row / column passes keep P in registers)
    __nv_bfloat16 *Xs = Xr + (size_t)LP * sv;                                 // FUSE: the output column (bf16, rows of sv elements)
    const int lines = (MODE == MODE_ROW) ? p.H : p.W;
    const int b = (active ? vb : 0) / lines, line = (active ? vb : 0) % lines;
    const size_t img = (size_t)b * p.H * p.W;
    uint8_t *scratch = reinterpret_cast<uint8_t *>(p.scratch);
    float *row_stats = reinterpret_cast<float *>(scratch + gm.stats_off);
    __nv_bfloat16 *Escr = reinterpret_cast<__nv_bfloat16 *>(scratch + gm.e_off);
    const int warp = tis >> 5, lane = tis & 31;
    const int g = lane >> 2, t4 = lane & 3;
    auto pix_of = [&](int i) -> size_t {       // pixel of line position i
        return (MODE == MODE_ROW) ? img + (size_t)line * p.W + i : img + (size_t)i * p.W + line;
    };
    auto x_src = [&](int pi, int d) -> uint4 {
        return __ldg(reinterpret_cast<const uint4 *>(p.x + pix_of(pi) * p.x_cs + p.x_off + d * 8));
    };

    // ---- stage operands: q / k / v computed from the line's x values (rows >= L are zero) ----
    if (active) {
        if (MODE == MODE_VPV) {
            // P[j][k] = E[b][i = line][j][k] (bf16 scratch written by the energy pass), rows j >= L are zero
            // general H x W (the reference's view chain, common.py:3763-3775): output column n' uses the H energy rows of the
            // flat pixels n'*H .. n'*H + H-1 (row-major) -- for H == W that is image row n'
            const __nv_bfloat16 *E = Escr + ((size_t)b * p.H * p.W + (size_t)line * p.H) * LP;
            const int vecs = LP / 8;
            const uint32_t ps_u = smem_addr(Ps);
            for (int idx = tis; idx < LP * vecs; idx += kThreads) {
                const int j = idx / vecs, c = (idx - j * vecs) * 8;
                cp_async16(ps_u + (uint32_t)(j * sp + c) * 2, E + (size_t)(j < L ? j : 0) * LP + c, j < L);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            stage_operands<false, true, false>(w1, p.Cq, gm, LP, tis, kThreads, x_src, Aq, Bk, Vs, Xr);
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        } else if (MODE == MODE_VE) {
            stage_operands<true, false, false>(w1, p.Cq, gm, LP, tis, kThreads, x_src, Aq, Bk, Vs, Xr);
        } else if (MODE == MODE_COL) {
            stage_operands<true, true, true>(w1, p.Cq, gm, LP, tis, kThreads, x_src, Aq, Bk, Vs, Xr);
        } else {
            stage_operands<true, true, false>(w1, p.Cq, gm, LP, tis, kThreads, x_src, Aq, Bk, Vs, Xr);
        }
    }
    __syncthreads();

    const int row0 = warp * 16;                                  // this warp's query rows
    float m_row[2] = {0.0f, 0.0f}, s_row[2] = {1.0f, 1.0f};
    // P of the row / column passes as the A fragments of the second product, straight from the soft-max registers: the fp32
    // accumulator layout of two adjacent 8-column tiles of m16n8k16 IS the bf16 A fragment of one 16-wide K step
    uint32_t pa[NT / 2][4];
    float2 rstat[2] = {make_float2(0.0f, 1.0f), make_float2(0.0f, 1.0f)};
    if (MODE == MODE_COL && active) {                             // row-pass statistics of my query rows: in flight during the energy product
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int i = row0 + g + 8 * hh;
            if (i < L) rstat[hh] = __ldg(reinterpret_cast<const float2 *>(row_stats + pix_of(i) * 2));
        }
    }
    if (MODE != MODE_VPV && active) {
        float e[NT][4];
        line_energies<NT>(e, Aq, Bk, sq, KQ, row0, lane);
        if (MODE == MODE_VE) {
            store_energies<NT>(e, Escr, b, p.H, p.W, line, L, row0, g, t4);
        } else {
            // ---- soft-max statistics per query row (columns >= L masked), P -> bf16 shared ----
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                float m = -INFINITY;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const int col = nt * 8 + 2 * t4;
                    if (col < L) m = fmaxf(m, e[nt][2 * hh]);
                    if (col + 1 < L) m = fmaxf(m, e[nt][2 * hh + 1]);
                }
                m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
                m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
                m *= kLog2e;                                     // statistics in the log2 domain: p = 2^(e*log2e - m)
                float s = 0.0f;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const int col = nt * 8 + 2 * t4;
                    const float p0 = col < L ? ex2_fast(fmaf(e[nt][2 * hh], kLog2e, -m)) : 0.0f;
                    const float p1 = col + 1 < L ? ex2_fast(fmaf(e[nt][2 * hh + 1], kLog2e, -m)) : 0.0f;
                    s += p0 + p1;
                    pa[nt >> 1][(nt & 1) * 2 + hh] = pack_bf16x2(p0, p1);
                }
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                m_row[hh] = m;
                s_row[hh] = s;
            }
            __syncwarp();
        }
    }
    if (MODE == MODE_VE) return;
    if (MODE != MODE_VPV) __syncthreads();                       // every warp is done with Aq / Bk: the region becomes the epilogue staging

    if (active) {
        // ---- O = P . V in chunks of 32 channels.  Epilogue per chunk through a per-warp staging tile (16 rows x 32 channels bf16,
        //      in the dead Aq/Bk region; the value pass has its own) so that every global access is a 16-byte vector of one pixel ----
        __nv_bfloat16 *wst = ((MODE == MODE_VPV) ? Ps + (size_t)LP * sp : Aq) + (size_t)warp * kStageElems;
        float *wsc = reinterpret_cast<float *>(wst + 16 * kStagePitch);                       // [16] per-row weight of the row partial
        float a_h[2] = {1.0f, 1.0f};
        if (MODE == MODE_COL) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int i = row0 + g + 8 * hh;
                float aw = 0.0f;
                if (i < L) {
                    const float mw = rstat[hh].x, sw = rstat[hh].y, mh = m_row[hh], sh = s_row[hh];
                    const float m = fmaxf(mh, mw);
                    const float fh = ex2_fast(mh - m), fw = ex2_fast(mw - m);          // (max, sum) of both passes are in the log2 domain
                    const float inv = 1.0f / (sh * fh + sw * fw);
                    a_h[hh] = fh * inv;
                    aw = fw * inv;
                }
                if (t4 == 0) wsc[g + 8 * hh] = aw;
            }
        }
        if (MODE == MODE_ROW && t4 == 0) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int i = row0 + g + 8 * hh;
                if (i < L) *reinterpret_cast<float2 *>(row_stats + pix_of(i) * 2) = make_float2(m_row[hh], s_row[hh]);
            }
        }
        const uint32_t pa_base = smem_addr(Ps + (size_t)(row0 + (lane & 15)) * sp + (lane >> 4) * 8);
        const uint32_t vb_base = smem_addr(Vs + (size_t)((lane & 7) + ((lane >> 3) & 1) * 8) * sv + (lane >> 4) * 8);
        for (int c0 = 0; c0 < C; c0 += 32) {
            float o[4][4];
#pragma unroll
            for (int ct = 0; ct < 4; ++ct) o[ct][0] = o[ct][1] = o[ct][2] = o[ct][3] = 0.0f;
            uint4 pwv[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
            if (MODE == MODE_COL || MODE == MODE_VPV) {           // row partials (COL) / x (VPV) of this chunk: in flight during the P.V product
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const int idx = lane + 32 * k, i = row0 + (idx >> 2);
                    if (i >= L) continue;
                    if (MODE == MODE_COL)
                        pwv[k] = __ldg(reinterpret_cast<const uint4 *>(reinterpret_cast<const __nv_bfloat16 *>(p.scratch) + pix_of(i) * C + c0 + (idx & 3) * 8));
                    else
                        pwv[k] = __ldg(reinterpret_cast<const uint4 *>(p.x + pix_of(i) * p.x_cs + p.x_off + c0 + (idx & 3) * 8));
                }
            }
#pragma unroll
            for (int kt = 0; kt < NT / 2; ++kt) {
                uint32_t a0, a1, a2, a3;
                if (MODE == MODE_VPV) ldsm_x4(pa_base + kt * 32, a0, a1, a2, a3);
                else { a0 = pa[kt][0]; a1 = pa[kt][1]; a2 = pa[kt][2]; a3 = pa[kt][3]; }
#pragma unroll
                for (int cp = 0; cp < 2; ++cp) {
                    uint32_t b0, b1, b2, b3;
                    ldsm_x4_trans(vb_base + (uint32_t)(kt * 16 * sv * 2) + (uint32_t)((c0 + cp * 16) * 2), b0, b1, b2, b3);
                    mma_bf16(o[2 * cp], a0, a1, a2, a3, b0, b1);
                    mma_bf16(o[2 * cp + 1], a0, a1, a2, a3, b2, b3);
                }
            }
            __syncwarp();                                            // previous chunk's staging reads are done
#pragma unroll
            for (int hh = 0; hh < 2; ++hh)
#pragma unroll
                for (int ct = 0; ct < 4; ++ct)
                    *reinterpret_cast<uint32_t *>(wst + (g + 8 * hh) * kStagePitch + ct * 8 + 2 * t4) =
                        pack_bf16x2(o[ct][2 * hh] * a_h[hh], o[ct][2 * hh + 1] * a_h[hh]);
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int idx = lane + 32 * k, r = idx >> 2, v = idx & 3;
                const int i = row0 + r;
                if (i >= L) continue;
                const size_t px = pix_of(i);
                const uint4 oh = *reinterpret_cast<const uint4 *>(wst + r * kStagePitch + v * 8);
                const int c = c0 + v * 8;
                if (MODE == MODE_ROW) {
                    // un-normalised row partial O_W as bf16 [pix][C] (it re-enters a bf16 result scaled by gamma)
                    *reinterpret_cast<uint4 *>(reinterpret_cast<__nv_bfloat16 *>(p.scratch) + px * C + c) = oh;
                } else {
                    const uint4 xw = (MODE == MODE_COL) ? *reinterpret_cast<const uint4 *>(Xr + (size_t)i * sv + c) : pwv[k];
                    const uint4 pw = (MODE == MODE_COL) ? pwv[k] : make_uint4(0, 0, 0, 0);
                    const float aw = (MODE == MODE_COL) ? wsc[r] : 0.0f;
                    const uint32_t ohw[4] = {oh.x, oh.y, oh.z, oh.w}, xww[4] = {xw.x, xw.y, xw.z, xw.w}, pww[4] = {pw.x, pw.y, pw.z, pw.w};
                    uint32_t res[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float2 fo = unpack_bf16x2(ohw[q]), fx = unpack_bf16x2(xww[q]), fp = unpack_bf16x2(pww[q]);
                        res[q] = pack_bf16x2(fmaf(p.gamma, fmaf(fp.x, aw, fo.x), fx.x), fmaf(p.gamma, fmaf(fp.y, aw, fo.y), fx.y));
                    }
                    const uint4 rv = make_uint4(res[0], res[1], res[2], res[3]);
                    *reinterpret_cast<uint4 *>(p.out + px * p.out_cs + p.out_off + c) = rv;
                    if (FUSE) *reinterpret_cast<uint4 *>(Xs + (size_t)i * sv + c) = rv;
                }
            }
        }
    }
    if (FUSE) {
        // ---- VerticalAttention energy pass on the column just written (kept in Xs) ----
        __syncthreads();                                         // all staging tiles (in the Aq / Bk region) are dead, Xs is complete
        if (active) {
            auto xs_src = [&](int pi, int d) -> uint4 { return *reinterpret_cast<const uint4 *>(Xs + (size_t)pi * sv + d * 8); };
            stage_operands<true, false, false>(w2, p.Cq, gm, LP, tis, kThreads, xs_src, Aq, Bk, Vs, Xr);
        }
        __syncthreads();
        if (active) {
            float e[NT][4];
            line_energies<NT>(e, Aq, Bk, sq, KQ, row0, lane);
            store_energies<NT>(e, Escr, b, p.H, p.W, line, L, row0, g, t4);
        }
    }
}

int pick_nt(int L) {
    static const int opts[] = {2, 4, 6, 8, 10, 12, 16, 20};
    for (int nt : opts)
        if (nt * 8 >= L) return nt;
    return 0;
}

size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

// scratch = [criss-cross row partials bf16 [B*H*W][C] | fp32 (max, sum) [B*H*W][2] | vertical energies bf16 [B][H][W][LP]]
// (the fused column pass reads row partials while it writes energies, so the two regions are disjoint)
struct ScratchLayout { size_t stats, e, total; };
ScratchLayout scratch_layout(int B, int H, int W, int C) {
    const size_t npix = (size_t)B * H * W;
    ScratchLayout s;
    s.stats = align256(npix * C * 2);
    s.e = align256(s.stats + npix * 8);
    s.total = s.e + align256(npix * (size_t)(pick_nt(H) * 8) * 2);
    return s;
}

AttnW weights_of(const AttnParams &p) { return AttnW{p.qk, p.wv, p.bv, p.s1, p.t1}; }

template <int MODE, int NT, bool FUSE>
int launch_nt(const AttnParams &p, const AttnW &w2, int L, cudaStream_t st) {
    constexpr int LPC = NT <= 4 ? 4 : (NT <= 8 ? 2 : 1);      // lines per CTA: keep CTAs at >= 4 warps
    if ((NT * 16) % p.Cq != 0) return 1;                      // staging: a thread owns one channel group
    Geom gm;
    gm.L = L;
    gm.LP = NT * 8;
    gm.KQ = (3 * p.Cq + 15) / 16 * 16;
    gm.sq = gm.KQ + 8;
    gm.sv = p.C + 8;
    gm.sp = gm.LP + 8;
    const ScratchLayout sl = scratch_layout(p.B, p.H, p.W, p.C);
    gm.stats_off = sl.stats;
    gm.e_off = sl.e;
    size_t smem = 0;
    if (MODE != MODE_VPV) smem += 2 * (size_t)gm.LP * gm.sq * 2;
    if (MODE != MODE_VE) smem += (size_t)gm.LP * gm.sv * 2;
    if (MODE == MODE_VPV) smem += (size_t)gm.LP * gm.sp * 2;                   // (row / column passes keep P in registers)
    if (MODE == MODE_VPV) smem += (size_t)(NT / 2) * kStageElems * 2;          // row / column passes stage in the dead q/k operand region
    if (MODE == MODE_COL) smem += (size_t)gm.LP * gm.sv * 2;                   // the raw line (the "+ x" of the epilogue)
    if (FUSE) smem += (size_t)gm.LP * gm.sv * 2;                               // the output column kept for the fused energy pass
    smem = (smem + 15) & ~size_t(15);
    if (smem * LPC > 227 * 1024) return 1;
    if (RY_ENSURE_DYN_SMEM((attn_mma_kernel<MODE, NT, LPC, FUSE>), 227 * 1024) != cudaSuccess) return 1;
    const int lines = (MODE == MODE_ROW) ? p.H : p.W;
    const int total = p.B * lines;
    launch_pdl(attn_mma_kernel<MODE, NT, LPC, FUSE>, dim3((total + LPC - 1) / LPC), dim3(NT * 16 * LPC), smem * LPC, st, p, weights_of(p), w2,
               gm, total, smem);
    return 0;
}

template <int MODE, bool FUSE>
int launch_mode(const AttnParams &p, const AttnW &w2, cudaStream_t st) {
    const int L = (MODE == MODE_ROW) ? p.W : p.H;
    if (p.C % 32 != 0 || p.C != p.Cq * 8) return 1;
    switch (pick_nt(L)) {
        case 2: return launch_nt<MODE, 2, FUSE>(p, w2, L, st);
        case 4: return launch_nt<MODE, 4, FUSE>(p, w2, L, st);
        case 6: return launch_nt<MODE, 6, FUSE>(p, w2, L, st);
        case 8: return launch_nt<MODE, 8, FUSE>(p, w2, L, st);
        case 10: return launch_nt<MODE, 10, FUSE>(p, w2, L, st);
        case 12: return launch_nt<MODE, 12, FUSE>(p, w2, L, st);
        case 16: return launch_nt<MODE, 16, FUSE>(p, w2, L, st);
        case 20: return launch_nt<MODE, 20, FUSE>(p, w2, L, st);
        default: return 1;
    }
}

}  // namespace

void attn_qk_launch(const __nv_bfloat16 *x, int x_cs, int x_off, int C, int Cq, size_t npix, const float *wq,
                    const float *bq, const float *wk, const float *bk, const float *s, const float *t, float *q, float *k,
                    cudaStream_t st) {
    (void)C;
    const size_t total = npix * Cq;
    size_t g = (total + 255) / 256;
    if (g > (size_t)num_sms() * 16) g = (size_t)num_sms() * 16;
    launch_pdl(attn_qk_kernel, dim3((int)g), dim3(256), 0, st, x, x_cs, x_off, Cq, npix, wq, bq, wk, bk, s, t, q, k);
}

size_t attn_scratch_bytes(int B, int H, int W, int C) { return scratch_layout(B, H, W, C).total; }

// next != NULL: the VerticalAttention that consumes this module's output (same geometry): its energy pass is fused into the
// column pass here, and vertical_launch(..., energies_ready = 1) then only runs the value pass.
int crisscross_launch(const AttnParams &p, const AttnParams *next, int *energies_done, cudaStream_t st) {
    const AttnW none{nullptr, nullptr, nullptr, nullptr, nullptr};
    if (energies_done) *energies_done = 0;
    if (launch_mode<MODE_ROW, false>(p, none, st)) return 1;
    if (next != nullptr && energies_done != nullptr && next->C == p.C && next->H == p.H && next->W == p.W && next->B == p.B &&
        launch_mode<MODE_COL, true>(p, weights_of(*next), st) == 0) {                  // (does not fit shared memory: unfused below)
        *energies_done = 1;
        return 0;
    }
    return launch_mode<MODE_COL, false>(p, none, st);
}

int vertical_launch(const AttnParams &p, int energies_ready, cudaStream_t st) {
    const AttnW none{nullptr, nullptr, nullptr, nullptr, nullptr};
    if (!energies_ready && launch_mode<MODE_VE, false>(p, none, st)) return 1;
    return launch_mode<MODE_VPV, false>(p, none, st);
}

}  // namespace ry
