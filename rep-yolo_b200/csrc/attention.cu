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

#include "common.cuh"

namespace ry {

namespace {

enum { MODE_ROW = 0, MODE_COL = 1, MODE_VE = 2, MODE_VPV = 3 };

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

struct Geom {
    int L, LP, KQ;              // line length, padded to 16, padded 3*Cq contraction length
    int sq, sv, sp;             // row strides (elements) of Aq/Bk, Vs, Ps
};

__device__ __forceinline__ float v_of(float x, float wv, float bv, float s1, float t1) {
    return relu6_f(fmaf(s1, silu_f(fmaf(wv, x, bv)), t1));
}

// LPC lines per CTA (short lines share a CTA for occupancy), one 16-row query tile per warp.  NT = LP / 8 (compile time:
// register arrays).  Each line slot is an independent "virtual CTA" of NT*16 threads; only __syncthreads is shared.
template <int MODE, int NT, int LPC>
__global__ void __launch_bounds__(NT * 16 * LPC) attn_mma_kernel(const AttnParams p, const Geom gm, int total_lines, size_t slot_bytes) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(16) uint8_t sm_all[];
    constexpr int LP = NT * 8;
    constexpr int kThreads = NT * 16;                        // threads per line slot
    const int slot = threadIdx.x / kThreads, tis = threadIdx.x - slot * kThreads;
    const int vb = blockIdx.x * LPC + slot;                  // global line index
    const bool active = vb < total_lines;
    uint8_t *sm_raw = sm_all + (size_t)slot * slot_bytes;
    const int L = gm.L, C = p.C, Cq = p.Cq, KQ = gm.KQ;
    const int sq = gm.sq, sv = gm.sv, sp = gm.sp;
    __nv_bfloat16 *Aq = reinterpret_cast<__nv_bfloat16 *>(sm_raw);
    __nv_bfloat16 *Bk = Aq + (size_t)LP * sq;
    __nv_bfloat16 *Vs = (MODE == MODE_VPV) ? Aq : Bk + (size_t)LP * sq;      // value pass has no q/k operands
    __nv_bfloat16 *Ps = (MODE == MODE_VE) ? Vs : Vs + (size_t)LP * sv;        // energy pass has no V / P
    const int lines = (MODE == MODE_ROW) ? p.H : p.W;
    const int b = (active ? vb : 0) / lines, line = (active ? vb : 0) % lines;
    const size_t img = (size_t)b * p.H * p.W;
    const int warp = tis >> 5, lane = tis & 31;
    const int g = lane >> 2, t4 = lane & 3;
    auto pix_of = [&](int i) -> size_t {       // pixel of line position i
        return (MODE == MODE_ROW) ? img + (size_t)line * p.W + i : img + (size_t)i * p.W + line;
    };

    // ---- stage operands ----
    if (MODE != MODE_VPV && active) {
        // Aq row = [Qhi | Qhi | Qlo | 0], Bk row = [Khi | Klo | Khi | 0]  ->  Aq.Bk^T = Qhi.Khi + Qhi.Klo + Qlo.Khi
        // q = ReLU6(bn(SiLU(gconv_q(x)))), k likewise with the SAME bn (common.py:3693-3701), computed here from the 8 input
        // channels of group d -- the same 16 bytes of x that give the 8 value channels v[8d..8d+7] (no q/k round trip through HBM)
        const __nv_bfloat16 zero = __float2bfloat16_rn(0.0f);
        const float *wq = p.qk, *bq = wq + Cq * 8, *wk = bq + Cq, *bk = wk + Cq * 8, *qs = bk + Cq, *qt = qs + Cq;
        for (int idx = tis; idx < LP * Cq; idx += kThreads) {
            const int pi = idx / Cq, d = idx - pi * Cq;
            float qa = 0.0f, kb = 0.0f;
            uint4 vo = make_uint4(0, 0, 0, 0);
            if (pi < L) {
                const uint4 u = __ldg(reinterpret_cast<const uint4 *>(p.x + pix_of(pi) * p.x_cs + p.x_off + d * 8));
                const float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
                const float xv[8] = {f0.x, f0.y, f1.x, f1.y, f2.x, f2.y, f3.x, f3.y};
                float aq = __ldg(bq + d), ak = __ldg(bk + d);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    aq = fmaf(__ldg(wq + d * 8 + j), xv[j], aq);
                    ak = fmaf(__ldg(wk + d * 8 + j), xv[j], ak);
                }
                const float sc = __ldg(qs + d), sh = __ldg(qt + d);
                qa = relu6_f(fmaf(sc, silu_f(aq), sh));
                kb = relu6_f(fmaf(sc, silu_f(ak), sh));
                if (MODE != MODE_VE) {
                    uint32_t ow[4];
#pragma unroll
                    for (int h = 0; h < 4; ++h) {
                        const int c0 = d * 8 + 2 * h;
                        ow[h] = pack_bf16x2(v_of(xv[2 * h], __ldg(p.wv + c0), __ldg(p.bv + c0), __ldg(p.s1 + c0), __ldg(p.t1 + c0)),
                                            v_of(xv[2 * h + 1], __ldg(p.wv + c0 + 1), __ldg(p.bv + c0 + 1), __ldg(p.s1 + c0 + 1),
                                                 __ldg(p.t1 + c0 + 1)));
                    }
                    vo = make_uint4(ow[0], ow[1], ow[2], ow[3]);
                }
            }
            const __nv_bfloat16 qh = __float2bfloat16_rn(qa), kh = __float2bfloat16_rn(kb);
            const __nv_bfloat16 ql = __float2bfloat16_rn(qa - __bfloat162float(qh)), kl = __float2bfloat16_rn(kb - __bfloat162float(kh));
            __nv_bfloat16 *ar = Aq + (size_t)pi * sq, *br = Bk + (size_t)pi * sq;
            ar[d] = qh; ar[Cq + d] = qh; ar[2 * Cq + d] = ql;
            br[d] = kh; br[Cq + d] = kl; br[2 * Cq + d] = kh;
            if (MODE != MODE_VE) *reinterpret_cast<uint4 *>(Vs + (size_t)pi * sv + d * 8) = vo;
        }
        const int padc = KQ - 3 * Cq;
        for (int idx = tis; idx < LP * padc; idx += kThreads) {
            const int pi = idx / padc, col = 3 * Cq + idx - pi * padc;
            Aq[(size_t)pi * sq + col] = zero;
            Bk[(size_t)pi * sq + col] = zero;
        }
    }
    if (MODE == MODE_VPV && active) {
        const int vecs = C / 8;
        for (int idx = tis; idx < LP * vecs; idx += kThreads) {
            const int pi = idx / vecs, c = (idx - pi * vecs) * 8;
            uint4 o = make_uint4(0, 0, 0, 0);
            if (pi < L) {
                const uint4 u = __ldg(reinterpret_cast<const uint4 *>(p.x + pix_of(pi) * p.x_cs + p.x_off + c));
                const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
                uint32_t ow[4];
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    const float2 f = unpack_bf16x2(uw[h]);
                    const int c0 = c + 2 * h;
                    ow[h] = pack_bf16x2(v_of(f.x, __ldg(p.wv + c0), __ldg(p.bv + c0), __ldg(p.s1 + c0), __ldg(p.t1 + c0)),
                                        v_of(f.y, __ldg(p.wv + c0 + 1), __ldg(p.bv + c0 + 1), __ldg(p.s1 + c0 + 1), __ldg(p.t1 + c0 + 1)));
                }
                o = make_uint4(ow[0], ow[1], ow[2], ow[3]);
            }
            *reinterpret_cast<uint4 *>(Vs + (size_t)pi * sv + c) = o;
        }
    }
    if (MODE == MODE_VPV && active) {
        // P[j][k] = E[b][i = line][j][k] (bf16 scratch written by the energy pass), rows j >= L are zero
        // general H x W (the reference's view chain, common.py:3763-3775): output column n' uses the H energy rows of the
        // flat pixels n'*H .. n'*H + H-1 (row-major) -- for H == W that is image row n'
        const __nv_bfloat16 *E = reinterpret_cast<const __nv_bfloat16 *>(p.scratch) + ((size_t)b * p.H * p.W + (size_t)line * p.H) * LP;
        const int vecs = LP / 8;
        for (int idx = tis; idx < LP * vecs; idx += kThreads) {
            const int j = idx / vecs, c = (idx - j * vecs) * 8;
            uint4 o = make_uint4(0, 0, 0, 0);
            if (j < L) o = __ldg(reinterpret_cast<const uint4 *>(E + (size_t)j * LP + c));
            *reinterpret_cast<uint4 *>(Ps + (size_t)j * sp + c) = o;
        }
    }
    __syncthreads();
    if (!active) return;

    const int row0 = warp * 16;                                  // this warp's query rows
    float m_row[2] = {0.0f, 0.0f}, s_row[2] = {1.0f, 1.0f};
    if (MODE != MODE_VPV) {
        // ---- E = Aq . Bk^T for rows row0..row0+15, all LP columns ----
        float e[NT][4];
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
        if (MODE == MODE_VE) {
            // raw energies -> bf16 scratch [b][i][j = line][k]; padded k columns are exact zeros (zero Bk rows)
            __nv_bfloat16 *E = reinterpret_cast<__nv_bfloat16 *>(p.scratch);
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int i = row0 + g + 8 * hh;
                if (i < L) {
                    uint32_t *dst = reinterpret_cast<uint32_t *>(E + (((size_t)b * p.H + i) * p.W + line) * LP) + t4;
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) dst[nt * 4] = pack_bf16x2(e[nt][2 * hh], e[nt][2 * hh + 1]);
                }
            }
            return;
        }
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
            float s = 0.0f;
            const int r = row0 + g + 8 * hh;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const int col = nt * 8 + 2 * t4;
                const float p0 = col < L ? __expf(e[nt][2 * hh] - m) : 0.0f;
                const float p1 = col + 1 < L ? __expf(e[nt][2 * hh + 1] - m) : 0.0f;
                s += p0 + p1;
                *reinterpret_cast<uint32_t *>(Ps + (size_t)r * sp + col) = pack_bf16x2(p0, p1);
            }
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            m_row[hh] = m;
            s_row[hh] = s;
        }
        __syncwarp();
    }

    // ---- O = P . V in chunks of 32 channels; epilogue per chunk ----
    const uint32_t pa_base = smem_addr(Ps + (size_t)(row0 + (lane & 15)) * sp + (lane >> 4) * 8);
    const uint32_t vb_base = smem_addr(Vs + (size_t)((lane & 7) + ((lane >> 3) & 1) * 8) * sv + (lane >> 4) * 8);
    for (int c0 = 0; c0 < C; c0 += 32) {
        float o[4][4];
#pragma unroll
        for (int ct = 0; ct < 4; ++ct) o[ct][0] = o[ct][1] = o[ct][2] = o[ct][3] = 0.0f;
#pragma unroll
        for (int kt = 0; kt < NT / 2; ++kt) {
            uint32_t a0, a1, a2, a3;
            ldsm_x4(pa_base + kt * 32, a0, a1, a2, a3);
#pragma unroll
            for (int cp = 0; cp < 2; ++cp) {
                uint32_t b0, b1, b2, b3;
                ldsm_x4_trans(vb_base + (uint32_t)(kt * 16 * sv * 2) + (uint32_t)((c0 + cp * 16) * 2), b0, b1, b2, b3);
                mma_bf16(o[2 * cp], a0, a1, a2, a3, b0, b1);
                mma_bf16(o[2 * cp + 1], a0, a1, a2, a3, b2, b3);
            }
        }
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int i = row0 + g + 8 * hh;
            if (i >= L) continue;
            const size_t px = pix_of(i);
            if (MODE == MODE_ROW) {
                float *sc = p.scratch + px * (C + 4);
#pragma unroll
                for (int ct = 0; ct < 4; ++ct)
                    *reinterpret_cast<float2 *>(sc + c0 + ct * 8 + 2 * t4) = make_float2(o[ct][2 * hh], o[ct][2 * hh + 1]);
                if (c0 == 0 && t4 == 0) { sc[C] = m_row[hh]; sc[C + 1] = s_row[hh]; }
            } else {
                float fh = 1.0f, fw = 0.0f, inv = 1.0f;
                const float *sc = p.scratch + px * (C + 4);
                if (MODE == MODE_COL) {
                    const float mw = sc[C], sw = sc[C + 1], mh = m_row[hh], sh = s_row[hh];
                    const float m = fmaxf(mh, mw);
                    fh = __expf(mh - m);
                    fw = __expf(mw - m);
                    inv = 1.0f / (sh * fh + sw * fw);
                }
#pragma unroll
                for (int ct = 0; ct < 4; ++ct) {
                    const int c = c0 + ct * 8 + 2 * t4;
                    float o0 = o[ct][2 * hh], o1 = o[ct][2 * hh + 1];
                    if (MODE == MODE_COL) {
                        const float2 ow = *reinterpret_cast<const float2 *>(sc + c);
                        o0 = (o0 * fh + ow.x * fw) * inv;
                        o1 = (o1 * fh + ow.y * fw) * inv;
                    }
                    const float2 xv = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t *>(p.x + px * p.x_cs + p.x_off + c)));
                    *reinterpret_cast<uint32_t *>(p.out + px * p.out_cs + p.out_off + c) =
                        pack_bf16x2(fmaf(p.gamma, o0, xv.x), fmaf(p.gamma, o1, xv.y));
                }
            }
        }
    }
}

int pick_nt(int L) {
    static const int opts[] = {2, 4, 6, 8, 10, 12, 16, 20};
    for (int nt : opts)
        if (nt * 8 >= L) return nt;
    return 0;
}

template <int MODE, int NT>
int launch_nt(const AttnParams &p, int L, cudaStream_t st) {
    constexpr int LPC = NT <= 4 ? 4 : (NT <= 8 ? 2 : 1);      // lines per CTA: keep CTAs at >= 4 warps
    Geom gm;
    gm.L = L;
    gm.LP = NT * 8;
    gm.KQ = (3 * p.Cq + 15) / 16 * 16;
    gm.sq = gm.KQ + 8;
    gm.sv = p.C + 8;
    gm.sp = gm.LP + 8;
    size_t smem = 0;
    if (MODE != MODE_VPV) smem += 2 * (size_t)gm.LP * gm.sq * 2;
    if (MODE != MODE_VE) smem += (size_t)gm.LP * gm.sv * 2 + (size_t)gm.LP * gm.sp * 2;
    smem = (smem + 15) & ~size_t(15);
    if (smem * LPC > 227 * 1024) return 1;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(attn_mma_kernel<MODE, NT, LPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        attr_set = true;
    }
    const int lines = (MODE == MODE_ROW) ? p.H : p.W;
    const int total = p.B * lines;
    launch_pdl(attn_mma_kernel<MODE, NT, LPC>, dim3((total + LPC - 1) / LPC), dim3(NT * 16 * LPC), smem * LPC, st, p, gm, total, smem);
    return 0;
}

template <int MODE>
int launch_mode(const AttnParams &p, cudaStream_t st) {
    const int L = (MODE == MODE_ROW) ? p.W : p.H;
    if (p.C % 32 != 0) return 1;
    switch (pick_nt(L)) {
        case 2: return launch_nt<MODE, 2>(p, L, st);
        case 4: return launch_nt<MODE, 4>(p, L, st);
        case 6: return launch_nt<MODE, 6>(p, L, st);
        case 8: return launch_nt<MODE, 8>(p, L, st);
        case 10: return launch_nt<MODE, 10>(p, L, st);
        case 12: return launch_nt<MODE, 12>(p, L, st);
        case 16: return launch_nt<MODE, 16>(p, L, st);
        case 20: return launch_nt<MODE, 20>(p, L, st);
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
    if (g > (size_t)kNumSMs * 16) g = (size_t)kNumSMs * 16;
    launch_pdl(attn_qk_kernel, dim3((int)g), dim3(256), 0, st, x, x_cs, x_off, Cq, npix, wq, bq, wk, bk, s, t, q, k);
}

// crisscross: [B*H*W][C + 4] fp32 row-pass partials; vertical: [B][H][W][LP] bf16 energies -- one shared region
size_t attn_scratch_bytes(int B, int H, int W, int C) {
    const size_t cc = (size_t)B * H * W * (C + 4) * 4;
    const size_t ve = (size_t)B * H * W * (size_t)(pick_nt(H) * 8) * 2;
    return cc > ve ? cc : ve;
}

int crisscross_launch(const AttnParams &p, cudaStream_t st) {
    if (launch_mode<MODE_ROW>(p, st)) return 1;
    return launch_mode<MODE_COL>(p, st);
}

int vertical_launch(const AttnParams &p, cudaStream_t st) {
    if (launch_mode<MODE_VE>(p, st)) return 1;
    return launch_mode<MODE_VPV>(p, st);
}

}  // namespace ry
