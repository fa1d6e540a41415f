// Attention kernels of CCVA (reference models/common.py:3675-3786).
//
//  attn_qk     : q = ReLU6(bn(SiLU(gconv_q(x)))), k likewise with the SAME bn (common.py:3693-3701); the grouped 1x1 convs
//                have 8 input channels per output channel (DWConv(c, c/8): groups = c/8, common.py:154-156).
//  crisscross  : eH[h,w,g] = q[h,w].k[g,w], eW[h,w,g] = q[h,w].k[h,g]; softmax over the H+W energies (no -inf diagonal);
//                out = sum_g v[g,w] aH + sum_g v[h,g] aW; gamma*out + x                     (common.py:3704-3726)
//                Done as two "line attention" passes that are merged like a split soft-max:
//                  row pass  (CTA = one image row)    -> un-normalised partial O_W, running max m_W, sum s_W (scratch)
//                  col pass  (CTA = one image column) -> O_H, m_H, s_H, merge with the row partials, gamma*out + x
//  vertical    : out[y,x] = sum_g v[g,x] * (q[x,y].k[g,y])  (raw energies, H == W; common.py:3763-3778)
//                CTA = one output column x.
//  v = ReLU6(bn1(SiLU(wv*x + bv))) is recomputed from x where needed (depthwise 1x1: purely per element).
// All arithmetic is fp32 on bf16 inputs; the [L x L] . [L x C] product is register-tiled from shared memory.
#include "memops.cuh"

#include "common.cuh"

namespace ry {

namespace {

constexpr int kAttnThreads = 256;
constexpr int kMaxPix = 10;   // query pixels per thread per chunk

enum { MODE_ROW = 0, MODE_COL = 1, MODE_VERT = 2 };

__global__ void __launch_bounds__(256) attn_qk_kernel(const __nv_bfloat16 *__restrict__ x, int x_cs, int x_off, int Cq,
                                                      size_t npix, const float *__restrict__ wq, const float *__restrict__ bq,
                                                      const float *__restrict__ wk, const float *__restrict__ bk,
                                                      const float *__restrict__ s, const float *__restrict__ t,
                                                      float *__restrict__ q, float *__restrict__ k) {
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

// shared memory: Q[L][Cq] | K[L][Cq+1] | V[L][C] | E[L][L+1] | m[L] | s[L]
template <int MODE>
__global__ void __launch_bounds__(kAttnThreads) attn_line_kernel(const AttnParams p) {
    extern __shared__ float sm[];
    const int L = (MODE == MODE_ROW) ? p.W : p.H;           // line length (MODE_VERT: H == W)
    const int lines = (MODE == MODE_ROW) ? p.H : p.W;       // lines per image
    const int b = blockIdx.x / lines, line = blockIdx.x % lines;
    const int C = p.C, Cq = p.Cq, EP = L + 1, KP = Cq + 1;   // padded strides: no bank conflicts
    float *Q = sm, *K = Q + L * Cq, *V = K + ((L * KP + 3) & ~3) + ((4 - ((L * Cq) & 3)) & 3), *E = V + (size_t)L * C, *mrow = E + (size_t)L * EP, *srow = mrow + L;
    const size_t img = (size_t)b * p.H * p.W;
    // pixel of line position i (values / outputs)
    auto pix_of = [&](int i) -> size_t {
        return (MODE == MODE_ROW) ? img + (size_t)line * p.W + i : img + (size_t)i * p.W + line;
    };
    // ---- stage Q, K (ROW/COL: along the line; VERT: Q = q of image row `line`) and V ----
    for (int i = threadIdx.x; i < L * Cq; i += blockDim.x) {
        const int pi = i / Cq, d = i - pi * Cq;
        if (MODE == MODE_VERT) {
            Q[i] = __ldg(p.q + (img + (size_t)line * p.W + pi) * Cq + d);
        } else {
            const size_t px = pix_of(pi);
            Q[i] = __ldg(p.q + px * Cq + d);
            K[pi * KP + d] = __ldg(p.k + px * Cq + d);
        }
    }
    const int vecs = C / 8;
    for (int i = threadIdx.x; i < L * vecs; i += blockDim.x) {
        const int pi = i / vecs, c = (i - pi * vecs) * 8;
        const uint4 u = __ldg(reinterpret_cast<const uint4 *>(p.x + pix_of(pi) * p.x_cs + p.x_off + c));
        const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const float2 f = unpack_bf16x2(uw[h]);
            const int c0 = c + 2 * h;
            V[(size_t)pi * C + c0] = relu6_f(fmaf(__ldg(p.s1 + c0), silu_f(fmaf(__ldg(p.wv + c0), f.x, __ldg(p.bv + c0))), __ldg(p.t1 + c0)));
            V[(size_t)pi * C + c0 + 1] =
                relu6_f(fmaf(__ldg(p.s1 + c0 + 1), silu_f(fmaf(__ldg(p.wv + c0 + 1), f.y, __ldg(p.bv + c0 + 1))), __ldg(p.t1 + c0 + 1)));
        }
    }
    __syncthreads();
    // ---- energies E[i][g] ----
    for (int e = threadIdx.x; e < L * L; e += blockDim.x) {
        float acc = 0.0f;
        if (MODE == MODE_VERT) {
            const int g = e / L, i = e - g * L;            // i fastest: consecutive threads read consecutive pixels of k
            const float *kp = p.k + (img + (size_t)g * p.W + i) * Cq;
            for (int d = 0; d < Cq; ++d) acc = fmaf(Q[i * Cq + d], __ldg(kp + d), acc);
            E[(size_t)i * EP + g] = acc;
        } else {
            const int i = e / L, g = e - i * L;
            for (int d = 0; d < Cq; ++d) acc = fmaf(Q[i * Cq + d], K[g * KP + d], acc);
            E[(size_t)i * EP + g] = acc;
        }
    }
    __syncthreads();
    // ---- soft-max statistics per query row (not for VERT: raw energies are used) ----
    if (MODE != MODE_VERT) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
        for (int i = warp; i < L; i += nwarps) {
            float m = -INFINITY;
            for (int g = lane; g < L; g += 32) m = fmaxf(m, E[(size_t)i * EP + g]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            float s = 0.0f;
            for (int g = lane; g < L; g += 32) {
                const float pe = __expf(E[(size_t)i * EP + g] - m);
                E[(size_t)i * EP + g] = pe;
                s += pe;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) { mrow[i] = m; srow[i] = s; }
        }
        __syncthreads();
    }
    // ---- O[i][c] = sum_g E[i][g] V[g][c]; thread = 4 channels x up to kMaxPix query pixels ----
    const int ncg = C / 4;                                  // channel groups
    const int npg = blockDim.x / ncg;                       // pixel groups working in parallel (>= 1 for C <= 1024)
    const int cg = threadIdx.x % ncg, pg = threadIdx.x / ncg;
    if (pg < npg) {
        for (int base = 0; base < L; base += npg * kMaxPix) {
            float acc[kMaxPix][4];
#pragma unroll
            for (int r = 0; r < kMaxPix; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.0f;
            for (int g = 0; g < L; ++g) {
                const float4 v = *reinterpret_cast<const float4 *>(V + (size_t)g * C + cg * 4);
#pragma unroll
                for (int r = 0; r < kMaxPix; ++r) {
                    const int i = base + pg + r * npg;
                    const float e = (i < L) ? E[(size_t)i * EP + g] : 0.0f;
                    acc[r][0] = fmaf(e, v.x, acc[r][0]);
                    acc[r][1] = fmaf(e, v.y, acc[r][1]);
                    acc[r][2] = fmaf(e, v.z, acc[r][2]);
                    acc[r][3] = fmaf(e, v.w, acc[r][3]);
                }
            }
#pragma unroll
            for (int r = 0; r < kMaxPix; ++r) {
                const int i = base + pg + r * npg;
                if (i >= L) continue;
                const size_t px = pix_of(i);
                const int c = cg * 4;
                if (MODE == MODE_ROW) {
                    float *sc = p.scratch + px * (C + 4);
                    *reinterpret_cast<float4 *>(sc + c) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
                    if (cg == 0) { sc[C] = mrow[i]; sc[C + 1] = srow[i]; }
                } else {
                    float o[4] = {acc[r][0], acc[r][1], acc[r][2], acc[r][3]};
                    if (MODE == MODE_COL) {
                        const float *sc = p.scratch + px * (C + 4);
                        const float4 ow = *reinterpret_cast<const float4 *>(sc + c);
                        const float mw = sc[C], sw = sc[C + 1], mh = mrow[i], sh = srow[i];
                        const float m = fmaxf(mh, mw);
                        const float fh = __expf(mh - m), fw = __expf(mw - m);
                        const float inv = 1.0f / (sh * fh + sw * fw);
                        o[0] = (o[0] * fh + ow.x * fw) * inv;
                        o[1] = (o[1] * fh + ow.y * fw) * inv;
                        o[2] = (o[2] * fh + ow.z * fw) * inv;
                        o[3] = (o[3] * fh + ow.w * fw) * inv;
                    }
                    const uint2 xu = __ldg(reinterpret_cast<const uint2 *>(p.x + px * p.x_cs + p.x_off + c));
                    const float2 x0 = unpack_bf16x2(xu.x), x1 = unpack_bf16x2(xu.y);
                    uint2 res;
                    res.x = pack_bf16x2(fmaf(p.gamma, o[0], x0.x), fmaf(p.gamma, o[1], x0.y));
                    res.y = pack_bf16x2(fmaf(p.gamma, o[2], x1.x), fmaf(p.gamma, o[3], x1.y));
                    *reinterpret_cast<uint2 *>(p.out + px * p.out_cs + p.out_off + c) = res;
                }
            }
        }
    }
}

size_t attn_smem_bytes(int L, int C, int Cq) {
    return ((size_t)L * (2 * Cq + 1) + 8 + (size_t)L * C + (size_t)L * (L + 1) + 2 * L) * sizeof(float);
}

template <int MODE>
int launch_line(const AttnParams &p, cudaStream_t st) {
    const int L = (MODE == MODE_ROW) ? p.W : p.H;
    const int lines = (MODE == MODE_ROW) ? p.H : p.W;
    const size_t smem = attn_smem_bytes(L, p.C, p.Cq);
    if (smem > 227 * 1024 || p.C % 8 != 0 || p.C / 4 > kAttnThreads) return 1;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(attn_line_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        attr_set = true;
    }
    attn_line_kernel<MODE><<<p.B * lines, kAttnThreads, smem, st>>>(p);
    return 0;
}

}  // namespace

void attn_qk_launch(const __nv_bfloat16 *x, int x_cs, int x_off, int C, int Cq, size_t npix, const float *wq,
                    const float *bq, const float *wk, const float *bk, const float *s, const float *t, float *q, float *k,
                    cudaStream_t st) {
    (void)C;
    const size_t total = npix * Cq;
    size_t g = (total + 255) / 256;
    if (g > (size_t)kNumSMs * 16) g = (size_t)kNumSMs * 16;
    attn_qk_kernel<<<(int)g, 256, 0, st>>>(x, x_cs, x_off, Cq, npix, wq, bq, wk, bk, s, t, q, k);
}

size_t crisscross_scratch_floats(int B, int H, int W, int C) { return (size_t)B * H * W * (C + 4); }

int crisscross_launch(const AttnParams &p, cudaStream_t st) {
    if (launch_line<MODE_ROW>(p, st)) return 1;
    return launch_line<MODE_COL>(p, st);
}

int vertical_launch(const AttnParams &p, cudaStream_t st) {
    if (p.H != p.W) return 2;   // the reference's view chain scrambles indices for H != W: not built yet
    return launch_line<MODE_VERT>(p, st);
}

}  // namespace ry
