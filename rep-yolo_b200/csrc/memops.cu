// Memory-bound kernels of the deploy graph: stem conv (fp32 NCHW image -> NHWC bf16), depthwise 5x5, 2x2 max-pool,
// SPP 5/9/13 pools, nearest x2 up-sampling, channel attention (global average + two tiny FCs).
// All activations are NHWC bf16; every access is a 16-byte vector of 8 channels; outputs go to (tensor, channel-offset)
// views so that concatenations never materialise.
#include "memops.cuh"

#include "common.cuh"

namespace ry {

namespace {

__device__ __forceinline__ uint4 ldg16(const __nv_bfloat16 *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }
__device__ __forceinline__ void stg16(__nv_bfloat16 *p, uint4 v) { *reinterpret_cast<uint4 *>(p) = v; }

// ------------------------------------------------------------------------------------------------------------------
// Stem: RepS_Block L0 deploy branch (reference models/common.py:3412-3416): SiLU(conv3x3 s2 p1 (x) + b), Cin = 3.
// One thread = one output pixel, all COUT channels in registers; weights [27][COUT] fp32 in shared memory.
// ------------------------------------------------------------------------------------------------------------------
template <int COUT>
__global__ void __launch_bounds__(128) stem_kernel(const float *__restrict__ img, const float *__restrict__ w,
                                                   const float *__restrict__ bias, __nv_bfloat16 *__restrict__ out,
                                                   int B, int H, int W, int out_cs, int out_off) {
    __shared__ float sw[27 * COUT];
    __shared__ float sb[COUT];
    for (int i = threadIdx.x; i < 27 * COUT; i += blockDim.x) sw[i] = w[i];
    for (int i = threadIdx.x; i < COUT; i += blockDim.x) sb[i] = bias[i];
    __syncthreads();
    const int Ho = H / 2, Wo = W / 2;
    const size_t total = (size_t)B * Ho * Wo;
    for (size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x; pix < total; pix += (size_t)gridDim.x * blockDim.x) {
        const int wo = (int)(pix % Wo);
        const int ho = (int)((pix / Wo) % Ho);
        const int b = (int)(pix / ((size_t)Wo * Ho));
        float acc[COUT];
#pragma unroll
        for (int c = 0; c < COUT; ++c) acc[c] = sb[c];
        const float *ib = img + (size_t)b * 3 * H * W;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) {
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const int hi = 2 * ho + kh - 1;
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const int wi = 2 * wo + kw - 1;
                    float x = 0.0f;
                    if (hi >= 0 && hi < H && wi >= 0 && wi < W) x = __ldg(ib + ((size_t)ci * H + hi) * W + wi);
                    const float *wr = sw + ((ci * 3 + kh) * 3 + kw) * COUT;
#pragma unroll
                    for (int c = 0; c < COUT; ++c) acc[c] = fmaf(x, wr[c], acc[c]);
                }
            }
        }
        __nv_bfloat16 *o = out + pix * out_cs + out_off;
#pragma unroll
        for (int c = 0; c < COUT; c += 8) {
            float s[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) s[i] = silu_f(acc[c + i]);
            stg16(o + c, make_uint4(pack_bf16x2(s[0], s[1]), pack_bf16x2(s[2], s[3]), pack_bf16x2(s[4], s[5]),
                                    pack_bf16x2(s[6], s[7])));
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Depthwise 5x5 s1 p2 + bias + act  (GSConv.cv2, reference models/common.py:3813, 3817).  Channels come as two halves
// (in_off0 / in_off1) and go to two halves (out_off0 / out_off1): the GSConv channel shuffle folded into addressing.
// weights: [25][C] fp32 (tap-major), one thread = 8 channels of one pixel.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dw5_kernel(const __nv_bfloat16 *__restrict__ in, int in_cs, int in_off0, int in_off1,
                                                  __nv_bfloat16 *__restrict__ out, int out_cs, int out_off0, int out_off1,
                                                  const float *__restrict__ w, const float *__restrict__ bias, int C,
                                                  int half, int B, int H, int W, int act) {
    const int vecs = C / 8;
    const size_t total = (size_t)B * H * W * vecs;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int v = (int)(i % vecs);
        const size_t pix = i / vecs;
        const int x = (int)(pix % W), y = (int)((pix / W) % H);
        const size_t img_base = (pix / ((size_t)W * H)) * (size_t)H * W;
        const int c = v * 8;
        const int ci = c < half ? in_off0 + c : in_off1 + (c - half);
        const int co = c < half ? out_off0 + c : out_off1 + (c - half);
        float acc[8];
        {
            const float4 b0 = __ldg(reinterpret_cast<const float4 *>(bias + c));
            const float4 b1 = __ldg(reinterpret_cast<const float4 *>(bias + c + 4));
            acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w;
            acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
        }
#pragma unroll
        for (int dy = -2; dy <= 2; ++dy) {
            const int yy = y + dy;
            if (yy < 0 || yy >= H) continue;
#pragma unroll
            for (int dx = -2; dx <= 2; ++dx) {
                const int xx = x + dx;
                if (xx < 0 || xx >= W) continue;
                const uint4 u = ldg16(in + (img_base + (size_t)yy * W + xx) * in_cs + ci);
                const float *wt = w + ((dy + 2) * 5 + (dx + 2)) * C + c;
                const float4 w0 = __ldg(reinterpret_cast<const float4 *>(wt));
                const float4 w1 = __ldg(reinterpret_cast<const float4 *>(wt + 4));
                const float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
                acc[0] = fmaf(f0.x, w0.x, acc[0]); acc[1] = fmaf(f0.y, w0.y, acc[1]);
                acc[2] = fmaf(f1.x, w0.z, acc[2]); acc[3] = fmaf(f1.y, w0.w, acc[3]);
                acc[4] = fmaf(f2.x, w1.x, acc[4]); acc[5] = fmaf(f2.y, w1.y, acc[5]);
                acc[6] = fmaf(f3.x, w1.z, acc[6]); acc[7] = fmaf(f3.y, w1.w, acc[7]);
            }
        }
        if (act == 1) {
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = silu_f(acc[k]);
        }
        stg16(out + pix * out_cs + co, make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]),
                                                  pack_bf16x2(acc[4], acc[5]), pack_bf16x2(acc[6], acc[7])));
    }
}

// MP (reference models/common.py:32-38): MaxPool2d(2, 2).  One thread = 8 channels of one output pixel.
__global__ void __launch_bounds__(256) maxpool2_kernel(const __nv_bfloat16 *__restrict__ in, int in_cs, int in_off,
                                                       __nv_bfloat16 *__restrict__ out, int out_cs, int out_off, int C,
                                                       int B, int H, int W) {
    const int vecs = C / 8, Ho = H / 2, Wo = W / 2;
    const size_t total = (size_t)B * Ho * Wo * vecs;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int v = (int)(i % vecs);
        const size_t opix = i / vecs;
        const int xo = (int)(opix % Wo), yo = (int)((opix / Wo) % Ho);
        const size_t b = opix / ((size_t)Wo * Ho);
        const __nv_bfloat16 *p = in + ((b * H + 2 * yo) * W + 2 * xo) * in_cs + in_off + v * 8;
        const uint4 a = ldg16(p), c = ldg16(p + in_cs), d = ldg16(p + (size_t)W * in_cs), e = ldg16(p + (size_t)(W + 1) * in_cs);
        stg16(out + opix * out_cs + out_off + v * 8, bf16x8_max(bf16x8_max(a, c), bf16x8_max(d, e)));
    }
}

// SPPCSPC pools (reference models/common.py:279, 286): MaxPool2d(k, 1, k//2) for k = 5, 9, 13 (-inf padding), one read
// of the 13x13 neighbourhood, three writes at channel offsets of the cv5 input buffer.
__global__ void __launch_bounds__(256) spp_kernel(const __nv_bfloat16 *__restrict__ in, int in_cs, int in_off,
                                                  __nv_bfloat16 *__restrict__ out, int out_cs, int off5, int off9,
                                                  int off13, int C, int B, int H, int W) {
    const int vecs = C / 8;
    const size_t total = (size_t)B * H * W * vecs;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int v = (int)(i % vecs);
        const size_t pix = i / vecs;
        const int x = (int)(pix % W), y = (int)((pix / W) % H);
        const size_t img_base = (pix / ((size_t)W * H)) * (size_t)H * W;
        const uint4 ctr = ldg16(in + pix * in_cs + in_off + v * 8);
        uint4 m5 = ctr, m9 = ctr, m13 = ctr;
        for (int dy = -6; dy <= 6; ++dy) {
            const int yy = y + dy;
            if (yy < 0 || yy >= H) continue;
            const int ady = dy < 0 ? -dy : dy;
            for (int dx = -6; dx <= 6; ++dx) {
                const int xx = x + dx;
                if (xx < 0 || xx >= W) continue;
                const int adx = dx < 0 ? -dx : dx;
                const int d = ady > adx ? ady : adx;
                const uint4 u = ldg16(in + (img_base + (size_t)yy * W + xx) * in_cs + in_off + v * 8);
                m13 = bf16x8_max(m13, u);
                if (d <= 4) m9 = bf16x8_max(m9, u);
                if (d <= 2) m5 = bf16x8_max(m5, u);
            }
        }
        __nv_bfloat16 *o = out + pix * out_cs + v * 8;
        stg16(o + off5, m5);
        stg16(o + off9, m9);
        stg16(o + off13, m13);
    }
}

// nn.Upsample(None, 2, 'nearest') (reference cfg/training/Rep-YOLO.yaml:43,51).
__global__ void __launch_bounds__(256) upsample2_kernel(const __nv_bfloat16 *__restrict__ in, int in_cs, int in_off,
                                                        __nv_bfloat16 *__restrict__ out, int out_cs, int out_off, int C,
                                                        int B, int H, int W) {
    const int vecs = C / 8, Ho = H * 2, Wo = W * 2;
    const size_t total = (size_t)B * Ho * Wo * vecs;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int v = (int)(i % vecs);
        const size_t opix = i / vecs;
        const int xo = (int)(opix % Wo), yo = (int)((opix / Wo) % Ho);
        const size_t b = opix / ((size_t)Wo * Ho);
        stg16(out + opix * out_cs + out_off + v * 8,
              ldg16(in + ((b * H + yo / 2) * W + xo / 2) * in_cs + in_off + v * 8));
    }
}

// CA (reference models/common.py:3797-3802): p = avgpool(x); out = p * sigmoid(f2(relu(f1(p)))) + p  -> [B, C] fp32.
// One CTA per image; fp32 accumulation of the mean.
__global__ void __launch_bounds__(512) ca_kernel(const __nv_bfloat16 *__restrict__ in, int in_cs, int in_off,
                                                 float *__restrict__ out, int out_cs, int out_off, const float *__restrict__ f1,
                                                 const float *__restrict__ f2, int C, int HW) {
    extern __shared__ float sm[];            // [C] mean, [C/16] hidden, [512*8] partials
    float *mean = sm, *hid = sm + C, *part = hid + C / 16;
    const int b = blockIdx.x, vecs = C / 8;
    const int groups = blockDim.x / vecs;    // pixel groups working in parallel
    const int v = threadIdx.x % vecs, g = threadIdx.x / vecs;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (g < groups) {
        const __nv_bfloat16 *base = in + (size_t)b * HW * in_cs + in_off + v * 8;
        for (int p = g; p < HW; p += groups) {
            const uint4 u = ldg16(base + (size_t)p * in_cs);
            const float2 f0 = unpack_bf16x2(u.x), f1v = unpack_bf16x2(u.y), f2v = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
            acc[0] += f0.x; acc[1] += f0.y; acc[2] += f1v.x; acc[3] += f1v.y;
            acc[4] += f2v.x; acc[5] += f2v.y; acc[6] += f3.x; acc[7] += f3.y;
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) part[threadIdx.x * 8 + k] = acc[k];
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const int vv = c / 8, k = c % 8;
        float s = 0.0f;
        for (int gg = 0; gg < groups; ++gg) s += part[(gg * vecs + vv) * 8 + k];
        mean[c] = s / (float)HW;
    }
    __syncthreads();
    const int Ch = C / 16;
    for (int j = threadIdx.x; j < Ch; j += blockDim.x) {
        float s = 0.0f;
        for (int c = 0; c < C; ++c) s = fmaf(__ldg(f1 + (size_t)j * C + c), mean[c], s);
        hid[j] = fmaxf(s, 0.0f);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.0f;
        for (int j = 0; j < Ch; ++j) s = fmaf(__ldg(f2 + (size_t)c * Ch + j), hid[j], s);
        const float a = 1.0f / (1.0f + __expf(-s));
        out[(size_t)b * out_cs + out_off + c] = mean[c] * a + mean[c];
    }
}

inline int grid_for(size_t total, int block) {
    size_t g = (total + block - 1) / block;
    const size_t cap = (size_t)kNumSMs * 16;
    return (int)(g < cap ? (g ? g : 1) : cap);
}

}  // namespace

int stem_launch(const float *img, const float *w27, const float *bias, __nv_bfloat16 *out, int out_cs, int out_off,
                int cout, int B, int H, int W, cudaStream_t st) {
    const size_t total = (size_t)B * (H / 2) * (W / 2);
    const int grid = grid_for(total, 128);
    switch (cout) {
        case 16: stem_kernel<16><<<grid, 128, 0, st>>>(img, w27, bias, out, B, H, W, out_cs, out_off); break;
        case 32: stem_kernel<32><<<grid, 128, 0, st>>>(img, w27, bias, out, B, H, W, out_cs, out_off); break;
        case 48: stem_kernel<48><<<grid, 128, 0, st>>>(img, w27, bias, out, B, H, W, out_cs, out_off); break;
        case 64: stem_kernel<64><<<grid, 128, 0, st>>>(img, w27, bias, out, B, H, W, out_cs, out_off); break;
        default: return 1;
    }
    return 0;
}

void dw5_launch(const __nv_bfloat16 *in, int in_cs, int in_off0, int in_off1, __nv_bfloat16 *out, int out_cs, int out_off0,
                int out_off1, const float *w, const float *bias, int C, int half, int B, int H, int W, int act,
                cudaStream_t st) {
    const size_t total = (size_t)B * H * W * (C / 8);
    dw5_kernel<<<grid_for(total, 256), 256, 0, st>>>(in, in_cs, in_off0, in_off1, out, out_cs, out_off0, out_off1, w, bias,
                                                     C, half, B, H, W, act);
}

void maxpool2_launch(const __nv_bfloat16 *in, int in_cs, int in_off, __nv_bfloat16 *out, int out_cs, int out_off, int C,
                     int B, int H, int W, cudaStream_t st) {
    const size_t total = (size_t)B * (H / 2) * (W / 2) * (C / 8);
    maxpool2_kernel<<<grid_for(total, 256), 256, 0, st>>>(in, in_cs, in_off, out, out_cs, out_off, C, B, H, W);
}

void spp_launch(const __nv_bfloat16 *in, int in_cs, int in_off, __nv_bfloat16 *out, int out_cs, int off5, int off9,
                int off13, int C, int B, int H, int W, cudaStream_t st) {
    const size_t total = (size_t)B * H * W * (C / 8);
    spp_kernel<<<grid_for(total, 256), 256, 0, st>>>(in, in_cs, in_off, out, out_cs, off5, off9, off13, C, B, H, W);
}

void upsample2_launch(const __nv_bfloat16 *in, int in_cs, int in_off, __nv_bfloat16 *out, int out_cs, int out_off, int C,
                      int B, int H, int W, cudaStream_t st) {
    const size_t total = (size_t)B * (2 * H) * (2 * W) * (C / 8);
    upsample2_kernel<<<grid_for(total, 256), 256, 0, st>>>(in, in_cs, in_off, out, out_cs, out_off, C, B, H, W);
}

void ca_launch(const __nv_bfloat16 *in, int in_cs, int in_off, float *out, int out_cs, int out_off, const float *f1,
               const float *f2, int C, int B, int HW, cudaStream_t st) {
    const int threads = 512;
    const size_t smem = (size_t)(C + C / 16 + threads * 8) * sizeof(float);
    ca_kernel<<<B, threads, smem, st>>>(in, in_cs, in_off, out, out_cs, out_off, f1, f2, C, HW);
}

}  // namespace ry
