// Memory-bound kernels of the deploy graph: stem conv (fp32 NCHW image -> NHWC bf16), depthwise 5x5, 2x2 max-pool,
// SPP 5/9/13 pools, nearest x2 up-sampling, channel attention (global average + two tiny FCs).
// All activations are NHWC bf16; every access is a 16-byte vector of 8 channels; outputs go to (tensor, channel-offset)
// views so that concatenations never materialise.
#include "memops.cuh"

#include <algorithm>

#include "common.cuh"
#include "ptx.cuh"

namespace ry {

namespace {

__device__ __forceinline__ float tanh_approx_f(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint4 ldg16(const __nv_bfloat16 *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }
__device__ __forceinline__ void stg16(__nv_bfloat16 *p, uint4 v) { *reinterpret_cast<uint4 *>(p) = v; }

// ------------------------------------------------------------------------------------------------------------------
// Stem: RepS_Block L0 deploy branch (reference models/common.py:3412-3416): SiLU(conv3x3 s2 p1 (x) + b), Cin = 3, on the
// fp32 NCHW image -> NHWC bf16.  K = 27 (padded to 32) is far too small for a tcgen05 tile pipeline but 1296 FMA per
// pixel would make a SIMT kernel FMA-bound at ~1.6x the HBM floor, so the product runs on mma.sync m16n8k16 (bf16 in,
// fp32 accumulate): one CTA tile = 2 output rows x 64 output pixels; the 5 x 130 x 3 input patch is staged in shared
// memory as bf16, each warp gathers the im2col A fragments of its 16 pixels with 16-bit shared loads, the weight B
// fragments live in registers for the whole kernel; bias + SiLU; the tile goes out through shared memory as coalesced
// 16-byte stores.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kStemTW = 64, kStemTH = 8, kStemPW = 2 * kStemTW + 2, kStemPitch = 136;   // patch row: 130 used of 136
// one CTA tile = 8 output rows x 64 pixels: the 17-row input patch is fetched once (26 independent loads per thread),
// then four row pairs are computed, staged and stored

__device__ __forceinline__ void mma_bf16_16816(float *c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

constexpr int kStemLines = 3 * (2 * kStemTH + 1);                       // (channel, input row) lines of one patch
constexpr int kStemPatchFloats = kStemLines * kStemPitch;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, bool ok) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? 16 : 0) : "memory");   // !ok: zero fill
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const float *src, bool ok) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(ok ? 4 : 0) : "memory");   // !ok: zero fill
}

// U8 = true: the image is uint8 NCHW (0..255) as the reference's detect.py ships it to the device (detect.py:73-78); the
// /255 of detect.py:76 is fused here.  The patch is then fetched as 34 aligned 4-byte words per line.
template <int COUT, bool U8>
__global__ void __launch_bounds__(256) stem_kernel(const void *__restrict__ img_v, const float *__restrict__ w,
                                                   const float *__restrict__ bias, __nv_bfloat16 *__restrict__ out,
                                                   int B, int H, int W, int out_cs, int out_off) {
    pdl_trigger();
    constexpr int NT = COUT / 8;
    constexpr int kElem = U8 ? 1 : 4;                                          // bytes per patch element
    constexpr int kX0 = 3;                                                     // patch column 0 sits at this element offset: both
                                                                               // paths fetch ALIGNED words (4 x u8 / 4 x fp32) from column wi0 - 3 on
    extern __shared__ __align__(16) uint8_t stem_smem[];
    uint8_t *patch = stem_smem;                                                // 2 x [lines][pitch] elements, filled by cp.async
    __nv_bfloat16 *stage = reinterpret_cast<__nv_bfloat16 *>(stem_smem + 2 * kStemPatchFloats * kElem);   // [2 rows][64 px][COUT]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
    const int Ho = H / 2, Wo = W / 2;
    // ---- per-thread constants: patch offsets of this thread's 8 K indices, weight fragments, bias ----
    int koff[8];
    bool kval[8];
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = s * 16 + 2 * t4 + (j & 1) + (j >> 1) * 8;
            kval[s * 4 + j] = k < 27;
            const int ci = k / 9, kh = (k % 9) / 3, kw = k % 3;
            koff[s * 4 + j] = k < 27 ? (ci * (2 * kStemTH + 1) + kh) * kStemPitch + kw + kX0 : 0;
        }
    uint32_t bw[NT][2][2];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int k0 = s * 16 + 2 * t4 + h * 8, n = nt * 8 + g;
                const float w0 = k0 < 27 ? __ldg(w + k0 * COUT + n) : 0.0f, w1 = k0 + 1 < 27 ? __ldg(w + (k0 + 1) * COUT + n) : 0.0f;
                bw[nt][s][h] = pack_bf16x2(w0, w1);
            }
    float bb[NT][2];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) { bb[nt][0] = __ldg(bias + nt * 8 + 2 * t4); bb[nt][1] = __ldg(bias + nt * 8 + 2 * t4 + 1); }
    pdl_wait();
    const int tiles_x = cdiv(Wo, kStemTW), tiles_y = cdiv(Ho, kStemTH);
    const int total = B * tiles_y * tiles_x;
    const uint32_t patch_u = (uint32_t)__cvta_generic_to_shared(patch);

    auto prefetch = [&](int tile, int buf) {       // asynchronous fill of one patch buffer (zero fill outside the image)
        const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
        const int hi0 = 2 * ty * kStemTH - 1, wi0 = 2 * tx * kStemTW - 1;
        for (int r = warp; r < kStemLines; r += 8) {   // one warp per (channel, input row) line
            const int ci = r / (2 * kStemTH + 1), yy = hi0 + r - ci * (2 * kStemTH + 1);
            const bool row_ok = yy >= 0 && yy < H;
            const size_t row_base = (((size_t)b * 3 + ci) * H + (row_ok ? yy : 0)) * W;
            const uint32_t dst = patch_u + (uint32_t)(buf * kStemPatchFloats + r * kStemPitch) * kElem;
            if (U8) {
                const uint8_t *src = static_cast<const uint8_t *>(img_v) + row_base;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int wd = lane + 32 * j, xx = wi0 - 3 + 4 * wd;       // aligned 4-byte word: columns xx .. xx+3
                    const bool ok = row_ok && xx >= 0 && xx < W;
                    if (wd < 34) cp_async4(dst + wd * 4, reinterpret_cast<const float *>(src + (ok ? xx : 0)), ok);
                }
            } else {
                const float *src = static_cast<const float *>(img_v) + row_base;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int wd = lane + 32 * j, xx = wi0 - 3 + 4 * wd;       // aligned 16-byte word: columns xx .. xx+3 (W % 4 == 0:
                    const bool ok = row_ok && xx >= 0 && xx < W;                // a word is entirely inside or entirely outside the row)
                    if (wd < 34) cp_async16(dst + wd * 16, src + (ok ? xx : 0), ok);
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    int buf = 0;
    if (blockIdx.x < total) prefetch(blockIdx.x, 0);
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, buf ^= 1) {
        const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
        const int wo0 = tx * kStemTW, ho0 = ty * kStemTH;
        const int next = tile + gridDim.x;
        if (next < total) {
            prefetch(next, buf ^ 1);                      // overlaps this tile's compute
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const uint8_t *pb = patch + (size_t)buf * kStemPatchFloats * kElem;
        for (int rp = 0; rp < kStemTH / 2; ++rp) {
            // ---- warp = 16 consecutive pixels of one row of this row pair ----
            const int trow = warp / 4, px0 = (warp % 4) * 16;
            const int base = (2 * (2 * rp + trow)) * kStemPitch + 2 * px0;
            uint32_t a[2][4];
#pragma unroll
            for (int s = 0; s < 2; ++s)
#pragma unroll
                for (int h = 0; h < 2; ++h) {               // h: k pair (2t,2t+1) / (2t+8,2t+9)
#pragma unroll
                    for (int r = 0; r < 2; ++r) {           // r: pixel g / g+8
                        const int pix = base + 2 * (g + 8 * r);
                        float lo = 0.0f, hi = 0.0f;
                        if (U8) {
                            if (kval[s * 4 + 2 * h]) lo = (float)pb[koff[s * 4 + 2 * h] + pix] * (1.0f / 255.0f);
                            if (kval[s * 4 + 2 * h + 1]) hi = (float)pb[koff[s * 4 + 2 * h + 1] + pix] * (1.0f / 255.0f);
                        } else {
                            const float *pf = reinterpret_cast<const float *>(pb);
                            if (kval[s * 4 + 2 * h]) lo = pf[koff[s * 4 + 2 * h] + pix];
                            if (kval[s * 4 + 2 * h + 1]) hi = pf[koff[s * 4 + 2 * h + 1] + pix];
                        }
                        a[s][h * 2 + r] = pack_bf16x2(lo, hi);   // a0:(g,k lo) a1:(g+8,k lo) a2:(g,k hi) a3:(g+8,k hi)
                    }
                }
            float acc[NT][4];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                acc[nt][0] = acc[nt][2] = bb[nt][0];
                acc[nt][1] = acc[nt][3] = bb[nt][1];
#pragma unroll
                for (int s = 0; s < 2; ++s) mma_bf16_16816(acc[nt], a[s][0], a[s][1], a[s][2], a[s][3], bw[nt][s][0], bw[nt][s][1]);
            }
            // ---- SiLU -> bf16 -> staging [pixel][COUT] ----
            if (rp > 0) __syncthreads();                    // previous row pair's staging fully stored
            uint32_t *st = reinterpret_cast<uint32_t *>(stage);
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const int pix = trow * kStemTW + px0 + g + 8 * r;
                    // SiLU(x) = h + h * tanh(h), h = x / 2 (one MUFU; error below the bf16 rounding of the result)
                    const float h0 = 0.5f * acc[nt][2 * r], h1 = 0.5f * acc[nt][2 * r + 1];
                    st[(pix * COUT + nt * 8 + 2 * t4) >> 1] = pack_bf16x2(fmaf(h0, tanh_approx_f(h0), h0), fmaf(h1, tanh_approx_f(h1), h1));
                }
            __syncthreads();
            constexpr int CPP = COUT / 8;                   // 16-byte chunks per pixel
            constexpr int RowV = kStemTW * CPP;             // 16-byte chunks per staged row
            if (out_cs == COUT && wo0 + kStemTW <= Wo) {
                // dense output, full-width tile: a staged row is one contiguous run of the output map
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const int ho = ho0 + 2 * rp + rr;
                    if (ho < Ho) {
                        uint4 *dst = reinterpret_cast<uint4 *>(out + (((size_t)b * Ho + ho) * Wo + wo0) * COUT + out_off);
                        const uint4 *srcv = reinterpret_cast<const uint4 *>(stage) + rr * RowV;
                        for (int i = threadIdx.x; i < RowV; i += 256) dst[i] = srcv[i];
                    }
                }
            } else {
                for (int i = threadIdx.x; i < 2 * RowV; i += 256) {
                    const int pix = i / CPP, ch = i % CPP;
                    const int ho = ho0 + 2 * rp + pix / kStemTW, wo = wo0 + pix % kStemTW;
                    if (ho < Ho && wo < Wo)
                        stg16(out + (((size_t)b * Ho + ho) * Wo + wo) * out_cs + out_off + ch * 8, reinterpret_cast<const uint4 *>(stage)[i]);
                }
            }
        }
        __syncthreads();                                    // this buffer / staging are free for the next iteration
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Depthwise 5x5 s1 p2 + bias + act  (GSConv.cv2, reference models/common.py:3813, 3817).  Channels come as two halves
// (in_off0 / in_off1) and go to two halves (out_off0 / out_off1): the GSConv channel shuffle folded into addressing.
// Persistent CTAs over (image, spatial tile, 32-channel group) work items, two cp.async stages: the (TH+4) x (TW+4)
// halo of the NEXT item (zero-filled outside the map = the conv padding), its 25 x 32 weights and bias land in shared
// memory while the current item is computed.  A warp = 8 rows x one strip of 4 pixels; lane = (row, 8-channel vector);
// the halo row pitch is an odd number of pixels so that the 8 lanes of a quarter warp hit 8 different 16-byte bank
// groups.  Every staged vector feeds up to 5 x 4 taps from registers; the MACs are packed fp32 pairs (FFMA2).
// ------------------------------------------------------------------------------------------------------------------
struct Dw5Args {
    const __nv_bfloat16 *in;
    __nv_bfloat16 *out;
    const float *w, *bias;          // [25][C] (tap-major), [C]
    int in_cs, in_off0, in_off1, out_cs, out_off0, out_off1;
    int C, half, H, W, act;
    int SX, RG;                     // warps: SX strips of 4 pixels x RG groups of 8 rows
    int pitch;                      // halo row pitch in pixels (odd)
    int tiles_x, tiles_y, n_items, stage_bytes, halo_bytes;
};

__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float2 unpack_f32x2(uint64_t v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
// two bf16 of one 32-bit word -> exact fp32 pair
__device__ __forceinline__ uint64_t bf16x2_to_f32x2(uint32_t u) {
    return pack_f32x2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}

__device__ __forceinline__ void dw5_stage_load(const Dw5Args &a, uint32_t base, int item) {
    const int ncg = a.C >> 5;
    const int cg = item % ncg; item /= ncg;
    const int tx = item % a.tiles_x; item /= a.tiles_x;
    const int ty = item % a.tiles_y;
    const int b = item / a.tiles_y;
    const int twp4 = 4 * a.SX + 4, th4 = 8 * a.RG + 4;
    const int x0 = tx * 4 * a.SX - 2, y0 = ty * 8 * a.RG - 2, c0 = cg * 32;
    const size_t img_base = (size_t)b * a.H * a.W;
    for (int i = threadIdx.x; i < th4 * twp4 * 4; i += blockDim.x) {
        const int v = i & 3, q = i >> 2;
        const int py = q / twp4, px = q - py * twp4;
        const int yy = y0 + py, xx = x0 + px;
        const int c = c0 + v * 8;
        const int ci = c < a.half ? a.in_off0 + c : a.in_off1 + (c - a.half);
        const bool ok = yy >= 0 && yy < a.H && xx >= 0 && xx < a.W;
        const __nv_bfloat16 *src = a.in + (ok ? (img_base + (size_t)yy * a.W + xx) * a.in_cs + ci : 0);
        cp_async16(base + (uint32_t)(((py * a.pitch + px) * 4 + v) * 16), src, ok);
    }
    for (int i = threadIdx.x; i < 25 * 8 + 8; i += blockDim.x) {         // weights [25][32] fp32, then bias [32]
        const float *src = i < 200 ? a.w + (size_t)(i >> 3) * a.C + c0 + (i & 7) * 4 : a.bias + c0 + (i - 200) * 4;
        cp_async16(base + (uint32_t)a.halo_bytes + (uint32_t)i * 16, src, true);
    }
}

__global__ void __launch_bounds__(512, 1) dw5_kernel(const __grid_constant__ Dw5Args a) {
    pdl_trigger();
    extern __shared__ __align__(128) uint8_t dw_smem[];
    const uint32_t smem_u = (uint32_t)__cvta_generic_to_shared(dw_smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int v = lane & 3, ry = lane >> 2;
    const int sxi = warp % a.SX, rgi = warp / a.SX;
    const int r = rgi * 8 + ry, sx = sxi * 4;
    const int ncg = a.C >> 5;
    pdl_wait();
    int item = blockIdx.x, stage = 0;
    if (item < a.n_items) dw5_stage_load(a, smem_u, item);
    asm volatile("cp.async.commit_group;" ::: "memory");
    for (; item < a.n_items; item += gridDim.x, stage ^= 1) {
        if (item + (int)gridDim.x < a.n_items) dw5_stage_load(a, smem_u + (stage ^ 1) * a.stage_bytes, item + gridDim.x);
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncthreads();
        const uint8_t *base = dw_smem + (size_t)stage * a.stage_bytes;
        const float *sw = reinterpret_cast<const float *>(base + a.halo_bytes) + v * 8;
        uint64_t acc[4][4];
        {
            const float4 b0 = *reinterpret_cast<const float4 *>(sw + 25 * 32), b1 = *reinterpret_cast<const float4 *>(sw + 25 * 32 + 4);
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                acc[p][0] = pack_f32x2(b0.x, b0.y); acc[p][1] = pack_f32x2(b0.z, b0.w);
                acc[p][2] = pack_f32x2(b1.x, b1.y); acc[p][3] = pack_f32x2(b1.z, b1.w);
            }
        }
        const uint4 *row = reinterpret_cast<const uint4 *>(base) + ((size_t)r * a.pitch + sx) * 4 + v;
#pragma unroll 1
        for (int dy = 0; dy < 5; ++dy) {
            uint64_t wk[5][4];
#pragma unroll
            for (int dx = 0; dx < 5; ++dx) {
                const float4 w0 = *reinterpret_cast<const float4 *>(sw + (dy * 5 + dx) * 32);
                const float4 w1 = *reinterpret_cast<const float4 *>(sw + (dy * 5 + dx) * 32 + 4);
                wk[dx][0] = pack_f32x2(w0.x, w0.y); wk[dx][1] = pack_f32x2(w0.z, w0.w);
                wk[dx][2] = pack_f32x2(w1.x, w1.y); wk[dx][3] = pack_f32x2(w1.z, w1.w);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {                       // input pixel sx - 2 + j of this halo row
                const uint4 u = row[j * 4];
                const uint64_t x0 = bf16x2_to_f32x2(u.x), x1 = bf16x2_to_f32x2(u.y), x2 = bf16x2_to_f32x2(u.z), x3 = bf16x2_to_f32x2(u.w);
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const int dx = j - p;
                    if (dx >= 0 && dx < 5) {
                        acc[p][0] = ffma2(x0, wk[dx][0], acc[p][0]);
                        acc[p][1] = ffma2(x1, wk[dx][1], acc[p][1]);
                        acc[p][2] = ffma2(x2, wk[dx][2], acc[p][2]);
                        acc[p][3] = ffma2(x3, wk[dx][3], acc[p][3]);
                    }
                }
            }
            row += (size_t)a.pitch * 4;
        }
        {
            int it = item;
            const int cg = it % ncg; it /= ncg;
            const int tx = it % a.tiles_x; it /= a.tiles_x;
            const int ty = it % a.tiles_y;
            const int b = it / a.tiles_y;
            const int y = ty * 8 * a.RG + r, xb = tx * 4 * a.SX + sx;
            const int c = cg * 32 + v * 8;
            const int co = c < a.half ? a.out_off0 + c : a.out_off1 + (c - a.half);
            if (y < a.H) {
                __nv_bfloat16 *orow = a.out + (((size_t)b * a.H + y) * a.W + xb) * a.out_cs + co;
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    if (xb + p >= a.W) break;
                    uint32_t o[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        float2 f = unpack_f32x2(acc[p][k]);
                        if (a.act == 1) {                        // SiLU(x) = h + h * tanh(h), h = x / 2
                            const float h0 = 0.5f * f.x, h1 = 0.5f * f.y;
                            f.x = fmaf(h0, tanh_approx_f(h0), h0);
                            f.y = fmaf(h1, tanh_approx_f(h1), h1);
                        }
                        o[k] = pack_bf16x2(f.x, f.y);
                    }
                    stg16(orow + (size_t)p * a.out_cs, make_uint4(o[0], o[1], o[2], o[3]));
                }
            }
        }
        __syncthreads();                                          // this stage is refilled by the next iteration's prefetch
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Depthwise 5x5 on the tensor cores (mma.sync m16n8k16, bf16 x bf16 -> fp32).  A depthwise conv is a GEMM with a
// block-diagonal weight matrix: for one group of 8 channels and one pair of horizontal taps (dx, dx+1) of kernel row dy
//     D[pixel][c] += sum_{k = (tap j, channel c')} A[pixel][k] * B[k][c],   A[pixel][(j, c')] = in[y + dy][x + dx + j][c'],
//                                                                          B[(j, c')][c]     = w[dy][dx + j][c] * (c' == c)
// 7/8 of the MACs multiply zeros, which the tensor pipe has room for (a 5x5 depthwise conv is 0.2 % of the model's FLOPs)
// while the FMA pipe does not: the SIMT version above spends ~40 issue slots per output, this one ~0.4.
//   unit  = 8 pixels x 2 blocks of R = 5 output rows x 8 channels: MMA rows 0-7 = the pixels of block 0, rows 8-15 = block 1
//           (ldmatrix addresses make that free), one fp32 accumulator fragment per output row;
//   step  = one input row s of the (R + 4)-row halo: three ldmatrix.x4 give the A fragments of the tap pairs (0,1) (2,3) (4,-),
//           each feeds every output row i with 0 <= s - i <= 4 (weights of kernel row s - i): a staged vector is read from
//           shared memory 3 times instead of 25;
//   B fragments (5 kernel rows x 3 tap pairs) live in registers for the whole work item.
// Work items, persistence and the cp.async double buffering are those of the SIMT kernel; the halo is stored as four planes
// (one per 8-channel vector) so that the 8 rows of an ldmatrix matrix are 128 contiguous bytes.  Weights enter as bf16
// (like every other conv of the model); bias, activation and accumulation are fp32.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kDwR = 5;                         // output rows per half block
struct Dw5mArgs {
    const __nv_bfloat16 *in;
    __nv_bfloat16 *out;
    const float *w, *bias;          // [25][C] (tap-major), [C]
    int in_cs, in_off0, in_off1, out_cs, out_off0, out_off1;
    int C, half, H, W, act;
    int TW, strips, swarps;         // tile width (multiple of 8), strips of 8 pixels, warps per channel vector
    int pitch;                      // halo row pitch in pixels
    int tiles_x, tiles_y, n_items, stage_bytes, halo_bytes;
    const CUtensorMap *imap;        // input map {C, W, H, B}, box {8, TW + 4, 2R + 4, 1}: one TMA load per channel vector (OOB = zero padding)
};

__device__ __forceinline__ void dw5m_stage_load(const Dw5mArgs &a, uint32_t base, int item) {
    const int ncg = a.C >> 5;
    const int cg = item % ncg; item /= ncg;
    const int tx = item % a.tiles_x; item /= a.tiles_x;
    const int ty = item % a.tiles_y;
    const int b = item / a.tiles_y;
    const int tw4 = a.TW + 4, th4 = 2 * kDwR + 4;
    const int x0 = tx * a.TW - 2, y0 = ty * 2 * kDwR - 2, c0 = cg * 32;
    // chunk index = (row py, pixel px, vector v) with v fastest; blockDim and the chunks per row are multiples of 4, so a thread
    // keeps its v (and its channel offset) for the whole item and walks (py, px) without a division
    const int v = threadIdx.x & 3;
    const int c = c0 + v * 8;
    const int ci = c < a.half ? a.in_off0 + c : a.in_off1 + (c - a.half);
    const __nv_bfloat16 *src0 = a.in + (size_t)b * a.H * a.W * a.in_cs + ci;
    const uint32_t dst0 = base + (uint32_t)(v * th4 * a.pitch * 16);
    const int row_chunks = tw4 * 4;
    int py = 0, rem = threadIdx.x;
    while (rem >= row_chunks) { rem -= row_chunks; ++py; }
    while (py < th4) {
        const int px = rem >> 2;
        const int yy = y0 + py, xx = x0 + px;
        const bool ok = yy >= 0 && yy < a.H && xx >= 0 && xx < a.W;
        cp_async16(dst0 + (uint32_t)((py * a.pitch + px) * 16), src0 + (ok ? (yy * a.W + xx) * a.in_cs : 0), ok);
        rem += blockDim.x;
        while (rem >= row_chunks) { rem -= row_chunks; ++py; }
    }
    for (int i = threadIdx.x; i < 25 * 8 + 8; i += blockDim.x) {         // weights [25][32] fp32, then bias [32]
        const float *src = i < 200 ? a.w + (size_t)(i >> 3) * a.C + c0 + (i & 7) * 4 : a.bias + c0 + (i - 200) * 4;
        cp_async16(base + (uint32_t)a.halo_bytes + (uint32_t)i * 16, src, true);
    }
}

__global__ void __launch_bounds__(128, 5) dw5_mma_kernel(const __grid_constant__ Dw5mArgs a) {
    pdl_trigger();
    extern __shared__ __align__(128) uint8_t dw_smem[];
    const uint32_t smem_u = (uint32_t)__cvta_generic_to_shared(dw_smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const int v = warp & 3, sw = warp >> 2;                      // channel vector of this warp, first strip
    const int ncg = a.C >> 5;
    const int th4 = 2 * kDwR + 4;
    // ldmatrix row address of this lane: matrix mi = lane / 8 -> (tap j = mi / 2, half block = mi % 2), row = lane % 8 = pixel
    const int mi = lane >> 3, mr = lane & 7;
    const uint32_t lm_off = (uint32_t)((((v * th4) + (mi & 1) * kDwR) * a.pitch + mr) * 16);
    const int joff = mi >> 1;                                    // second tap of the pair; the pair (4, -) re-reads tap 4 (zero weights)
    pdl_wait();
    // one work item per CTA, one staging buffer: several CTAs per SM overlap each other's load and compute phases
    {
        const int item = blockIdx.x;
        uint64_t *bar = reinterpret_cast<uint64_t *>(dw_smem + a.stage_bytes);
        if (a.imap != nullptr) {
            // halo planes by TMA (four boxes of 8 channels; out-of-bounds pixels arrive as zeros = the conv padding)
            if (threadIdx.x == 0) {
                ptx::mbar_init(bar, 1);
                ptx::fence_mbar_init();
                int it = item;
                const int cg = it % ncg; it /= ncg;
                const int tx = it % a.tiles_x; it /= a.tiles_x;
                const int ty = it % a.tiles_y;
                const int b = it / a.tiles_y;
                const uint32_t plane = (uint32_t)(th4 * a.pitch * 16);
                ptx::mbar_expect_tx(bar, 4u * plane);
#pragma unroll
                for (int vv = 0; vv < 4; ++vv) {
                    const int c = cg * 32 + vv * 8;
                    const int ci = c < a.half ? a.in_off0 + c : a.in_off1 + (c - a.half);
                    ptx::tma_load_4d(dw_smem + (size_t)vv * plane, a.imap, bar, ci, tx * a.TW - 2, ty * 2 * kDwR - 2, b);
                }
            }
            {
                const int c0 = (item % ncg) * 32;
                for (int i = threadIdx.x; i < 25 * 8 + 8; i += blockDim.x) {         // weights [25][32] fp32, then bias [32]
                    const float *src = i < 200 ? a.w + (size_t)(i >> 3) * a.C + c0 + (i & 7) * 4 : a.bias + c0 + (i - 200) * 4;
                    cp_async16(smem_u + (uint32_t)a.halo_bytes + (uint32_t)i * 16, src, true);
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();                                      // barrier initialised + weights visible
            ptx::mbar_wait(bar, 0);
        } else {
            dw5m_stage_load(a, smem_u, item);
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();
        }
        const int stage = 0;
        const uint32_t base_u = smem_u + (uint32_t)stage * a.stage_bytes;
        const float *swt = reinterpret_cast<const float *>(dw_smem + (size_t)stage * a.stage_bytes + a.halo_bytes) + v * 8;
        // ---- B fragments: column n = g of the block-diagonal tap-pair matrices; k rows 2*t4, 2*t4+1 (tap j = 0) and +8 (tap j = 1) ----
        uint32_t bf[5][3][2];
        {
            const bool lo = g == 2 * t4, hi = g == 2 * t4 + 1;
#pragma unroll
            for (int dy = 0; dy < 5; ++dy)
#pragma unroll
                for (int p = 0; p < 3; ++p)
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int dx = 2 * p + j;
                        const float wv = dx < 5 ? swt[(dy * 5 + dx) * 32 + g] : 0.0f;
                        bf[dy][p][j] = pack_bf16x2(lo ? wv : 0.0f, hi ? wv : 0.0f);
                    }
        }
        const float bias0 = swt[25 * 32 + 2 * t4], bias1 = swt[25 * 32 + 2 * t4 + 1];
        int it = item;
        const int cg = it % ncg; it /= ncg;
        const int tx = it % a.tiles_x; it /= a.tiles_x;
        const int ty = it % a.tiles_y;
        const int b = it / a.tiles_y;
        for (int s8 = sw; s8 < a.strips; s8 += a.swarps) {       // strips of 8 pixels
            float acc[kDwR][4];
#pragma unroll
            for (int i = 0; i < kDwR; ++i) { acc[i][0] = acc[i][2] = bias0; acc[i][1] = acc[i][3] = bias1; }
            const uint32_t a_base = base_u + lm_off + (uint32_t)(s8 * 8 * 16);
#pragma unroll
            for (int s = 0; s < kDwR + 4; ++s) {                 // input rows of the halo
                uint32_t af[3][4];
#pragma unroll
                for (int p = 0; p < 3; ++p) {
                    const uint32_t addr = a_base + (uint32_t)((s * a.pitch + (p < 2 ? 2 * p + joff : 4)) * 16);
                    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                                 : "=r"(af[p][0]), "=r"(af[p][1]), "=r"(af[p][2]), "=r"(af[p][3]) : "r"(addr));
                }
#pragma unroll
                for (int i = 0; i < kDwR; ++i) {
                    const int dy = s - i;
                    if (dy < 0 || dy > 4) continue;
#pragma unroll
                    for (int p = 0; p < 3; ++p)
                        mma_bf16_16816(acc[i], af[p][0], af[p][1], af[p][2], af[p][3], bf[dy][p][0], bf[dy][p][1]);
                }
            }
            // ---- epilogue: activation, bf16, 4-byte stores (the 4 lanes of a pixel write 16 contiguous bytes) ----
            const int x = tx * a.TW + s8 * 8 + g;
            const int c = cg * 32 + v * 8 + 2 * t4;
            const int co = c < a.half ? a.out_off0 + c : a.out_off1 + (c - a.half);
            if (x < a.W) {
#pragma unroll
                for (int hb = 0; hb < 2; ++hb)
#pragma unroll
                    for (int i = 0; i < kDwR; ++i) {
                        const int y = ty * 2 * kDwR + hb * kDwR + i;
                        if (y >= a.H) continue;
                        float f0 = acc[i][2 * hb], f1 = acc[i][2 * hb + 1];
                        if (a.act == 1) {                        // SiLU(x) = h + h * tanh(h), h = x / 2
                            const float h0 = 0.5f * f0, h1 = 0.5f * f1;
                            f0 = fmaf(h0, tanh_approx_f(h0), h0);
                            f1 = fmaf(h1, tanh_approx_f(h1), h1);
                        }
                        *reinterpret_cast<uint32_t *>(a.out + (((size_t)b * a.H + y) * a.W + x) * a.out_cs + co) = pack_bf16x2(f0, f1);
                    }
            }
        }
    }
}

// MP (reference models/common.py:32-38): MaxPool2d(2, 2).  One thread = 8 channels of one output pixel.
__global__ void __launch_bounds__(256) maxpool2_kernel(const __nv_bfloat16 *__restrict__ in, int in_cs, int in_off,
                                                       __nv_bfloat16 *__restrict__ out, int out_cs, int out_off, int C,
                                                       int B, int H, int W) {
    pdl_trigger();
    pdl_wait();
    const int vecs = C / 8, Ho = H / 2, Wo = W / 2;
    const unsigned total = (unsigned)B * Ho * Wo * vecs;                 // 32-bit index arithmetic (maps of < 2^31 vectors)
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const unsigned opix = i / vecs, v = i - opix * vecs;
        const unsigned orow = opix / Wo, xo = opix - orow * Wo;          // orow = b * Ho + yo
        const unsigned b = orow / Ho, yo = orow - b * Ho;
        const __nv_bfloat16 *p = in + (((size_t)b * H + 2 * yo) * W + 2 * xo) * in_cs + in_off + v * 8;
        const uint4 a = ldg16(p), c = ldg16(p + in_cs), d = ldg16(p + (size_t)W * in_cs), e = ldg16(p + (size_t)(W + 1) * in_cs);
        stg16(out + (size_t)opix * out_cs + out_off + v * 8, bf16x8_max(bf16x8_max(a, c), bf16x8_max(d, e)));
    }
}

// SPPCSPC pools (reference models/common.py:279, 286): MaxPool2d(k, 1, k//2) for k = 5, 9, 13 (-inf padding).  Max is
// exact, so 9 = 5 o 5 and 13 = 5 o 5 o 5, and each 5x5 is a row pass followed by a column pass.  One CTA = one image x CH
// channels held in two shared-memory planes: one read of the map, three writes at channel offsets of the cv5 input.
__global__ void __launch_bounds__(256) spp_kernel(const __nv_bfloat16 *__restrict__ in, int in_cs, int in_off,
                                                  __nv_bfloat16 *__restrict__ out, int out_cs, int off5, int off9,
                                                  int off13, int C, int H, int W, int CH) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(16) uint8_t spp_smem[];
    const int vecs = CH / 8, HW = H * W, n = HW * vecs;
    uint4 *A = reinterpret_cast<uint4 *>(spp_smem), *Bf = A + n;
    const int chunks = C / CH;
    const int b = blockIdx.x / chunks, c0 = (blockIdx.x % chunks) * CH;
    const size_t img_base = (size_t)b * HW;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int v = i % vecs, pix = i / vecs;
        A[i] = ldg16(in + (img_base + pix) * in_cs + in_off + c0 + v * 8);
    }
    __syncthreads();
    const int offs[3] = {off5, off9, off13};
    for (int pass = 0; pass < 3; ++pass) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {            // row pass: A -> Bf
            const int v = i % vecs, pix = i / vecs, x = pix % W;
            uint4 m = A[i];
#pragma unroll
            for (int d = 1; d <= 2; ++d) {
                if (x - d >= 0) m = bf16x8_max(m, A[i - d * vecs]);
                if (x + d < W) m = bf16x8_max(m, A[i + d * vecs]);
            }
            (void)v;
            Bf[i] = m;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) {            // column pass: Bf -> A, and out
            const int v = i % vecs, pix = i / vecs, y = pix / W;
            uint4 m = Bf[i];
#pragma unroll
            for (int d = 1; d <= 2; ++d) {
                if (y - d >= 0) m = bf16x8_max(m, Bf[i - d * W * vecs]);
                if (y + d < H) m = bf16x8_max(m, Bf[i + d * W * vecs]);
            }
            A[i] = m;
            stg16(out + (img_base + pix) * out_cs + offs[pass] + c0 + v * 8, m);
        }
        __syncthreads();
    }
}

// nn.Upsample(None, 2, 'nearest') (reference cfg/training/Rep-YOLO.yaml:43,51).
__global__ void __launch_bounds__(256) upsample2_kernel(const __nv_bfloat16 *__restrict__ in, int in_cs, int in_off,
                                                        __nv_bfloat16 *__restrict__ out, int out_cs, int out_off, int C,
                                                        int B, int H, int W) {
    pdl_trigger();
    pdl_wait();
    // one thread = one INPUT vector (8 channels of one pixel) -> its four output pixels; 32-bit index arithmetic
    const int vecs = C / 8, Wo = W * 2;
    const unsigned total = (unsigned)B * H * W * vecs;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const unsigned ipix = i / vecs, v = i - ipix * vecs;
        const unsigned row = ipix / W, x = ipix - row * W;        // row = b * H + y
        const uint4 u = ldg16(in + (size_t)ipix * in_cs + in_off + v * 8);
        __nv_bfloat16 *o = out + ((size_t)row * 2 * Wo + 2 * x) * out_cs + out_off + v * 8;
        stg16(o, u);
        stg16(o + out_cs, u);
        stg16(o + (size_t)Wo * out_cs, u);
        stg16(o + (size_t)(Wo + 1) * out_cs, u);
    }
}

// CA (reference models/common.py:3797-3802): p = avgpool(x); out = p * sigmoid(f2(relu(f1(p)))) + p  -> [B, C] fp32.
// Stage 1: kCaSplits CTAs per image each sum a pixel range in fp32 -> partial[B][S][C] (fixed order: deterministic).
// Stage 2: one CTA per image adds the partials in order, then the two tiny FCs.
constexpr int kCaSplits = 16;

__global__ void __launch_bounds__(256) ca_partial_kernel(const __nv_bfloat16 *__restrict__ in, int in_cs, int in_off,
                                                         float *__restrict__ partial, int C, int HW) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float ca_sm[];         // [256 * 8] partials
    const int b = blockIdx.x / kCaSplits, sp = blockIdx.x % kCaSplits, vecs = C / 8;
    const int groups = blockDim.x / vecs;    // pixel groups working in parallel (C <= 2048)
    const int v = threadIdx.x % vecs, g = threadIdx.x / vecs;
    const int p0 = (int)((long)HW * sp / kCaSplits), p1 = (int)((long)HW * (sp + 1) / kCaSplits);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (g < groups) {
        const __nv_bfloat16 *base = in + (size_t)b * HW * in_cs + in_off + v * 8;
        for (int p = p0 + g; p < p1; p += groups) {
            const uint4 u = ldg16(base + (size_t)p * in_cs);
            const float2 f0 = unpack_bf16x2(u.x), f1v = unpack_bf16x2(u.y), f2v = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
            acc[0] += f0.x; acc[1] += f0.y; acc[2] += f1v.x; acc[3] += f1v.y;
            acc[4] += f2v.x; acc[5] += f2v.y; acc[6] += f3.x; acc[7] += f3.y;
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) ca_sm[threadIdx.x * 8 + k] = acc[k];
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const int vv = c / 8, k = c % 8;
        float s = 0.0f;
        for (int gg = 0; gg < groups; ++gg) s += ca_sm[(gg * vecs + vv) * 8 + k];
        partial[((size_t)b * kCaSplits + sp) * C + c] = s;
    }
}

__global__ void __launch_bounds__(256) ca_finish_kernel(const float *__restrict__ partial, float *__restrict__ out, int out_cs,
                                                        int out_off, const float *__restrict__ f1, const float *__restrict__ f2,
                                                        int C, int HW) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float ca_sm[];         // [C] mean, [C/16] hidden
    float *mean = ca_sm, *hid = ca_sm + C;
    const int b = blockIdx.x;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.0f;
        for (int sp = 0; sp < kCaSplits; ++sp) s += partial[((size_t)b * kCaSplits + sp) * C + c];
        mean[c] = s / (float)HW;
    }
    __syncthreads();
    const int Ch = C / 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int j = warp; j < Ch; j += nw) {
        float s = 0.0f;
        for (int c = lane; c < C; c += 32) s = fmaf(__ldg(f1 + (size_t)j * C + c), mean[c], s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) hid[j] = fmaxf(s, 0.0f);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.0f;
        if ((Ch & 3) == 0) {                                  // the row of f2 as independent 16-byte loads (the loop is L2-latency bound)
            const float4 *row = reinterpret_cast<const float4 *>(f2 + (size_t)c * Ch);
            for (int j0 = 0; j0 < Ch; j0 += 32) {
                float4 w[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) w[q] = j0 + 4 * q < Ch ? __ldg(row + (j0 >> 2) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int j = j0 + 4 * q;
                    if (j < Ch) {                             // same summation order as the scalar loop
                        s = fmaf(w[q].x, hid[j], s); s = fmaf(w[q].y, hid[j + 1], s);
                        s = fmaf(w[q].z, hid[j + 2], s); s = fmaf(w[q].w, hid[j + 3], s);
                    }
                }
            }
        } else {
            for (int j = 0; j < Ch; ++j) s = fmaf(__ldg(f2 + (size_t)c * Ch + j), hid[j], s);
        }
        const float a = 1.0f / (1.0f + __expf(-s));
        out[(size_t)b * out_cs + out_off + c] = mean[c] * a + mean[c];
    }
}

inline int grid_for(size_t total, int block) {
    size_t g = (total + block - 1) / block;
    const size_t cap = (size_t)num_sms() * 16;
    return (int)(g < cap ? (g ? g : 1) : cap);
}

}  // namespace

template <int COUT, bool U8>
void stem_launch_t(const void *img, const float *w27, const float *bias, __nv_bfloat16 *out, int out_cs, int out_off, int B, int H,
                   int W, cudaStream_t st) {
    const long tiles = (long)B * cdiv(H / 2, kStemTH) * cdiv(W / 2, kStemTW);
    const size_t smem = (size_t)2 * kStemPatchFloats * (U8 ? 1 : 4) + (size_t)2 * kStemTW * COUT * 2;
    RY_ENSURE_DYN_SMEM((stem_kernel<COUT, U8>), 100 * 1024);
    const int grid = (int)std::min<long>(tiles, (long)num_sms() * 2);
    launch_pdl(stem_kernel<COUT, U8>, dim3(grid), dim3(256), smem, st, img, w27, bias, out, B, H, W, out_cs, out_off);
}

int stem_launch(const void *img, int img_u8, const float *w27, const float *bias, __nv_bfloat16 *out, int out_cs, int out_off,
                int cout, int B, int H, int W, cudaStream_t st) {
#define RY_STEM_CASE(C)                                                                                   \
    case C:                                                                                               \
        if (img_u8) stem_launch_t<C, true>(img, w27, bias, out, out_cs, out_off, B, H, W, st);            \
        else stem_launch_t<C, false>(img, w27, bias, out, out_cs, out_off, B, H, W, st);                  \
        break;
    switch (cout) {
        RY_STEM_CASE(16) RY_STEM_CASE(32) RY_STEM_CASE(48) RY_STEM_CASE(64)
        default: return 1;
    }
#undef RY_STEM_CASE
    return 0;
}

int dw5_tile_w(int W) {
    const int w8 = (W + 7) / 8 * 8;
    int tw = std::min(w8, 40);
    if (w8 > 40 && w8 % 40 != 0 && w8 % 32 == 0) tw = 32;
    return tw;
}
int dw5_halo_rows() { return 2 * kDwR + 4; }

void dw5_launch(const __nv_bfloat16 *in, int in_cs, int in_off0, int in_off1, __nv_bfloat16 *out, int out_cs, int out_off0,
                int out_off1, const float *w, const float *bias, int C, int half, int B, int H, int W, int act,
                const CUtensorMap *imap, cudaStream_t st) {
    static const bool use_ffma = getenv("RY_DW5_FFMA") != nullptr;
    if (!use_ffma && half % 8 == 0) {
        Dw5mArgs m = {};
        m.in = in; m.out = out; m.w = w; m.bias = bias;
        m.in_cs = in_cs; m.in_off0 = in_off0; m.in_off1 = in_off1; m.out_cs = out_cs; m.out_off0 = out_off0; m.out_off1 = out_off1;
        m.C = C; m.half = half; m.H = H; m.W = W; m.act = act;
        m.TW = dw5_tile_w(W);
        m.imap = imap;
        m.strips = m.TW / 8;
        m.swarps = 1;                                            // 4 warps (one per channel vector), every warp walks all strips
        m.pitch = m.TW + 4;
        m.tiles_x = cdiv(W, m.TW); m.tiles_y = cdiv(H, 2 * kDwR);
        m.halo_bytes = 4 * (2 * kDwR + 4) * m.pitch * 16;
        m.stage_bytes = (m.halo_bytes + 208 * 16 + 127) & ~127;
        m.n_items = B * m.tiles_y * m.tiles_x * (C / 32);
        const int threads = 128 * m.swarps;
        RY_ENSURE_DYN_SMEM(dw5_mma_kernel, 200 * 1024);
        launch_pdl(dw5_mma_kernel, dim3(m.n_items), dim3(threads), (size_t)m.stage_bytes + 16, st, m);   // + the TMA barrier
        return;
    }
    // tile = (4 SX) x (8 RG) pixels, SX * RG warps.  Model: an SM issues for ~16 warps at a time, so one round of n resident
    // CTAs costs n * warps; rounds = ceil(items / (148 n)); staged halo pixels per output pixel add L2 -> smem traffic.
    Dw5Args a = {};
    const int n_sm = num_sms();
    double best = -1.0;
    for (int sx = 1; sx <= 16; ++sx)
        for (int rg = 1; rg * sx <= 16; ++rg) {
            const int warps = sx * rg;
            if (warps < 4 && !(4 * sx >= W && 8 * rg >= H)) continue;
            const int pitch = 4 * sx + 5;
            const size_t stage = ((size_t)(8 * rg + 4) * pitch * 64 + 208 * 16 + 127) & ~(size_t)127;
            if (2 * stage > 200 * 1024) continue;
            const int n_res = std::max(1, std::min((int)(220 * 1024 / (2 * stage)), 16 / warps));
            const long items = (long)B * cdiv(W, 4 * sx) * cdiv(H, 8 * rg) * (C / 32);
            const double rounds = (double)((items + (long)n_sm * n_res - 1) / ((long)n_sm * n_res));
            const double halo = (double)(8 * rg + 4) * (4 * sx + 4) / (32.0 * warps);
            const double cost = rounds * n_res * warps * (1.0 + 0.08 * halo);
            if (best < 0 || cost < best) {
                best = cost;
                a.SX = sx; a.RG = rg; a.pitch = pitch; a.stage_bytes = (int)stage;
                a.halo_bytes = (8 * rg + 4) * pitch * 64;
                a.tiles_x = cdiv(W, 4 * sx); a.tiles_y = cdiv(H, 8 * rg);
                a.n_items = (int)items;
            }
        }
    a.in = in; a.out = out; a.w = w; a.bias = bias;
    a.in_cs = in_cs; a.in_off0 = in_off0; a.in_off1 = in_off1; a.out_cs = out_cs; a.out_off0 = out_off0; a.out_off1 = out_off1;
    a.C = C; a.half = half; a.H = H; a.W = W; a.act = act;
    const int warps = a.SX * a.RG;
    const int n_res = std::max(1, std::min((int)(220 * 1024 / (2 * (size_t)a.stage_bytes)), 16 / warps));
    RY_ENSURE_DYN_SMEM(dw5_kernel, 200 * 1024);
    launch_pdl(dw5_kernel, dim3(std::min(a.n_items, n_sm * n_res)), dim3(32 * warps), 2 * (size_t)a.stage_bytes, st, a);
}

void maxpool2_launch(const __nv_bfloat16 *in, int in_cs, int in_off, __nv_bfloat16 *out, int out_cs, int out_off, int C,
                     int B, int H, int W, cudaStream_t st) {
    const size_t total = (size_t)B * (H / 2) * (W / 2) * (C / 8);
    launch_pdl(maxpool2_kernel, dim3(grid_for(total, 256)), dim3(256), 0, st, in, in_cs, in_off, out, out_cs, out_off, C, B, H, W);
}

void spp_launch(const __nv_bfloat16 *in, int in_cs, int in_off, __nv_bfloat16 *out, int out_cs, int off5, int off9,
                int off13, int C, int B, int H, int W, cudaStream_t st) {
    static const int ch0 = getenv("RY_SPP_CH") ? atoi(getenv("RY_SPP_CH")) : 32;   // channels per CTA: 32 -> 51 KB of planes at 20x20, 4 CTAs per SM
    int CH = ch0;
    while (CH > 8 && (C % CH != 0 || (size_t)2 * H * W * CH * 2 > 200 * 1024)) CH -= 8;
    const size_t smem = (size_t)2 * H * W * CH * 2;
    RY_ENSURE_DYN_SMEM(spp_kernel, 227 * 1024);
    launch_pdl(spp_kernel, dim3(B * (C / CH)), dim3(256), smem, st, in, in_cs, in_off, out, out_cs, off5, off9, off13, C, H, W, CH);
}

void upsample2_launch(const __nv_bfloat16 *in, int in_cs, int in_off, __nv_bfloat16 *out, int out_cs, int out_off, int C,
                      int B, int H, int W, cudaStream_t st) {
    const size_t total = (size_t)B * H * W * (C / 8);
    launch_pdl(upsample2_kernel, dim3(grid_for(total, 256)), dim3(256), 0, st, in, in_cs, in_off, out, out_cs, out_off, C, B, H, W);
}

size_t ca_scratch_bytes(int B, int C) { return (size_t)B * kCaSplits * C * sizeof(float); }

void ca_launch(const __nv_bfloat16 *in, int in_cs, int in_off, float *out, int out_cs, int out_off, const float *f1,
               const float *f2, int C, int B, int HW, float *scratch, cudaStream_t st) {
    launch_pdl(ca_partial_kernel, dim3(B * kCaSplits), dim3(256), 256 * 8 * sizeof(float), st, in, in_cs, in_off, scratch, C, HW);
    launch_pdl(ca_finish_kernel, dim3(B), dim3(256), (size_t)(C + C / 16) * sizeof(float), st, scratch, out, out_cs, out_off, f1, f2, C, HW);
}

}  // namespace ry
