// Pre-/post-processing either side of the hot path, bit-exact with the reference's host code:
//   ry_letterbox_u8   utils/datasets.py:984-1014 letterbox (cv2.resize INTER_LINEAR + cv2.copyMakeBorder) fused with the
//                     BGR->RGB / HWC->CHW packing of LoadImages.__next__ (datasets.py:191-195)
//   ry_scale_coords   utils/general.py:319-340 scale_coords + clip_coords (+ the .round() of detect.py:114)
// cv2's 8-bit linear resize is fixed point (OpenCV modules/imgproc/src/resize.cpp, HResizeLinear / VResizeLinear <uchar, int, short>):
//   fx = float((dx + 0.5) * scale_x - 0.5); sx = floor(fx); fx -= sx; border columns clamp sx and zero fx; rows clamp the index
//   a = (rint((1 - fx) * 2048), rint(fx * 2048));  H = S[sx] * a0 + S[sx + 1] * a1;
//   dst = (((b0 * (H0 >> 4)) >> 16) + ((b1 * (H1 >> 4)) >> 16) + 2) >> 2
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <stdint.h>

#include "common.cuh"
#include "../../include/repyolo_b200.h"

namespace ry {
namespace {

struct LbArgs {
    const uint8_t *src;       // HWC, 3 channels, row pitch src_row_bytes
    uint8_t *dst;
    int H0, W0, src_row_bytes;
    int H1, W1;               // letterboxed extent
    int new_w, new_h, left, top;
    int planar_swap;          // 1: dst = [3][H1][W1] with channel order reversed (BGR -> RGB); 0: dst = HWC, same channel order
    int pad[3];               // border value per SOURCE channel
    double scale_x, scale_y;  // 1 / (new / old), computed in double on the host exactly as resize.cpp does
    int resize;               // 0: the reference skips cv2.resize when the shape already matches
};

__device__ __forceinline__ void lin_coeff(int d, double scale, int n_src, bool clamp_f, int &s, int &c0, int &c1) {
    float f = __double2float_rn(__dsub_rn(__dmul_rn(__dadd_rn((double)d, 0.5), scale), 0.5));
    s = (int)floorf(f);
    f = __fsub_rn(f, (float)s);
    if (clamp_f) {                                            // columns: clamp the index AND drop the fraction
        if (s < 0) { f = 0.0f; s = 0; }
        if (s >= n_src - 1) { f = 0.0f; s = n_src - 1; }
    }
    c0 = __float2int_rn(__fmul_rn(__fsub_rn(1.0f, f), 2048.0f));
    c1 = __float2int_rn(__fmul_rn(f, 2048.0f));
}

__global__ void __launch_bounds__(256) letterbox_kernel(const LbArgs a) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= a.W1) return;
    int v[3] = {a.pad[0], a.pad[1], a.pad[2]};
    const int dx = x - a.left, dy = y - a.top;
    if (dx >= 0 && dx < a.new_w && dy >= 0 && dy < a.new_h) {
        if (!a.resize) {
            const uint8_t *p = a.src + (size_t)dy * a.src_row_bytes + dx * 3;
            v[0] = p[0]; v[1] = p[1]; v[2] = p[2];
        } else {
            int sx, a0, a1, sy, b0, b1;
            lin_coeff(dx, a.scale_x, a.W0, true, sx, a0, a1);
            lin_coeff(dy, a.scale_y, a.H0, false, sy, b0, b1);
            const int sx1 = min(sx + 1, a.W0 - 1);            // a1 == 0 wherever sx + 1 is outside
            const int y0 = min(max(sy, 0), a.H0 - 1), y1 = min(max(sy + 1, 0), a.H0 - 1);
            const uint8_t *r0 = a.src + (size_t)y0 * a.src_row_bytes, *r1 = a.src + (size_t)y1 * a.src_row_bytes;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int h0 = (int)r0[sx * 3 + c] * a0 + (int)r0[sx1 * 3 + c] * a1;
                const int h1 = (int)r1[sx * 3 + c] * a0 + (int)r1[sx1 * 3 + c] * a1;
                v[c] = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
            }
        }
    }
    if (a.planar_swap) {
        const size_t plane = (size_t)a.H1 * a.W1, o = (size_t)y * a.W1 + x;
        a.dst[o] = (uint8_t)v[2];
        a.dst[plane + o] = (uint8_t)v[1];
        a.dst[2 * plane + o] = (uint8_t)v[0];
    } else {
        uint8_t *p = a.dst + ((size_t)y * a.W1 + x) * 3;
        p[0] = (uint8_t)v[0]; p[1] = (uint8_t)v[1]; p[2] = (uint8_t)v[2];
    }
}

// one thread per box; the arithmetic is the reference's torch fp32 chain: (x - pad) / gain, clamp, optional round-half-even
__global__ void __launch_bounds__(256) scale_coords_kernel(float *coords, const int *count, int n_max, int row_stride, float pad_x,
                                                           float pad_y, float gain, float w0, float h0, int round_result) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = count ? min(*count, n_max) : n_max;
    if (i >= n) return;
    float *c = coords + (size_t)i * row_stride;
    float x1 = __fdiv_rn(__fsub_rn(c[0], pad_x), gain), y1 = __fdiv_rn(__fsub_rn(c[1], pad_y), gain);
    float x2 = __fdiv_rn(__fsub_rn(c[2], pad_x), gain), y2 = __fdiv_rn(__fsub_rn(c[3], pad_y), gain);
    x1 = fminf(fmaxf(x1, 0.0f), w0); y1 = fminf(fmaxf(y1, 0.0f), h0);
    x2 = fminf(fmaxf(x2, 0.0f), w0); y2 = fminf(fmaxf(y2, 0.0f), h0);
    if (round_result) { x1 = rintf(x1); y1 = rintf(y1); x2 = rintf(x2); y2 = rintf(y2); }
    c[0] = x1; c[1] = y1; c[2] = x2; c[3] = y2;
}

// fp32 NCHW feature map -> a channel range of an NHWC bf16 tensor (IDetect.fuseforward called on its own, models/yolo.py:135: the
// caller hands over NCHW maps; the head kernels read NHWC bf16).  Lanes walk pixels (coalesced reads of each channel plane),
// every thread writes the 8 channels of its pixel as one 16-byte vector.
__global__ void __launch_bounds__(256) nchw_to_nhwc_bf16_kernel(const float *__restrict__ src, int C, size_t hw, size_t npix,
                                                                 __nv_bfloat16 *__restrict__ dst, int dst_cs, int dst_off) {
    const int groups = C >> 3;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix * groups; i += (size_t)gridDim.x * blockDim.x) {
        const size_t pix = i % npix;
        const int g = (int)(i / npix);
        const size_t b = pix / hw, r = pix - b * hw;
        const float *p = src + (b * C + (size_t)g * 8) * hw + r;
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) w[j] = pack_bf16x2(__ldg(p + (size_t)(2 * j) * hw), __ldg(p + (size_t)(2 * j + 1) * hw));
        *reinterpret_cast<uint4 *>(dst + pix * dst_cs + dst_off + g * 8) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

}  // namespace
}  // namespace ry

extern "C" {

int ry_letterbox_u8(const uint8_t *src_hwc, int H0, int W0, int src_row_bytes, uint8_t *dst, int H1, int W1, int new_w, int new_h,
                    int left, int top, const int32_t *pad_value3_host, int planar_rgb, void *stream) {
    using namespace ry;
    if (!src_hwc || !dst || !pad_value3_host) RY_FAIL("ry_letterbox_u8: NULL pointer");
    if (H0 <= 0 || W0 <= 0 || H1 <= 0 || W1 <= 0 || new_w <= 0 || new_h <= 0 || src_row_bytes < 3 * W0) RY_FAIL("ry_letterbox_u8: bad shape");
    if (left < 0 || top < 0 || left + new_w > W1 || top + new_h > H1) RY_FAIL("ry_letterbox_u8: the resized image does not fit the output");
    LbArgs a;
    a.src = src_hwc; a.dst = dst; a.H0 = H0; a.W0 = W0; a.src_row_bytes = src_row_bytes; a.H1 = H1; a.W1 = W1;
    a.new_w = new_w; a.new_h = new_h; a.left = left; a.top = top; a.planar_swap = planar_rgb ? 1 : 0;
    for (int i = 0; i < 3; ++i) a.pad[i] = pad_value3_host[i] & 255;
    a.scale_x = 1.0 / ((double)new_w / (double)W0);
    a.scale_y = 1.0 / ((double)new_h / (double)H0);
    a.resize = (new_w != W0 || new_h != H0) ? 1 : 0;
    letterbox_kernel<<<dim3((W1 + 255) / 256, H1), 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
    RY_CUDA(cudaGetLastError());
    return 0;
}

int ry_scale_coords(float *coords, const int32_t *count_dev, int n_max, int row_stride, float pad_x, float pad_y, float gain, int w0,
                    int h0, int round_result, void *stream) {
    using namespace ry;
    if (!coords) RY_FAIL("ry_scale_coords: NULL pointer");
    if (n_max < 0 || row_stride < 4) RY_FAIL("ry_scale_coords: bad shape");
    if (n_max == 0) return 0;
    scale_coords_kernel<<<(n_max + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(coords, count_dev, n_max, row_stride, pad_x, pad_y,
                                                                                       gain, (float)w0, (float)h0, round_result);
    RY_CUDA(cudaGetLastError());
    return 0;
}

int ry_nchw_to_nhwc_bf16(const float *src_nchw, int B, int C, int H, int W, void *dst_nhwc_bf16, int dst_channels, int dst_c_off,
                         void *stream) {
    using namespace ry;
    if (!src_nchw || !dst_nhwc_bf16) RY_FAIL("ry_nchw_to_nhwc_bf16: NULL pointer");
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || (C & 7) || (dst_c_off & 7) || (dst_channels & 7) || dst_c_off + C > dst_channels)
        RY_FAIL("ry_nchw_to_nhwc_bf16: channel counts / offsets must be multiples of 8 and fit the destination");
    const size_t hw = (size_t)H * W, npix = (size_t)B * hw, total = npix * (size_t)(C >> 3);
    const int grid = (int)std::min<size_t>((total + 255) / 256, (size_t)num_sms() * 16);
    nchw_to_nhwc_bf16_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(src_nchw, C, hw, npix, static_cast<__nv_bfloat16 *>(dst_nhwc_bf16),
                                                                                 dst_channels, dst_c_off);
    RY_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
