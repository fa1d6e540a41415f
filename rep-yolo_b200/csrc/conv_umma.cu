// tcgen05 / TMEM / TMA implicit-GEMM convolution kernel (see conv_umma.cuh).
#include "conv_umma.cuh"

#include "common.cuh"
#include "ptx.cuh"

namespace ry {

namespace {

constexpr int kStageA = 128 * 128;   // 128 rows x 128 B (64 bf16 of K) per stage, whatever the swizzle span

struct TileCoord {
    int w0, h0, n0, nc0;
};

__device__ __forceinline__ TileCoord tile_coord(const ConvArgs &p, int t) {
    const int nt = t % p.n_ntiles;
    int m = t / p.n_ntiles;
    const int wi = m % p.tiles_w;
    m /= p.tiles_w;
    const int hi = m % p.tiles_h;
    const int ni = m / p.tiles_h;
    return {wi * p.tw, hi * p.th, ni * p.tn, nt * p.BN};
}

__device__ __forceinline__ void store_epilogue(const ConvArgs &p, const float *v, int ng, size_t pix, int img) {
    // v[16]: fp32 accumulators of output channels ng..ng+15 of one pixel
    float x[16];
    const float4 *b4 = reinterpret_cast<const float4 *>(p.bias + ng);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 b = __ldg(b4 + q);
        x[4 * q + 0] = v[4 * q + 0] + b.x;
        x[4 * q + 1] = v[4 * q + 1] + b.y;
        x[4 * q + 2] = v[4 * q + 2] + b.z;
        x[4 * q + 3] = v[4 * q + 3] + b.w;
    }
    if (p.act == 1) {
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = silu_f(x[i]);
    }
    const bool second = ng + 8 < p.cout;
    if (p.res != nullptr) {
        const uint4 *r = reinterpret_cast<const uint4 *>(p.res + pix * p.res_cs + p.res_off + ng);
        uint4 r0 = __ldg(r), r1 = second ? __ldg(r + 1) : make_uint4(0, 0, 0, 0);
        const uint32_t rw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float2 f = unpack_bf16x2(rw[i]);
            x[2 * i] += f.x;
            x[2 * i + 1] += f.y;
        }
    }
    if (p.bvec != nullptr) {
        const float *bv = p.bvec + (size_t)img * p.bvec_cs + p.bvec_off + ng;
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] += (ng + i < p.cout) ? __ldg(bv + i) : 0.0f;
    }
    const int ch = ng < p.split_at ? p.off0 + ng : p.off1 + (ng - p.split_at);
    uint4 *dst = reinterpret_cast<uint4 *>(p.out + pix * p.out_cs + ch);
    dst[0] = make_uint4(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]), pack_bf16x2(x[4], x[5]), pack_bf16x2(x[6], x[7]));
    if (second)
        dst[1] = make_uint4(pack_bf16x2(x[8], x[9]), pack_bf16x2(x[10], x[11]), pack_bf16x2(x[12], x[13]),
                            pack_bf16x2(x[14], x[15]));
}

// IDetect.fuseforward decode (reference models/yolo.py:139-156): same op order, fp32.
__device__ __forceinline__ void detect_epilogue(const ConvArgs &p, const float *v, int j0, size_t pix) {
    const int b = (int)(pix / p.img_hw);
    const int rem = (int)(pix - (size_t)b * p.img_hw);
    const int gy = rem / p.img_w, gx = rem - gy * p.img_w;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int j = j0 + i;
        if (j >= p.na * p.no) break;
        const int a = j / p.no, o = j - a * p.no;
        const float t = v[i] + __ldg(p.bias + j);
        if (p.raw != nullptr) p.raw[(((size_t)b * p.na + a) * p.img_hw + rem) * p.no + o] = t;
        const float s = 1.0f / (1.0f + expf(-t));
        float y;
        if (o == 0) y = (s * 2.0f - 0.5f + (float)gx) * p.det_stride;
        else if (o == 1) y = (s * 2.0f - 0.5f + (float)gy) * p.det_stride;
        else if (o == 2) { const float u = s * 2.0f; y = u * u * p.anchors[2 * a]; }
        else if (o == 3) { const float u = s * 2.0f; y = u * u * p.anchors[2 * a + 1]; }
        else y = s;
        p.pred[((size_t)b * p.rows_total + p.row_off + (size_t)a * p.img_hw + rem) * p.no + o] = y;
    }
}

__global__ void __launch_bounds__(kConvThreads, 1) conv_umma_kernel(const __grid_constant__ ConvArgs p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int stage_bytes = kStageA + p.BN * 128;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)p.stages * stage_bytes);
    uint64_t *empty = full + p.stages;
    uint64_t *tfull = empty + p.stages;
    uint64_t *tempty = tfull + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = 64 / p.kb;                       // K blocks per pipeline stage
    const int row_bytes = p.kb * 2;
    const int a_sub = 128 * row_bytes, b_sub = p.BN * row_bytes;
    const int n_iters = (p.kblocks + G - 1) / G;
    const int total_tiles = p.tiles_w * p.tiles_h * p.tiles_n * p.n_ntiles;
    uint32_t tmem_cols = 32;
    while (tmem_cols < 2u * p.BN) tmem_cols <<= 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            ptx::mbar_init(full + s, 1);
            ptx::mbar_init(empty + s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            ptx::mbar_init(tfull + a, 1);
            ptx::mbar_init(tempty + a, 4);
        }
        ptx::fence_mbar_init();
        for (int i = 0; i < p.ntaps && i < 4; ++i) ptx::prefetch_tmap(p.amap + i);
        ptx::prefetch_tmap(p.wmap);
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, tmem_cols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            const uint32_t box_rows = p.tw * p.th * p.tn;
            int s = 0;
            uint32_t ph = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const TileCoord tc = tile_coord(p, t);
                for (int it = 0; it < n_iters; ++it) {
                    const int kb0 = it * G;
                    const int nsub = min(G, p.kblocks - kb0);
                    ptx::mbar_wait(empty + s, ph ^ 1);
                    ptx::mbar_expect_tx(full + s, nsub * (box_rows + p.BN) * row_bytes);
                    uint8_t *sa = smem + (size_t)s * stage_bytes, *sb = sa + kStageA;
                    for (int j = 0; j < nsub; ++j) {
                        const int kbi = kb0 + j;
                        const int tap = kbi / p.cblk, cb = kbi - tap * p.cblk;
                        ptx::tma_load_4d(sa + j * a_sub, p.amap + p.tap_map[tap], full + s, cb * p.kb,
                                         tc.w0 + p.tap_dw[tap], tc.h0 + p.tap_dh[tap], tc.n0);
                        ptx::tma_load_2d(sb + j * b_sub, p.wmap, full + s, kbi * p.kb, tc.nc0);
                    }
                    if (++s == p.stages) { s = 0; ph ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = ptx::umma_idesc_bf16(128, p.BN);
            int s = 0, acc = 0;
            uint32_t ph = 0, aph = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                ptx::mbar_wait(tempty + acc, aph ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * p.BN;
                uint32_t accumulate = 0;
                for (int it = 0; it < n_iters; ++it) {
                    const int nsub = min(G, p.kblocks - it * G);
                    ptx::mbar_wait(full + s, ph);
                    ptx::tc_fence_after();
                    const uint32_t sa = ptx::smem_u32(smem + (size_t)s * stage_bytes), sb = sa + kStageA;
                    for (int j = 0; j < nsub; ++j) {
                        for (int k = 0; k < p.kb / 16; ++k) {
                            const uint64_t da = ptx::umma_smem_desc(sa + j * a_sub + k * 32, row_bytes);
                            const uint64_t db = ptx::umma_smem_desc(sb + j * b_sub + k * 32, row_bytes);
                            ptx::umma_bf16(d_tmem, da, db, idesc, accumulate);
                            accumulate = 1;
                        }
                    }
                    ptx::umma_commit(empty + s);      // frees the smem stage once these MMAs retire
                    if (++s == p.stages) { s = 0; ph ^= 1; }
                }
                ptx::umma_commit(tfull + acc);        // accumulator ready for the epilogue
                if (++acc == 2) { acc = 0; aph ^= 1; }
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue (4 warps, TMEM lane quarter = warp % 4) =====================
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const int w_in = row % p.tw, h_in = (row / p.tw) % p.th, n_in = row / (p.tw * p.th);
        int acc = 0;
        uint32_t aph = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            const TileCoord tc = tile_coord(p, t);
            const int w = tc.w0 + w_in, h = tc.h0 + h_in, n = tc.n0 + n_in;
            const bool valid = (n_in < p.tn) && (w < p.Wo) && (h < p.Ho) && (n < p.Bo);
            const size_t pix = ((size_t)n * p.Ho + h) * p.Wo + w;
            const int img = (int)(pix / p.img_hw);
            ptx::mbar_wait(tfull + acc, aph);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + acc * p.BN + ((uint32_t)(quarter * 32) << 16);
            for (int c = 0; c < p.BN / 16; ++c) {
                float v[16];
                ptx::tmem_ld16(taddr + c * 16, v);
                const int ng = tc.nc0 + c * 16;
                if (valid && ng < p.cout) {
                    if (p.mode == 0) store_epilogue(p, v, ng, pix, img);
                    else detect_epilogue(p, v, ng, pix);
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(tempty + acc);
            if (++acc == 2) { acc = 0; aph ^= 1; }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, tmem_cols);
    }
}

}  // namespace

int conv_pick_stages(int BN) {
    const int stage_bytes = kStageA + BN * 128;
    int s = (200 * 1024) / stage_bytes;
    return s > 8 ? 8 : (s < 2 ? 2 : s);
}

size_t conv_smem_bytes(int BN, int stages) {
    return (size_t)stages * (kStageA + BN * 128) + (2 * stages + 4) * sizeof(uint64_t) + 16 + 1024;
}

void conv_launch(const ConvArgs &a, int grid, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(conv_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        attr_set = true;
    }
    conv_umma_kernel<<<grid, kConvThreads, conv_smem_bytes(a.BN, a.stages), stream>>>(a);
}

}  // namespace ry
