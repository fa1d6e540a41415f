// Probe (run on a B200): can a tcgen05.mma A-operand descriptor address a SHIFTED window of a halo tile that one TMA box
// load wrote with SWIZZLE_128B?  The halo tile is [HR rows][HP pixels][64 bf16] (128 B per pixel); output tile = 16 image
// rows x 8 pixels = 128 GEMM rows; tap (dh, dw) starts at pixel ((dh+1)*HP + (dw+1)) -> start address not 1024-aligned,
// 8-row groups HP*128 B apart (SBO).  Tries descriptor base_offset = 0 and = (start >> 7) & 7.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o umma_halo_probe tools/umma_halo_probe.cu -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../rep-yolo_b200/csrc/ptx.cuh"

using namespace ry;

constexpr int HP = 10, HR = 18, KC = 64, NB = 64;

__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, uint32_t sbo, uint32_t base_off) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(base_off & 7) << 49;
    d |= (uint64_t)2 << 61;
    return d;
}

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap bmap,
                                                float *out, int mode) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sA = smem;                          // HR*HP*128 = 23040 B
    uint8_t *sB = smem + 24 * 1024;              // 64 x 128 B
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 40 * 1024);
    uint64_t *mbar = bar + 1;
    uint32_t *slot = reinterpret_cast<uint32_t *>(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        ptx::mbar_init(bar, 1);
        ptx::mbar_init(mbar, 1);
        ptx::fence_mbar_init();
    }
    if (warp == 0) {
        ptx::tmem_alloc(slot, 64);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        ptx::mbar_expect_tx(bar, HR * HP * 128 + NB * 128);
        ptx::tma_load_4d(sA, &amap, bar, 0, 0, 0, 0);
        ptx::tma_load_2d(sB, &bmap, bar, 0, 0);
        ptx::mbar_wait(bar, 0);
        ptx::tc_fence_after();
    }
    __syncthreads();
    uint32_t ph = 0;
    for (int tap = 0; tap < 9; ++tap) {
        const int dh = tap / 3 - 1, dw = tap % 3 - 1;
        if (threadIdx.x == 0) {
            const uint32_t a0 = ptx::smem_u32(sA) + ((dh + 1) * HP + (dw + 1)) * 128;
            const uint32_t b0 = ptx::smem_u32(sB);
            const uint32_t idesc = ptx::umma_idesc_bf16(128, NB);
            for (int k = 0; k < KC / 16; ++k) {
                const uint32_t aa = a0 + k * 32;
                const uint32_t bo = mode == 0 ? 0u : ((aa >> 7) & 7u);
                ptx::umma_bf16(tmem, desc_sw128(aa, HP * 128, bo), desc_sw128(b0 + k * 32, 1024, 0), idesc, k > 0);
            }
            ptx::umma_commit(mbar);
            ptx::mbar_wait(mbar, ph);
            ptx::tc_fence_after();
        }
        ph ^= 1;
        __syncthreads();
        ptx::tc_fence_after();
        const int row = warp * 32 + lane;
        for (int c = 0; c < NB / 16; ++c) {
            float v[16];
            ptx::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c * 16, v);
            for (int i = 0; i < 16; ++i) out[((size_t)tap * 128 + row) * NB + c * 16 + i] = v[i];
        }
        ptx::tc_fence_before();
        __syncthreads();
    }
    if (warp == 0) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem, 64);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    void *fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fp);
    std::vector<__nv_bfloat16> hA((size_t)HR * HP * KC), hB((size_t)NB * KC);
    std::vector<float> fA(hA.size()), fB(hB.size());
    srand(1);
    for (size_t i = 0; i < hA.size(); ++i) { hA[i] = __float2bfloat16((rand() % 17 - 8) / 8.0f); fA[i] = __bfloat162float(hA[i]); }
    for (size_t i = 0; i < hB.size(); ++i) { hB[i] = __float2bfloat16((rand() % 13 - 6) / 4.0f); fB[i] = __bfloat162float(hB[i]); }
    __nv_bfloat16 *dA, *dB;
    float *dO;
    cudaMalloc(&dA, hA.size() * 2);
    cudaMalloc(&dB, hB.size() * 2);
    cudaMalloc(&dO, 9 * 128 * NB * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap amap, bmap;
    cuuint32_t es[4] = {1, 1, 1, 1};
    {
        cuuint64_t dims[4] = {KC, HP, HR, 1};
        cuuint64_t str[3] = {KC * 2, (cuuint64_t)HP * KC * 2, (cuuint64_t)HR * HP * KC * 2};
        cuuint32_t box[4] = {KC, HP, HR, 1};
        CUresult r = enc(&amap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dA, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r) { printf("encode A failed %d\n", (int)r); return 1; }
    }
    {
        cuuint64_t dims[2] = {KC, NB};
        cuuint64_t str[1] = {KC * 2};
        cuuint32_t box[2] = {KC, NB};
        CUresult r = enc(&bmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r) { printf("encode B failed %d\n", (int)r); return 1; }
    }
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
    std::vector<float> hO((size_t)9 * 128 * NB);
    for (int mode = 0; mode < 2; ++mode) {
        cudaMemset(dO, 0, hO.size() * 4);
        probe<<<1, 128, 44 * 1024>>>(amap, bmap, dO, mode);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d: kernel failed: %s\n", mode, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost);
        for (int tap = 0; tap < 9; ++tap) {
            const int dh = tap / 3 - 1, dw = tap % 3 - 1;
            int bad = 0;
            double maxd = 0;
            for (int r = 0; r < 128; ++r) {
                const int ih = r / 8 + dh + 1, iw = r % 8 + dw + 1;
                for (int n = 0; n < NB; ++n) {
                    float ref = 0;
                    for (int k = 0; k < KC; ++k) ref += fA[((size_t)ih * HP + iw) * KC + k] * fB[(size_t)n * KC + k];
                    const double d = fabs((double)ref - hO[((size_t)tap * 128 + r) * NB + n]);
                    if (d > 1e-3) ++bad;
                    if (d > maxd) maxd = d;
                }
            }
            printf("mode %d (base_offset %s) tap (%+d,%+d): mismatches %d / %d, max |d| %.4f\n", mode,
                   mode ? "= (addr>>7)&7" : "= 0", dh, dw, bad, 128 * NB, maxd);
        }
    }
    return 0;
}
