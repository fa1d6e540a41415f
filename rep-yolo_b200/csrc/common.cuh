// Shared host/device helpers for the repyolo_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <string>

namespace ry {

void set_error(const std::string &msg);   // plan.cu

#define RY_CUDA(expr)                                                                                       \
    do {                                                                                                    \
        cudaError_t err__ = (expr);                                                                         \
        if (err__ != cudaSuccess) {                                                                         \
            ry::set_error(std::string(#expr) + ": " + cudaGetErrorString(err__) + " (" + __FILE__ + ":" +   \
                          std::to_string(__LINE__) + ")");                                                  \
            return 1;                                                                                       \
        }                                                                                                   \
    } while (0)

#define RY_FAIL(msg)                        \
    do {                                    \
        ry::set_error(std::string(msg));    \
        return 1;                           \
    } while (0)

// Multiprocessor count of the CURRENT device (cached per device; 148 on B200).  Grids of the persistent kernels are sized
// from this, never from a compile-time constant.
int num_sms();                              // plan.cu

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device function attribute: set it once per (kernel, device),
// not once per process.  `kernel` may be a template instantiation (parenthesise it).  Evaluates to a cudaError_t.
#define RY_ENSURE_DYN_SMEM(kernel, bytes)                                                                         \
    ([&]() -> cudaError_t {                                                                                        \
        static std::atomic<unsigned long long> done__{0ull};                                                       \
        int dev__ = 0;                                                                                             \
        cudaGetDevice(&dev__);                                                                                     \
        const unsigned long long bit__ = 1ull << (dev__ & 63);                                                     \
        if (done__.load(std::memory_order_acquire) & bit__) return cudaSuccess;                                    \
        const cudaError_t e__ = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);  \
        if (e__ == cudaSuccess) done__.fetch_or(bit__, std::memory_order_release);                                 \
        return e__;                                                                                                \
    }())

// ---- programmatic dependent launch (PDL) ----
// Every kernel of the forward pass is launched with the programmatic-stream-serialization attribute and
//   * calls pdl_trigger() at its top: the NEXT kernel's CTAs may be scheduled as soon as SM resources free up, so its
//     prologue (barrier init, TMEM alloc, constant weight loads) overlaps this kernel's tail;
//   * calls pdl_wait() before it touches anything a previous kernel wrote (waits for full completion + visibility of
//     the prerequisite grid, so ordering is transitively that of a plain stream).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}


__host__ __device__ inline int cdiv(int a, int b) { return (a + b - 1) / b; }

__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
__device__ __forceinline__ float relu6_f(float x) { return fminf(fmaxf(x, 0.0f), 6.0f); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
    __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162 *>(&u);
    return __bfloat1622float2(v);
}
__device__ __forceinline__ uint32_t bf16x2_max(uint32_t a, uint32_t b) {
    __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162 *>(&a), *reinterpret_cast<__nv_bfloat162 *>(&b));
    return *reinterpret_cast<uint32_t *>(&r);
}
__device__ __forceinline__ uint4 bf16x8_max(uint4 a, uint4 b) {
    return make_uint4(bf16x2_max(a.x, b.x), bf16x2_max(a.y, b.y), bf16x2_max(a.z, b.z), bf16x2_max(a.w, b.w));
}

}  // namespace ry
