// Shared host-side helpers of libskyb200: error reporting and launch checks.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>

#include "../../include/skyb200.h"

namespace skyb200 {

char* last_error_buffer();  // thread-local, 512 bytes

inline int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(last_error_buffer(), 512, fmt, ap);
    va_end(ap);
    return code;
}

#define SKYB200_CUDA_CHECK(expr)                                                                         \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            return ::skyb200::fail(SKYB200_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                   __FILE__, __LINE__);                                                  \
    } while (0)

#define SKYB200_REQUIRE(cond, msg)                                                    \
    do {                                                                              \
        if (!(cond)) return ::skyb200::fail(SKYB200_EINVAL, "%s: %s", __func__, msg); \
    } while (0)

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(SKYB200_ECUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
    return SKYB200_OK;
}

int num_sms();  // cached SM count of the current device

// ---- programmatic dependent launch (PDL): a kernel launched with launch_pdl() may start while its predecessor in the
// stream is still running; it must call pdl_wait() before touching anything the predecessor writes.  The predecessor
// calls pdl_launch_dependents() as early as it likes (it only releases the launch, not the data).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- device helpers shared by the memory-bound kernels
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum of `v`; result valid in every thread.  `sh` must hold >= 33 elements.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* sh) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();  // protect sh from a previous use
    if (lane == 0) sh[w] = v;
    __syncthreads();
    if (w == 0) {
        T x = (lane < nw) ? sh[lane] : T(0);
        x = warp_sum(x);
        if (lane == 0) sh[32] = x;
    }
    __syncthreads();
    return sh[32];
}
__device__ __forceinline__ float block_max(float v, float* sh) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    if (w == 0) {
        float x = (lane < nw) ? sh[lane] : -INFINITY;
        x = warp_max(x);
        if (lane == 0) sh[32] = x;
    }
    __syncthreads();
    return sh[32];
}
__device__ __forceinline__ float block_min(float v, float* sh) { return -block_max(-v, sh); }

// 128-bit streaming load (read once, do not pollute L1)
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

}  // namespace skyb200
