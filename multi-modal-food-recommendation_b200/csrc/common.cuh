// Shared helpers for the foodrec_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "foodrec_b200.h"

namespace fr {

void set_error(const char *fmt, ...);
void count_launch(int n = 1);
bool profiling();
// RAII: when profiling is enabled, brackets the kernel launches in its scope with CUDA events.
struct LaunchTimer {
    LaunchTimer(const char *name, cudaStream_t st);
    ~LaunchTimer();
    const char *name_;
    cudaStream_t st_;
    bool on_;
    cudaEvent_t a_, b_;
};

inline int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return FR_ECUDA;
    }
    count_launch();
    return FR_OK;
}

#define FR_REQUIRE(cond, ...)       \
    do {                            \
        if (!(cond)) {              \
            fr::set_error(__VA_ARGS__); \
            return FR_EINVAL;       \
        }                           \
    } while (0)

// 128-bit read-only gather of one embedding-row slice (rows written by an earlier launch).
__device__ __forceinline__ float4 ldg_f4(const float *p) {
    return __ldg(reinterpret_cast<const float4 *>(p));
}
// L2-only load: data produced by other CTAs of the SAME launch (never cached in L1).
__device__ __forceinline__ float4 ldcg_f4(const float *p) {
    return __ldcg(reinterpret_cast<const float4 *>(p));
}
__device__ __forceinline__ void fma4(float4 &a, float s, const float4 &x) {
    a.x = fmaf(s, x.x, a.x);
    a.y = fmaf(s, x.y, a.y);
    a.z = fmaf(s, x.z, a.z);
    a.w = fmaf(s, x.w, a.w);
}
__device__ __forceinline__ void add4(float4 &a, const float4 &x) {
    a.x += x.x; a.y += x.y; a.z += x.z; a.w += x.w;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

inline int num_sms() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

}  // namespace fr
