// Shared device/host helpers for the Caption-HN sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define CAPHN_OK 0
#define CAPHN_EINVAL (-1)

// Number of kernels this library has launched (bench.py reports it as gpu_launches); defined in api_misc.cu.
extern unsigned long long caphn_launch_counter;

// After a kernel launch at the end of an entry point: count it and return the cudaError_t as int (0 == ok).
#define CAPHN_RETURN_LAST()                    \
    do {                                       \
        ++caphn_launch_counter;                \
        cudaError_t e__ = cudaGetLastError();  \
        return (int)e__;                       \
    } while (0)

// After a kernel launch in the middle of an entry point.
#define CAPHN_LAUNCH_CHECK()                       \
    do {                                           \
        ++caphn_launch_counter;                    \
        cudaError_t e__ = cudaGetLastError();      \
        if (e__ != cudaSuccess) return (int)e__;   \
    } while (0)

#define CAPHN_CHECK(expr)                          \
    do {                                           \
        cudaError_t e__ = (expr);                  \
        if (e__ != cudaSuccess) return (int)e__;   \
    } while (0)

namespace caphn {

constexpr int kNumSMs = 148;  // B200

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Streaming 128-bit load that does not allocate in L1 (weights touched once per step).
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ float ldg_stream1(const float* p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream4(float* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

static inline int ceil_div(long a, long b) { return (int)((a + b - 1) / b); }

}  // namespace caphn
