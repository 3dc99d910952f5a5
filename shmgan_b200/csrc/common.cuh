// common.cuh -- shared device/host helpers for libshmgan (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/shmgan.h"

typedef __nv_bfloat16 bf16;

// ---- error plumbing -------------------------------------------------------------------------
void shm_set_error(const char* fmt, ...);
#define SHM_FAIL(code, ...) do { shm_set_error(__VA_ARGS__); return (code); } while (0)
#define SHM_REQUIRE(cond, ...) do { if (!(cond)) SHM_FAIL(SHM_EINVAL, __VA_ARGS__); } while (0)
#define SHM_CHECK_LAUNCH(name) do { cudaError_t e__ = cudaGetLastError(); \
    if (e__ != cudaSuccess) SHM_FAIL(SHM_ECUDA, "%s: %s", name, cudaGetErrorString(e__)); } while (0)

static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
int shm_num_sms();

// TF 'SAME' padding rule: pad_before for (size, k, stride)
static inline int same_pad_before(int size, int k, int s) {
    int out = (size + s - 1) / s;
    int total = (out - 1) * s + k - size;
    if (total < 0) total = 0;
    return total / 2;
}

// ---- typed element access -------------------------------------------------------------------
__device__ __forceinline__ float ldf(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ldf(const bf16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stf(float* p, float v) { *p = v; }
__device__ __forceinline__ void stf(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

// 4 consecutive elements (16B for float, 8B for bf16); pointer must be suitably aligned
__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld4(const bf16* p) {
    uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&u.x);
    __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&u.y);
    float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st4(bf16* p, float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
    __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&a);
    u.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = u;
}
template <typename T> __device__ __forceinline__ bool aligned4(const T* p) {
    return (reinterpret_cast<uintptr_t>(p) & (sizeof(T) * 4 - 1)) == 0;
}

__device__ __forceinline__ float act_fwd(float v, int act) {
    switch (act) {
        case SHM_ACT_LRELU:   return v > 0.f ? v : 0.2f * v;
        case SHM_ACT_RELU:    return v > 0.f ? v : 0.f;
        case SHM_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
        default:              return v;
    }
}
// derivative expressed through the POST-activation value y
__device__ __forceinline__ float act_grad_from_post(float y, int act) {
    switch (act) {
        case SHM_ACT_LRELU:   return y > 0.f ? 1.f : 0.2f;
        case SHM_ACT_RELU:    return y > 0.f ? 1.f : 0.f;
        case SHM_ACT_SIGMOID: return y * (1.f - y);
        default:              return 1.f;
    }
}

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

// block-wide sum of one double per thread; result valid in thread 0.  blockDim.x multiple of 32, <= 1024.
__device__ __forceinline__ double block_sum(double v, double* smem32) {
    v = warp_sum(v);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) smem32[w] = v;
    __syncthreads();
    double r = 0.0;
    if (w == 0) {
        int nw = (blockDim.x + 31) >> 5;
        r = lane < nw ? smem32[lane] : 0.0;
        r = warp_sum(r);
    }
    __syncthreads();
    return r;
}

#define DISPATCH_DTYPE(dtype, T, ...) \
    if ((dtype) == SHM_F32) { typedef float T; __VA_ARGS__ } \
    else if ((dtype) == SHM_BF16) { typedef bf16 T; __VA_ARGS__ } \
    else SHM_FAIL(SHM_EINVAL, "bad dtype %d", (int)(dtype));
