// dense.cu -- the discriminator's domain-classification head: Flatten (H,W,C order) -> Dense(5, no bias).
// ShmGANwithSSpecSeg.py:371-375.  K = (S/32)^2 * 1024 (65536 at 256x256), J = 5: a bandwidth-bound GEMV family,
// done with warp-shuffle reductions instead of a tensor-core GEMM.
#include "common.cuh"

namespace {

constexpr int DJ = 8;   // max outputs supported

template <typename T>
__global__ void __launch_bounds__(256) dense_fwd_kernel(const T* __restrict__ x, const float* __restrict__ w, float* __restrict__ out,
                                                        int K, int J, int kpb) {
    __shared__ float red[8][DJ];
    const int b = blockIdx.y;
    const int kbeg = blockIdx.x * kpb, kend = min(kbeg + kpb, K);
    float acc[DJ];
#pragma unroll
    for (int j = 0; j < DJ; ++j) acc[j] = 0.f;
    const T* xb = x + (long long)b * K;
    for (int k = kbeg + threadIdx.x; k < kend; k += blockDim.x) {
        const float xv = ldf(xb + k);
        const float* wr = w + (long long)k * J;
#pragma unroll
        for (int j = 0; j < DJ; ++j) if (j < J) acc[j] = fmaf(xv, __ldg(wr + j), acc[j]);
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < DJ; ++j) {
        const float s = warp_sum(acc[j]);
        if (lane == 0) red[wid][j] = s;
    }
    __syncthreads();
    if (threadIdx.x < J) {
        float s = 0.f;
        for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x];
        atomicAdd(out + b * J + threadIdx.x, s);
    }
}

template <typename T>
__global__ void dense_dgrad_kernel(const float* __restrict__ dout, const float* __restrict__ w, T* __restrict__ dx, int K, int J, long long total) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / K), k = (int)(i - (long long)b * K);
        float acc = 0.f;
        for (int j = 0; j < J; ++j) acc = fmaf(__ldg(dout + b * J + j), __ldg(w + (long long)k * J + j), acc);
        stf(dx + i, acc);
    }
}

template <typename T>
__global__ void dense_wgrad_kernel(const T* __restrict__ x, const float* __restrict__ dout, float* __restrict__ dw, int B, int K, int J) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    float acc[DJ];
#pragma unroll
    for (int j = 0; j < DJ; ++j) acc[j] = 0.f;
    for (int b = 0; b < B; ++b) {
        const float xv = ldf(x + (long long)b * K + k);
#pragma unroll
        for (int j = 0; j < DJ; ++j) if (j < J) acc[j] = fmaf(xv, __ldg(dout + b * J + j), acc[j]);
    }
#pragma unroll
    for (int j = 0; j < DJ; ++j) if (j < J) dw[(long long)k * J + j] += acc[j];
}

}  // namespace

extern "C" int shm_dense_fwd(const void* x, const float* w, float* out, int B, int K, int J, int dtype, void* stream) {
    SHM_REQUIRE(x && w && out && B > 0 && K > 0 && J > 0 && J <= DJ, "shm_dense_fwd: bad args (J <= %d)", DJ);
    cudaStream_t st = (cudaStream_t)stream;
    if (cudaMemsetAsync(out, 0, sizeof(float) * B * J, st) != cudaSuccess) SHM_FAIL(SHM_ECUDA, "shm_dense_fwd: memset failed");
    int kpb = cdiv(K, cdiv(shm_num_sms() * 4, B));
    if (kpb < 2048) kpb = 2048;
    DISPATCH_DTYPE(dtype, T, {
        dense_fwd_kernel<T><<<dim3(cdiv(K, kpb), B), 256, 0, st>>>((const T*)x, w, out, K, J, kpb);
        SHM_CHECK_LAUNCH("dense_fwd_kernel");
        return SHM_OK;
    })
}

extern "C" int shm_dense_dgrad(const float* dout, const float* w, void* dx, int B, int K, int J, int dtype, void* stream) {
    SHM_REQUIRE(dout && w && dx && B > 0 && K > 0 && J > 0 && J <= DJ, "shm_dense_dgrad: bad args");
    const long long total = (long long)B * K;
    long long g = cdiv64(total, 256);
    if (g > shm_num_sms() * 16) g = shm_num_sms() * 16;
    DISPATCH_DTYPE(dtype, T, {
        dense_dgrad_kernel<T><<<(int)g, 256, 0, (cudaStream_t)stream>>>(dout, w, (T*)dx, K, J, total);
        SHM_CHECK_LAUNCH("dense_dgrad_kernel");
        return SHM_OK;
    })
}

extern "C" int shm_dense_wgrad(const void* x, const float* dout, float* dw, int B, int K, int J, int dtype, void* stream) {
    SHM_REQUIRE(x && dout && dw && B > 0 && K > 0 && J > 0 && J <= DJ, "shm_dense_wgrad: bad args");
    DISPATCH_DTYPE(dtype, T, {
        dense_wgrad_kernel<T><<<cdiv(K, 128), 128, 0, (cudaStream_t)stream>>>((const T*)x, dout, dw, B, K, J);
        SHM_CHECK_LAUNCH("dense_wgrad_kernel");
        return SHM_OK;
    })
}
