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

// ---------------------------------------------------------------------------------------------------------------------
// 1x1 convolution to ONE output channel (the generator's output layer, ShmGANwithSSpecSeg.py:326, and SpecSeg's sigmoid
// head, SpecSeg.py:88): a per-pixel dot product over C channels -- pure HBM bandwidth (C*e bytes in, e bytes out per pixel),
// so it is served by a coalesced 128-bit-vectorised kernel with sub-warp shuffle reductions instead of a GEMM tile that
// would waste 63/64 of its columns.  bf16 activations, C in {8,16,32,64,128,256}.
// ---------------------------------------------------------------------------------------------------------------------
namespace {

__device__ __forceinline__ void unpack8(const uint4& u, float v[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(h[j]); v[2 * j] = f.x; v[2 * j + 1] = f.y; }
}

// TPP threads per pixel, each owning 8 consecutive channels
template <int TPP>
__global__ void __launch_bounds__(256) pw1_fwd_kernel(const bf16* __restrict__ x, int ldx, const float* __restrict__ w,
                                                      const float* __restrict__ bias, int act, bf16* __restrict__ y, long long npix) {
    const int part = threadIdx.x % TPP;
    float wv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) wv[j] = __ldg(w + part * 8 + j);
    const float b = bias ? __ldg(bias) : 0.f;
    const long long ppb = 256 / TPP;
    // the trip count is uniform over the block (p0), so the full-mask shuffles below are always converged.
    // Four pixels per trip while four full strides remain: one 16-byte load in flight per thread left the kernel latency-bound (4.1 TB/s).
    const long long stride = (long long)gridDim.x * ppb;
    long long p0 = (long long)blockIdx.x * ppb;
    for (; p0 + 3 * stride + ppb <= npix; p0 += 4 * stride) {
        const long long p = p0 + threadIdx.x / TPP;
        uint4 xv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) xv[u] = __ldg(reinterpret_cast<const uint4*>(x + (p + u * stride) * ldx + part * 8));
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float v[8], s = 0.f;
            unpack8(xv[u], v);
#pragma unroll
            for (int j = 0; j < 8; ++j) s = fmaf(v[j], wv[j], s);
#pragma unroll
            for (int o = TPP / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (part == 0) y[p + u * stride] = __float2bfloat16_rn(act_fwd(s + b, act));
        }
    }
    for (; p0 < npix; p0 += stride) {
        const long long p = p0 + threadIdx.x / TPP;
        const bool ok = p < npix;
        float s = 0.f;
        if (ok) {
            float v[8];
            unpack8(__ldg(reinterpret_cast<const uint4*>(x + p * ldx + part * 8)), v);
#pragma unroll
            for (int j = 0; j < 8; ++j) s = fmaf(v[j], wv[j], s);
        }
#pragma unroll
        for (int o = TPP / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (ok && part == 0) y[p] = __float2bfloat16_rn(act_fwd(s + b, act));
    }
}

// fused backward: dpre = dy * act'(y);  dx[p, c] = dpre * w[c];  dw[c] += sum_p x[p, c] dpre;  dbias += sum_p dpre
template <int TPP>
__global__ void __launch_bounds__(256) pw1_bwd_kernel(const bf16* __restrict__ x, int ldx, const float* __restrict__ w,
        const bf16* __restrict__ dy, const bf16* __restrict__ y, int act, bf16* __restrict__ dx, int lddx,
        float* __restrict__ dw, float* __restrict__ dbias, long long npix) {
    __shared__ float sdw[TPP * 8 + 1];
    for (int i = threadIdx.x; i < TPP * 8 + 1; i += 256) sdw[i] = 0.f;
    __syncthreads();
    const int part = threadIdx.x % TPP;
    float wv[8], acc[8], accb = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { wv[j] = __ldg(w + part * 8 + j); acc[j] = 0.f; }
    const long long ppb = 256 / TPP;
    const long long stride = (long long)gridDim.x * ppb;
    long long p = (long long)blockIdx.x * ppb + threadIdx.x / TPP;
    // four pixels per trip (four 16-byte loads in flight per thread), then the remainder one by one
    for (; p + 3 * stride < npix; p += 4 * stride) {
        uint4 xv[4];
        float g[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) xv[u] = __ldg(reinterpret_cast<const uint4*>(x + (p + u * stride) * ldx + part * 8));
#pragma unroll
        for (int u = 0; u < 4; ++u) g[u] = __bfloat162float(dy[p + u * stride]) * act_grad_from_post(__bfloat162float(y[p + u * stride]), act);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float v[8];
            unpack8(xv[u], v);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaf(v[j], g[u], acc[j]);
            if (part == 0) accb += g[u];
            if (dx) {
                __nv_bfloat162 h[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(g[u] * wv[2 * j], g[u] * wv[2 * j + 1]);
                *reinterpret_cast<uint4*>(dx + (p + u * stride) * lddx + part * 8) = *reinterpret_cast<uint4*>(h);
            }
        }
    }
    for (; p < npix; p += stride) {
        const float g = __bfloat162float(dy[p]) * act_grad_from_post(__bfloat162float(y[p]), act);
        float v[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(x + p * ldx + part * 8)), v);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(v[j], g, acc[j]);
        if (part == 0) accb += g;
        if (dx) {
            __nv_bfloat162 h[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(g * wv[2 * j], g * wv[2 * j + 1]);
            *reinterpret_cast<uint4*>(dx + p * lddx + part * 8) = *reinterpret_cast<uint4*>(h);
        }
    }
    // lanes sharing `part` (stride TPP inside the warp) hold partials of the same channels
#pragma unroll
    for (int o = 16; o >= TPP; o >>= 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
        accb += __shfl_xor_sync(0xffffffffu, accb, o);
    }
    if ((threadIdx.x & 31) < TPP) {
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(&sdw[part * 8 + j], acc[j]);
        if (part == 0) atomicAdd(&sdw[TPP * 8], accb);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < TPP * 8; i += 256) atomicAdd(dw + i, sdw[i]);
    if (threadIdx.x == 0 && dbias) atomicAdd(dbias, sdw[TPP * 8]);
}

inline int pw1_grid(long long npix, int tpp) {
    long long g = cdiv64(npix, 256 / tpp);
    const long long cap = (long long)shm_num_sms() * 8;
    return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}

inline int pw1_check(const void* x, int ldx, int C, int dtype, const char* who) {
    if (dtype != SHM_BF16) SHM_FAIL(SHM_EUNSUPPORTED, "%s: bf16 only (fp32 mode uses shm_conv2d_*)", who);
    if (!(C == 8 || C == 16 || C == 32 || C == 64 || C == 128 || C == 256)) SHM_FAIL(SHM_EUNSUPPORTED, "%s: C=%d not in {8,...,256}", who, C);
    if (ldx % 8 != 0 || (reinterpret_cast<uintptr_t>(x) & 15) != 0) SHM_FAIL(SHM_EINVAL, "%s: x must be 16-byte aligned with ld %% 8 == 0", who);
    return SHM_OK;
}

#define PW1_DISPATCH(C, KERNEL, ...) \
    switch ((C) / 8) { \
        case 1:  KERNEL<1> __VA_ARGS__; break;  case 2:  KERNEL<2> __VA_ARGS__; break; \
        case 4:  KERNEL<4> __VA_ARGS__; break;  case 8:  KERNEL<8> __VA_ARGS__; break; \
        case 16: KERNEL<16> __VA_ARGS__; break; default: KERNEL<32> __VA_ARGS__; break; }

}  // namespace

extern "C" int shm_pw1_fwd(const void* x, int ldx, int C, const float* w, const float* bias, int act, void* y, int64_t npix, int dtype, void* stream) {
    SHM_REQUIRE(x && w && y && npix > 0, "shm_pw1_fwd: bad args");
    if (int rc = pw1_check(x, ldx, C, dtype, "shm_pw1_fwd")) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int g = pw1_grid(npix, C / 8);
    PW1_DISPATCH(C, pw1_fwd_kernel, <<<g, 256, 0, st>>>((const bf16*)x, ldx, w, bias, act, (bf16*)y, npix))
    SHM_CHECK_LAUNCH("pw1_fwd_kernel");
    return SHM_OK;
}

extern "C" int shm_pw1_bwd(const void* x, int ldx, int C, const float* w, const void* dy, const void* y, int act, void* dx, int lddx,
                           float* dw, float* dbias, int64_t npix, int dtype, void* stream) {
    SHM_REQUIRE(x && w && dy && y && dw && npix > 0, "shm_pw1_bwd: bad args");
    if (int rc = pw1_check(x, ldx, C, dtype, "shm_pw1_bwd")) return rc;
    SHM_REQUIRE(!dx || (lddx % 8 == 0 && (reinterpret_cast<uintptr_t>(dx) & 15) == 0), "shm_pw1_bwd: dx must be 16-byte aligned with ld %% 8 == 0");
    cudaStream_t st = (cudaStream_t)stream;
    const int g = pw1_grid(npix, C / 8);
    PW1_DISPATCH(C, pw1_bwd_kernel, <<<g, 256, 0, st>>>((const bf16*)x, ldx, w, (const bf16*)dy, (const bf16*)y, act, (bf16*)dx, lddx, dw, dbias, npix))
    SHM_CHECK_LAUNCH("pw1_bwd_kernel");
    return SHM_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// 3x3 stride-1 SAME convolution to ONE output channel (the discriminator's real/fake head, ShmGANwithSSpecSeg.py:365-369:
// Conv2D(1, 3, use_bias=False) + LeakyReLU on the [B, S/32, S/32, 1024] feature map).  K = 9*C = 9216 but N = 1: a GEMM tile
// would waste 127/128 of its columns and the exact-fp32 SIMT kernel took 1.6 ms per step on it; as a warp-per-pixel dot
// product it is a ~25 MB read.  bf16 activations, C % 8 == 0; w = the Keras (3,3,C,1) kernel as fp32 [9][C].
// ---------------------------------------------------------------------------------------------------------------------
namespace {

__global__ void __launch_bounds__(256) c3to1_fwd_kernel(const bf16* __restrict__ x, int N, int H, int W, int C, int ldx,
        const float* __restrict__ w, const float* __restrict__ bias, int act, bf16* __restrict__ y) {
    const int lane = threadIdx.x & 31;
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const long long total = (long long)N * H * W;
    const float b = bias ? __ldg(bias) : 0.f;
    for (long long pix = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); pix < total; pix += nwarps) {
        const int n = (int)(pix / (H * W)); const int r = (int)(pix - (long long)n * H * W);
        const int oy = r / W, ox = r - oy * W;
        float s = 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const int iy = oy + t / 3 - 1, ix = ox + t % 3 - 1;
            if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;          // warp-uniform
            const bf16* xp = x + ((long long)(n * H + iy) * W + ix) * ldx;
            const float* wp = w + t * C;
            for (int g = lane; g < C / 8; g += 32) {
                float v[8];
                unpack8(__ldg(reinterpret_cast<const uint4*>(xp + g * 8)), v);
                const float4 w0 = __ldg(reinterpret_cast<const float4*>(wp + g * 8)), w1 = __ldg(reinterpret_cast<const float4*>(wp + g * 8 + 4));
                s = fmaf(v[0], w0.x, s); s = fmaf(v[1], w0.y, s); s = fmaf(v[2], w0.z, s); s = fmaf(v[3], w0.w, s);
                s = fmaf(v[4], w1.x, s); s = fmaf(v[5], w1.y, s); s = fmaf(v[6], w1.z, s); s = fmaf(v[7], w1.w, s);
            }
        }
        s = warp_sum(s);
        if (lane == 0) y[pix] = __float2bfloat16_rn(act_fwd(s + b, act));
    }
}

// dx[n,iy,ix,c] = sum_t dpre[n, iy-(ty-1), ix-(tx-1)] * w[t][c]
__global__ void __launch_bounds__(256) c3to1_dgrad_kernel(const bf16* __restrict__ dpre, int N, int H, int W, int C,
        const float* __restrict__ w, bf16* __restrict__ dx, int lddx) {
    const int G = C / 8;
    const long long total = (long long)N * H * W * G;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(i % G); const long long pix = i / G;
        const int n = (int)(pix / (H * W)); const int r = (int)(pix - (long long)n * H * W);
        const int iy = r / W, ix = r - iy * W;
        float a[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const int oy = iy - (t / 3 - 1), ox = ix - (t % 3 - 1);
            if (oy < 0 || oy >= H || ox < 0 || ox >= W) continue;
            const float d = __bfloat162float(dpre[(long long)(n * H + oy) * W + ox]);
            const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + t * C + g * 8)), w1 = __ldg(reinterpret_cast<const float4*>(w + t * C + g * 8 + 4));
            a[0] = fmaf(d, w0.x, a[0]); a[1] = fmaf(d, w0.y, a[1]); a[2] = fmaf(d, w0.z, a[2]); a[3] = fmaf(d, w0.w, a[3]);
            a[4] = fmaf(d, w1.x, a[4]); a[5] = fmaf(d, w1.y, a[5]); a[6] = fmaf(d, w1.z, a[6]); a[7] = fmaf(d, w1.w, a[7]);
        }
        __nv_bfloat162 h[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(a[2 * j], a[2 * j + 1]);
        *reinterpret_cast<uint4*>(dx + pix * lddx + g * 8) = *reinterpret_cast<uint4*>(h);
    }
}

// dw[t][c] += sum over pixels of x[n, oy+ty-1, ox+tx-1, c] * dpre[n, oy, ox]; one block per chunk of input pixels of one image
__global__ void __launch_bounds__(128) c3to1_wgrad_kernel(const bf16* __restrict__ x, int N, int H, int W, int C, int ldx,
        const bf16* __restrict__ dpre, float* __restrict__ dw, int chunks_per_img, int ppc) {
    const int n = blockIdx.x / chunks_per_img, ch = blockIdx.x - n * chunks_per_img;
    const int pbeg = ch * ppc, pend = min(pbeg + ppc, H * W);
    for (int g = threadIdx.x; g < C / 8; g += blockDim.x) {
        float acc[9][8];
#pragma unroll
        for (int t = 0; t < 9; ++t)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[t][j] = 0.f;
        for (int p = pbeg; p < pend; ++p) {
            const int iy = p / W, ix = p - iy * W;
            float v[8];
            unpack8(__ldg(reinterpret_cast<const uint4*>(x + ((long long)n * H * W + p) * ldx + g * 8)), v);
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const int oy = iy - (t / 3 - 1), ox = ix - (t % 3 - 1);
                if (oy < 0 || oy >= H || ox < 0 || ox >= W) continue;      // block-uniform
                const float d = __bfloat162float(dpre[(long long)(n * H + oy) * W + ox]);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[t][j] = fmaf(v[j], d, acc[t][j]);
            }
        }
#pragma unroll
        for (int t = 0; t < 9; ++t)
#pragma unroll
            for (int j = 0; j < 8; ++j) atomicAdd(dw + t * C + g * 8 + j, acc[t][j]);
    }
}

inline int c3to1_check(const void* x, int C, int ldx, int dtype, const char* who) {
    if (dtype != SHM_BF16) SHM_FAIL(SHM_EUNSUPPORTED, "%s: bf16 only (fp32 mode uses shm_conv2d_*)", who);
    if (C % 8 != 0 || ldx % 8 != 0 || (reinterpret_cast<uintptr_t>(x) & 15) != 0) SHM_FAIL(SHM_EINVAL, "%s: needs C %% 8 == 0, ld %% 8 == 0 and 16-byte alignment", who);
    return SHM_OK;
}

}  // namespace

extern "C" int shm_c3to1_fwd(const void* x, int N, int H, int W, int C, int ldx, const float* w, const float* bias, int act, void* y, int dtype, void* stream) {
    SHM_REQUIRE(x && w && y && N > 0 && H > 0 && W > 0 && C > 0, "shm_c3to1_fwd: bad args");
    if (int rc = c3to1_check(x, C, ldx, dtype, "shm_c3to1_fwd")) return rc;
    SHM_REQUIRE((reinterpret_cast<uintptr_t>(w) & 15) == 0, "shm_c3to1_fwd: w must be 16-byte aligned");
    long long g = cdiv64((long long)N * H * W, 8);
    if (g > shm_num_sms() * 8) g = shm_num_sms() * 8;
    c3to1_fwd_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, N, H, W, C, ldx, w, bias, act, (bf16*)y);
    SHM_CHECK_LAUNCH("c3to1_fwd_kernel");
    return SHM_OK;
}

extern "C" int shm_c3to1_dgrad(const void* dpre, int N, int H, int W, int C, const float* w, void* dx, int lddx, int dtype, void* stream) {
    SHM_REQUIRE(dpre && w && dx && N > 0 && H > 0 && W > 0 && C > 0, "shm_c3to1_dgrad: bad args");
    if (int rc = c3to1_check(dx, C, lddx, dtype, "shm_c3to1_dgrad")) return rc;
    SHM_REQUIRE((reinterpret_cast<uintptr_t>(w) & 15) == 0, "shm_c3to1_dgrad: w must be 16-byte aligned");
    long long g = cdiv64((long long)N * H * W * (C / 8), 256);
    if (g > shm_num_sms() * 16) g = shm_num_sms() * 16;
    c3to1_dgrad_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>((const bf16*)dpre, N, H, W, C, w, (bf16*)dx, lddx);
    SHM_CHECK_LAUNCH("c3to1_dgrad_kernel");
    return SHM_OK;
}

extern "C" int shm_c3to1_wgrad(const void* x, int N, int H, int W, int C, int ldx, const void* dpre, float* dw, int dtype, void* stream) {
    SHM_REQUIRE(x && dpre && dw && N > 0 && H > 0 && W > 0 && C > 0, "shm_c3to1_wgrad: bad args");
    if (int rc = c3to1_check(x, C, ldx, dtype, "shm_c3to1_wgrad")) return rc;
    // ~2 waves of blocks; at least 16 pixels per block so that the 72 atomics per thread amortise
    int chunks = cdiv(shm_num_sms() * 2, N);
    int ppc = cdiv(H * W, chunks < 1 ? 1 : chunks);
    if (ppc < 16) ppc = 16;
    chunks = cdiv(H * W, ppc);
    c3to1_wgrad_kernel<<<N * chunks, 128, 0, (cudaStream_t)stream>>>((const bf16*)x, N, H, W, C, ldx, (const bf16*)dpre, dw, chunks, ppc);
    SHM_CHECK_LAUNCH("c3to1_wgrad_kernel");
    return SHM_OK;
}
