// norm.cu -- bandwidth kernels around the convolutions: instance norm (stats / apply / backward) fused with
// AvgPool2, skip-add and concat-slice placement; BatchNorm(eval)+MaxPool2 for SpecSeg; activation backward;
// max pooling; strided adds.  All 128-bit vectorised over the NHWC channel axis (4 x fp32 or 4 x bf16 per access).
// Replaces the ~13 eager TF ops per tfa InstanceNormalization (Generator_summary.txt:9-37) and the pooling /
// Add / Concatenate layers of ShmGANwithSSpecSeg.py:245-323, :358-359, :388 and SpecSeg.py:37-83.
#include "common.cuh"
#include "../../include/shmgan_tools.h"

namespace {

// channel-group geometry: TX groups of 4 channels per block row, TY pixel lanes, TX*TY == 256
struct CG { int TX, TY, tiles; };
inline CG cgeom(int C) {
    int g4 = C / 4;
    int tx = 1;
    while (tx < g4 && tx < 64) tx <<= 1;
    CG g; g.TX = tx; g.TY = 256 / tx; g.tiles = cdiv(g4, tx);
    return g;
}

__device__ __forceinline__ void stat_ab(const double* __restrict__ sums, int n, int C, int c, int HW, float eps,
                                        float& mean, float& rstd) {
    const double s = sums[((long long)n * C + c) * 2 + 0], q = sums[((long long)n * C + c) * 2 + 1];
    const double m = s / HW;
    double var = q / HW - m * m;
    if (var < 0.0) var = 0.0;
    mean = (float)m;
    rstd = (float)(1.0 / sqrt(var + (double)eps));
}

// cheap variant for the bf16 fast paths: one 16-byte load, fp64 only for the cancellation-prone variance, float rsqrt.
// (the exact version above costs ~100+ instructions per channel in fp64 division / sqrt sequences, which dominated the deep
// levels where a thread owns 8 channels but only ~20 pixels)
__device__ __forceinline__ void stat_ab_fast(const double* __restrict__ sums, int n, int C, int c, double inv_hw, float eps,
                                             float& mean, float& rstd) {
    const double2 sq = __ldg(reinterpret_cast<const double2*>(sums + ((long long)n * C + c) * 2));
    const double m = sq.x * inv_hw;
    double var = fma(-m, m, sq.y * inv_hw);
    if (var < 0.0) var = 0.0;
    mean = (float)m;
    rstd = rsqrtf((float)var + eps);
}

// ---------------------------------------------------------------------------------------------
// stats: sums[n][c] += (sum x, sum x^2)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) inorm_stats_kernel(const T* __restrict__ x, int HW, int C, int ldx,
                                                          double* __restrict__ sums, int ppb, int TX) {
    __shared__ float red[256][9];
    const int TY = 256 / TX;
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const int cg = blockIdx.y * TX + tx;
    const int n = blockIdx.z;
    const int pbeg = blockIdx.x * ppb;
    const int pend = min(pbeg + ppb, HW);
    float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
    if (cg * 4 < C) {
        const T* base = x + (long long)n * HW * ldx + cg * 4;
        for (int p = pbeg + ty; p < pend; p += TY) {
            const float4 v = ld4(base + (long long)p * ldx);
            s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
            q[0] = fmaf(v.x, v.x, q[0]); q[1] = fmaf(v.y, v.y, q[1]); q[2] = fmaf(v.z, v.z, q[2]); q[3] = fmaf(v.w, v.w, q[3]);
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { red[threadIdx.x][j] = s[j]; red[threadIdx.x][4 + j] = q[j]; }
    __syncthreads();
    for (int e = threadIdx.x; e < TX * 8; e += 256) {      // TX * 8 partial columns (up to 512) over 256 threads
        const int t = e / 8, j = e % 8;
        const int cgo = blockIdx.y * TX + t;
        if (cgo * 4 < C) {
            double acc = 0.0;
            for (int y = 0; y < TY; ++y) acc += (double)red[y * TX + t][j];
            const int c = cgo * 4 + (j & 3);
            atomicAdd(&sums[((long long)n * C + c) * 2 + (j >> 2)], acc);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// apply: y = (x-mean)*rstd*gamma + beta ; out = y (+ add) ; pooled = AvgPool2x2(y)
// ---------------------------------------------------------------------------------------------
template <typename T, bool POOL>
__global__ void __launch_bounds__(256) inorm_apply_kernel(const T* __restrict__ x, int H, int W, int C, int ldx,
        const double* __restrict__ sums, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
        const T* __restrict__ add, int ldadd, int nadd, T* __restrict__ out, int ldo, T* __restrict__ pooled, int ldp,
        int upb, int TX) {
    const int TY = 256 / TX;
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const int cg = blockIdx.y * TX + tx;
    if (cg * 4 >= C) return;
    const int n = blockIdx.z;
    const int c0 = cg * 4;
    const int HW = H * W;
    float a[4], b[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float mean, rstd;
        stat_ab(sums, n, C, c0 + j, HW, eps, mean, rstd);
        a[j] = rstd * __ldg(gamma + c0 + j);
        b[j] = __ldg(beta + c0 + j) - mean * a[j];
    }
    const T* xb = x + (long long)n * HW * ldx + c0;
    T* ob = out ? out + (long long)n * HW * ldo + c0 : nullptr;
    const T* ab = add ? add + (long long)(n % nadd) * HW * ldadd + c0 : nullptr;
    if (POOL) {
        const int Wq = W / 2, Hq = H / 2, nunits = Hq * Wq;
        T* pb = pooled + (long long)n * nunits * ldp + c0;
        const int ubeg = blockIdx.x * upb, uend = min(ubeg + upb, nunits);
        for (int u = ubeg + ty; u < uend; u += TY) {
            const int qy = u / Wq, qx = u - qy * Wq;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const long long p = (long long)(2 * qy + (i >> 1)) * W + 2 * qx + (i & 1);
                const float4 v = ld4(xb + p * ldx);
                float4 y = make_float4(fmaf(v.x, a[0], b[0]), fmaf(v.y, a[1], b[1]), fmaf(v.z, a[2], b[2]), fmaf(v.w, a[3], b[3]));
                acc.x += y.x; acc.y += y.y; acc.z += y.z; acc.w += y.w;
                if (ob) {
                    if (ab) { const float4 r = ld4(ab + p * ldadd); y.x += r.x; y.y += r.y; y.z += r.z; y.w += r.w; }
                    st4(ob + p * ldo, y);
                }
            }
            st4(pb + (long long)u * ldp, make_float4(0.25f * acc.x, 0.25f * acc.y, 0.25f * acc.z, 0.25f * acc.w));
        }
    } else {
        const int ubeg = blockIdx.x * upb, uend = min(ubeg + upb, HW);
        for (int p = ubeg + ty; p < uend; p += TY) {
            const float4 v = ld4(xb + (long long)p * ldx);
            float4 y = make_float4(fmaf(v.x, a[0], b[0]), fmaf(v.y, a[1], b[1]), fmaf(v.z, a[2], b[2]), fmaf(v.w, a[3], b[3]));
            if (ab) { const float4 r = ld4(ab + (long long)p * ldadd); y.x += r.x; y.y += r.y; y.z += r.z; y.w += r.w; }
            st4(ob + (long long)p * ldo, y);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// backward.  dy = dyA + 0.25 * upsample2(dyP)
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float4 load_dy(const T* dyA, int ldA, const T* dyP, int ldP, int W, int p) {
    float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
    if (dyA) d = ld4(dyA + (long long)p * ldA);
    if (dyP) {
        const int y = p / W, xx = p - y * W;
        const long long u = (long long)(y >> 1) * (W >> 1) + (xx >> 1);
        const float4 e = ld4(dyP + u * ldP);
        d.x = fmaf(0.25f, e.x, d.x); d.y = fmaf(0.25f, e.y, d.y); d.z = fmaf(0.25f, e.z, d.z); d.w = fmaf(0.25f, e.w, d.w);
    }
    return d;
}

template <typename T>
__global__ void __launch_bounds__(256) inorm_bwd_stats_kernel(const T* __restrict__ x, int H, int W, int C, int ldx,
        const double* __restrict__ sums, float eps, const T* __restrict__ dyA, int ldA, const T* __restrict__ dyP, int ldP,
        double* __restrict__ bsums, int ppb, int TX) {
    __shared__ float red[256][9];
    const int TY = 256 / TX;
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const int cg = blockIdx.y * TX + tx;
    const int n = blockIdx.z;
    const int HW = H * W;
    const int pbeg = blockIdx.x * ppb, pend = min(pbeg + ppb, HW);
    float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
    if (cg * 4 < C) {
        const int c0 = cg * 4;
        float mean[4], rstd[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) stat_ab(sums, n, C, c0 + j, HW, eps, mean[j], rstd[j]);
        const T* xb = x + (long long)n * HW * ldx + c0;
        const T* da = dyA ? dyA + (long long)n * HW * ldA + c0 : nullptr;
        const T* dp = dyP ? dyP + (long long)n * (HW / 4) * ldP + c0 : nullptr;
        for (int p = pbeg + ty; p < pend; p += TY) {
            const float4 v = ld4(xb + (long long)p * ldx);
            const float4 d = load_dy(da, ldA, dp, ldP, W, p);
            s[0] += d.x; s[1] += d.y; s[2] += d.z; s[3] += d.w;
            q[0] = fmaf(d.x, (v.x - mean[0]) * rstd[0], q[0]);
            q[1] = fmaf(d.y, (v.y - mean[1]) * rstd[1], q[1]);
            q[2] = fmaf(d.z, (v.z - mean[2]) * rstd[2], q[2]);
            q[3] = fmaf(d.w, (v.w - mean[3]) * rstd[3], q[3]);
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { red[threadIdx.x][j] = s[j]; red[threadIdx.x][4 + j] = q[j]; }
    __syncthreads();
    for (int e = threadIdx.x; e < TX * 8; e += 256) {      // TX * 8 partial columns (up to 512) over 256 threads
        const int t = e / 8, j = e % 8;
        const int cgo = blockIdx.y * TX + t;
        if (cgo * 4 < C) {
            double acc = 0.0;
            for (int y = 0; y < TY; ++y) acc += (double)red[y * TX + t][j];
            const int c = cgo * 4 + (j & 3);
            atomicAdd(&bsums[((long long)n * C + c) * 2 + (j >> 2)], acc);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) inorm_bwd_apply_kernel(const T* __restrict__ x, int H, int W, int C, int ldx,
        const double* __restrict__ sums, const float* __restrict__ gamma, float eps,
        const T* __restrict__ dyA, int ldA, const T* __restrict__ dyP, int ldP, const double* __restrict__ bsums, int act,
        T* __restrict__ dx, int lddx, int ppb, int TX) {
    const int TY = 256 / TX;
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const int cg = blockIdx.y * TX + tx;
    if (cg * 4 >= C) return;
    const int n = blockIdx.z;
    const int HW = H * W;
    const int c0 = cg * 4;
    float mean[4], rstd[4], a[4], m1[4], m2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        stat_ab(sums, n, C, c0 + j, HW, eps, mean[j], rstd[j]);
        a[j] = rstd[j] * __ldg(gamma + c0 + j);
        m1[j] = (float)(bsums[((long long)n * C + c0 + j) * 2 + 0] / HW);
        m2[j] = (float)(bsums[((long long)n * C + c0 + j) * 2 + 1] / HW);
    }
    const T* xb = x + (long long)n * HW * ldx + c0;
    const T* da = dyA ? dyA + (long long)n * HW * ldA + c0 : nullptr;
    const T* dp = dyP ? dyP + (long long)n * (HW / 4) * ldP + c0 : nullptr;
    T* ob = dx + (long long)n * HW * lddx + c0;
    const int pbeg = blockIdx.x * ppb, pend = min(pbeg + ppb, HW);
    for (int p = pbeg + ty; p < pend; p += TY) {
        const float4 v = ld4(xb + (long long)p * ldx);
        const float4 d = load_dy(da, ldA, dp, ldP, W, p);
        float4 r;
        r.x = act_grad_from_post(v.x, act) * a[0] * (d.x - m1[0] - (v.x - mean[0]) * rstd[0] * m2[0]);
        r.y = act_grad_from_post(v.y, act) * a[1] * (d.y - m1[1] - (v.y - mean[1]) * rstd[1] * m2[1]);
        r.z = act_grad_from_post(v.z, act) * a[2] * (d.z - m1[2] - (v.z - mean[2]) * rstd[2] * m2[2]);
        r.w = act_grad_from_post(v.w, act) * a[3] * (d.w - m1[3] - (v.w - mean[3]) * rstd[3] * m2[3]);
        st4(ob + (long long)p * lddx, r);
    }
}

// ---------------------------------------------------------------------------------------------
// bf16 fast paths: 8 channels (one 16-byte access) per thread, 4 pixels in flight per thread.
// The generic kernels above move 8 bytes per access with one load in flight and reach ~35 % of the HBM roofline
// (profiles/r01_launches_halo.txt); these serve every instance-norm tensor of the bf16 training step
// (C in {64..1024}: C/8 is a power of two <= 256, 16-byte aligned channel slices).
// ---------------------------------------------------------------------------------------------
struct f8 { float v[8]; };
__device__ __forceinline__ f8 ld8(const bf16* p) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    f8 r;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
        r.v[2 * i] = f.x; r.v[2 * i + 1] = f.y;
    }
    return r;
}
__device__ __forceinline__ void st8(bf16* p, const f8& a) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(a.v[2 * i], a.v[2 * i + 1]);
        w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ f8 load_dy8(const bf16* dyA, int ldA, const bf16* dyP, int ldP, int W, int p) {
    f8 d;
    if (dyA) d = ld8(dyA + (long long)p * ldA);
    else {
#pragma unroll
        for (int j = 0; j < 8; ++j) d.v[j] = 0.f;
    }
    if (dyP) {
        const int y = p / W, xx = p - y * W;
        const f8 e = ld8(dyP + ((long long)(y >> 1) * (W >> 1) + (xx >> 1)) * ldP);
#pragma unroll
        for (int j = 0; j < 8; ++j) d.v[j] = fmaf(0.25f, e.v[j], d.v[j]);
    }
    return d;
}

#include "norm_pipe.cuh"

// block reduction of NV partial columns per thread over the TY pixel lanes; calls sink(column c8*8 + j-th value index, total)
template <int NV, typename Sink>
__device__ __forceinline__ void reduce_lanes(const float* vals, int TX, float (*red)[NV + 1], Sink sink) {
    const int TY = 256 / TX;
#pragma unroll
    for (int j = 0; j < NV; ++j) red[threadIdx.x][j] = vals[j];
    __syncthreads();
    for (int e = threadIdx.x; e < TX * NV; e += 256) {
        const int t = e / NV, j = e % NV;
        float acc = 0.f;
        for (int y = 0; y < TY; ++y) acc += red[y * TX + t][j];
        sink(t, j, acc);
    }
}

__global__ void __launch_bounds__(256) inorm_stats8_kernel(const bf16* __restrict__ x, int HW, int C, int ldx,
                                                           double* __restrict__ sums, int ppb, int TX) {
    __shared__ float red[256][17];
    const int TY = 256 / TX;
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const int n = blockIdx.z;
    const int pbeg = blockIdx.x * ppb, pend = min(pbeg + ppb, HW);
    float a[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = 0.f;
    const bf16* xb = x + (long long)n * HW * ldx + tx * 8;
    int p = pbeg + ty;
    for (; p + 3 * TY < pend; p += 4 * TY) {
        f8 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = ld8(xb + (long long)(p + u * TY) * ldx);
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int j = 0; j < 8; ++j) { a[j] += v[u].v[j]; a[8 + j] = fmaf(v[u].v[j], v[u].v[j], a[8 + j]); }
    }
    for (; p < pend; p += TY) {
        const f8 v = ld8(xb + (long long)p * ldx);
#pragma unroll
        for (int j = 0; j < 8; ++j) { a[j] += v.v[j]; a[8 + j] = fmaf(v.v[j], v.v[j], a[8 + j]); }
    }
    reduce_lanes<16>(a, TX, red, [&](int t, int j, float acc) {
        atomicAdd(&sums[((long long)n * C + t * 8 + (j & 7)) * 2 + (j >> 3)], (double)acc);
    });
}

template <bool POOL>
__global__ void __launch_bounds__(256) inorm_apply8_kernel(const bf16* __restrict__ x, int H, int W, int C, int ldx,
        const double* __restrict__ sums, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
        const bf16* __restrict__ add, int ldadd, int nadd, bf16* __restrict__ out, int ldo, bf16* __restrict__ pooled, int ldp,
        int upb, int TX) {
    const int TY = 256 / TX;
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const int n = blockIdx.z;
    const int c0 = tx * 8;
    const int HW = H * W;
    const double inv_hw = 1.0 / (double)HW;
    float a[8], b[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float mean, rstd;
        stat_ab_fast(sums, n, C, c0 + j, inv_hw, eps, mean, rstd);
        a[j] = rstd * __ldg(gamma + c0 + j);
        b[j] = __ldg(beta + c0 + j) - mean * a[j];
    }
    const bf16* xb = x + (long long)n * HW * ldx + c0;
    bf16* ob = out ? out + (long long)n * HW * ldo + c0 : nullptr;
    const bf16* ab = add ? add + (long long)(n % nadd) * HW * ldadd + c0 : nullptr;
    if (POOL) {
        const int Wq = W / 2, nunits = (H / 2) * Wq;
        bf16* pb = pooled + (long long)n * nunits * ldp + c0;
        const int ubeg = blockIdx.x * upb, uend = min(ubeg + upb, nunits);
        for (int u = ubeg + ty; u < uend; u += TY) {
            const int qy = u / Wq, qx = u - qy * Wq;
            long long pp[4];
            f8 v[4], r[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                pp[i] = (long long)(2 * qy + (i >> 1)) * W + 2 * qx + (i & 1);
                v[i] = ld8(xb + pp[i] * ldx);
                if (ab && ob) r[i] = ld8(ab + pp[i] * ldadd);
            }
            f8 acc;
#pragma unroll
            for (int j = 0; j < 8; ++j) acc.v[j] = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                f8 y;
#pragma unroll
                for (int j = 0; j < 8; ++j) { y.v[j] = fmaf(v[i].v[j], a[j], b[j]); acc.v[j] += y.v[j]; }
                if (ob) {
                    if (ab) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) y.v[j] += r[i].v[j];
                    }
                    st8(ob + pp[i] * ldo, y);
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) acc.v[j] *= 0.25f;
            st8(pb + (long long)u * ldp, acc);
        }
    } else {
        const int ubeg = blockIdx.x * upb, uend = min(ubeg + upb, HW);
        int p = ubeg + ty;
        for (; p + 3 * TY < uend; p += 4 * TY) {
            f8 v[4], r[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                v[u] = ld8(xb + (long long)(p + u * TY) * ldx);
                if (ab) r[u] = ld8(ab + (long long)(p + u * TY) * ldadd);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                f8 y;
#pragma unroll
                for (int j = 0; j < 8; ++j) { y.v[j] = fmaf(v[u].v[j], a[j], b[j]); if (ab) y.v[j] += r[u].v[j]; }
                st8(ob + (long long)(p + u * TY) * ldo, y);
            }
        }
        for (; p < uend; p += TY) {
            const f8 v = ld8(xb + (long long)p * ldx);
            f8 y;
#pragma unroll
            for (int j = 0; j < 8; ++j) y.v[j] = fmaf(v.v[j], a[j], b[j]);
            if (ab) {
                const f8 r = ld8(ab + (long long)p * ldadd);
#pragma unroll
                for (int j = 0; j < 8; ++j) y.v[j] += r.v[j];
            }
            st8(ob + (long long)p * ldo, y);
        }
    }
}

__global__ void __launch_bounds__(256) inorm_bwd_stats8_kernel(const bf16* __restrict__ x, int H, int W, int C, int ldx,
        const double* __restrict__ sums, float eps, const bf16* __restrict__ dyA, int ldA, const bf16* __restrict__ dyP, int ldP,
        double* __restrict__ bsums, int ppb, int TX) {
    __shared__ float red[256][17];
    const int TY = 256 / TX;
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const int n = blockIdx.z;
    const int HW = H * W;
    const int c0 = tx * 8;
    const int pbeg = blockIdx.x * ppb, pend = min(pbeg + ppb, HW);
    float mean[8], rstd[8], a[16];
    const double inv_hw = 1.0 / (double)HW;
#pragma unroll
    for (int j = 0; j < 8; ++j) { stat_ab_fast(sums, n, C, c0 + j, inv_hw, eps, mean[j], rstd[j]); a[j] = 0.f; a[8 + j] = 0.f; }
    const bf16* xb = x + (long long)n * HW * ldx + c0;
    const bf16* da = dyA ? dyA + (long long)n * HW * ldA + c0 : nullptr;
    const bf16* dp = dyP ? dyP + (long long)n * (HW / 4) * ldP + c0 : nullptr;
    int p = pbeg + ty;
    for (; p + 3 * TY < pend; p += 4 * TY) {
        f8 v[4], d[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { v[u] = ld8(xb + (long long)(p + u * TY) * ldx); d[u] = load_dy8(da, ldA, dp, ldP, W, p + u * TY); }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int j = 0; j < 8; ++j) { a[j] += d[u].v[j]; a[8 + j] = fmaf(d[u].v[j], v[u].v[j] - mean[j], a[8 + j]); }
    }
    for (; p < pend; p += TY) {
        const f8 v = ld8(xb + (long long)p * ldx);
        const f8 d = load_dy8(da, ldA, dp, ldP, W, p);
#pragma unroll
        for (int j = 0; j < 8; ++j) { a[j] += d.v[j]; a[8 + j] = fmaf(d.v[j], v.v[j] - mean[j], a[8 + j]); }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) a[8 + j] *= rstd[j];          // sum dy * xhat
    reduce_lanes<16>(a, TX, red, [&](int t, int j, float acc) {
        atomicAdd(&bsums[((long long)n * C + t * 8 + (j & 7)) * 2 + (j >> 3)], (double)acc);
    });
}

// dx = act'(x) * rstd*gamma*(dy - mean(dy) - xhat*mean(dy*xhat)); dbias[c] += sum over pixels of dx (the producing conv's bias
// gradient, ShmGANwithSSpecSeg.py:244: the conv bias sits right behind this tensor) when dbias != NULL
__global__ void __launch_bounds__(256) inorm_bwd_apply8_kernel(const bf16* __restrict__ x, int H, int W, int C, int ldx,
        const double* __restrict__ sums, const float* __restrict__ gamma, float eps,
        const bf16* __restrict__ dyA, int ldA, const bf16* __restrict__ dyP, int ldP, const double* __restrict__ bsums, int act,
        bf16* __restrict__ dx, int lddx, float* __restrict__ dbias, int ppb, int TX) {
    __shared__ float red[256][9];
    const int TY = 256 / TX;
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const int n = blockIdx.z;
    const int HW = H * W;
    const int c0 = tx * 8;
    // r = g * (k0 + k1 * d + k2 * v) with g = act'(v):  a*(d - m1 - (v-mean)*rstd*m2)
    float k0[8], k1[8], k2[8], s[8];
    const double inv_hw = 1.0 / (double)HW;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float mean, rstd;
        stat_ab_fast(sums, n, C, c0 + j, inv_hw, eps, mean, rstd);
        const float a = rstd * __ldg(gamma + c0 + j);
        const double2 bs = __ldg(reinterpret_cast<const double2*>(bsums + ((long long)n * C + c0 + j) * 2));
        const float m1 = (float)(bs.x * inv_hw);
        const float m2 = (float)(bs.y * inv_hw);
        k1[j] = a; k2[j] = -a * rstd * m2; k0[j] = -a * m1 + a * rstd * m2 * mean;
        s[j] = 0.f;
    }
    const float neg = act == SHM_ACT_LRELU ? 0.2f : (act == SHM_ACT_RELU ? 0.f : 1.f);
    const bf16* xb = x + (long long)n * HW * ldx + c0;
    const bf16* da = dyA ? dyA + (long long)n * HW * ldA + c0 : nullptr;
    const bf16* dp = dyP ? dyP + (long long)n * (HW / 4) * ldP + c0 : nullptr;
    bf16* ob = dx + (long long)n * HW * lddx + c0;
    const int pbeg = blockIdx.x * ppb, pend = min(pbeg + ppb, HW);
    int p = pbeg + ty;
    for (; p + 3 * TY < pend; p += 4 * TY) {
        f8 v[4], d[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { v[u] = ld8(xb + (long long)(p + u * TY) * ldx); d[u] = load_dy8(da, ldA, dp, ldP, W, p + u * TY); }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            f8 r;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float g = v[u].v[j] > 0.f ? 1.f : neg;
                r.v[j] = g * fmaf(k2[j], v[u].v[j], fmaf(k1[j], d[u].v[j], k0[j]));
                s[j] += r.v[j];
            }
            st8(ob + (long long)(p + u * TY) * lddx, r);
        }
    }
    for (; p < pend; p += TY) {
        const f8 v = ld8(xb + (long long)p * ldx);
        const f8 d = load_dy8(da, ldA, dp, ldP, W, p);
        f8 r;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float g = v.v[j] > 0.f ? 1.f : neg;
            r.v[j] = g * fmaf(k2[j], v.v[j], fmaf(k1[j], d.v[j], k0[j]));
            s[j] += r.v[j];
        }
        st8(ob + (long long)p * lddx, r);
    }
    if (dbias != nullptr)
        reduce_lanes<8>(s, TX, red, [&](int t, int j, float acc) { atomicAdd(dbias + t * 8 + j, acc); });
}

// dpre = dy * act'(y) with the bias gradient fused: dbias[c] += sum over pixels of dpre
__global__ void __launch_bounds__(256) act_bwd8_kernel(const bf16* __restrict__ dy, int lddy, const bf16* __restrict__ y, int ldy,
        bf16* __restrict__ dpre, int ldd, long long npix, int act, float* __restrict__ dbias, int ppb, int TX) {
    __shared__ float red[256][9];
    const int TY = 256 / TX;
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const int c0 = tx * 8;
    const float neg = act == SHM_ACT_LRELU ? 0.2f : (act == SHM_ACT_RELU ? 0.f : 1.f);
    float s[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = 0.f;
    const long long pbeg = (long long)blockIdx.x * ppb;
    long long pend = pbeg + ppb; if (pend > npix) pend = npix;
    long long p = pbeg + ty;
    for (; p + 3 * TY < pend; p += 4 * TY) {
        f8 v[4], d[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { v[u] = ld8(y + (p + u * TY) * ldy + c0); d[u] = ld8(dy + (p + u * TY) * lddy + c0); }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            f8 r;
#pragma unroll
            for (int j = 0; j < 8; ++j) { r.v[j] = d[u].v[j] * (v[u].v[j] > 0.f ? 1.f : neg); s[j] += r.v[j]; }
            st8(dpre + (p + u * TY) * ldd + c0, r);
        }
    }
    for (; p < pend; p += TY) {
        const f8 v = ld8(y + p * ldy + c0), d = ld8(dy + p * lddy + c0);
        f8 r;
#pragma unroll
        for (int j = 0; j < 8; ++j) { r.v[j] = d.v[j] * (v.v[j] > 0.f ? 1.f : neg); s[j] += r.v[j]; }
        st8(dpre + p * ldd + c0, r);
    }
    if (dbias != nullptr)
        reduce_lanes<8>(s, TX, red, [&](int t, int j, float acc) { atomicAdd(dbias + t * 8 + j, acc); });
}

// the fast paths serve bf16 tensors whose channel count is 8 * 2^k <= 2048 with 16-byte aligned pixels
inline bool fast8_ok(int dtype, int C) {
    if (dtype != SHM_BF16 || C % 8 != 0) return false;
    const int tx = C / 8;
    return tx <= 256 && (tx & (tx - 1)) == 0;
}
inline bool al16(const void* p, int ld) { return p == nullptr || (((reinterpret_cast<uintptr_t>(p) & 15) == 0) && ld % 8 == 0); }
inline int ppb8(long long units, int TY, int N) {
    long long want = (long long)shm_num_sms() * 8 / (N > 0 ? N : 1);
    if (want < 1) want = 1;
    long long upb = cdiv64(units, want);
    const long long lo = (long long)TY * 32, hi = (long long)TY * 128;
    if (upb < lo) upb = lo;
    if (upb > hi) upb = hi;
    return (int)upb;
}

// ---------------------------------------------------------------------------------------------
// BatchNorm(eval) + optional MaxPool2 (SpecSeg.py:37-38)
// ---------------------------------------------------------------------------------------------
template <typename T, bool POOL>
__global__ void __launch_bounds__(256) bn_eval_kernel(const T* __restrict__ x, int H, int W, int C, int ldx,
        const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ mean,
        const float* __restrict__ var, float eps, T* __restrict__ out, int ldo, T* __restrict__ pooled, int ldp, int upb, int TX) {
    const int TY = 256 / TX;
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const int cg = blockIdx.y * TX + tx;
    if (cg * 4 >= C) return;
    const int n = blockIdx.z;
    const int c0 = cg * 4, HW = H * W;
    float a[4], b[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        a[j] = __ldg(gamma + c0 + j) * (1.0f / sqrtf(__ldg(var + c0 + j) + eps));
        b[j] = __ldg(beta + c0 + j) - __ldg(mean + c0 + j) * a[j];
    }
    const T* xb = x + (long long)n * HW * ldx + c0;
    T* ob = out + (long long)n * HW * ldo + c0;
    if (POOL) {
        const int Wq = W / 2, nunits = (H / 2) * Wq;
        T* pb = pooled + (long long)n * nunits * ldp + c0;
        const int ubeg = blockIdx.x * upb, uend = min(ubeg + upb, nunits);
        for (int u = ubeg + ty; u < uend; u += TY) {
            const int qy = u / Wq, qx = u - qy * Wq;
            float4 mx = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const long long p = (long long)(2 * qy + (i >> 1)) * W + 2 * qx + (i & 1);
                const float4 v = ld4(xb + p * ldx);
                const float4 y = make_float4(fmaf(v.x, a[0], b[0]), fmaf(v.y, a[1], b[1]), fmaf(v.z, a[2], b[2]), fmaf(v.w, a[3], b[3]));
                st4(ob + p * ldo, y);
                mx.x = fmaxf(mx.x, y.x); mx.y = fmaxf(mx.y, y.y); mx.z = fmaxf(mx.z, y.z); mx.w = fmaxf(mx.w, y.w);
            }
            // the pooled value must be the max of the STORED (rounded) values so that bf16 matches a separate pool pass
            // (rounding is monotone, so max-then-round == round-then-max)
            st4(pb + (long long)u * ldp, mx);
        }
    } else {
        const int ubeg = blockIdx.x * upb, uend = min(ubeg + upb, HW);
        for (int p = ubeg + ty; p < uend; p += TY) {
            const float4 v = ld4(xb + (long long)p * ldx);
            st4(ob + (long long)p * ldo, make_float4(fmaf(v.x, a[0], b[0]), fmaf(v.y, a[1], b[1]), fmaf(v.z, a[2], b[2]), fmaf(v.w, a[3], b[3])));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// generic strided pointwise kernels (any C; scalar per element, still coalesced over the channel axis)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void act_bwd_kernel(const T* __restrict__ dy, int lddy, const T* __restrict__ y, int ldy, T* __restrict__ dpre, int ldd,
                               long long npix, int C, int act) {
    const long long total = npix * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / C; const int c = (int)(i - p * C);
        stf(dpre + p * ldd + c, ldf(dy + p * lddy + c) * act_grad_from_post(ldf(y + p * ldy + c), act));
    }
}
template <typename T>
__global__ void act_bwd_vec_kernel(const T* __restrict__ dy, int lddy, const T* __restrict__ y, int ldy, T* __restrict__ dpre, int ldd,
                                   long long npix, int C4, int act) {
    const long long total = npix * C4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / C4; const int c = (int)(i - p * C4) * 4;
        const float4 d = ld4(dy + p * lddy + c), v = ld4(y + p * ldy + c);
        st4(dpre + p * ldd + c, make_float4(d.x * act_grad_from_post(v.x, act), d.y * act_grad_from_post(v.y, act),
                                            d.z * act_grad_from_post(v.z, act), d.w * act_grad_from_post(v.w, act)));
    }
}

template <typename T>
__global__ void add_kernel(const T* __restrict__ a, int lda, const T* __restrict__ b, int ldb, T* __restrict__ out, int ldo,
                           long long npix, int C) {
    const long long total = npix * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / C; const int c = (int)(i - p * C);
        stf(out + p * ldo + c, ldf(a + p * lda + c) + ldf(b + p * ldb + c));
    }
}

// dst[b][pix][c] (+)= sum_r src[r*nb + b][pix][c]   (gradient of a batch-broadcast add)
template <typename T>
__global__ void group_sum_kernel(const T* __restrict__ src, int lds, int reps, long long pix_per_group, int C4,
                                 T* __restrict__ dst, int ldd, int accumulate) {
    const long long total = pix_per_group * C4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / C4; const int c = (int)(i - p * C4) * 4;
        float4 acc = accumulate ? ld4(dst + p * ldd + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < reps; ++r) {
            const float4 v = ld4(src + ((long long)r * pix_per_group + p) * lds + c);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        st4(dst + p * ldd + c, acc);
    }
}

// bf16 fast path: 8 channels (16 bytes) per thread, four group images in flight
__global__ void __launch_bounds__(256) group_sum8_kernel(const bf16* __restrict__ src, int lds, int reps, long long pix_per_group, int C8,
                                                         bf16* __restrict__ dst, int ldd, int accumulate) {
    const long long total = pix_per_group * C8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / C8; const int c = (int)(i - p * C8) * 8;
        f8 acc;
        if (accumulate) acc = ld8(dst + p * ldd + c);
        else {
#pragma unroll
            for (int j = 0; j < 8; ++j) acc.v[j] = 0.f;
        }
        const bf16* s0 = src + p * lds + c;
        const long long gstep = pix_per_group * lds;
        int r = 0;
        for (; r + 4 <= reps; r += 4) {
            f8 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = ld8(s0 + (long long)(r + u) * gstep);
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc.v[j] += v[u].v[j];
        }
        for (; r < reps; ++r) {
            const f8 v = ld8(s0 + (long long)r * gstep);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc.v[j] += v.v[j];
        }
        st8(dst + p * ldd + c, acc);
    }
}

template <typename T>
__global__ void maxpool_kernel(const T* __restrict__ x, int H, int W, int C, int ldx, int k, T* __restrict__ y, int ldy, long long total) {
    const int Ho = H / k, Wo = W / k;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        long long t = i / C;
        const int ox = (int)(t % Wo); t /= Wo;
        const int oy = (int)(t % Ho);
        const int n = (int)(t / Ho);
        float m = -INFINITY;
        for (int dy = 0; dy < k; ++dy)
            for (int dx = 0; dx < k; ++dx)
                m = fmaxf(m, ldf(x + ((long long)(n * H + oy * k + dy) * W + ox * k + dx) * ldx + c));
        stf(y + ((long long)(n * Ho + oy) * Wo + ox) * ldy + c, m);
    }
}

template <typename T>
__global__ void mul_mask_kernel(const T* __restrict__ x, const T* __restrict__ keep, T* __restrict__ out, long long n, float scale) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        stf(out + i, ldf(x + i) * ldf(keep + i) * scale);
}

template <typename S, typename D>
__global__ void cast_kernel(const S* __restrict__ s, D* __restrict__ d, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        stf(d + i, ldf(s + i));
}
template <typename S, typename D>
__global__ void cast_vec_kernel(const S* __restrict__ s, D* __restrict__ d, long long n4) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
        st4(d + i * 4, ld4(s + i * 4));
}

// strided copy / convert: dst[p*ldd + c] = src[p*lds + c]
template <typename S, typename D>
__global__ void cast2d_kernel(const S* __restrict__ s, int lds, D* __restrict__ d, int ldd, long long npix, int C) {
    const long long total = npix * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / C; const int c = (int)(i - p * C);
        stf(d + p * ldd + c, ldf(s + p * lds + c));
    }
}

template <typename T>
__global__ void axpy_kernel(float alpha, const T* __restrict__ x, T* __restrict__ y, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        stf(y + i, fmaf(alpha, ldf(x + i), ldf(y + i)));
}

inline int flat_grid(long long total, int block = 256) {
    long long g = cdiv64(total, block);
    const long long cap = (long long)shm_num_sms() * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

inline bool vec_ok(const void* p, int ld, int esize) {
    return p == nullptr || (((reinterpret_cast<uintptr_t>(p) & (uintptr_t)(esize * 4 - 1)) == 0) && (ld % 4 == 0));
}

// pick units-per-block so that the grid has a few waves of 148 SMs but each thread still loops a little
inline int units_per_block(long long units, int TY, int other_blocks) {
    long long want = (long long)shm_num_sms() * 8 / (other_blocks > 0 ? other_blocks : 1);
    if (want < 1) want = 1;
    long long upb = cdiv64(units, want);
    const long long lo = (long long)TY * 4, hi = (long long)TY * 256;   // <= 256 values per fp32 partial sum
    if (upb < lo) upb = lo;
    if (upb > hi) upb = hi;
    return (int)upb;
}

}  // namespace

// tuning hook of tools/bench_norm.py: pipe_off = 1 selects the register-staged kernels; depth / grid_mul = 0 keep the defaults
extern "C" int shm_norm_tune(int pipe_off, int depth, int grid_mul) {
    SHM_REQUIRE(depth == 0 || depth == 1 || depth == 2 || depth == 4 || depth == 8, "shm_norm_tune: depth must be 0, 1, 2, 4 or 8");
    g_pipe_off = pipe_off; g_pipe_depth = depth; g_grid_mul = grid_mul;
    return SHM_OK;
}

#define REQ_VEC(p, ld, T, name) SHM_REQUIRE(vec_ok(p, ld, sizeof(T)), name ": pointer/ld not aligned to 4 elements")

extern "C" int shm_inorm_stats(const void* x, int N, int HW, int C, int ldx, int dtype, double* sums, void* stream) {
    SHM_REQUIRE(x && sums && N > 0 && HW > 0 && C > 0 && ldx >= C, "shm_inorm_stats: bad args");
    SHM_REQUIRE(C % 4 == 0, "shm_inorm_stats: C=%d must be a multiple of 4", C);
    if (pipe_ok(dtype, C, HW) && al16(x, ldx)) {
        const int TX = C / 8, S__ = pipe_depth(4);
        const long long NP = (long long)N * HW, PB = NP * C * 2;
        cudaStream_t st = (cudaStream_t)stream;
#define CALL(S_) PIPE_LAUNCH(in_stats_p<S_>, 1, NP, 256 / TX, PB, 512 << 10, (const bf16*)x, HW, C, ldx, NP, sums, TX)
        PIPE_DEPTHS(S__, CALL)
#undef CALL
        SHM_CHECK_LAUNCH("in_stats_p");
        return SHM_OK;
    }
    if (fast8_ok(dtype, C) && al16(x, ldx)) {
        const int TX = C / 8, pp = ppb8(HW, 256 / TX, N);
        inorm_stats8_kernel<<<dim3(cdiv(HW, pp), 1, N), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, HW, C, ldx, sums, pp, TX);
        SHM_CHECK_LAUNCH("inorm_stats8_kernel");
        return SHM_OK;
    }
    const CG g = cgeom(C);
    const int ppb = units_per_block(HW, g.TY, g.tiles * N);
    dim3 grid(cdiv(HW, ppb), g.tiles, N);
    DISPATCH_DTYPE(dtype, T, {
        REQ_VEC(x, ldx, T, "shm_inorm_stats");
        inorm_stats_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>((const T*)x, HW, C, ldx, sums, ppb, g.TX);
        SHM_CHECK_LAUNCH("inorm_stats_kernel");
        return SHM_OK;
    })
}

extern "C" int shm_inorm_apply(const void* x, int N, int H, int W, int C, int ldx, int dtype, const double* sums,
                               const float* gamma, const float* beta, float eps, const void* add, int ldadd, int nadd,
                               void* out, int ldo, void* pooled, int ldp, void* stream) {
    SHM_REQUIRE(x && sums && gamma && beta && N > 0 && H > 0 && W > 0 && C > 0, "shm_inorm_apply: bad args");
    SHM_REQUIRE(out || pooled, "shm_inorm_apply: no output");
    SHM_REQUIRE(C % 4 == 0, "shm_inorm_apply: C=%d must be a multiple of 4", C);
    SHM_REQUIRE(!pooled || (H % 2 == 0 && W % 2 == 0), "shm_inorm_apply: pooling needs even H, W");
    SHM_REQUIRE(!add || out, "shm_inorm_apply: add without out");
    if (nadd <= 0) nadd = N;
    SHM_REQUIRE(N % nadd == 0, "shm_inorm_apply: N %% nadd != 0");
    if (pipe_ok(dtype, C, pooled ? (long long)(H / 2) * (W / 2) : (long long)H * W) && al16(x, ldx) && al16(add, ldadd) && al16(out, ldo) && al16(pooled, ldp)) {
        const int TX = C / 8;
        cudaStream_t st = (cudaStream_t)stream;
        if (pooled) {
            const long long NQ = (long long)N * (H / 2) * (W / 2);
            const int S__ = pipe_depth((out && add) ? 2 : 4) > 4 ? 4 : pipe_depth((out && add) ? 2 : 4);
            const long long PB = NQ * 4 * C * 2;
#define POOL_ARGS (const bf16*)x, H, W, C, ldx, NQ, sums, gamma, beta, eps, (const bf16*)add, ldadd, nadd, (bf16*)out, ldo, (bf16*)pooled, ldp, TX
#define CALL(S_) { if (out && add) PIPE_LAUNCH((in_apply_pool_p<S_, true, true>), 8, NQ, 256 / TX, PB, 128 << 10, POOL_ARGS) \
                   else if (out) PIPE_LAUNCH((in_apply_pool_p<S_, true, false>), 4, NQ, 256 / TX, PB, 128 << 10, POOL_ARGS) \
                   else PIPE_LAUNCH((in_apply_pool_p<S_, false, false>), 4, NQ, 256 / TX, PB, 256 << 10, POOL_ARGS) }
            PIPE_DEPTHS3(S__, CALL)
#undef CALL
#undef POOL_ARGS
        } else {
            const long long NP = (long long)N * H * W;
            const int S__ = pipe_depth(4);
            const long long PB = NP * C * 2;
#define APPLY_ARGS (const bf16*)x, H * W, C, ldx, NP, sums, gamma, beta, eps, (const bf16*)add, ldadd, nadd, (bf16*)out, ldo, TX
#define CALL(S_) { if (add) PIPE_LAUNCH((in_apply_p<S_, true>), 2, NP, 256 / TX, PB, 128 << 10, APPLY_ARGS) \
                   else PIPE_LAUNCH((in_apply_p<S_, false>), 1, NP, 256 / TX, PB, 128 << 10, APPLY_ARGS) }
            PIPE_DEPTHS(S__, CALL)
#undef CALL
#undef APPLY_ARGS
        }
        SHM_CHECK_LAUNCH("in_apply_p");
        return SHM_OK;
    }
    if (fast8_ok(dtype, C) && al16(x, ldx) && al16(add, ldadd) && al16(out, ldo) && al16(pooled, ldp)) {
        const int TX = C / 8;
        const long long un = pooled ? (long long)(H / 2) * (W / 2) : (long long)H * W;
        const int pp = ppb8(un, 256 / TX, N);
        dim3 grid8((unsigned)cdiv64(un, pp), 1, N);
        if (pooled) inorm_apply8_kernel<true><<<grid8, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, H, W, C, ldx, sums, gamma, beta, eps, (const bf16*)add, ldadd, nadd, (bf16*)out, ldo, (bf16*)pooled, ldp, pp, TX);
        else        inorm_apply8_kernel<false><<<grid8, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, H, W, C, ldx, sums, gamma, beta, eps, (const bf16*)add, ldadd, nadd, (bf16*)out, ldo, nullptr, 0, pp, TX);
        SHM_CHECK_LAUNCH("inorm_apply8_kernel");
        return SHM_OK;
    }
    const CG g = cgeom(C);
    const long long units = pooled ? (long long)(H / 2) * (W / 2) : (long long)H * W;
    const int upb = units_per_block(units, g.TY, g.tiles * N);
    dim3 grid((unsigned)cdiv64(units, upb), g.tiles, N);
    cudaStream_t st = (cudaStream_t)stream;
    DISPATCH_DTYPE(dtype, T, {
        REQ_VEC(x, ldx, T, "shm_inorm_apply"); REQ_VEC(add, ldadd, T, "shm_inorm_apply");
        REQ_VEC(out, ldo, T, "shm_inorm_apply"); REQ_VEC(pooled, ldp, T, "shm_inorm_apply");
        if (pooled) inorm_apply_kernel<T, true><<<grid, 256, 0, st>>>((const T*)x, H, W, C, ldx, sums, gamma, beta, eps, (const T*)add, ldadd, nadd, (T*)out, ldo, (T*)pooled, ldp, upb, g.TX);
        else        inorm_apply_kernel<T, false><<<grid, 256, 0, st>>>((const T*)x, H, W, C, ldx, sums, gamma, beta, eps, (const T*)add, ldadd, nadd, (T*)out, ldo, nullptr, 0, upb, g.TX);
        SHM_CHECK_LAUNCH("inorm_apply_kernel");
        return SHM_OK;
    })
}

extern "C" int shm_inorm_bwd_stats(const void* x, int N, int H, int W, int C, int ldx, int dtype, const double* sums, float eps,
                                   const void* dyA, int ldA, const void* dyP, int ldP, double* bsums, void* stream) {
    SHM_REQUIRE(x && sums && bsums && (dyA || dyP) && N > 0 && H > 0 && W > 0, "shm_inorm_bwd_stats: bad args");
    SHM_REQUIRE(C % 4 == 0, "shm_inorm_bwd_stats: C=%d must be a multiple of 4", C);
    SHM_REQUIRE(!dyP || (H % 2 == 0 && W % 2 == 0), "shm_inorm_bwd_stats: pooled gradient needs even H, W");
    if (pipe_ok(dtype, C, (long long)H * W) && al16(x, ldx) && al16(dyA, ldA) && al16(dyP, ldP)) {
        const int TX = C / 8, S__ = pipe_depth(4) > 4 ? 4 : pipe_depth(4);
        const long long NP = (long long)N * H * W, PB = NP * C * 2;
        cudaStream_t st = (cudaStream_t)stream;
#define BS_ARGS (const bf16*)x, H, W, C, ldx, NP, sums, eps, (const bf16*)dyA, ldA, (const bf16*)dyP, ldP, bsums, TX
#define CALL(S_) { if (dyA && dyP) PIPE_LAUNCH((in_bwd_stats_p<S_, 3>), 3, NP, 256 / TX, PB, 512 << 10, BS_ARGS) \
                   else if (dyA) PIPE_LAUNCH((in_bwd_stats_p<S_, 1>), 2, NP, 256 / TX, PB, 512 << 10, BS_ARGS) \
                   else PIPE_LAUNCH((in_bwd_stats_p<S_, 2>), 2, NP, 256 / TX, PB, 512 << 10, BS_ARGS) }
        PIPE_DEPTHS3(S__, CALL)
#undef CALL
#undef BS_ARGS
        SHM_CHECK_LAUNCH("in_bwd_stats_p");
        return SHM_OK;
    }
    if (fast8_ok(dtype, C) && al16(x, ldx) && al16(dyA, ldA) && al16(dyP, ldP)) {
        const int TX = C / 8, pp = ppb8((long long)H * W, 256 / TX, N);
        inorm_bwd_stats8_kernel<<<dim3(cdiv(H * W, pp), 1, N), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, H, W, C, ldx, sums, eps, (const bf16*)dyA, ldA, (const bf16*)dyP, ldP, bsums, pp, TX);
        SHM_CHECK_LAUNCH("inorm_bwd_stats8_kernel");
        return SHM_OK;
    }
    const CG g = cgeom(C);
    const int ppb = units_per_block((long long)H * W, g.TY, g.tiles * N);
    dim3 grid(cdiv(H * W, ppb), g.tiles, N);
    DISPATCH_DTYPE(dtype, T, {
        REQ_VEC(x, ldx, T, "shm_inorm_bwd_stats"); REQ_VEC(dyA, ldA, T, "shm_inorm_bwd_stats"); REQ_VEC(dyP, ldP, T, "shm_inorm_bwd_stats");
        inorm_bwd_stats_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>((const T*)x, H, W, C, ldx, sums, eps, (const T*)dyA, ldA, (const T*)dyP, ldP, bsums, ppb, g.TX);
        SHM_CHECK_LAUNCH("inorm_bwd_stats_kernel");
        return SHM_OK;
    })
}

extern "C" int shm_inorm_bwd_apply(const void* x, int N, int H, int W, int C, int ldx, int dtype, const double* sums,
                                   const float* gamma, float eps, const void* dyA, int ldA, const void* dyP, int ldP,
                                   const double* bsums, int act, void* dx, int lddx, float* dbias, void* stream) {
    SHM_REQUIRE(x && sums && gamma && bsums && dx && (dyA || dyP) && N > 0 && H > 0 && W > 0, "shm_inorm_bwd_apply: bad args");
    SHM_REQUIRE(C % 4 == 0, "shm_inorm_bwd_apply: C=%d must be a multiple of 4", C);
    SHM_REQUIRE(!dyP || (H % 2 == 0 && W % 2 == 0), "shm_inorm_bwd_apply: pooled gradient needs even H, W");
    if (pipe_ok(dtype, C, (long long)H * W) && al16(x, ldx) && al16(dyA, ldA) && al16(dyP, ldP) && al16(dx, lddx) && act != SHM_ACT_SIGMOID) {
        const int TX = C / 8, S__ = pipe_depth(4) > 4 ? 4 : pipe_depth(4);
        const long long NP = (long long)N * H * W, PB = NP * C * 2;
        cudaStream_t st = (cudaStream_t)stream;
#define BA_ARGS (const bf16*)x, H, W, C, ldx, NP, sums, gamma, eps, (const bf16*)dyA, ldA, (const bf16*)dyP, ldP, bsums, act, (bf16*)dx, lddx, dbias, TX
#define CALL(S_) { if (dyA && dyP) PIPE_LAUNCH((in_bwd_apply_p<S_, 3>), 3, NP, 256 / TX, PB, 128 << 10, BA_ARGS) \
                   else if (dyA) PIPE_LAUNCH((in_bwd_apply_p<S_, 1>), 2, NP, 256 / TX, PB, 128 << 10, BA_ARGS) \
                   else PIPE_LAUNCH((in_bwd_apply_p<S_, 2>), 2, NP, 256 / TX, PB, 128 << 10, BA_ARGS) }
        PIPE_DEPTHS3(S__, CALL)
#undef CALL
#undef BA_ARGS
        SHM_CHECK_LAUNCH("in_bwd_apply_p");
        return SHM_OK;
    }
    if (fast8_ok(dtype, C) && al16(x, ldx) && al16(dyA, ldA) && al16(dyP, ldP) && al16(dx, lddx) && act != SHM_ACT_SIGMOID) {
        const int TX = C / 8, pp = ppb8((long long)H * W, 256 / TX, N);
        inorm_bwd_apply8_kernel<<<dim3(cdiv(H * W, pp), 1, N), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, H, W, C, ldx, sums, gamma, eps, (const bf16*)dyA, ldA, (const bf16*)dyP, ldP, bsums, act, (bf16*)dx, lddx, dbias, pp, TX);
        SHM_CHECK_LAUNCH("inorm_bwd_apply8_kernel");
        return SHM_OK;
    }
    const CG g = cgeom(C);
    const int ppb = units_per_block((long long)H * W, g.TY, g.tiles * N);
    dim3 grid(cdiv(H * W, ppb), g.tiles, N);
    DISPATCH_DTYPE(dtype, T, {
        REQ_VEC(x, ldx, T, "shm_inorm_bwd_apply"); REQ_VEC(dyA, ldA, T, "shm_inorm_bwd_apply");
        REQ_VEC(dyP, ldP, T, "shm_inorm_bwd_apply"); REQ_VEC(dx, lddx, T, "shm_inorm_bwd_apply");
        inorm_bwd_apply_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>((const T*)x, H, W, C, ldx, sums, gamma, eps, (const T*)dyA, ldA, (const T*)dyP, ldP, bsums, act, (T*)dx, lddx, ppb, g.TX);
        SHM_CHECK_LAUNCH("inorm_bwd_apply_kernel");
        if (dbias) return shm_colsum(dx, (int64_t)N * H * W, C, lddx, dtype, dbias, stream);
        return SHM_OK;
    })
}

extern "C" int shm_bn_eval(const void* x, int N, int H, int W, int C, int ldx, int dtype, const float* gamma, const float* beta,
                           const float* mean, const float* var, float eps, void* out, int ldo, void* pooled, int ldp, void* stream) {
    SHM_REQUIRE(x && gamma && beta && mean && var && out && N > 0 && H > 0 && W > 0, "shm_bn_eval: bad args");
    SHM_REQUIRE(C % 4 == 0, "shm_bn_eval: C=%d must be a multiple of 4", C);
    SHM_REQUIRE(!pooled || (H % 2 == 0 && W % 2 == 0), "shm_bn_eval: pooling needs even H, W");
    const CG g = cgeom(C);
    const long long units = pooled ? (long long)(H / 2) * (W / 2) : (long long)H * W;
    const int upb = units_per_block(units, g.TY, g.tiles * N);
    dim3 grid((unsigned)cdiv64(units, upb), g.tiles, N);
    cudaStream_t st = (cudaStream_t)stream;
    DISPATCH_DTYPE(dtype, T, {
        REQ_VEC(x, ldx, T, "shm_bn_eval"); REQ_VEC(out, ldo, T, "shm_bn_eval"); REQ_VEC(pooled, ldp, T, "shm_bn_eval");
        if (pooled) bn_eval_kernel<T, true><<<grid, 256, 0, st>>>((const T*)x, H, W, C, ldx, gamma, beta, mean, var, eps, (T*)out, ldo, (T*)pooled, ldp, upb, g.TX);
        else        bn_eval_kernel<T, false><<<grid, 256, 0, st>>>((const T*)x, H, W, C, ldx, gamma, beta, mean, var, eps, (T*)out, ldo, nullptr, 0, upb, g.TX);
        SHM_CHECK_LAUNCH("bn_eval_kernel");
        return SHM_OK;
    })
}

extern "C" int shm_act_bwd(const void* dy, int lddy, const void* y, int ldy, void* dpre, int ldd, int64_t npix, int C, int act,
                           int dtype, float* dbias, void* stream) {
    SHM_REQUIRE(dy && y && dpre && npix >= 0 && C > 0, "shm_act_bwd: bad args");
    if (npix == 0) return SHM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (pipe_ok(dtype, C, npix) && al16(dy, lddy) && al16(y, ldy) && al16(dpre, ldd) && act != SHM_ACT_SIGMOID) {
        const int TX = C / 8, S__ = pipe_depth(4);
        cudaStream_t st = (cudaStream_t)stream;
#define CALL(S_) PIPE_LAUNCH(act_bwd_p<S_>, 2, npix, 256 / TX, (long long)npix * C * 2, 128 << 10, (const bf16*)dy, lddy, (const bf16*)y, ldy, (bf16*)dpre, ldd, (long long)npix, act, dbias, TX)
        PIPE_DEPTHS(S__, CALL)
#undef CALL
        SHM_CHECK_LAUNCH("act_bwd_p");
        return SHM_OK;
    }
    if (fast8_ok(dtype, C) && al16(dy, lddy) && al16(y, ldy) && al16(dpre, ldd) && act != SHM_ACT_SIGMOID) {
        const int TX = C / 8, pp = ppb8(npix, 256 / TX, 1);
        act_bwd8_kernel<<<(unsigned)cdiv64(npix, pp), 256, 0, st>>>((const bf16*)dy, lddy, (const bf16*)y, ldy, (bf16*)dpre, ldd, npix, act, dbias, pp, TX);
        SHM_CHECK_LAUNCH("act_bwd8_kernel");
        return SHM_OK;
    }
    DISPATCH_DTYPE(dtype, T, {
        if (C % 4 == 0 && vec_ok(dy, lddy, sizeof(T)) && vec_ok(y, ldy, sizeof(T)) && vec_ok(dpre, ldd, sizeof(T)))
            act_bwd_vec_kernel<T><<<flat_grid(npix * (C / 4)), 256, 0, st>>>((const T*)dy, lddy, (const T*)y, ldy, (T*)dpre, ldd, npix, C / 4, act);
        else
            act_bwd_kernel<T><<<flat_grid(npix * C), 256, 0, st>>>((const T*)dy, lddy, (const T*)y, ldy, (T*)dpre, ldd, npix, C, act);
        SHM_CHECK_LAUNCH("act_bwd_kernel");
        if (dbias) return shm_colsum(dpre, npix, C, ldd, dtype, dbias, stream);
        return SHM_OK;
    })
}

extern "C" int shm_add(const void* a, int lda, const void* b, int ldb, void* out, int ldo, int64_t npix, int C, int dtype, void* stream) {
    SHM_REQUIRE(a && b && out && npix >= 0 && C > 0, "shm_add: bad args");
    if (npix == 0) return SHM_OK;
    DISPATCH_DTYPE(dtype, T, {
        add_kernel<T><<<flat_grid(npix * C), 256, 0, (cudaStream_t)stream>>>((const T*)a, lda, (const T*)b, ldb, (T*)out, ldo, npix, C);
        SHM_CHECK_LAUNCH("add_kernel");
        return SHM_OK;
    })
}

extern "C" int shm_group_sum(const void* src, int lds, int reps, int64_t pix_per_group, int C, void* dst, int ldd, int accumulate,
                             int dtype, void* stream) {
    SHM_REQUIRE(src && dst && reps > 0 && pix_per_group > 0 && C > 0 && C % 4 == 0, "shm_group_sum: bad args");
    if (dtype == SHM_BF16 && C % 8 == 0 && al16(src, lds) && al16(dst, ldd)) {
        long long g = cdiv64(pix_per_group * (C / 8), 256);
        const long long cap = (long long)shm_num_sms() * 16;
        if (g > cap) g = cap;
        group_sum8_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>((const bf16*)src, lds, reps, pix_per_group, C / 8, (bf16*)dst, ldd, accumulate);
        SHM_CHECK_LAUNCH("group_sum8_kernel");
        return SHM_OK;
    }
    DISPATCH_DTYPE(dtype, T, {
        REQ_VEC(src, lds, T, "shm_group_sum"); REQ_VEC(dst, ldd, T, "shm_group_sum");
        group_sum_kernel<T><<<flat_grid(pix_per_group * (C / 4)), 256, 0, (cudaStream_t)stream>>>((const T*)src, lds, reps, pix_per_group, C / 4, (T*)dst, ldd, accumulate);
        SHM_CHECK_LAUNCH("group_sum_kernel");
        return SHM_OK;
    })
}

extern "C" int shm_maxpool(const void* x, int N, int H, int W, int C, int ldx, int k, void* y, int ldy, int dtype, void* stream) {
    SHM_REQUIRE(x && y && N > 0 && H > 0 && W > 0 && C > 0 && k >= 1, "shm_maxpool: bad args");
    SHM_REQUIRE(H % k == 0 && W % k == 0, "shm_maxpool: H, W must be divisible by k=%d", k);
    const long long total = (long long)N * (H / k) * (W / k) * C;
    DISPATCH_DTYPE(dtype, T, {
        maxpool_kernel<T><<<flat_grid(total), 256, 0, (cudaStream_t)stream>>>((const T*)x, H, W, C, ldx, k, (T*)y, ldy, total);
        SHM_CHECK_LAUNCH("maxpool_kernel");
        return SHM_OK;
    })
}

extern "C" int shm_mul_mask(const void* x, const void* keep, void* out, int64_t n, float scale, int dtype, void* stream) {
    SHM_REQUIRE(x && keep && out && n >= 0, "shm_mul_mask: bad args");
    if (n == 0) return SHM_OK;
    DISPATCH_DTYPE(dtype, T, {
        mul_mask_kernel<T><<<flat_grid(n), 256, 0, (cudaStream_t)stream>>>((const T*)x, (const T*)keep, (T*)out, n, scale);
        SHM_CHECK_LAUNCH("mul_mask_kernel");
        return SHM_OK;
    })
}

extern "C" int shm_cast(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n, void* stream) {
    SHM_REQUIRE(src && dst && n >= 0, "shm_cast: bad args");
    if (n == 0) return SHM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const bool v = (n % 4 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
#define CAST_CASE(S, D) { if (v) cast_vec_kernel<S, D><<<flat_grid(n / 4), 256, 0, st>>>((const S*)src, (D*)dst, n / 4); \
                          else cast_kernel<S, D><<<flat_grid(n), 256, 0, st>>>((const S*)src, (D*)dst, n); }
    if (src_dtype == SHM_F32 && dst_dtype == SHM_BF16) CAST_CASE(float, bf16)
    else if (src_dtype == SHM_BF16 && dst_dtype == SHM_F32) CAST_CASE(bf16, float)
    else if (src_dtype == SHM_F32 && dst_dtype == SHM_F32) CAST_CASE(float, float)
    else if (src_dtype == SHM_BF16 && dst_dtype == SHM_BF16) CAST_CASE(bf16, bf16)
    else SHM_FAIL(SHM_EINVAL, "shm_cast: bad dtypes %d -> %d", src_dtype, dst_dtype);
#undef CAST_CASE
    SHM_CHECK_LAUNCH("cast_kernel");
    return SHM_OK;
}

extern "C" int shm_axpy(float alpha, const void* x, void* y, int64_t n, int dtype, void* stream) {
    SHM_REQUIRE(x && y && n >= 0, "shm_axpy: bad args");
    if (n == 0) return SHM_OK;
    DISPATCH_DTYPE(dtype, T, {
        axpy_kernel<T><<<flat_grid(n), 256, 0, (cudaStream_t)stream>>>(alpha, (const T*)x, (T*)y, n);
        SHM_CHECK_LAUNCH("axpy_kernel");
        return SHM_OK;
    })
}

extern "C" int shm_cast2d(const void* src, int src_dtype, int lds, void* dst, int dst_dtype, int ldd, int64_t npix, int C, void* stream) {
    SHM_REQUIRE(src && dst && npix >= 0 && C > 0 && lds >= C && ldd >= C, "shm_cast2d: bad args");
    if (npix == 0) return SHM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = flat_grid(npix * C);
    if (src_dtype == SHM_F32 && dst_dtype == SHM_BF16) cast2d_kernel<float, bf16><<<grid, 256, 0, st>>>((const float*)src, lds, (bf16*)dst, ldd, npix, C);
    else if (src_dtype == SHM_BF16 && dst_dtype == SHM_F32) cast2d_kernel<bf16, float><<<grid, 256, 0, st>>>((const bf16*)src, lds, (float*)dst, ldd, npix, C);
    else if (src_dtype == SHM_F32 && dst_dtype == SHM_F32) cast2d_kernel<float, float><<<grid, 256, 0, st>>>((const float*)src, lds, (float*)dst, ldd, npix, C);
    else if (src_dtype == SHM_BF16 && dst_dtype == SHM_BF16) cast2d_kernel<bf16, bf16><<<grid, 256, 0, st>>>((const bf16*)src, lds, (bf16*)dst, ldd, npix, C);
    else SHM_FAIL(SHM_EINVAL, "shm_cast2d: bad dtypes %d -> %d", src_dtype, dst_dtype);
    SHM_CHECK_LAUNCH("cast2d_kernel");
    return SHM_OK;
}
