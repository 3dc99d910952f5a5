// extras.cu -- the rows SURVEY.md section 8(f) marks "next", as bandwidth kernels next to the hot path:
//   dataset-loader contract   datasetLoader.py:48-62   uint8 -> bilinear resize (TF2 half-pixel) -> /255 -> optional vertical flip
//   degree of polarisation    ShmGANwithSSpecSeg.py:1157-1169
//   test-time metrics         test.py:332-352          per-image squared error (MSE / PSNR), sRGB -> Lab + Delta-E 76 / 94 sums
#include "common.cuh"

namespace {

inline int flat_grid(long long total, int block = 256) {
    long long g = cdiv64(total, block);
    const long long cap = (long long)shm_num_sms() * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// TF 2.8 compute_interpolation_weights with HalfPixelScaler, float32 throughout (no fused multiply-add: TF rounds each step)
__device__ __forceinline__ void interp(int o, float scale, int in_size, int& lo, int& hi, float& lerp) {
    const float src = __fsub_rn(__fmul_rn(__fadd_rn((float)o, 0.5f), scale), 0.5f);
    const float f = floorf(src);
    lo = max((int)f, 0);
    hi = min((int)ceilf(src), in_size - 1);
    lerp = __fsub_rn(src, f);
}

// one thread per output pixel (C = 3 channels): 4 source pixels -> top / bottom lerps -> / 255 -> row (flipped or not)
__global__ void load_u8_bilinear_kernel(const unsigned char* __restrict__ src, int Hs, int Ws, float* __restrict__ dst, int Ho, int Wo,
                                        int flip_ud, float sy, float sx, long long total) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % Wo);
        const long long t = i / Wo;
        const int y = (int)(t % Ho);
        const long long n = t / Ho;
        int ylo, yhi, xlo, xhi;
        float yl, xl;
        interp(y, sy, Hs, ylo, yhi, yl);
        interp(x, sx, Ws, xlo, xhi, xl);
        const unsigned char* im = src + n * Hs * Ws * 3;
        const unsigned char* tl = im + ((long long)ylo * Ws + xlo) * 3;
        const unsigned char* tr = im + ((long long)ylo * Ws + xhi) * 3;
        const unsigned char* bl = im + ((long long)yhi * Ws + xlo) * 3;
        const unsigned char* br = im + ((long long)yhi * Ws + xhi) * 3;
        const int yo = flip_ud ? Ho - 1 - y : y;
        float* o = dst + ((n * Ho + yo) * Wo + x) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float a = (float)tl[c], b = (float)tr[c], d = (float)bl[c], e = (float)br[c];
            const float top = __fadd_rn(a, __fmul_rn(__fsub_rn(b, a), xl));
            const float bot = __fadd_rn(d, __fmul_rn(__fsub_rn(e, d), xl));
            o[c] = __fdiv_rn(__fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), yl)), 255.0f);
        }
    }
}

// same-size fast path: u8 -> f32 / 255 (+ flip), 16 source bytes per thread
__global__ void load_u8_scale_kernel(const uint4* __restrict__ src, float4* __restrict__ dst, int H, long long row16, int flip_ud, long long total16) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total16; i += (long long)gridDim.x * blockDim.x) {
        const uint4 u = __ldg(src + i);
        long long o = i;
        if (flip_ud) {
            const long long r = i / row16, c = i - r * row16;
            const long long n = r / H, y = r - n * H;
            o = (n * H + (H - 1 - y)) * row16 + c;
        }
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
            dst[o * 4 + k] = make_float4(__fdiv_rn((float)(w[k] & 255u), 255.0f), __fdiv_rn((float)((w[k] >> 8) & 255u), 255.0f),
                                         __fdiv_rn((float)((w[k] >> 16) & 255u), 255.0f), __fdiv_rn((float)(w[k] >> 24), 255.0f));
    }
}

__global__ void dop_kernel(const float* __restrict__ i0, const float* __restrict__ i45, const float* __restrict__ i90,
                           const float* __restrict__ i135, float* __restrict__ dopo, float* __restrict__ aop, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float a = __ldg(i0 + i), b = __ldg(i45 + i), c = __ldg(i90 + i), d = __ldg(i135 + i);
        const float s0 = a + c, s1 = a - c, s2 = b - d;
        const float pol = sqrtf(__fadd_rn(__fmul_rn(s1, s1), __fmul_rn(s2, s2)));
        dopo[i] = s0 == 0.f ? 0.f : __fdiv_rn(pol, s0);            // tf.math.divide_no_nan
        if (aop) aop[i] = 0.5f * atan2f(s2, s1);
    }
}

// per-image sum of squared differences (fp64 accumulation): out[n] += sum (a - b)^2
__global__ void __launch_bounds__(256) sqerr_kernel(const float* __restrict__ a, const float* __restrict__ b, long long per, double* __restrict__ out) {
    __shared__ double sm[32];
    const long long n = blockIdx.y;
    const float* pa = a + n * per;
    const float* pb = b + n * per;
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per; i += (long long)gridDim.x * blockDim.x) {
        const float d = __ldg(pa + i) - __ldg(pb + i);
        acc += (double)d * (double)d;
    }
    acc = block_sum(acc, sm);
    if (threadIdx.x == 0) atomicAdd(out + n, acc);
}

// sRGB -> CIE Lab (D65 / 2 degree), the algorithm of tfio.experimental.color.rgb_to_lab / skimage.color.rgb2lab
__device__ __forceinline__ void rgb2lab(float r, float g, float b, float& L, float& A, float& B) {
    auto lin = [](float v) { return v > 0.04045f ? powf((v + 0.055f) / 1.055f, 2.4f) : v / 12.92f; };
    auto f = [](float t) { return t > 0.008856f ? cbrtf(t) : 7.787f * t + 16.0f / 116.0f; };
    const float R = lin(r), G = lin(g), Bl = lin(b);
    const float fx = f((0.412453f * R + 0.357580f * G + 0.180423f * Bl) / 0.95047f);
    const float fy = f(0.212671f * R + 0.715160f * G + 0.072169f * Bl);
    const float fz = f((0.019334f * R + 0.119193f * G + 0.950227f * Bl) / 1.08883f);
    L = 116.0f * fy - 16.0f;
    A = 500.0f * (fx - fy);
    B = 200.0f * (fy - fz);
}

// sums[n][0] += sum of dE76, sums[n][1] += sum of dE94 (skimage.color.deltaE_cie76 / deltaE_ciede94 defaults; image 1 is the reference colour)
__global__ void __launch_bounds__(256) delta_e_kernel(const float* __restrict__ rgb1, const float* __restrict__ rgb2, long long HW, double* __restrict__ sums) {
    __shared__ double sm[32];
    const long long n = blockIdx.y;
    const float* p1 = rgb1 + n * HW * 3;
    const float* p2 = rgb2 + n * HW * 3;
    double s76 = 0.0, s94 = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (long long)gridDim.x * blockDim.x) {
        float l1, a1, b1, l2, a2, b2;
        rgb2lab(__ldg(p1 + 3 * i), __ldg(p1 + 3 * i + 1), __ldg(p1 + 3 * i + 2), l1, a1, b1);
        rgb2lab(__ldg(p2 + 3 * i), __ldg(p2 + 3 * i + 1), __ldg(p2 + 3 * i + 2), l2, a2, b2);
        const float dL = l1 - l2, da = a1 - a2, db = b1 - b2;
        s76 += (double)sqrtf(dL * dL + da * da + db * db);
        const float c1 = hypotf(a1, b1), c2 = hypotf(a2, b2);
        const float dC = c1 - c2;
        const float dH2 = 2.0f * (c1 * c2 - (a1 * a2 + b1 * b2));
        const float sc = 1.0f + 0.045f * c1, sh = 1.0f + 0.015f * c1;
        const float de2 = dL * dL + (dC / sc) * (dC / sc) + dH2 / (sh * sh);
        s94 += (double)sqrtf(fmaxf(de2, 0.f));
    }
    s76 = block_sum(s76, sm);
    s94 = block_sum(s94, sm);
    if (threadIdx.x == 0) { atomicAdd(sums + 2 * n, s76); atomicAdd(sums + 2 * n + 1, s94); }
}

}  // namespace

extern "C" int shm_load_u8_bilinear(const void* src, int N, int Hs, int Ws, float* dst, int Ho, int Wo, int flip_ud, void* stream) {
    SHM_REQUIRE(src && dst && N >= 0 && Hs > 0 && Ws > 0 && Ho > 0 && Wo > 0, "shm_load_u8_bilinear: bad args");
    if (N == 0) return SHM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const long long row = (long long)Ws * 3;
    if (Hs == Ho && Ws == Wo && row % 16 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        const long long total16 = (long long)N * Hs * row / 16;
        load_u8_scale_kernel<<<flat_grid(total16), 256, 0, st>>>((const uint4*)src, (float4*)dst, Hs, row / 16, flip_ud, total16);
        SHM_CHECK_LAUNCH("load_u8_scale_kernel");
        return SHM_OK;
    }
    const long long total = (long long)N * Ho * Wo;
    load_u8_bilinear_kernel<<<flat_grid(total), 256, 0, st>>>((const unsigned char*)src, Hs, Ws, dst, Ho, Wo, flip_ud,
                                                              (float)Hs / (float)Ho, (float)Ws / (float)Wo, total);
    SHM_CHECK_LAUNCH("load_u8_bilinear_kernel");
    return SHM_OK;
}

extern "C" int shm_dop(const float* i0, const float* i45, const float* i90, const float* i135, float* dop, float* aop, int64_t n, void* stream) {
    SHM_REQUIRE(i0 && i45 && i90 && i135 && dop && n >= 0, "shm_dop: bad args");
    if (n == 0) return SHM_OK;
    dop_kernel<<<flat_grid(n), 256, 0, (cudaStream_t)stream>>>(i0, i45, i90, i135, dop, aop, n);
    SHM_CHECK_LAUNCH("dop_kernel");
    return SHM_OK;
}

extern "C" int shm_sqerr_per_image(const float* a, const float* b, int N, int64_t per, double* out, void* stream) {
    SHM_REQUIRE(a && b && out && N >= 0 && per >= 0, "shm_sqerr_per_image: bad args");
    if (N == 0 || per == 0) return SHM_OK;
    int gx = flat_grid(per) / (N > 0 ? N : 1);
    if (gx < 1) gx = 1;
    sqerr_kernel<<<dim3(gx, N), 256, 0, (cudaStream_t)stream>>>(a, b, per, out);
    SHM_CHECK_LAUNCH("sqerr_kernel");
    return SHM_OK;
}

extern "C" int shm_delta_e(const float* rgb1, const float* rgb2, int N, int64_t HW, double* sums, void* stream) {
    SHM_REQUIRE(rgb1 && rgb2 && sums && N >= 0 && HW >= 0, "shm_delta_e: bad args");
    if (N == 0 || HW == 0) return SHM_OK;
    int gx = flat_grid(HW) / (N > 0 ? N : 1);
    if (gx < 1) gx = 1;
    delta_e_kernel<<<dim3(gx, N), 256, 0, (cudaStream_t)stream>>>(rgb1, rgb2, HW, sums);
    SHM_CHECK_LAUNCH("delta_e_kernel");
    return SHM_OK;
}
