// prep.cu -- polarimetric preprocessing and colour-space kernels (pure HBM-bandwidth work).
//   pseudo-diffuse min-of-4          utils.py:102-106
//   rgb->yuv + whole-image std scale  ShmGANwithSSpecSeg.py:480-484, :1271-1309
//   averageCbCr                       :505
//   generator input assembly          :509-531, :576-594, test.py:227-235
//   yuv->rgb of concat(Y, CbCr)       :544-553, :613-624
#include "common.cuh"

namespace {

// tf.image.rgb_to_yuv / yuv_to_rgb kernels (out = in @ K)
__device__ __forceinline__ void rgb2yuv(float r, float g, float b, float& y, float& u, float& v) {
    y = 0.299f * r + 0.587f * g + 0.114f * b;
    u = -0.14714119f * r + -0.28886916f * g + 0.43601035f * b;
    v = 0.61497538f * r + -0.51496512f * g + -0.10001026f * b;
}
__device__ __forceinline__ void yuv2rgb(float y, float u, float v, float& r, float& g, float& b) {
    r = y + 1.13988303f * v;
    g = y + -0.394642334f * u + -0.58062185f * v;
    b = y + 2.03206185f * u;
}

inline int flat_grid(long long total, int block = 256) {
    long long g = cdiv64(total, block);
    const long long cap = (long long)shm_num_sms() * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// ---- pseudo diffuse ---------------------------------------------------------------------------
__global__ void min4_f32_kernel(const float4* __restrict__ a, const float4* __restrict__ b, const float4* __restrict__ c,
                                const float4* __restrict__ d, float4* __restrict__ o, long long n4) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 x = __ldg(a + i), y = __ldg(b + i), z = __ldg(c + i), w = __ldg(d + i);
        float4 r;
        r.x = fminf(fminf(x.x, y.x), fminf(z.x, w.x));
        r.y = fminf(fminf(x.y, y.y), fminf(z.y, w.y));
        r.z = fminf(fminf(x.z, y.z), fminf(z.z, w.z));
        r.w = fminf(fminf(x.w, y.w), fminf(z.w, w.w));
        o[i] = r;
    }
}
__global__ void min4_u8_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, const uint4* __restrict__ c,
                               const uint4* __restrict__ d, uint4* __restrict__ o, long long n16) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x) {
        const uint4 x = __ldg(a + i), y = __ldg(b + i), z = __ldg(c + i), w = __ldg(d + i);
        uint4 r;
        r.x = __vminu4(__vminu4(x.x, y.x), __vminu4(z.x, w.x));
        r.y = __vminu4(__vminu4(x.y, y.y), __vminu4(z.y, w.y));
        r.z = __vminu4(__vminu4(x.z, y.z), __vminu4(z.z, w.z));
        r.w = __vminu4(__vminu4(x.w, y.w), __vminu4(z.w, w.w));
        o[i] = r;
    }
}
__global__ void min4_bf16_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, const uint4* __restrict__ c,
                                 const uint4* __restrict__ d, uint4* __restrict__ o, long long n8) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        uint4 x = __ldg(a + i), y = __ldg(b + i), z = __ldg(c + i), w = __ldg(d + i), r;
        const __nv_bfloat162* px = reinterpret_cast<const __nv_bfloat162*>(&x);
        const __nv_bfloat162* py = reinterpret_cast<const __nv_bfloat162*>(&y);
        const __nv_bfloat162* pz = reinterpret_cast<const __nv_bfloat162*>(&z);
        const __nv_bfloat162* pw = reinterpret_cast<const __nv_bfloat162*>(&w);
        __nv_bfloat162* pr = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
        for (int j = 0; j < 4; ++j) pr[j] = __hmin2(__hmin2(px[j], py[j]), __hmin2(pz[j], pw[j]));
        o[i] = r;
    }
}
template <typename T>
__global__ void min4_tail_kernel(const T* a, const T* b, const T* c, const T* d, T* o, long long beg, long long n) {
    for (long long i = beg + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        T m = a[i];
        if (b[i] < m) m = b[i];
        if (c[i] < m) m = c[i];
        if (d[i] < m) m = d[i];
        o[i] = m;
    }
}

// ---- yuv stats / standardise -------------------------------------------------------------------
// one block handles a pixel range of one image; sums[n] += (sum v, sum v^2) over the 3 yuv channels
__global__ void __launch_bounds__(256) yuv_stats_kernel(const float* __restrict__ rgb, int HW, double* __restrict__ sums, int ppb) {
    __shared__ double sm[32];
    const int n = blockIdx.y;
    const float* base = rgb + (long long)n * HW * 3;
    const int pbeg = blockIdx.x * ppb, pend = min(pbeg + ppb, HW);
    float s = 0.f, q = 0.f;
    // 4 pixels = 12 floats = 3 float4 per thread iteration (HW % 4 == 0 and ppb % 4 == 0 enforced by the host)
    for (int p = pbeg + threadIdx.x * 4; p < pend; p += blockDim.x * 4) {
        const float4* src = reinterpret_cast<const float4*>(base + (long long)p * 3);
        const float4 a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2);
        const float px[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float y, u, v;
            rgb2yuv(px[3 * j], px[3 * j + 1], px[3 * j + 2], y, u, v);
            s += y + u + v;
            q = fmaf(y, y, fmaf(u, u, fmaf(v, v, q)));
        }
    }
    const double bs = block_sum((double)s, sm);
    const double bq = block_sum((double)q, sm);
    if (threadIdx.x == 0) { atomicAdd(&sums[n * 2 + 0], bs); atomicAdd(&sums[n * 2 + 1], bq); }
}

__global__ void __launch_bounds__(256) yuv_standardize_kernel(const float* __restrict__ rgb, int HW, const double* __restrict__ sums,
                                                              float* __restrict__ yuv, float* __restrict__ scale, int ppb) {
    const int n = blockIdx.y;
    const double cnt = (double)HW * 3.0;
    const double mean = sums[n * 2] / cnt;
    double var = sums[n * 2 + 1] / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    float sc = (float)sqrt(var);
    sc = fmaxf(sc, 1.0f / 256.0f);          // rsqrt(num_pixels = 65536), ShmGANwithSSpecSeg.py:1280,1299
    if (blockIdx.x == 0 && threadIdx.x == 0) scale[n] = sc;
    const float inv = 1.0f / sc;
    const float* base = rgb + (long long)n * HW * 3;
    float* ob = yuv + (long long)n * HW * 3;
    const int pbeg = blockIdx.x * ppb, pend = min(pbeg + ppb, HW);
    for (int p = pbeg + threadIdx.x * 4; p < pend; p += blockDim.x * 4) {
        const float4* src = reinterpret_cast<const float4*>(base + (long long)p * 3);
        const float4 a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2);
        const float px[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
        float o[12];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float y, u, v;
            rgb2yuv(px[3 * j], px[3 * j + 1], px[3 * j + 2], y, u, v);
            o[3 * j] = y / sc; o[3 * j + 1] = u / sc; o[3 * j + 2] = v / sc;
        }
        (void)inv;
        float4* dst = reinterpret_cast<float4*>(ob + (long long)p * 3);
        dst[0] = make_float4(o[0], o[1], o[2], o[3]);
        dst[1] = make_float4(o[4], o[5], o[6], o[7]);
        dst[2] = make_float4(o[8], o[9], o[10], o[11]);
    }
}

__global__ void avg_cbcr_kernel(const float* __restrict__ y0, const float* __restrict__ y1, const float* __restrict__ y2,
                                const float* __restrict__ y3, const float* __restrict__ y4, float* __restrict__ out, long long npix) {
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (long long)gridDim.x * blockDim.x) {
        const long long i = p * 3;
        const float u = (((y0[i + 1] + y1[i + 1]) + y2[i + 1]) + y3[i + 1]) + y4[i + 1];
        const float v = (((y0[i + 2] + y1[i + 2]) + y2[i + 2]) + y3[i + 2]) + y4[i + 2];
        reinterpret_cast<float2*>(out)[p] = make_float2(u / 5.0f, v / 5.0f);
    }
}

struct AsmSrc { const float* p[5]; int ld[5]; };

template <typename T>
__global__ void assemble_kernel(AsmSrc s, int onehot, T* __restrict__ out, long long npix, int ldo) {
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (long long)gridDim.x * blockDim.x) {
        T* o = out + p * ldo;
#pragma unroll
        for (int j = 0; j < 5; ++j) stf(o + j, s.p[j] ? __ldg(s.p[j] + p * s.ld[j]) : 0.f);
#pragma unroll
        for (int j = 0; j < 5; ++j) stf(o + 5 + j, j == onehot ? 1.f : 0.f);
        for (int j = 10; j < ldo; ++j) stf(o + j, 0.f);
    }
}

// bf16 rows of 64 channels (the zero-padded tensor-core input): 8 threads per pixel, one 16-byte store each
// (parts = 8; parts = 2 serves the 16-channel rows of the thin first layer)
__global__ void assemble64_kernel(AsmSrc s, int onehot, uint4* __restrict__ out, long long npix, int parts) {
    const long long total = npix * parts;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / parts; const int part = (int)(i - p * parts);
        uint4 u = make_uint4(0u, 0u, 0u, 0u);
        if (part < 2) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = part * 8 + j;
                v[j] = c < 5 ? (s.p[c] ? __ldg(s.p[c] + p * s.ld[c]) : 0.f) : (c < 10 ? (c - 5 == onehot ? 1.f : 0.f) : 0.f);
            }
            __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
            __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4], v[5]), h3 = __floats2bfloat162_rn(v[6], v[7]);
            u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
            u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
        }
        out[i] = u;
    }
}

// dst[p, 0..C) = src[p, 0..C), dst[p, C..64) = 0: the zero-padded input of a tensor-core first layer
template <typename S>
__global__ void pad64_kernel(const S* __restrict__ src, int lds, int C, uint4* __restrict__ out, long long npix, int parts) {
    const long long total = npix * parts;                  // parts = padded width / 8 (8 for 64 channels, 4 for 32, 2 for 16)
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / parts; const int part = (int)(i - p * parts);
        uint4 u = make_uint4(0u, 0u, 0u, 0u);
        if (part * 8 < C) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = part * 8 + j < C ? ldf(src + p * lds + part * 8 + j) : 0.f;
            __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
            __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4], v[5]), h3 = __floats2bfloat162_rn(v[6], v[7]);
            u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
            u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
        }
        out[i] = u;
    }
}


// im2col of a 3x3 stride-2 SAME convolution on a few-channel image (the discriminator's first layer d1: Conv 3 -> 64, s2,
// ShmGANwithSSpecSeg.py:353,:387).  Presenting the 3-channel image zero-padded to 64 channels moved 128 B per INPUT pixel through
// every d1 kernel (wgrad ran at 10 TFLOP/s); folding the 27 patch values into the channel axis makes d1 a 1x1 convolution over
// [N, H/2, W/2, 64] (27 real channels), one 128-byte row per OUTPUT pixel.  channel k = (ky*3 + kx)*C + c.
template <typename S, int CT>
__global__ void im2col_k3s2_kernel(const S* __restrict__ x, int ldx, int H, int W, int Crt, int pb, uint4* __restrict__ out, long long nout) {
    const int C = CT > 0 ? CT : Crt;                     // CT = 3 (the discriminator's RGB input): tap / channel split by a constant
    const int Ho = H / 2, Wo = W / 2;
    const long long total = nout * 8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long q = i >> 3; const int part = (int)(i & 7);
        uint4 u = make_uint4(0u, 0u, 0u, 0u);
        if (part * 8 < 9 * C) {
            const int ox = (int)(q % Wo); const long long t = q / Wo; const int oy = (int)(t % Ho); const long long n = t / Ho;
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int k = part * 8 + j;
                v[j] = 0.f;
                if (k < 9 * C) {
                    const int tap = k / C, c = k - tap * C;
                    const int iy = 2 * oy + tap / 3 - pb, ix = 2 * ox + tap % 3 - pb;
                    if (iy >= 0 && iy < H && ix >= 0 && ix < W) v[j] = ldf(x + ((n * H + iy) * W + ix) * ldx + c);
                }
            }
            __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
            __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4], v[5]), h3 = __floats2bfloat162_rn(v[6], v[7]);
            u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
            u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
        }
        out[i] = u;
    }
}

// The same for the discriminator's dense bf16 RGB input (C = 3, ldx = 3, Wo % 32 == 0), the only shape the training step uses: a warp takes 32
// consecutive output pixels of one output row.  The generic kernel above gathers 27 scalars per output pixel with a division chain each and ran
// at 1.1 TB/s (0.37 ms for the 160-image pass); here the three input rows of the strip (65 pixels x 3 channels, 4-byte aligned because W is
// even) are read once with coalesced 4-byte loads into shared memory, and every lane writes one 16-byte chunk per iteration -- 512 contiguous
// bytes per warp store -- from eight shared-memory offsets that do not depend on the iteration (chunk = 32 j + lane: part = lane & 7).
constexpr int IM2C_RS = 200;                              // bf16 elements per staged input row (195 used)
__global__ void __launch_bounds__(256) im2col3_strip_kernel(const bf16* __restrict__ x, int H, int W, long long strips, uint4* __restrict__ out) {
    __shared__ __align__(16) bf16 stage[8][3 * IM2C_RS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Ho = H >> 1, Wo = W >> 1, spr = Wo >> 5;   // strips per output row
    const long long strip = (long long)blockIdx.x * 8 + warp;
    if (strip >= strips) return;
    const int sx = (int)(strip % spr); const long long t = strip / spr; const int oy = (int)(t % Ho); const long long n = t / Ho;
    const int ox0 = sx << 5;
    bf16* st = stage[warp];
    const bool tail = 2 * ox0 + 64 < W;                   // the 65th input pixel of the strip exists (else it is the SAME padding: zero)
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const int iy = 2 * oy + r;
        uint32_t* dst = reinterpret_cast<uint32_t*>(st + r * IM2C_RS);
        if (iy < H) {
            const uint32_t* src = reinterpret_cast<const uint32_t*>(x + ((n * H + iy) * W + 2 * ox0) * 3);
            dst[lane] = __ldg(src + lane); dst[lane + 32] = __ldg(src + lane + 32); dst[lane + 64] = __ldg(src + lane + 64);
            if (lane < 2) dst[96 + lane] = tail ? __ldg(src + 96 + lane) : 0u;
        } else {
            dst[lane] = 0u; dst[lane + 32] = 0u; dst[lane + 64] = 0u;
            if (lane < 2) dst[96 + lane] = 0u;
        }
    }
    __syncwarp();
    const int part = lane & 7;
    int off[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int k = part * 8 + i;
        off[i] = k < 27 ? (k / 9) * IM2C_RS + (k % 9) : -1;
    }
    uint4* o = out + ((n * Ho + oy) * Wo + ox0) * 8;
    const unsigned short* s16 = reinterpret_cast<const unsigned short*>(st);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int p6 = (j * 4 + (lane >> 3)) * 6;         // pixel of this chunk x (2 input pixels x 3 channels)
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t lo = off[2 * i] >= 0 ? s16[off[2 * i] + p6] : 0u;
            const uint32_t hi = off[2 * i + 1] >= 0 ? s16[off[2 * i + 1] + p6] : 0u;
            w[i] = lo | (hi << 16);
        }
        o[j * 32 + lane] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// transpose of the above: dx[n, iy, ix, c] = sum over taps (ky, kx) with 2*oy + ky - pb == iy, 2*ox + kx - pb == ix of dP[n, oy, ox, (ky*3+kx)*C + c]
template <typename D>
__global__ void col2im_k3s2_kernel(const bf16* __restrict__ dP, int H, int W, int C, int pb, D* __restrict__ dx, int lddx, long long npix) {
    const int Ho = H / 2, Wo = W / 2;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (long long)gridDim.x * blockDim.x) {
        const int ix = (int)(p % W); const long long t = p / W; const int iy = (int)(t % H); const long long n = t / H;
        float acc[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = 0.f;
        for (int ky = 0; ky < 3; ++ky) {
            const int ty = iy + pb - ky;
            if (ty < 0 || (ty & 1) || (ty >> 1) >= Ho) continue;
            for (int kx = 0; kx < 3; ++kx) {
                const int tx = ix + pb - kx;
                if (tx < 0 || (tx & 1) || (tx >> 1) >= Wo) continue;
                const bf16* row = dP + ((n * Ho + (ty >> 1)) * Wo + (tx >> 1)) * 64 + (ky * 3 + kx) * C;
                for (int c = 0; c < C; ++c) acc[c] += __bfloat162float(row[c]);
            }
        }
        for (int c = 0; c < C; ++c) stf(dx + p * lddx + c, acc[c]);
    }
}

// The same for the training step's shape (C = 3, dense bf16 dx, W % 64 == 0): a warp produces 64 pixels of one input row.  Lane l loads the nine
// values (kx, c) of patch row ky of output pixel ox0 + l -- from the one or two output rows that reach this input row -- and owns input pixels
// 2 (ox0 + l) (taps kx = 0 of its own output pixel and kx = 2 of its left neighbour's, by shuffle) and 2 (ox0 + l) + 1 (tap kx = 1).  Same
// summation order as the generic kernel (ky, then kx ascending): identical results, 384 contiguous bytes per warp store.
__global__ void __launch_bounds__(256) col2im3_strip_kernel(const bf16* __restrict__ dP, int H, int W, long long strips, bf16* __restrict__ dx) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Ho = H >> 1, Wo = W >> 1, spr = W >> 6;
    const long long strip = (long long)blockIdx.x * 8 + warp;
    if (strip >= strips) return;
    const int sx = (int)(strip % spr); const long long t = strip / spr; const int iy = (int)(t % H); const long long n = t / H;
    const int ox = (sx << 5) + lane;
    float even[3] = {0.f, 0.f, 0.f}, odd[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        const int ty = iy - ky;
        const bool rowok = ty >= 0 && !(ty & 1) && (ty >> 1) < Ho;      // warp-uniform
        if (!rowok) continue;
        // the nine values sit at elements 9 ky .. 9 ky + 8 of the 64-element patch row: two aligned 16-byte loads (window of 16 elements starting
        // at element 8 * (9 ky / 8)) instead of nine 2-byte ones
        const bf16* prow = dP + ((n * Ho + (ty >> 1)) * Wo + ox) * 64;
        const int j0 = (ky * 9) >> 3, e0 = ky * 9 - (j0 << 3);          // ky = 0, 1, 2 -> window 0, 1, 2; first element 0, 1, 2
        auto window = [&](const bf16* r, float* w9) {
            const uint4 a = __ldg(reinterpret_cast<const uint4*>(r) + j0), b = __ldg(reinterpret_cast<const uint4*>(r) + j0 + 1);
            const uint32_t u[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                const int e = e0 + i;
                w9[i] = __uint_as_float((e & 1) ? (u[e >> 1] & 0xffff0000u) : (u[e >> 1] << 16));
            }
        };
        float v[9];
        window(prow, v);
        float left[3];                                                   // kx = 2 of output pixel ox - 1
#pragma unroll
        for (int c = 0; c < 3; ++c) left[c] = __shfl_up_sync(0xffffffffu, v[6 + c], 1);
        if (lane == 0) {
            float lv[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (ox > 0) window(prow - 64, lv);
#pragma unroll
            for (int c = 0; c < 3; ++c) left[c] = lv[6 + c];
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) { even[c] += v[c]; even[c] += left[c]; odd[c] += v[3 + c]; }
    }
    // six bf16 = three 4-byte words per lane
    __nv_bfloat162 w0 = __floats2bfloat162_rn(even[0], even[1]), w1 = __floats2bfloat162_rn(even[2], odd[0]), w2 = __floats2bfloat162_rn(odd[1], odd[2]);
    uint32_t* o = reinterpret_cast<uint32_t*>(dx + ((n * H + iy) * W + 2 * ox) * 3);
    o[0] = *reinterpret_cast<uint32_t*>(&w0); o[1] = *reinterpret_cast<uint32_t*>(&w1); o[2] = *reinterpret_cast<uint32_t*>(&w2);
}

struct Slots { int s[5]; int n; };
template <typename T>
__global__ void assemble_bwd_kernel(const T* __restrict__ din, int ldin, Slots sl, float* __restrict__ dgen, long long npix) {
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (long long)gridDim.x * blockDim.x) {
        float acc = 0.f;
        for (int j = 0; j < sl.n; ++j) acc += ldf(din + p * ldin + sl.s[j]);
        dgen[p] += acc;
    }
}

template <typename T>
__global__ void yuv2rgb_kernel(const float* __restrict__ Y, const float* __restrict__ cbcr, long long npix_c, float* __restrict__ rgb,
                               T* __restrict__ rgb_lp, int ld_lp, long long npix) {
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (long long)gridDim.x * blockDim.x) {
        const float2 uv = __ldg(reinterpret_cast<const float2*>(cbcr) + (p % npix_c));
        float r, g, b;
        yuv2rgb(__ldg(Y + p), uv.x, uv.y, r, g, b);
        if (rgb) { rgb[p * 3] = r; rgb[p * 3 + 1] = g; rgb[p * 3 + 2] = b; }
        if (rgb_lp) {
            T* o = rgb_lp + p * ld_lp;
            if (sizeof(T) == 2 && ld_lp == 64) {          // zero-padded 64-channel bf16 row: 8 x 16-byte stores
                __nv_bfloat162 h0 = __floats2bfloat162_rn(r, g), h1 = __floats2bfloat162_rn(b, 0.f);
                uint4 u = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1), 0u, 0u);
                uint4* o4 = reinterpret_cast<uint4*>(o);
                o4[0] = u;
#pragma unroll
                for (int j = 1; j < 8; ++j) o4[j] = make_uint4(0u, 0u, 0u, 0u);
            } else {
                stf(o, r); stf(o + 1, g); stf(o + 2, b);
                for (int j = 3; j < ld_lp; ++j) stf(o + j, 0.f);
            }
        }
    }
}

template <typename T>
__global__ void yuv2rgb_bwd_kernel(const float* __restrict__ da, const T* __restrict__ db, int ldb, float* __restrict__ dY, long long npix, int accumulate) {
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (long long)gridDim.x * blockDim.x) {
        float acc = accumulate ? dY[p] : 0.f;
        // d/dY of (r,g,b) = (1,1,1): first row of the yuv->rgb kernel
        if (da) acc += (da[p * 3] + da[p * 3 + 1]) + da[p * 3 + 2];
        if (db) acc += (ldf(db + p * ldb) + ldf(db + p * ldb + 1)) + ldf(db + p * ldb + 2);
        dY[p] = acc;
    }
}

}  // namespace

extern "C" int shm_pseudo_diffuse_min4(const void* i0, const void* i45, const void* i90, const void* i135, void* out, int64_t n,
                                       int dtype, void* stream) {
    SHM_REQUIRE(i0 && i45 && i90 && i135 && out && n >= 0, "shm_pseudo_diffuse_min4: bad args");
    if (n == 0) return SHM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const uintptr_t al = reinterpret_cast<uintptr_t>(i0) | reinterpret_cast<uintptr_t>(i45) | reinterpret_cast<uintptr_t>(i90) |
                         reinterpret_cast<uintptr_t>(i135) | reinterpret_cast<uintptr_t>(out);
    const bool aligned = (al & 15) == 0;
    const int esize = dtype == 0 ? 4 : (dtype == 1 ? 2 : 1);
    SHM_REQUIRE(dtype >= 0 && dtype <= 2, "shm_pseudo_diffuse_min4: dtype %d (0 f32, 1 bf16, 2 u8)", dtype);
    const int per = 16 / esize;
    const long long nv = aligned ? n / per : 0;
    if (nv > 0) {
        const int grid = flat_grid(nv);
        if (dtype == 0) min4_f32_kernel<<<grid, 256, 0, st>>>((const float4*)i0, (const float4*)i45, (const float4*)i90, (const float4*)i135, (float4*)out, nv);
        else if (dtype == 1) min4_bf16_kernel<<<grid, 256, 0, st>>>((const uint4*)i0, (const uint4*)i45, (const uint4*)i90, (const uint4*)i135, (uint4*)out, nv);
        else min4_u8_kernel<<<grid, 256, 0, st>>>((const uint4*)i0, (const uint4*)i45, (const uint4*)i90, (const uint4*)i135, (uint4*)out, nv);
        SHM_CHECK_LAUNCH("min4_kernel");
    }
    const long long done = nv * per;
    if (done < n) {
        const int grid = flat_grid(n - done);
        if (dtype == 0) min4_tail_kernel<float><<<grid, 256, 0, st>>>((const float*)i0, (const float*)i45, (const float*)i90, (const float*)i135, (float*)out, done, n);
        else if (dtype == 1) min4_tail_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)i0, (const bf16*)i45, (const bf16*)i90, (const bf16*)i135, (bf16*)out, done, n);
        else min4_tail_kernel<unsigned char><<<grid, 256, 0, st>>>((const unsigned char*)i0, (const unsigned char*)i45, (const unsigned char*)i90, (const unsigned char*)i135, (unsigned char*)out, done, n);
        SHM_CHECK_LAUNCH("min4_tail_kernel");
    }
    return SHM_OK;
}

static inline int yuv_ppb(int HW, int N) {
    long long want = (long long)shm_num_sms() * 8 / (N > 0 ? N : 1);
    if (want < 1) want = 1;
    long long ppb = cdiv64(HW, want);
    if (ppb < 1024) ppb = 1024;
    if (ppb > 65536) ppb = 65536;       // <= 256 pixels (768 values) per fp32 partial
    return (int)(cdiv64(ppb, 4) * 4);
}

extern "C" int shm_yuv_stats(const float* rgb, int N, int HW, double* sums, void* stream) {
    SHM_REQUIRE(rgb && sums && N > 0 && HW > 0, "shm_yuv_stats: bad args");
    SHM_REQUIRE(HW % 4 == 0 && (reinterpret_cast<uintptr_t>(rgb) & 15) == 0, "shm_yuv_stats: HW must be a multiple of 4 and rgb 16B-aligned");
    const int ppb = yuv_ppb(HW, N);
    yuv_stats_kernel<<<dim3(cdiv(HW, ppb), N), 256, 0, (cudaStream_t)stream>>>(rgb, HW, sums, ppb);
    SHM_CHECK_LAUNCH("yuv_stats_kernel");
    return SHM_OK;
}

extern "C" int shm_yuv_standardize(const float* rgb, int N, int HW, const double* sums, float* yuv, float* scale, void* stream) {
    SHM_REQUIRE(rgb && sums && yuv && scale && N > 0 && HW > 0, "shm_yuv_standardize: bad args");
    SHM_REQUIRE(HW % 4 == 0 && ((reinterpret_cast<uintptr_t>(rgb) | reinterpret_cast<uintptr_t>(yuv)) & 15) == 0,
                "shm_yuv_standardize: HW must be a multiple of 4 and buffers 16B-aligned");
    const int ppb = yuv_ppb(HW, N);
    yuv_standardize_kernel<<<dim3(cdiv(HW, ppb), N), 256, 0, (cudaStream_t)stream>>>(rgb, HW, sums, yuv, scale, ppb);
    SHM_CHECK_LAUNCH("yuv_standardize_kernel");
    return SHM_OK;
}

extern "C" int shm_avg_cbcr(const float* y0, const float* y1, const float* y2, const float* y3, const float* y4, float* out,
                            int64_t npix, void* stream) {
    SHM_REQUIRE(y0 && y1 && y2 && y3 && y4 && out && npix > 0, "shm_avg_cbcr: bad args");
    avg_cbcr_kernel<<<flat_grid(npix), 256, 0, (cudaStream_t)stream>>>(y0, y1, y2, y3, y4, out, npix);
    SHM_CHECK_LAUNCH("avg_cbcr_kernel");
    return SHM_OK;
}

extern "C" int shm_assemble_input(const float* const src[5], const int32_t src_ld[5], int onehot, void* out, int ldo, int64_t npix, int dtype, void* stream) {
    SHM_REQUIRE(src && src_ld && out && npix > 0 && onehot >= 0 && onehot < 5 && ldo >= 10, "shm_assemble_input: bad args");
    AsmSrc s;
    for (int j = 0; j < 5; ++j) { s.p[j] = src[j]; s.ld[j] = src_ld[j]; }
    if (dtype == SHM_BF16 && (ldo == 64 || ldo == 16) && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
        assemble64_kernel<<<flat_grid(npix * (ldo / 8)), 256, 0, (cudaStream_t)stream>>>(s, onehot, (uint4*)out, npix, ldo / 8);
        SHM_CHECK_LAUNCH("assemble64_kernel");
        return SHM_OK;
    }
    DISPATCH_DTYPE(dtype, T, {
        assemble_kernel<T><<<flat_grid(npix), 256, 0, (cudaStream_t)stream>>>(s, onehot, (T*)out, npix, ldo);
        SHM_CHECK_LAUNCH("assemble_kernel");
        return SHM_OK;
    })
}

extern "C" int shm_pad_channels64(const void* src, int src_dtype, int lds, int C, void* dst_bf16, int64_t npix, void* stream) {
    SHM_REQUIRE(src && dst_bf16 && npix > 0 && C >= 1 && C <= 64 && lds >= C, "shm_pad_channels64: bad args (1 <= C <= 64)");
    SHM_REQUIRE((reinterpret_cast<uintptr_t>(dst_bf16) & 15) == 0, "shm_pad_channels64: dst must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    if (src_dtype == SHM_F32) pad64_kernel<float><<<flat_grid(npix * 8), 256, 0, st>>>((const float*)src, lds, C, (uint4*)dst_bf16, npix, 8);
    else if (src_dtype == SHM_BF16) pad64_kernel<bf16><<<flat_grid(npix * 8), 256, 0, st>>>((const bf16*)src, lds, C, (uint4*)dst_bf16, npix, 8);
    else SHM_FAIL(SHM_EINVAL, "shm_pad_channels64: bad dtype %d", src_dtype);
    SHM_CHECK_LAUNCH("pad64_kernel");
    return SHM_OK;
}

extern "C" int shm_pad_channels(const void* src, int src_dtype, int lds, int C, void* dst_bf16, int Cpad, int64_t npix, void* stream) {
    SHM_REQUIRE(src && dst_bf16 && npix >= 0 && C >= 1 && (Cpad == 16 || Cpad == 32 || Cpad == 64) && C <= Cpad && lds >= C, "shm_pad_channels: bad args (Cpad in {16, 32, 64})");
    if (npix == 0) return SHM_OK;
    SHM_REQUIRE((reinterpret_cast<uintptr_t>(dst_bf16) & 15) == 0, "shm_pad_channels: dst must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int parts = Cpad / 8;
    if (src_dtype == SHM_F32) pad64_kernel<float><<<flat_grid(npix * parts), 256, 0, st>>>((const float*)src, lds, C, (uint4*)dst_bf16, npix, parts);
    else if (src_dtype == SHM_BF16) pad64_kernel<bf16><<<flat_grid(npix * parts), 256, 0, st>>>((const bf16*)src, lds, C, (uint4*)dst_bf16, npix, parts);
    else SHM_FAIL(SHM_EINVAL, "shm_pad_channels: bad dtype %d", src_dtype);
    SHM_CHECK_LAUNCH("pad64_kernel");
    return SHM_OK;
}

extern "C" int shm_im2col_k3s2(const void* x, int src_dtype, int ldx, int N, int H, int W, int C, void* out_bf16, void* stream) {
    SHM_REQUIRE(x && out_bf16 && N > 0 && H > 0 && W > 0 && C >= 1 && 9 * C <= 64 && ldx >= C, "shm_im2col_k3s2: bad args (9*C <= 64)");
    SHM_REQUIRE(H % 2 == 0 && W % 2 == 0, "shm_im2col_k3s2: H, W must be even");
    SHM_REQUIRE((reinterpret_cast<uintptr_t>(out_bf16) & 15) == 0, "shm_im2col_k3s2: out must be 16-byte aligned");
    const long long nout = (long long)N * (H / 2) * (W / 2);
    const int pb = same_pad_before(H, 3, 2);
    cudaStream_t st = (cudaStream_t)stream;
    if (src_dtype == SHM_F32) {
        if (C == 3) im2col_k3s2_kernel<float, 3><<<flat_grid(nout * 8), 256, 0, st>>>((const float*)x, ldx, H, W, C, pb, (uint4*)out_bf16, nout);
        else im2col_k3s2_kernel<float, 0><<<flat_grid(nout * 8), 256, 0, st>>>((const float*)x, ldx, H, W, C, pb, (uint4*)out_bf16, nout);
    } else if (src_dtype == SHM_BF16) {
        if (C == 3 && ldx == 3 && pb == 0 && (W / 2) % 32 == 0 && (reinterpret_cast<uintptr_t>(x) & 3) == 0) {
            const long long strips = (long long)N * (H / 2) * (W / 64);
            im2col3_strip_kernel<<<(unsigned)((strips + 7) / 8), 256, 0, st>>>((const bf16*)x, H, W, strips, (uint4*)out_bf16);
        } else
        if (C == 3) im2col_k3s2_kernel<bf16, 3><<<flat_grid(nout * 8), 256, 0, st>>>((const bf16*)x, ldx, H, W, C, pb, (uint4*)out_bf16, nout);
        else im2col_k3s2_kernel<bf16, 0><<<flat_grid(nout * 8), 256, 0, st>>>((const bf16*)x, ldx, H, W, C, pb, (uint4*)out_bf16, nout);
    }
    else SHM_FAIL(SHM_EINVAL, "shm_im2col_k3s2: bad dtype %d", src_dtype);
    SHM_CHECK_LAUNCH("im2col_k3s2_kernel");
    return SHM_OK;
}

extern "C" int shm_col2im_k3s2(const void* dP_bf16, int N, int H, int W, int C, void* dx, int dst_dtype, int lddx, void* stream) {
    SHM_REQUIRE(dP_bf16 && dx && N > 0 && H > 0 && W > 0 && C >= 1 && C <= 7 && lddx >= C, "shm_col2im_k3s2: bad args (C <= 7)");
    SHM_REQUIRE(H % 2 == 0 && W % 2 == 0, "shm_col2im_k3s2: H, W must be even");
    const long long npix = (long long)N * H * W;
    const int pb = same_pad_before(H, 3, 2);
    cudaStream_t st = (cudaStream_t)stream;
    if (dst_dtype == SHM_F32) col2im_k3s2_kernel<float><<<flat_grid(npix), 256, 0, st>>>((const bf16*)dP_bf16, H, W, C, pb, (float*)dx, lddx, npix);
    else if (dst_dtype == SHM_BF16 && C == 3 && lddx == 3 && pb == 0 && W % 64 == 0 && (reinterpret_cast<uintptr_t>(dx) & 3) == 0) {
        const long long strips = (long long)N * H * (W / 64);
        col2im3_strip_kernel<<<(unsigned)((strips + 7) / 8), 256, 0, st>>>((const bf16*)dP_bf16, H, W, strips, (bf16*)dx);
    }
    else if (dst_dtype == SHM_BF16) col2im_k3s2_kernel<bf16><<<flat_grid(npix), 256, 0, st>>>((const bf16*)dP_bf16, H, W, C, pb, (bf16*)dx, lddx, npix);
    else SHM_FAIL(SHM_EINVAL, "shm_col2im_k3s2: bad dtype %d", dst_dtype);
    SHM_CHECK_LAUNCH("col2im_k3s2_kernel");
    return SHM_OK;
}

extern "C" int shm_assemble_bwd(const void* din, int dtype, int ldin, const int32_t slots[5], int nslots, float* dgen, int64_t npix, void* stream) {
    SHM_REQUIRE(din && dgen && npix > 0 && nslots >= 0 && nslots <= 5 && ldin >= 10, "shm_assemble_bwd: bad args");
    if (nslots == 0) return SHM_OK;
    Slots sl; sl.n = nslots;
    for (int j = 0; j < 5; ++j) sl.s[j] = j < nslots ? slots[j] : 0;
    for (int j = 0; j < nslots; ++j) SHM_REQUIRE(slots[j] >= 0 && slots[j] < 5, "shm_assemble_bwd: slot out of range");
    DISPATCH_DTYPE(dtype, T, {
        assemble_bwd_kernel<T><<<flat_grid(npix), 256, 0, (cudaStream_t)stream>>>((const T*)din, ldin, sl, dgen, npix);
        SHM_CHECK_LAUNCH("assemble_bwd_kernel");
        return SHM_OK;
    })
}

extern "C" int shm_yuv2rgb(const float* Y, const float* cbcr, int64_t npix_cbcr, float* rgb, void* rgb_lp, int dtype_lp, int ld_lp, int64_t npix, void* stream) {
    SHM_REQUIRE(Y && cbcr && (rgb || rgb_lp) && npix > 0 && npix_cbcr > 0, "shm_yuv2rgb: bad args");
    SHM_REQUIRE(!rgb_lp || ld_lp >= 3, "shm_yuv2rgb: ld_lp must be >= 3");
    SHM_REQUIRE(!(rgb_lp && dtype_lp == SHM_BF16 && ld_lp == 64) || (reinterpret_cast<uintptr_t>(rgb_lp) & 15) == 0, "shm_yuv2rgb: padded output must be 16-byte aligned");
    SHM_REQUIRE(npix % npix_cbcr == 0, "shm_yuv2rgb: npix %% npix_cbcr != 0");
    DISPATCH_DTYPE(dtype_lp, T, {
        yuv2rgb_kernel<T><<<flat_grid(npix), 256, 0, (cudaStream_t)stream>>>(Y, cbcr, npix_cbcr, rgb, (T*)rgb_lp, ld_lp, npix);
        SHM_CHECK_LAUNCH("yuv2rgb_kernel");
        return SHM_OK;
    })
}

// running mean of the standardisation scales (the reference appends every pixel_value_scale to self.stddev_arr, ShmGANwithSSpecSeg.py:1306,
// and reads tf.reduce_mean(self.stddev_arr) at :548 / test.py:246): acc[0] += sum(src), acc[1] += n.  One block; n is a handful of values.
__global__ void sum_count_kernel(const float* __restrict__ src, long long n, double* __restrict__ acc) {
    __shared__ double sm[32];
    double s = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) s += (double)src[i];
    s = block_sum(s, sm);
    if (threadIdx.x == 0) { acc[0] += s; acc[1] += (double)n; }
}
// out = x * mul * (acc[0] / acc[1]): gen_rgb_output = yuv_to_rgb(gen_YCbCr * mean(stddev_arr) * 255) (:550) -- yuv_to_rgb is linear
__global__ void scale_by_mean_kernel(const float* __restrict__ x, const double* __restrict__ acc, float mul, float* __restrict__ out, long long n) {
    float f = mul;
    if (acc != nullptr) f *= acc[1] > 0.0 ? (float)(acc[0] / acc[1]) : 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = x[i] * f;
}

extern "C" int shm_sum_count(const float* src, int64_t n, double* acc, void* stream) {
    SHM_REQUIRE(src && acc && n > 0, "shm_sum_count: bad args");
    sum_count_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(src, n, acc);
    SHM_CHECK_LAUNCH("sum_count_kernel");
    return SHM_OK;
}

extern "C" int shm_scale_by_mean(const float* x, const double* acc, float mul, float* out, int64_t n, void* stream) {
    SHM_REQUIRE(x && out && n > 0, "shm_scale_by_mean: bad args");
    scale_by_mean_kernel<<<flat_grid(n), 256, 0, (cudaStream_t)stream>>>(x, acc, mul, out, n);
    SHM_CHECK_LAUNCH("scale_by_mean_kernel");
    return SHM_OK;
}

extern "C" int shm_yuv2rgb_bwd(const float* drgb_f32, const void* drgb_lp, int dtype_lp, int ld_lp, float* dY, int64_t npix, int accumulate, void* stream) {
    SHM_REQUIRE((drgb_f32 || drgb_lp) && dY && npix > 0 && (!drgb_lp || ld_lp >= 3), "shm_yuv2rgb_bwd: bad args");
    DISPATCH_DTYPE(dtype_lp, T, {
        yuv2rgb_bwd_kernel<T><<<flat_grid(npix), 256, 0, (cudaStream_t)stream>>>(drgb_f32, (const T*)drgb_lp, ld_lp, dY, npix, accumulate);
        SHM_CHECK_LAUNCH("yuv2rgb_bwd_kernel");
        return SHM_OK;
    })
}
