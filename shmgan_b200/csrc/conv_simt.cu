// conv_simt.cu -- exact-fp32 (SIMT FFMA) implicit-GEMM convolutions: the parity-mode path.
//
// One generalised "lattice convolution" kernel covers Conv2D fwd (stride 1/2, TF SAME), its dgrad,
// Conv2DTranspose fwd (sub-pixel decomposition by output parity) and its dgrad:
//     out[n, qy*OS+py, qx*OS+px, j] = act(bias[j] + sum_t sum_k in[n, qy*IS+dy_t, qx*IS+dx_t, k] * W[woff_t + k*w_ks + j*w_ns])
// and one generalised wgrad kernel covers both weight gradients:
//     dW[woff_t + a*w_as + b*w_bs] += sum_{n,q} A[n, q*SA+da_t, a] * B[n, q*SB+db_t, b]
// Replaces the cuDNN kernels TF dispatches for ShmGANwithSSpecSeg.py:244-326, :365, :387, :410-411 and SpecSeg.py:34-88.
#include "common.cuh"

namespace {

struct GTap { int dy, dx, woff; };

struct GConvParams {
    const void* in; void* out; const float* w; const float* bias;
    int N, Hin, Win;         // input tensor spatial dims
    int Qh, Qw;              // lattice dims
    int Hout, Wout;          // output tensor spatial dims
    int IS, OS, py, px;
    int K, Nn;               // reduction channels, output channels
    int ldin, ldout;
    int w_ks, w_ns;
    int ntaps;
    GTap taps[9];
    int act, accumulate;
};

constexpr int BM = 128, BK = 16, NT = 256;

template <typename T, int BN, bool VECA>
__global__ void __launch_bounds__(NT) gconv_kernel(const GConvParams p) {
    constexpr int TN = BN / 16;              // 8 or 4 columns per thread
    __shared__ __align__(16) float As[2][BK][BM];
    __shared__ __align__(16) float Bs[2][BK][BN];

    const T* __restrict__ in = reinterpret_cast<const T*>(p.in);
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const long long M = (long long)p.N * p.Qh * p.Qw;
    const long long m0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;

    // ---- A loader: row r = tid/2, 8 consecutive k at (tid&1)*8
    const int ar = tid >> 1, ak = (tid & 1) * 8;
    const long long am = m0 + ar;
    const bool arow_ok = am < M;
    int a_n = 0, a_iy0 = 0, a_ix0 = 0;
    if (arow_ok) {
        int qx = (int)(am % p.Qw);
        long long t = am / p.Qw;
        int qy = (int)(t % p.Qh);
        a_n = (int)(t / p.Qh);
        a_iy0 = qy * p.IS; a_ix0 = qx * p.IS;
    }
    // ---- B loader
    const bool b_ncontig = (p.w_ns == 1);
    const bool b_vec = b_ncontig ? ((p.Nn & 3) == 0 && (p.w_ks & 3) == 0) : ((p.K & 3) == 0 && (p.w_ns & 3) == 0);

    const int kchunks = (p.K + BK - 1) / BK;
    const int niter = p.ntaps * kchunks;

    float areg[8];
    float breg[BN / 16];     // BN*BK/NT elements per thread: 8 (BN=128) or 4 (BN=64)

    auto load_tiles = [&](int it) {
        const int t = it / kchunks;
        const int c0 = (it - t * kchunks) * BK;
        const GTap tap = p.taps[t];
        // A
        {
            const int iy = a_iy0 + tap.dy, ix = a_ix0 + tap.dx;
            const bool ok = arow_ok && iy >= 0 && iy < p.Hin && ix >= 0 && ix < p.Win;
            const int kk = c0 + ak;
            if (ok) {
                const T* src = in + ((long long)(a_n * p.Hin + iy) * p.Win + ix) * p.ldin + kk;
                if (VECA && kk + 8 <= p.K) {
                    float4 v0 = ld4(src), v1 = ld4(src + 4);
                    areg[0] = v0.x; areg[1] = v0.y; areg[2] = v0.z; areg[3] = v0.w;
                    areg[4] = v1.x; areg[5] = v1.y; areg[6] = v1.z; areg[7] = v1.w;
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) areg[j] = (kk + j < p.K) ? ldf(src + j) : 0.f;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) areg[j] = 0.f;
            }
        }
        // B: tile [BK][BN] from W[woff + k*w_ks + n*w_ns]
        {
            const float* wb = p.w + tap.woff;
            if (b_ncontig) {
                // threads sweep n fastest: element e = tid*TNB.. ; TNB = BN/16 consecutive n for one k
                constexpr int PER = BN / 16;                 // 8 or 4
                const int k = (tid * PER) / BN;              // 0..15
                const int n = (tid * PER) % BN;
                const int kk = c0 + k, nn = n0 + n;
                const float* src = wb + (long long)kk * p.w_ks + nn;
                if (kk < p.K && b_vec && nn + PER <= p.Nn) {
#pragma unroll
                    for (int j = 0; j < PER; j += 4) {
                        float4 v = __ldg(reinterpret_cast<const float4*>(src + j));
                        breg[j] = v.x; breg[j + 1] = v.y; breg[j + 2] = v.z; breg[j + 3] = v.w;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < PER; ++j) breg[j] = (kk < p.K && nn + j < p.Nn) ? __ldg(src + j) : 0.f;
                }
            } else {
                // k contiguous (dgrad forms): thread owns one n and PER consecutive k
                constexpr int PER = BN / 16;
                constexpr int TPN = BK / PER;                // threads per n: 2 (BN=128) or 4 (BN=64)
                const int n = tid / TPN;
                const int k = (tid % TPN) * PER;
                const int kk = c0 + k, nn = n0 + n;
                const float* src = wb + (long long)nn * p.w_ns + kk;
                if (nn < p.Nn && b_vec && kk + PER <= p.K) {
#pragma unroll
                    for (int j = 0; j < PER; j += 4) {
                        float4 v = __ldg(reinterpret_cast<const float4*>(src + j));
                        breg[j] = v.x; breg[j + 1] = v.y; breg[j + 2] = v.z; breg[j + 3] = v.w;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < PER; ++j) breg[j] = (nn < p.Nn && kk + j < p.K) ? __ldg(src + j) : 0.f;
                }
            }
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int j = 0; j < 8; ++j) As[buf][ak + j][ar] = areg[j];
        constexpr int PER = BN / 16;
        if (b_ncontig) {
            const int k = (tid * PER) / BN, n = (tid * PER) % BN;
#pragma unroll
            for (int j = 0; j < PER; ++j) Bs[buf][k][n + j] = breg[j];
        } else {
            constexpr int TPN = BK / PER;
            const int n = tid / TPN, k = (tid % TPN) * PER;
#pragma unroll
            for (int j = 0; j < PER; ++j) Bs[buf][k + j][n] = breg[j];
        }
    };

    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    for (int it = 0; it < niter; ++it) {
        const int buf = it & 1;
        if (it + 1 < niter) load_tiles(it + 1);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[8], b[TN];
            float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            if (TN == 8) {
                float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
                float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
                b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
                b[TN - 4] = b1.x; b[TN - 3] = b1.y; b[TN - 2] = b1.z; b[TN - 1] = b1.w;
            } else {
                float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
                b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (it + 1 < niter) store_tiles(buf ^ 1);
        __syncthreads();
    }

    // ---- epilogue: thread owns rows ty*8+i, columns {tx*4..+3} (+ {64+tx*4..+3} when TN == 8)
    T* __restrict__ out = reinterpret_cast<T*>(p.out);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const long long m = m0 + ty * 8 + i;
        if (m >= M) continue;
        int qx = (int)(m % p.Qw);
        long long t = m / p.Qw;
        int qy = (int)(t % p.Qh);
        int n = (int)(t / p.Qh);
        const int oy = qy * p.OS + p.py, ox = qx * p.OS + p.px;
        if (oy >= p.Hout || ox >= p.Wout) continue;
        T* dst = out + ((long long)(n * p.Hout + oy) * p.Wout + ox) * p.ldout;
#pragma unroll
        for (int h = 0; h < TN / 4; ++h) {
            const int col = n0 + h * 64 + tx * 4;
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float x = acc[i][h * 4 + j];
                if (p.bias != nullptr && col + j < p.Nn) x += __ldg(p.bias + col + j);
                v[j] = act_fwd(x, p.act);
            }
            if (col + 4 <= p.Nn && !p.accumulate && aligned4(dst + col)) {
                st4(dst + col, make_float4(v[0], v[1], v[2], v[3]));
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (col + j < p.Nn) {
                        float x = v[j];
                        if (p.accumulate) x += ldf(dst + col + j);
                        stf(dst + col + j, x);
                    }
            }
        }
    }
}

template <typename T>
int launch_gconv(const GConvParams& p, cudaStream_t st) {
    const long long M = (long long)p.N * p.Qh * p.Qw;
    if (M == 0 || p.Nn == 0) return SHM_OK;
    const bool veca = (p.K % 8 == 0) && (p.ldin % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.in) & (sizeof(T) * 4 - 1)) == 0);
    const bool bn64 = p.Nn <= 64;
    dim3 grid((unsigned)cdiv64(M, BM), (unsigned)cdiv(p.Nn, bn64 ? 64 : 128));
    if (bn64) {
        if (veca) gconv_kernel<T, 64, true><<<grid, NT, 0, st>>>(p);
        else      gconv_kernel<T, 64, false><<<grid, NT, 0, st>>>(p);
    } else {
        if (veca) gconv_kernel<T, 128, true><<<grid, NT, 0, st>>>(p);
        else      gconv_kernel<T, 128, false><<<grid, NT, 0, st>>>(p);
    }
    SHM_CHECK_LAUNCH("gconv_kernel");
    return SHM_OK;
}

// -------------------------------------------------------------------------------------------------
// generalised wgrad
// -------------------------------------------------------------------------------------------------
struct WTap { int day, dax, dby, dbx, woff; };
struct WgradParams {
    const void* A; const void* B; float* dW;
    int N, Qh, Qw;
    int HA, WA, SA, HB, WB, SB;
    int Ca, Cb, lda, ldb;
    int w_as, w_bs;
    int ntaps;
    WTap taps[9];
    int q_per_block;
};

template <typename T, bool VEC>
__global__ void __launch_bounds__(256) wgrad_kernel(const WgradParams p) {
    constexpr int TA = 64, TB = 64, KC = 16;
    __shared__ __align__(16) float As[2][KC][TA];
    __shared__ __align__(16) float Bs[2][KC][TB];
    const T* __restrict__ A = reinterpret_cast<const T*>(p.A);
    const T* __restrict__ B = reinterpret_cast<const T*>(p.B);
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int tiles_b = (p.Cb + TB - 1) / TB;
    const int a0 = (blockIdx.x / tiles_b) * TA, b0 = (blockIdx.x % tiles_b) * TB;
    const WTap tap = p.taps[blockIdx.y];
    const long long Q = (long long)p.N * p.Qh * p.Qw;
    const long long qbeg = (long long)blockIdx.z * p.q_per_block;
    long long qend = qbeg + p.q_per_block; if (qend > Q) qend = Q;
    if (qbeg >= qend) return;
    const int nchunks = (int)((qend - qbeg + KC - 1) / KC);

    const int lp = tid >> 4;           // pixel within chunk 0..15
    const int lc = (tid & 15) * 4;     // channel offset 0..60
    float4 ra, rb;

    auto load = [&](int ch) {
        const long long q = qbeg + (long long)ch * KC + lp;
        ra = make_float4(0.f, 0.f, 0.f, 0.f); rb = ra;
        if (q < qend) {
            int qx = (int)(q % p.Qw);
            long long t = q / p.Qw;
            int qy = (int)(t % p.Qh);
            int n = (int)(t / p.Qh);
            const int ay = qy * p.SA + tap.day, ax = qx * p.SA + tap.dax;
            const int by = qy * p.SB + tap.dby, bx = qx * p.SB + tap.dbx;
            const bool ok = ay >= 0 && ay < p.HA && ax >= 0 && ax < p.WA && by >= 0 && by < p.HB && bx >= 0 && bx < p.WB;
            if (ok) {
                const T* pa = A + ((long long)(n * p.HA + ay) * p.WA + ax) * p.lda + a0 + lc;
                const T* pb = B + ((long long)(n * p.HB + by) * p.WB + bx) * p.ldb + b0 + lc;
                if (VEC && a0 + lc + 4 <= p.Ca) ra = ld4(pa);
                else {
                    if (a0 + lc + 0 < p.Ca) ra.x = ldf(pa + 0);
                    if (a0 + lc + 1 < p.Ca) ra.y = ldf(pa + 1);
                    if (a0 + lc + 2 < p.Ca) ra.z = ldf(pa + 2);
                    if (a0 + lc + 3 < p.Ca) ra.w = ldf(pa + 3);
                }
                if (VEC && b0 + lc + 4 <= p.Cb) rb = ld4(pb);
                else {
                    if (b0 + lc + 0 < p.Cb) rb.x = ldf(pb + 0);
                    if (b0 + lc + 1 < p.Cb) rb.y = ldf(pb + 1);
                    if (b0 + lc + 2 < p.Cb) rb.z = ldf(pb + 2);
                    if (b0 + lc + 3 < p.Cb) rb.w = ldf(pb + 3);
                }
            }
        }
    };
    auto store = [&](int buf) {
        *reinterpret_cast<float4*>(&As[buf][lp][lc]) = ra;
        *reinterpret_cast<float4*>(&Bs[buf][lp][lc]) = rb;
    };

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    load(0); store(0); __syncthreads();
    for (int ch = 0; ch < nchunks; ++ch) {
        const int buf = ch & 1;
        if (ch + 1 < nchunks) load(ch + 1);
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            float4 a = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (ch + 1 < nchunks) store(buf ^ 1);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int a = a0 + ty * 4 + i;
        if (a >= p.Ca) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int b = b0 + tx * 4 + j;
            if (b < p.Cb) atomicAdd(p.dW + tap.woff + (long long)a * p.w_as + (long long)b * p.w_bs, acc[i][j]);
        }
    }
}

template <typename T>
int launch_wgrad(WgradParams& p, cudaStream_t st) {
    const long long Q = (long long)p.N * p.Qh * p.Qw;
    if (Q == 0) return SHM_OK;
    const int tiles = cdiv(p.Ca, 64) * cdiv(p.Cb, 64);
    // aim at >= 4 waves of 148 SMs x 2 resident blocks, but keep >= 256 pixels per block
    long long want_blocks = 148LL * 8;
    long long per_tile = want_blocks / ((long long)tiles * p.ntaps);
    if (per_tile < 1) per_tile = 1;
    long long qpb = cdiv64(Q, per_tile);
    if (qpb < 256) qpb = 256;
    qpb = cdiv64(qpb, 16) * 16;
    p.q_per_block = (int)qpb;
    const int nsplit = (int)cdiv64(Q, qpb);
    const bool vec = (p.lda % 4 == 0) && (p.ldb % 4 == 0) && (p.Ca % 4 == 0) && (p.Cb % 4 == 0) &&
                     ((reinterpret_cast<uintptr_t>(p.A) & (sizeof(T) * 4 - 1)) == 0) &&
                     ((reinterpret_cast<uintptr_t>(p.B) & (sizeof(T) * 4 - 1)) == 0);
    dim3 grid(tiles, p.ntaps, nsplit);
    if (vec) wgrad_kernel<T, true><<<grid, 256, 0, st>>>(p);
    else     wgrad_kernel<T, false><<<grid, 256, 0, st>>>(p);
    SHM_CHECK_LAUNCH("wgrad_kernel");
    return SHM_OK;
}

// column sums: out[c] += sum_pix x[pix*ld + c]
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ x, long long npix, int C, int ld, float* __restrict__ out, int pix_per_block) {
    // blockDim = (32 channels, 8 pixel lanes)
    __shared__ float red[8][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    const long long pbeg = (long long)blockIdx.y * pix_per_block;
    long long pend = pbeg + pix_per_block; if (pend > npix) pend = npix;
    float s = 0.f;
    if (c < C)
        for (long long pi = pbeg + threadIdx.y; pi < pend; pi += 8) s += ldf(x + pi * ld + c);
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
        atomicAdd(out + c, t);
    }
}

int check_desc(const shm_conv_desc* d) {
    SHM_REQUIRE(d != nullptr, "conv desc is NULL");
    SHM_REQUIRE(d->N > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0, "conv desc: non-positive dims");
    SHM_REQUIRE(d->kh >= 1 && d->kh <= 3 && d->kw >= 1 && d->kw <= 3, "conv desc: kernel %dx%d unsupported", d->kh, d->kw);
    SHM_REQUIRE(d->stride == 1 || d->stride == 2, "conv desc: stride %d unsupported", d->stride);
    SHM_REQUIRE(!d->transposed || d->stride == 2, "conv desc: transposed conv needs stride 2");
    SHM_REQUIRE(d->dtype == SHM_F32 || d->dtype == SHM_BF16, "conv desc: bad dtype");
    SHM_REQUIRE(d->ldx >= d->Cin && d->ldy >= d->Cout, "conv desc: ld smaller than channels");
    return SHM_OK;
}

inline void out_dims(const shm_conv_desc* d, int& Ho, int& Wo) {
    if (d->transposed) { Ho = d->H * d->stride; Wo = d->W * d->stride; }
    else { Ho = cdiv(d->H, d->stride); Wo = cdiv(d->W, d->stride); }
}

// "gather at stride" form: out lattice = small image; used by Conv2D fwd and ConvT dgrad
template <typename T>
int run_gather(const shm_conv_desc* d, const void* in, int Hin, int Win, int ldin, int K,
               void* out, int Hq, int Wq, int ldout, int Nn, const float* w, int w_ks, int w_ns,
               const float* bias, int act, int accumulate, cudaStream_t st) {
    // out[q] = sum_k in[q*s + k - pb]
    GConvParams p{};
    p.in = in; p.out = out; p.w = w; p.bias = bias;
    p.N = d->N; p.Hin = Hin; p.Win = Win; p.Qh = Hq; p.Qw = Wq; p.Hout = Hq; p.Wout = Wq;
    p.IS = d->stride; p.OS = 1; p.py = 0; p.px = 0; p.K = K; p.Nn = Nn; p.ldin = ldin; p.ldout = ldout;
    p.w_ks = w_ks; p.w_ns = w_ns; p.act = act; p.accumulate = accumulate;
    const int pby = same_pad_before(Hin, d->kh, d->stride), pbx = same_pad_before(Win, d->kw, d->stride);
    p.ntaps = 0;
    for (int ky = 0; ky < d->kh; ++ky)
        for (int kx = 0; kx < d->kw; ++kx) {
            p.taps[p.ntaps].dy = ky - pby; p.taps[p.ntaps].dx = kx - pbx;
            p.taps[p.ntaps].woff = (ky * d->kw + kx) * d->Cin * d->Cout;
            ++p.ntaps;
        }
    return launch_gconv<T>(p, st);
}

// "scatter by parity" form: out lattice = big image split in s*s parity classes; used by ConvT fwd and strided-conv dgrad
template <typename T>
int run_scatter(const shm_conv_desc* d, const void* in, int Hs, int Ws, int ldin, int K,
                void* out, int Hb, int Wb, int ldout, int Nn, const float* w, int w_ks, int w_ns,
                const float* bias, int act, int accumulate, cudaStream_t st) {
    // out[p] = sum_{o,k : s*o + k - pb = p} in[o] W[k];  p = s*q + r  =>  k = r + pb (mod s), o = q + (r + pb - k)/s
    const int s = d->stride;
    const int pby = same_pad_before(Hb, d->kh, s), pbx = same_pad_before(Wb, d->kw, s);
    for (int ry = 0; ry < s; ++ry)
        for (int rx = 0; rx < s; ++rx) {
            GConvParams p{};
            p.in = in; p.out = out; p.w = w; p.bias = bias;
            p.N = d->N; p.Hin = Hs; p.Win = Ws; p.Qh = cdiv(Hb - ry, s); p.Qw = cdiv(Wb - rx, s);
            p.Hout = Hb; p.Wout = Wb; p.IS = 1; p.OS = s; p.py = ry; p.px = rx;
            p.K = K; p.Nn = Nn; p.ldin = ldin; p.ldout = ldout; p.w_ks = w_ks; p.w_ns = w_ns;
            p.act = act; p.accumulate = accumulate;
            p.ntaps = 0;
            for (int ky = 0; ky < d->kh; ++ky) {
                if (((ry + pby - ky) % s) != 0) continue;
                for (int kx = 0; kx < d->kw; ++kx) {
                    if (((rx + pbx - kx) % s) != 0) continue;
                    p.taps[p.ntaps].dy = (ry + pby - ky) / s; p.taps[p.ntaps].dx = (rx + pbx - kx) / s;
                    p.taps[p.ntaps].woff = (ky * d->kw + kx) * d->Cin * d->Cout;
                    ++p.ntaps;
                }
            }
            if (p.ntaps == 0) {
                // a parity class no tap reaches (k < s): output = act(bias); emit with a zero-tap launch
                p.ntaps = 0;
            }
            int rc = launch_gconv<T>(p, st);
            if (rc) return rc;
        }
    return SHM_OK;
}

}  // namespace

extern "C" int shm_conv2d_fwd(const shm_conv_desc* d, const void* x, const float* w, const float* bias, void* y, void* stream) {
    if (int rc = check_desc(d)) return rc;
    SHM_REQUIRE(x && w && y, "shm_conv2d_fwd: NULL buffer");
    cudaStream_t st = (cudaStream_t)stream;
    int Ho, Wo; out_dims(d, Ho, Wo);
    DISPATCH_DTYPE(d->dtype, T, {
        if (!d->transposed)   // Conv2D kernel (kh,kw,Cin,Cout): k stride Cout, n stride 1
            return run_gather<T>(d, x, d->H, d->W, d->ldx, d->Cin, y, Ho, Wo, d->ldy, d->Cout, w, d->Cout, 1, bias, d->act, 0, st);
        // Conv2DTranspose kernel (kh,kw,Cout,Cin): k (=ci) stride 1, n (=co) stride Cin
        return run_scatter<T>(d, x, d->H, d->W, d->ldx, d->Cin, y, Ho, Wo, d->ldy, d->Cout, w, 1, d->Cin, bias, d->act, 0, st);
    })
}

extern "C" int shm_conv2d_dgrad(const shm_conv_desc* d, const void* dy, const float* w, void* dx, int accumulate, void* stream) {
    if (int rc = check_desc(d)) return rc;
    SHM_REQUIRE(dy && w && dx, "shm_conv2d_dgrad: NULL buffer");
    SHM_REQUIRE(!accumulate || d->dtype == SHM_F32, "shm_conv2d_dgrad: accumulate needs fp32");
    cudaStream_t st = (cudaStream_t)stream;
    int Ho, Wo; out_dims(d, Ho, Wo);
    DISPATCH_DTYPE(d->dtype, T, {
        if (!d->transposed)   // reduction over co (stride 1), output ci (stride Cout); dx = scatter of dy
            return run_scatter<T>(d, dy, Ho, Wo, d->ldy, d->Cout, dx, d->H, d->W, d->ldx, d->Cin, w, 1, d->Cout, nullptr, SHM_ACT_NONE, accumulate, st);
        // ConvT dgrad: dx[o,ci] = sum_k dy[2o+k-pb, co] W[k][co][ci]: gather at stride 2 from the big image
        return run_gather<T>(d, dy, Ho, Wo, d->ldy, d->Cout, dx, d->H, d->W, d->ldx, d->Cin, w, d->Cin, 1, nullptr, SHM_ACT_NONE, accumulate, st);
    })
}

extern "C" int shm_conv2d_wgrad(const shm_conv_desc* d, const void* x, const void* dy, float* dw, float* dbias, void* stream) {
    if (int rc = check_desc(d)) return rc;
    SHM_REQUIRE(x && dy && dw, "shm_conv2d_wgrad: NULL buffer");
    cudaStream_t st = (cudaStream_t)stream;
    int Ho, Wo; out_dims(d, Ho, Wo);
    WgradParams p{};
    p.A = x; p.B = dy; p.dW = dw; p.N = d->N;
    p.HA = d->H; p.WA = d->W; p.HB = Ho; p.WB = Wo; p.Ca = d->Cin; p.Cb = d->Cout; p.lda = d->ldx; p.ldb = d->ldy;
    p.ntaps = 0;
    if (!d->transposed) {
        const int pby = same_pad_before(d->H, d->kh, d->stride), pbx = same_pad_before(d->W, d->kw, d->stride);
        p.Qh = Ho; p.Qw = Wo; p.SA = d->stride; p.SB = 1; p.w_as = d->Cout; p.w_bs = 1;
        for (int ky = 0; ky < d->kh; ++ky)
            for (int kx = 0; kx < d->kw; ++kx) {
                WTap& t = p.taps[p.ntaps++];
                t.day = ky - pby; t.dax = kx - pbx; t.dby = 0; t.dbx = 0; t.woff = (ky * d->kw + kx) * d->Cin * d->Cout;
            }
    } else {
        const int pby = same_pad_before(Ho, d->kh, d->stride), pbx = same_pad_before(Wo, d->kw, d->stride);
        p.Qh = d->H; p.Qw = d->W; p.SA = 1; p.SB = d->stride; p.w_as = 1; p.w_bs = d->Cin;
        for (int ky = 0; ky < d->kh; ++ky)
            for (int kx = 0; kx < d->kw; ++kx) {
                WTap& t = p.taps[p.ntaps++];
                t.day = 0; t.dax = 0; t.dby = ky - pby; t.dbx = kx - pbx; t.woff = (ky * d->kw + kx) * d->Cin * d->Cout;
            }
    }
    DISPATCH_DTYPE(d->dtype, T, {
        if (int rc = launch_wgrad<T>(p, st)) return rc;
        if (dbias) return shm_colsum(dy, (long long)d->N * Ho * Wo, d->Cout, d->ldy, d->dtype, dbias, stream);
        return SHM_OK;
    })
}

extern "C" int shm_colsum(const void* dy, int64_t npix, int C, int ld, int dtype, float* out, void* stream) {
    SHM_REQUIRE(dy && out && npix > 0 && C > 0 && ld >= C, "shm_colsum: bad args");
    int ppb = (int)cdiv64(npix, 148 * 4 / cdiv(C, 32) + 1);
    if (ppb < 64) ppb = 64;
    dim3 grid(cdiv(C, 32), (unsigned)cdiv64(npix, ppb));
    DISPATCH_DTYPE(dtype, T, {
        colsum_kernel<T><<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(reinterpret_cast<const T*>(dy), npix, C, ld, out, ppb);
        SHM_CHECK_LAUNCH("colsum_kernel");
        return SHM_OK;
    })
}
