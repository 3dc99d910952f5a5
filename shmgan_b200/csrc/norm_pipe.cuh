// norm_pipe.cuh -- the bf16 instance-norm / activation-backward streams as cp.async-pipelined range kernels.
// (included by norm.cu inside its anonymous namespace, after f8 / ld8 / st8 / stat_ab_fast)
//
// Why: the register-staged kernels (`*8_kernel` in norm.cu) keep every in-flight 16-byte load in registers.  With the
// per-channel coefficients next to them they need 80-137 registers -> one to three 256-thread blocks per SM -> 32-48 KB in
// flight per SM, and measured 3.3-4.7 TB/s of the 6.5 TB/s copy peak.  Here the in-flight loads live in shared memory:
// every thread owns a private ring of S slots per input stream and fills it with `cp.async.cg` (LDGSTS: global -> shared
// without a destination register); it only ever reads back what it wrote itself, so the pipeline needs no block barrier,
// only `cp.async.wait_group`.  Registers hold the coefficients and ONE pixel -> 4 blocks per SM x 256 threads x S x 16 B x streams
// in flight per SM; ring depth S = 4 (about 64 KB per SM) measured best -- deeper rings LOSE bandwidth.
//
// Work split: the tensor is ONE pixel range [0, N*HW) cut into gridDim.x equal contiguous pieces; a block walks its piece image by
// image (coefficients reloaded at an image boundary, fp32 partial sums flushed to the fp64 atomics at least every SEG_MAX values).
// The grid is sized by bytes per block (range_grid below: ~128 KB of the primary stream, ~512 KB for the statistics kernels whose
// blocks end in atomics), not by the SM count: several short waves beat one long balanced wave because the pipeline-fill and
// flush phases of different blocks then overlap (sweep: profiles/r01_norm_pipe_sweep.txt).  HW % TY == 0 is required so that the
// TY pixel lanes of one iteration lie in one image (the host falls back to the *8 kernels otherwise).

__device__ __forceinline__ void cp16(uint32_t saddr, const void* g) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(g) : "memory");
}
// L1-allocating variant for the quarter-resolution gradient: the four pixels of a quad read the same 16 bytes
__device__ __forceinline__ void cp16_ca(uint32_t saddr, const void* g) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ f8 lds8(uint32_t saddr) {
    uint4 u;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(saddr) : "memory");
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    f8 r;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
        r.v[2 * i] = f.x; r.v[2 * i + 1] = f.y;
    }
    return r;
}

constexpr int SEG_MAX = 512;        // fp32 partial sums hold at most this many values before they go to the atomics

// this block's share [beg, end) of `total` units, boundaries on multiples of `align`
__device__ __forceinline__ void block_range(long long total, int align, long long& beg, long long& end) {
    const long long chunks = total / align;
    beg = chunks * blockIdx.x / gridDim.x * align;
    end = chunks * (blockIdx.x + 1) / gridDim.x * align;
}

// per-thread partial columns -> block totals -> sink(channel group t, value j, total), eight columns per pass; zeroes vals
template <int NV, typename Sink>
__device__ __forceinline__ void flush_lanes(float* vals, int TX, float (*red)[9], Sink sink) {
    const int TY = 256 / TX;
#pragma unroll
    for (int h = 0; h < NV; h += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { red[threadIdx.x][j] = vals[h + j]; vals[h + j] = 0.f; }
        __syncthreads();
        for (int e = threadIdx.x; e < TX * 8; e += 256) {
            const int t = e >> 3, j = e & 7;
            float acc = 0.f;
            for (int y = 0; y < TY; ++y) acc += red[y * TX + t][j];
            sink(t, h + j, acc);
        }
        __syncthreads();
    }
}

// Ring geometry: slot (stage s, stream k) of thread t sits at ring + ((s * K + k) * 256 + t) * 16 bytes (conflict-free LDS.128).
template <int S, int K>
struct Ring {
    uint32_t base, off;                 // shared address of this thread's stage-0 / stream-0 slot; byte offset of the current stage
    __device__ __forceinline__ Ring(const void* smem) : base((uint32_t)__cvta_generic_to_shared(smem) + threadIdx.x * 16), off(0) {}
    __device__ __forceinline__ uint32_t slot(int k) const { return base + off + k * 4096; }
    __device__ __forceinline__ void advance() { off += K * 4096; if (off == S * K * 4096) off = 0; }
};

// ---- stats: sums[n][c] += (sum x, sum x^2) -----------------------------------------------------------------------------------
template <int S>
__global__ void __launch_bounds__(256, 4) in_stats_p(const bf16* __restrict__ x, int HW, int C, int ldx, long long NP,
                                                     double* __restrict__ sums, int TX) {
    extern __shared__ __align__(16) unsigned char ring_mem[];
    __shared__ float red[256][9];
    const int TY = 256 / TX, tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    long long beg, end;
    block_range(NP, TY, beg, end);
    const int iters = (int)((end - beg) / TY);
    const long long step = (long long)TY * ldx;
    const bf16* src = x + (beg + ty) * ldx + tx * 8;
    Ring<S, 1> fill(ring_mem), use(ring_mem);
#pragma unroll
    for (int s = 0; s < S; ++s) {
        if (s < iters) cp16(fill.slot(0), src);
        src += step; fill.advance(); cp_commit();
    }
    float a[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = 0.f;
    int n = (int)(beg / HW);
    int left = (int)min((long long)(n + 1) * HW - beg, (long long)SEG_MAX * TY) / TY;      // iterations until the next flush
    for (int i = 0; i < iters; ++i) {
        cp_wait<S - 1>();
        const f8 v = lds8(use.slot(0));
        use.advance();
#pragma unroll
        for (int j = 0; j < 8; ++j) { a[j] += v.v[j]; a[8 + j] = fmaf(v.v[j], v.v[j], a[8 + j]); }
        if (i + S < iters) cp16(fill.slot(0), src);
        src += step; fill.advance(); cp_commit();
        if (--left == 0 || i + 1 == iters) {
            flush_lanes<16>(a, TX, red, [&](int t, int j, float acc) {
                atomicAdd(&sums[((long long)n * C + t * 8 + (j & 7)) * 2 + (j >> 3)], (double)acc);
            });
            const long long g = beg + (long long)(i + 1) * TY;
            n = (int)(g / HW);
            left = (int)min((long long)(n + 1) * HW - g, (long long)SEG_MAX * TY) / TY;
        }
    }
}

// ---- apply: out = IN(x) (+ add) ----------------------------------------------------------------------------------------------
template <int S, bool ADD>
__global__ void __launch_bounds__(256, 4) in_apply_p(const bf16* __restrict__ x, int HW, int C, int ldx, long long NP,
        const double* __restrict__ sums, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
        const bf16* __restrict__ add, int ldadd, int nadd, bf16* __restrict__ out, int ldo, int TX) {
    extern __shared__ __align__(16) unsigned char ring_mem[];
    constexpr int K = ADD ? 2 : 1;
    const int TY = 256 / TX, tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const int c0 = tx * 8;
    long long beg, end;
    block_range(NP, TY, beg, end);
    const int iters = (int)((end - beg) / TY);
    const double inv_hw = 1.0 / (double)HW;
    const long long step = (long long)TY * ldx, stepo = (long long)TY * ldo;
    const bf16* src = x + (beg + ty) * ldx + c0;
    bf16* dst = out + (beg + ty) * ldo + c0;
    // broadcast operand: image n % nadd, pixel p.  `fp` walks (image, pixel) of the fill position.
    int fn = (int)(beg / HW), fp = (int)(beg - (long long)fn * HW) + ty;
    Ring<S, K> fill(ring_mem), use(ring_mem);
    auto issue = [&](bool on) {
        if (on) {
            cp16(fill.slot(0), src);
            if (ADD) cp16(fill.slot(1), add + ((long long)(fn % nadd) * HW + fp) * ldadd + c0);
        }
        src += step; fill.advance(); cp_commit();
        if (ADD) { fp += TY; if (fp >= HW) { fp -= HW; ++fn; } }
    };
#pragma unroll
    for (int s = 0; s < S; ++s) issue(s < iters);
    int n = (int)(beg / HW) - 1;
    int left = 0;                                   // iterations left in the current image
    float a[8], b[8];
    for (int i = 0; i < iters; ++i) {
        if (left == 0) {
            const long long g = beg + (long long)i * TY;
            n = (int)(g / HW);
            left = (int)(((long long)(n + 1) * HW - g) / TY);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float mean, rstd;
                stat_ab_fast(sums, n, C, c0 + j, inv_hw, eps, mean, rstd);
                a[j] = rstd * __ldg(gamma + c0 + j);
                b[j] = __ldg(beta + c0 + j) - mean * a[j];
            }
        }
        --left;
        cp_wait<S - 1>();
        f8 y = lds8(use.slot(0));
#pragma unroll
        for (int j = 0; j < 8; ++j) y.v[j] = fmaf(y.v[j], a[j], b[j]);
        if (ADD) {
            const f8 e = lds8(use.slot(1));
#pragma unroll
            for (int j = 0; j < 8; ++j) y.v[j] += e.v[j];
        }
        use.advance();
        st8(dst, y);
        dst += stepo;
        issue(i + S < iters);
    }
}

// ---- apply + AvgPool2: units are 2x2 quads.  pooled = mean of the four IN(x); out (optional) = IN(x) (+ add) ------------------
template <int S, bool OUT, bool ADD>
__global__ void __launch_bounds__(256, 3) in_apply_pool_p(const bf16* __restrict__ x, int H, int W, int C, int ldx, long long NQ,
        const double* __restrict__ sums, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
        const bf16* __restrict__ add, int ldadd, int nadd, bf16* __restrict__ out, int ldo, bf16* __restrict__ pooled, int ldp, int TX) {
    extern __shared__ __align__(16) unsigned char ring_mem[];
    constexpr int K = ADD ? 8 : 4;
    const int TY = 256 / TX, tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const int c0 = tx * 8;
    const int HW = H * W, Wq = W >> 1, Q = HW >> 2;
    long long beg, end;
    block_range(NQ, TY, beg, end);
    const int iters = (int)((end - beg) / TY);
    const double inv_hw = 1.0 / (double)HW;
    bf16* pdst = pooled + (beg + ty) * ldp + c0;
    const long long stepp = (long long)TY * ldp;
    int fn = (int)(beg / Q), fu = (int)(beg - (long long)fn * Q) + ty;          // fill position: image, quad
    Ring<S, K> fill(ring_mem), use(ring_mem);
    auto issue = [&](bool on) {
        if (on) {
            const int qy = fu / Wq, qx = fu - qy * Wq;
            const long long p0 = (long long)fn * HW + (2 * qy) * W + 2 * qx;
            const bf16* s0 = x + p0 * ldx + c0;
            cp16(fill.slot(0), s0);
            cp16(fill.slot(1), s0 + ldx);
            cp16(fill.slot(2), s0 + (long long)W * ldx);
            cp16(fill.slot(3), s0 + (long long)(W + 1) * ldx);
            if (ADD) {
                const bf16* a0 = add + ((long long)(fn % nadd) * HW + (2 * qy) * W + 2 * qx) * ldadd + c0;
                cp16(fill.slot(4), a0);
                cp16(fill.slot(5), a0 + ldadd);
                cp16(fill.slot(6), a0 + (long long)W * ldadd);
                cp16(fill.slot(7), a0 + (long long)(W + 1) * ldadd);
            }
        }
        fill.advance(); cp_commit();
        fu += TY; if (fu >= Q) { fu -= Q; ++fn; }
    };
#pragma unroll
    for (int s = 0; s < S; ++s) issue(s < iters);
    int n = 0, u = 0, left = 0;
    float a[8], b[8];
    for (int i = 0; i < iters; ++i) {
        if (left == 0) {
            const long long g = beg + (long long)i * TY;
            n = (int)(g / Q);
            u = (int)(g - (long long)n * Q) + ty;
            left = (Q - (u - ty)) / TY;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float mean, rstd;
                stat_ab_fast(sums, n, C, c0 + j, inv_hw, eps, mean, rstd);
                a[j] = rstd * __ldg(gamma + c0 + j);
                b[j] = __ldg(beta + c0 + j) - mean * a[j];
            }
        }
        --left;
        cp_wait<S - 1>();
        f8 acc;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc.v[j] = 0.f;
        bf16* o0 = nullptr;
        if (OUT) {
            const int qy = u / Wq, qx = u - qy * Wq;
            o0 = out + ((long long)n * HW + (2 * qy) * W + 2 * qx) * ldo + c0;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            f8 y = lds8(use.slot(q));
#pragma unroll
            for (int j = 0; j < 8; ++j) { y.v[j] = fmaf(y.v[j], a[j], b[j]); acc.v[j] += y.v[j]; }
            if (OUT) {
                if (ADD) {
                    const f8 e = lds8(use.slot(4 + q));
#pragma unroll
                    for (int j = 0; j < 8; ++j) y.v[j] += e.v[j];
                }
                st8(o0 + ((long long)(q >> 1) * W + (q & 1)) * ldo, y);
            }
        }
        use.advance();
#pragma unroll
        for (int j = 0; j < 8; ++j) acc.v[j] *= 0.25f;
        st8(pdst, acc);
        pdst += stepp;
        u += TY;
        issue(i + S < iters);
    }
}

// ---- backward.  dy = dyA + 0.25 * upsample2(dyP); MODE bit 0: dyA present, bit 1: dyP present -----------------------------------
template <int MODE> struct DyStreams { static constexpr int K = 1 + ((MODE & 1) ? 1 : 0) + ((MODE & 2) ? 1 : 0); };

// fill position shared by the two backward kernels: issues x / dyA / dyP of one pixel into the current fill stage
template <int S, int MODE>
struct BwdFill {
    Ring<S, DyStreams<MODE>::K> ring;
    const bf16 *x, *dyA, *dyP;
    int ldx, ldA, ldP, W, Wq, HW, TY, c0;
    int fn, fp;
    __device__ __forceinline__ BwdFill(const void* smem) : ring(smem) {}
    __device__ __forceinline__ void issue(bool on) {
        if (on) {
            const long long g = (long long)fn * HW + fp;
            cp16(ring.slot(0), x + g * ldx + c0);
            if (MODE & 1) cp16(ring.slot(1), dyA + g * ldA + c0);
            if (MODE & 2) {
                const int y = fp / W, xx = fp - y * W;
                cp16_ca(ring.slot((MODE & 1) ? 2 : 1), dyP + ((long long)fn * (HW >> 2) + (y >> 1) * Wq + (xx >> 1)) * ldP + c0);
            }
        }
        ring.advance(); cp_commit();
        fp += TY; if (fp >= HW) { fp -= HW; ++fn; }
    }
};
template <int S, int MODE>
__device__ __forceinline__ f8 read_dy(const Ring<S, DyStreams<MODE>::K>& use) {
    f8 d;
    if (MODE & 1) d = lds8(use.slot(1));
    if (MODE & 2) {
        const f8 e = lds8(use.slot((MODE & 1) ? 2 : 1));
#pragma unroll
        for (int j = 0; j < 8; ++j) d.v[j] = (MODE & 1) ? fmaf(0.25f, e.v[j], d.v[j]) : 0.25f * e.v[j];
    }
    return d;
}

template <int S, int MODE>
__global__ void __launch_bounds__(256, 4) in_bwd_stats_p(const bf16* __restrict__ x, int H, int W, int C, int ldx, long long NP,
        const double* __restrict__ sums, float eps, const bf16* __restrict__ dyA, int ldA, const bf16* __restrict__ dyP, int ldP,
        double* __restrict__ bsums, int TX) {
    extern __shared__ __align__(16) unsigned char ring_mem[];
    __shared__ float red[256][9];
    constexpr int K = DyStreams<MODE>::K;
    const int TY = 256 / TX, tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const int HW = H * W;
    long long beg, end;
    block_range(NP, TY, beg, end);
    const int iters = (int)((end - beg) / TY);
    const double inv_hw = 1.0 / (double)HW;
    BwdFill<S, MODE> fill(ring_mem);
    fill.x = x; fill.dyA = dyA; fill.dyP = dyP; fill.ldx = ldx; fill.ldA = ldA; fill.ldP = ldP;
    fill.W = W; fill.Wq = W >> 1; fill.HW = HW; fill.TY = TY; fill.c0 = tx * 8;
    fill.fn = (int)(beg / HW); fill.fp = (int)(beg - (long long)fill.fn * HW) + ty;
    Ring<S, K> use(ring_mem);
#pragma unroll
    for (int s = 0; s < S; ++s) fill.issue(s < iters);
    float a[16], mean[8], rstd[8];
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = 0.f;
    int n = -1, left = 0;
    for (int i = 0; i < iters; ++i) {
        if (left == 0) {
            const long long g = beg + (long long)i * TY;
            const int nn = (int)(g / HW);
            left = (int)min((long long)(nn + 1) * HW - g, (long long)SEG_MAX * TY) / TY;
            if (nn != n) {
                n = nn;
#pragma unroll
                for (int j = 0; j < 8; ++j) stat_ab_fast(sums, n, C, tx * 8 + j, inv_hw, eps, mean[j], rstd[j]);
            }
        }
        cp_wait<S - 1>();
        const f8 v = lds8(use.slot(0));
        const f8 d = read_dy<S, MODE>(use);
        use.advance();
#pragma unroll
        for (int j = 0; j < 8; ++j) { a[j] += d.v[j]; a[8 + j] = fmaf(d.v[j], (v.v[j] - mean[j]) * rstd[j], a[8 + j]); }
        fill.issue(i + S < iters);
        if (--left == 0 || i + 1 == iters) {
            flush_lanes<16>(a, TX, red, [&](int t, int j, float acc) {
                atomicAdd(&bsums[((long long)n * C + t * 8 + (j & 7)) * 2 + (j >> 3)], (double)acc);
            });
            left = 0;
        }
    }
}

// dx = act'(x) * rstd*gamma*(dy - mean(dy) - xhat*mean(dy*xhat)); dbias[c] += sum over pixels of dx (the producing conv's bias
// gradient, ShmGANwithSSpecSeg.py:244: the conv bias sits right behind this tensor) when dbias != NULL
template <int S, int MODE>
__global__ void __launch_bounds__(256, 4) in_bwd_apply_p(const bf16* __restrict__ x, int H, int W, int C, int ldx, long long NP,
        const double* __restrict__ sums, const float* __restrict__ gamma, float eps,
        const bf16* __restrict__ dyA, int ldA, const bf16* __restrict__ dyP, int ldP, const double* __restrict__ bsums, int act,
        bf16* __restrict__ dx, int lddx, float* __restrict__ dbias, int TX) {
    extern __shared__ __align__(16) unsigned char ring_mem[];
    __shared__ float red[256][9];
    constexpr int K = DyStreams<MODE>::K;
    const int TY = 256 / TX, tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const int HW = H * W, c0 = tx * 8;
    long long beg, end;
    block_range(NP, TY, beg, end);
    const int iters = (int)((end - beg) / TY);
    const double inv_hw = 1.0 / (double)HW;
    const float neg = act == SHM_ACT_LRELU ? 0.2f : (act == SHM_ACT_RELU ? 0.f : 1.f);
    BwdFill<S, MODE> fill(ring_mem);
    fill.x = x; fill.dyA = dyA; fill.dyP = dyP; fill.ldx = ldx; fill.ldA = ldA; fill.ldP = ldP;
    fill.W = W; fill.Wq = W >> 1; fill.HW = HW; fill.TY = TY; fill.c0 = c0;
    fill.fn = (int)(beg / HW); fill.fp = (int)(beg - (long long)fill.fn * HW) + ty;
    Ring<S, K> use(ring_mem);
#pragma unroll
    for (int s = 0; s < S; ++s) fill.issue(s < iters);
    bf16* dst = dx + (beg + ty) * lddx + c0;
    const long long stepo = (long long)TY * lddx;
    // r = g * (k0 + k1 * d + k2 * v) with g = act'(v):  a*(d - m1 - (v-mean)*rstd*m2)
    float k0[8], k1[8], k2[8], s[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = 0.f;
    int n = -1, left = 0;
    for (int i = 0; i < iters; ++i) {
        if (left == 0) {
            const long long g = beg + (long long)i * TY;
            const int nn = (int)(g / HW);
            left = (int)min((long long)(nn + 1) * HW - g, (long long)SEG_MAX * TY) / TY;
            if (nn != n) {
                n = nn;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float mean, rstd;
                    stat_ab_fast(sums, n, C, c0 + j, inv_hw, eps, mean, rstd);
                    const float a = rstd * __ldg(gamma + c0 + j);
                    const double2 bs = __ldg(reinterpret_cast<const double2*>(bsums + ((long long)n * C + c0 + j) * 2));
                    const float m1 = (float)(bs.x * inv_hw);
                    const float m2 = (float)(bs.y * inv_hw);
                    k1[j] = a; k2[j] = -a * rstd * m2; k0[j] = -a * m1 + a * rstd * m2 * mean;
                }
            }
        }
        cp_wait<S - 1>();
        const f8 v = lds8(use.slot(0));
        const f8 d = read_dy<S, MODE>(use);
        use.advance();
        f8 r;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float g = v.v[j] > 0.f ? 1.f : neg;
            r.v[j] = g * fmaf(k2[j], v.v[j], fmaf(k1[j], d.v[j], k0[j]));
            s[j] += r.v[j];
        }
        st8(dst, r);
        dst += stepo;
        fill.issue(i + S < iters);
        if (--left == 0 || i + 1 == iters) {
            if (dbias != nullptr)
                flush_lanes<8>(s, TX, red, [&](int t, int j, float acc) { atomicAdd(dbias + t * 8 + j, acc); });
            left = 0;
        }
    }
}

// ---- dpre = dy * act'(y) with the bias gradient fused: dbias[c] += sum over pixels of dpre -----------------------------------
template <int S>
__global__ void __launch_bounds__(256, 4) act_bwd_p(const bf16* __restrict__ dy, int lddy, const bf16* __restrict__ y, int ldy,
        bf16* __restrict__ dpre, int ldd, long long NP, int act, float* __restrict__ dbias, int TX) {
    extern __shared__ __align__(16) unsigned char ring_mem[];
    __shared__ float red[256][9];
    const int TY = 256 / TX, tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const int c0 = tx * 8;
    const float neg = act == SHM_ACT_LRELU ? 0.2f : (act == SHM_ACT_RELU ? 0.f : 1.f);
    long long beg, end;
    block_range(NP, TY, beg, end);
    const int iters = (int)((end - beg) / TY);
    const bf16* sy = y + (beg + ty) * ldy + c0;
    const bf16* sd = dy + (beg + ty) * lddy + c0;
    bf16* dst = dpre + (beg + ty) * ldd + c0;
    const long long stepy = (long long)TY * ldy, stepd = (long long)TY * lddy, stepo = (long long)TY * ldd;
    Ring<S, 2> fill(ring_mem), use(ring_mem);
    auto issue = [&](bool on) {
        if (on) { cp16(fill.slot(0), sy); cp16(fill.slot(1), sd); }
        sy += stepy; sd += stepd; fill.advance(); cp_commit();
    };
#pragma unroll
    for (int s = 0; s < S; ++s) issue(s < iters);
    float s[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = 0.f;
    int left = SEG_MAX;
    for (int i = 0; i < iters; ++i) {
        cp_wait<S - 1>();
        const f8 v = lds8(use.slot(0)), d = lds8(use.slot(1));
        use.advance();
        f8 r;
#pragma unroll
        for (int j = 0; j < 8; ++j) { r.v[j] = d.v[j] * (v.v[j] > 0.f ? 1.f : neg); s[j] += r.v[j]; }
        st8(dst, r);
        dst += stepo;
        issue(i + S < iters);
        if (--left == 0 || i + 1 == iters) {
            if (dbias != nullptr)
                flush_lanes<8>(s, TX, red, [&](int t, int j, float acc) { atomicAdd(dbias + t * 8 + j, acc); });
            left = SEG_MAX;
        }
    }
}

// ---- launch plumbing ---------------------------------------------------------------------------------------------------------
int g_pipe_off = 0;              // 1: use the register-staged *8 kernels instead (comparison runs of tools/bench_norm.py)
int g_pipe_depth = 0;            // 0: per-kernel default ring depth; else forced (2, 4, 8)
int g_grid_mul = 0;              // 0: one resident wave (SMs x occupancy); k > 0: k x SMs blocks

inline int occ_of(const void* kernel, int smem) {
    static const void* keys[128];
    static int smems[128], vals[128];
    static int used = 0;
    for (int i = 0; i < used; ++i) if (keys[i] == kernel && smems[i] == smem) return vals[i];
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);      // dynamic + static may pass 48 KB
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, 256, smem) != cudaSuccess || occ < 1) occ = 1;
    if (used < 128) { keys[used] = kernel; smems[used] = smem; vals[used] = occ; ++used; }
    return occ;
}
// Grid: about `target` bytes of the primary stream per block (measured sweet spot, profiles/r01_norm_pipe_sweep.txt: ~128 KB for the
// read+write streams, ~512 KB for the statistics kernels whose blocks end in atomics), at least one resident wave when the tensor
// is big enough to give every block 64 KB, at most 32 blocks per SM.
inline int range_grid(const void* kernel, int smem, long long units, int TY, long long bytes, long long target) {
    const int occ = occ_of(kernel, smem);
    const long long sms = shm_num_sms();
    long long g;
    if (g_grid_mul > 0) g = sms * g_grid_mul;
    else {
        g = bytes / target;
        long long lo = bytes / (64 << 10);
        if (lo > sms * occ) lo = sms * occ;
        if (g < lo) g = lo;
        if (g > sms * 32) g = sms * 32;
    }
    const long long cap = units / ((long long)TY * 4);            // at least four pixel rows per thread
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}
inline int pipe_depth(int dflt) { return g_pipe_depth > 0 ? g_pipe_depth : dflt; }

// the pipelined kernels serve bf16 tensors with C = 8 * 2^k <= 2048, 16-byte aligned pixels, and a pixel count per image that
// the TY = 256 / (C/8) pixel lanes divide
inline bool pipe_ok(int dtype, int C, long long HW) {
    if (g_pipe_off || dtype != SHM_BF16 || C % 8 != 0) return false;
    const int tx = C / 8;
    if (tx > 256 || (tx & (tx - 1)) != 0) return false;
    return HW % (256 / tx) == 0;
}

#define PIPE_DEPTHS(S_, CALL) if (S_ == 1) { CALL(1) } else if (S_ == 2) { CALL(2) } else if (S_ == 4) { CALL(4) } else { CALL(8) }
#define PIPE_DEPTHS3(S_, CALL) if (S_ == 1) { CALL(1) } else if (S_ == 2) { CALL(2) } else { CALL(4) }
#define PIPE_LAUNCH(KERNEL, K_, UNITS, TY_, BYTES, TARGET, ...) { auto k__ = KERNEL; const int sm__ = (K_) * 4096 * S__; \
    k__<<<range_grid((const void*)k__, sm__, UNITS, TY_, BYTES, TARGET), 256, sm__, st>>>(__VA_ARGS__); }
