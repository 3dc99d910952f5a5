// loss.cu -- fused loss value + seed-gradient kernels, clip+Adam and the counter-based RNG.
//   LSGAN / softmax-CE / L1 / content / style / SSIM / Spec terms   ShmGANwithSSpecSeg.py:669-844
//   rescale_01                                                        utils.py:190-195
//   clip_by_value + Keras Adam                                        ShmGANwithSSpecSeg.py:860-871, :169-175
// Every loss kernel ADDS weight * value into loss_out[0] (fp32, device) so a step reads all scalars back with one copy.
#include "common.cuh"
#include <cooperative_groups.h>

namespace {

inline int flat_grid(long long total, int block = 256) {
    long long g = cdiv64(total, block);
    const long long cap = (long long)shm_num_sms() * 8;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// ---- elementwise reductions ----------------------------------------------------------------------
// mode 0: (a - target)^2 (target = *b when b != NULL: a device-resident scalar) ; mode 1: |a - b| ; mode 2: (a - b)^2
template <int MODE>
__global__ void __launch_bounds__(256) ew_loss_kernel(const float* __restrict__ a, const float* __restrict__ b, float target, long long n,
                                                      float* __restrict__ loss_out, float weight, float* __restrict__ da, float gscale, int accumulate) {
    __shared__ double sm[32];
    float acc = 0.f;
    const float inv_n = 1.0f / (float)n;
    if (MODE == 0 && b != nullptr) target = __ldg(b);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float d = a[i] - (MODE == 0 ? target : b[i]);
        float g;
        if (MODE == 1) { acc += fabsf(d); g = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f); }
        else { acc = fmaf(d, d, acc); g = 2.f * d; }
        if (da) { const float v = gscale * g * inv_n; da[i] = accumulate ? da[i] + v : v; }
    }
    const double s = block_sum((double)acc, sm);
    if (threadIdx.x == 0 && loss_out) atomicAdd(loss_out, (float)(s / (double)n) * weight);
}

struct Labels5 { float v[5]; };
__global__ void softmax_ce_kernel(const float* __restrict__ logits, int B, Labels5 lab, const float* __restrict__ lab_dev,
                                  float* __restrict__ loss_out, float weight, float* __restrict__ dlogits, float gscale, int accumulate) {
    __shared__ double sm[32];
    float acc = 0.f;
    if (lab_dev != nullptr) {
#pragma unroll
        for (int j = 0; j < 5; ++j) lab.v[j] = __ldg(lab_dev + j);
    }
    const float lsum = lab.v[0] + lab.v[1] + lab.v[2] + lab.v[3] + lab.v[4];
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
        float z[5], m = -INFINITY;
#pragma unroll
        for (int j = 0; j < 5; ++j) { z[j] = logits[b * 5 + j]; m = fmaxf(m, z[j]); }
        float se = 0.f;
#pragma unroll
        for (int j = 0; j < 5; ++j) se += expf(z[j] - m);
        const float lse = m + logf(se);
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            acc += lab.v[j] * (lse - z[j]);
            if (dlogits) {
                const float g = gscale * (lsum * expf(z[j] - lse) - lab.v[j]) / (float)B;
                dlogits[b * 5 + j] = accumulate ? dlogits[b * 5 + j] + g : g;
            }
        }
    }
    const double s = block_sum((double)acc, sm);
    if (threadIdx.x == 0 && loss_out) atomicAdd(loss_out, (float)(s / (double)B) * weight);
}

// content loss on concat(Y, cbcr) vs yuv: mean over [npix, 3]; gradient only through Y (:814)
__global__ void __launch_bounds__(256) mse_ycc_kernel(const float* __restrict__ Y, const float* __restrict__ cbcr, const float* __restrict__ yuv,
        long long npix, float* __restrict__ loss_out, float weight, float* __restrict__ dY, float gscale) {
    __shared__ double sm[32];
    float acc = 0.f;
    const float inv_n = 1.0f / (float)(npix * 3);
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (long long)gridDim.x * blockDim.x) {
        const float d0 = Y[p] - yuv[p * 3], d1 = cbcr[p * 2] - yuv[p * 3 + 1], d2 = cbcr[p * 2 + 1] - yuv[p * 3 + 2];
        acc = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, acc)));
        if (dY) dY[p] += gscale * 2.f * d0 * inv_n;
    }
    const double s = block_sum((double)acc, sm);
    if (threadIdx.x == 0 && loss_out) atomicAdd(loss_out, (float)(s / (double)(npix * 3)) * weight);
}

// masked "Spec" L2 (:792-806): mean(((gen*mask) - (ref*mask))^2) over [npix,3]
__global__ void __launch_bounds__(256) spec_kernel(const float* __restrict__ Y, const float* __restrict__ cbcr, const float* __restrict__ yuv,
        const float* __restrict__ mask, long long npix, float* __restrict__ loss_out, float weight) {
    __shared__ double sm[32];
    float acc = 0.f;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (long long)gridDim.x * blockDim.x) {
        const float m = mask[p];
        const float d0 = Y[p] * m - yuv[p * 3] * m, d1 = cbcr[p * 2] * m - yuv[p * 3 + 1] * m, d2 = cbcr[p * 2 + 1] * m - yuv[p * 3 + 2] * m;
        acc = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, acc)));
    }
    const double s = block_sum((double)acc, sm);
    if (threadIdx.x == 0) atomicAdd(loss_out, (float)(s / (double)(npix * 3)) * weight);
}

__device__ __forceinline__ void load3(const float* Yp, const float* cbcr, int HW, int n, int p, float v[3]) {
    if (cbcr) { v[0] = Yp[(long long)n * HW + p]; v[1] = cbcr[((long long)n * HW + p) * 2]; v[2] = cbcr[((long long)n * HW + p) * 2 + 1]; }
    else { const float* s = Yp + ((long long)n * HW + p) * 3; v[0] = s[0]; v[1] = s[1]; v[2] = s[2]; }
}

// per-image min / max / argmin / argmax over the [HW,3] values.  One CLUSTER of eight 1024-thread blocks per image (one block per image left
// 132 of the 148 SMs idle at batch 16: 39 us per call, five calls per step): every block scans an eighth of the image, block 0 of the cluster
// collects the eight partial results through distributed shared memory.
constexpr int MM_CLUSTER = 8;
__global__ void __cluster_dims__(MM_CLUSTER, 1, 1) __launch_bounds__(1024) minmax3_kernel(const float* __restrict__ Yp, const float* __restrict__ cbcr, int HW,
                                                                                          float* __restrict__ mm, int* __restrict__ idx) {
    __shared__ float smn[32], smx[32];
    __shared__ int simn[32], simx[32];
    __shared__ float res_v[2];
    __shared__ int res_i[2];
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int n = blockIdx.x / MM_CLUSTER;
    float mn = INFINITY, mx = -INFINITY;
    int imn = 0x7fffffff, imx = 0x7fffffff;
    for (int p = rank * (int)blockDim.x + threadIdx.x; p < HW; p += MM_CLUSTER * blockDim.x) {
        float v[3];
        load3(Yp, cbcr, HW, n, p, v);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int id = p * 3 + c;
            if (v[c] < mn) { mn = v[c]; imn = id; }
            if (v[c] > mx) { mx = v[c]; imx = id; }
        }
    }
    // warp then block reduction, ties -> smallest flat index
    for (int o = 16; o > 0; o >>= 1) {
        const float omn = __shfl_xor_sync(0xffffffffu, mn, o), omx = __shfl_xor_sync(0xffffffffu, mx, o);
        const int oimn = __shfl_xor_sync(0xffffffffu, imn, o), oimx = __shfl_xor_sync(0xffffffffu, imx, o);
        if (omn < mn || (omn == mn && oimn < imn)) { mn = omn; imn = oimn; }
        if (omx > mx || (omx == mx && oimx < imx)) { mx = omx; imx = oimx; }
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { smn[w] = mn; smx[w] = mx; simn[w] = imn; simx[w] = imx; }
    __syncthreads();
    if (w == 0) {
        const int nw = blockDim.x >> 5;
        mn = lane < nw ? smn[lane] : INFINITY; mx = lane < nw ? smx[lane] : -INFINITY;
        imn = lane < nw ? simn[lane] : 0x7fffffff; imx = lane < nw ? simx[lane] : 0x7fffffff;
        for (int o = 16; o > 0; o >>= 1) {
            const float omn = __shfl_xor_sync(0xffffffffu, mn, o), omx = __shfl_xor_sync(0xffffffffu, mx, o);
            const int oimn = __shfl_xor_sync(0xffffffffu, imn, o), oimx = __shfl_xor_sync(0xffffffffu, imx, o);
            if (omn < mn || (omn == mn && oimn < imn)) { mn = omn; imn = oimn; }
            if (omx > mx || (omx == mx && oimx < imx)) { mx = omx; imx = oimx; }
        }
        if (lane == 0) { res_v[0] = mn; res_v[1] = mx; res_i[0] = imn; res_i[1] = imx; }
    }
    cluster.sync();                                        // every block's partial result is in its shared memory
    if (rank == 0 && w == 0) {
        mn = INFINITY; mx = -INFINITY; imn = 0x7fffffff; imx = 0x7fffffff;
        if (lane < MM_CLUSTER) {
            const float* rv = cluster.map_shared_rank(res_v, lane);
            const int* ri = cluster.map_shared_rank(res_i, lane);
            mn = rv[0]; mx = rv[1]; imn = ri[0]; imx = ri[1];
        }
        for (int o = 16; o > 0; o >>= 1) {
            const float omn = __shfl_xor_sync(0xffffffffu, mn, o), omx = __shfl_xor_sync(0xffffffffu, mx, o);
            const int oimn = __shfl_xor_sync(0xffffffffu, imn, o), oimx = __shfl_xor_sync(0xffffffffu, imx, o);
            if (omn < mn || (omn == mn && oimn < imn)) { mn = omn; imn = oimn; }
            if (omx > mx || (omx == mx && oimx < imx)) { mx = omx; imx = oimx; }
        }
        if (lane == 0) { mm[n * 2] = mn; mm[n * 2 + 1] = mx; idx[n * 2] = imn; idx[n * 2 + 1] = imx; }
    }
    cluster.sync();                                        // no block leaves while block 0 may still read its shared memory
}

// gram[n][c*3+d] += sum_p x_c x_d   (un-normalised; the 1/HW is applied by the consumers)
__global__ void __launch_bounds__(256) gram3_kernel(const float* __restrict__ Yp, const float* __restrict__ cbcr, int HW,
                                                    double* __restrict__ gram, int ppb) {
    __shared__ double sm[32];
    const int n = blockIdx.y;
    const int pbeg = blockIdx.x * ppb, pend = min(pbeg + ppb, HW);
    float g[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // 00 01 02 11 12 22
    for (int p = pbeg + threadIdx.x; p < pend; p += blockDim.x) {
        float v[3];
        load3(Yp, cbcr, HW, n, p, v);
        g[0] = fmaf(v[0], v[0], g[0]); g[1] = fmaf(v[0], v[1], g[1]); g[2] = fmaf(v[0], v[2], g[2]);
        g[3] = fmaf(v[1], v[1], g[3]); g[4] = fmaf(v[1], v[2], g[4]); g[5] = fmaf(v[2], v[2], g[5]);
    }
    const int map[6][2] = {{0, 0}, {1, 3}, {2, 6}, {4, 4}, {5, 7}, {8, 8}};
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const double s = block_sum((double)g[k], sm);
        if (threadIdx.x == 0) {
            atomicAdd(&gram[n * 9 + map[k][0]], s);
            if (map[k][1] != map[k][0]) atomicAdd(&gram[n * 9 + map[k][1]], s);
        }
    }
}

// style = factor * mean_{b,c,d} (GA/HW - GB/HW)^2 ; dgramA[b][cd] = gscale * factor * 2 (GA-GB)/HW / (9B)  (w.r.t. the normalised gram)
__global__ void style_kernel(const double* __restrict__ gA, const double* __restrict__ gB, int N, int HW, double factor,
                             float* __restrict__ loss_out, float weight, float* __restrict__ dgramA, float gscale) {
    __shared__ double sm[32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < N * 9; i += blockDim.x) {
        const double d = (gA[i] - gB[i]) / (double)HW;
        acc += d * d;
        if (dgramA) dgramA[i] = (float)((double)gscale * factor * 2.0 * d / (9.0 * N));
    }
    const double s = block_sum(acc, sm);
    if (threadIdx.x == 0 && loss_out) atomicAdd(loss_out, (float)(factor * s / (9.0 * N)) * weight);
}

// dY[n,p] += sum_d (dG[0,d] + dG[d,0]) x_d(p) / HW
__global__ void gram3_bwd_kernel(const float* __restrict__ Y, const float* __restrict__ cbcr, int HW, const float* __restrict__ dgram,
                                 float* __restrict__ dY, long long total) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(i / HW);
        const float* dg = dgram + n * 9;
        const float x0 = Y[i], x1 = cbcr[i * 2], x2 = cbcr[i * 2 + 1];
        dY[i] += ((dg[0] + dg[0]) * x0 + (dg[1] + dg[3]) * x1 + (dg[2] + dg[6]) * x2) / (float)HW;
    }
}

// ---- SSIM (tf.image.ssim: 11x11 Gaussian sigma 1.5, VALID) on rescale_01'd images ------------------
constexpr int SS_K = 11, SS_T = 16, SS_IN = SS_T + SS_K - 1;   // 16x16 outputs from a 26x26 input tile

struct Gauss { float g[SS_K]; };
inline Gauss make_gauss() {
    Gauss r; double s = 0.0, t[SS_K];
    for (int i = 0; i < SS_K; ++i) { const double c = i - (SS_K - 1) / 2.0; t[i] = exp(-0.5 * c * c / (1.5 * 1.5)); s += t[i]; }
    for (int i = 0; i < SS_K; ++i) r.g[i] = (float)(t[i] / s);   // softmax over the 2-D grid == outer product of normalised 1-D
    return r;
}

// grid (tiles_x, tiles_y, N*3).  imgA = concat(Y,cbcr) rescaled with mmA, imgB = yuv rescaled with mmB
__global__ void __launch_bounds__(256) ssim_fwd_kernel(const float* __restrict__ Y, const float* __restrict__ cbcr, const float* __restrict__ mmA,
        const float* __restrict__ imgB, const float* __restrict__ mmB, int H, int W, float c1, float c2, Gauss gs,
        float* __restrict__ ssim_out, float* __restrict__ maps) {
    __shared__ float su[SS_IN][SS_IN + 1], sv[SS_IN][SS_IN + 1];
    __shared__ float hs[5][SS_IN][SS_T + 1];
    __shared__ double sm[32];
    const int n = blockIdx.z / 3, c = blockIdx.z % 3;
    const int HW = H * W, Ho = H - SS_K + 1, Wo = W - SS_K + 1;
    const float mnA = mmA[n * 2], dA = mmA[n * 2 + 1] - mnA, iA = dA != 0.f ? 1.f / dA : 0.f;
    const float mnB = mmB[n * 2], dB = mmB[n * 2 + 1] - mnB, iB = dB != 0.f ? 1.f / dB : 0.f;
    const int x0 = blockIdx.x * SS_T, y0 = blockIdx.y * SS_T;
    for (int i = threadIdx.x; i < SS_IN * SS_IN; i += 256) {
        const int ly = i / SS_IN, lx = i - ly * SS_IN;
        const int gy = y0 + ly, gx = x0 + lx;
        float u = 0.f, v = 0.f;
        if (gy < H && gx < W) {
            const long long p = (long long)n * HW + (long long)gy * W + gx;
            const float a = cbcr == nullptr ? Y[p * 3 + c] : (c == 0 ? Y[p] : cbcr[p * 2 + (c - 1)]);   // cbcr == NULL: A is a packed [N,H,W,3] image
            u = (a - mnA) * iA;
            v = (imgB[p * 3 + c] - mnB) * iB;
        }
        su[ly][lx] = u; sv[ly][lx] = v;
    }
    __syncthreads();
    // separable 11x11 Gaussian: horizontal pass over the 26 tile rows into shared memory, then the vertical pass per output
    // (144 FMAs per output instead of 605; the non-separable form was shared-memory-load bound at 0.2 ms per call)
    for (int i = threadIdx.x; i < SS_IN * SS_T; i += 256) {
        const int r = i / SS_T, cx = i - r * SS_T;
        float r1 = 0.f, r2 = 0.f, r11 = 0.f, r22 = 0.f, r12 = 0.f;
#pragma unroll
        for (int j = 0; j < SS_K; ++j) {
            const float u = su[r][cx + j], v = sv[r][cx + j], g = gs.g[j];
            r1 = fmaf(g, u, r1); r2 = fmaf(g, v, r2);
            r11 = fmaf(g, u * u, r11); r22 = fmaf(g, v * v, r22); r12 = fmaf(g, u * v, r12);
        }
        hs[0][r][cx] = r1; hs[1][r][cx] = r2; hs[2][r][cx] = r11; hs[3][r][cx] = r22; hs[4][r][cx] = r12;
    }
    __syncthreads();
    const int tx = threadIdx.x % SS_T, ty = threadIdx.x / SS_T;
    const int ox = x0 + tx, oy = y0 + ty;
    float S = 0.f;
    if (ox < Wo && oy < Ho) {
        float m1 = 0.f, m2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
#pragma unroll
        for (int i = 0; i < SS_K; ++i) {
            const float g = gs.g[i];
            m1 = fmaf(g, hs[0][ty + i][tx], m1); m2 = fmaf(g, hs[1][ty + i][tx], m2);
            s11 = fmaf(g, hs[2][ty + i][tx], s11); s22 = fmaf(g, hs[3][ty + i][tx], s22); s12 = fmaf(g, hs[4][ty + i][tx], s12);
        }
        const float num0 = 2.f * m1 * m2, den0 = m1 * m1 + m2 * m2;
        const float L = (num0 + c1) / (den0 + c1);
        const float A = 2.f * s12 - num0 + c2, Bc = s11 + s22 - den0 + c2;
        const float CS = A / Bc;
        S = L * CS;
        if (maps) {
            const float dL = (2.f * m2 * (den0 + c1) - (num0 + c1) * 2.f * m1) / ((den0 + c1) * (den0 + c1));
            const float dCS = (-2.f * m2 * Bc + 2.f * m1 * A) / (Bc * Bc);
            float* mp = maps + (((long long)(n * 3 + c) * Ho + oy) * Wo + ox) * 3;
            mp[0] = CS * dL + L * dCS;        // dS/dmu1
            mp[1] = -L * A / (Bc * Bc);       // dS/ds11
            mp[2] = L * 2.f / Bc;             // dS/ds12
        }
    }
    const double bs = block_sum((double)S, sm);
    if (threadIdx.x == 0) atomicAdd(&ssim_out[n], (float)(bs / (3.0 * Ho * Wo)));
}

__global__ void ssim_loss_kernel(const float* __restrict__ ssim, int N, float* __restrict__ loss_out, float weight,
                                 float* __restrict__ dssim, float gscale) {
    __shared__ double sm[32];
    float acc = 0.f;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        acc += -logf((1.f + ssim[i]) * 0.5f);
        if (dssim) dssim[i] = gscale * (-1.f / (1.f + ssim[i])) / (float)N;
    }
    const double s = block_sum((double)acc, sm);
    if (threadIdx.x == 0 && loss_out) atomicAdd(loss_out, (float)(s / N) * weight);
}

// du_c(p) = coef * sum_w g(p-w) [P(w) + 2 u(p) Q(w) + v(p) R(w)] ; dY += du_0 / (max-min) ; scratch[n] += (sum du (u-1), sum du u)
__global__ void __launch_bounds__(256) ssim_bwd_kernel(const float* __restrict__ Y, const float* __restrict__ cbcr, const float* __restrict__ mmA,
        const float* __restrict__ imgB, const float* __restrict__ mmB, int H, int W, Gauss gs, const float* __restrict__ maps,
        const float* __restrict__ dssim, float* __restrict__ dY, double* __restrict__ scratch) {
    __shared__ float sP[SS_IN][SS_IN + 1], sQ[SS_IN][SS_IN + 1], sR[SS_IN][SS_IN + 1];
    __shared__ float hs[3][SS_IN][SS_T + 1];
    __shared__ double sm[32];
    const int n = blockIdx.z / 3, c = blockIdx.z % 3;
    const int HW = H * W, Ho = H - SS_K + 1, Wo = W - SS_K + 1;
    const float mnA = mmA[n * 2], dA = mmA[n * 2 + 1] - mnA, iA = dA != 0.f ? 1.f / dA : 0.f;
    const float mnB = mmB[n * 2], dB = mmB[n * 2 + 1] - mnB, iB = dB != 0.f ? 1.f / dB : 0.f;
    const int x0 = blockIdx.x * SS_T, y0 = blockIdx.y * SS_T;
    // window origins w = p - 10 .. p  ->  tile rows y0-10 .. y0+15
    for (int i = threadIdx.x; i < SS_IN * SS_IN; i += 256) {
        const int ly = i / SS_IN, lx = i - ly * SS_IN;
        const int wy = y0 - (SS_K - 1) + ly, wx = x0 - (SS_K - 1) + lx;
        float P = 0.f, Q = 0.f, R = 0.f;
        if (wy >= 0 && wy < Ho && wx >= 0 && wx < Wo) {
            const float* mp = maps + (((long long)(n * 3 + c) * Ho + wy) * Wo + wx) * 3;
            P = mp[0]; Q = mp[1]; R = mp[2];
        }
        sP[ly][lx] = P; sQ[ly][lx] = Q; sR[ly][lx] = R;
    }
    __syncthreads();
    // separable correlation with the (symmetric) Gaussian: sum_k g(k) X(p - k) == sum_i g(i) X[tile row ty + i] by symmetry
    for (int i = threadIdx.x; i < SS_IN * SS_T; i += 256) {
        const int r = i / SS_T, cx = i - r * SS_T;
        float rP = 0.f, rQ = 0.f, rR = 0.f;
#pragma unroll
        for (int j = 0; j < SS_K; ++j) {
            const float g = gs.g[j];
            rP = fmaf(g, sP[r][cx + j], rP); rQ = fmaf(g, sQ[r][cx + j], rQ); rR = fmaf(g, sR[r][cx + j], rR);
        }
        hs[0][r][cx] = rP; hs[1][r][cx] = rQ; hs[2][r][cx] = rR;
    }
    __syncthreads();
    const int tx = threadIdx.x % SS_T, ty = threadIdx.x / SS_T;
    const int px = x0 + tx, py = y0 + ty;
    float t1 = 0.f, t2 = 0.f;
    if (px < W && py < H) {
        float aP = 0.f, aQ = 0.f, aR = 0.f;
#pragma unroll
        for (int i = 0; i < SS_K; ++i) {
            const float g = gs.g[i];
            aP = fmaf(g, hs[0][ty + i][tx], aP); aQ = fmaf(g, hs[1][ty + i][tx], aQ); aR = fmaf(g, hs[2][ty + i][tx], aR);
        }
        const long long p = (long long)n * HW + (long long)py * W + px;
        const float a = c == 0 ? Y[p] : cbcr[p * 2 + (c - 1)];
        const float u = (a - mnA) * iA, v = (imgB[p * 3 + c] - mnB) * iB;
        const float coef = dssim[n] / (3.0f * Ho * Wo);
        const float du = coef * (aP + 2.f * u * aQ + v * aR);
        if (c == 0) dY[p] += du * iA;
        t1 = du * (u - 1.f);
        t2 = du * u;
    }
    const double b1 = block_sum((double)t1, sm);
    const double b2 = block_sum((double)t2, sm);
    if (threadIdx.x == 0) { atomicAdd(&scratch[n * 2], b1); atomicAdd(&scratch[n * 2 + 1], b2); }
}

// gradient through the min / max of rescale_01: u = (x-mn)/d : du/dmn = (u-1)/d, du/dmx = -u/d
__global__ void ssim_bwd_minmax_kernel(const float* __restrict__ mmA, const int* __restrict__ idxA, int N, int HW,
                                       const double* __restrict__ scratch, float* __restrict__ dY) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float d = mmA[n * 2 + 1] - mmA[n * 2];
    if (d == 0.f) return;
    const int imn = idxA[n * 2], imx = idxA[n * 2 + 1];
    if (imn % 3 == 0) atomicAdd(&dY[(long long)n * HW + imn / 3], (float)(scratch[n * 2] / d));
    if (imx % 3 == 0) atomicAdd(&dY[(long long)n * HW + imx / 3], (float)(-scratch[n * 2 + 1] / d));
}

// ---- clip + Adam --------------------------------------------------------------------------------
__global__ void clip_adam_kernel(float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ m, float* __restrict__ v,
                                 long long n, float lr_t, const float* __restrict__ lr_dev, float b1, float b2, float eps, float clip, float gscale) {
    if (lr_dev != nullptr) lr_t = __ldg(lr_dev);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float g = grad[i] * gscale;
        g = fminf(fmaxf(g, -clip), clip);
        const float mi = b1 * m[i] + (1.f - b1) * g;
        const float vi = b2 * v[i] + (1.f - b2) * g * g;
        m[i] = mi; v[i] = vi;
        param[i] = param[i] - lr_t * mi / (sqrtf(vi) + eps);
    }
}
__global__ void clip_adam_vec_kernel(float4* __restrict__ param, const float4* __restrict__ grad, float4* __restrict__ m, float4* __restrict__ v,
                                     long long n4, float lr_t, const float* __restrict__ lr_dev, float b1, float b2, float eps, float clip, float gscale) {
    if (lr_dev != nullptr) lr_t = __ldg(lr_dev);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 g4 = __ldg(grad + i);
        float4 p4 = param[i], m4 = m[i], v4 = v[i];
        float* pp = &p4.x; float* pm = &m4.x; float* pv = &v4.x; const float* pg = &g4.x;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float g = pg[j] * gscale;
            g = fminf(fmaxf(g, -clip), clip);
            pm[j] = b1 * pm[j] + (1.f - b1) * g;
            pv[j] = b2 * pv[j] + (1.f - b2) * g * g;
            pp[j] = pp[j] - lr_t * pm[j] / (sqrtf(pv[j]) + eps);
        }
        param[i] = p4; m[i] = m4; v[i] = v4;
    }
}

// ---- Philox-4x32-10 --------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox(uint64_t seed, uint64_t ctr) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = 0x5851F42Du, c3 = 0x4C957F2Du;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = h1 ^ c1 ^ k0, n1 = l1, n2 = h0 ^ c3 ^ k1, n3 = l0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }

template <typename T, bool NORMAL>
__global__ void rng_kernel(T* __restrict__ out, long long n, uint64_t seed, uint64_t offset, const uint64_t* __restrict__ offset_dev, float param) {
    if (offset_dev != nullptr) offset = __ldg(reinterpret_cast<const unsigned long long*>(offset_dev));
    const long long n4 = (n + 3) / 4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const uint4 r = philox(seed, offset + (uint64_t)i);
        float v[4];
        if (NORMAL) {
            const float r0 = sqrtf(-2.f * logf(u01(r.x))), r1 = sqrtf(-2.f * logf(u01(r.z)));
            float s0, c0, s1, c1;
            sincosf(6.283185307179586f * u01(r.y), &s0, &c0);
            sincosf(6.283185307179586f * u01(r.w), &s1, &c1);
            v[0] = r0 * c0 * param; v[1] = r0 * s0 * param; v[2] = r1 * c1 * param; v[3] = r1 * s1 * param;
        } else {
            v[0] = u01(r.x) < param ? 1.f : 0.f; v[1] = u01(r.y) < param ? 1.f : 0.f;
            v[2] = u01(r.z) < param ? 1.f : 0.f; v[3] = u01(r.w) < param ? 1.f : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) if (i * 4 + j < n) stf(out + i * 4 + j, v[j]);
    }
}

}  // namespace

extern "C" int shm_lsgan(const float* a, int64_t n, float target, float* loss_out, float weight, float* da, float gscale, int accumulate, void* stream) {
    SHM_REQUIRE(a && n > 0, "shm_lsgan: bad args");
    ew_loss_kernel<0><<<flat_grid(n), 256, 0, (cudaStream_t)stream>>>(a, nullptr, target, n, loss_out, weight, da, gscale, accumulate);
    SHM_CHECK_LAUNCH("lsgan_kernel");
    return SHM_OK;
}
extern "C" int shm_l1(const float* a, const float* b, int64_t n, float* loss_out, float weight, float* da, float gscale, int accumulate, void* stream) {
    SHM_REQUIRE(a && b && n > 0, "shm_l1: bad args");
    ew_loss_kernel<1><<<flat_grid(n), 256, 0, (cudaStream_t)stream>>>(a, b, 0.f, n, loss_out, weight, da, gscale, accumulate);
    SHM_CHECK_LAUNCH("l1_kernel");
    return SHM_OK;
}
extern "C" int shm_mse(const float* a, const float* b, int64_t n, float* loss_out, float weight, float* da, float gscale, int accumulate, void* stream) {
    SHM_REQUIRE(a && b && n > 0, "shm_mse: bad args");
    ew_loss_kernel<2><<<flat_grid(n), 256, 0, (cudaStream_t)stream>>>(a, b, 0.f, n, loss_out, weight, da, gscale, accumulate);
    SHM_CHECK_LAUNCH("mse_kernel");
    return SHM_OK;
}
extern "C" int shm_softmax_ce(const float* logits, int B, const float labels[5], float* loss_out, float weight, float* dlogits,
                              float gscale, int accumulate, void* stream) {
    SHM_REQUIRE(logits && labels && B > 0, "shm_softmax_ce: bad args");
    Labels5 l;
    for (int j = 0; j < 5; ++j) l.v[j] = labels[j];
    softmax_ce_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(logits, B, l, nullptr, loss_out, weight, dlogits, gscale, accumulate);
    SHM_CHECK_LAUNCH("softmax_ce_kernel");
    return SHM_OK;
}
// target / labels read from device memory when the kernel runs: TARGET_LABELS is redrawn every step (ShmGANwithSSpecSeg.py:986) and a step
// captured in a CUDA graph must see the current value
extern "C" int shm_lsgan_dev(const float* a, int64_t n, const float* target_dev, float* loss_out, float weight, float* da, float gscale, int accumulate, void* stream) {
    SHM_REQUIRE(a && target_dev && n > 0, "shm_lsgan_dev: bad args");
    ew_loss_kernel<0><<<flat_grid(n), 256, 0, (cudaStream_t)stream>>>(a, target_dev, 0.f, n, loss_out, weight, da, gscale, accumulate);
    SHM_CHECK_LAUNCH("lsgan_kernel");
    return SHM_OK;
}
extern "C" int shm_softmax_ce_dev(const float* logits, int B, const float* labels5_dev, float* loss_out, float weight, float* dlogits,
                                  float gscale, int accumulate, void* stream) {
    SHM_REQUIRE(logits && labels5_dev && B > 0, "shm_softmax_ce_dev: bad args");
    Labels5 l{};
    softmax_ce_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(logits, B, l, labels5_dev, loss_out, weight, dlogits, gscale, accumulate);
    SHM_CHECK_LAUNCH("softmax_ce_kernel");
    return SHM_OK;
}
extern "C" int shm_mse_ycc(const float* Y, const float* cbcr, const float* yuv, int64_t npix, float* loss_out, float weight,
                           float* dY, float gscale, void* stream) {
    SHM_REQUIRE(Y && cbcr && yuv && npix > 0, "shm_mse_ycc: bad args");
    mse_ycc_kernel<<<flat_grid(npix), 256, 0, (cudaStream_t)stream>>>(Y, cbcr, yuv, npix, loss_out, weight, dY, gscale);
    SHM_CHECK_LAUNCH("mse_ycc_kernel");
    return SHM_OK;
}
extern "C" int shm_spec_loss(const float* Y, const float* cbcr, const float* yuv, const float* mask, int64_t npix, float* loss_out,
                             float weight, void* stream) {
    SHM_REQUIRE(Y && cbcr && yuv && mask && loss_out && npix > 0, "shm_spec_loss: bad args");
    spec_kernel<<<flat_grid(npix), 256, 0, (cudaStream_t)stream>>>(Y, cbcr, yuv, mask, npix, loss_out, weight);
    SHM_CHECK_LAUNCH("spec_kernel");
    return SHM_OK;
}
extern "C" int shm_minmax3(const float* Y_or_yuv, const float* cbcr, int N, int HW, float* mm, int32_t* idx, void* stream) {
    SHM_REQUIRE(Y_or_yuv && mm && idx && N > 0 && HW > 0, "shm_minmax3: bad args");
    SHM_REQUIRE((long long)HW * 3 < 0x7fffffffLL, "shm_minmax3: image too large");
    minmax3_kernel<<<N * MM_CLUSTER, 1024, 0, (cudaStream_t)stream>>>(Y_or_yuv, cbcr, HW, mm, idx);
    SHM_CHECK_LAUNCH("minmax3_kernel");
    return SHM_OK;
}
extern "C" int shm_gram3(const float* Y_or_yuv, const float* cbcr, int N, int HW, double* gram, void* stream) {
    SHM_REQUIRE(Y_or_yuv && gram && N > 0 && HW > 0, "shm_gram3: bad args");
    int ppb = cdiv(HW, cdiv(shm_num_sms() * 4, N));
    if (ppb < 1024) ppb = 1024;
    if (ppb > 65536) ppb = 65536;
    gram3_kernel<<<dim3(cdiv(HW, ppb), N), 256, 0, (cudaStream_t)stream>>>(Y_or_yuv, cbcr, HW, gram, ppb);
    SHM_CHECK_LAUNCH("gram3_kernel");
    return SHM_OK;
}
extern "C" int shm_style_loss(const double* gramA, const double* gramB, int N, int HW, int S, float* loss_out, float weight,
                              float* dgramA, float gscale, void* stream) {
    SHM_REQUIRE(gramA && gramB && N > 0 && HW > 0 && S > 0, "shm_style_loss: bad args");
    const double f = 1.0 / ((2.0 * 9.0 * S * S) * (2.0 * 9.0 * S * S));     // :817 factor = 1/(2*filter_count*size)^2
    style_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(gramA, gramB, N, HW, f, loss_out, weight, dgramA, gscale);
    SHM_CHECK_LAUNCH("style_kernel");
    return SHM_OK;
}
extern "C" int shm_gram3_bwd(const float* Y, const float* cbcr, int N, int HW, const float* dgram, float* dY, void* stream) {
    SHM_REQUIRE(Y && cbcr && dgram && dY && N > 0 && HW > 0, "shm_gram3_bwd: bad args");
    const long long total = (long long)N * HW;
    gram3_bwd_kernel<<<flat_grid(total), 256, 0, (cudaStream_t)stream>>>(Y, cbcr, HW, dgram, dY, total);
    SHM_CHECK_LAUNCH("gram3_bwd_kernel");
    return SHM_OK;
}
extern "C" int64_t shm_ssim_map_elems(int N, int H, int W) {
    if (H < SS_K || W < SS_K) return 0;
    return (int64_t)N * 3 * (H - SS_K + 1) * (W - SS_K + 1) * 3;
}
extern "C" int shm_ssim_fwd(const float* Y, const float* cbcr, const float* mmA, const float* imgB, const float* mmB, int N, int H, int W,
                            float max_val, float* ssim_out, float* maps, void* stream) {
    SHM_REQUIRE(Y && mmA && imgB && mmB && ssim_out && N > 0, "shm_ssim_fwd: bad args");
    SHM_REQUIRE(cbcr || !maps, "shm_ssim_fwd: the packed-image form (cbcr == NULL) is value-only");
    SHM_REQUIRE(H >= SS_K && W >= SS_K, "shm_ssim_fwd: image smaller than the 11x11 window");
    const float c1 = (0.01f * max_val) * (0.01f * max_val), c2 = (0.03f * max_val) * (0.03f * max_val);
    cudaStream_t st = (cudaStream_t)stream;
    if (cudaMemsetAsync(ssim_out, 0, sizeof(float) * N, st) != cudaSuccess) SHM_FAIL(SHM_ECUDA, "shm_ssim_fwd: memset failed");
    dim3 grid(cdiv(W - SS_K + 1, SS_T), cdiv(H - SS_K + 1, SS_T), N * 3);
    ssim_fwd_kernel<<<grid, 256, 0, st>>>(Y, cbcr, mmA, imgB, mmB, H, W, c1, c2, make_gauss(), ssim_out, maps);
    SHM_CHECK_LAUNCH("ssim_fwd_kernel");
    return SHM_OK;
}
extern "C" int shm_ssim_loss(const float* ssim, int N, float* loss_out, float weight, float* dssim, float gscale, void* stream) {
    SHM_REQUIRE(ssim && N > 0, "shm_ssim_loss: bad args");
    ssim_loss_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(ssim, N, loss_out, weight, dssim, gscale);
    SHM_CHECK_LAUNCH("ssim_loss_kernel");
    return SHM_OK;
}
extern "C" int shm_ssim_bwd(const float* Y, const float* cbcr, const float* mmA, const int32_t* idxA, const float* imgB, const float* mmB,
                            int N, int H, int W, const float* maps, const float* dssim, float* dY, double* scratch, void* stream) {
    SHM_REQUIRE(Y && cbcr && mmA && idxA && imgB && mmB && maps && dssim && dY && scratch && N > 0, "shm_ssim_bwd: bad args");
    SHM_REQUIRE(H >= SS_K && W >= SS_K, "shm_ssim_bwd: image smaller than the 11x11 window");
    cudaStream_t st = (cudaStream_t)stream;
    if (cudaMemsetAsync(scratch, 0, sizeof(double) * 2 * N, st) != cudaSuccess) SHM_FAIL(SHM_ECUDA, "shm_ssim_bwd: memset failed");
    dim3 grid(cdiv(W, SS_T), cdiv(H, SS_T), N * 3);
    ssim_bwd_kernel<<<grid, 256, 0, st>>>(Y, cbcr, mmA, imgB, mmB, H, W, make_gauss(), maps, dssim, dY, scratch);
    SHM_CHECK_LAUNCH("ssim_bwd_kernel");
    ssim_bwd_minmax_kernel<<<cdiv(N, 64), 64, 0, st>>>(mmA, idxA, N, H * W, scratch, dY);
    SHM_CHECK_LAUNCH("ssim_bwd_minmax_kernel");
    return SHM_OK;
}

namespace {
int clip_adam_launch(float* param, const float* grad, float* m, float* v, int64_t n, float lr_t, const float* lr_dev, float beta1, float beta2,
                     float eps, float clip, float gscale, void* stream) {
    if (n == 0) return SHM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const uintptr_t al = reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v);
    const long long n4 = ((al & 15) == 0) ? n / 4 : 0;
    if (n4 > 0) {
        clip_adam_vec_kernel<<<flat_grid(n4), 256, 0, st>>>((float4*)param, (const float4*)grad, (float4*)m, (float4*)v, n4, lr_t, lr_dev, beta1, beta2, eps, clip, gscale);
        SHM_CHECK_LAUNCH("clip_adam_vec_kernel");
    }
    if (n4 * 4 < n) {
        const long long off = n4 * 4;
        clip_adam_kernel<<<flat_grid(n - off), 256, 0, st>>>(param + off, grad + off, m + off, v + off, n - off, lr_t, lr_dev, beta1, beta2, eps, clip, gscale);
        SHM_CHECK_LAUNCH("clip_adam_kernel");
    }
    return SHM_OK;
}
}  // namespace

extern "C" int shm_clip_adam(float* param, const float* grad, float* m, float* v, int64_t n, float lr_t, float beta1, float beta2,
                             float eps, float clip, float gscale, void* stream) {
    SHM_REQUIRE(param && grad && m && v && n >= 0, "shm_clip_adam: bad args");
    return clip_adam_launch(param, grad, m, v, n, lr_t, nullptr, beta1, beta2, eps, clip, gscale, stream);
}
// the step size read from device memory at run time: a step captured in a CUDA graph replays with the current bias-corrected learning rate
extern "C" int shm_clip_adam_dev(float* param, const float* grad, float* m, float* v, int64_t n, const float* lr_t_dev, float beta1, float beta2,
                                 float eps, float clip, float gscale, void* stream) {
    SHM_REQUIRE(param && grad && m && v && lr_t_dev && n >= 0, "shm_clip_adam_dev: bad args");
    return clip_adam_launch(param, grad, m, v, n, 0.f, lr_t_dev, beta1, beta2, eps, clip, gscale, stream);
}

namespace {
int rng_launch(void* out, int64_t n, uint64_t seed, uint64_t offset, const uint64_t* offset_dev, float param, int dtype, bool normal, void* stream) {
    if (n == 0) return SHM_OK;
    DISPATCH_DTYPE(dtype, T, {
        if (normal) rng_kernel<T, true><<<flat_grid((n + 3) / 4), 256, 0, (cudaStream_t)stream>>>((T*)out, n, seed, offset, offset_dev, param);
        else        rng_kernel<T, false><<<flat_grid((n + 3) / 4), 256, 0, (cudaStream_t)stream>>>((T*)out, n, seed, offset, offset_dev, param);
        SHM_CHECK_LAUNCH("rng_kernel");
        return SHM_OK;
    })
}
}  // namespace

extern "C" int shm_rng_normal(void* out, int64_t n, uint64_t seed, uint64_t offset, float sigma, int dtype, void* stream) {
    SHM_REQUIRE(out && n >= 0, "shm_rng_normal: bad args");
    return rng_launch(out, n, seed, offset, nullptr, sigma, dtype, true, stream);
}
extern "C" int shm_rng_keep(void* out, int64_t n, uint64_t seed, uint64_t offset, float keep_prob, int dtype, void* stream) {
    SHM_REQUIRE(out && n >= 0, "shm_rng_keep: bad args");
    return rng_launch(out, n, seed, offset, nullptr, keep_prob, dtype, false, stream);
}
// the Philox counter offset read from device memory at run time (CUDA-graph replays draw a fresh stream every step)
extern "C" int shm_rng_normal_dev(void* out, int64_t n, uint64_t seed, const uint64_t* offset_dev, float sigma, int dtype, void* stream) {
    SHM_REQUIRE(out && offset_dev && n >= 0, "shm_rng_normal_dev: bad args");
    return rng_launch(out, n, seed, 0, offset_dev, sigma, dtype, true, stream);
}
extern "C" int shm_rng_keep_dev(void* out, int64_t n, uint64_t seed, const uint64_t* offset_dev, float keep_prob, int dtype, void* stream) {
    SHM_REQUIRE(out && offset_dev && n >= 0, "shm_rng_keep_dev: bad args");
    return rng_launch(out, n, seed, 0, offset_dev, keep_prob, dtype, false, stream);
}
