// umma_rate_probe.cu -- hardware probe: issue-to-completion rate of back-to-back tcgen05.mma (kind::f16, bf16 x bf16 -> fp32, K = 16 per
// instruction, both operands from shared memory with SWIZZLE_128B K-major descriptors) for
//     cta_group::1   M = 128, N in {64, 128, 256}          (one SM; what every kernel of csrc/conv_tc.cu issues today)
//     cta_group::2   M = 256, N in {64, 128, 256}          (a CTA pair: each SM holds its 128 rows of A and HALF of B)
// The question (profiles/r01_negative_results.txt #4): N = 64 tiles run at ~62 cycles per M128 x N64 x K16 MMA against a
// 32-cycle tensor floor because every MMA pulls 4 KB (A) + 2 KB (B) through the 128 B/clk shared-memory port.  With cta_group::2 a pair
// reads 4 KB (A) + 1 KB (half of B) per SM per MMA: does the per-SM cost per 128 x 64 x 16 block drop, and by how much?
// Operand contents are irrelevant (uninitialised shared memory); only timing is measured: clock64 around `nmma` MMAs + commit + wait.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o umma_rate_probe umma_rate_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0, tries = 0;
    while (!ok) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (++tries > (1u << 22)) __trap();
    }
}
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

constexpr int A_BYTES = 128 * 128;       // 128 rows x 64 bf16 per CTA
constexpr int NBUF = 4;                  // operand buffers cycled through (a real main loop walks a ring)

template <int CG>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if (CG == 1)
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// CG = 1: grid of independent CTAs.  CG = 2: clusters of two CTAs, the leader (rank 0) issues for the pair.
template <int CG, int N>
__global__ void __launch_bounds__(128, 1) rate_kernel(int nmma, long long* cycles) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;                                   // NBUF x 16 KB
    uint8_t* sB = smem + NBUF * A_BYTES;                  // NBUF x (N / CG) rows x 128 B
    constexpr int B_BYTES = (N / CG) * 128;
    uint64_t* bar = (uint64_t*)(sB + NBUF * B_BYTES);
    uint32_t* slot = (uint32_t*)(bar + 1);
    const int warp = threadIdx.x >> 5;
    uint32_t rank = 0;
    if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        if (CG == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(N < 32 ? 32 : N) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(N < 32 ? 32 : N) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (CG == 2) { asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory"); } else __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    long long t0 = 0, t1 = 0;
    if (threadIdx.x == 0 && rank == 0) {
        constexpr uint32_t idesc = make_idesc(128 * CG, N);
        t0 = clock64();
        for (int i = 0; i < nmma; ++i) {
            const int buf = (i >> 2) % NBUF;                               // 4 x (K = 16) per 64-channel operand tile, then the next buffer
            const uint64_t ad = make_desc_sw128(smem_u32(sA + buf * A_BYTES)) + (uint64_t)((i & 3) * 2);
            const uint64_t bd = make_desc_sw128(smem_u32(sB + buf * B_BYTES)) + (uint64_t)((i & 3) * 2);
            mma<CG>(tmem, ad, bd, idesc, i != 0);
        }
        if (CG == 1)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
        else
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
    }
    mbar_wait(bar, 0);                                                     // both CTAs of a pair wait on their own copy of the barrier
    if (threadIdx.x == 0 && rank == 0) { t1 = clock64(); cycles[blockIdx.x / CG] = t1 - t0; }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (CG == 2) { asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory"); } else __syncthreads();
    if (warp == 0) {
        if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(N < 32 ? 32 : N) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(N < 32 ? 32 : N) : "memory");
    }
}

template <int CG, int N>
void run(int nmma, int ctas, long long* dcyc) {
    const int smem = NBUF * A_BYTES + NBUF * (N / CG) * 128 + 1024 + 64;
    cudaFuncSetAttribute(rate_kernel<CG, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    for (int rep = 0; rep < 2; ++rep) {
        cudaError_t e = cudaLaunchKernelEx(&cfg, rate_kernel<CG, N>, nmma, dcyc);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("cta_group::%d N=%d: CUDA error %s\n", CG, N, cudaGetErrorString(e)); return; }
    }
    long long h[256];
    cudaMemcpy(h, dcyc, sizeof(long long) * (ctas / CG), cudaMemcpyDeviceToHost);
    long long mx = 0; double avg = 0;
    for (int i = 0; i < ctas / CG; ++i) { if (h[i] > mx) mx = h[i]; avg += (double)h[i]; }
    avg /= (ctas / CG);
    const double per = avg / nmma;                        // cycles per MMA instruction (the pair's instruction covers M = 256)
    const double flop_per_clk_sm = 2.0 * 128 * N * 16 / per;   // per SM: each SM of a pair computes its own 128 x N x 16 block
    printf("cta_group::%d  M=%3d N=%3d  %5d MMAs on %3d SMs: %7.1f cycles/MMA (max %7.1f)  -> %6.0f FLOP/clk/SM  (%4.1f%% of 8192)  operand bytes/SM/MMA %5d\n",
           CG, 128 * CG, N, nmma, ctas, per, (double)mx / nmma, flop_per_clk_sm, 100.0 * flop_per_clk_sm / 8192.0, 4096 + (N / CG) * 32);
}

int main() {
    long long* dcyc;
    cudaMalloc(&dcyc, sizeof(long long) * 256);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int ctas = sms & ~1;
    for (int nmma : {256, 2048}) {
        run<1, 64>(nmma, ctas, dcyc);
        run<2, 64>(nmma, ctas, dcyc);
        run<1, 128>(nmma, ctas, dcyc);
        run<2, 128>(nmma, ctas, dcyc);
        run<1, 256>(nmma, ctas, dcyc);
        run<2, 256>(nmma, ctas, dcyc);
    }
    // one SM (pair) alone: no chip-level effects (power, L2) -- the pure per-SM pipeline rate
    run<1, 64>(2048, 2, dcyc);
    run<2, 64>(2048, 2, dcyc);
    return 0;
}
