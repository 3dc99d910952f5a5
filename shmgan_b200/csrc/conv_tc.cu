// conv_tc.cu -- bf16 implicit-GEMM convolutions on the 5th-gen tensor cores (tcgen05.mma, fp32 accumulators in TMEM),
// operands staged in shared memory by TMA with 128-byte swizzle, one persistent warp-specialised CTA per SM.
//
//   D[m, n] = sum_{tap} sum_{k} A[lattice(m) * IS + off_tap, k] * Wt[tap][n][k]
//
//   m  : 128 lattice points per tile = a BI x BH x BW box of (image, row, col) -- loaded per tap by ONE 4-D TMA box whose
//        out-of-range coordinates are zero-filled by the hardware, which IS the TF 'SAME' padding
//   n  : output channels (BN = 64 or 128 per tile), k : reduction channels in chunks of 64 (= one 128-byte swizzle row)
//
// Replaces the cuDNN kernels TF dispatches for Conv2D / Conv2DTranspose forward and dgrad
// (ShmGANwithSSpecSeg.py:244-326, :365, :387, :410-411; tape.gradient :859,:868).
#include "common.cuh"
#include <string.h>
#include <cuda.h>
#include <stdlib.h>
#include <type_traits>

namespace {

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok;
}
// try_wait suspends in hardware for a bounded time per attempt; the attempt cap turns a pipeline deadlock into a trap
// (reported as a launch failure by the next CUDA call) instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t tries = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++tries > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// the same box into L2 only: no shared-memory destination, no barrier -- hides the HBM part of a later tma_load_4d's latency when the kernel
// cannot afford another shared-memory stage
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* tm, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
                 ::"l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* tm) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}
// one lane of the (converged) warp: the tcgen05 issue loops run warp-uniformly and only the MMA / commit instructions are guarded by
// this predicate -- guarding the whole loop with `lane == 0` made ptxas wrap every UTCHMMA in a per-lane ELECT / R2UR.BROADCAST /
// BRA.U.ANY loop (~90 issue cycles per MMA against 32-64 tensor cycles: every kernel was issue-bound)
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred = 0;
    asm volatile(
        "{\n"
        ".reg .pred px;\n"
        "elect.sync _|px, 0xffffffff;\n"
        "@px mov.s32 %0, 1;\n"
        "}\n" : "+r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// ---- CTA-pair (cta_group::2) forms: a cluster of two CTAs on one TPC; the leader (cluster rank 0) issues one MMA for both ----
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (a shared::cta pointer of this CTA) as seen in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(const void* p, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads whose completion is signalled on a barrier that may live in the PEER CTA (the leader's full barrier)
__device__ __forceinline__ void tma_load_4d_2sm(void* dst, const CUtensorMap* tm, uint32_t bar_cluster_addr, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(tm), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* tm, uint32_t bar_cluster_addr, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(tm), "r"(bar_cluster_addr), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrives (once all MMAs issued so far have completed) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp reads TMEM lane (base_lane + i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}

// smem matrix descriptor, 128-byte swizzle, rows of 128 bytes, 8-row groups 1024 bytes apart (K-major operand)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;      // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;      // SWIZZLE_128B
    return d;
}
// same with an explicit swizzle mode for rows of 128 / 64 / 32 bytes (descriptor layout-type codes 2 / 4 / 6)
__device__ __forceinline__ uint64_t make_desc_swz(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_code) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout_code << 61;
    return d;
}
// instruction descriptor: bf16 x bf16 -> fp32, M x N, A/B major bits
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}


// 32 lanes x 32 columns WITHOUT the wait: the registers are valid only after tmem_wait_ld32 on the same array
__device__ __forceinline__ void tmem_ld32_nw(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
// waits for every outstanding tcgen05.ld of this thread; the array is an in/out operand so that no consumer of r is scheduled
// above the wait
__device__ __forceinline__ void tmem_wait_ld32(uint32_t* r) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
        : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
          "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
          "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
          "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
        :: "memory");
}

// 32 lanes x 16 columns (thin layers with 16 output channels), without / with the wait
__device__ __forceinline__ void tmem_ld16_nw(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld16(uint32_t* r) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
        : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
          "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
        :: "memory");
}
// bias + activation + bf16 of NV accumulator columns -> NV / 8 packed 16-byte chunks (registers)
template <int NV>
__device__ __forceinline__ void epi_pack(const uint32_t* r, const float* sbias, int act, uint4* out) {
    float v[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = __uint_as_float(r[j]);
    if (sbias != nullptr) {
        const float4* b4 = reinterpret_cast<const float4*>(sbias);
#pragma unroll
        for (int j = 0; j < NV / 4; ++j) {
            const float4 b = b4[j];
            v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
        }
    }
    if (act == SHM_ACT_LRELU) {
#pragma unroll
        for (int j = 0; j < NV; ++j) v[j] = fmaxf(v[j], 0.2f * v[j]);
    } else if (act == SHM_ACT_RELU) {
#pragma unroll
        for (int j = 0; j < NV; ++j) v[j] = fmaxf(v[j], 0.f);
    } else if (act == SHM_ACT_SIGMOID) {
#pragma unroll
        for (int j = 0; j < NV; ++j) v[j] = 1.f / (1.f + __expf(-v[j]));
    }
#pragma unroll
    for (int j = 0; j < NV; j += 8) {
        __nv_bfloat162 h0 = __floats2bfloat162_rn(v[j], v[j + 1]), h1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
        __nv_bfloat162 h2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]), h3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
        out[j / 8] = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                                *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
    }
}

// bias (shared memory, broadcast reads) + activation + bf16 + 64 contiguous bytes for 32 accumulator columns in registers
__device__ __forceinline__ void epi_store32(const uint32_t* r, const float* sbias, int act, bf16* __restrict__ dst, bool ok, int ncols = 32) {
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
    if (sbias != nullptr) {
        const float4* b4 = reinterpret_cast<const float4*>(sbias);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 b = b4[j];
            v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
        }
    }
    if (act == SHM_ACT_LRELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.2f * v[j]);
    } else if (act == SHM_ACT_RELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
    } else if (act == SHM_ACT_SIGMOID) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 1.f / (1.f + __expf(-v[j]));
    }
    if (ok) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
            if (j >= ncols) break;                       // partial store: only the first ncols channels of this chunk exist in dst
            uint4 u;
            __nv_bfloat162 h0 = __floats2bfloat162_rn(v[j], v[j + 1]), h1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
            __nv_bfloat162 h2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]), h3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
            u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
            u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
            *reinterpret_cast<uint4*>(dst + j) = u;
        }
    }
}

// Epilogue of one accumulator row of NC x 32 columns.  The TMEM load of chunk c+1 is in flight while chunk c is converted and
// stored: the first version loaded, waited and processed chunk by chunk, and ncu's source page put 45 % of the epilogue
// warps' samples on the first use of the tcgen05.ld result (the load competes with the MMAs for TMEM bandwidth), which made
// every 64-channel layer epilogue-bound (3300 cycles per 128 x 64 tile against 1360 tensor cycles).
template <int NC>
__device__ __forceinline__ void epi_row(uint32_t taddr, const float* sbias, int act, bf16* __restrict__ dst, bool ok, int ncols = NC * 32) {
    uint32_t r[2][32];
    tmem_ld32_nw(taddr, r[0]);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        tmem_wait_ld32(r[c & 1]);
        if (c + 1 < NC) tmem_ld32_nw(taddr + (uint32_t)((c + 1) * 32), r[(c + 1) & 1]);
        epi_store32(r[c & 1], sbias ? sbias + c * 32 : nullptr, act, dst + c * 32, ok, ncols - c * 32);
    }
}

// ------------------------------------------------------------------------------------------------
// Instance-norm statistics produced by the convolution epilogue (SURVEY 2b "IN stats fused into the producing conv's epilogue").
//
// Every Conv -> LeakyReLU -> InstanceNorm block of the reference (ShmGANwithSSpecSeg.py:244-245, :386-389) needs sum / sum of squares per
// (image, channel) of the tensor the epilogue is about to store; computing them here removes one full HBM read of that tensor (in_stats_p).
// A thread owns one pixel ROW of the accumulator, so per-channel sums are COLUMN sums across the lanes of a warp.  Reducing across lanes per
// tile (a recursive-halving shuffle butterfly, 31 shuffles + ~90 ALU instructions per 32 x 32 block and moment) made every epilogue the
// bottleneck of its kernel: measured +0.85 ms on the halo kernels, +0.78 ms on the big-halo kernel, +0.37 ms on the generic kernel per training
// step against 1.65 ms of in_stats_p saved (profiles/r02_fused_stats_ab.txt) -- a net loss.  What does pay: every thread keeps PRIVATE fp32
// running sums of its own pixel row over the consecutive tiles its CTA handles of one image (the tile order of a statistics launch is
// contiguous per CTA for that reason; 2 FMA-class instructions per value), and the cross-lane butterfly + one fp64 atomic per (image, channel,
// moment) run only when the image changes -- two or three times per CTA.  That costs 2 x BN registers per thread, so it exists for the
// 64-column halo variants only (the full-resolution 64-channel tensors: the largest share of the statistics traffic); every other layer keeps
// the separate pass.  The statistics are taken from the bf16-ROUNDED values, exactly what in_stats_p would read back.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_colsum32(float* v, int lane) {
#pragma unroll
    for (int h = 16; h >= 1; h >>= 1) {
        const bool up = (lane & h) != 0;
#pragma unroll
        for (int j = 0; j < h; ++j) {
            const float send = up ? v[j] : v[j + h];
            const float keep = up ? v[j + h] : v[j];
            v[j] = keep + __shfl_xor_sync(0xffffffffu, send, h);
        }
    }
    return v[0];
}

template <int BN>                      // accumulator columns of one pixel row (a multiple of 32)
struct EpiStats {
    float s[BN], q[BN];                // this thread's running sum / sum of squares per column, over the tiles of image `img`
    int img;
    __device__ __forceinline__ void init() {
        img = -1;
#pragma unroll
        for (int c = 0; c < BN; ++c) { s[c] = 0.f; q[c] = 0.f; }
    }
    // cross-lane column sums of the running totals -> lane L owns columns L, 32 + L, ...; one fp64 atomic per column and moment
    __device__ __forceinline__ void flush(double* __restrict__ stats, int lane) {
        if (img >= 0) {
#pragma unroll
            for (int c = 0; c < BN / 32; ++c) {
                const float ts = warp_colsum32(s + c * 32, lane);
                const float tq = warp_colsum32(q + c * 32, lane);
                double* d = stats + ((long long)img * BN + c * 32 + lane) * 2;
                atomicAdd(d, (double)ts);
                atomicAdd(d + 1, (double)tq);
            }
#pragma unroll
            for (int c = 0; c < BN; ++c) { s[c] = 0.f; q[c] = 0.f; }
        }
    }
    // warp-uniform: the running sums now belong to image nimg
    __device__ __forceinline__ void own(double* __restrict__ stats, int nimg, int lane) {
        if (nimg != img) { flush(stats, lane); img = nimg; }
    }
    // pk: the packed bf16 of this lane's pixel row, columns [col0, col0 + 8 * NU)
    template <int NU>
    __device__ __forceinline__ void add(int col0, const uint4* pk) {
#pragma unroll
        for (int i = 0; i < NU; ++i) {
            const uint32_t u[4] = {pk[i].x, pk[i].y, pk[i].z, pk[i].w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float lo = __uint_as_float(u[k] << 16), hi = __uint_as_float(u[k] & 0xffff0000u);
                const int c = col0 + i * 8 + 2 * k;
                s[c] += lo; q[c] = fmaf(lo, lo, q[c]);
                s[c + 1] += hi; q[c + 1] = fmaf(hi, hi, q[c + 1]);
            }
        }
    }
};
struct NoStats {                       // stand-in for the kernel variants without a statistics epilogue (no registers)
    __device__ __forceinline__ void init() {}
    __device__ __forceinline__ void flush(double*, int) {}
    __device__ __forceinline__ void own(double*, int, int) {}
    template <int NU> __device__ __forceinline__ void add(int, const uint4*) {}
};

// contiguous share of `total` work items for this CTA (statistics launches), or the strided order (everything else)
__device__ __forceinline__ void tile_walk(int total, bool contig, int& first, int& end, int& step) {
    if (contig) {
        first = (int)((long long)total * blockIdx.x / gridDim.x);
        end = (int)((long long)total * (blockIdx.x + 1) / gridDim.x);
        step = 1;
    } else { first = blockIdx.x; end = total; step = gridDim.x; }
}

// copies the layer's bias vector into shared memory (all threads of the CTA; call before the first __syncthreads)
__device__ __forceinline__ void stage_bias(float* sbias, const float* __restrict__ bias, int n) {
    if (bias != nullptr)
        for (int i = threadIdx.x; i < n; i += blockDim.x) sbias[i] = __ldg(bias + i);
}
constexpr int BIAS_SMEM = 4096;          // up to 1024 output channels

// ------------------------------------------------------------------------------------------------
// forward / dgrad kernel
// ------------------------------------------------------------------------------------------------
struct TcParams {
    int ntaps; int dy[9], dx[9], wrow[9];
    int kchunks;                 // K / 64
    int Qh, Qw, BW, BH, BI;      // lattice size and the tile box (BI*BH*BW == 128)
    int tiles_x, tiles_y;        // tiles per image row / column
    int m_tiles, n_tiles;
    int IS;                      // input traversal stride
    int Hout, Wout, OS, py, px, ldout, Nn;
    const float* bias; int act;
    bf16* out;
};

constexpr int TC_THREADS = 192;          // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2..5: epilogue
constexpr int A_BYTES = 128 * 128;       // 128 rows x 64 bf16

template <int BN>
struct TcCfg {
    static constexpr int B_BYTES = BN * 128;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (BN == 64) ? 8 : 6;
    static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + BIAS_SMEM;
    static constexpr int TMEM_COLS = 2 * BN;
};

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
    using Cfg = TcCfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = smem + Cfg::STAGES * A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
    uint64_t* full = bars;                          // [STAGES]
    uint64_t* empty = bars + Cfg::STAGES;           // [STAGES]
    uint64_t* tfull = bars + 2 * Cfg::STAGES;       // [2]
    uint64_t* tempty = tfull + 2;                   // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    float* sbias = reinterpret_cast<float*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = p.m_tiles * p.n_tiles;
    const int niter = p.ntaps * p.kchunks;
    int tile_first, tile_end, tile_step;
    tile_walk(total_tiles, false, tile_first, tile_end, tile_step);
    stage_bias(sbias, p.bias, p.Nn);

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA); prefetch_tmap(&tmB);
        for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 4); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer (warp-uniform loop, one elected lane issues) =====
        {
            int stage = 0; uint32_t phase = 0;
            for (int tile = tile_first; tile < tile_end; tile += tile_step) {
                const int nt = tile % p.n_tiles, mt = tile / p.n_tiles;
                const int per_img = p.tiles_x * p.tiles_y;
                int img0, qy0, qx0;
                if (p.BI > 1) { img0 = mt * p.BI; qy0 = 0; qx0 = 0; }
                else { img0 = mt / per_img; const int r = mt - img0 * per_img; qy0 = (r / p.tiles_x) * p.BH; qx0 = (r % p.tiles_x) * p.BW; }
                for (int t = 0; t < p.ntaps; ++t) {
                    const int cy = qy0 * p.IS + p.dy[t], cx = qx0 * p.IS + p.dx[t];
                    for (int kc = 0; kc < p.kchunks; ++kc) {
                        mbar_wait(&empty[stage], phase ^ 1);
                        if (elect_one_sync()) {
                            mbar_expect_tx(&full[stage], Cfg::STAGE_BYTES);
                            tma_load_4d(sA + stage * A_BYTES, &tmA, &full[stage], kc * 64, cx, cy, img0);
                            tma_load_2d(sB + stage * Cfg::B_BYTES, &tmB, &full[stage], kc * 64, p.wrow[t] + nt * BN);
                        }
                        __syncwarp();
                        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (warp-uniform loop, one elected lane issues) =====
        constexpr uint32_t idesc = make_idesc(128, BN, 0, 0);
        int stage = 0; uint32_t phase = 0;
        int local = 0;
        for (int tile = tile_first; tile < tile_end; tile += tile_step, ++local) {
            const int as = local & 1;
            mbar_wait(&tempty[as], ((local >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + as * BN;
            for (int it = 0; it < niter; ++it) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint64_t adesc = make_desc_sw128(smem_u32(sA + stage * A_BYTES), 16, 1024);
                    const uint64_t bdesc = make_desc_sw128(smem_u32(sB + stage * Cfg::B_BYTES), 16, 1024);
#pragma unroll
                    for (int k = 0; k < 4; ++k)      // 4 x (K = 16) per 64-channel chunk: +32 bytes inside the swizzle row
                        umma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (it | k) != 0);
                    umma_commit(&empty[stage]);
                }
                __syncwarp();
                if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
            }
            if (elect_one_sync()) umma_commit(&tfull[as]);
            __syncwarp();
        }
    } else {
        // ===== epilogue: TMEM -> registers -> bias + activation -> bf16 -> global =====
        const int quad = warp & 3;                    // TMEM lane quadrant this warp may access
        const int row = quad * 32 + lane;             // accumulator row = lattice point within the tile
        int local = 0;
        for (int tile = tile_first; tile < tile_end; tile += tile_step, ++local) {
            const int as = local & 1;
            const int nt = tile % p.n_tiles, mt = tile / p.n_tiles;
            const int per_img = p.tiles_x * p.tiles_y;
            int img, qy, qx;
            {
                const int x = row % p.BW, y = (row / p.BW) % p.BH, i = row / (p.BW * p.BH);
                if (p.BI > 1) { img = mt * p.BI + i; qy = y; qx = x; }
                else { const int im = mt / per_img; const int r = mt - im * per_img; img = im; qy = (r / p.tiles_x) * p.BH + y; qx = (r % p.tiles_x) * p.BW + x; }
            }
            const int oy = qy * p.OS + p.py, ox = qx * p.OS + p.px;
            const bool ok = oy < p.Hout && ox < p.Wout;
            bf16* dst = p.out + ((long long)(img * p.Hout + oy) * p.Wout + ox) * p.ldout + nt * BN;
            mbar_wait(&tfull[as], (local >> 1) & 1);
            tc_fence_after();
            epi_row<BN / 32>(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * BN), p.bias ? sbias + nt * BN : nullptr, p.act, dst, ok);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[as]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}


// ------------------------------------------------------------------------------------------------
// halo kernel: stride-1 3x3 convolutions with few channels (Cin*Cout <= 8192: the full-resolution 64/128-channel layers).
//
// The generic kernel above fetches the A tile once PER TAP (9x) and the weight tile once per (tile, tap): for a 64->64 layer
// that is 216 KB of L2->SM traffic per 128x64 output tile against 1152 tensor-core cycles, i.e. the kernel runs at the
// ~10 TB/s TMA/L2 fabric limit and 16 % of the tensor peak (profiles/r01_conv_layers_first.json).  Here
//   * the weights of ALL taps stay resident in shared memory for the life of the persistent CTA (<= 144 KB), and
//   * ONE TMA box brings the (16+2) x (8+2)-pixel halo of a 16 x 8 output tile; the nine taps are nine smem matrix descriptors
//     into that same halo tile: start address shifted by (dy * 10 + dx) rows of 128 B, 8-row groups SBO = 10 rows apart.
//     tcgen05.mma applies the 128-byte swizzle to the final absolute address, so unaligned starts and SBO = 1280 address the
//     TMA-written tile correctly (measured: profiles/r01_umma_shifted_descriptor_probe.txt).
// L2->SM traffic drops to 23 KB per k-chunk per tile (9.4x less); the layer becomes shared-memory-operand / HBM bound.
// ------------------------------------------------------------------------------------------------
constexpr int HALO_W = 10, HALO_H = 18;                       // (8 + 2) x (16 + 2) pixels
constexpr int HALO_BYTES = HALO_W * HALO_H * 128;             // 23040
constexpr int HALO_STAGE = 23552;                             // rounded up to 1024
struct HaloParams {
    int tdy[9], tdx[9], wrow[9];     // tap offsets relative to the halo origin (0..2) and first weight row of the tap
    int oy, ox;                      // halo origin relative to the tile origin (-1 for SAME 3x3)
    int tiles_x, tiles_y, total_tiles;
    int H, W, ldout, Nn;
    const float* bias; int act;
    bf16* out;
    double* stats;                   // optional instance-norm statistics [N][Nn][2] (see EpiStats)
};

// CPX = bytes of one pixel row of the A operand = 2 x (reduction channels per k-chunk): 128 (64 channels, SWIZZLE_128B) for the
// regular layers; 64 / 32 (32 / 16 channels, SWIZZLE_64B / 32B) for the THIN layers -- SpecSeg's 16/32-channel levels
// (SpecSeg.py:34-44, :76-86), the 1-channel mask inputs of the attention branches (ShmGANwithSSpecSeg.py:404-412) and the
// 10-channel generator input (:243) -- which otherwise run zero-padded to 64 channels (4x the bytes, up to 16x the MMAs).
// The shifted-descriptor trick is unchanged: the swizzle is a function of absolute shared-memory address bits in every mode.
// T = 16 x 8 sub-tiles stacked vertically in one work item (one (16 T + 2) x 10 halo box, T accumulators, one epilogue pass over
// T x BN columns).  A thin tile has 9-18 MMAs of 8-16 cycles against ~1000 cycles of fixed per-tile latency (barrier round trips,
// tcgen05.ld, commit): with T = 1 the thin layers ran at 1.8-2.7 TB/s; T = 4 amortises that latency over 512 pixels.
template <int KC, int BN, int CPX = 128, int T = 1>
struct HaloCfg {
    static_assert(CPX == 128 || KC == 1, "thin rows hold the whole reduction dimension");
    static_assert(2 * T * BN <= 512, "accumulators exceed TMEM");
    static constexpr int W_TILE = BN * CPX;                       // one tap, one k-chunk
    static constexpr int W_BYTES = 9 * KC * W_TILE;
    static constexpr int W_SPACE = (W_BYTES + 1023) / 1024 * 1024;
    static constexpr int A_BYTES = HALO_W * (16 * T + 2) * CPX;
    static constexpr int A_STAGE = (A_BYTES + 1023) / 1024 * 1024;
    // TMA-store epilogue (BN <= 64, one k-chunk).  A thread owns one pixel row of the accumulator, so direct stores put the 32 lanes
    // of every STG.128 on 32 different 128-byte lines (32 L1 wavefronts per instruction: ~1000 cycles per 128 x 64 tile), and the
    // store address arithmetic made the four epilogue warps a serial latency chain (ncu on a stacked thin layer: 0.19 IPC per
    // epilogue warp while the MMA warp waits for the accumulator).  Instead each group of four epilogue warps packs a sub-tile,
    // writes it into a swizzled shared-memory tile (conflict-free STS.128), and one thread hands the 128-pixel x BN tile to the TMA
    // engine (cp.async.bulk.tensor store): whole lines, no per-thread global addresses, asynchronous.  Two staging tiles per group.
    static constexpr int EPI_WARPS = T > 1 ? 8 : 4;                       // stacked items: two warps per TMEM lane quadrant
    static constexpr int EPI_GROUPS = EPI_WARPS / 4;
    static constexpr bool TSTORE = BN <= 64 && KC == 1;
    static constexpr int SROW = 2 * BN;                                   // bytes of one staged pixel row: 128 / 64 / 32
    static constexpr int STG_TILE = 128 * SROW;                           // one sub-tile: 16 / 8 / 4 KB
    static constexpr int STG_BYTES = TSTORE ? ((2 * EPI_GROUPS * STG_TILE + 1023) / 1024 * 1024) : 0;
    static constexpr int FIXED = 1024 + 256 + BIAS_SMEM + STG_BYTES;
    static constexpr int FIT = (227 * 1024 - FIXED - W_SPACE) / A_STAGE;  // halo stages that fit beside the resident weights
    static constexpr int WANT = T > 1 ? 4 : ((W_BYTES <= 73728) ? 6 : 3);
    static constexpr int STAGES = FIT < WANT ? FIT : WANT;
    static_assert(STAGES >= 2, "halo kernel shared memory");
    static constexpr int SMEM = W_SPACE + STAGES * A_STAGE + FIXED;
    static constexpr int TMEM_COLS = 2 * T * BN < 32 ? 32 : 2 * T * BN;
    static constexpr uint32_t LAYOUT = CPX == 128 ? 2u : (CPX == 64 ? 4u : 6u);
    static constexpr int THREADS = 64 + 32 * EPI_WARPS;               // warps 2-5 take the even sub-tiles of a stacked item, 6-9 the odd ones
};

template <int KC, int BN, int CPX = 128, int T = 1>
__global__ void __launch_bounds__((HaloCfg<KC, BN, CPX, T>::THREADS), 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmO,
                 const HaloParams p) {
    using Cfg = HaloCfg<KC, BN, CPX, T>;
    constexpr int HALO_STAGE = Cfg::A_STAGE;
    constexpr int HALO_BYTES = Cfg::A_BYTES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sW = smem;                                   // [9 taps][KC][BN rows x CPX B]
    uint8_t* sStage = smem + Cfg::W_SPACE;                // [EPI_GROUPS][2][128 rows x SROW B] output staging (1024-aligned)
    uint8_t* sA = sStage + Cfg::STG_BYTES;                // [STAGES][HALO_STAGE]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sA + Cfg::STAGES * HALO_STAGE);
    uint64_t* full = bars;
    uint64_t* empty = bars + Cfg::STAGES;
    uint64_t* tfull = bars + 2 * Cfg::STAGES;             // [2]
    uint64_t* tempty = tfull + 2;                         // [2]
    uint64_t* wbar = tempty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);
    float* sbias = reinterpret_cast<float*>(sA + Cfg::STAGES * HALO_STAGE + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per_img = p.tiles_x * p.tiles_y;
    int tile_first, tile_end, tile_step;
    tile_walk(p.total_tiles, BN == 64 && T == 1 && p.stats != nullptr, tile_first, tile_end, tile_step);
    stage_bias(sbias, p.bias, BN);

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA); prefetch_tmap(&tmB);
        if (Cfg::TSTORE) prefetch_tmap(&tmO);
        for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], Cfg::EPI_WARPS); }
        mbar_init(wbar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        {
            // weights of every tap, once
            if (elect_one_sync()) {
                mbar_expect_tx(wbar, Cfg::W_BYTES);
                for (int t = 0; t < 9; ++t)
                    for (int kc = 0; kc < KC; ++kc)
                        tma_load_2d(sW + (t * KC + kc) * Cfg::W_TILE, &tmB, wbar, kc * 64, p.wrow[t]);
            }
            __syncwarp();
            int stage = 0; uint32_t phase = 0;
            for (int tile = tile_first; tile < tile_end; tile += tile_step) {
                const int img = tile / per_img; const int r = tile - img * per_img;
                const int y0 = (r / p.tiles_x) * (16 * T) + p.oy, x0 = (r % p.tiles_x) * 8 + p.ox;
                for (int kc = 0; kc < KC; ++kc) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    if (elect_one_sync()) {
                        mbar_expect_tx(&full[stage], HALO_BYTES);
                        tma_load_4d(sA + stage * HALO_STAGE, &tmA, &full[stage], kc * 64, x0, y0, img);
                    }
                    __syncwarp();
                    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = make_idesc(128, BN, 0, 0);
        mbar_wait(wbar, 0);
        int stage = 0; uint32_t phase = 0;
        int local = 0;
        for (int tile = tile_first; tile < tile_end; tile += tile_step, ++local) {
            const int as = local & 1;
            mbar_wait(&tempty[as], ((local >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + as * (T * BN);
            for (int kc = 0; kc < KC; ++kc) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint32_t a0 = smem_u32(sA + stage * HALO_STAGE);
                if (elect_one_sync()) {
#pragma unroll
                    for (int j = 0; j < T; ++j) {
#pragma unroll
                        for (int t = 0; t < 9; ++t) {
                            const uint64_t adesc = make_desc_swz(a0 + (uint32_t)((p.tdy[t] + 16 * j) * HALO_W + p.tdx[t]) * (uint32_t)CPX, 16, HALO_W * CPX, Cfg::LAYOUT);
                            const uint64_t bdesc = make_desc_swz(smem_u32(sW + (t * KC + kc) * Cfg::W_TILE), 16, 8 * CPX, Cfg::LAYOUT);
#pragma unroll
                            for (int k = 0; k < CPX / 32; ++k)
                                umma_bf16(d_tmem + j * BN, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kc | t | k) != 0);
                        }
                    }
                    umma_commit(&empty[stage]);
                }
                __syncwarp();
                if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
            }
            if (elect_one_sync()) umma_commit(&tfull[as]);
            __syncwarp();
        }
    } else {
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const int ty = row >> 3, tx = row & 7;
        int local = 0;
        uint32_t nstore = 0;                              // sub-tiles this epilogue group has handed to the TMA engine
        constexpr bool STATS = BN == 64 && T == 1;        // see EpiStats: private running sums cost 2 x BN registers per thread
        typename std::conditional<STATS, EpiStats<64>, NoStats>::type st;
        st.init();
        for (int tile = tile_first; tile < tile_end; tile += tile_step, ++local) {
            const int as = local & 1;
            const int img = tile / per_img; const int r = tile - img * per_img;
            const int y0t = (r / p.tiles_x) * (16 * T), x0t = (r % p.tiles_x) * 8;
            mbar_wait(&tfull[as], (local >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * (T * BN));
            if constexpr (Cfg::TSTORE) {
                constexpr int CH = BN / 8;                                // 16-byte chunks per staged pixel row
                constexpr int G = Cfg::EPI_GROUPS, TJ = T / G;            // sub-tiles of this item handled by my group
                const int eg = (warp - 2) >> 2;                           // epilogue group; its sub-tiles are eg, eg + G, ...
                const bool leader = ((warp - 2) & 3) == 0 && lane == 0;
                uint4 pk[TJ][CH];
                if constexpr (BN == 16) {
                    uint32_t r16[TJ][16];
#pragma unroll
                    for (int jj = 0; jj < TJ; ++jj) tmem_ld16_nw(taddr + (uint32_t)((G * jj + eg) * BN), r16[jj]);
#pragma unroll
                    for (int jj = 0; jj < TJ; ++jj) { tmem_wait_ld16(r16[jj]); epi_pack<16>(r16[jj], p.bias ? sbias : nullptr, p.act, pk[jj]); }
                } else {
#pragma unroll
                    for (int jj = 0; jj < TJ; ++jj) {
                        uint32_t r32[BN / 32][32];
#pragma unroll
                        for (int c = 0; c < BN / 32; ++c) tmem_ld32_nw(taddr + (uint32_t)((G * jj + eg) * BN + c * 32), r32[c]);
#pragma unroll
                        for (int c = 0; c < BN / 32; ++c) {
                            tmem_wait_ld32(r32[c]);
                            epi_pack<32>(r32[c], p.bias ? sbias + c * 32 : nullptr, p.act, pk[jj] + c * 4);
                        }
                    }
                }
                // the accumulators are in registers: hand the TMEM buffer back before the store phase
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[as]);
#pragma unroll
                for (int jj = 0; jj < TJ; ++jj) {
                    const uint32_t sb = smem_u32(sStage + (eg * 2 + (nstore & 1)) * Cfg::STG_TILE);
                    // the bulk store that read this staging tile two sub-tiles ago must be done reading it
                    if (leader) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    asm volatile("bar.sync %0, 128;" ::"r"(1 + eg) : "memory");
                    const uint32_t rowa = sb + (uint32_t)row * Cfg::SROW;
#pragma unroll
                    for (int c = 0; c < CH; ++c) {
                        const uint32_t a = rowa + c * 16;
                        const uint32_t sw = a ^ (((a >> 7) & (uint32_t)(Cfg::SROW / 16 - 1)) << 4);   // TMA swizzle of the row width
                        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(sw),
                                     "r"(pk[jj][c].x), "r"(pk[jj][c].y), "r"(pk[jj][c].z), "r"(pk[jj][c].w) : "memory");
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    asm volatile("bar.sync %0, 128;" ::"r"(1 + eg) : "memory");
                    if (leader) {
                        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                                     ::"l"(&tmO), "r"(sb), "r"(0), "r"(x0t), "r"(y0t + 16 * (G * jj + eg)), "r"(img) : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                    ++nstore;
                }
                if constexpr (STATS) {
                    if (p.stats != nullptr) {
                        st.own(p.stats, img, lane);
                        st.template add<CH>(0, pk[0]);
                    }
                }
            } else {
                const int oy = y0t + ty, ox = x0t + tx;
                bf16* dst = p.out + ((long long)(img * p.H + oy) * p.W + ox) * p.ldout;
                if constexpr (T > 1) {
                    const long long rowstep = (long long)16 * p.W * p.ldout;
                    const int eg = (warp - 2) >> 2;                   // epilogue warp group: sub-tiles eg, eg + 2, ...
                    if constexpr (BN == 16) {
                        uint32_t r16[T / 2][16];
#pragma unroll
                        for (int jj = 0; jj < T / 2; ++jj) tmem_ld16_nw(taddr + (uint32_t)((2 * jj + eg) * BN), r16[jj]);
#pragma unroll
                        for (int jj = 0; jj < T / 2; ++jj) {
                            tmem_wait_ld16(r16[jj]);
                            uint4 pk[2];
                            epi_pack<16>(r16[jj], p.bias ? sbias : nullptr, p.act, pk);
                            uint4* g = reinterpret_cast<uint4*>(dst + (2 * jj + eg) * rowstep);
                            g[0] = pk[0]; g[1] = pk[1];
                        }
                    } else {
#pragma unroll
                        for (int jj = 0; jj < T / 2; ++jj)
                            epi_row<BN / 32>(taddr + (uint32_t)((2 * jj + eg) * BN), p.bias ? sbias : nullptr, p.act, dst + (2 * jj + eg) * rowstep, true);
                    }
                } else if constexpr (BN == 16) {
                    uint32_t r16[16];
                    tmem_ld16_nw(taddr, r16);
                    tmem_wait_ld16(r16);
                    uint4 pk[2];
                    epi_pack<16>(r16, p.bias ? sbias : nullptr, p.act, pk);
                    uint4* g = reinterpret_cast<uint4*>(dst);
                    g[0] = pk[0]; g[1] = pk[1];
                } else {
                    constexpr int NC = BN / 32;
                    uint32_t r[NC][32];
#pragma unroll
                    for (int c = 0; c < NC; ++c) tmem_ld32_nw(taddr + (uint32_t)(c * 32), r[c]);
#pragma unroll
                    for (int c = 0; c < NC; ++c) tmem_wait_ld32(r[c]);
                    if (STATS && p.stats != nullptr) {
                        st.own(p.stats, img, lane);
#pragma unroll
                        for (int c = 0; c < NC; ++c) {
                            uint4 pk4[4];
                            epi_pack<32>(r[c], p.bias ? sbias + c * 32 : nullptr, p.act, pk4);
                            uint4* g = reinterpret_cast<uint4*>(dst + c * 32);
                            g[0] = pk4[0]; g[1] = pk4[1]; g[2] = pk4[2]; g[3] = pk4[3];
                            st.template add<4>(c * 32, pk4);
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < NC; ++c) epi_store32(r[c], p.bias ? sbias + c * 32 : nullptr, p.act, dst + c * 32, true);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[as]);
            }
        }
        if (STATS && p.stats != nullptr) st.flush(p.stats, lane);
        if (Cfg::TSTORE) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // every bulk store of this thread has completed
        (void)nstore; (void)ty; (void)tx;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// Resident-weight stride-2 scatter kernel: Conv2DTranspose(3, s=2) forward and the dgrad of a stride-2 Conv2D(3) whose N = 64 output columns
// and K <= 128 reduction channels let ALL nine weight tiles stay in shared memory (9 x K/64 x 8 KB <= 147 KB): up4T (128 -> 64) and d2's dgrad.
// The streamed form (conv_multi_kernel<64, 2, 1>) pulls an 8 KB weight tile per four N = 64 MMAs (~35 B/clk/SM on top of the halo tiles) and ran
// at 470-520 TFLOP/s, half of what an N = 64 MMA stream can do; here the only stream is ONE 18 x 10 halo box per k-chunk, exactly as in
// conv_halo_kernel, and the nine taps scatter into four parity-class accumulators (4 x 64 columns, double-buffered = all 512 TMEM columns).
// ------------------------------------------------------------------------------------------------
struct ScatResParams {
    int tdy[9], tdx[9], wrow[9], acc[9], first[9];   // tap offset inside the halo (0..1), first weight row, parity class 2*ry+rx, first tap of its class
    int tiles_x, tiles_y, total_tiles;
    const float* bias; int act;
};
// one output map per parity class (py, px): the class's pixels (2*qy + py, 2*qx + px) as a dense (C, Wq, Hq, N) tensor with doubled strides
struct ScatOutMaps { CUtensorMap m[4]; };
template <int KC>
struct ScatResCfg {
    static constexpr int W_TILE = 64 * 128;
    static constexpr int W_BYTES = 9 * KC * W_TILE;
    static constexpr int STG_TILE = 128 * 128;                       // one staged accumulator: 128 pixel rows x 64 bf16
    static constexpr int STG_BYTES = 2 * STG_TILE;
    static constexpr int FIXED = 1024 + 256 + 256;                   // alignment slack, barriers, 64 bias floats
    static constexpr int FIT = (227 * 1024 - FIXED - W_BYTES - STG_BYTES) / HALO_STAGE;
    static constexpr int STAGES = FIT < 4 ? FIT : 4;
    static_assert(STAGES >= 2, "resident scatter kernel shared memory");
    static constexpr int SMEM = W_BYTES + STG_BYTES + STAGES * HALO_STAGE + FIXED;
};

constexpr int SCAT_RES_THREADS = 320;   // producer + MMA issuer + EIGHT epilogue warps (two per TMEM lane quadrant, 32 columns each)
template <int KC>
__global__ void __launch_bounds__(SCAT_RES_THREADS, 1)
conv_scat_res_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ ScatOutMaps om,
                     const ScatResParams p) {
    using Cfg = ScatResCfg<KC>;
    constexpr int BN = 64, SET = 4 * BN;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sW = smem;                                   // [9 taps][KC][64 rows x 128 B]
    uint8_t* sStage = smem + Cfg::W_BYTES;                // [2][128 rows x 128 B] output staging (1024-aligned)
    uint8_t* sA = sStage + Cfg::STG_BYTES;                // [STAGES][HALO_STAGE]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sA + Cfg::STAGES * HALO_STAGE);
    uint64_t* full = bars;
    uint64_t* empty = bars + Cfg::STAGES;
    uint64_t* tfull = bars + 2 * Cfg::STAGES;             // [2]
    uint64_t* tempty = tfull + 2;                         // [2]
    uint64_t* wbar = tempty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);
    float* sbias = reinterpret_cast<float*>(sA + Cfg::STAGES * HALO_STAGE + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per_img = p.tiles_x * p.tiles_y;
    stage_bias(sbias, p.bias, BN);
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA); prefetch_tmap(&tmB);
        for (int a = 0; a < 4; ++a) prefetch_tmap(&om.m[a]);
        for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 8); }
        mbar_init(wbar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one_sync()) {
            mbar_expect_tx(wbar, Cfg::W_BYTES);
            for (int t = 0; t < 9; ++t)
                for (int kc = 0; kc < KC; ++kc)
                    tma_load_2d(sW + (t * KC + kc) * Cfg::W_TILE, &tmB, wbar, kc * 64, p.wrow[t]);
        }
        __syncwarp();
        int stage = 0; uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            const int img = tile / per_img; const int r = tile - img * per_img;
            const int y0 = (r / p.tiles_x) * 16 - 1, x0 = (r % p.tiles_x) * 8 - 1;
            // With the weights and the store staging resident there is room for two halo stages only: a stage is refilled when its 36 MMAs
            // (~2100 cycles) retire, and the refill takes a full HBM round trip (~3000 cycles) -- ncu showed the tensor pipe at 32 % of a 55 %
            // ceiling with DRAM at 38 %.  The halo boxes of this CTA's NEXT tile are prefetched into L2 here, one tile time ahead.
            const int nxt = tile + gridDim.x;
            if (nxt < p.total_tiles && elect_one_sync()) {
                const int nimg = nxt / per_img; const int nr = nxt - nimg * per_img;
                for (int kc = 0; kc < KC; ++kc)
                    tma_prefetch_4d(&tmA, kc * 64, (nr % p.tiles_x) * 8 - 1, (nr / p.tiles_x) * 16 - 1, nimg);
            }
            __syncwarp();
            for (int kc = 0; kc < KC; ++kc) {
                mbar_wait(&empty[stage], phase ^ 1);
                if (elect_one_sync()) {
                    mbar_expect_tx(&full[stage], HALO_BYTES);
                    tma_load_4d(sA + stage * HALO_STAGE, &tmA, &full[stage], kc * 64, x0, y0, img);
                }
                __syncwarp();
                if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = make_idesc(128, BN, 0, 0);
        mbar_wait(wbar, 0);
        int stage = 0; uint32_t phase = 0;
        int local = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++local) {
            const int as = local & 1;
            mbar_wait(&tempty[as], ((local >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + as * SET;
            for (int kc = 0; kc < KC; ++kc) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint32_t a0 = smem_u32(sA + stage * HALO_STAGE);
                if (elect_one_sync()) {
#pragma unroll
                    for (int t = 0; t < 9; ++t) {
                        const uint64_t adesc = make_desc_sw128(a0 + (uint32_t)(p.tdy[t] * HALO_W + p.tdx[t]) * 128u, 16, HALO_W * 128);
                        const uint64_t bdesc = make_desc_sw128(smem_u32(sW + (t * KC + kc) * Cfg::W_TILE), 16, 1024);
                        const uint32_t dcol = d_tmem + p.acc[t] * BN;
                        umma_bf16(dcol, adesc, bdesc, idesc, (kc == 0 && p.first[t]) ? 0u : 1u);
#pragma unroll
                        for (int k = 1; k < 4; ++k)
                            umma_bf16(dcol, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, 1u);
                    }
                    umma_commit(&empty[stage]);
                }
                __syncwarp();
                if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
            }
            if (elect_one_sync()) umma_commit(&tfull[as]);
            __syncwarp();
        }
    } else {
        // ===== epilogue.  A thread owns one pixel ROW of an accumulator, and the four parity classes land on output pixels two apart, so direct
        // stores put the 32 lanes of every STG.128 on 32 different 128-byte lines (32 L1 wavefronts per instruction, ~4000 per tile -- as long as the
        // tile's MMAs: ncu showed the tensor pipe 27 % active and L1 75 % busy).  Instead each accumulator is packed, staged in shared memory in the
        // TMA swizzle and handed to the TMA engine as ONE 16 x 8-pixel box of its parity class (om.m[class]: the class's pixels as a dense tensor
        // with doubled pixel / row strides; image borders are clipped by the map).
        // EIGHT warps: with four, ncu's source view had the epilogue warps issuing ~95 % of the time (bias + LeakyReLU + pack of 4 x 64 columns per
        // thread and tile is ~1300 instructions against ~4200 cycles of MMAs, plus the tcgen05.ld / barrier / store latencies): 7900 cycles per tile.
        // Warps 2-5 take columns 0-31 of their TMEM lane quadrant, warps 6-9 columns 32-63.
        const int quad = warp & 3;
        const int half = (warp - 2) >> 2;
        const int row = quad * 32 + lane;
        const bool leader = warp == 2 && lane == 0;
        int local = 0;
        uint32_t cnt = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++local) {
            const int as = local & 1;
            const int img = tile / per_img; const int r = tile - img * per_img;
            const int y0t = (r / p.tiles_x) * 16, x0t = (r % p.tiles_x) * 8;
            mbar_wait(&tfull[as], (local >> 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int a = 0; a < 4; ++a, ++cnt) {
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * SET + a * BN + half * 32);
                uint32_t r32[32];
                tmem_ld32_nw(taddr, r32);
                uint4 pk[4];
                tmem_wait_ld32(r32);
                epi_pack<32>(r32, p.bias ? sbias + half * 32 : nullptr, p.act, pk);
                if (a == 3) {                              // the last accumulator of the set is in registers: hand the TMEM buffer back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty[as]);
                }
                const uint32_t sb = smem_u32(sStage + (cnt & 1) * Cfg::STG_TILE);
                // the bulk store that read this staging tile two accumulators ago must be done reading it
                if (leader) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                asm volatile("bar.sync 1, 256;" ::: "memory");
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint32_t addr = sb + (uint32_t)row * 128u + (uint32_t)(((half * 4 + c) ^ (row & 7)) << 4);
                    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[c].x), "r"(pk[c].y), "r"(pk[c].z), "r"(pk[c].w) : "memory");
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (leader) {
                    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                                 ::"l"(&om.m[a]), "r"(sb), "r"(0), "r"(x0t), "r"(y0t), "r"(img) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
        }
        if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // every bulk store has completed
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// multi-accumulator halo kernel.  One TMA box per 64-channel k-chunk brings the halo of a tile of lattice points; every tap is
// a shifted smem descriptor into it (as in conv_halo_kernel); each 64-channel weight tile (tap, k-chunk) is streamed through
// its own ring and feeds one or two accumulators.  Two configurations:
//   * "big": stride-1 3x3 convolutions with many channels (K * N > 8192: every 128..1024-channel generator layer, fwd and
//     dgrad).  32 x 8-pixel tile = TWO 128-row accumulators per weight tile: 191 KB of operands per 72 MMAs (2.7 KB/MMA against
//     the generic kernel's 8 KB/MMA) -> 1220-1410 TFLOP/s instead of 900-1100 (profiles/r01_conv_big_layers.txt).
//   * "scatter": Conv2DTranspose(k, s=2) forward and the dgrad of a stride-2 Conv2D: the FOUR output-parity classes of a
//     16 x 8 input tile are four accumulators fed from ONE halo tile (the generic path launched one kernel per class, each
//     re-reading the input), written with output stride 2.
// Accumulator sets are double-buffered in TMEM when two sets fit in the 512 columns.
// ------------------------------------------------------------------------------------------------
constexpr int BIG_H = 34;                                      // (32 + 2) halo rows of HALO_W pixels
constexpr int BIG_A_ST = 44032;                                // 34 * 10 * 128 = 43520 rounded up to 1024
constexpr int BIG_A_STAGES = 2;
template <int BN>
struct MultiCfg {
    static constexpr int B_ST = BN * 128;
    static constexpr int B_STAGES = 131072 / B_ST;             // 8 (BN = 128) or 16 (BN = 64)
    static constexpr int SMEM = BIG_A_STAGES * BIG_A_ST + B_STAGES * B_ST + 1024 + 512 + BIAS_SMEM;
};
struct MultiParams {
    int ntaps; int wrow[9]; int npairs[9]; int aoff[9][2]; int acc[9][2]; int first[9][2];
    int nacc; int row_dy[4], py[4], px[4];
    int OS, TH, a_bytes, hy, hx;              // output stride, tile height, bytes of the halo box, halo origin relative to the tile origin
    int kchunks, tiles_x, tiles_y, m_tiles, n_tiles;
    int Hout, Wout, ldout;
    int nstore;                               // output channels that exist in `out` (<= Nn; the rest are zero-padding columns)
    const float* bias; int act;
    bf16* out;
    // CTA-pair scatter only: a tile is worked as `halves` (1 or 2) items; item half h runs the taps with half[t] == h into nacc accumulators whose
    // parity classes are py / px [h * nacc + a] -- two classes per half keep the set at 2 x 128 columns, so the sets double-buffer in TMEM
    int halves, half[9];
};

template <int BN, int NBUF, int NPAIR>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_multi_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const MultiParams p) {
    using Cfg = MultiCfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = smem + BIG_A_STAGES * BIG_A_ST;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + Cfg::B_STAGES * Cfg::B_ST);
    uint64_t* fullA = bars;
    uint64_t* emptyA = fullA + BIG_A_STAGES;
    uint64_t* fullB = emptyA + BIG_A_STAGES;
    uint64_t* emptyB = fullB + Cfg::B_STAGES;
    uint64_t* tfull = emptyB + Cfg::B_STAGES;            // [2]
    uint64_t* tempty = tfull + 2;                        // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    float* sbias = reinterpret_cast<float*>(sB + Cfg::B_STAGES * Cfg::B_ST + 512);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total = p.m_tiles * p.n_tiles;
    const int per_img = p.tiles_x * p.tiles_y;
    const int set_cols = p.nacc * BN;
    int item_first, item_end, item_step;
    tile_walk(total, false, item_first, item_end, item_step);
    stage_bias(sbias, p.bias, p.n_tiles * BN);

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA); prefetch_tmap(&tmB);
        for (int i = 0; i < BIG_A_STAGES; ++i) { mbar_init(&fullA[i], 1); mbar_init(&emptyA[i], 1); }
        for (int i = 0; i < Cfg::B_STAGES; ++i) { mbar_init(&fullB[i], 1); mbar_init(&emptyB[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        {
            int sa = 0, sb = 0; uint32_t pha = 0, phb = 0;
            for (int item = item_first; item < item_end; item += item_step) {
                const int nt = item % p.n_tiles, mt = item / p.n_tiles;
                const int img = mt / per_img; const int r = mt - img * per_img;
                const int y0 = (r / p.tiles_x) * p.TH + p.hy, x0 = (r % p.tiles_x) * 8 + p.hx;
                for (int kc = 0; kc < p.kchunks; ++kc) {
                    mbar_wait(&emptyA[sa], pha ^ 1);
                    if (elect_one_sync()) {
                        mbar_expect_tx(&fullA[sa], p.a_bytes);
                        tma_load_4d(sA + sa * BIG_A_ST, &tmA, &fullA[sa], kc * 64, x0, y0, img);
                    }
                    __syncwarp();
                    if (++sa == BIG_A_STAGES) { sa = 0; pha ^= 1; }
                    for (int t = 0; t < p.ntaps; ++t) {
                        mbar_wait(&emptyB[sb], phb ^ 1);
                        if (elect_one_sync()) {
                            mbar_expect_tx(&fullB[sb], Cfg::B_ST);
                            tma_load_2d(sB + sb * Cfg::B_ST, &tmB, &fullB[sb], kc * 64, p.wrow[t] + nt * BN);
                        }
                        __syncwarp();
                        if (++sb == Cfg::B_STAGES) { sb = 0; phb ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = make_idesc(128, BN, 0, 0);
        int sa = 0, sb = 0; uint32_t pha = 0, phb = 0;
        int local = 0;
        for (int item = item_first; item < item_end; item += item_step, ++local) {
            const int as = local % NBUF;
            mbar_wait(&tempty[as], ((local / NBUF) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d0 = tmem_base + as * set_cols;
            for (int kc = 0; kc < p.kchunks; ++kc) {
                mbar_wait(&fullA[sa], pha);
                tc_fence_after();
                const uint32_t a0 = smem_u32(sA + sa * BIG_A_ST);
                // fully unrolled over the (at most nine) taps: the per-tap parameters become constant-bank operands
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    if (t < p.ntaps) {
                        mbar_wait(&fullB[sb], phb);
                        tc_fence_after();
                        if (elect_one_sync()) {
                            const uint64_t bdesc = make_desc_sw128(smem_u32(sB + sb * Cfg::B_ST), 16, 1024);
#pragma unroll
                            for (int j = 0; j < NPAIR; ++j) {
                                const uint64_t adesc = make_desc_sw128(a0 + (uint32_t)p.aoff[t][j] * 128u, 16, HALO_W * 128);
                                const uint32_t keep = (kc == 0 && p.first[t][j]) ? 0u : 1u;
                                const uint32_t dcol = d0 + p.acc[t][j] * BN;
                                umma_bf16(dcol, adesc, bdesc, idesc, keep);
#pragma unroll
                                for (int k = 1; k < 4; ++k)
                                    umma_bf16(dcol, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, 1u);
                            }
                            umma_commit(&emptyB[sb]);
                        }
                        __syncwarp();
                        if (++sb == Cfg::B_STAGES) { sb = 0; phb ^= 1; }
                    }
                }
                if (elect_one_sync()) umma_commit(&emptyA[sa]);
                __syncwarp();
                if (++sa == BIG_A_STAGES) { sa = 0; pha ^= 1; }
            }
            if (elect_one_sync()) umma_commit(&tfull[as]);
            __syncwarp();
        }
    } else {
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const int ty = row >> 3, tx = row & 7;
        int local = 0;
        for (int item = item_first; item < item_end; item += item_step, ++local) {
            const int as = local % NBUF;
            const int nt = item % p.n_tiles, mt = item / p.n_tiles;
            const int img = mt / per_img; const int r = mt - img * per_img;
            const int qy = (r / p.tiles_x) * p.TH + ty, qx = (r % p.tiles_x) * 8 + tx;
            mbar_wait(&tfull[as], (local / NBUF) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int a = 0; a < p.nacc; ++a) {
                const int oy = (qy + p.row_dy[a]) * p.OS + p.py[a], ox = qx * p.OS + p.px[a];
                const bool ok = oy < p.Hout && ox < p.Wout;
                bf16* dst = p.out + ((long long)(img * p.Hout + oy) * p.Wout + ox) * p.ldout + nt * BN;
                epi_row<BN / 32>(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * set_cols + a * BN), p.bias ? sbias + nt * BN : nullptr,
                                 p.act, dst, ok, p.nstore - nt * BN);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[as]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// CTA-pair form of the "big" configuration: stride-1 3x3 convolutions with Nn % 128 == 0 (every 128..1024-channel generator layer, forward and
// dgrad -- the largest single share of the training step).
//
// Measured on B200 (tools/umma_rate_probe.cu, profiles/r02_umma_rate_probe.txt): back-to-back M128 x N128 x K16 MMAs of one SM retire every
// 72.9 cycles (87.8 % of the tensor peak), the M256 x N128 x K16 MMA of a CTA PAIR every 64.1 cycles (99.8 %) -- each SM of a pair reads its 128
// rows of A plus HALF of B from its own shared memory.  (The same probe shows that N = 64 tiles gain nothing from pairing: 58.4 vs 59.6 cycles,
// a fixed per-instruction floor, not operand bandwidth.)  Pairing also halves the weight stream per SM: the single-CTA kernel pulls a 16 KB
// weight tile per 512 MMA cycles = 32 B/clk/SM on top of the halo tiles -- ~6 KB/clk over 148 SMs, the L2 -> SM limit.
//
// Structure (one cluster = two CTAs = two SMs of a TPC; each CTA is the warp-specialised CTA of conv_multi_kernel):
//   * each CTA owns one 32 x 8-pixel tile (two 128-row accumulators, double-buffered in its own TMEM) and loads that tile's halo itself;
//   * each CTA loads 64 of the 128 rows of every streamed weight tile (8 KB instead of 16 KB);
//   * every TMA load signals the LEADER's full barrier (cp.async.bulk.tensor ... cta_group::2); the leader's MMA warp issues ONE
//     tcgen05.mma.cta_group::2 per (accumulator, k-step) for both tiles; tcgen05.commit ... multicast::cluster releases the operand slots and
//     publishes the accumulators in BOTH CTAs; the peer's epilogue warps hand the accumulator buffer back with a remote mbarrier arrive.
// ------------------------------------------------------------------------------------------------
constexpr int B2_ST = 64 * 128;                                // this CTA's half of a 128-row weight tile
constexpr int B2_STAGES = 16;
constexpr int BIG2_SMEM = BIG_A_STAGES * BIG_A_ST + B2_STAGES * B2_ST + 1024 + 512 + BIAS_SMEM;

// NBUF / NPAIR as in conv_multi_kernel: <2, 2> = "big" (two accumulators per weight tile, double-buffered), <1, 1> = stride-2 "scatter" with
// 128 columns (four parity accumulators = all of TMEM; pairing halves its weight stream, which is what bounds it)
// Epilogue warps per CTA: four for the stride-1 "big" configuration, EIGHT (two per TMEM lane quadrant, 64 of the 128 columns each) for the
// stride-2 scatter configuration.  With four, the scatter epilogue was bound by its warps' issue rate, like conv_scat_res_kernel before it
// (2 accumulators x 128 columns of bias + activation + pack + scattered stores per thread against 64-80 MMAs of 64 cycles per item; ncu: tensor
// pipe 42-46 %): +5 %.  The big configuration has 144+ MMAs per item to hide its epilogue behind and LOST 1.7 % with eight warps (eight spinning
// mbarrier waiters instead of four beside the MMA issuer), so it keeps four.
template <int NPAIR> struct Big2Epi { static constexpr int WARPS = NPAIR == 1 ? 8 : 4, THREADS = 64 + 32 * WARPS, COLS = 512 / WARPS; };
template <int NBUF, int NPAIR>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Big2Epi<NPAIR>::THREADS, 1)
conv_big2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const MultiParams p) {
    constexpr int BN = 128;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = smem + BIG_A_STAGES * BIG_A_ST;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + B2_STAGES * B2_ST);
    uint64_t* fullA = bars;                              // used in the leader only (both CTAs' loads land here)
    uint64_t* emptyA = fullA + BIG_A_STAGES;
    uint64_t* fullB = emptyA + BIG_A_STAGES;             // leader only
    uint64_t* emptyB = fullB + B2_STAGES;
    uint64_t* tfull = emptyB + B2_STAGES;                // [2]
    uint64_t* tempty = tfull + 2;                        // [2] leader only: the epilogue warps of both CTAs
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    float* sbias = reinterpret_cast<float*>(sB + B2_STAGES * B2_ST + 512);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pairs_m = (p.m_tiles + 1) >> 1;
    const int halves = p.halves > 1 ? p.halves : 1;
    const int total = pairs_m * p.n_tiles * halves;
    const int per_img = p.tiles_x * p.tiles_y;
    const int nclusters = gridDim.x >> 1, cid = blockIdx.x >> 1;
    const int set_cols = p.nacc * BN;
    stage_bias(sbias, p.bias, p.n_tiles * BN);

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA); prefetch_tmap(&tmB);
        // full barriers: ONE arrival (the leader's arrive.expect_tx for the bytes of both CTAs); the peer only contributes transaction bytes.
        // (A per-stage remote mbarrier.arrive.release.cluster from the peer's producer cost ~500 cycles per weight tile and halved the kernel's
        // throughput: 780 instead of 1400 TFLOP/s.)  A peer's bytes can land before the leader's expect_tx of the same phase -- the transaction
        // count goes negative for a moment while the pending arrival keeps the phase open -- but never in an earlier phase: the peer refills a
        // slot only after the multicast commit that released it.
        for (int i = 0; i < BIG_A_STAGES; ++i) { mbar_init(&fullA[i], 1); mbar_init(&emptyA[i], 1); }
        for (int i = 0; i < B2_STAGES; ++i) { mbar_init(&fullB[i], 1); mbar_init(&emptyB[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 2 * Big2Epi<NPAIR>::WARPS); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_2sm(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                   // barriers of BOTH CTAs are initialised before anyone signals across the pair
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer (both CTAs): own halo tile + own half of every weight tile, signalled on the leader's full barriers =====
        int sa = 0, sb = 0; uint32_t pha = 0, phb = 0;
        for (int item = cid; item < total; item += nclusters) {
            const int h = item % halves, base = item / halves;
            const int nt = base % p.n_tiles, mt = 2 * (base / p.n_tiles) + (int)rank;
            const int img = mt / per_img; const int r = mt - img * per_img;       // mt == m_tiles (odd tile count): img == N, zero-filled box
            const int y0 = (r / p.tiles_x) * p.TH + p.hy, x0 = (r % p.tiles_x) * 8 + p.hx;
            for (int kc = 0; kc < p.kchunks; ++kc) {
                mbar_wait(&emptyA[sa], pha ^ 1);
                if (elect_one_sync()) {
                    const uint32_t lead = map_to_cta(&fullA[sa], 0);
                    if (rank == 0) mbar_expect_tx(&fullA[sa], 2 * p.a_bytes);
                    tma_load_4d_2sm(sA + sa * BIG_A_ST, &tmA, lead, kc * 64, x0, y0, img);
                }
                __syncwarp();
                if (++sa == BIG_A_STAGES) { sa = 0; pha ^= 1; }
                for (int t = 0; t < p.ntaps; ++t) {
                    if (p.half[t] != h) continue;
                    mbar_wait(&emptyB[sb], phb ^ 1);
                    if (elect_one_sync()) {
                        const uint32_t lead = map_to_cta(&fullB[sb], 0);
                        if (rank == 0) mbar_expect_tx(&fullB[sb], 2 * B2_ST);
                        tma_load_2d_2sm(sB + sb * B2_ST, &tmB, lead, kc * 64, p.wrow[t] + nt * BN + (int)rank * 64);
                    }
                    __syncwarp();
                    if (++sb == B2_STAGES) { sb = 0; phb ^= 1; }
                }
            }
        }
    } else if (warp == 1 && rank == 0) {
        // ===== MMA issuer (leader only): M = 256 over the pair =====
        constexpr uint32_t idesc = make_idesc(256, BN, 0, 0);
        int sa = 0, sb = 0; uint32_t pha = 0, phb = 0;
        int local = 0;
        for (int item = cid; item < total; item += nclusters, ++local) {
            const int as = local % NBUF;
            const int h = item % halves;
            mbar_wait(&tempty[as], ((local / NBUF) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d0 = tmem_base + as * set_cols;
            for (int kc = 0; kc < p.kchunks; ++kc) {
                mbar_wait(&fullA[sa], pha);
                tc_fence_after();
                const uint32_t a0 = smem_u32(sA + sa * BIG_A_ST);
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    if (t < p.ntaps && p.half[t] == h) {
                        mbar_wait(&fullB[sb], phb);
                        tc_fence_after();
                        if (elect_one_sync()) {
                            const uint64_t bdesc = make_desc_sw128(smem_u32(sB + sb * B2_ST), 16, 1024);
#pragma unroll
                            for (int j = 0; j < NPAIR; ++j) {
                                const uint64_t adesc = make_desc_sw128(a0 + (uint32_t)p.aoff[t][j] * 128u, 16, HALO_W * 128);
                                const uint32_t keep = (kc == 0 && p.first[t][j]) ? 0u : 1u;
                                const uint32_t dcol = d0 + p.acc[t][j] * BN;
                                umma_bf16_2sm(dcol, adesc, bdesc, idesc, keep);
#pragma unroll
                                for (int k = 1; k < 4; ++k)
                                    umma_bf16_2sm(dcol, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, 1u);
                            }
                            umma_commit_2sm(&emptyB[sb]);
                        }
                        __syncwarp();
                        if (++sb == B2_STAGES) { sb = 0; phb ^= 1; }
                    }
                }
                if (elect_one_sync()) umma_commit_2sm(&emptyA[sa]);
                __syncwarp();
                if (++sa == BIG_A_STAGES) { sa = 0; pha ^= 1; }
            }
            if (elect_one_sync()) umma_commit_2sm(&tfull[as]);
            __syncwarp();
        }
    } else if (warp >= 2) {
        // ===== epilogue (both CTAs): own accumulators -> bias + activation -> bf16 -> global; hands the buffer back to the LEADER's MMA warp =====
        const int quad = warp & 3;
        constexpr int CW = Big2Epi<NPAIR>::COLS;         // columns of the accumulator this warp handles: all 128, or 64 (eight warps)
        const int ch = ((warp - 2) >> 2) * CW;
        const int row = quad * 32 + lane;
        const int ty = row >> 3, tx = row & 7;
        int local = 0;
        for (int item = cid; item < total; item += nclusters, ++local) {
            const int as = local % NBUF;
            const int h = item % halves, base = item / halves;
            const int nt = base % p.n_tiles, mt = 2 * (base / p.n_tiles) + (int)rank;
            const bool valid = mt < p.m_tiles;
            const int img = mt / per_img; const int r = mt - img * per_img;
            const int qy = (r / p.tiles_x) * p.TH + ty, qx = (r % p.tiles_x) * 8 + tx;
            mbar_wait(&tfull[as], (local / NBUF) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int a = 0; a < p.nacc; ++a) {
                const int c = h * p.nacc + a;
                const int oy = (qy + p.row_dy[c]) * p.OS + p.py[c], ox = qx * p.OS + p.px[c];
                const bool ok = valid && oy < p.Hout && ox < p.Wout;
                bf16* dst = p.out + ((long long)((valid ? img : 0) * p.Hout + (ok ? oy : 0)) * p.Wout + (ok ? ox : 0)) * p.ldout + nt * BN + ch;
                epi_row<CW / 32>(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * set_cols + a * BN + ch), p.bias ? sbias + nt * BN + ch : nullptr,
                                 p.act, dst, ok, p.nstore - nt * BN - ch);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (rank == 0) mbar_arrive(&tempty[as]); else mbar_arrive_cluster(map_to_cta(&tempty[as], 0));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                   // no CTA leaves (or frees TMEM) while its peer may still signal it or read its operands
    if (warp == 1) tmem_dealloc_2sm(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 4-D activation map: dims (C, W, H, N), bf16, box (64, bw, bh, bi), traversal stride `is` on W and H
int encode_act(CUtensorMap* tm, const void* base, int C, int W, int H, int N, int ld, int bw, int bh, int bi, int is) {
    EncodeTiledFn enc = get_encode();
    if (!enc) SHM_FAIL(SHM_ECUDA, "cuTensorMapEncodeTiled entry point not found");
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)(bw * is - (is - 1)), (cuuint32_t)(bh * is - (is - 1)), (cuuint32_t)bi};
    cuuint32_t estr[4] = {1, (cuuint32_t)is, (cuuint32_t)is, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) SHM_FAIL(SHM_ECUDA, "cuTensorMapEncodeTiled(activation C=%d W=%d H=%d N=%d ld=%d box=%d,%d,%d is=%d) failed: %d", C, W, H, N, ld, bw, bh, bi, is, (int)r);
    return SHM_OK;
}
// 2-D weight map: dims (K, rows), bf16, box (64, bn)
inline CUtensorMapSwizzle swizzle_for(int inner) {
    return inner == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (inner == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}
int encode_w(CUtensorMap* tm, const void* base, int K, long long rows, int bn, int inner = 64) {
    EncodeTiledFn enc = get_encode();
    if (!enc) SHM_FAIL(SHM_ECUDA, "cuTensorMapEncodeTiled entry point not found");
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)inner, (cuuint32_t)bn};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(inner), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) SHM_FAIL(SHM_ECUDA, "cuTensorMapEncodeTiled(weights K=%d rows=%lld) failed: %d", K, rows, (int)r);
    return SHM_OK;
}

inline void out_dims(const shm_conv_desc* d, int& Ho, int& Wo) {
    if (d->transposed) { Ho = d->H * d->stride; Wo = d->W * d->stride; }
    else { Ho = cdiv(d->H, d->stride); Wo = cdiv(d->W, d->stride); }
}

// choose the 128-point tile box for a lattice Qh x Qw over N images; returns false when it does not tile exactly
bool pick_box(int N, int Qh, int Qw, int& BW, int& BH, int& BI) {
    BW = Qw >= 128 ? 128 : Qw;
    if (BW <= 0 || 128 % BW != 0 || Qw % BW != 0) return false;
    BH = 128 / BW;
    if (BH > Qh) BH = Qh;
    if (Qh % BH != 0) return false;
    BI = 128 / (BW * BH);
    if (BI * BW * BH != 128) return false;
    if (BI > 1 && (N % BI != 0)) return false;
    return true;
}

struct Geometry {
    // operand A tensor (what TMA reads), lattice, output tensor
    int Hin, Win, ldin, K;
    int Qh, Qw, IS;
    int Hout, Wout, ldout, Nn, OS;
};

int launch_tc(const Geometry& g, int N, const void* in, const void* w_tc, int wrows_total, const float* bias, int act, void* out,
              const int* tdy, const int* tdx, const int* twrow, int ntaps, int py, int px, cudaStream_t st) {
    TcParams p{};
    if (!pick_box(N, g.Qh, g.Qw, p.BW, p.BH, p.BI)) SHM_FAIL(SHM_EUNSUPPORTED, "conv_tc: lattice %dx%d (N=%d) does not tile into 128-point boxes", g.Qh, g.Qw, N);
    p.ntaps = ntaps;
    for (int t = 0; t < ntaps; ++t) { p.dy[t] = tdy[t]; p.dx[t] = tdx[t]; p.wrow[t] = twrow[t]; }
    p.kchunks = g.K / 64;
    p.Qh = g.Qh; p.Qw = g.Qw; p.IS = g.IS;
    p.tiles_x = g.Qw / p.BW; p.tiles_y = g.Qh / p.BH;
    p.m_tiles = (int)((long long)N * g.Qh * g.Qw / 128);
    const int BN = (g.Nn % 128 == 0) ? 128 : 64;
    p.n_tiles = g.Nn / BN;
    p.Hout = g.Hout; p.Wout = g.Wout; p.OS = g.OS; p.py = py; p.px = px; p.ldout = g.ldout; p.Nn = g.Nn;
    p.bias = bias; p.act = act; p.out = (bf16*)out;
    CUtensorMap tmA, tmB;
    if (int rc = encode_act(&tmA, in, g.K, g.Win, g.Hin, N, g.ldin, p.BW, p.BH, p.BI, g.IS)) return rc;
    if (int rc = encode_w(&tmB, w_tc, g.K, wrows_total, BN)) return rc;
    const int total = p.m_tiles * p.n_tiles;
    int grid = shm_num_sms();
    if (grid > total) grid = total;
    if (ntaps == 0 || total == 0) SHM_FAIL(SHM_EUNSUPPORTED, "conv_tc: empty problem");
    if (BN == 128) {
        static bool attr = false;
        if (!attr) { cudaFuncSetAttribute(conv_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<128>::SMEM); attr = true; }
        conv_tc_kernel<128><<<grid, TC_THREADS, TcCfg<128>::SMEM, st>>>(tmA, tmB, p);
    } else {
        static bool attr = false;
        if (!attr) { cudaFuncSetAttribute(conv_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<64>::SMEM); attr = true; }
        conv_tc_kernel<64><<<grid, TC_THREADS, TcCfg<64>::SMEM, st>>>(tmA, tmB, p);
    }
    SHM_CHECK_LAUNCH("conv_tc_kernel");
    return SHM_OK;
}


// encode a 4-D activation map with an explicit box (halo kernel)
int encode_act_box(CUtensorMap* tm, const void* base, int C, int W, int H, int N, int ld, int bx, int by, int inner = 64) {
    EncodeTiledFn enc = get_encode();
    if (!enc) SHM_FAIL(SHM_ECUDA, "cuTensorMapEncodeTiled entry point not found");
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
    cuuint32_t box[4] = {(cuuint32_t)inner, (cuuint32_t)bx, (cuuint32_t)by, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(inner), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) SHM_FAIL(SHM_ECUDA, "cuTensorMapEncodeTiled(halo C=%d W=%d H=%d N=%d ld=%d) failed: %d", C, W, H, N, ld, (int)r);
    return SHM_OK;
}

// output map of ONE parity class of a stride-2 scatter: `base` = the class's first pixel; dims (C, Wq, Hq, N) with pixel stride 2 * ld and row
// stride 2 * Wout * ld; box (C, 8, 16, 1) in the 128-byte swizzle (C = 64)
int encode_out_parity(CUtensorMap* tm, const void* base, int C, int Wq, int Hq, int N, int Hout, int Wout, int ld) {
    EncodeTiledFn enc = get_encode();
    if (!enc) SHM_FAIL(SHM_ECUDA, "cuTensorMapEncodeTiled entry point not found");
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)Wq, (cuuint64_t)Hq, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)ld * 4, (cuuint64_t)Wout * ld * 4, (cuuint64_t)Hout * Wout * ld * 2};
    cuuint32_t box[4] = {(cuuint32_t)C, 8, 16, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) SHM_FAIL(SHM_ECUDA, "cuTensorMapEncodeTiled(scatter output Wq=%d Hq=%d N=%d ld=%d) failed: %d", Wq, Hq, N, ld, (int)r);
    return SHM_OK;
}

// stride-1 3x3 layers the halo kernel serves: K, Nn in {64, 128} with all nine weight tiles resident (K * Nn <= 8192)
bool halo_ok(int H, int W, int K, int Nn, int kh, int kw, int stride) {
    return kh == 3 && kw == 3 && stride == 1 && H % 16 == 0 && W % 8 == 0 && (K == 64 || K == 128) && (Nn == 64 || Nn == 128) && K * Nn <= 8192;
}

template <int BN, int NBUF, int NPAIR>
int launch_multi_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const MultiParams& p, cudaStream_t st) {
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(conv_multi_kernel<BN, NBUF, NPAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, MultiCfg<BN>::SMEM); attr = true; }
    const int total = p.m_tiles * p.n_tiles;
    int grid = shm_num_sms();
    if (grid > total) grid = total;
    conv_multi_kernel<BN, NBUF, NPAIR><<<grid, TC_THREADS, MultiCfg<BN>::SMEM, st>>>(tmA, tmB, p);
    SHM_CHECK_LAUNCH("conv_multi_kernel");
    return SHM_OK;
}

inline bool pair_enabled() {
    static const bool on = []() { const char* e = getenv("SHM_BIG2"); return !(e && e[0] == '0'); }();
    return on;
}
template <int NBUF, int NPAIR>
int launch_pair_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const MultiParams& p, cudaStream_t st) {
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(conv_big2_kernel<NBUF, NPAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, BIG2_SMEM); attr = true; }
    // a cluster needs both SMs of one TPC; a persistent kernel launched with more clusters than fit at once would run the excess as a second
    // wave AFTER the first has finished all of its items, so the grid is capped by what the occupancy query says (74 pairs on a 148-SM B200)
    static int max_clusters = 0;
    if (max_clusters == 0) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(shm_num_sms() & ~1); cfg.blockDim = dim3(Big2Epi<NPAIR>::THREADS); cfg.dynamicSmemBytes = BIG2_SMEM;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, conv_big2_kernel<NBUF, NPAIR>, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); n = shm_num_sms() / 2; }
        max_clusters = n < shm_num_sms() / 2 ? n : shm_num_sms() / 2;
        if (getenv("SHM_DEBUG")) fprintf(stderr, "[shmgan] conv_big2_kernel<%d, %d>: %d CTA pairs fit on %d SMs\n", NBUF, NPAIR, max_clusters, shm_num_sms());
    }
    const int pairs = ((p.m_tiles + 1) / 2) * p.n_tiles * (p.halves > 1 ? p.halves : 1);
    const int clusters = max_clusters < pairs ? max_clusters : pairs;
    conv_big2_kernel<NBUF, NPAIR><<<2 * clusters, Big2Epi<NPAIR>::THREADS, BIG2_SMEM, st>>>(tmA, tmB, p);
    SHM_CHECK_LAUNCH("conv_big2_kernel");
    return SHM_OK;
}

// stride-1 3x3 layers the "big" configuration serves
bool big_ok(int H, int W, int K, int Nn, int kh, int kw, int stride) {
    return kh == 3 && kw == 3 && stride == 1 && H % 32 == 0 && W % 8 == 0 && K % 64 == 0 && Nn % 128 == 0;
}

int launch_big(int N, int H, int W, int K, int Nn, const void* in, int ldin, const void* w_tc, int wrows_total, const float* bias, int act,
               void* out, int ldout, const int* tdy, const int* tdx, const int* twrow, cudaStream_t st) {
    MultiParams p{};
    int miny = 9, minx = 9;
    for (int t = 0; t < 9; ++t) { if (tdy[t] < miny) miny = tdy[t]; if (tdx[t] < minx) minx = tdx[t]; }
    p.ntaps = 9;
    for (int t = 0; t < 9; ++t) {
        const int dy = tdy[t] - miny, dx = tdx[t] - minx;
        if (dy > 2 || dx > 2) SHM_FAIL(SHM_EUNSUPPORTED, "conv_big: tap offsets exceed the 1-pixel halo");
        p.wrow[t] = twrow[t]; p.npairs[t] = 2;
        p.aoff[t][0] = dy * HALO_W + dx; p.aoff[t][1] = p.aoff[t][0] + 16 * HALO_W;
        p.acc[t][0] = 0; p.acc[t][1] = 1;
        p.first[t][0] = p.first[t][1] = (t == 0);
    }
    p.nacc = 2; p.row_dy[0] = 0; p.row_dy[1] = 16;
    p.OS = 1; p.TH = 32; p.a_bytes = BIG_H * HALO_W * 128; p.hy = miny; p.hx = minx;
    p.kchunks = K / 64;
    p.tiles_x = W / 8; p.tiles_y = H / 32;
    p.m_tiles = N * p.tiles_x * p.tiles_y; p.n_tiles = Nn / 128;
    p.nstore = Nn;
    p.Hout = H; p.Wout = W; p.ldout = ldout; p.bias = bias; p.act = act; p.out = (bf16*)out;
    CUtensorMap tmA, tmB;
    if (int rc = encode_act_box(&tmA, in, K, W, H, N, ldin, HALO_W, BIG_H)) return rc;
    // CTA-pair form (conv_big2_kernel): every layer with at least one pair of tiles.  SHM_BIG2=0 in the environment keeps the single-CTA
    // kernel (A/B measurements, tests of both paths).
    if (pair_enabled() && p.m_tiles >= 2) {
        if (int rc = encode_w(&tmB, w_tc, K, wrows_total, 64)) return rc;
        return launch_pair_t<2, 2>(tmA, tmB, p, st);
    }
    if (int rc = encode_w(&tmB, w_tc, K, wrows_total, 128)) return rc;
    return launch_multi_t<128, 2, 2>(tmA, tmB, p, st);
}

// "scatter" configuration: out[(q + .) * 2 + r] classes of a stride-2 scatter (Conv2DTranspose fwd, strided-Conv2D dgrad).
// Per class (ry, rx) the caller lists its taps as offsets (dy, dx) in {-1, 0} on the input lattice and their weight rows.
bool scatter_ok(int Hq, int Wq, int K, int Nn) {
    return Hq % 16 == 0 && Wq % 8 == 0 && K % 64 == 0 && Nn % 64 == 0;
}

struct ScatterTap { int ry, rx, dy, dx, wrow; };

int launch_scatter(int N, int Hq, int Wq, int K, int Nn, const void* in, int ldin, const void* w_tc, int wrows_total, const float* bias, int act,
                   void* out, int Hout, int Wout, int ldout, const ScatterTap* taps, int ntaps, cudaStream_t st, int nstore = 0) {
    MultiParams p{};
    p.nstore = nstore > 0 ? nstore : Nn;
    if (ntaps > 9) SHM_FAIL(SHM_EUNSUPPORTED, "conv_scatter: more than 9 taps");
    bool seen[4] = {false, false, false, false};
    p.ntaps = ntaps;
    for (int t = 0; t < ntaps; ++t) {
        const ScatterTap& tp = taps[t];
        if (tp.dy < -1 || tp.dy > 0 || tp.dx < -1 || tp.dx > 0) SHM_FAIL(SHM_EUNSUPPORTED, "conv_scatter: tap offset outside {-1, 0}");
        const int a = tp.ry * 2 + tp.rx;
        p.wrow[t] = tp.wrow; p.npairs[t] = 1;
        p.aoff[t][0] = (tp.dy + 1) * HALO_W + (tp.dx + 1);
        p.acc[t][0] = a; p.first[t][0] = seen[a] ? 0 : 1;
        seen[a] = true;
    }
    for (int a = 0; a < 4; ++a) {
        if (!seen[a]) SHM_FAIL(SHM_EUNSUPPORTED, "conv_scatter: parity class without taps");
        p.row_dy[a] = 0; p.py[a] = a >> 1; p.px[a] = a & 1;
    }
    p.nacc = 4;
    p.OS = 2; p.TH = 16; p.a_bytes = HALO_H * HALO_W * 128; p.hy = -1; p.hx = -1;
    p.kchunks = K / 64;
    p.tiles_x = Wq / 8; p.tiles_y = Hq / 16;
    p.m_tiles = N * p.tiles_x * p.tiles_y;
    p.Hout = Hout; p.Wout = Wout; p.ldout = ldout; p.bias = bias; p.act = act; p.out = (bf16*)out;
    const int BN = (Nn % 128 == 0) ? 128 : 64;
    p.n_tiles = Nn / BN;
    CUtensorMap tmA, tmB;
    // (Tried: each parity class of a 128-column layer as its own pass in the "big" configuration -- two accumulators per streamed weight tile,
    // double-buffered TMEM.  No gain: up3T 0.397 -> 0.363 ms but d3 dgrad 0.316 -> 0.381 ms; a class with one or two taps streams a 43 KB halo
    // tile per 8-16 MMAs and is bound by the L2 -> SM path just like the four-accumulator form.  profiles/r02_negative_results.txt)
    if (int rc = encode_act_box(&tmA, in, K, Wq, Hq, N, ldin, HALO_W, HALO_H)) return rc;
    if (Nn == 64 && p.nstore == 64 && ntaps == 9 && (K == 64 || K == 128)) {   // all nine weight tiles fit in shared memory: resident-weight kernel
        static const bool res_on = []() { const char* e = getenv("SHM_SCAT_RES"); return !(e && e[0] == '0'); }();
        if (res_on) {
            ScatResParams q{};
            bool seen[4] = {false, false, false, false};
            for (int t = 0; t < 9; ++t) {
                const ScatterTap& tp = taps[t];
                const int a = tp.ry * 2 + tp.rx;
                q.tdy[t] = tp.dy + 1; q.tdx[t] = tp.dx + 1; q.wrow[t] = tp.wrow; q.acc[t] = a; q.first[t] = seen[a] ? 0 : 1;
                seen[a] = true;
            }
            q.tiles_x = Wq / 8; q.tiles_y = Hq / 16; q.total_tiles = N * q.tiles_x * q.tiles_y;
            q.bias = bias; q.act = act;
            if (int rc = encode_w(&tmB, w_tc, K, wrows_total, 64)) return rc;
            ScatOutMaps om;
            for (int a = 0; a < 4; ++a) {
                const int py = a >> 1, px = a & 1;
                if (int rc = encode_out_parity(&om.m[a], (const bf16*)out + ((long long)py * Wout + px) * ldout, 64, (Wout - px + 1) / 2, (Hout - py + 1) / 2, N,
                                               Hout, Wout, ldout)) return rc;
            }
            int grid = shm_num_sms();
            if (grid > q.total_tiles) grid = q.total_tiles;
            if (K == 64) {
                static bool attr = false;
                if (!attr) { cudaFuncSetAttribute(conv_scat_res_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, ScatResCfg<1>::SMEM); attr = true; }
                conv_scat_res_kernel<1><<<grid, SCAT_RES_THREADS, ScatResCfg<1>::SMEM, st>>>(tmA, tmB, om, q);
            } else {
                static bool attr = false;
                if (!attr) { cudaFuncSetAttribute(conv_scat_res_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, ScatResCfg<2>::SMEM); attr = true; }
                conv_scat_res_kernel<2><<<grid, SCAT_RES_THREADS, ScatResCfg<2>::SMEM, st>>>(tmA, tmB, om, q);
            }
            SHM_CHECK_LAUNCH("conv_scat_res_kernel");
            return SHM_OK;
        }
    }
    if (BN == 128 && pair_enabled() && p.m_tiles >= 2) {                // CTA pair: each SM streams half of every weight tile
        if (int rc = encode_w(&tmB, w_tc, K, wrows_total, 64)) return rc;
        // two items per tile: parity classes {(0,0), (1,1)} (4 + 1 taps of a 3 x 3 kernel) and {(0,1), (1,0)} (2 + 2 taps).  Two accumulators per
        // item leave room for a second set in TMEM, so the epilogue of one item overlaps the MMAs of the next (with all four classes in one item
        // the set filled TMEM and the tensor pipe idled through every epilogue: 32-34 % active).  The halo tile is loaded once per half.
        static const bool halves_on = []() { const char* e = getenv("SHM_SCAT_HALVES"); return !(e && e[0] == '0'); }();
        if (halves_on) {
            MultiParams q = p;
            const int cls_half[4] = {0, 1, 1, 0}, cls_acc[4] = {0, 0, 1, 1};          // class 2*ry+rx -> (half, accumulator)
            for (int t = 0; t < ntaps; ++t) {
                const int a = taps[t].ry * 2 + taps[t].rx;
                q.half[t] = cls_half[a]; q.acc[t][0] = cls_acc[a];
            }
            q.halves = 2; q.nacc = 2;
            const int order[4] = {0, 3, 1, 2};                                         // [half * 2 + accumulator] -> class
            for (int c = 0; c < 4; ++c) { q.row_dy[c] = 0; q.py[c] = order[c] >> 1; q.px[c] = order[c] & 1; }
            return launch_pair_t<2, 1>(tmA, tmB, q, st);
        }
        return launch_pair_t<1, 1>(tmA, tmB, p, st);
    }
    if (int rc = encode_w(&tmB, w_tc, K, wrows_total, BN)) return rc;
    if (BN == 128) return launch_multi_t<128, 1, 1>(tmA, tmB, p, st);   // 4 x 128 columns: one accumulator set
    return launch_multi_t<64, 2, 1>(tmA, tmB, p, st);                   // 4 x 64 columns, double-buffered
}

// thin stride-1 3x3 layers (fewer than 64 reduction or output channels) the halo kernel serves with 32- / 64-byte pixel rows
bool thin_ok(int H, int W, int K, int Nn, int kh, int kw, int stride) {
    if (!(kh == 3 && kw == 3 && stride == 1 && H % 16 == 0 && W % 8 == 0)) return false;
    if (K == 16) return Nn == 16 || Nn == 32 || Nn == 64 || Nn == 128;
    if (K == 32) return Nn == 16 || Nn == 32 || Nn == 64;
    if (K == 64) return Nn == 16 || Nn == 32;
    return false;
}

template <int KC, int BN, int CPX = 128, int T = 1>
int launch_halo_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const HaloParams& p, cudaStream_t st) {
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(conv_halo_kernel<KC, BN, CPX, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, HaloCfg<KC, BN, CPX, T>::SMEM); attr = true; }
    int grid = shm_num_sms();
    if (grid > p.total_tiles) grid = p.total_tiles;
    // output map for the TMA-store epilogue: (BN channels, W, H, N) over the destination slice, box = one 16 x 8-pixel sub-tile
    CUtensorMap tmO = tmA;
    if (HaloCfg<KC, BN, CPX, T>::TSTORE) {
        if (int rc = encode_act_box(&tmO, p.out, BN, p.W, p.H, p.total_tiles / (p.tiles_x * p.tiles_y), p.ldout, 8, 16, BN)) return rc;
    }
    conv_halo_kernel<KC, BN, CPX, T><<<grid, HaloCfg<KC, BN, CPX, T>::THREADS, HaloCfg<KC, BN, CPX, T>::SMEM, st>>>(tmA, tmB, tmO, p);
    SHM_CHECK_LAUNCH("conv_halo_kernel");
    return SHM_OK;
}
// thin layer: stacked sub-tiles per work item when the image height allows it
template <int BN, int CPX>
int launch_thin(const CUtensorMap& tmA1, const CUtensorMap& tmAT, const CUtensorMap& tmB, HaloParams& p, bool stack, cudaStream_t st) {
    constexpr int T = (BN <= 64 && CPX < 128) ? 4 : 2;
    if (!stack) return launch_halo_t<1, BN, CPX, 1>(tmA1, tmB, p, st);
    p.tiles_y /= T; p.total_tiles /= T;
    return launch_halo_t<1, BN, CPX, T>(tmAT, tmB, p, st);
}

// in: [N,H,W,K] (ld ldin), out: [N,H,W,Nn] (ld ldout); taps (tdy, tdx) in {-1,0,1} with weight rows twrow
// which halo-kernel variants carry the statistics epilogue (EpiStats): the un-stacked 64-column ones -- 64 -> 64 (TMA-store epilogue) and
// 128 -> 64 (two k-chunks, direct stores)
bool halo_stats_ok(int H, int K, int Nn) {
    (void)H;
    return Nn == 64 && (K == 64 || K == 128);
}

int launch_halo(int N, int H, int W, int K, int Nn, const void* in, int ldin, const void* w_tc, int wrows_total, const float* bias, int act,
                void* out, int ldout, const int* tdy, const int* tdx, const int* twrow, cudaStream_t st, double* stats = nullptr) {
    HaloParams p{};
    if (stats != nullptr && !halo_stats_ok(H, K, Nn)) SHM_FAIL(SHM_EUNSUPPORTED, "conv_halo: this variant (K=%d, Nn=%d) has no statistics epilogue", K, Nn);
    p.stats = stats;
    int miny = 9, minx = 9;
    for (int t = 0; t < 9; ++t) { if (tdy[t] < miny) miny = tdy[t]; if (tdx[t] < minx) minx = tdx[t]; }
    for (int t = 0; t < 9; ++t) {
        p.tdy[t] = tdy[t] - miny; p.tdx[t] = tdx[t] - minx; p.wrow[t] = twrow[t];
        if (p.tdy[t] > 2 || p.tdx[t] > 2) SHM_FAIL(SHM_EUNSUPPORTED, "conv_halo: tap offsets exceed the 1-pixel halo");
    }
    p.oy = miny; p.ox = minx;
    p.tiles_x = W / 8; p.tiles_y = H / 16; p.total_tiles = N * p.tiles_x * p.tiles_y;
    p.H = H; p.W = W; p.ldout = ldout; p.Nn = Nn; p.bias = bias; p.act = act; p.out = (bf16*)out;
    CUtensorMap tmA, tmB;
    const int inner = K < 64 ? K : 64;
    if (int rc = encode_act_box(&tmA, in, K, W, H, N, ldin, HALO_W, HALO_H, inner)) return rc;
    if (int rc = encode_w(&tmB, w_tc, K, wrows_total, Nn, inner)) return rc;
    if (K < 64 || Nn < 64) {
        const int T = (Nn <= 64 && K < 64) ? 4 : 2;
        const bool stack = H % (16 * T) == 0;
        CUtensorMap tmS = tmA;
        if (stack) { if (int rc = encode_act_box(&tmS, in, K, W, H, N, ldin, HALO_W, 16 * T + 2, inner)) return rc; }
        if (K == 16) {
            if (Nn == 16) return launch_thin<16, 32>(tmA, tmS, tmB, p, stack, st);
            if (Nn == 32) return launch_thin<32, 32>(tmA, tmS, tmB, p, stack, st);
            if (Nn == 64) return launch_thin<64, 32>(tmA, tmS, tmB, p, stack, st);
            if (Nn == 128) return launch_thin<128, 32>(tmA, tmS, tmB, p, stack, st);
        } else if (K == 32) {
            if (Nn == 16) return launch_thin<16, 64>(tmA, tmS, tmB, p, stack, st);
            if (Nn == 32) return launch_thin<32, 64>(tmA, tmS, tmB, p, stack, st);
            if (Nn == 64) return launch_thin<64, 64>(tmA, tmS, tmB, p, stack, st);
        } else if (K == 64 && Nn == 16) return launch_thin<16, 128>(tmA, tmS, tmB, p, stack, st);
        else if (K == 64 && Nn == 32) return launch_thin<32, 128>(tmA, tmS, tmB, p, stack, st);
        SHM_FAIL(SHM_EUNSUPPORTED, "conv_halo (thin): K=%d Nn=%d", K, Nn);
    }
    if (K == 64 && Nn == 64) return launch_halo_t<1, 64>(tmA, tmB, p, st);
    if (K == 128 && Nn == 64) return launch_halo_t<2, 64>(tmA, tmB, p, st);
    if (K == 64 && Nn == 128) return launch_halo_t<1, 128>(tmA, tmB, p, st);
    SHM_FAIL(SHM_EUNSUPPORTED, "conv_halo: K=%d Nn=%d", K, Nn);
}


// ------------------------------------------------------------------------------------------------
// wgrad kernel:  dW[tap][a][b] += sum_{lattice q} A[q*SA + offA_tap, a] * B[q*SB + offB_tap, b]
//   GEMM M = 128 "a" channels (two 64-channel M-blocks; for 64-channel layers the two blocks are two TAPS),
//   N = BN "b" channels, K = lattice points in chunks of 64 (split over CTAs; fp32 atomics into dW).
//   Both operands are MN-major in shared memory: a TMA box is [64 points][64 channels] = rows of 128 bytes.
// ------------------------------------------------------------------------------------------------
struct WgParams {
    int nblocks;                       // M-blocks = ntaps * Ca/64
    int cblocks;                       // Ca / 64
    int ntaps; int day[9], dax[9], dby[9], dbx[9]; long long woff[9];
    int Qh, Qw, BW, BH, BI, tiles_x, tiles_y;
    int SA, SB;
    int kblocks, kb_per_split;
    int m_tiles, n_tiles;
    int ldw;                           // elements between consecutive "a" rows of dW (= Cb)
    float* dW;
};

constexpr int WG_STAGES = 6;
template <int BN>
struct WgCfg {
    static constexpr int A_ST = 2 * 8192;
    static constexpr int B_ST = (BN / 64) * 8192;
    static constexpr int STAGE_BYTES = A_ST + B_ST;
    static constexpr int SMEM = WG_STAGES * STAGE_BYTES + 1024 + 256;
};

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const WgParams p) {
    using Cfg = WgCfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = smem + WG_STAGES * Cfg::A_ST;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WG_STAGES * Cfg::STAGE_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + WG_STAGES;
    uint64_t* tfull = bars + 2 * WG_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nt = blockIdx.x % p.n_tiles;
    const int mt = (blockIdx.x / p.n_tiles) % p.m_tiles;
    const int split = blockIdx.x / (p.n_tiles * p.m_tiles);
    const int kb0 = split * p.kb_per_split;
    int kb1 = kb0 + p.kb_per_split; if (kb1 > p.kblocks) kb1 = p.kblocks;
    // the two M-blocks of this tile: (tap, channel block)
    const int blk0 = mt * 2, blk1 = (mt * 2 + 1 < p.nblocks) ? mt * 2 + 1 : mt * 2;
    const int tap0 = blk0 / p.cblocks, cb0 = blk0 % p.cblocks, tap1 = blk1 / p.cblocks, cb1 = blk1 % p.cblocks;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA); prefetch_tmap(&tmB);
        for (int s = 0; s < WG_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tfull, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, BN < 32 ? 32 : BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        {
            int stage = 0; uint32_t phase = 0;
            const int per_img = p.tiles_x * p.tiles_y;
            for (int kb = kb0; kb < kb1; ++kb) {
                int img0, qy0, qx0;
                if (p.BI > 1) { img0 = kb * p.BI; qy0 = 0; qx0 = 0; }
                else { img0 = kb / per_img; const int r = kb - img0 * per_img; qy0 = (r / p.tiles_x) * p.BH; qx0 = (r % p.tiles_x) * p.BW; }
                mbar_wait(&empty[stage], phase ^ 1);
                if (elect_one_sync()) {
                    mbar_expect_tx(&full[stage], Cfg::STAGE_BYTES);
                    uint8_t* a = sA + stage * Cfg::A_ST;
                    uint8_t* b = sB + stage * Cfg::B_ST;
                    tma_load_4d(a, &tmA, &full[stage], cb0 * 64, qx0 * p.SA + p.dax[tap0], qy0 * p.SA + p.day[tap0], img0);
                    tma_load_4d(a + 8192, &tmA, &full[stage], cb1 * 64, qx0 * p.SA + p.dax[tap1], qy0 * p.SA + p.day[tap1], img0);
                    // the B operand is shifted by its own tap offset (zero for Conv2D where B = dy): both M-blocks of a tile must
                    // therefore share the B offset, which holds because offB != 0 only when cblocks >= 2 pairs blocks of ONE tap
#pragma unroll
                    for (int j = 0; j < BN / 64; ++j)
                        tma_load_4d(b + j * 8192, &tmB, &full[stage], nt * BN + j * 64, qx0 * p.SB + p.dbx[tap0], qy0 * p.SB + p.dby[tap0], img0);
                }
                __syncwarp();
                if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = make_idesc(128, BN, 1, 1);
        int stage = 0; uint32_t phase = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            if (elect_one_sync()) {
                // MN-major SW128: 64-channel atoms LBO = 8192 bytes apart, 8-point K groups SBO = 1024 bytes apart
                const uint64_t adesc = make_desc_sw128(smem_u32(sA + stage * Cfg::A_ST), 8192, 1024);
                const uint64_t bdesc = make_desc_sw128(smem_u32(sB + stage * Cfg::B_ST), 8192, 1024);
#pragma unroll
                for (int k = 0; k < 4; ++k)      // 16 points per MMA = 2048 bytes
                    umma_bf16(tmem_base, adesc + (uint64_t)(k * 128), bdesc + (uint64_t)(k * 128), idesc, (kb > kb0) || (k > 0));
                umma_commit(&empty[stage]);
            }
            __syncwarp();
            if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
        }
        if (elect_one_sync()) umma_commit(tfull);
        __syncwarp();
    } else if (kb1 > kb0) {
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const int half = row >> 6;
        const bool ok = !(half == 1 && blk1 == blk0);          // odd M-block count: upper half is a duplicate
        const int tap = half ? tap1 : tap0, cb = half ? cb1 : cb0;
        float* dst = p.dW + p.woff[tap] + (long long)(cb * 64 + (row & 63)) * p.ldw + nt * BN;
        mbar_wait(tfull, 0);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            uint32_t r[32];
            tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(c * 32), r);
            if (ok) {
#pragma unroll
                for (int j = 0; j < 32; ++j) atomicAdd(dst + c * 32 + j, __uint_as_float(r[j]));
            }
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, BN < 32 ? 32 : BN);
}

// ------------------------------------------------------------------------------------------------
// halo wgrad kernel: stride-1 3x3 Conv2D weight gradients (every 3x3 layer of the generator).
//
//   dW[ky][kx][ci][co] += sum_q x[q + (ky-1, kx-1), ci] * dy[q, co]
//
// The generic kernel above re-reads the x tile once per tap and the dy tile once per (tap, ci-block): at 128 x 128 x 64 per stage
// it needs 128 B/clk/SM from L2 and runs at 220-600 TFLOP/s.  Here one TMA box brings the (16+2) x (8+2)-pixel halo of a
// 16 x 8-pixel tile ONCE, and every tap is a shifted MN-major smem descriptor into it (start + (ty*10+tx) rows, 8-point K groups
// SBO = 10 rows apart), against ONE dy tile.  All the taps a CTA owns accumulate in TMEM over the CTA's whole pixel range
// (split-K over pixels across CTAs), then one vectorised red.global.add epilogue.
//   MODE 0 (Cin or Cout == 64):  unit = (64 ci, 64 co), M = 128 = TWO TAPS (the second tap is LBO = tap distance away in the
//           same halo), N = 64, five accumulators (taps 01 23 45 67 8-) = 320 TMEM columns; 39 KB of operands per 40 MMAs.
//   MODE 1 (Cin, Cout % 128 == 0): unit = (128 ci, 128 co, one filter row), M = 128 = two 64-channel halos, N = 128,
//           three accumulators = 384 TMEM columns; 72 KB of operands per 24 MMAs (47 B/clk/SM instead of 128).
// ------------------------------------------------------------------------------------------------
struct WhParams {
    int units, splits;                 // grid = units * splits
    int cblocks, nblocks;              // MODE 0: Cin/64, Cout/64;  MODE 1: Cin/128, Cout/128
    int tiles_x, tiles_y, total_tiles, tiles_per_split;
    int Cin, Cout;
    float* dW;
};

template <int MODE>
struct WhCfg {
    static constexpr int A_ROWS = MODE == 0 ? 18 : 16;
    static constexpr int A_ONE = MODE == 0 ? HALO_STAGE : 16 * HALO_W * 128;       // 23552 / 20480 (both multiples of 1024)
    static constexpr int A_BYTES_TX = (MODE == 0 ? 1 : 2) * A_ROWS * HALO_W * 128;  // bytes TMA actually writes
    static constexpr int A_ST = (MODE == 0 ? 1 : 2) * A_ONE;
    static constexpr int B_ST = (MODE == 0 ? 1 : 2) * 16384;
    static constexpr int STAGE_BYTES = A_ST + B_ST;
    static constexpr int STAGES = MODE == 0 ? 5 : 3;
    static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + 256;
    static constexpr int BN = MODE == 0 ? 64 : 128;
    static constexpr int NACC = MODE == 0 ? 5 : 3;
};

__device__ __forceinline__ void red_add_v4(float* dst, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1)
wgrad_halo_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY, const WhParams p) {
    using Cfg = WhCfg<MODE>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = smem + Cfg::STAGES * Cfg::A_ST;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + Cfg::STAGES;
    uint64_t* tfull = bars + 2 * Cfg::STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int unit = blockIdx.x % p.units, split = blockIdx.x / p.units;
    // unit -> (ci block, co block, filter row)
    const int nb = unit % p.nblocks;
    const int cb = (unit / p.nblocks) % p.cblocks;
    const int frow = unit / (p.nblocks * p.cblocks);            // MODE 1 only (0..2)
    const int t0 = split * p.tiles_per_split;
    int t1 = t0 + p.tiles_per_split; if (t1 > p.total_tiles) t1 = p.total_tiles;
    const int per_img = p.tiles_x * p.tiles_y;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmX); prefetch_tmap(&tmDY);
        for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tfull, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        {
            int stage = 0; uint32_t phase = 0;
            for (int t = t0; t < t1; ++t) {
                const int img = t / per_img; const int r = t - img * per_img;
                const int y0 = (r / p.tiles_x) * 16, x0 = (r % p.tiles_x) * 8;
                mbar_wait(&empty[stage], phase ^ 1);
                if (elect_one_sync()) {
                    mbar_expect_tx(&full[stage], Cfg::A_BYTES_TX + Cfg::B_ST);
                    uint8_t* a = sA + stage * Cfg::A_ST;
                    uint8_t* b = sB + stage * Cfg::B_ST;
                    if (MODE == 0) {
                        tma_load_4d(a, &tmX, &full[stage], cb * 64, x0 - 1, y0 - 1, img);
                        tma_load_4d(b, &tmDY, &full[stage], nb * 64, x0, y0, img);
                    } else {
                        tma_load_4d(a, &tmX, &full[stage], cb * 128, x0 - 1, y0 - 1 + frow, img);
                        tma_load_4d(a + Cfg::A_ONE, &tmX, &full[stage], cb * 128 + 64, x0 - 1, y0 - 1 + frow, img);
                        tma_load_4d(b, &tmDY, &full[stage], nb * 128, x0, y0, img);
                        tma_load_4d(b + 16384, &tmDY, &full[stage], nb * 128 + 64, x0, y0, img);
                    }
                }
                __syncwarp();
                if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (t1 > t0) {
            constexpr uint32_t idesc = make_idesc(128, Cfg::BN, 1, 1);
            int stage = 0; uint32_t phase = 0;
            for (int t = t0; t < t1; ++t) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint32_t a0 = smem_u32(sA + stage * Cfg::A_ST);
                    const uint32_t b0 = smem_u32(sB + stage * Cfg::B_ST);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {          // 16 pixels = tile rows 2j, 2j+1 per MMA
                        const uint64_t bdesc = make_desc_sw128(b0 + j * 2048, 16384, 1024);
                        const uint32_t acc = (t > t0 || j > 0) ? 1u : 0u;
                        if (MODE == 0) {
#pragma unroll
                            for (int pr = 0; pr < 5; ++pr) {
                                const int ta = 2 * pr, tb = pr < 4 ? 2 * pr + 1 : 8;
                                const int offa = (ta / 3) * HALO_W + ta % 3, offb = (tb / 3) * HALO_W + tb % 3;
                                const uint64_t adesc = make_desc_sw128(a0 + (uint32_t)(2 * j * HALO_W + offa) * 128u, (uint32_t)(offb - offa) * 128u, HALO_W * 128);
                                umma_bf16(tmem_base + pr * 64, adesc, bdesc, idesc, acc);
                            }
                        } else {
#pragma unroll
                            for (int tx = 0; tx < 3; ++tx) {
                                const uint64_t adesc = make_desc_sw128(a0 + (uint32_t)(2 * j * HALO_W + tx) * 128u, Cfg::A_ONE, HALO_W * 128);
                                umma_bf16(tmem_base + tx * 128, adesc, bdesc, idesc, acc);
                            }
                        }
                    }
                    umma_commit(&empty[stage]);
                }
                __syncwarp();
                if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
            }
            if (elect_one_sync()) umma_commit(tfull);
            __syncwarp();
        }
    } else if (t1 > t0) {
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        mbar_wait(tfull, 0);
        tc_fence_after();
#pragma unroll 1
        for (int a = 0; a < Cfg::NACC; ++a) {
            int tap, ci;
            bool ok = true;
            if (MODE == 0) { tap = 2 * a + (row >> 6); ci = cb * 64 + (row & 63); ok = tap < 9; }
            else { tap = frow * 3 + a; ci = cb * 128 + row; }
            float* dst = p.dW + ((long long)tap * p.Cin + ci) * p.Cout + nb * Cfg::BN;
#pragma unroll 1
            for (int c = 0; c < Cfg::BN / 32; ++c) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(a * Cfg::BN + c * 32), r);
                if (ok) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        red_add_v4(dst + c * 32 + j, __uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                }
            }
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// CTA-pair form of wgrad_halo_kernel<1> (Cin % 256 == 0, Cout % 128 == 0): M = 256 input channels over the pair (each CTA loads the halo of its own
// 128 channels), N = 128 output channels of ONE dy tile of which each CTA loads 64 -- tcgen05.mma.cta_group::2 retires an M256 x N128 x K16 MMA
// every 64 cycles against 73 for M128 x N128 on one SM (tools/umma_rate_probe.cu), and the dy stream per SM halves.  Barrier protocol as in
// conv_big2_kernel: all loads signal the leader's full barriers (one arrival: the leader's expect_tx for both CTAs' bytes), commits are multicast.
// There is one accumulator set per CTA for the whole kernel (split-K over the pixel range), so no accumulator hand-back is needed.
// ------------------------------------------------------------------------------------------------
constexpr int WH2_A_ONE = 16 * HALO_W * 128;                   // 20480: one 64-channel halo (16 rows of 10 pixels)
constexpr int WH2_A_ST = 2 * WH2_A_ONE;
constexpr int WH2_B_ST = 16384;                                // this CTA's 64 output channels of the 128-pixel dy tile
constexpr int WH2_STAGE = WH2_A_ST + WH2_B_ST;
constexpr int WH2_STAGES = 3;
constexpr int WH2_SMEM = WH2_STAGES * WH2_STAGE + 1024 + 256;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
wgrad_halo2_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY, const WhParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = smem + WH2_STAGES * WH2_A_ST;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WH2_STAGES * WH2_STAGE);
    uint64_t* full = bars;                                 // leader only
    uint64_t* empty = bars + WH2_STAGES;
    uint64_t* tfull = bars + 2 * WH2_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cid = blockIdx.x >> 1;
    const int unit = cid % p.units, split = cid / p.units;          // p.units counts PAIRS of 128-channel ci blocks here
    const int nb = unit % p.nblocks;
    const int cb = 2 * ((unit / p.nblocks) % (p.cblocks >> 1)) + (int)rank;
    const int frow = unit / (p.nblocks * (p.cblocks >> 1));
    const int t0 = split * p.tiles_per_split;
    int t1 = t0 + p.tiles_per_split; if (t1 > p.total_tiles) t1 = p.total_tiles;
    const int per_img = p.tiles_x * p.tiles_y;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmX); prefetch_tmap(&tmDY);
        for (int s = 0; s < WH2_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tfull, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_2sm(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        int stage = 0; uint32_t phase = 0;
        for (int t = t0; t < t1; ++t) {
            const int img = t / per_img; const int r = t - img * per_img;
            const int y0 = (r / p.tiles_x) * 16, x0 = (r % p.tiles_x) * 8;
            mbar_wait(&empty[stage], phase ^ 1);
            if (elect_one_sync()) {
                const uint32_t lead = map_to_cta(&full[stage], 0);
                if (rank == 0) mbar_expect_tx(&full[stage], 2 * (WH2_A_ST + WH2_B_ST));
                uint8_t* a = sA + stage * WH2_A_ST;
                tma_load_4d_2sm(a, &tmX, lead, cb * 128, x0 - 1, y0 - 1 + frow, img);
                tma_load_4d_2sm(a + WH2_A_ONE, &tmX, lead, cb * 128 + 64, x0 - 1, y0 - 1 + frow, img);
                tma_load_4d_2sm(sB + stage * WH2_B_ST, &tmDY, lead, nb * 128 + (int)rank * 64, x0, y0, img);
            }
            __syncwarp();
            if (++stage == WH2_STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1 && rank == 0) {
        if (t1 > t0) {
            constexpr uint32_t idesc = make_idesc(256, 128, 1, 1);
            int stage = 0; uint32_t phase = 0;
            for (int t = t0; t < t1; ++t) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint32_t a0 = smem_u32(sA + stage * WH2_A_ST);
                    const uint32_t b0 = smem_u32(sB + stage * WH2_B_ST);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {          // 16 pixels = tile rows 2j, 2j+1 per MMA
                        const uint64_t bdesc = make_desc_sw128(b0 + j * 2048, 16384, 1024);
                        const uint32_t acc = (t > t0 || j > 0) ? 1u : 0u;
#pragma unroll
                        for (int tx = 0; tx < 3; ++tx) {
                            const uint64_t adesc = make_desc_sw128(a0 + (uint32_t)(2 * j * HALO_W + tx) * 128u, WH2_A_ONE, HALO_W * 128);
                            umma_bf16_2sm(tmem_base + tx * 128, adesc, bdesc, idesc, acc);
                        }
                    }
                    umma_commit_2sm(&empty[stage]);
                }
                __syncwarp();
                if (++stage == WH2_STAGES) { stage = 0; phase ^= 1; }
            }
            if (elect_one_sync()) umma_commit_2sm(tfull);
            __syncwarp();
        }
    } else if (warp >= 2 && t1 > t0) {
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        mbar_wait(tfull, 0);
        tc_fence_after();
#pragma unroll 1
        for (int a = 0; a < 3; ++a) {
            const int tap = frow * 3 + a, ci = cb * 128 + row;
            float* dst = p.dW + ((long long)tap * p.Cin + ci) * p.Cout + nb * 128;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(a * 128 + c * 32), r);
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    red_add_v4(dst + c * 32 + j, __uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
            }
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) tmem_dealloc_2sm(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// stride-2 halo wgrad kernel: weight gradients of Conv2DTranspose(3, s=2) and of stride-2 Conv2D(3) layers.
//
//   dW[ky][kx][cg][cs] += sum_o G[2*o + (ky, kx), cg] * S[o, cs]
//
// G = the big image (dy of a transposed conv / x of a strided conv), S = the small one (x / dy), o runs over S's lattice.
// The generic wgrad kernel re-gathers G at stride 2 once per tap (350-450 TFLOP/s).  Here the four PARITY PLANES of G's halo are
// loaded once per 16 x 8 tile of o by TMA boxes with traversal stride 2 (dense 18 x 10 smem tiles), and tap (ky, kx) is a shifted
// MN-major descriptor into plane (ky & 1, kx & 1) at offset (ky >> 1, kx >> 1), exactly as in wgrad_halo_kernel:
//   MODE 0: unit = (64 cg, 64 cs); M = 128 = two taps of one plane (LBO = tap distance), five accumulators x 64 columns
//   MODE 1: unit = (128 cg, 128 cs, filter row); M = 128 = two 64-channel tiles of a plane, three accumulators x 128 columns
// ------------------------------------------------------------------------------------------------
struct Ws2Params {
    int units, splits, cblocks, nblocks;
    int tiles_x, tiles_y, total_tiles, tiles_per_split;
    int CG, CS;
    float* dW;
};

template <int MODE>
struct Ws2Cfg {
    static constexpr int A_ONE = MODE == 0 ? HALO_STAGE : 16 * HALO_W * 128;             // one plane tile: 18 x 10 or 16 x 10 rows
    static constexpr int A_TX = 4 * (MODE == 0 ? HALO_BYTES : 16 * HALO_W * 128);        // bytes written by TMA
    static constexpr int A_ST = 4 * A_ONE;
    static constexpr int B_ST = (MODE == 0 ? 1 : 2) * 16384;
    static constexpr int STAGE_BYTES = A_ST + B_ST;
    static constexpr int STAGES = 2;
    static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + 256;
    static constexpr int BN = MODE == 0 ? 64 : 128;
    static constexpr int NACC = MODE == 0 ? 5 : 3;
};

template <int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1)
wgrad_s2_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmS, const Ws2Params p) {
    using Cfg = Ws2Cfg<MODE>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = smem + Cfg::STAGES * Cfg::A_ST;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + Cfg::STAGES;
    uint64_t* tfull = bars + 2 * Cfg::STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int unit = blockIdx.x % p.units, split = blockIdx.x / p.units;
    const int nb = unit % p.nblocks;
    const int cb = (unit / p.nblocks) % p.cblocks;
    const int frow = unit / (p.nblocks * p.cblocks);            // MODE 1: filter row ky
    const int t0 = split * p.tiles_per_split;
    int t1 = t0 + p.tiles_per_split; if (t1 > p.total_tiles) t1 = p.total_tiles;
    const int per_img = p.tiles_x * p.tiles_y;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmG); prefetch_tmap(&tmS);
        for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tfull, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        int stage = 0; uint32_t phase = 0;
        for (int t = t0; t < t1; ++t) {
            const int img = t / per_img; const int r = t - img * per_img;
            const int y0 = (r / p.tiles_x) * 16, x0 = (r % p.tiles_x) * 8;
            mbar_wait(&empty[stage], phase ^ 1);
            if (elect_one_sync()) {
                mbar_expect_tx(&full[stage], Cfg::A_TX + Cfg::B_ST);
                uint8_t* a = sA + stage * Cfg::A_ST;
                uint8_t* b = sB + stage * Cfg::B_ST;
                if (MODE == 0) {
                    // planes (py, px) = (0,0) (0,1) (1,0) (1,1)
#pragma unroll
                    for (int pl = 0; pl < 4; ++pl)
                        tma_load_4d(a + pl * Cfg::A_ONE, &tmG, &full[stage], cb * 64, 2 * x0 + (pl & 1), 2 * y0 + (pl >> 1), img);
                    tma_load_4d(b, &tmS, &full[stage], nb * 64, x0, y0, img);
                } else {
                    // planes (frow & 1, px) for px = 0, 1, rows starting at (frow >> 1); two 64-channel tiles each
                    const int gy = 2 * (y0 + (frow >> 1)) + (frow & 1);
#pragma unroll
                    for (int px = 0; px < 2; ++px)
#pragma unroll
                        for (int h = 0; h < 2; ++h)
                            tma_load_4d(a + (px * 2 + h) * Cfg::A_ONE, &tmG, &full[stage], cb * 128 + h * 64, 2 * x0 + px, gy, img);
                    tma_load_4d(b, &tmS, &full[stage], nb * 128, x0, y0, img);
                    tma_load_4d(b + 16384, &tmS, &full[stage], nb * 128 + 64, x0, y0, img);
                }
            }
            __syncwarp();
            if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1) {
        if (t1 > t0) {
            constexpr uint32_t idesc = make_idesc(128, Cfg::BN, 1, 1);
            int stage = 0; uint32_t phase = 0;
            for (int t = t0; t < t1; ++t) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint32_t a0 = smem_u32(sA + stage * Cfg::A_ST);
                    const uint32_t b0 = smem_u32(sB + stage * Cfg::B_ST);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {          // 16 lattice points = tile rows 2j, 2j+1 per MMA
                        const uint64_t bdesc = make_desc_sw128(b0 + j * 2048, 16384, 1024);
                        const uint32_t acc = (t > t0 || j > 0) ? 1u : 0u;
                        const uint32_t rowoff = (uint32_t)(2 * j * HALO_W) * 128u;
                        if (MODE == 0) {
                            // (plane, first offset, LBO in rows): taps {0,2} {6,8} {1,7} {3,5} {4,-}
                            constexpr int PL[5] = {0, 0, 1, 2, 3};
                            constexpr int OFF[5] = {0, HALO_W, 0, 0, 0};
                            constexpr int LBO[5] = {1, 1, HALO_W, 1, 0};
#pragma unroll
                            for (int pr = 0; pr < 5; ++pr) {
                                const uint64_t adesc = make_desc_sw128(a0 + PL[pr] * Cfg::A_ONE + rowoff + OFF[pr] * 128u, LBO[pr] * 128u, HALO_W * 128);
                                umma_bf16(tmem_base + pr * 64, adesc, bdesc, idesc, acc);
                            }
                        } else {
#pragma unroll
                            for (int kx = 0; kx < 3; ++kx) {
                                const uint32_t base = a0 + ((kx & 1) * 2) * Cfg::A_ONE + rowoff + (kx >> 1) * 128u;
                                const uint64_t adesc = make_desc_sw128(base, Cfg::A_ONE, HALO_W * 128);
                                umma_bf16(tmem_base + kx * 128, adesc, bdesc, idesc, acc);
                            }
                        }
                    }
                    umma_commit(&empty[stage]);
                }
                __syncwarp();
                if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
            }
            if (elect_one_sync()) umma_commit(tfull);
            __syncwarp();
        }
    } else if (t1 > t0) {
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        mbar_wait(tfull, 0);
        tc_fence_after();
#pragma unroll 1
        for (int a = 0; a < Cfg::NACC; ++a) {
            int tap, cg;
            bool ok = true;
            if (MODE == 0) {
                const int TA[5] = {0, 6, 1, 3, 4}, TB[5] = {2, 8, 7, 5, -1};
                tap = (row >> 6) ? TB[a] : TA[a];
                ok = tap >= 0;
                cg = cb * 64 + (row & 63);
            } else { tap = frow * 3 + a; cg = cb * 128 + row; }
            float* dst = p.dW + ((long long)(ok ? tap : 0) * p.CG + cg) * p.CS + nb * Cfg::BN;
#pragma unroll 1
            for (int c = 0; c < Cfg::BN / 32; ++c) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(a * Cfg::BN + c * 32), r);
                if (ok) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        red_add_v4(dst + c * 32 + j, __uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                }
            }
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

bool pick_box64(int N, int Qh, int Qw, int& BW, int& BH, int& BI) {
    BW = Qw >= 64 ? 64 : Qw;
    if (BW <= 0 || 64 % BW != 0 || Qw % BW != 0) return false;
    BH = 64 / BW;
    if (BH > Qh) BH = Qh;
    if (Qh % BH != 0) return false;
    BI = 64 / (BW * BH);
    if (BI * BW * BH != 64) return false;
    if (BI > 1 && (N % BI != 0)) return false;
    return true;
}

// weights -> bf16 [tap][n][k].  The device reduction dimension may be wider than the Keras one: device channel k maps to source
// channel (k / seg_pad) * seg_real + (k % seg_pad) when (k % seg_pad) < seg_real (zero otherwise), which covers a zero-padded
// input (one segment) and a concat of two zero-padded halves (two segments, SpecSeg decoder).
__global__ void prep_w_kernel(const float* __restrict__ w, bf16* __restrict__ o, int K, int Nn, int k_real, int n_real, long long tap_elems,
                              int w_ks, int w_ns, int seg_real, int seg_pad, long long total) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i % K);
        const long long t2 = i / K;
        const int n = (int)(t2 % Nn);
        const long long tap = t2 / Nn;
        const int kin = k % seg_pad, ks = (k / seg_pad) * seg_real + kin;
        const float v = (kin < seg_real && ks < k_real && n < n_real) ? __ldg(w + tap * tap_elems + (long long)ks * w_ks + (long long)n * w_ns) : 0.f;
        o[i] = __float2bfloat16_rn(v);
    }
}

// both layouts of one layer in ONE launch (blockIdx.y = 0: forward [tap][Cout][Cin], 1: dgrad [tap][Cin][Cout]): the per-step weight
// refresh of ~70 layers was 140 tiny launches
struct PrepSet { bf16* o; int K, Nn, k_real, n_real, w_ks, w_ns; };
__global__ void prep_w_both_kernel(const float* __restrict__ w, PrepSet a, PrepSet b, long long tap_elems, long long total) {
    const PrepSet s = blockIdx.y == 0 ? a : b;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i % s.K);
        const long long t2 = i / s.K;
        const int n = (int)(t2 % s.Nn);
        const long long tap = t2 / s.Nn;
        const float v = (k < s.k_real && n < s.n_real) ? __ldg(w + tap * tap_elems + (long long)k * s.w_ks + (long long)n * s.w_ns) : 0.f;
        s.o[i] = __float2bfloat16_rn(v);
    }
}

// the same for MANY layers in one launch: the per-step weight refresh of a network (~40 layers) is launch-bound otherwise.
// blockIdx.x -> (job, chunk of 2048 elements) through the jobs' block_begin prefix; blockIdx.y = layout (0 forward, 1 dgrad).
struct PrepJob { const float* w; PrepSet a, b; long long tap_elems, total; int block_begin, nblocks; };
constexpr int PREP_CHUNK = 8192;
__global__ void __launch_bounds__(256) prep_w_multi_kernel(const PrepJob* __restrict__ jobs, int njobs) {
    __shared__ float tile[PREP_CHUNK + 128];
    int lo = 0, hi = njobs - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (jobs[mid].block_begin <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    const PrepJob J = jobs[lo];
    const PrepSet s = blockIdx.y == 0 ? J.a : J.b;
    const long long beg = (long long)((int)blockIdx.x - J.block_begin) * PREP_CHUNK;
    const long long end = beg + PREP_CHUNK < J.total ? beg + PREP_CHUNK : J.total;
    // The layout whose k runs along the source's SLOW axis (forward of a Conv2D, dgrad of a Conv2DTranspose: w_ns == 1) is a transpose:
    // read straight, consecutive threads fetched 4 bytes of 32 different sectors (0.13 ms per network and step).  A chunk is whole rows of K
    // outputs, i.e. R = chunk / K consecutive n per k in the source: read those runs with n fastest into shared memory, write k fastest.
    if (s.w_ns == 1 && s.w_ks != 1 && s.K <= 1024 && PREP_CHUNK % s.K == 0 && beg % s.K == 0) {
        const int K = s.K, len = (int)(end - beg), R = len / K;                 // len is a multiple of K (total = taps * Nn * K)
        const long long row0 = beg / K;
        for (int j = threadIdx.x; j < len; j += 256) {
            const int kk = j / R, rr = j - kk * R;
            const long long row = row0 + rr;
            const long long tap = row / s.Nn; const int n = (int)(row - tap * s.Nn);
            tile[rr * (K + 1) + kk] = (kk < s.k_real && n < s.n_real) ? __ldg(J.w + tap * J.tap_elems + (long long)kk * s.w_ks + n) : 0.f;
        }
        __syncthreads();
        for (int j = threadIdx.x; j < len; j += 256) {
            const int rr = j / K, kk = j - rr * K;
            s.o[beg + j] = __float2bfloat16_rn(tile[rr * (K + 1) + kk]);
        }
        return;
    }
    for (long long i = beg + threadIdx.x; i < end; i += 256) {
        const int k = (int)(i % s.K);
        const long long t2 = i / s.K;
        const int n = (int)(t2 % s.Nn);
        const long long tap = t2 / s.Nn;
        const float v = (k < s.k_real && n < s.n_real) ? __ldg(J.w + tap * J.tap_elems + (long long)k * s.w_ks + (long long)n * s.w_ns) : 0.f;
        s.o[i] = __float2bfloat16_rn(v);
    }
}

bool wgrad_halo_ok(const shm_conv_desc* d) {
    return !d->transposed && d->stride == 1 && d->kh == 3 && d->kw == 3 && d->H % 16 == 0 && d->W % 8 == 0 &&
           d->Cin % 64 == 0 && d->Cout % 64 == 0;
}

template <int MODE>
int launch_wgrad_halo_t(const CUtensorMap& tmX, const CUtensorMap& tmDY, const WhParams& p, cudaStream_t st) {
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(wgrad_halo_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, WhCfg<MODE>::SMEM); attr = true; }
    wgrad_halo_kernel<MODE><<<p.units * p.splits, TC_THREADS, WhCfg<MODE>::SMEM, st>>>(tmX, tmDY, p);
    SHM_CHECK_LAUNCH("wgrad_halo_kernel");
    return SHM_OK;
}

int launch_wgrad_halo(const shm_conv_desc* d, const void* x, const void* dy, float* dw, cudaStream_t st) {
    WhParams p{};
    const int mode = (d->Cin % 128 == 0 && d->Cout % 128 == 0) ? 1 : 0;
    p.cblocks = d->Cin / (mode ? 128 : 64);
    p.nblocks = d->Cout / (mode ? 128 : 64);
    p.units = p.cblocks * p.nblocks * (mode ? 3 : 1);
    p.tiles_x = d->W / 8; p.tiles_y = d->H / 16;
    p.total_tiles = d->N * p.tiles_x * p.tiles_y;
    p.Cin = d->Cin; p.Cout = d->Cout; p.dW = dw;
    // split the pixel range so that the grid fills whole waves of SMs; the epilogue (TMEM -> red.add) costs about 6 tiles
    const int sms = shm_num_sms();
    long long best_cost = -1; int best_s = 1;
    const int smax = p.total_tiles < 4 * sms ? p.total_tiles : 4 * sms;
    for (int s = 1; s <= smax; ++s) {
        const long long ctas = (long long)p.units * s;
        if (ctas > 4LL * sms && s > 1) break;
        const long long waves = (ctas + sms - 1) / sms;
        const long long cost = waves * (cdiv(p.total_tiles, s) + 6);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_s = s; }
    }
    p.tiles_per_split = cdiv(p.total_tiles, best_s);
    p.splits = cdiv(p.total_tiles, p.tiles_per_split);
    CUtensorMap tmX, tmDY;
    // ldx < Cin: the tensor holds only ldx channels per pixel (a first layer's input padded to 16 instead of 64); the 64-channel TMA box
    // then runs past the channel extent and the out-of-bounds part is zero-filled in shared memory -- the padding costs no HBM bytes
    if (int rc = encode_act_box(&tmX, x, d->ldx < d->Cin ? d->ldx : d->Cin, d->W, d->H, d->N, d->ldx, HALO_W, mode ? 16 : 18)) return rc;
    if (int rc = encode_act_box(&tmDY, dy, d->Cout, d->W, d->H, d->N, d->ldy, 8, 16)) return rc;
    if (mode == 1 && pair_enabled() && p.cblocks % 2 == 0) {
        // CTA pairs: a unit = TWO 128-channel ci blocks; the pixel range is split as before, over half as many (twice as wide) units
        p.units /= 2;
        long long bc = -1; int bs = 1;
        for (int s = 1; s <= smax; ++s) {
            const long long ctas = 2LL * p.units * s;
            if (ctas > 4LL * sms && s > 1) break;
            const long long waves = (ctas + sms - 1) / sms;
            const long long cost = waves * (cdiv(p.total_tiles, s) + 6);
            if (bc < 0 || cost < bc) { bc = cost; bs = s; }
        }
        p.tiles_per_split = cdiv(p.total_tiles, bs);
        p.splits = cdiv(p.total_tiles, p.tiles_per_split);
        static bool attr = false;
        if (!attr) { cudaFuncSetAttribute(wgrad_halo2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WH2_SMEM); attr = true; }
        wgrad_halo2_kernel<<<2 * p.units * p.splits, TC_THREADS, WH2_SMEM, st>>>(tmX, tmDY, p);
        SHM_CHECK_LAUNCH("wgrad_halo2_kernel");
        return SHM_OK;
    }
    return mode ? launch_wgrad_halo_t<1>(tmX, tmDY, p, st) : launch_wgrad_halo_t<0>(tmX, tmDY, p, st);
}

// stride-2 3x3 weight gradients: Conv2DTranspose (G = dy, S = x) and strided Conv2D (G = x, S = dy) at even sizes
bool wgrad_s2_ok(const shm_conv_desc* d) {
    if (d->kh != 3 || d->kw != 3 || d->stride != 2 || d->Cin % 64 != 0 || d->Cout % 64 != 0) return false;
    int Ho, Wo; out_dims(d, Ho, Wo);
    const int Hs = d->transposed ? d->H : Ho, Ws = d->transposed ? d->W : Wo;       // the small image's lattice
    const int Hg = d->transposed ? Ho : d->H, Wg = d->transposed ? Wo : d->W;
    return Hg == 2 * Hs && Wg == 2 * Ws && Hs % 16 == 0 && Ws % 8 == 0;
}

template <int MODE>
int launch_wgrad_s2_t(const CUtensorMap& tmG, const CUtensorMap& tmS, const Ws2Params& p, cudaStream_t st) {
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(wgrad_s2_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, Ws2Cfg<MODE>::SMEM); attr = true; }
    wgrad_s2_kernel<MODE><<<p.units * p.splits, TC_THREADS, Ws2Cfg<MODE>::SMEM, st>>>(tmG, tmS, p);
    SHM_CHECK_LAUNCH("wgrad_s2_kernel");
    return SHM_OK;
}

int launch_wgrad_s2(const shm_conv_desc* d, const void* x, const void* dy, float* dw, cudaStream_t st) {
    int Ho, Wo; out_dims(d, Ho, Wo);
    const void* G = d->transposed ? dy : x;  const void* S = d->transposed ? x : dy;
    const int CG = d->transposed ? d->Cout : d->Cin, CS = d->transposed ? d->Cin : d->Cout;
    const int ldG = d->transposed ? d->ldy : d->ldx, ldS = d->transposed ? d->ldx : d->ldy;
    const int Hs = d->transposed ? d->H : Ho, Ws = d->transposed ? d->W : Wo;
    Ws2Params p{};
    const int mode = (CG % 128 == 0 && CS % 128 == 0) ? 1 : 0;
    p.cblocks = CG / (mode ? 128 : 64);
    p.nblocks = CS / (mode ? 128 : 64);
    p.units = p.cblocks * p.nblocks * (mode ? 3 : 1);
    p.tiles_x = Ws / 8; p.tiles_y = Hs / 16;
    p.total_tiles = d->N * p.tiles_x * p.tiles_y;
    p.CG = CG; p.CS = CS; p.dW = dw;
    const int sms = shm_num_sms();
    long long best_cost = -1; int best_s = 1;
    const int smax = p.total_tiles < 4 * sms ? p.total_tiles : 4 * sms;
    for (int s = 1; s <= smax; ++s) {
        const long long ctas = (long long)p.units * s;
        if (ctas > 4LL * sms && s > 1) break;
        const long long waves = (ctas + sms - 1) / sms;
        const long long cost = waves * (cdiv(p.total_tiles, s) + 6);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_s = s; }
    }
    p.tiles_per_split = cdiv(p.total_tiles, best_s);
    p.splits = cdiv(p.total_tiles, p.tiles_per_split);
    CUtensorMap tmG, tmS;
    // parity-plane boxes: 10 x 18 (MODE 0) or 10 x 16 (MODE 1) plane pixels at traversal stride 2
    if (int rc = encode_act(&tmG, G, CG, 2 * Ws, 2 * Hs, d->N, ldG, HALO_W, mode ? 16 : HALO_H, 1, 2)) return rc;
    if (int rc = encode_act_box(&tmS, S, CS, Ws, Hs, d->N, ldS, 8, 16)) return rc;
    return mode ? launch_wgrad_s2_t<1>(tmG, tmS, p, st) : launch_wgrad_s2_t<0>(tmG, tmS, p, st);
}

int tc_check(const shm_conv_desc* d) {
    SHM_REQUIRE(d != nullptr, "conv desc is NULL");
    if (d->dtype != SHM_BF16) SHM_FAIL(SHM_EUNSUPPORTED, "conv_tc: needs dtype bf16");
    if (d->Cin % 64 != 0 || d->Cout % 64 != 0) {
        // thin layers: the forward pass (K = Cin, N = Cout) or the dgrad (K = Cout, N = Cin) of a stride-1 3x3 conv goes through the
        // halo kernel with narrow pixel rows / few accumulator columns
        const bool thin = !d->transposed && (thin_ok(d->H, d->W, d->Cin, d->Cout, d->kh, d->kw, d->stride) ||
                                             thin_ok(d->H, d->W, d->Cout, d->Cin, d->kh, d->kw, d->stride));
        if (!thin) SHM_FAIL(SHM_EUNSUPPORTED, "conv_tc: Cin=%d / Cout=%d must be multiples of 64 (or a thin 3x3 stride-1 layer)", d->Cin, d->Cout);
    }
    if (d->ldx % 8 != 0 || d->ldy % 8 != 0) SHM_FAIL(SHM_EUNSUPPORTED, "conv_tc: ld must be a multiple of 8");
    if (d->kh < 1 || d->kh > 3 || d->kw < 1 || d->kw > 3 || (d->stride != 1 && d->stride != 2)) SHM_FAIL(SHM_EUNSUPPORTED, "conv_tc: kernel/stride unsupported");
    if (d->transposed && d->stride != 2) SHM_FAIL(SHM_EUNSUPPORTED, "conv_tc: transposed conv needs stride 2");
    return SHM_OK;
}

}  // namespace

extern "C" int64_t shm_conv2d_tc_weight_elems(const shm_conv_desc* d) {
    if (!d) return 0;
    return (int64_t)d->kh * d->kw * d->Cin * d->Cout;
}

extern "C" int shm_conv2d_tc_prep_weights(const shm_conv_desc* d, const float* w, int cin_real, void* w_tc, int for_dgrad, void* stream) {
    if (int rc = tc_check(d)) return rc;
    SHM_REQUIRE(w && w_tc, "shm_conv2d_tc_prep_weights: NULL buffer");
    if (cin_real <= 0) cin_real = d->Cin;
    SHM_REQUIRE(cin_real <= d->Cin, "shm_conv2d_tc_prep_weights: cin_real > Cin");
    SHM_REQUIRE(cin_real == d->Cin || !d->transposed, "shm_conv2d_tc_prep_weights: channel padding is for Conv2D only");
    // GEMM reduction dim K / output dim Nn and the strides of (k, n) inside one Keras tap slice (cin_real x Cout elements)
    int K, Nn, w_ks, w_ns, k_real, n_real;
    if (!d->transposed) {          // (kh,kw,Cin,Cout)
        if (!for_dgrad) { K = d->Cin; Nn = d->Cout; w_ks = d->Cout; w_ns = 1; k_real = cin_real; n_real = d->Cout; }
        else            { K = d->Cout; Nn = d->Cin; w_ks = 1; w_ns = d->Cout; k_real = d->Cout; n_real = cin_real; }
    } else {                       // (kh,kw,Cout,Cin)
        if (!for_dgrad) { K = d->Cin; Nn = d->Cout; w_ks = 1; w_ns = d->Cin; }
        else            { K = d->Cout; Nn = d->Cin; w_ks = d->Cin; w_ns = 1; }
        k_real = K; n_real = Nn;
    }
    const long long total = (long long)d->kh * d->kw * d->Cin * d->Cout;
    long long g = cdiv64(total, 256);
    if (g > shm_num_sms() * 16) g = shm_num_sms() * 16;
    prep_w_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(w, (bf16*)w_tc, K, Nn, k_real, n_real, (long long)cin_real * d->Cout, w_ks, w_ns, K, K, total);
    SHM_CHECK_LAUNCH("prep_w_kernel");
    return SHM_OK;
}

extern "C" int shm_conv2d_tc_prep_weights_both(const shm_conv_desc* d, const float* w, int cin_real, void* w_tc_fwd, void* w_tc_dgrad, void* stream) {
    if (int rc = tc_check(d)) return rc;
    SHM_REQUIRE(w && w_tc_fwd && w_tc_dgrad, "shm_conv2d_tc_prep_weights_both: NULL buffer");
    if (cin_real <= 0) cin_real = d->Cin;
    SHM_REQUIRE(cin_real <= d->Cin, "shm_conv2d_tc_prep_weights_both: cin_real > Cin");
    SHM_REQUIRE(cin_real == d->Cin || !d->transposed, "shm_conv2d_tc_prep_weights_both: channel padding is for Conv2D only");
    PrepSet a, b;
    a.o = (bf16*)w_tc_fwd; b.o = (bf16*)w_tc_dgrad;
    if (!d->transposed) {          // (kh,kw,Cin,Cout)
        a.K = d->Cin; a.Nn = d->Cout; a.w_ks = d->Cout; a.w_ns = 1; a.k_real = cin_real; a.n_real = d->Cout;
        b.K = d->Cout; b.Nn = d->Cin; b.w_ks = 1; b.w_ns = d->Cout; b.k_real = d->Cout; b.n_real = cin_real;
    } else {                       // (kh,kw,Cout,Cin)
        a.K = d->Cin; a.Nn = d->Cout; a.w_ks = 1; a.w_ns = d->Cin; a.k_real = a.K; a.n_real = a.Nn;
        b.K = d->Cout; b.Nn = d->Cin; b.w_ks = d->Cin; b.w_ns = 1; b.k_real = b.K; b.n_real = b.Nn;
    }
    const long long total = (long long)d->kh * d->kw * d->Cin * d->Cout;
    long long g = cdiv64(total, 256);
    if (g > shm_num_sms() * 8) g = shm_num_sms() * 8;
    prep_w_both_kernel<<<dim3((unsigned)g, 2), 256, 0, (cudaStream_t)stream>>>(w, a, b, (long long)cin_real * d->Cout, total);
    SHM_CHECK_LAUNCH("prep_w_both_kernel");
    return SHM_OK;
}

// ---- batched form of shm_conv2d_tc_prep_weights_both: fill one job record per layer (host memory, shm_conv2d_tc_prep_job_bytes()
//      each), finalize the array (block prefix), copy it to the device once, then ONE launch per optimiser step refreshes every layer.
extern "C" int shm_conv2d_tc_prep_job_bytes(void) { return (int)sizeof(PrepJob); }
extern "C" int shm_conv2d_tc_prep_job(const shm_conv_desc* d, const float* w, int cin_real, void* w_tc_fwd, void* w_tc_dgrad, void* job_out) {
    if (int rc = tc_check(d)) return rc;
    SHM_REQUIRE(w && w_tc_fwd && w_tc_dgrad && job_out, "shm_conv2d_tc_prep_job: NULL buffer");
    if (cin_real <= 0) cin_real = d->Cin;
    SHM_REQUIRE(cin_real <= d->Cin, "shm_conv2d_tc_prep_job: cin_real > Cin");
    SHM_REQUIRE(cin_real == d->Cin || !d->transposed, "shm_conv2d_tc_prep_job: channel padding is for Conv2D only");
    PrepJob J{};
    J.w = w;
    PrepSet& a = J.a; PrepSet& b = J.b;
    a.o = (bf16*)w_tc_fwd; b.o = (bf16*)w_tc_dgrad;
    if (!d->transposed) {          // (kh,kw,Cin,Cout)
        a.K = d->Cin; a.Nn = d->Cout; a.w_ks = d->Cout; a.w_ns = 1; a.k_real = cin_real; a.n_real = d->Cout;
        b.K = d->Cout; b.Nn = d->Cin; b.w_ks = 1; b.w_ns = d->Cout; b.k_real = d->Cout; b.n_real = cin_real;
    } else {                       // (kh,kw,Cout,Cin)
        a.K = d->Cin; a.Nn = d->Cout; a.w_ks = 1; a.w_ns = d->Cin; a.k_real = a.K; a.n_real = a.Nn;
        b.K = d->Cout; b.Nn = d->Cin; b.w_ks = d->Cin; b.w_ns = 1; b.k_real = b.K; b.n_real = b.Nn;
    }
    J.tap_elems = (long long)cin_real * d->Cout;
    J.total = (long long)d->kh * d->kw * d->Cin * d->Cout;
    J.nblocks = (int)cdiv64(J.total, PREP_CHUNK);
    memcpy(job_out, &J, sizeof(J));
    return SHM_OK;
}
extern "C" int shm_conv2d_tc_prep_jobs_finalize(void* jobs_host, int njobs) {
    if (!jobs_host || njobs <= 0) return 0;
    PrepJob* J = reinterpret_cast<PrepJob*>(jobs_host);
    int acc = 0;
    for (int i = 0; i < njobs; ++i) { J[i].block_begin = acc; acc += J[i].nblocks; }
    return acc;                                       // total blocks of the launch
}
extern "C" int shm_conv2d_tc_prep_multi(const void* jobs_dev, int njobs, int total_blocks, void* stream) {
    SHM_REQUIRE(jobs_dev && njobs > 0 && total_blocks > 0, "shm_conv2d_tc_prep_multi: bad args");
    prep_w_multi_kernel<<<dim3((unsigned)total_blocks, 2), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const PrepJob*>(jobs_dev), njobs);
    SHM_CHECK_LAUNCH("prep_w_multi_kernel");
    return SHM_OK;
}

// forward-layout weights of a layer whose device geometry (d->Cin, d->Cout) is a zero-padded version of the Keras kernel
// (cin_real, cout_real): input channels in segments of seg_pad device channels holding seg_real real ones each.
extern "C" int shm_conv2d_tc_prep_weights_padded(const shm_conv_desc* d, const float* w, int cin_real, int seg_real, int seg_pad,
                                                 int cout_real, void* w_tc, void* stream) {
    if (int rc = tc_check(d)) return rc;
    SHM_REQUIRE(w && w_tc, "shm_conv2d_tc_prep_weights_padded: NULL buffer");
    SHM_REQUIRE(cin_real > 0 && cout_real > 0 && cout_real <= d->Cout && seg_real > 0 && seg_pad >= seg_real && d->Cin % seg_pad == 0 &&
                (d->Cin / seg_pad) * seg_real >= cin_real, "shm_conv2d_tc_prep_weights_padded: inconsistent padding geometry");
    const int K = d->Cin, Nn = d->Cout;
    int w_ks, w_ns;
    if (!d->transposed) { w_ks = cout_real; w_ns = 1; }      // (kh,kw,cin,cout)
    else { w_ks = 1; w_ns = cin_real; }                      // (kh,kw,cout,cin)
    const long long total = (long long)d->kh * d->kw * K * Nn;
    long long g = cdiv64(total, 256);
    if (g > shm_num_sms() * 16) g = shm_num_sms() * 16;
    prep_w_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(w, (bf16*)w_tc, K, Nn, cin_real, cout_real, (long long)cin_real * cout_real, w_ks, w_ns,
                                                            seg_real, seg_pad, total);
    SHM_CHECK_LAUNCH("prep_w_kernel");
    return SHM_OK;
}

extern "C" int shm_conv2d_tc_supported(const shm_conv_desc* d, int for_dgrad) {
    if (tc_check(d) != SHM_OK) return 0;
    if (d->Cin % 64 != 0 || d->Cout % 64 != 0)                                // thin layer (tc_check admitted one orientation)
        return for_dgrad ? thin_ok(d->H, d->W, d->Cout, d->Cin, d->kh, d->kw, d->stride) : thin_ok(d->H, d->W, d->Cin, d->Cout, d->kh, d->kw, d->stride);
    int Ho, Wo; out_dims(d, Ho, Wo);
    int BW, BH, BI;
    // in every form the lattice is the SMALL image: Ho x Wo of a strided conv, H x W of a transposed conv
    const int Qh = d->transposed ? d->H : Ho, Qw = d->transposed ? d->W : Wo;
    if (d->stride == 2 && ((d->transposed ? Ho : d->H) % 2 != 0 || (d->transposed ? Wo : d->W) % 2 != 0)) return 0;
    return pick_box(d->N, Qh, Qw, BW, BH, BI) ? 1 : 0;
}

// which kernel serves a layer (for per-kernel accounting in bench.py): pass 0 = fwd, 1 = dgrad, 2 = wgrad.
// 0 conv_tc (generic)  1 conv_halo  2 conv_multi/big  3 conv_multi/scatter  4 wgrad_tc (generic)  5 wgrad_halo<0>  6 wgrad_halo<1>
extern "C" int shm_conv2d_tc_route(const shm_conv_desc* d, int pass) {
    if (tc_check(d) != SHM_OK) return -1;
    int Ho, Wo; out_dims(d, Ho, Wo);
    const int s = d->stride;
    if (pass == 2) {
        if (wgrad_halo_ok(d)) return (d->Cin % 128 == 0 && d->Cout % 128 == 0) ? 6 : 5;
        if (wgrad_s2_ok(d)) return (d->Cin % 128 == 0 && d->Cout % 128 == 0) ? 8 : 7;
        return 4;
    }
    if (pass == 0) {
        if (!d->transposed) {
            if (halo_ok(d->H, d->W, d->Cin, d->Cout, d->kh, d->kw, s) || thin_ok(d->H, d->W, d->Cin, d->Cout, d->kh, d->kw, s)) return 1;
            if (big_ok(d->H, d->W, d->Cin, d->Cout, d->kh, d->kw, s)) return 2;
            return 0;
        }
        return (s == 2 && scatter_ok(d->H, d->W, d->Cin, d->Cout)) ? 3 : 0;
    }
    if (d->transposed) return 0;
    if (s == 1) {
        if (halo_ok(d->H, d->W, d->Cout, d->Cin, d->kh, d->kw, s) || thin_ok(d->H, d->W, d->Cout, d->Cin, d->kh, d->kw, s)) return 1;
        if (big_ok(d->H, d->W, d->Cout, d->Cin, d->kh, d->kw, s)) return 2;
        return 0;
    }
    return (s == 2 && d->H % 2 == 0 && d->W % 2 == 0 && scatter_ok(Ho, Wo, d->Cout, d->Cin)) ? 3 : 0;
}

// forward: Conv2D (gather form) or Conv2DTranspose (scatter-by-parity form, 4 launches)
static int tc_fwd_impl(const shm_conv_desc* d, const void* x, const void* w_tc, const float* bias, void* y, int nstore, void* stream, double* stats = nullptr);
extern "C" int shm_conv2d_tc_fwd(const shm_conv_desc* d, const void* x, const void* w_tc, const float* bias, void* y, void* stream) {
    return tc_fwd_impl(d, x, w_tc, bias, y, 0, stream);
}
// forward + instance-norm statistics of the stored output in the same kernel: stats[n][c] = (sum, sum of squares) over the pixels of image n,
// ADDED to the caller's (zeroed) fp64 buffer -- what shm_inorm_stats would compute from y, without reading y back
extern "C" int shm_conv2d_tc_fwd_stats(const shm_conv_desc* d, const void* x, const void* w_tc, const float* bias, void* y, double* stats, void* stream) {
    SHM_REQUIRE(stats != nullptr, "shm_conv2d_tc_fwd_stats: stats is NULL");
    if (!shm_conv2d_tc_stats_supported(d)) SHM_FAIL(SHM_EUNSUPPORTED, "shm_conv2d_tc_fwd_stats: no statistics epilogue for this layer (ask shm_conv2d_tc_stats_supported)");
    return tc_fwd_impl(d, x, w_tc, bias, y, 0, stream, stats);
}
extern "C" int shm_conv2d_tc_stats_supported(const shm_conv_desc* d) {
    if (tc_check(d) != SHM_OK || d->transposed) return 0;
    if (!shm_conv2d_tc_supported(d, 0)) return 0;
    const int s = d->stride;
    if (halo_ok(d->H, d->W, d->Cin, d->Cout, d->kh, d->kw, s) || thin_ok(d->H, d->W, d->Cin, d->Cout, d->kh, d->kw, s))
        return halo_stats_ok(d->H, d->Cin, d->Cout) ? 1 : 0;
    return 0;                                          // big-halo / generic kernels: the separate pass is cheaper (see EpiStats)
}
extern "C" int shm_conv2d_tc_fwd_cols(const shm_conv_desc* d, const void* x, const void* w_tc, const float* bias, void* y, int nstore, void* stream) {
    SHM_REQUIRE(d && nstore > 0 && nstore <= d->Cout && nstore % 8 == 0, "shm_conv2d_tc_fwd_cols: nstore must be a multiple of 8 in (0, Cout]");
    return tc_fwd_impl(d, x, w_tc, bias, y, nstore, stream);
}
static int tc_fwd_impl(const shm_conv_desc* d, const void* x, const void* w_tc, const float* bias, void* y, int nstore, void* stream, double* stats) {
    if (int rc = tc_check(d)) return rc;
    SHM_REQUIRE(x && w_tc && y, "shm_conv2d_tc_fwd: NULL buffer");
    if (nstore > 0 && nstore < d->Cout && !(d->transposed && d->stride == 2 && scatter_ok(d->H, d->W, d->Cin, d->Cout)))
        SHM_FAIL(SHM_EUNSUPPORTED, "shm_conv2d_tc_fwd_cols: partial column stores are served by the Conv2DTranspose scatter kernel only");
    SHM_REQUIRE((reinterpret_cast<uintptr_t>(bias) & 15) == 0, "shm_conv2d_tc_fwd: bias must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    int Ho, Wo; out_dims(d, Ho, Wo);
    const int s = d->stride;
    const int wrows = d->kh * d->kw * d->Cout;
    int tdy[9], tdx[9], twr[9];
    if (!d->transposed) {
        const int pby = same_pad_before(d->H, d->kh, s), pbx = same_pad_before(d->W, d->kw, s);
        int nt = 0;
        for (int ky = 0; ky < d->kh; ++ky)
            for (int kx = 0; kx < d->kw; ++kx) { tdy[nt] = ky - pby; tdx[nt] = kx - pbx; twr[nt] = (ky * d->kw + kx) * d->Cout; ++nt; }
        if (halo_ok(d->H, d->W, d->Cin, d->Cout, d->kh, d->kw, s) || thin_ok(d->H, d->W, d->Cin, d->Cout, d->kh, d->kw, s))
            return launch_halo(d->N, d->H, d->W, d->Cin, d->Cout, x, d->ldx, w_tc, wrows, bias, d->act, y, d->ldy, tdy, tdx, twr, st, stats);
        if (d->Cin % 64 != 0 || d->Cout % 64 != 0) SHM_FAIL(SHM_EUNSUPPORTED, "conv_tc fwd: thin layer %d -> %d not servable", d->Cin, d->Cout);
        if (big_ok(d->H, d->W, d->Cin, d->Cout, d->kh, d->kw, s))
            return launch_big(d->N, d->H, d->W, d->Cin, d->Cout, x, d->ldx, w_tc, wrows, bias, d->act, y, d->ldy, tdy, tdx, twr, st);
        Geometry g{d->H, d->W, d->ldx, d->Cin, Ho, Wo, s, Ho, Wo, d->ldy, d->Cout, 1};
        return launch_tc(g, d->N, x, w_tc, wrows, bias, d->act, y, tdy, tdx, twr, nt, 0, 0, st);
    }
    if (stats != nullptr) SHM_FAIL(SHM_EUNSUPPORTED, "shm_conv2d_tc_fwd_stats: transposed convolutions have no statistics epilogue");
    // transposed: out[p] = sum_{o,k: s*o + k - pb = p} x[o] W[k];  p = s*q + r
    const int pby = same_pad_before(Ho, d->kh, s), pbx = same_pad_before(Wo, d->kw, s);
    if (s == 2 && scatter_ok(d->H, d->W, d->Cin, d->Cout)) {
        ScatterTap taps[9]; int n = 0; bool fits = true;
        for (int ry = 0; ry < 2; ++ry) for (int rx = 0; rx < 2; ++rx)
            for (int ky = 0; ky < d->kh; ++ky) {
                if (((ry + pby - ky) % 2) != 0) continue;
                for (int kx = 0; kx < d->kw; ++kx) {
                    if (((rx + pbx - kx) % 2) != 0) continue;
                    const int oy = (ry + pby - ky) / 2, ox = (rx + pbx - kx) / 2;
                    if (oy < -1 || oy > 0 || ox < -1 || ox > 0) fits = false;
                    taps[n++] = ScatterTap{ry, rx, oy, ox, (ky * d->kw + kx) * d->Cout};
                }
            }
        if (fits) return launch_scatter(d->N, d->H, d->W, d->Cin, d->Cout, x, d->ldx, w_tc, wrows, bias, d->act, y, Ho, Wo, d->ldy, taps, n, st, nstore);
    }
    if (nstore > 0 && nstore < d->Cout) SHM_FAIL(SHM_EUNSUPPORTED, "shm_conv2d_tc_fwd_cols: this transposed conv does not fit the scatter kernel");
    for (int ry = 0; ry < s; ++ry)
        for (int rx = 0; rx < s; ++rx) {
            int nt = 0;
            for (int ky = 0; ky < d->kh; ++ky) {
                if (((ry + pby - ky) % s) != 0) continue;
                for (int kx = 0; kx < d->kw; ++kx) {
                    if (((rx + pbx - kx) % s) != 0) continue;
                    tdy[nt] = (ry + pby - ky) / s; tdx[nt] = (rx + pbx - kx) / s; twr[nt] = (ky * d->kw + kx) * d->Cout; ++nt;
                }
            }
            if (nt == 0) SHM_FAIL(SHM_EUNSUPPORTED, "conv_tc: parity class without taps (k < stride)");
            Geometry g{d->H, d->W, d->ldx, d->Cin, d->H, d->W, 1, Ho, Wo, d->ldy, d->Cout, s};
            if (int rc = launch_tc(g, d->N, x, w_tc, wrows, bias, d->act, y, tdy, tdx, twr, nt, ry, rx, st)) return rc;
        }
    return SHM_OK;
}

// dgrad: dx = dL/dx from dy = dL/d(pre-activation); w_tc prepared with for_dgrad = 1
extern "C" int shm_conv2d_tc_dgrad(const shm_conv_desc* d, const void* dy, const void* w_tc, void* dx, void* stream) {
    if (int rc = tc_check(d)) return rc;
    SHM_REQUIRE(dy && w_tc && dx, "shm_conv2d_tc_dgrad: NULL buffer");
    if ((d->Cin % 64 != 0 || d->Cout % 64 != 0) && !thin_ok(d->H, d->W, d->Cout, d->Cin, d->kh, d->kw, d->stride))
        SHM_FAIL(SHM_EUNSUPPORTED, "conv_tc dgrad: thin layer %d -> %d not servable (needs K = Cout in {16, 32, 64}, N = Cin in {16, 32, 64, 128})", d->Cin, d->Cout);
    cudaStream_t st = (cudaStream_t)stream;
    int Ho, Wo; out_dims(d, Ho, Wo);
    const int s = d->stride;
    const int wrows = d->kh * d->kw * d->Cin;
    int tdy[9], tdx[9], twr[9];
    if (d->transposed) {
        // dx[o] = sum_k dy[s*o + k - pb] W[k]^T : gather at stride s from the big image
        const int pby = same_pad_before(Ho, d->kh, s), pbx = same_pad_before(Wo, d->kw, s);
        int nt = 0;
        for (int ky = 0; ky < d->kh; ++ky)
            for (int kx = 0; kx < d->kw; ++kx) { tdy[nt] = ky - pby; tdx[nt] = kx - pbx; twr[nt] = (ky * d->kw + kx) * d->Cin; ++nt; }
        Geometry g{Ho, Wo, d->ldy, d->Cout, d->H, d->W, s, d->H, d->W, d->ldx, d->Cin, 1};
        return launch_tc(g, d->N, dy, w_tc, wrows, nullptr, SHM_ACT_NONE, dx, tdy, tdx, twr, nt, 0, 0, st);
    }
    // Conv2D dgrad: dx[p] = sum_{o,k: s*o + k - pb = p} dy[o] W[k]^T  (scatter by parity; one class when s == 1)
    const int pby = same_pad_before(d->H, d->kh, s), pbx = same_pad_before(d->W, d->kw, s);
    if (s == 2 && d->H % 2 == 0 && d->W % 2 == 0 && scatter_ok(Ho, Wo, d->Cout, d->Cin)) {
        ScatterTap taps[9]; int n = 0; bool fits = true;
        for (int ry = 0; ry < 2; ++ry) for (int rx = 0; rx < 2; ++rx)
            for (int ky = 0; ky < d->kh; ++ky) {
                if (((ry + pby - ky) % 2) != 0) continue;
                for (int kx = 0; kx < d->kw; ++kx) {
                    if (((rx + pbx - kx) % 2) != 0) continue;
                    const int oy = (ry + pby - ky) / 2, ox = (rx + pbx - kx) / 2;
                    if (oy < -1 || oy > 0 || ox < -1 || ox > 0) fits = false;
                    taps[n++] = ScatterTap{ry, rx, oy, ox, (ky * d->kw + kx) * d->Cin};
                }
            }
        if (fits && n <= 9) return launch_scatter(d->N, Ho, Wo, d->Cout, d->Cin, dy, d->ldy, w_tc, wrows, nullptr, SHM_ACT_NONE, dx, d->H, d->W, d->ldx, taps, n, st);
    }
    for (int ry = 0; ry < s; ++ry)
        for (int rx = 0; rx < s; ++rx) {
            int nt = 0;
            for (int ky = 0; ky < d->kh; ++ky) {
                if (((ry + pby - ky) % s) != 0) continue;
                for (int kx = 0; kx < d->kw; ++kx) {
                    if (((rx + pbx - kx) % s) != 0) continue;
                    tdy[nt] = (ry + pby - ky) / s; tdx[nt] = (rx + pbx - kx) / s; twr[nt] = (ky * d->kw + kx) * d->Cin; ++nt;
                }
            }
            if (nt == 0) SHM_FAIL(SHM_EUNSUPPORTED, "conv_tc: parity class without taps");
            if (halo_ok(d->H, d->W, d->Cout, d->Cin, d->kh, d->kw, s) || thin_ok(d->H, d->W, d->Cout, d->Cin, d->kh, d->kw, s))
                return launch_halo(d->N, d->H, d->W, d->Cout, d->Cin, dy, d->ldy, w_tc, wrows, nullptr, SHM_ACT_NONE, dx, d->ldx, tdy, tdx, twr, st);
            if (big_ok(d->H, d->W, d->Cout, d->Cin, d->kh, d->kw, s))
                return launch_big(d->N, d->H, d->W, d->Cout, d->Cin, dy, d->ldy, w_tc, wrows, nullptr, SHM_ACT_NONE, dx, d->ldx, tdy, tdx, twr, st);
            Geometry g{Ho, Wo, d->ldy, d->Cout, cdiv(d->H - ry, s), cdiv(d->W - rx, s), 1, d->H, d->W, d->ldx, d->Cin, s};
            if (int rc = launch_tc(g, d->N, dy, w_tc, wrows, nullptr, SHM_ACT_NONE, dx, tdy, tdx, twr, nt, ry, rx, st)) return rc;
        }
    return SHM_OK;
}

// wgrad: dw += dL/dw in the Keras layout (fp32 atomics).  x and dy are bf16.
extern "C" int shm_conv2d_tc_wgrad(const shm_conv_desc* d, const void* x, const void* dy, float* dw, void* stream) {
    if (int rc = tc_check(d)) return rc;
    SHM_REQUIRE(x && dy && dw, "shm_conv2d_tc_wgrad: NULL buffer");
    if (d->Cin % 64 != 0 || d->Cout % 64 != 0) SHM_FAIL(SHM_EUNSUPPORTED, "conv_tc wgrad: thin layers (%d -> %d) are forward-only", d->Cin, d->Cout);
    cudaStream_t st = (cudaStream_t)stream;
    int Ho, Wo; out_dims(d, Ho, Wo);
    const int s = d->stride;
    WgParams p{};
    const void *A, *B;
    int Ca, Cb, HA, WA, lda, HB, WB, ldb;
    p.ntaps = 0;
    if (!d->transposed) {
        // dW[tap][ci][co] = sum_q x[q*s + k - pb, ci] dy[q, co]
        A = x; Ca = d->Cin; HA = d->H; WA = d->W; lda = d->ldx; p.SA = s;
        B = dy; Cb = d->Cout; HB = Ho; WB = Wo; ldb = d->ldy; p.SB = 1;
        p.Qh = Ho; p.Qw = Wo;
        const int pby = same_pad_before(d->H, d->kh, s), pbx = same_pad_before(d->W, d->kw, s);
        for (int ky = 0; ky < d->kh; ++ky)
            for (int kx = 0; kx < d->kw; ++kx) {
                const int t = p.ntaps++;
                p.day[t] = ky - pby; p.dax[t] = kx - pbx; p.dby[t] = 0; p.dbx[t] = 0;
                p.woff[t] = (long long)(ky * d->kw + kx) * d->Cin * d->Cout;
            }
    } else {
        // dW[tap][co][ci] = sum_o dy[o*s + k - pb, co] x[o, ci]
        A = dy; Ca = d->Cout; HA = Ho; WA = Wo; lda = d->ldy; p.SA = s;
        B = x; Cb = d->Cin; HB = d->H; WB = d->W; ldb = d->ldx; p.SB = 1;
        p.Qh = d->H; p.Qw = d->W;
        const int pby = same_pad_before(Ho, d->kh, s), pbx = same_pad_before(Wo, d->kw, s);
        for (int ky = 0; ky < d->kh; ++ky)
            for (int kx = 0; kx < d->kw; ++kx) {
                const int t = p.ntaps++;
                p.day[t] = ky - pby; p.dax[t] = kx - pbx; p.dby[t] = 0; p.dbx[t] = 0;
                p.woff[t] = (long long)(ky * d->kw + kx) * d->Cin * d->Cout;
            }
    }
    if (wgrad_halo_ok(d)) return launch_wgrad_halo(d, x, dy, dw, st);
    if (d->ldx < d->Cin) SHM_FAIL(SHM_EUNSUPPORTED, "conv_tc wgrad: a channel-truncated input (ldx=%d < Cin=%d) is served by the halo wgrad kernel only", d->ldx, d->Cin);
    if (wgrad_s2_ok(d)) return launch_wgrad_s2(d, x, dy, dw, st);
    if (!pick_box64(d->N, p.Qh, p.Qw, p.BW, p.BH, p.BI)) SHM_FAIL(SHM_EUNSUPPORTED, "wgrad_tc: lattice %dx%d does not tile into 64-point boxes", p.Qh, p.Qw);
    p.tiles_x = p.Qw / p.BW; p.tiles_y = p.Qh / p.BH;
    p.cblocks = Ca / 64;
    p.nblocks = p.ntaps * p.cblocks;
    p.m_tiles = (p.nblocks + 1) / 2;
    const int BN = (Cb % 128 == 0) ? 128 : 64;
    p.n_tiles = Cb / BN;
    p.kblocks = (int)((long long)d->N * p.Qh * p.Qw / 64);
    p.ldw = Cb; p.dW = dw;
    // split K so that the grid covers the SMs ~2x, keeping >= 8 K-blocks per CTA
    const int tiles = p.m_tiles * p.n_tiles;
    int splits = cdiv(shm_num_sms() * 2, tiles);
    if (splits > cdiv(p.kblocks, 8)) splits = cdiv(p.kblocks, 8);
    if (splits < 1) splits = 1;
    p.kb_per_split = cdiv(p.kblocks, splits);
    splits = cdiv(p.kblocks, p.kb_per_split);
    CUtensorMap tmA, tmB;
    if (int rc = encode_act(&tmA, A, Ca, WA, HA, d->N, lda, p.BW, p.BH, p.BI, p.SA)) return rc;
    if (int rc = encode_act(&tmB, B, Cb, WB, HB, d->N, ldb, p.BW, p.BH, p.BI, p.SB)) return rc;
    const int grid = tiles * splits;
    if (BN == 128) {
        static bool attr = false;
        if (!attr) { cudaFuncSetAttribute(wgrad_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, WgCfg<128>::SMEM); attr = true; }
        wgrad_tc_kernel<128><<<grid, TC_THREADS, WgCfg<128>::SMEM, st>>>(tmA, tmB, p);
    } else {
        static bool attr = false;
        if (!attr) { cudaFuncSetAttribute(wgrad_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, WgCfg<64>::SMEM); attr = true; }
        wgrad_tc_kernel<64><<<grid, TC_THREADS, WgCfg<64>::SMEM, st>>>(tmA, tmB, p);
    }
    SHM_CHECK_LAUNCH("wgrad_tc_kernel");
    return SHM_OK;
}
