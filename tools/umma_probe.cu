// umma_probe.cu -- hardware probe: how does tcgen05.mma address a K-major SWIZZLE_128B A operand whose start address is NOT
// 1024-byte aligned and whose 8-row groups are SBO apart with SBO != 1024?  (Needed to reuse one halo tile for all 9 taps.)
// B = 64x64 identity, so D[m][n] = A_window[m][n]: the accumulator shows which smem row / column each MMA row read.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o umma_probe umma_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>
typedef __nv_bfloat16 bf16;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0, tries = 0;
    while (!ok) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (++tries > (1u << 22)) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\ntcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

struct Probe { int row_off, sbo, base_off; };

__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, Probe pr, float* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;                // 256 rows x 128 B
    uint8_t* sB = smem + 32768;        // 64 rows x 128 B
    uint64_t* bars = (uint64_t*)(smem + 32768 + 8192);
    uint32_t* slot = (uint32_t*)(bars + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
    if (warp == 0) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(64) : "memory"); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bars[0], 32768 + 8192);
        tma_load_2d(sA, &tmA, &bars[0], 0, 0);
        tma_load_2d(sB, &tmB, &bars[0], 0, 0);
        mbar_wait(&bars[0], 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        auto mk = [&](uint32_t saddr, uint32_t sbo, uint32_t bo) {
            uint64_t d = 0;
            d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
            d |= (uint64_t)1 << 16;
            d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
            d |= (uint64_t)1 << 46;
            d |= (uint64_t)(bo & 7) << 49;
            d |= (uint64_t)2 << 61;
            return d;
        };
        const uint64_t ad = mk(smem_u32(sA) + pr.row_off * 128, pr.sbo, pr.base_off);
        const uint64_t bd = mk(smem_u32(sB), 1024, 0);
        for (int k = 0; k < 4; ++k) umma_bf16(tmem, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), make_idesc(128, 64), k != 0);
        umma_commit(&bars[1]);
    }
    mbar_wait(&bars[1], 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c * 32, r);
        for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + c * 32 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64) : "memory");
}

typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                          CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    EncFn enc = (EncFn)fp;
    const int R = 256;
    std::vector<bf16> hA(R * 64), hB(64 * 64);
    bf16 *dA, *dB; float* dO;
    cudaMalloc(&dA, R * 64 * 2); cudaMalloc(&dB, 64 * 64 * 2); cudaMalloc(&dO, 128 * 64 * 4);
    for (int n = 0; n < 64; ++n) for (int k = 0; k < 64; ++k) hB[n * 64 + k] = __float2bfloat16(n == k ? 1.f : 0.f);
    cudaMemcpy(dB, hB.data(), 64 * 64 * 2, cudaMemcpyHostToDevice);
    auto mkmap = [&](CUtensorMap* tm, void* base, int rows) {
        cuuint64_t dims[2] = {64, (cuuint64_t)rows}; cuuint64_t strides[1] = {128}; cuuint32_t box[2] = {64, (cuuint32_t)rows}; cuuint32_t es[2] = {1, 1};
        CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) printf("encode failed %d\n", (int)r);
    };
    CUtensorMap tmA, tmB; mkmap(&tmA, dA, R); mkmap(&tmB, dB, 64);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 + 8192 + 1024 + 256);
    Probe cfgs[] = {{0, 1024, 0}, {1, 1024, 0}, {1, 1024, 1}, {3, 1024, 0}, {3, 1024, 3}, {0, 1280, 0}, {11, 1280, 0}, {11, 1280, 3}, {21, 1280, 5}, {21, 1280, 0}, {8, 1280, 0}, {16, 2304, 0}};
    std::vector<float> o1(128 * 64), o2(128 * 64);
    for (auto pr : cfgs) {
        for (int pass = 0; pass < 2; ++pass) {
            for (int r = 0; r < R; ++r) for (int k = 0; k < 64; ++k) hA[r * 64 + k] = __float2bfloat16(pass == 0 ? (float)r : (float)k);
            cudaMemcpy(dA, hA.data(), R * 64 * 2, cudaMemcpyHostToDevice);
            probe_kernel<<<1, 128, 32768 + 8192 + 1024 + 256>>>(tmA, tmB, pr, dO);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("cfg row_off=%d sbo=%d bo=%d: CUDA error %s\n", pr.row_off, pr.sbo, pr.base_off, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(pass == 0 ? o1.data() : o2.data(), dO, 128 * 64 * 4, cudaMemcpyDeviceToHost);
        }
        // expected under the "swizzle = function of the final absolute address" model
        int ok_rows = 0, ok_cols = 0, uniform = 0;
        for (int m = 0; m < 128; ++m) {
            const int exp_row = pr.row_off + (m / 8) * (pr.sbo / 128) + (m % 8);
            bool u = true, okr = true, okc = true;
            for (int n = 0; n < 64; ++n) { if (o1[m * 64 + n] != o1[m * 64]) u = false; if ((int)o1[m * 64 + n] != exp_row) okr = false; if ((int)o2[m * 64 + n] != n) okc = false; }
            uniform += u; ok_rows += okr; ok_cols += okc;
        }
        printf("row_off=%2d sbo=%4d base_off=%d : rows as expected %3d/128, columns identity %3d/128, row-uniform %3d/128 |", pr.row_off, pr.sbo, pr.base_off, ok_rows, ok_cols, uniform);
        for (int m = 0; m < 12; ++m) printf(" m%d:r%d,c%d..%d", m, (int)o1[m * 64], (int)o2[m * 64], (int)o2[m * 64 + 8]);
        printf("\n");
    }
    return 0;
}
