// Microbenchmark: cycles per tcgen05.mma (kind::f16, bf16, M=128, K=16, cta_group::1) for several N / accumulator
// patterns, operands in SWIZZLE_128B K-major shared memory.  One CTA per SM, 128 threads; warp 0 issues.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../patchaugnet_b200/csrc -I../../include mma_bench.cu -o mma_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace tc;

// mode 0: one accumulator, every MMA depends on the previous one
// mode 1: two accumulators alternating
// mode 2: four accumulators round robin (N <= 128)
__device__ __forceinline__ void umma_ts(uint32_t issue, uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
    asm volatile(
        "{\n.reg .pred q;\n.reg .b64 db;\nsetp.ne.b32 q, %5, 0;\nmov.b64 db, {%2, %3};\n"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, 1;\n}\n" ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(issue) : "memory");
}
// mode 3: A operand from tensor memory (TS form), one accumulator; mode 4: TS with a commit + barrier wait every 8 MMAs
__global__ void __launch_bounds__(128, 1) k(int N, int mode, int iters, int distinct_ab, long long *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = uniform_warp_idx();
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, slot, 0);
    if (warp == 0) {
        const uint32_t leader = elect_one();
        const uint32_t a_lo = umma_desc_lo(smem_u32(smem)), b_lo = umma_desc_lo(smem_u32(smem + 64 * 1024));
        const uint32_t idesc = umma_idesc(N);
        const int nacc = mode == 0 ? 1 : (mode == 1 ? 2 : 4);
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int j = 0; j < 12; ++j) {
                const uint32_t d = tmem + (uint32_t)((j % nacc) * N);
                const uint32_t ao = distinct_ab ? (uint32_t)(j % 4) * 2 + (uint32_t)((j / 4) & 1) * (A_CHUNK >> 4) : 0;
                const uint32_t bo = distinct_ab ? (uint32_t)(j % 4) * 2 + (uint32_t)((j / 8) & 1) * ((uint32_t)N * 8) : 0;
                if (mode >= 3) umma_ts(leader, tmem, tmem + 256 + (uint32_t)(j % 4) * 8 + (uint32_t)((j / 4) & 1) * 128, b_lo + bo, UMMA_DESC_HI, idesc);
                else umma_f16_if(leader, d, a_lo + ao, UMMA_DESC_HI, b_lo + bo, UMMA_DESC_HI, idesc, 1);
            }
        }
        umma_commit_if(leader, &bar);
        long long t1 = clock64();
        mbar_wait(&bar, 0);
        long long t2 = clock64();
        if (leader && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

int main() {
    long long *out;
    cudaMallocManaged(&out, 16);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 200;
    for (int grid : {1, 148})
        for (int N : {64, 128, 256})
            for (int mode : {0, 1, 2, 3})
                for (int dab : {0, 1}) {
                    if (mode == 2 && N > 128) continue;
                    if (mode == 1 && N > 256) continue;
                    if (mode == 3 && (N > 256 || grid == 1)) continue;
                    k<<<grid, 128, 200 * 1024>>>(N, mode, iters, dab, out);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                    printf("grid %3d N %3d acc-mode %d distinct %d: issue %.1f clk/MMA, complete %.1f clk/MMA\n", grid, N, mode, dab,
                           (double)out[0] / (12.0 * iters), (double)out[1] / (12.0 * iters));
                }
    return 0;
}
