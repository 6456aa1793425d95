// common.cuh — shared helpers for libpatchaug_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/patchaug_b200.h"

#define PAB_API extern "C" __attribute__((visibility("default")))

extern int g_pab_launches;  // counted in api.cu

#define PAB_LAUNCH_CHECK()                                  \
    do {                                                    \
        ++g_pab_launches;                                   \
        cudaError_t e__ = cudaGetLastError();               \
        if (e__ != cudaSuccess) return (int)e__;            \
    } while (0)

#define PAB_CUDA(x)                                         \
    do {                                                    \
        cudaError_t e__ = (x);                              \
        if (e__ != cudaSuccess) return (int)e__;            \
    } while (0)

static inline int pab_divup(long a, long b) { return (int)((a + b - 1) / b); }

// The reference's squared distance as nvcc 12.9 -O2 contracts it for sm_100:
//   (ax-bx)^2 + (ay-by)^2 + (az-bz)^2  ->  fma(dz,dz, fma(dx,dx, dy*dy)).
// Intrinsics are used so the order can never be re-associated or re-contracted by this build.
__device__ __forceinline__ float ref_sqdist(float ax, float ay, float az, float bx, float by, float bz) {
    const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }
