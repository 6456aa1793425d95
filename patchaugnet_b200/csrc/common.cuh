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

// Two reference distances at once with Blackwell's packed fp32 pair instructions (sub/mul/fma .f32x2): the same IEEE
// operations in the same order on each half, so both results are bit-identical to ref_sqdist — at half the issue slots
// of the fp32 pipe.  a* hold two points (lo, hi halves of a 64-bit register pair), b* the same point in both halves.
__device__ __forceinline__ unsigned long long pack_f32x2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void ref_sqdist_x2(unsigned long long ax, unsigned long long ay, unsigned long long az, unsigned long long bx,
                                              unsigned long long by, unsigned long long bz, float &d_lo, float &d_hi) {
    asm("{\n"
        ".reg .b64 dx, dy, dz, t;\n"
        "sub.rn.f32x2 dx, %2, %5;\n"
        "sub.rn.f32x2 dy, %3, %6;\n"
        "sub.rn.f32x2 dz, %4, %7;\n"
        "mul.rn.f32x2 t, dy, dy;\n"
        "fma.rn.f32x2 t, dx, dx, t;\n"
        "fma.rn.f32x2 t, dz, dz, t;\n"
        "mov.b64 {%0, %1}, t;\n"
        "}\n" : "=f"(d_lo), "=f"(d_hi) : "l"(ax), "l"(ay), "l"(az), "l"(bx), "l"(by), "l"(bz));
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }
