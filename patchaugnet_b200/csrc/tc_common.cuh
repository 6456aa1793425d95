// tc_common.cuh — tcgen05 / TMEM / TMA / mbarrier PTX wrappers and the bf16 hi/lo operand helpers shared by the
// tensor-core kernels (mlp_tc.cu, vlad_tc.cu).  sm_100a only.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include "common.cuh"

namespace tc {

constexpr int TM = 128;                       // rows per tile (UMMA M)
constexpr int KCH = 64;                       // bf16 per K chunk = one 128-byte swizzle row
constexpr int A_CHUNK = TM * 128;             // bytes of one K chunk of one operand plane

// ---- PTX wrappers ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// multicast variant: the box lands at the same CTA-relative offset in every CTA of `mask`, and each destination CTA's
// barrier (same offset) receives the complete_tx
__device__ __forceinline__ void tma_load_2d_mc(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar, uint16_t mask) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5}], [%2], %3;"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "h"(mask), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4 in [0,14),
// LBO (unused for swizzled K-major) = 1 in [16,30), SBO = 1024 B (8 rows x 128 B) >> 4 in [32,46), version 1 in [46,48),
// layout type SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D = f32 [4,6)=1, A = B = bf16 [7,10)=[10,13)=1,
// both K-major, N>>3 in [17,23), M>>4 in [24,29).
__device__ __forceinline__ uint32_t umma_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}
// ---- warp-uniform issue path -----------------------------------------------------------------------------------
// The MMA warp runs its loops with all 32 lanes converged (warp index broadcast with a shuffle so the compiler treats it
// as uniform) and predicates only the tcgen05 instructions on an elected lane: the descriptors then live in uniform
// registers and each MMA costs a handful of uniform-datapath instructions.  Issuing from inside an `if (lane == 0)` block
// instead makes the compiler wrap EVERY tcgen05.mma in an elect/branch "uniformisation" loop (~150 clk per MMA, measured:
// the tensor pipe then starves behind its own issue thread).
__device__ __forceinline__ int uniform_warp_idx() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }
__device__ __forceinline__ uint32_t elect_one() {           // 1 in exactly one lane of a converged warp
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(pred));
    return pred;
}
constexpr uint32_t UMMA_DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);      // SBO 1024 B, version 1, SWIZZLE_128B
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t saddr) { return ((saddr & 0x3FFFF) >> 4) | (1u << 16); }
// descriptor low words advance by (bytes >> 4); shared memory is < 256 KB so the 14-bit address field never carries
__device__ __forceinline__ void umma_f16_if(uint32_t issue, uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                            uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p, q;\n"
        ".reg .b64 da, db;\n"
        "setp.ne.b32 p, %6, 0;\n"
        "setp.ne.b32 q, %7, 0;\n"
        "mov.b64 da, {%1, %2};\n"
        "mov.b64 db, {%3, %4};\n"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n"
        "}\n" ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate), "r"(issue) : "memory");
}
__device__ __forceinline__ void umma_commit_if(uint32_t issue, uint64_t *bar) {
    asm volatile(
        "{\n"
        ".reg .pred q;\n"
        "setp.ne.b32 q, %1, 0;\n"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(issue) : "memory");
}
// commit that arrives on the barrier at this offset in every CTA of `mask` (a weight stage is shared by the cluster)
__device__ __forceinline__ void umma_commit_mc_if(uint32_t issue, uint64_t *bar, uint16_t mask) {
    asm volatile(
        "{\n"
        ".reg .pred q;\n"
        "setp.ne.b32 q, %1, 0;\n"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %2;\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(issue), "h"(mask) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// x = hi + lo with hi = bf16_rn(x), lo = bf16_rn(x - hi); packs element pairs (low half = lower index)
__device__ __forceinline__ void split_pack(float a, float b, uint32_t &hi, uint32_t &lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    const float2 hf = __bfloat1622float2(h);
    const __nv_bfloat162 l = __floats2bfloat162_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<const uint32_t *>(&h);
    lo = *reinterpret_cast<const uint32_t *>(&l);
}

// byte offset of the 16-byte unit holding channels [8*j, 8*j+8) of row r inside the operand planes
__device__ __forceinline__ uint32_t a_unit_offset(int r, int j) {
    return (uint32_t)((j >> 3) * A_CHUNK + r * 128 + (((j & 7) ^ (r & 7)) << 4));
}
// a2 == nullptr: single-plane (plain bf16) operands, only the rounded values are stored
__device__ __forceinline__ void store_units(uint8_t *a1, uint8_t *a2, int r, int j, const float (&v)[8]) {
    uint4 h, l;
    split_pack(v[0], v[1], h.x, l.x);
    split_pack(v[2], v[3], h.y, l.y);
    split_pack(v[4], v[5], h.z, l.z);
    split_pack(v[6], v[7], h.w, l.w);
    const uint32_t off = a_unit_offset(r, j);
    *reinterpret_cast<uint4 *>(a1 + off) = h;
    if (a2) *reinterpret_cast<uint4 *>(a2 + off) = l;
}


}  // namespace tc
