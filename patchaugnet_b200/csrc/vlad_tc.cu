// vlad_tc.cu — NetVLAD soft-assignment + residual aggregation on tcgen05 tensor cores (sm_100a).
//
// Same maths and partial-sum contract as vlad.cu (NetVLADBase.forward, patch_aug_net/models/loupe.py:191-222), with both
// contractions as tcgen05.mma on bf16 hi/lo operand planes (three MMAs per product, fp32 accumulation in TMEM):
//
//   GEMM1  logits[128 pts][K]      = x[128][256] . Wc^T          A = x planes (K-major), B = Wc planes (K-major)
//   softmax over the K clusters, one thread per point straight out of TMEM (tcgen05.ld), written as act^T planes
//   GEMM2  vlad^T[256 ch][K]      += x^T[256][128 pts] . act      A = THE SAME x planes read MN-major (no transpose copy:
//                                                                 a 128-byte swizzled row of 64 channels is a K-major
//                                                                 row of GEMM1 and an MN-major row of GEMM2),
//                                                                 B = act^T planes (K-major); accumulated in TMEM over
//                                                                 all tiles of the CTA's work item
// A work item is (cloud, 256- or 512-row chunk); persistent CTAs loop over items.  Each item writes a partial (K, C) block and
// partial a_sum (K) exactly like vlad_partial_kernel, so vlad_finalize_kernel (vlad.cu) is shared.
#include <math.h>
#include "tc_common.cuh"

extern int g_tc_max_ctas;   // mlp_tc.cu
extern long long *g_tc_trace;   // mlp_tc.cu: optional timeline buffer (pab_tune_tc_trace), here [tile < 16][8 events] of CTA 0

namespace {

using namespace tc;

constexpr int VT_WORK = 256;                  // worker threads (8 warps); warp 0: idle helper, warp 1: MMA issuer
constexpr int VT_THREADS = 64 + VT_WORK;
constexpr int VT_ROWS_MIN = 256;               // smallest work item (rows of one cloud); the workspace is sized for it
#ifndef VT_ROWS_BIG
#define VT_ROWS_BIG 512                        // work item of large clouds (n >= 2048)
#endif
#define VT_TRACE(ev)                                                                             \
    do {                                                                                         \
        if (a.trace && blockIdx.x == 0 && tcount < 16 && wt == 0) a.trace[tcount * 8 + (ev)] = clock64(); \
    } while (0)
constexpr int VT_C = 256;                     // channels (4 chunks of 64)

struct VtArgs {
    int n, K, Kp, nchunk, nitems, rows_per_item;
    int planes;                               // 2: bf16 hi/lo operands (three MMAs per product); 1: plain bf16 (one MMA)
    const float *x, *shift;
    const __nv_bfloat16 *wc_hi, *wc_lo;       // (Kp, 256) K-major, bn1 scale folded, rows >= K zero
    float *part, *asum;
    long long *trace;
};

// MN-major SWIZZLE_128B descriptor over the x planes: 64-channel blocks (one K chunk each) are LBO = A_CHUNK bytes apart,
// groups of 8 points (K direction) are SBO = 1024 bytes apart.
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(A_CHUNK >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
           (2ull << 61);
}
__device__ __forceinline__ uint32_t umma_desc_mn_lo(uint32_t saddr) { return ((saddr & 0x3FFFF) >> 4) | ((uint32_t)(A_CHUNK >> 4) << 16); }
constexpr uint32_t UMMA_DESC_MN_HI = UMMA_DESC_HI;            // same SBO / version / swizzle fields as the K-major descriptor
__device__ __forceinline__ uint32_t umma_idesc_amn(int n) {   // as umma_idesc, with A MN-major (bit 15)
    return umma_idesc(n) | (1u << 15);
}

__global__ void __launch_bounds__(VT_THREADS, 1) vlad_tc_kernel(const VtArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *x1 = smem, *x2 = x1 + 4 * A_CHUNK;                    // x hi / lo planes: 4 chunks x [128][64] bf16
    uint8_t *w1 = x2 + 4 * A_CHUNK;                                // Wc hi: 4 chunks x [Kp][64]
    const int wchunk = a.Kp * 128;
    uint8_t *w2 = w1 + 4 * wchunk;
    uint8_t *p1 = w2 + 4 * wchunk;                                 // act hi: [128 pts][64 clusters] bf16, 128-byte swizzled rows —
    uint8_t *p2 = p1 + A_CHUNK;                                    // GEMM2's B operand read MN-major (a point's clusters are contiguous)
    uint8_t *misc = p2 + A_CHUNK;
    uint64_t *a_ready = reinterpret_cast<uint64_t *>(misc);        // 64-channel chunk 0 of the x planes staged (256 arrivals)
    uint64_t *a_ready_c = reinterpret_cast<uint64_t *>(misc + 40) - 1;   // chunks 1..3 at a_ready_c[1..3]: GEMM1 starts on chunk 0
                                                                   // while the later chunks are still being loaded
    uint64_t *g2_half = reinterpret_cast<uint64_t *>(misc + 64 + 256 + 4 * TM * 4);   // GEMM2 block 0 done: chunks 0,1 are free
    uint64_t *d1_ready = a_ready + 1;                              // logits in TMEM             (tcgen05.commit)
    uint64_t *p_ready = a_ready + 2;                               // act^T planes staged        (128 arrivals)
    uint64_t *g2_done = a_ready + 3;                               // GEMM2 of the tile finished (tcgen05.commit)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(a_ready + 4);
    volatile int *progress = reinterpret_cast<volatile int *>(misc + 36);   // tile the workers are staging (paces the prefetching helper)
    float *shift_s = reinterpret_cast<float *>(misc + 64);         // [64] bn1 shift of the clusters (zero beyond K)
    float *xmax = reinterpret_cast<float *>(misc + 64 + 256);      // [2][128] row maxima of the two column halves; after the softmax
    float *psum = xmax;                                            //   [4 lane quarters][64] per-warp column sums of act (a_sum)
    float *xsum = xmax + 2 * TM;                                   // [2][128] row sums
    const bool split = a.Kp > 32;

    const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;
    if (tid == 0) {
        mbar_init(a_ready, VT_WORK); mbar_init(d1_ready, 1);
        for (int c = 1; c < 4; ++c) mbar_init(a_ready_c + c, VT_WORK);
        mbar_init(g2_half, 1);
        *progress = -1; mbar_init(p_ready, a.Kp > 32 ? VT_WORK : 128); mbar_init(g2_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // cluster weights into the K-major swizzled planes (once per CTA)
    for (int u = tid; u < a.Kp * 32; u += VT_THREADS) {            // 16-byte units: Kp rows x 32 units (256 ch)
        const int k = u >> 5, j = u & 31;
        const uint32_t off = (uint32_t)((j >> 3) * wchunk + k * 128 + (((j & 7) ^ (k & 7)) << 4));
        *reinterpret_cast<uint4 *>(w1 + off) = __ldg(reinterpret_cast<const uint4 *>(a.wc_hi + (size_t)k * VT_C) + j);
        if (a.planes == 2) *reinterpret_cast<uint4 *>(w2 + off) = __ldg(reinterpret_cast<const uint4 *>(a.wc_lo + (size_t)k * VT_C) + j);
    }
    if (tid < 64) shift_s[tid] = tid < a.K ? __ldg(a.shift + tid) : 0.f;
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    const uint32_t d1 = tmem, d2 = tmem + 64;                      // D2 block cb at d2 + cb*64

    if (warp == 1) {
        // all lanes run the loops, one elected lane issues (tc_common.cuh: warp-uniform issue path)
        const uint32_t leader = elect_one();
        const bool two = a.planes == 2;
        uint32_t tcount = 0;
        const uint32_t id1 = umma_idesc(a.Kp), id2 = umma_idesc_amn(a.Kp) | (1u << 16);   // GEMM2: A and B MN-major
        const uint32_t x1_lo = umma_desc_lo(smem_u32(x1)), x2_lo = umma_desc_lo(smem_u32(x2));
        const uint32_t w1_lo = umma_desc_lo(smem_u32(w1)), w2_lo = umma_desc_lo(smem_u32(w2));
        const uint32_t p1_lo = umma_desc_mn_lo(smem_u32(p1)), p2_lo = umma_desc_mn_lo(smem_u32(p2));
        const uint32_t xm1_lo = umma_desc_mn_lo(smem_u32(x1)), xm2_lo = umma_desc_mn_lo(smem_u32(x2));
        const uint32_t wstep = (uint32_t)wchunk >> 4;
        for (int item = blockIdx.x; item < a.nitems; item += gridDim.x) {
            const int chunk = item % a.nchunk;
            const int r_begin = chunk * a.rows_per_item, r_end = min(a.n, r_begin + a.rows_per_item);
            int t_in_item = 0;
            for (int r0 = r_begin; r0 < r_end; r0 += TM, ++tcount, ++t_in_item) {
#pragma unroll
                for (int kc = 0; kc < 4; ++kc) {
                    mbar_wait(kc == 0 ? a_ready : a_ready_c + kc, tcount & 1);
                    tc_fence_after();
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint32_t xa = (uint32_t)kc * (A_CHUNK >> 4) + 2 * ks, wb = (uint32_t)kc * wstep + 2 * ks;
                        umma_f16_if(leader, d1, x1_lo + xa, UMMA_DESC_HI, w1_lo + wb, UMMA_DESC_HI, id1, (kc | ks) != 0);
                        if (two) {
                            umma_f16_if(leader, d1, x2_lo + xa, UMMA_DESC_HI, w1_lo + wb, UMMA_DESC_HI, id1, 1);
                            umma_f16_if(leader, d1, x1_lo + xa, UMMA_DESC_HI, w2_lo + wb, UMMA_DESC_HI, id1, 1);
                        }
                    }
                }
                umma_commit_if(leader, d1_ready);
                mbar_wait(p_ready, tcount & 1);
                tc_fence_after();
#pragma unroll
                for (int cb = 0; cb < 2; ++cb) {
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {            // 16 points per MMA
                        const uint32_t xa = (uint32_t)(2 * cb) * (A_CHUNK >> 4) + (uint32_t)ks * (2048 >> 4);
                        const uint32_t pb = (uint32_t)ks * (2048 >> 4);   // 16 points = two 8-row swizzle atoms
                        umma_f16_if(leader, d2 + cb * 64, xm1_lo + xa, UMMA_DESC_MN_HI, p1_lo + pb, UMMA_DESC_MN_HI, id2, (t_in_item | ks) != 0);
                        if (two) {
                            umma_f16_if(leader, d2 + cb * 64, xm2_lo + xa, UMMA_DESC_MN_HI, p1_lo + pb, UMMA_DESC_MN_HI, id2, 1);
                            umma_f16_if(leader, d2 + cb * 64, xm1_lo + xa, UMMA_DESC_MN_HI, p2_lo + pb, UMMA_DESC_MN_HI, id2, 1);
                        }
                    }
                    if (cb == 0) umma_commit_if(leader, g2_half);      // channels 0..127 = chunks 0,1 of the x planes are free
                }
                umma_commit_if(leader, g2_done);
            }
        }
        __syncwarp();
    } else if (warp == 0) {
        // helper: while tile t is processed, pull tile t+1 of this CTA's sequence (128 contiguous KB of one cloud) into L2 with
        // one bulk prefetch, so the workers' loads — the longest phase of the per-tile chain — find it there instead of in HBM
        if (lane == 0) {
            auto prefetch_tile = [&](int item, int r0) {
                const int cloud = item / a.nchunk, chunk = item % a.nchunk;
                const int r_end = min(a.n, chunk * a.rows_per_item + a.rows_per_item);
                const float *src = a.x + ((size_t)cloud * a.n + r0) * VT_C;
                const uint32_t bytes = (uint32_t)(min(TM, r_end - r0)) * VT_C * 4;
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
            };
            uint32_t tcount = 0;
            for (int item = blockIdx.x; item < a.nitems; item += gridDim.x) {
                const int chunk = item % a.nchunk;
                const int r_begin = chunk * a.rows_per_item, r_end = min(a.n, r_begin + a.rows_per_item);
                for (int r0 = r_begin; r0 < r_end; r0 += TM, ++tcount) {
                    // the tile after (item, r0) in this CTA's sequence
                    int n_item = item, n_r0 = r0 + TM;
                    if (n_r0 >= r_end) { n_item = item + gridDim.x; n_r0 = (n_item % a.nchunk) * a.rows_per_item; }
                    // pace: one tile ahead of the workers.  A monotonic counter, not an mbarrier phase: nothing waits for this
                    // warp, so it may fall a whole tile behind — and a parity wait that is two phases late never wakes
                    while (*progress < (int)tcount) {}
                    if (n_item < a.nitems) prefetch_tile(n_item, n_r0);
                }
            }
        }
        __syncwarp();
    } else if (warp >= 2) {
        const int wt = tid - 64, wwarp = warp - 2;
        const int q = warp & 3, half = wwarp >> 2;
        const int row = q * 32 + lane;
        const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
        uint32_t tcount = 0;
        for (int item = blockIdx.x; item < a.nitems; item += gridDim.x) {
            const int cloud = item / a.nchunk, chunk = item % a.nchunk;
            const int r_begin = chunk * a.rows_per_item, r_end = min(a.n, r_begin + a.rows_per_item);
            const float *xg = a.x + (size_t)cloud * a.n * VT_C;
            float asum_k = 0.f;                                    // cluster (wt - 128)'s running sum (threads of half 1)
            for (int r0 = r_begin; r0 < r_end; r0 += TM, ++tcount) {
                VT_TRACE(0);
                // The tile is staged one 64-channel chunk at a time (a warp covers 16 rows of a chunk: 4 steps of 4 rows x 256
                // contiguous bytes), two chunks in flight: chunk c+2 is requested before chunk c is split and stored, and GEMM1
                // starts on chunk c as soon as all warps have staged it.  Chunks 0,1 may be overwritten once GEMM2's first block
                // (channels 0..127) of the previous tile has completed, chunks 2,3 after its second block.
                const int lr = lane >> 3, lu = lane & 7;                     // row within a 4-row step, 32-byte unit within 256 B
                float4 q0[2][4], q1[2][4];
                auto request = [&](int c, float4 (&a0)[4], float4 (&a1)[4]) {
#pragma unroll
                    for (int st4 = 0; st4 < 4; ++st4) {
                        const int r = wwarp * 16 + st4 * 4 + lr;
                        a0[st4] = a1[st4] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (r0 + r < r_end) {
                            const float4 *src = reinterpret_cast<const float4 *>(xg + (size_t)(r0 + r) * VT_C + c * 64) + 2 * lu;
                            a0[st4] = __ldg(src); a1[st4] = __ldg(src + 1);
                        }
                    }
                };
                auto stage = [&](int c, const float4 (&a0)[4], const float4 (&a1)[4]) {
#pragma unroll
                    for (int st4 = 0; st4 < 4; ++st4) {
                        const float v[8] = {a0[st4].x, a0[st4].y, a0[st4].z, a0[st4].w, a1[st4].x, a1[st4].y, a1[st4].z, a1[st4].w};
                        store_units(x1, a.planes == 2 ? x2 : nullptr, wwarp * 16 + st4 * 4 + lr, c * 8 + lu, v);
                    }
                    fence_proxy_async();
                    mbar_arrive(c == 0 ? a_ready : a_ready_c + c);
                };
                request(0, q0[0], q1[0]);
                request(1, q0[1], q1[1]);
                if (tcount > 0) mbar_wait(g2_half, (tcount - 1) & 1);
                if (wt == 0) *progress = (int)tcount;
                VT_TRACE(1);
                stage(0, q0[0], q1[0]);
                request(2, q0[0], q1[0]);
                stage(1, q0[1], q1[1]);
                request(3, q0[1], q1[1]);
                if (tcount > 0) mbar_wait(g2_done, (tcount - 1) & 1);      // previous tile's GEMM2 no longer reads chunks 2,3
                stage(2, q0[0], q1[0]);
                stage(3, q0[1], q1[1]);
                VT_TRACE(2);
                if (half == 0 || split) {
                    // softmax over the clusters of this point: with more than 32 clusters the two warps that share a TMEM lane
                    // quarter take 32 columns each and exchange the row maximum and the row sum through shared memory
                    mbar_wait(d1_ready, tcount & 1);
                    tc_fence_after();
                    VT_TRACE(3);
                    float lg[32];
                    tmem_ld32(trow + (uint32_t)(half * 32), lg);
                    const int kb = half * 32;
                    const bool valid = r0 + row < r_end;
                    float mx = -INFINITY;
#pragma unroll
                    for (int k = 0; k < 32; ++k)
                        if (kb + k < a.K) { lg[k] += shift_s[kb + k]; mx = fmaxf(mx, lg[k]); }
                    if (split) {
                        xmax[half * TM + row] = mx;
                        asm volatile("bar.sync 2, %0;" ::"n"(VT_WORK) : "memory");
                        mx = fmaxf(mx, xmax[(half ^ 1) * TM + row]);
                    }
                    float sum = 0.f;
#pragma unroll
                    for (int k = 0; k < 32; ++k)
                        if (kb + k < a.K) { lg[k] = __expf(lg[k] - mx); sum += lg[k]; }
                    if (split) {
                        xsum[half * TM + row] = sum;
                        asm volatile("bar.sync 2, %0;" ::"n"(VT_WORK) : "memory");
                        sum = xsum[row] + xsum[TM + row];                    // same order in both threads of the row
                    }
                    const float inv = valid ? 1.f / sum : 0.f;
                    // this point's probabilities as bf16 hi/lo, eight clusters per 16-byte unit of its 128-byte row; lg[] keeps
                    // the rounded value hi + lo (exact in fp32): a_sum must add up what the tensor cores add up
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (kb + 8 * u < a.Kp) {
                            uint4 h, l;
                            uint32_t *hw = &h.x, *lw = &l.x;
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const int k0 = 8 * u + 2 * i;
                                const float pa = kb + k0 < a.K ? lg[k0] * inv : 0.f, pb = kb + k0 + 1 < a.K ? lg[k0 + 1] * inv : 0.f;
                                const __nv_bfloat162 hh = __floats2bfloat162_rn(pa, pb);
                                const float2 hf = __bfloat1622float2(hh);
                                const __nv_bfloat162 ll = __floats2bfloat162_rn(pa - hf.x, pb - hf.y);
                                const float2 lf = __bfloat1622float2(ll);
                                hw[i] = *reinterpret_cast<const uint32_t *>(&hh);
                                lw[i] = *reinterpret_cast<const uint32_t *>(&ll);
                                if (a.planes == 2) { lg[k0] = hf.x + lf.x; lg[k0 + 1] = hf.y + lf.y; }
                                else { lg[k0] = hf.x; lg[k0 + 1] = hf.y; }
                            }
                            const uint32_t off = (uint32_t)(row * 128 + ((((kb >> 3) + u) ^ (row & 7)) << 4));
                            *reinterpret_cast<uint4 *>(p1 + off) = h;
                            if (a.planes == 2) *reinterpret_cast<uint4 *>(p2 + off) = l;
                        } else {
#pragma unroll
                            for (int i = 0; i < 8; ++i) lg[8 * u + i] = 0.f;
                        }
                    }
                    tc_fence_before();
                    fence_proxy_async();
                    VT_TRACE(4);
                    mbar_arrive(p_ready);
                    // column sums of act over this warp's 32 points: transposing butterfly (31 shuffles), lane k ends up with
                    // cluster kb + k; fixed tree, so the result depends on the tile only
#pragma unroll
                    for (int off = 16, n = 32; off >= 1; off >>= 1, n >>= 1) {
                        const bool up = (lane & off) != 0;
#pragma unroll
                        for (int i = 0; i < n / 2; ++i) {
                            const float send = up ? lg[i] : lg[i + n / 2];
                            const float keep = up ? lg[i + n / 2] : lg[i];
                            lg[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                        }
                    }
                    psum[q * 64 + kb + lane] = lg[0];
                }
                asm volatile("bar.sync 3, %0;" ::"n"(VT_WORK) : "memory");       // every warp's column sums are in psum
                if (half == 1) {
                    // a_sum[k] += the four lane quarters' column sums, in a fixed order
                    const int k = wt - 128;
                    if (k < a.K) asum_k += (psum[k] + psum[64 + k]) + (psum[128 + k] + psum[192 + k]);
                }
            }
            // ---- item done: accumulators -> partial (K, C) block, a_sum -> partial (K) ------------------------------
            mbar_wait(g2_done, (tcount - 1) & 1);
            tc_fence_after();
            {
                const int ch = half * 128 + row;                               // D2 block `half`, TMEM lane = channel
                float *part = a.part + (size_t)item * a.K * VT_C;
                float v[32];
                tmem_ld32(trow + 64 + (uint32_t)half * 64, v);
#pragma unroll
                for (int k = 0; k < 32; ++k)
                    if (k < a.K) part[(size_t)k * VT_C + ch] = v[k];
                if (a.K > 32) {
                    tmem_ld32(trow + 64 + (uint32_t)half * 64 + 32, v);
#pragma unroll
                    for (int k = 0; k < 32; ++k)
                        if (32 + k < a.K) part[(size_t)(32 + k) * VT_C + ch] = v[k];
                }
            }
            tc_fence_before();
            if (half == 1 && wt - 128 < a.K) a.asum[(size_t)item * a.K + (wt - 128)] = asum_k;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
    }
}

}  // namespace

// host entry used by vlad.cu: returns 0 on success, PAB_EINVAL if the shape is not supported (caller falls back to SIMT)
int pab_vlad_tc_partial(int b, int n, int c, int K, const float *x, const void *wc_hi, const void *wc_lo, const float *shift,
                        float *part, float *asum, int *nchunk_out, cudaStream_t st) {
    if (c != VT_C || K <= 0 || K > 64 || !wc_hi) return PAB_EINVAL;
    VtArgs a;
    static int n_sm = 0;
    if (!n_sm) {
        int dev = 0;
        PAB_CUDA(cudaGetDevice(&dev));
        PAB_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    }
    // work items = (cloud, chunk of rows).  Every chunk is another (K, 256) partial block through HBM, so large clouds use
    // 512-row chunks (1024-row chunks make the kernel itself 6 % faster, but 512 fill the SMs the other streams leave free
    // better: +1.7 % descriptors/s in stream mode, A/B on the same box).  The chunking depends on n ONLY: the fp32 summation order — and with it every output bit — must not
    // change with the batch size (a database built in batches of 32 has to match single-cloud queries exactly).
    const int rpi = n >= 2048 ? VT_ROWS_BIG : VT_ROWS_MIN;
    a.rows_per_item = rpi;
    a.n = n; a.K = K; a.Kp = (K + 15) / 16 * 16; a.nchunk = (n + rpi - 1) / rpi; a.nitems = b * a.nchunk;
    a.trace = g_tc_trace;
    a.planes = wc_lo ? 2 : 1;
    a.x = x; a.shift = shift; a.wc_hi = (const __nv_bfloat16 *)wc_hi; a.wc_lo = (const __nv_bfloat16 *)wc_lo; a.part = part; a.asum = asum;
    *nchunk_out = a.nchunk;
    const size_t smem = 10 * (size_t)A_CHUNK + 8 * (size_t)a.Kp * 128 + 64 + 64 * 4 + 4 * TM * 4 + 64;
    PAB_CUDA(cudaFuncSetAttribute(vlad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = a.nitems < n_sm ? a.nitems : n_sm;
    if (g_tc_max_ctas > 0 && grid > g_tc_max_ctas) grid = g_tc_max_ctas;
    vlad_tc_kernel<<<grid, VT_THREADS, smem, st>>>(a);
    PAB_LAUNCH_CHECK();
    return 0;
}
