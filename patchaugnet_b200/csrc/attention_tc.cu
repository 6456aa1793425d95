// attention_tc.cu — the two N x N passes of PPT-Net's SA_Layer on the tcgen05 tensor cores (sm_100a).
//
// Reference: place_recognition/pptnet_origin/models/pptnet.py:261-282.  Same algebra as attention.cu (energy = Q^T Q is a
// symmetric Gram matrix because q_conv and k_conv share one weight; nothing of size N x N leaves the SM), but both
// contractions of each pass run as tcgen05.mma with fp32 accumulators in tensor memory:
//
//   stats pass   for a tile of 128 points j:  S[j][i] = Q_j . Q_i over all tiles i  ->  online row max / row sum of row j
//   apply pass   for a tile of 128 points j:  S[j][i] (= E[i][j], symmetric), P^T[j][i] = exp(S - max_i) / sum_i  (softmax of
//                ROW i, pptnet.py:277), colsum_j += sum_i P^T[j][i] (pptnet.py:278),  O[j][c] += P^T[j][i] . V[i][c]
//                (pptnet.py:279), finally d[j] = x[j] - O[j] / (1e-9 + colsum_j)
//
// One CTA per (cloud, 128-point tile j).  Warp 0 issues the MMAs (warp-uniform loop, one elected lane); warps 1..8 stage the
// operands — fp32 rows of [Q | V] split into bf16 hi/lo planes in the canonical 128-byte-swizzled layout — and run the
// softmax straight out of tensor memory: a thread owns row j (a TMEM lane), the statistics of the columns i come from a
// small shared-memory table, and P^T goes back INTO tensor memory as the A operand of the second GEMM (TS form), whose B
// operand is the V tile read MN-major (a point's channels are contiguous).  S is double-buffered in TMEM.
// TMEM columns: S0 [0,128)  S1 [128,256)  O [256,256+C)  P^T hi [384,448)  P^T lo [448,512).
// precision 2 (fp32 contract): three MMAs per product on the hi/lo planes; precision 1: plain bf16 operands, one MMA.
#include <math.h>
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int AW = 256;                       // worker threads (warps 1..8)
constexpr int AT_THREADS = 32 + AW;
constexpr uint32_t O_COL = 256, PH_COL = 384, PL_COL = 448;

struct AtArgs {
    int n, C, planes, apply;
    const float *q, *v;                       // row stride ld (floats)
    long ld;
    const float *x;                           // (b, n, C), apply only
    float *rowmax, *rowsum;                   // (b, n): written by the stats pass, read by the apply pass
    float *d;                                 // (b, n, C), apply only
};

__device__ __forceinline__ uint32_t umma_desc_mn_lo(uint32_t saddr) { return ((saddr & 0x3FFFF) >> 4) | ((uint32_t)(A_CHUNK >> 4) << 16); }

__device__ __forceinline__ void umma_ts_if(uint32_t issue, uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p, q;\n"
        ".reg .b64 db;\n"
        "setp.ne.b32 p, %5, 0;\n"
        "setp.ne.b32 q, %6, 0;\n"
        "mov.b64 db, {%2, %3};\n"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n"
        "}\n" ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate), "r"(issue) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
          "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}

__global__ void __launch_bounds__(AT_THREADS, 1) attn_tc_kernel(const AtArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int nchunk = a.C / KCH;
    const size_t plane = (size_t)nchunk * A_CHUNK;                 // one bf16 plane of a [128][C] tile
    uint8_t *qj1 = smem, *qj2 = qj1 + plane;                       // Q_j hi / lo  (A of GEMM1, K-major)
    uint8_t *qi1 = qj2 + plane, *qi2 = qi1 + plane;                // Q_i hi / lo  (B of GEMM1, K-major)
    uint8_t *vi1 = qi2 + plane, *vi2 = vi1 + plane;                // V_i hi / lo  (B of GEMM2, MN-major); apply pass only
    uint8_t *misc = a.apply ? vi2 + plane : vi1;
    uint64_t *ops_ready = reinterpret_cast<uint64_t *>(misc);      // workers -> MMA: Q_i (and V_i) staged               (AW arrivals)
    uint64_t *s_ready = ops_ready + 1;                             // MMA -> workers: S[buf] accumulated                  [2], commit
    uint64_t *s_free = ops_ready + 3;                              // workers -> MMA: S[buf] read                         [2], AW arrivals
    uint64_t *p_ready = ops_ready + 5;                             // workers -> MMA: P^T planes written                  (AW arrivals)
    uint64_t *g_done = ops_ready + 6;                              // MMA -> workers: GEMM1 (stats) / GEMM2 (apply) of the tile done: operands free
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(ops_ready + 7);
    float *cmax = reinterpret_cast<float *>(misc + 64);            // [128] row max of the points i of the tile
    float *cinv = cmax + TM;                                       // [128] 1 / row sum (0 for padded points)
    float *comb = cinv + TM;                                       // [2][128] exchange between the two column halves

    const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;
    const int cloud = blockIdx.y, j0 = blockIdx.x * TM;
    const float *q = a.q + (size_t)cloud * a.n * a.ld, *v = a.v + (size_t)cloud * a.n * a.ld;
    const int ntile = (a.n + TM - 1) / TM;
    const bool two = a.planes == 2;

    if (tid == 0) {
        mbar_init(ops_ready, AW);
        mbar_init(s_ready, 1); mbar_init(s_ready + 1, 1);
        mbar_init(s_free, AW); mbar_init(s_free + 1, AW);
        mbar_init(p_ready, AW);
        mbar_init(g_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    // rows [r0, r0+128) of src (row stride ld, C channels) -> hi/lo planes; rows beyond n are zero
    auto stage_tile = [&](const float *src, int r0, uint8_t *p1, uint8_t *p2, int wt) {
        const int wwarp = wt >> 5, lr = lane >> 3, lu = lane & 7;
        for (int c = 0; c < nchunk; ++c)
#pragma unroll
            for (int st4 = 0; st4 < 4; ++st4) {
                const int r = wwarp * 16 + st4 * 4 + lr;
                float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
                if (r0 + r < a.n) {
                    const float4 *s4 = reinterpret_cast<const float4 *>(src + (size_t)(r0 + r) * a.ld + c * 64) + 2 * lu;
                    a0 = __ldg(s4); a1 = __ldg(s4 + 1);
                }
                const float vv[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                store_units(p1, two ? p2 : nullptr, r, c * 8 + lu, vv);
            }
    };

    if (warp == 0) {
        // ================= MMA issuer =====================================================================================
        const uint32_t leader = elect_one();
        const uint32_t id1 = umma_idesc(TM);                                  // GEMM1: N = 128 points i
        const uint32_t id2 = umma_idesc(64) | (1u << 16);                     // GEMM2: N = 64 channels, B MN-major
        const uint32_t qj1_lo = umma_desc_lo(smem_u32(qj1)), qj2_lo = umma_desc_lo(smem_u32(qj2));
        const uint32_t qi1_lo = umma_desc_lo(smem_u32(qi1)), qi2_lo = umma_desc_lo(smem_u32(qi2));
        const uint32_t vi1_lo = umma_desc_mn_lo(smem_u32(vi1)), vi2_lo = umma_desc_mn_lo(smem_u32(vi2));
        for (int it = 0; it < ntile; ++it) {
            const uint32_t buf = (uint32_t)it & 1u, use = (uint32_t)it >> 1;
            mbar_wait(ops_ready, it & 1);
            if (it >= 2) mbar_wait(s_free + buf, (use - 1) & 1);              // S[buf] of tile it-2 has been read
            tc_fence_after();
            const uint32_t dS = tmem + buf * TM;
            for (int kc = 0; kc < nchunk; ++kc)
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint32_t o = (uint32_t)kc * (A_CHUNK >> 4) + 2 * ks;
                    umma_f16_if(leader, dS, qj1_lo + o, UMMA_DESC_HI, qi1_lo + o, UMMA_DESC_HI, id1, (kc | ks) != 0);
                    if (two) {
                        umma_f16_if(leader, dS, qj2_lo + o, UMMA_DESC_HI, qi1_lo + o, UMMA_DESC_HI, id1, 1);
                        umma_f16_if(leader, dS, qj1_lo + o, UMMA_DESC_HI, qi2_lo + o, UMMA_DESC_HI, id1, 1);
                    }
                }
            umma_commit_if(leader, s_ready + buf);
            if (!a.apply) {
                umma_commit_if(leader, g_done);                               // Q_i may be overwritten
                continue;
            }
            mbar_wait(p_ready, it & 1);
            tc_fence_after();
            for (int cb = 0; cb < nchunk; ++cb)                               // 64 output channels per block
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {                              // 16 points i per MMA
                    const uint32_t vb = (uint32_t)cb * (A_CHUNK >> 4) + (uint32_t)ks * (2048 >> 4);
                    const uint32_t dO = tmem + O_COL + (uint32_t)cb * 64;
                    umma_ts_if(leader, dO, tmem + PH_COL + ks * 8, vi1_lo + vb, UMMA_DESC_HI, id2, (it | ks) != 0);
                    if (two) {
                        umma_ts_if(leader, dO, tmem + PL_COL + ks * 8, vi1_lo + vb, UMMA_DESC_HI, id2, 1);
                        umma_ts_if(leader, dO, tmem + PH_COL + ks * 8, vi2_lo + vb, UMMA_DESC_HI, id2, 1);
                    }
                }
            umma_commit_if(leader, g_done);                                   // Q_i, V_i and the P^T planes are free
        }
        __syncwarp();
    } else {
        // ================= workers: staging + softmax ======================================================================
        const int wt = tid - 32, wwarp = wt >> 5;
        const int qd = warp & 3, half = wwarp >> 2;                           // TMEM lane quarter of this warp; column half
        const int row = qd * 32 + lane;                                       // row j owned by this thread
        const uint32_t trow = tmem + ((uint32_t)(qd * 32) << 16);
        float m_run = -INFINITY, s_run = 0.f, csum = 0.f;
        stage_tile(q, j0, qj1, qj2, wt);                                      // Q_j once
        for (int it = 0; it < ntile; ++it) {
            const int i0 = it * TM;
            const uint32_t buf = (uint32_t)it & 1u, use = (uint32_t)it >> 1;
            if (it > 0) mbar_wait(g_done, (it - 1) & 1);                      // previous tile's MMAs no longer read Q_i / V_i / P^T
            stage_tile(q, i0, qi1, qi2, wt);
            if (a.apply) {
                stage_tile(v, i0, vi1, vi2, wt);
                if (wt < TM) {
                    const bool ok = i0 + wt < a.n;
                    cmax[wt] = ok ? __ldg(a.rowmax + (size_t)cloud * a.n + i0 + wt) : 0.f;
                    cinv[wt] = ok ? 1.f / __ldg(a.rowsum + (size_t)cloud * a.n + i0 + wt) : 0.f;     // 0 kills the padded points
                }
            }
            fence_proxy_async();
            mbar_arrive(ops_ready);
            if (a.apply) asm volatile("bar.sync 1, %0;" ::"n"(AW) : "memory");   // cmax / cinv visible to every worker
            mbar_wait(s_ready + buf, use & 1);
            tc_fence_after();
            const uint32_t sbase = trow + buf * TM + (uint32_t)half * 64;
            if (!a.apply) {
                // online softmax statistics of row j over this tile's columns (64 per thread)
                float sv[2][32];
                tmem_ld32(sbase, sv[0]);
                tmem_ld32(sbase + 32, sv[1]);
                tc_fence_before();
                mbar_arrive(s_free + buf);
                float tm = m_run;
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int k = 0; k < 32; ++k)
                        if (i0 + half * 64 + h * 32 + k < a.n) tm = fmaxf(tm, sv[h][k]);
                float ts = s_run * __expf(m_run - tm);
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int k = 0; k < 32; ++k)
                        if (i0 + half * 64 + h * 32 + k < a.n) ts += __expf(sv[h][k] - tm);
                m_run = tm; s_run = ts;
            } else {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float sv[32];
                    tmem_ld32(sbase + h * 32, sv);
                    if (h == 1) {                                             // every S column of this thread has been read
                        tc_fence_before();
                        mbar_arrive(s_free + buf);
                    }
                    const int cb = half * 64 + h * 32;
                    uint32_t hi[16], lo[16];
#pragma unroll
                    for (int k = 0; k < 32; k += 2) {
                        const float pa = __expf(sv[k] - cmax[cb + k]) * cinv[cb + k];
                        const float pb = __expf(sv[k + 1] - cmax[cb + k + 1]) * cinv[cb + k + 1];
                        const __nv_bfloat162 hh = __floats2bfloat162_rn(pa, pb);
                        const float2 hf = __bfloat1622float2(hh);
                        hi[k >> 1] = *reinterpret_cast<const uint32_t *>(&hh);
                        if (two) {
                            const __nv_bfloat162 ll = __floats2bfloat162_rn(pa - hf.x, pb - hf.y);
                            const float2 lf = __bfloat1622float2(ll);
                            lo[k >> 1] = *reinterpret_cast<const uint32_t *>(&ll);
                            csum += (hf.x + lf.x) + (hf.y + lf.y);            // what the tensor cores add up
                        } else {
                            csum += hf.x + hf.y;
                        }
                    }
                    tmem_st16(trow + PH_COL + (uint32_t)(cb >> 1), hi);
                    if (two) tmem_st16(trow + PL_COL + (uint32_t)(cb >> 1), lo);
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                mbar_arrive(p_ready);
            }
        }
        if (!a.apply) {
            // combine the two column halves of row j
            comb[half * TM + row] = m_run;
            comb[2 * TM + half * TM + row] = s_run;
            asm volatile("bar.sync 1, %0;" ::"n"(AW) : "memory");
            if (half == 0 && j0 + row < a.n) {
                const float m0 = comb[row], m1 = comb[TM + row], s0 = comb[2 * TM + row], s1 = comb[3 * TM + row];
                const float m = fmaxf(m0, m1);
                a.rowmax[(size_t)cloud * a.n + j0 + row] = m;
                a.rowsum[(size_t)cloud * a.n + j0 + row] = s0 * __expf(m0 - m) + s1 * __expf(m1 - m);
            }
        } else {
            comb[half * TM + row] = csum;
            asm volatile("bar.sync 1, %0;" ::"n"(AW) : "memory");
            const float inv = 1.f / (1e-9f + (comb[row] + comb[TM + row]));
            mbar_wait(g_done, (ntile - 1) & 1);
            tc_fence_after();
            // O[j][c]: this thread takes channels [half * C/2, (half+1) * C/2) of its row
            const int cw = a.C >> 1;
            const bool ok = j0 + row < a.n;
            const size_t ro = ((size_t)cloud * a.n + (ok ? j0 + row : 0)) * a.C + half * cw;
            const float *xr = a.x + ro;
            float *dr = a.d + ro;
            for (int c0 = 0; c0 < cw; c0 += 32) {
                float o[32];
                tmem_ld32(trow + O_COL + (uint32_t)(half * cw + c0), o);     // warp-collective: outside the row predicate
                if (ok) {
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const float4 xv = __ldg(reinterpret_cast<const float4 *>(xr + c0) + u);
                        *reinterpret_cast<float4 *>(dr + c0 + 4 * u) =
                            make_float4(xv.x - o[4 * u] * inv, xv.y - o[4 * u + 1] * inv, xv.z - o[4 * u + 2] * inv, xv.w - o[4 * u + 3] * inv);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

}  // namespace

// host entry used by attention.cu.  precision: 2 = bf16 hi/lo (fp32 contract), 1 = plain bf16.  Returns PAB_EINVAL when the
// shape is not supported (the caller then takes the SIMT kernels).
int pab_attention_tc(int b, int n, int c, const float *q, const float *v, long ld, const float *x, float *rowmax, float *rowsum,
                     float *d, int precision, cudaStream_t st) {
    if (!(c == 64 || c == 128) || n < 64 || b <= 0 || b > 65535) return PAB_EINVAL;
    AtArgs a;
    a.n = n; a.C = c; a.planes = precision == 1 ? 1 : 2; a.q = q; a.v = v; a.ld = ld; a.x = x; a.rowmax = rowmax; a.rowsum = rowsum; a.d = d;
    const size_t plane = (size_t)(c / KCH) * A_CHUNK;
    const dim3 grid(pab_divup(n, TM), b);
    const size_t misc = 64 + 5 * TM * 4 + 64;
    const size_t smem_stats = 4 * plane + misc + 1024, smem_apply = 6 * plane + misc + 1024;
    PAB_CUDA(cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_apply));
    a.apply = 0;
    attn_tc_kernel<<<grid, AT_THREADS, smem_stats, st>>>(a);
    PAB_LAUNCH_CHECK();
    a.apply = 1;
    attn_tc_kernel<<<grid, AT_THREADS, smem_apply, st>>>(a);
    PAB_LAUNCH_CHECK();
    return 0;
}
