// afa_tc.cu — AdaptiveFeatureAggregator head (patch_aug_net/models/loupe.py:23-60: MLPAttentionLayer conv1d -> max over
// channels -> softmax over clusters, re-weighting, ReLU, fc, BatchNorm, L2 norm) with its two products on the 5th-generation
// tensor cores, sm_100a.  Same arithmetic contract as the other tcgen05 kernels: every operand is split into bf16 hi/lo
// planes, hi*hi + lo*hi + hi*lo per product, fp32 accumulation in tensor memory.
//
// The fp32 SIMT kernels of vlad.cu (afa_att_kernel 37 us, afa_fc_kernel 39 us for 32 clouds) are latency-bound chains of
// dependent weight loads: the fc layer streams its 22 MB weight at 0.6 TB/s.  Here
//   * afa_att_tc_kernel: logits[(cloud, k), c'] = sum_c v[cloud][c][k] * W_att[c'][c] as a GEMM with the (cloud, cluster) pairs as
//     rows: 128-row tiles x 64-column slices of c' per CTA, A staged from v (cluster index contiguous -> coalesced), the row
//     maximum over c' taken by the row's thread straight from the accumulator;
//   * afa_fc_tc_kernel: part[cta][cloud][o] = sum_{f in the CTA's 64-wide chunks} relu(v + v * wsm)[cloud][f] * W_fc[o][f], one
//     persistent CTA per SM, the weight chunks (256 x 64 x hi/lo = 64 KB) double-buffered with cp.async under the MMAs of the
//     previous chunk, the accumulator staying in tensor memory across the CTA's chunks;
//   * the softmax over clusters is recomputed by every fc CTA for the clouds of its tile (a few thousand exponentials) instead of
//     a launch of its own; the final sum over CTAs + BatchNorm + L2 norm stays afa_finalize_kernel of vlad.cu: three launches.
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int AT_THREADS = 512;
constexpr int AT_NCOL = 64;                   // logit columns (c') per CTA

__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}

// mx[z][cloud][k] = max over c' in [64 z, 64 z + 64) of the attention logit of (cloud, k)
__global__ void __launch_bounds__(AT_THREADS, 1) afa_att_tc_kernel(int b, int c, int K, const float *__restrict__ v,
                                                                  const uint16_t *__restrict__ w_hi, const uint16_t *__restrict__ w_lo,
                                                                  float *__restrict__ mx) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int nchunk = c / KCH;
    uint8_t *a1 = smem, *a2 = a1 + (size_t)nchunk * A_CHUNK;                   // A planes: [chunk][128 rows][128 B]
    uint8_t *b1 = a2 + (size_t)nchunk * A_CHUNK;                               // B planes: [chunk][64 rows][128 B]
    uint8_t *b2 = b1 + (size_t)nchunk * (AT_NCOL * 128);
    uint64_t *bar = reinterpret_cast<uint64_t *>(b2 + (size_t)nchunk * (AT_NCOL * 128));
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar + 1);

    const int tid = threadIdx.x, warp = uniform_warp_idx();
    const long m0 = (long)blockIdx.x * TM, rows = (long)b * K;
    const int n0 = blockIdx.y * AT_NCOL;
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // B: rows n0 .. n0+63 of the weight planes ([c'][c] bf16), 16-byte units into the swizzled K-major layout
    for (int e = tid; e < AT_NCOL * (c / 8); e += AT_THREADS) {
        const int o = e / (c / 8), j = e - o * (c / 8);
        const uint32_t off = (uint32_t)((j >> 3) * (AT_NCOL * 128) + o * 128 + (((j & 7) ^ (o & 7)) << 4));
        cp_async16(b1 + off, w_hi + (size_t)(n0 + o) * c + 8 * j);
        cp_async16(b2 + off, w_lo + (size_t)(n0 + o) * c + 8 * j);
    }
    // A: row r = (cloud, k) pair m0 + r; thread (r, q) stages the channels of quarter q — for a fixed channel the 32 rows of a warp
    // are consecutive cluster indices of v (b, c, K): coalesced
    {
        const int r = tid & (TM - 1), q = tid >> 7;
        const long gr = m0 + r;
        const bool valid = gr < rows;
        const long cloud = valid ? gr / K : 0;
        const float *src = v + cloud * (long)c * K + (valid ? gr - cloud * K : 0);
        const int upq = (c / 8) / 4;                                           // 16-byte units per quarter (<= 8)
        float x[8][8];                                                         // every load of the thread in flight at once
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) x[u][i] = (valid && u < upq) ? __ldg(src + (long)(8 * (q * upq + u) + i) * K) : 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (u < upq) store_units(a1, a2, r, q * upq + u, x[u]);
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (warp == 0) {
        const uint32_t leader = elect_one();
        const uint32_t idesc = umma_idesc(AT_NCOL);
        const uint32_t a1d = umma_desc_lo(smem_u32(a1)), a2d = umma_desc_lo(smem_u32(a2));
        const uint32_t b1d = umma_desc_lo(smem_u32(b1)), b2d = umma_desc_lo(smem_u32(b2));
        for (int kc = 0; kc < nchunk; ++kc) {
            const uint32_t ao = (uint32_t)kc * (A_CHUNK >> 4), bo = (uint32_t)kc * ((AT_NCOL * 128) >> 4);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                umma_f16_if(leader, tmem, a1d + ao + 2 * ks, UMMA_DESC_HI, b1d + bo + 2 * ks, UMMA_DESC_HI, idesc, (kc | ks) != 0);
                umma_f16_if(leader, tmem, a2d + ao + 2 * ks, UMMA_DESC_HI, b1d + bo + 2 * ks, UMMA_DESC_HI, idesc, 1);
            }
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
                umma_f16_if(leader, tmem, a1d + ao + 2 * ks, UMMA_DESC_HI, b2d + bo + 2 * ks, UMMA_DESC_HI, idesc, 1);
        }
        umma_commit_if(leader, bar);
        __syncwarp();
    }
    if (warp < 4) {                                                            // one thread per accumulator row
        mbar_wait(bar, 0);
        tc_fence_after();
        const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
        float m = -INFINITY;
#pragma unroll
        for (int cb = 0; cb < AT_NCOL; cb += 32) {
            float x[32];
            tmem_ld32(trow + (uint32_t)cb, x);
#pragma unroll
            for (int i = 0; i < 32; ++i) m = fmaxf(m, x[i]);
        }
        const long gr = m0 + tid;
        if (gr < rows) mx[(size_t)blockIdx.y * rows + gr] = m;                 // (z, cloud, k) with (cloud, k) = gr
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
    }
}

// part[cta][cloud][o] = sum over the CTA's chunks of relu(v + v * wsm)[cloud][f] * W_fc[o][f]
constexpr int FC_BSTAGE = 256 * 128;          // bytes of one weight plane of one chunk (256 output rows x 64 k)
// wsm: softmax weights (b, K) from afa_softmax_kernel — or, with mx != NULL, the kernel computes them itself from the nparts
// partial maxima of the logits (every CTA for the clouds of its tile, into shared memory: one launch and one dependent
// round trip less; needs nb * K floats of shared memory)
__global__ void __launch_bounds__(AT_THREADS, 1) afa_fc_tc_kernel(int b, int c, int K, int c_out, const float *__restrict__ v,
                                                                 const float *__restrict__ wsm, const float *__restrict__ mx, int nparts,
                                                                 const uint16_t *__restrict__ w_hi,
                                                                 const uint16_t *__restrict__ w_lo, float *__restrict__ part) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *bst = smem;                                                       // [stage][plane][256 rows][128 B]
    uint8_t *ast = smem + 4 * (size_t)FC_BSTAGE;                               // [stage][plane][128 rows][128 B]
    uint64_t *done = reinterpret_cast<uint64_t *>(ast + 4 * (size_t)A_CHUNK);  // [2] MMAs of the chunk in this stage completed
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(done + 2);
    float *wsm_s = reinterpret_cast<float *>(done + 8);                        // fused softmax: [clouds of the tile][K]

    const int tid = threadIdx.x, warp = uniform_warp_idx();
    const long F = (long)c * K;
    const int nchunks = (int)(F / KCH);
    const int cloud0 = blockIdx.y * TM, nb = min(TM, b - cloud0);
    const int my_n = blockIdx.x < nchunks ? (nchunks - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;   // chunks x, x + grid, ...
    if (tid == 0) {
        mbar_init(done, 1);
        mbar_init(done + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (mx) {                                                                  // softmax over the clusters, one warp per cloud
        const int lane = tid & 31;
        for (int row = warp; row < nb; row += AT_THREADS / 32) {
            const long base = (long)(cloud0 + row) * K;
            float mm = -INFINITY;
            for (int k = lane; k < K; k += 32) {
                float a = __ldg(mx + base + k);
                for (int z = 1; z < nparts; ++z) a = fmaxf(a, __ldg(mx + (long)z * b * K + base + k));
                wsm_s[row * K + k] = a;
                mm = fmaxf(mm, a);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mm = fmaxf(mm, __shfl_xor_sync(0xffffffffu, mm, o));
            float sum = 0.f;
            for (int k = lane; k < K; k += 32) {
                const float e = expf(wsm_s[row * K + k] - mm);
                wsm_s[row * K + k] = e;
                sum += e;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            for (int k = lane; k < K; k += 32) wsm_s[row * K + k] = wsm_s[row * K + k] / sum;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    auto stage_b = [&](int ch, int s) {                                        // 2 planes x c_out rows x 8 units, cp.async
        for (int e = tid; e < c_out * 8; e += AT_THREADS) {
            const int o = e >> 3, j = e & 7;
            const uint32_t off = (uint32_t)(o * 128 + ((j ^ (o & 7)) << 4));
            const size_t src = (size_t)o * F + (size_t)ch * KCH + 8 * j;
            cp_async16(bst + (size_t)(2 * s) * FC_BSTAGE + off, w_hi + src);
            cp_async16(bst + (size_t)(2 * s + 1) * FC_BSTAGE + off, w_lo + src);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // A: y = relu(v + v * wsm[cloud][f % K]) for the chunk's 64 features of every cloud of the tile; (cloud, unit) pairs over the
    // threads, the loads of a pair issued together, the stores after the caller's wait
    const bool plain = !wsm && !mx;                                            // y = x: a plain fc over the flattened vector (PPT-Net head)
    constexpr int AR = (TM * 8 + AT_THREADS - 1) / AT_THREADS;                 // pairs per thread (2)
    float ax[AR][8], aw[AR][8];
    auto load_a = [&](int ch) {
#pragma unroll
        for (int rr = 0; rr < AR; ++rr) {
            const int e = tid + rr * AT_THREADS, row = e >> 3, j = e & 7;
            if (row < nb) {
                const long f0 = (long)ch * KCH + 8 * j;
                const float4 *p = reinterpret_cast<const float4 *>(v + (long)(cloud0 + row) * F + f0);
                const float4 x0 = __ldg(p), x1 = __ldg(p + 1);
                ax[rr][0] = x0.x; ax[rr][1] = x0.y; ax[rr][2] = x0.z; ax[rr][3] = x0.w;
                ax[rr][4] = x1.x; ax[rr][5] = x1.y; ax[rr][6] = x1.z; ax[rr][7] = x1.w;
                if (!plain) {
                    int k = (int)(f0 % K);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        aw[rr][i] = mx ? wsm_s[row * K + k] : __ldg(wsm + (long)(cloud0 + row) * K + k);
                        if (++k == K) k = 0;
                    }
                }
            }
        }
    };
    auto store_a = [&](int s) {
        uint8_t *a1 = ast + (size_t)(2 * s) * A_CHUNK, *a2 = a1 + A_CHUNK;
#pragma unroll
        for (int rr = 0; rr < AR; ++rr) {
            const int e = tid + rr * AT_THREADS, row = e >> 3, j = e & 7;
            if (row < nb) {
                float y[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) y[i] = plain ? ax[rr][i] : fmaxf(ax[rr][i] + ax[rr][i] * aw[rr][i], 0.f);
                store_units(a1, a2, row, j, y);
            }
        }
    };

    if (my_n > 0) {
        stage_b((int)blockIdx.x, 0);
        load_a((int)blockIdx.x);
        store_a(0);
    }
    const uint32_t leader = elect_one();
    const uint32_t idesc = umma_idesc(c_out);
    const uint32_t ad = umma_desc_lo(smem_u32(ast)), bd = umma_desc_lo(smem_u32(bst));
    for (int it = 0; it < my_n; ++it) {
        const int s = it & 1;
        const bool more = it + 1 < my_n;
        if (more) {
            if (it >= 1) mbar_wait(done + (s ^ 1), ((it - 1) >> 1) & 1);       // the chunk that used the other stage has been multiplied
            const int chn = (int)blockIdx.x + (it + 1) * (int)gridDim.x;
            stage_b(chn, s ^ 1);
            load_a(chn);
        }
        if (more) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (more) store_a(s ^ 1);
        fence_proxy_async();
        __syncthreads();
        if (warp == 0) {
            tc_fence_after();
            const uint32_t a1d = ad + (uint32_t)(2 * s) * (A_CHUNK >> 4), a2d = a1d + (A_CHUNK >> 4);
            const uint32_t b1d = bd + (uint32_t)(2 * s) * (FC_BSTAGE >> 4), b2d = b1d + (FC_BSTAGE >> 4);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                umma_f16_if(leader, tmem, a1d + 2 * ks, UMMA_DESC_HI, b1d + 2 * ks, UMMA_DESC_HI, idesc, (it | ks) != 0);
                umma_f16_if(leader, tmem, a2d + 2 * ks, UMMA_DESC_HI, b1d + 2 * ks, UMMA_DESC_HI, idesc, 1);
            }
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) umma_f16_if(leader, tmem, a1d + 2 * ks, UMMA_DESC_HI, b2d + 2 * ks, UMMA_DESC_HI, idesc, 1);
            umma_commit_if(leader, done + s);
            __syncwarp();
        }
    }
    if (warp < 4) {
        const int row = tid;                                                   // accumulator row = cloud of the tile
        float *dst = part + ((size_t)blockIdx.x * b + cloud0 + row) * c_out;
        if (my_n > 0) {
            const int last = my_n - 1;
            mbar_wait(done + (last & 1), (last >> 1) & 1);                     // tcgen05.commit covers every earlier MMA as well
            tc_fence_after();
            const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
            for (int cb = 0; cb < c_out; cb += 32) {
                float x[32];
                tmem_ld32(trow + (uint32_t)cb, x);                             // warp-convergent; only the stores are predicated
                if (row < nb) {
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        if (cb + 4 * u < c_out)
                            *reinterpret_cast<float4 *>(dst + cb + 4 * u) = make_float4(x[4 * u], x[4 * u + 1], x[4 * u + 2], x[4 * u + 3]);
                }
            }
        } else if (row < nb) {
            for (int o = 0; o < c_out; ++o) dst[o] = 0.f;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
    }
}

int g_afa_tc = 1;

int afa_sm_count() {
    static int n_sm = 0;
    if (!n_sm) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n_sm = 148;
    }
    return n_sm;
}

}  // namespace

PAB_API void pab_tune_afa_tc(int on) { g_afa_tc = on; }

// shapes the two tensor-core kernels take
int pab_afa_tc_eligible(int c, int K, int c_out) {
    return g_afa_tc && c % 64 == 0 && c >= 64 && c <= 256 && ((long)c * K) % 64 == 0 && c_out % 32 == 0 && c_out >= 32 && c_out <= 256;
}
int pab_afa_tc_att_parts(int c) { return c / AT_NCOL; }
int pab_afa_tc_fc_slices(int c, int K) {
    const long nchunks = (long)c * K / KCH;
    return (int)(nchunks < afa_sm_count() ? nchunks : afa_sm_count());
}

int pab_afa_tc_att(int b, int c, int K, const float *v, const void *w_hi, const void *w_lo, float *mx, cudaStream_t st) {
    const int nchunk = c / KCH;
    const size_t smem = (size_t)nchunk * (2 * A_CHUNK + 2 * AT_NCOL * 128) + 64;
    static size_t configured = 0;
    if (smem > configured) {
        PAB_CUDA(cudaFuncSetAttribute(afa_att_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    const long rows = (long)b * K;
    afa_att_tc_kernel<<<dim3((unsigned)((rows + TM - 1) / TM), c / AT_NCOL), AT_THREADS, smem, st>>>(
        b, c, K, v, (const uint16_t *)w_hi, (const uint16_t *)w_lo, mx);
    PAB_LAUNCH_CHECK();
    return 0;
}

// 1: the fc kernel computes the softmax itself (the tile's clouds x K weights fit next to its stages in shared memory)
int pab_afa_tc_fused_softmax(int b, int K) {
    const long nb = b < TM ? b : TM;
    return nb * K * 4 <= 32 * 1024;
}

int pab_afa_tc_fc(int b, int c, int K, int c_out, const float *v, const float *wsm, const float *mx, int nparts, const void *w_hi,
                  const void *w_lo, float *part, cudaStream_t st) {
    const long nb = b < TM ? b : TM;
    const size_t smem = 4 * (size_t)FC_BSTAGE + 4 * (size_t)A_CHUNK + 64 + (mx ? (size_t)nb * K * 4 : 0);
    static size_t configured = 0;
    if (smem > configured) {
        PAB_CUDA(cudaFuncSetAttribute(afa_fc_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    afa_fc_tc_kernel<<<dim3(pab_afa_tc_fc_slices(c, K), (b + TM - 1) / TM), AT_THREADS, smem, st>>>(
        b, c, K, c_out, v, wsm, mx, nparts, (const uint16_t *)w_hi, (const uint16_t *)w_lo, part);
    PAB_LAUNCH_CHECK();
    return 0;
}
