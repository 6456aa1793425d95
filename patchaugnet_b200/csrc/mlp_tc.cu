// mlp_tc.cu — fused SharedMLP on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
// Same fusion as mlp.cu (neighbour gather / 3-NN interpolation -> every MLP layer -> max over K or row store, tile
// resident on chip), but the layer GEMMs run as tcgen05.mma with fp32 accumulators in tensor memory:
//
//   * a persistent CTA owns 128-row tiles; the layer input lives in shared memory as TWO bf16 planes
//     (x = hi + lo, hi = bf16(x), lo = bf16(x - hi)) in the canonical K-major SWIZZLE_128B layout; the folded weights
//     are split the same way on the host.  Each 16-wide k-step issues three MMAs
//         D += hi(A) * hi(W)  +  lo(A) * hi(W)  +  hi(A) * lo(W)
//     which carries ~16 mantissa bits through every product (error ~2^-17 per term, fp32 accumulation) — the
//     descriptors stay within the 1e-4 contract where a single TF32/bf16 pass does not.
//   * warp 0 streams weight blocks (128 output channels x 64 k, hi+lo) with TMA (cp.async.bulk.tensor, mbarrier
//     complete_tx) through a ring of stages, running ahead across layers and tiles; warp 1 (one elected thread)
//     issues the MMAs and frees stages with tcgen05.commit; warps 2..9 (256 threads) gather the tile, and after
//     each layer read the accumulators with tcgen05.ld, add the folded BN shift (+ the rank-3 xyz update of
//     layer 0, whose K would otherwise not be a multiple of 64), ReLU, split to bf16 hi/lo and write the next layer's
//     operand in place.  The last layer is staged as fp32 in the (now free) operand region and either max-pooled
//     over the K neighbours (SA) or stored as coalesced rows (FP).
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int NBLK_MAX = 128;                 // max output channels per weight stage (UMMA N); runtime a.nblk is 64 or 128
constexpr int NWORK = 256;                    // worker threads
constexpr int TC_THREADS = 64 + NWORK;
constexpr int MAX_STAGES = 6;
constexpr int MAX_LAYERS = 3;

enum { TC_SA = 1, TC_FP = 2 };

struct alignas(64) TcArgs {
    CUtensorMap tm[MAX_LAYERS][2];
    const float *shift[MAX_LAYERS];
    int K[MAX_LAYERS], N[MAX_LAYERS], relu[MAX_LAYERS];
    int n_layers, n_stages, kchunks_max, mode;
    int nblk, stage_bytes;                     // output channels per weight stage, bytes per stage (hi + lo)
    int coff[MAX_LAYERS];                      // offset of each layer's shift vector in the smem constant table
    int a_region;                              // bytes of the operand region (>= the 128 KB fp32 staging of the last layer)
    long rows;                                 // SA: centres, FP: points
    int ntiles;
    // layer-0 "extra" channels (the xyz part), applied as a rank-n update from the fp32 weight rows
    const float *w_extra;                      // layers[0].wt + extra_row0 * N0
    int n_extra;
    // optional "pre" layer: the module's first layer has so few inputs (<= 8: SA1's [xyz_rel ; feat_rel]) that the loader
    // evaluates it in fp32 and stages its OUTPUT as the first tensor-core operand (zero-padded to 64 channels)
    int pre_cin, pre_cout, pre_relu;
    const float *pre_wt, *pre_shift;           // (pre_cin_pad, pre_cout) fp32 folded weight, shift
    int pre_off;                               // offset of [weights | shift] in the smem constant table
    // SA
    int n, m, k, nbr_stride, c;
    const float *xyz, *feat;
    const int *center_idx, *nbr_idx;
    // FP
    int c_known, c_skip;
    const float *known_feat, *skip_feat;
    const int *idx3;
    const float *w3;
    float *out;
};

__global__ void __launch_bounds__(TC_THREADS, 1) mlp_tc_kernel(const __grid_constant__ TcArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *a1 = smem;
    uint8_t *a2 = a1 + (size_t)a.kchunks_max * A_CHUNK;
    uint8_t *stages = a1 + (size_t)a.a_region;
    uint8_t *misc = stages + (size_t)a.n_stages * a.stage_bytes;
    uint64_t *full = reinterpret_cast<uint64_t *>(misc);
    uint64_t *empty = full + MAX_STAGES;
    uint64_t *a_ready = empty + MAX_STAGES;
    uint64_t *d_ready = a_ready + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(d_ready + 1);
    float *ctab = reinterpret_cast<float *>(misc + 256);   // [shift of every layer | 3 x N0 extra weight rows], 16-byte aligned

    const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < a.n_stages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(a_ready, NWORK);
        mbar_init(d_ready, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    {
        int off = 0;
        for (int l = 0; l < a.n_layers; ++l) {
            for (int i = tid; i < a.N[l]; i += TC_THREADS) ctab[off + i] = __ldg(a.shift[l] + i);
            off += a.N[l];
        }
        for (int i = tid; i < a.n_extra * a.N[0]; i += TC_THREADS) ctab[off + i] = __ldg(a.w_extra + i);
        if (a.pre_cout > 0) {
            for (int i = tid; i < a.pre_cin * a.pre_cout; i += TC_THREADS) ctab[a.pre_off + i] = __ldg(a.pre_wt + i);
            for (int i = tid; i < a.pre_cout; i += TC_THREADS) ctab[a.pre_off + a.pre_cin * a.pre_cout + i] = __ldg(a.pre_shift + i);
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    if (warp == 0) {
        // ================= TMA producer: weight blocks, in (tile, layer, n-block, k-chunk) order =================
        if (lane == 0) {
            uint32_t s = 0, ph = 0;
            for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
                for (int l = 0; l < a.n_layers; ++l) {
                    const int nkc = a.K[l] / KCH, nbr = min(a.nblk, a.N[l]), nnb = a.N[l] / nbr;
                    for (int nb = 0; nb < nnb; ++nb)
                        for (int kc = 0; kc < nkc; ++kc) {
                            mbar_wait(empty + s, ph ^ 1);
                            uint8_t *dst = stages + (size_t)s * a.stage_bytes;
                            mbar_expect_tx(full + s, 2u * nbr * 128u);
                            tma_load_2d(dst, &a.tm[l][0], kc * KCH, nb * nbr, full + s);
                            tma_load_2d(dst + nbr * 128, &a.tm[l][1], kc * KCH, nb * nbr, full + s);
                            if (++s == (uint32_t)a.n_stages) { s = 0; ph ^= 1; }
                        }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================= MMA issuer: all lanes run the loops, one elected lane issues ===========================
        const uint32_t leader = elect_one();
        const uint32_t a1_lo = umma_desc_lo(smem_u32(a1)), a2_lo = umma_desc_lo(smem_u32(a2));
        const uint32_t st_lo = umma_desc_lo(smem_u32(stages)), st_step = (uint32_t)a.stage_bytes >> 4;
        uint32_t s = 0, ph = 0, lcount = 0;
        for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
            for (int l = 0; l < a.n_layers; ++l, ++lcount) {
                const int nkc = a.K[l] / KCH, nbr = min(a.nblk, a.N[l]), nnb = a.N[l] / nbr;
                const uint32_t idesc = umma_idesc(nbr);
                mbar_wait(a_ready, lcount & 1);
                tc_fence_after();
                for (int nb = 0; nb < nnb; ++nb) {
                    const uint32_t d = tmem + (uint32_t)(nb * nbr);
                    for (int kc = 0; kc < nkc; ++kc) {
                        mbar_wait(full + s, ph);
                        tc_fence_after();
                        const uint32_t ka = (uint32_t)kc * (A_CHUNK >> 4);
                        const uint32_t sb1 = st_lo + s * st_step, sb2 = sb1 + (uint32_t)nbr * 8;     // lo plane follows hi (nbr x 128 B)
#pragma unroll
                        for (int ks = 0; ks < KCH / 16; ++ks) {
                            umma_f16_if(leader, d, a1_lo + ka + 2 * ks, UMMA_DESC_HI, sb1 + 2 * ks, UMMA_DESC_HI, idesc, (kc | ks) != 0);
                            umma_f16_if(leader, d, a2_lo + ka + 2 * ks, UMMA_DESC_HI, sb1 + 2 * ks, UMMA_DESC_HI, idesc, 1);
                            umma_f16_if(leader, d, a1_lo + ka + 2 * ks, UMMA_DESC_HI, sb2 + 2 * ks, UMMA_DESC_HI, idesc, 1);
                        }
                        umma_commit_if(leader, empty + s);   // stage reusable once these MMAs have read it
                        if (++s == (uint32_t)a.n_stages) { s = 0; ph ^= 1; }
                    }
                }
                umma_commit_if(leader, d_ready);             // accumulators of this layer complete
            }
        }
        __syncwarp();
    } else {
        // ================= workers: gather, epilogues, output ====================================================
        const int wt = tid - 64;                             // 0..255
        const int wwarp = warp - 2;                          // 0..7
        const int q = warp & 3;                              // TMEM lane quarter this warp may access
        const int half = wwarp >> 2;                         // column half handled by this warp
        const int row = q * 32 + lane;                       // accumulator row owned in the epilogue
        const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
        uint32_t lcount = 0;
        const int units0 = a.K[0] / 8;                       // 16-byte units per row of the layer-0 operand
        const float *wext = ctab + a.coff[a.n_layers - 1] + a.N[a.n_layers - 1];   // extra weight rows follow the shifts

        for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
            // ---- stage the layer-0 operand (hi/lo planes): one warp per row, lanes across the row's 16-byte units -----
            float xe[3] = {0.f, 0.f, 0.f};
            if (a.mode == TC_SA && a.pre_cout > 0) {
                // two threads per row: inputs [xyz_j - xyz_i ; f_j - f_i] (pre_cin <= 8 values), each thread evaluates half of
                // the pre-layer's outputs in fp32 and stages them (hi/lo); the rest of the 64-channel chunk is zeroed
                const int G = TM / a.k;
                const int r = wt >> 1, part = wt & 1;
                const int g = r / a.k, sidx = r - g * a.k;
                const long ci = (long)tile * G + g;
                float in[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) in[i] = 0.f;
                const bool valid = g < G && ci < a.rows;
                if (valid) {
                    const long cloud = ci / a.m;
                    const long pc = cloud * a.n + __ldg(a.center_idx + ci);
                    const long pn = cloud * a.n + __ldg(a.nbr_idx + ci * a.nbr_stride + sidx);
#pragma unroll
                    for (int i = 0; i < 3; ++i) in[i] = __ldg(a.xyz + pn * 3 + i) - __ldg(a.xyz + pc * 3 + i);
#pragma unroll
                    for (int i = 0; i < 5; ++i)
                        if (i < a.c) in[3 + i] = __ldg(a.feat + pn * a.c + i) - __ldg(a.feat + pc * a.c + i);
                }
                const float *pw = ctab + a.pre_off, *ps = pw + a.pre_cin * a.pre_cout;
                const int half_out = a.pre_cout / 2;                           // outputs per thread (multiple of 8)
                for (int u = 0; u < half_out / 8; ++u) {
                    float v[8];
                    const int o0 = part * half_out + u * 8;
#pragma unroll
                    for (int o = 0; o < 8; ++o) {
                        float acc = ps[o0 + o];
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            if (i < a.pre_cin) acc = fmaf(in[i], pw[i * a.pre_cout + o0 + o], acc);
                        v[o] = valid ? (a.pre_relu ? fmaxf(acc, 0.f) : acc) : 0.f;
                    }
                    store_units(a1, a2, r, o0 >> 3, v);
                }
                const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                for (int j = a.pre_cout / 8 + part; j < units0; j += 2) store_units(a1, a2, r, j, z);
            } else if (a.mode == TC_SA) {
                const int G = TM / a.k;
                for (int r = wwarp; r < TM; r += 8) {
                    const int g = r / a.k, sidx = r - g * a.k;
                    const long ci = (long)tile * G + g;
                    const bool valid = g < G && ci < a.rows;
                    long pc = 0, pn = 0;
                    if (valid) {
                        const long cloud = ci / a.m;
                        pc = cloud * a.n + __ldg(a.center_idx + ci);
                        pn = cloud * a.n + __ldg(a.nbr_idx + ci * a.nbr_stride + sidx);
                    }
                    for (int j = lane; j < units0; j += 32) {
                        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                        if (valid) {
                            const float4 *fn = reinterpret_cast<const float4 *>(a.feat + pn * a.c) + 2 * j;
                            const float4 *fc = reinterpret_cast<const float4 *>(a.feat + pc * a.c) + 2 * j;
                            const float4 n0 = __ldg(fn), n1 = __ldg(fn + 1), c0 = __ldg(fc), c1 = __ldg(fc + 1);
                            v[0] = n0.x - c0.x; v[1] = n0.y - c0.y; v[2] = n0.z - c0.z; v[3] = n0.w - c0.w;
                            v[4] = n1.x - c1.x; v[5] = n1.y - c1.y; v[6] = n1.z - c1.z; v[7] = n1.w - c1.w;
                        }
                        store_units(a1, a2, r, j, v);
                    }
                }
                {   // xyz_j - xyz_i of the row this thread owns in the epilogue
                    const int g = row / a.k, sidx = row - g * a.k;
                    const long ci = (long)tile * G + g;
                    if (g < G && ci < a.rows) {
                        const long cloud = ci / a.m;
                        const long pc = cloud * a.n + __ldg(a.center_idx + ci);
                        const long pn = cloud * a.n + __ldg(a.nbr_idx + ci * a.nbr_stride + sidx);
#pragma unroll
                        for (int e = 0; e < 3; ++e) xe[e] = __ldg(a.xyz + pn * 3 + e) - __ldg(a.xyz + pc * 3 + e);
                    }
                }
            } else {
                const int ku = a.c_known / 8;
                // two rows per iteration so that twelve 16-byte loads are in flight per lane
                for (int r0 = wwarp * 2; r0 < TM; r0 += 16) {
                    long pp[2]; bool ok[2]; long base[2]; int i0[2], i1[2], i2[2]; float w0[2], w1[2], w2[2];
#pragma unroll
                    for (int t2 = 0; t2 < 2; ++t2) {
                        pp[t2] = (long)tile * TM + r0 + t2;
                        ok[t2] = pp[t2] < a.rows;
                        base[t2] = 0; i0[t2] = i1[t2] = i2[t2] = 0; w0[t2] = w1[t2] = w2[t2] = 0.f;
                        if (ok[t2]) {
                            base[t2] = (pp[t2] / a.n) * a.m;
                            i0[t2] = __ldg(a.idx3 + pp[t2] * 3); i1[t2] = __ldg(a.idx3 + pp[t2] * 3 + 1); i2[t2] = __ldg(a.idx3 + pp[t2] * 3 + 2);
                            w0[t2] = __ldg(a.w3 + pp[t2] * 3); w1[t2] = __ldg(a.w3 + pp[t2] * 3 + 1); w2[t2] = __ldg(a.w3 + pp[t2] * 3 + 2);
                        }
                    }
                    for (int j = lane; j < units0; j += 32) {
                        float v[2][8];
                        if (j < ku) {
                            float4 x0[2], x1[2], y0[2], y1[2], z0[2], z1[2];
#pragma unroll
                            for (int t2 = 0; t2 < 2; ++t2) {
                                const float4 *f0 = reinterpret_cast<const float4 *>(a.known_feat + (base[t2] + i0[t2]) * a.c_known) + 2 * j;
                                const float4 *f1 = reinterpret_cast<const float4 *>(a.known_feat + (base[t2] + i1[t2]) * a.c_known) + 2 * j;
                                const float4 *f2 = reinterpret_cast<const float4 *>(a.known_feat + (base[t2] + i2[t2]) * a.c_known) + 2 * j;
                                x0[t2] = __ldg(f0); x1[t2] = __ldg(f0 + 1); y0[t2] = __ldg(f1); y1[t2] = __ldg(f1 + 1);
                                z0[t2] = __ldg(f2); z1[t2] = __ldg(f2 + 1);
                            }
#pragma unroll
                            for (int t2 = 0; t2 < 2; ++t2) {
                                // interpolation_forward: fma(w2,p2, fma(w0,p0, w1*p1)) (interpolation_cuda_kernel.cu:194)
                                v[t2][0] = __fmaf_rn(w2[t2], z0[t2].x, __fmaf_rn(w0[t2], x0[t2].x, __fmul_rn(w1[t2], y0[t2].x)));
                                v[t2][1] = __fmaf_rn(w2[t2], z0[t2].y, __fmaf_rn(w0[t2], x0[t2].y, __fmul_rn(w1[t2], y0[t2].y)));
                                v[t2][2] = __fmaf_rn(w2[t2], z0[t2].z, __fmaf_rn(w0[t2], x0[t2].z, __fmul_rn(w1[t2], y0[t2].z)));
                                v[t2][3] = __fmaf_rn(w2[t2], z0[t2].w, __fmaf_rn(w0[t2], x0[t2].w, __fmul_rn(w1[t2], y0[t2].w)));
                                v[t2][4] = __fmaf_rn(w2[t2], z1[t2].x, __fmaf_rn(w0[t2], x1[t2].x, __fmul_rn(w1[t2], y1[t2].x)));
                                v[t2][5] = __fmaf_rn(w2[t2], z1[t2].y, __fmaf_rn(w0[t2], x1[t2].y, __fmul_rn(w1[t2], y1[t2].y)));
                                v[t2][6] = __fmaf_rn(w2[t2], z1[t2].z, __fmaf_rn(w0[t2], x1[t2].z, __fmul_rn(w1[t2], y1[t2].z)));
                                v[t2][7] = __fmaf_rn(w2[t2], z1[t2].w, __fmaf_rn(w0[t2], x1[t2].w, __fmul_rn(w1[t2], y1[t2].w)));
                            }
                        } else {
#pragma unroll
                            for (int t2 = 0; t2 < 2; ++t2) {
                                float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
                                if (ok[t2]) {
                                    const float4 *sk = reinterpret_cast<const float4 *>(a.skip_feat + pp[t2] * a.c_skip) + 2 * (j - ku);
                                    s0 = __ldg(sk); s1 = __ldg(sk + 1);
                                }
                                v[t2][0] = s0.x; v[t2][1] = s0.y; v[t2][2] = s0.z; v[t2][3] = s0.w;
                                v[t2][4] = s1.x; v[t2][5] = s1.y; v[t2][6] = s1.z; v[t2][7] = s1.w;
                            }
                        }
#pragma unroll
                        for (int t2 = 0; t2 < 2; ++t2) store_units(a1, a2, r0 + t2, j, v[t2]);
                    }
                }
                if (a.n_extra > 0) {
                    const long p = (long)tile * TM + row;
                    if (p < a.rows)
                        for (int e = 0; e < a.n_extra; ++e) xe[e] = __ldg(a.skip_feat + p * a.c_skip + e);
                }
            }
            fence_proxy_async();
            mbar_arrive(a_ready);

            // ---- per-layer epilogues -----------------------------------------------------------------------------
            for (int l = 0; l < a.n_layers; ++l, ++lcount) {
                const int N = a.N[l];
                const bool last = l == a.n_layers - 1;
                const bool relu = a.relu[l] != 0;
                const bool extras = l == 0 && a.n_extra > 0;
                const float *shl = ctab + a.coff[l];
                mbar_wait(d_ready, lcount & 1);
                tc_fence_after();
                const int npass = last ? (N + 255) / 256 : 1;
                for (int pass = 0; pass < npass; ++pass) {
                    const int ncols = min(256, N - pass * 256);        // columns of this pass
                    const int per = ncols >= 64 ? ncols / 2 : (half == 0 ? ncols : 0);   // columns per worker half (batches of 32)
                    for (int cb = 0; cb < per; cb += 32) {
                        const int col = pass * 256 + (ncols >= 64 ? half * per : 0) + cb;  // first accumulator column of this batch
                        float v[32];
                        tmem_ld32(trow + (uint32_t)col, v);
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const float4 sh = *reinterpret_cast<const float4 *>(shl + col + 4 * u);
                            v[4 * u] += sh.x; v[4 * u + 1] += sh.y; v[4 * u + 2] += sh.z; v[4 * u + 3] += sh.w;
                        }
                        if (extras) {
#pragma unroll
                            for (int e = 0; e < 3; ++e) {
                                if (e < a.n_extra) {
#pragma unroll
                                    for (int u = 0; u < 8; ++u) {
                                        const float4 w = *reinterpret_cast<const float4 *>(wext + e * N + col + 4 * u);
                                        v[4 * u] = fmaf(xe[e], w.x, v[4 * u]); v[4 * u + 1] = fmaf(xe[e], w.y, v[4 * u + 1]);
                                        v[4 * u + 2] = fmaf(xe[e], w.z, v[4 * u + 2]); v[4 * u + 3] = fmaf(xe[e], w.w, v[4 * u + 3]);
                                    }
                                }
                            }
                        }
                        if (relu) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
                        }
                        if (!last) {
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                float w8[8];
#pragma unroll
                                for (int i = 0; i < 8; ++i) w8[i] = v[u * 8 + i];
                                store_units(a1, a2, row, (col >> 3) + u, w8);
                            }
                        } else {
                            const int c0 = col - pass * 256;
#pragma unroll
                            for (int u = 0; u < 8; ++u)
                                *reinterpret_cast<float4 *>(a1 + stage_offset(row, c0 + 4 * u)) =
                                    make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
                        }
                    }
                    if (last) {
                        tc_fence_before();
                        asm volatile("bar.sync 1, %0;" ::"n"(NWORK) : "memory");   // staging complete (workers only)
                        if (a.mode == TC_SA) {
                            const int G = TM / a.k;
                            for (int e = wt; e < G * ncols; e += NWORK) {
                                const int g = e / ncols, c = e - g * ncols;
                                const long ci = (long)tile * G + g;
                                if (ci >= a.rows) continue;
                                float mx = *reinterpret_cast<const float *>(a1 + stage_offset(g * a.k, c));
                                for (int sidx = 1; sidx < a.k; ++sidx)
                                    mx = fmaxf(mx, *reinterpret_cast<const float *>(a1 + stage_offset(g * a.k + sidx, c)));
                                a.out[ci * N + pass * 256 + c] = mx;
                            }
                        } else {
                            const int c4 = ncols / 4;                  // one warp per row: 1 KB contiguous per store wave
                            for (int r = wwarp; r < TM; r += 8) {
                                const long p = (long)tile * TM + r;
                                if (p >= a.rows) break;
                                for (int cq = lane; cq < c4; cq += 32)
                                    *reinterpret_cast<float4 *>(a.out + p * N + pass * 256 + 4 * cq) =
                                        *reinterpret_cast<const float4 *>(a1 + stage_offset(r, 4 * cq));
                            }
                        }
                        asm volatile("bar.sync 1, %0;" ::"n"(NWORK) : "memory");   // staging consumed before it is rewritten
                    }
                }
                if (!last) {
                    tc_fence_before();
                    fence_proxy_async();
                    mbar_arrive(a_ready);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

// ---- host side ---------------------------------------------------------------------------------------------------

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// weight plane (N rows, K bf16 contiguous) -> boxes of (64 k) x (box_n rows), 128-byte swizzle
int make_weight_map(CUtensorMap *map, const void *w, int N, int K, int box_n) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return PAB_EINVAL;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)KCH, (cuuint32_t)box_n};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(w), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : PAB_EINVAL;
}

int g_tc_enabled = 1;

// Shared-memory plan of one launch: operand region, constant table, weight-stage granularity and count.
// `layers` are the TENSOR-CORE layers only (the optional pre-layer is passed separately).
struct TcPlan { int kchunks_max, a_region, nblk, stage_bytes, n_stages, coff[MAX_LAYERS], pre_off; size_t misc, smem; };

bool tc_plan(const pab_layer_t *layers, int n_layers, const pab_layer_t *pre, TcPlan *p) {
    int kmax = 0, ctab = 0;
    bool need64 = false;
    for (int l = 0; l < n_layers; ++l) {
        if (layers[l].tc_k > kmax) kmax = layers[l].tc_k;
        p->coff[l] = ctab;
        ctab += layers[l].c_out;
        if (layers[l].c_out > 64 && layers[l].c_out % 128) need64 = true;   // e.g. 192: only 64-wide blocks divide it
    }
    if (!pre) ctab += (layers[0].c_in - layers[0].tc_k) * layers[0].c_out;
    p->pre_off = ctab;
    if (pre) ctab += pre->c_in * pre->c_out + pre->c_out;
    p->kchunks_max = kmax / KCH;
    p->a_region = 2 * p->kchunks_max * A_CHUNK;
    if (p->a_region < TM * 256 * 4) p->a_region = TM * 256 * 4;      // last-layer fp32 staging [128][256]
    p->misc = 256 + (size_t)ctab * 4 + 64;
    const long budget = 227L * 1024 - p->a_region - (long)p->misc;
    if (budget < 2 * 2 * 64 * 128) return false;
    // 128 output channels per stage when at least 3 such stages fit, else 64
    p->nblk = (!need64 && budget / (2 * 128 * 128) >= 3) ? 128 : 64;
    p->stage_bytes = 2 * p->nblk * 128;
    p->n_stages = (int)(budget / p->stage_bytes);
    if (p->n_stages > MAX_STAGES) p->n_stages = MAX_STAGES;
    p->smem = (size_t)p->a_region + (size_t)p->n_stages * p->stage_bytes + p->misc;
    return p->n_stages >= 2;
}

bool tc_layers_ok(const pab_layer_t *layers, int n_layers, bool first_is_module_input) {
    if (n_layers < 1 || n_layers > MAX_LAYERS) return false;
    for (int l = 0; l < n_layers; ++l) {
        const pab_layer_t &L = layers[l];
        if (!L.w_hi || !L.w_lo || L.tc_k <= 0 || L.tc_k % KCH) return false;
        if (!(L.c_out == 32 || L.c_out % 64 == 0) || L.c_out > 512) return false;
        if (l < n_layers - 1 && L.c_out > 256) return false;
        const bool module_input = l == 0 && first_is_module_input;
        if (!module_input && (L.tc_k0 != 0 || L.tc_k < L.c_in)) return false;          // K may be zero-padded to 64
        if (l > 0 && L.c_in != layers[l - 1].c_out) return false;
    }
    return true;
}

}  // namespace

// 0: not eligible; 1: every layer on tensor cores; 2: first layer evaluated by the loader (tiny input), rest on tensor cores
int pab_tc_eligible(const pab_layer_t *layers, int n_layers, int k_group, int allow_pre) {
    if (!g_tc_enabled || k_group > TM) return 0;
    TcPlan p;
    if (tc_layers_ok(layers, n_layers, true)) {
        const int n_extra = layers[0].c_in - layers[0].tc_k;
        if (n_extra >= 0 && n_extra <= 3 && (n_extra == 0 || layers[0].tc_k0 == 0 || layers[0].tc_k0 == n_extra) &&
            tc_plan(layers, n_layers, nullptr, &p))
            return 1;
    }
    if (allow_pre && n_layers >= 2 && layers[0].c_in <= 8 && layers[0].c_out % 16 == 0 && layers[0].c_out <= 64 &&
        tc_layers_ok(layers + 1, n_layers - 1, false) && tc_plan(layers + 1, n_layers - 1, &layers[0], &p))
        return 2;
    return 0;
}

namespace {

int pab_tc_launch(int mode, long rows, int k_group, const pab_layer_t *all_layers, int n_all, int kind, TcArgs &a, cudaStream_t st) {
    const pab_layer_t *pre = kind == 2 ? &all_layers[0] : nullptr;
    const pab_layer_t *layers = kind == 2 ? all_layers + 1 : all_layers;
    const int n_layers = kind == 2 ? n_all - 1 : n_all;
    TcPlan p;
    if (!tc_plan(layers, n_layers, pre, &p)) return PAB_EINVAL;
    a.kchunks_max = p.kchunks_max; a.a_region = p.a_region; a.nblk = p.nblk; a.stage_bytes = p.stage_bytes; a.n_stages = p.n_stages;
    for (int l = 0; l < n_layers; ++l) {
        const pab_layer_t &L = layers[l];
        const int nbr = L.c_out < a.nblk ? L.c_out : a.nblk;
        if (make_weight_map(&a.tm[l][0], L.w_hi, L.c_out, L.tc_k, nbr)) return PAB_EINVAL;
        if (make_weight_map(&a.tm[l][1], L.w_lo, L.c_out, L.tc_k, nbr)) return PAB_EINVAL;
        a.shift[l] = L.shift; a.K[l] = L.tc_k; a.N[l] = L.c_out; a.relu[l] = L.relu; a.coff[l] = p.coff[l];
    }
    a.n_layers = n_layers; a.mode = mode; a.rows = rows;
    if (pre) {
        a.n_extra = 0; a.w_extra = nullptr;
        a.pre_cin = pre->c_in; a.pre_cout = pre->c_out; a.pre_relu = pre->relu; a.pre_wt = pre->wt; a.pre_shift = pre->shift;
        a.pre_off = p.pre_off;
    } else {
        a.pre_cout = 0;
        a.n_extra = layers[0].c_in - layers[0].tc_k;
        // extra channels sit before (SA: xyz first) or after (FP: skip last) the tensor-core part
        const int extra_row0 = layers[0].tc_k0 == 0 ? layers[0].tc_k : 0;
        a.w_extra = layers[0].wt + (size_t)extra_row0 * layers[0].c_out;
    }
    const long per_tile = mode == TC_SA ? (TM / k_group) : TM;
    a.ntiles = (int)((rows + per_tile - 1) / per_tile);
    static int n_sm = 0;
    if (!n_sm) {
        int dev = 0;
        PAB_CUDA(cudaGetDevice(&dev));
        PAB_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    }
    PAB_CUDA(cudaFuncSetAttribute(mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
    const int grid = a.ntiles < n_sm ? a.ntiles : n_sm;
    if (grid == 0) return 0;
    mlp_tc_kernel<<<grid, TC_THREADS, p.smem, st>>>(a);
    PAB_LAUNCH_CHECK();
    return 0;
}

}  // namespace

PAB_API void pab_tune_tensor_core(int enable) { g_tc_enabled = enable; }

int pab_tc_sa(int kind, int b, int n, int m, int k, int nbr_stride, int c, const float *xyz, const float *feat, const int *center_idx,
              const int *nbr_idx, const pab_layer_t *layers, int n_layers, float *out, cudaStream_t st) {
    TcArgs a{};
    a.n = n; a.m = m; a.k = k; a.nbr_stride = nbr_stride; a.c = c; a.xyz = xyz; a.feat = feat; a.center_idx = center_idx;
    a.nbr_idx = nbr_idx; a.out = out;
    return pab_tc_launch(TC_SA, (long)b * m, k, layers, n_layers, kind, a, st);
}

int pab_tc_fp(int b, int n, int m, int c_known, int c_skip, const float *known_feat, const float *skip_feat, const int *idx,
              const float *weight, const pab_layer_t *layers, int n_layers, float *out, cudaStream_t st) {
    TcArgs a{};
    a.n = n; a.m = m; a.c_known = c_known; a.c_skip = c_skip; a.known_feat = known_feat; a.skip_feat = skip_feat;
    a.idx3 = idx; a.w3 = weight; a.out = out;
    return pab_tc_launch(TC_FP, (long)b * n, 0, layers, n_layers, 1, a, st);
}
