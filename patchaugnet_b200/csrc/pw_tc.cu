// pw_tc.cu — one point-wise layer out = act(x W^T + shift) (+ residual) on the tcgen05 tensor cores, sm_100a: the q / v / trans_conv
// projections of PPT-Net's SA_Layer (pptnet.py:261-282; rows = clouds x points, C in {64, 128, 256, 512}).
//
// The fp32 tile kernel of mlp.cu spends 0.15 ms per SA_Layer on these three products at level 0 and most of the layer's time at
// the deep levels (64 x 256 and 16 x 512 per cloud: tiny point counts, 1-MB weights).  Same contract as the other tcgen05 kernels:
// rows and weights as bf16 hi/lo planes, hi*hi + lo*hi + hi*lo per product, fp32 accumulation in tensor memory.
//
// One CTA per (128-row tile, slice of <= 128 output channels).  K runs through shared memory in 64-channel chunks, two stages: the
// chunk's rows are read as fp32, split and stored into the swizzled K-major layout by the CTA's threads, the chunk of the weight
// planes arrives by cp.async, both while the MMAs of the previous chunk run; the accumulator stays in TMEM across the chunks.
// Epilogue: thread = row, shift, ReLU, residual, 16-byte stores.
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int PW_THREADS = 256;
constexpr int PW_NT = 128;                    // output channels per CTA
constexpr int PW_BSTAGE = PW_NT * 128;        // bytes of one weight plane of one chunk

__device__ __forceinline__ void pw_cp_async16(void *dst_smem, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}

struct PwArgs {
    long rows;
    int K, N, wk, relu;                        // wk: bf16 per weight row in global memory (tc_k)
    const float *x;                            // (rows, K)
    const uint16_t *w_hi, *w_lo;               // (N, wk)
    const float *shift, *residual;             // (N); (rows, N) or NULL
    float *out;
    long out_ld;
};

__global__ void __launch_bounds__(PW_THREADS, 1) pw_tc_kernel(const __grid_constant__ PwArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *ast = smem;                                                       // [stage][plane][128 rows][128 B]
    uint8_t *bst = smem + 4 * (size_t)A_CHUNK;                                 // [stage][plane][NT rows][128 B]
    uint64_t *done = reinterpret_cast<uint64_t *>(bst + 4 * (size_t)PW_BSTAGE);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(done + 2);

    const int tid = threadIdx.x, warp = uniform_warp_idx();
    const long r0 = (long)blockIdx.x * TM;
    const int n0 = blockIdx.y * PW_NT, nt = min(PW_NT, a.N - n0);
    const int nchunks = a.K / KCH;
    if (tid == 0) {
        mbar_init(done, 1);
        mbar_init(done + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    auto stage_b = [&](int ch, int s) {
        for (int e = tid; e < nt * 8; e += PW_THREADS) {
            const int o = e >> 3, j = e & 7;
            const uint32_t off = (uint32_t)(o * 128 + ((j ^ (o & 7)) << 4));
            const size_t src = (size_t)(n0 + o) * a.wk + (size_t)ch * KCH + 8 * j;
            pw_cp_async16(bst + (size_t)(2 * s) * PW_BSTAGE + off, a.w_hi + src);
            pw_cp_async16(bst + (size_t)(2 * s + 1) * PW_BSTAGE + off, a.w_lo + src);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // A: 128 rows x 8 units of a chunk = 1024 (row, unit) pairs, four per thread; a warp covers four whole rows per pass (256
    // contiguous bytes each): loads issued together, stored after the caller's wait
    constexpr int AR = TM * 8 / PW_THREADS;
    float4 ax[AR][2];
    auto load_a = [&](int ch) {
#pragma unroll
        for (int rr = 0; rr < AR; ++rr) {
            const int e = tid + rr * PW_THREADS, row = e >> 3, j = e & 7;
            ax[rr][0] = ax[rr][1] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r0 + row < a.rows) {
                const float4 *p = reinterpret_cast<const float4 *>(a.x + (r0 + row) * a.K + (long)ch * KCH + 8 * j);
                ax[rr][0] = __ldg(p); ax[rr][1] = __ldg(p + 1);
            }
        }
    };
    auto store_a = [&](int s) {
        uint8_t *a1 = ast + (size_t)(2 * s) * A_CHUNK, *a2 = a1 + A_CHUNK;
#pragma unroll
        for (int rr = 0; rr < AR; ++rr) {
            const int e = tid + rr * PW_THREADS, row = e >> 3, j = e & 7;
            const float y[8] = {ax[rr][0].x, ax[rr][0].y, ax[rr][0].z, ax[rr][0].w, ax[rr][1].x, ax[rr][1].y, ax[rr][1].z, ax[rr][1].w};
            store_units(a1, a2, row, j, y);
        }
    };

    stage_b(0, 0);
    load_a(0);
    store_a(0);
    const uint32_t leader = elect_one();
    const uint32_t idesc = umma_idesc(nt);
    const uint32_t ad = umma_desc_lo(smem_u32(ast)), bd = umma_desc_lo(smem_u32(bst));
    for (int it = 0; it < nchunks; ++it) {
        const int s = it & 1;
        const bool more = it + 1 < nchunks;
        if (more) {
            if (it >= 1) mbar_wait(done + (s ^ 1), ((it - 1) >> 1) & 1);       // the chunk that used the other stage has been multiplied
            stage_b(it + 1, s ^ 1);
            load_a(it + 1);
        }
        if (more) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (more) store_a(s ^ 1);
        fence_proxy_async();
        __syncthreads();
        if (warp == 0) {
            tc_fence_after();
            const uint32_t a1d = ad + (uint32_t)(2 * s) * (A_CHUNK >> 4), a2d = a1d + (A_CHUNK >> 4);
            const uint32_t b1d = bd + (uint32_t)(2 * s) * (PW_BSTAGE >> 4), b2d = b1d + (PW_BSTAGE >> 4);
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
        const int last = nchunks - 1;
        mbar_wait(done + (last & 1), (last >> 1) & 1);
        tc_fence_after();
        const long row = r0 + tid;
        const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
        const long old = a.out_ld ? a.out_ld : a.N;
        for (int cb = 0; cb < nt; cb += 32) {
            float v[32];
            tmem_ld32(trow + (uint32_t)cb, v);                                 // warp-convergent; only the memory accesses are predicated
            if (row < a.rows) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const float4 sh = __ldg(reinterpret_cast<const float4 *>(a.shift + n0 + cb) + u);
                    float4 o = make_float4(v[4 * u] + sh.x, v[4 * u + 1] + sh.y, v[4 * u + 2] + sh.z, v[4 * u + 3] + sh.w);
                    if (a.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                    if (a.residual) {
                        const float4 rs = __ldg(reinterpret_cast<const float4 *>(a.residual + row * a.N + n0 + cb) + u);
                        o.x += rs.x; o.y += rs.y; o.z += rs.z; o.w += rs.w;
                    }
                    *reinterpret_cast<float4 *>(a.out + row * old + n0 + cb + 4 * u) = o;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
    }
}

int g_pw_tc = 1;

}  // namespace

PAB_API void pab_tune_pointwise_tc(int on) { g_pw_tc = on; }

// 1 when the layer can run on pw_tc_kernel: hi/lo planes covering every input, K a multiple of 64, N a multiple of 32, 16-byte
// aligned rows of out / residual
int pab_pw_tc_eligible(const pab_layer_t *L, long out_ld) {
    if (!g_pw_tc || !L->w_hi || !L->w_lo || L->tc_k0 != 0 || L->tc_k < L->c_in || L->tc_k % 8) return 0;
    if (L->c_in % KCH || L->c_in < KCH || L->c_out % 32 || L->c_out < 32) return 0;
    if (out_ld % 4) return 0;
    return 1;
}

int pab_pw_tc_launch(long rows, const float *x, const pab_layer_t *L, const float *residual, float *out, long out_ld, cudaStream_t st) {
    if (rows == 0) return 0;
    PwArgs a;
    a.rows = rows; a.K = L->c_in; a.N = L->c_out; a.wk = L->tc_k; a.relu = L->relu;
    a.x = x; a.w_hi = (const uint16_t *)L->w_hi; a.w_lo = (const uint16_t *)L->w_lo; a.shift = L->shift; a.residual = residual;
    a.out = out; a.out_ld = out_ld;
    const size_t smem = 4 * (size_t)A_CHUNK + 4 * (size_t)PW_BSTAGE + 64;
    static bool configured = false;
    if (!configured) {
        PAB_CUDA(cudaFuncSetAttribute(pw_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const long tiles = (rows + TM - 1) / TM;
    if (tiles > 0x7fffffffL) return PAB_EINVAL;
    pw_tc_kernel<<<dim3((unsigned)tiles, (a.N + PW_NT - 1) / PW_NT), PW_THREADS, smem, st>>>(a);
    PAB_LAUNCH_CHECK();
    return 0;
}
