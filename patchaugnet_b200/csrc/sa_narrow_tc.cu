// sa_narrow_tc.cu — the first set-abstraction module (tiny input, narrow layers: PatchAugNet / PPT-Net SA0 = [xyz_rel ; feat_rel]
// (6) -> 32 -> 32 -> 64, max over k neighbours) as a SMALL-CTA tensor-core kernel, sm_100a.
//
// mlp_tc.cu runs this module with one 512-thread CTA per SM and warp-specialised roles; for layers this narrow each 128-row tile
// is a chain of barrier hand-offs (loader -> MMA -> epilogue -> MMA -> epilogue -> staged max-pool) whose latencies add up to
// ~5.8 k cycles while the MMAs themselves take < 300 — the tensor pipe is 4 % busy (profiles/r02_share_of_step.md).  Here the
// chain stays serial inside a CTA, but the CTA is small enough that THREE are resident per SM and the hardware interleaves
// their chains, and the chain itself is shorter:
//
//   * 128 threads = 128 rows = 128 TMEM lanes.  A step's rows are (32 centres) x (4 neighbour slots), so a thread keeps the
//     SAME centre for all ceil(k / 4) steps of an item: the max over the k neighbours is a running max of the last layer's raw
//     accumulators in the thread's registers — no per-step staging, no per-step CTA barriers for the pooling; shift and ReLU are
//     monotonic and are applied once after the max (bit-identical to max of relu(x + shift));
//   * TMEM: 128 columns per CTA (accumulator 64 + bf16 hi plane 32 + lo plane 32);
//   * the folded weights (bf16 hi/lo planes, 24 KB) are copied ONCE per CTA into shared memory in the canonical K-major
//     128-byte-swizzled layout and stay there: no weight streaming, no per-tile barrier round trips with a producer;
//   * operands never touch shared memory: the pre-layer (<= 8 inputs) is evaluated in fp32 by the row's thread and written to the
//     TMEM planes (tcgen05.st), the layers read A from TMEM (TS form) and their epilogues write the next operand back there;
//   * at the end of an item the four slots of each centre are merged through a [128][32] fp32 staging tile, coalesced stores;
//   * items come from a global counter (self-resetting), so CTAs that start late — SMs busy with another stream's kernel —
//     simply take fewer; the gathers of step t+1 (rows) and t+2 (indices) are in flight while step t is evaluated.
//
// Arithmetic identical to mlp_tc.cu's pre-layer mode (same fp32 pre-layer order, same hi/lo split, same MMA order per k-chunk):
// the results are bit-identical to that kernel (tests/test_mlp_tc_gpu.py).
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int SN_THREADS = 128;
constexpr int SN_MAXL = 3;                    // tensor-core layers after the pre-layer
constexpr uint32_t SN_D = 0, SN_AH = 64, SN_AL = 96;   // TMEM columns
constexpr int SN_TMEM_COLS = 128;
constexpr int SN_PITCH = 36;                  // floats per staged row (32 + 4: 16-byte row writes and column reads conflict-free)
constexpr int SN_STG_BYTES = TM * SN_PITCH * 4;
constexpr int SN_CTAB_BYTES = 3328;             // pre-layer 8 x 64 + 64, three shifts of 64, rounded up

struct alignas(16) SnArgs {
    const uint16_t *w_hi[SN_MAXL], *w_lo[SN_MAXL];
    const float *shift[SN_MAXL];
    int wk[SN_MAXL];                           // bf16 per weight row in global memory (tc_k)
    int N[SN_MAXL], K[SN_MAXL], relu[SN_MAXL];
    int woff[SN_MAXL];                         // byte offset of the layer's hi plane in shared memory (lo plane follows)
    int soff[SN_MAXL];                         // float offset of the layer's shift in the constant table
    int n_layers, planes, wbytes;
    int pre_cin, pre_cout, pre_relu;
    const float *pre_wt, *pre_shift;
    int n, m, k, nbr_stride, c;
    const float *xyz, *feat;
    const int *center_idx, *nbr_idx;
    float *out;
    long rows;
    int nitems;                                // items of C = 128 >> slog centres; steps = ceil(k / S) tensor-core steps each
    int slog, steps;                           // S = 1 << slog neighbour slots per step
    unsigned int *counter;                     // [0] next tile, [1] CTAs finished (the last one zeroes both)
    int dbg;                                   // debugging: 1 = no gathers (zero rows)
    long long *trace;                          // optional timeline of CTA 0's first 16 steps (pab_tune_sa_narrow_trace), [step][8] clock64
};

__device__ __forceinline__ void sn_umma_ts(uint32_t issue, uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n"
        ".reg .pred p, q;\n"
        ".reg .b64 db;\n"
        "setp.ne.b32 p, %5, 0;\n"
        "setp.ne.b32 q, %6, 0;\n"
        "mov.b64 db, {%2, %3};\n"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n"
        "}\n" ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(UMMA_DESC_HI), "r"(idesc), "r"(acc), "r"(issue) : "memory");
}
__device__ __forceinline__ void sn_tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
          "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}

#define SN_TRACE(ev)                                                                      \
    do {                                                                                  \
        if (a.trace && blockIdx.x == 0 && tid == 0 && nstep < 16) a.trace[nstep * 8 + (ev)] = clock64(); \
    } while (0)

// CIN: pre-layer inputs evaluated per row (the module's count rounded up to 3 / 6 / 8, surplus weight rows are zero);
// SLOG: log2 of the neighbour slots per step.  Compile-time so that the pre-layer is one branch-free block of multiply-adds whose
// weight loads the compiler batches (with a run-time input count every input sat behind its own branch: 3 k cycles per step).
template <int CIN, int SLOG>
__global__ void __launch_bounds__(SN_THREADS, 3) sa_narrow_tc_kernel(const __grid_constant__ SnArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *wsm = smem;                                           // weight planes, each N x 128 bytes, 1024-aligned
    float *stg = reinterpret_cast<float *>(smem + a.wbytes);
    float *ctab = stg + TM * SN_PITCH;                             // [pre weights (cin x cout) | pre shift | layer shifts]
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + a.wbytes + SN_STG_BYTES + SN_CTAB_BYTES);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar + 1);
    int *item_ring = reinterpret_cast<int *>(bar + 2);             // [4] items drawn from the counter (-1: none left)

    const int tid = threadIdx.x, warp = uniform_warp_idx();
    const int pre_n = CIN * a.pre_cout;

    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // the first item is static (grid <= items), the counter hands out the items from gridDim.x on
        const unsigned v = atomicAdd(a.counter, 2u) + gridDim.x;
        item_ring[0] = (int)blockIdx.x;
        item_ring[1] = v < (unsigned)a.nitems ? (int)v : -1;
        item_ring[2] = v + 1 < (unsigned)a.nitems ? (int)(v + 1) : -1;
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(SN_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // Row r of a step = (centre r % C of the item, neighbour step * S + r / C): a thread keeps ONE centre for the whole item, so
    // the max over the k neighbours is a running max in its registers and nothing is staged per step.
    constexpr int C = TM >> SLOG, S = 1 << SLOG;
    const int cc = tid & (C - 1), slot = tid >> (7 - SLOG);
    // Gather pipeline, two steps deep.  Each stage only ISSUES loads and keeps the raw values; the arithmetic on them (point
    // index = cloud base + loaded index, input = neighbour - centre) is done one step later, so no load latency is waited for
    // inside a step (a subtraction written next to its loads stalls the warp there for a memory round trip).
    struct Idx { long base; int c, n; bool valid; };            // stage 2: centre / neighbour indices of the step two ahead
    struct Raw { float nx[3], cx[3], nf[5], cf[5]; bool valid; };   // stage 1: xyz / feature values of the next step
    auto issue_idx = [&](int item, int step, Idx &o) {
        const long ci = (long)item * C + cc;
        o.valid = item >= 0 && ci < a.rows && !(a.dbg & 1);
        o.base = 0; o.c = 0; o.n = 0;
        if (o.valid) {
            o.base = (long)((unsigned)ci / (unsigned)a.m) * a.n;   // rows < 2^31 (checked by the launcher): 32-bit division
            o.c = __ldg(a.center_idx + ci);
            o.n = __ldg(a.nbr_idx + ci * a.nbr_stride + min(step * S + slot, a.k - 1));   // surplus slots repeat the last neighbour
        }
    };
    auto issue_rows = [&](const Idx &ix, Raw &o) {
        o.valid = ix.valid;
#pragma unroll
        for (int i = 0; i < 3; ++i) o.nx[i] = o.cx[i] = 0.f;
#pragma unroll
        for (int i = 0; i < 5; ++i) o.nf[i] = o.cf[i] = 0.f;
        if (ix.valid) {
            const long pc = ix.base + ix.c, pn = ix.base + ix.n;
#pragma unroll
            for (int i = 0; i < 3; ++i) { o.nx[i] = __ldg(a.xyz + pn * 3 + i); o.cx[i] = __ldg(a.xyz + pc * 3 + i); }
#pragma unroll
            for (int i = 0; i < 5; ++i)
                if (i < a.c) { o.nf[i] = __ldg(a.feat + pn * a.c + i); o.cf[i] = __ldg(a.feat + pc * a.c + i); }
        }
    };
    // cursors over the CTA's (item, step) sequence: cur and the step two ahead of it
    uint32_t j0 = 0, j2 = 0;                                       // ring positions
    int st0 = 0, st2 = 0;
    auto advance = [&](uint32_t &j, int &st) {
        if (++st == a.steps) { st = 0; ++j; }
    };

    Idx idx2;
    issue_idx((int)blockIdx.x, 0, idx2);                           // the first item's indices are in flight during the set-up

    // constants and weights, once per CTA
    for (int i = tid; i < pre_n; i += SN_THREADS) ctab[i] = i < a.pre_cin * a.pre_cout ? __ldg(a.pre_wt + i) : 0.f;
    for (int i = tid; i < a.pre_cout; i += SN_THREADS) ctab[pre_n + i] = __ldg(a.pre_shift + i);
    for (int l = 0; l < a.n_layers; ++l) {
        for (int i = tid; i < a.N[l]; i += SN_THREADS) ctab[a.soff[l] + i] = __ldg(a.shift[l] + i);
        // weight row o (N rows), 16-byte unit j (8 of them = the 64-channel chunk that carries K <= 64) -> o*128 + ((j ^ (o&7)) << 4)
        for (int pl = 0; pl < a.planes; ++pl) {
            const uint16_t *src = pl == 0 ? a.w_hi[l] : a.w_lo[l];
            uint8_t *dst = wsm + a.woff[l] + pl * a.N[l] * 128;
            for (int e = tid; e < a.N[l] * 8; e += SN_THREADS) {
                const int o = e >> 3, j = e & 7;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst + o * 128 + ((j ^ (o & 7)) << 4))),
                             "l"(reinterpret_cast<const uint4 *>(src + (size_t)o * a.wk[l]) + j) : "memory");
            }
        }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");                // every weight copy of this thread has landed
    fence_proxy_async();                                           // the weights are read by the tensor cores (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    const uint32_t leader = elect_one();
    const uint32_t w_lo32 = umma_desc_lo(smem_u32(wsm));

    Raw raw1;
    int item = (int)blockIdx.x;
    issue_rows(idx2, raw1);
    advance(j2, st2);
    issue_idx(item_ring[j2 & 3], st2, idx2);
    uint32_t ph = 0;
    const float *pw = ctab, *ps = ctab + pre_n;
    float m0[32], m1[32];                                          // running max of the last layer's RAW accumulators (shift and
                                                                   // ReLU are monotonic: applied once, after the max)
    int nstep = 0;
    unsigned drawn = 0;
    while (item >= 0) {
        SN_TRACE(0);
        if (st0 == 0) {
            // draw the item three ahead.  Its ring entry is first read when the two-ahead cursor enters it — with more than one
            // step per item at least one whole step (several CTA barriers) after the END of this step, where the value is
            // stored, so the atomic's round trip is off the chain; with one step per item it is stored at once
            if (tid == 0) {
                drawn = atomicAdd(a.counter, 1u) + gridDim.x;
                if (a.steps == 1) item_ring[(j0 + 3) & 3] = drawn < (unsigned)a.nitems ? (int)drawn : -1;
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) m0[i] = m1[i] = -INFINITY;
        }
        float in[8];
#pragma unroll
        for (int i = 0; i < 3; ++i) in[i] = raw1.nx[i] - raw1.cx[i];
#pragma unroll
        for (int i = 0; i < 5; ++i) in[3 + i] = raw1.nf[i] - raw1.cf[i];
        const bool valid = raw1.valid;
        issue_rows(idx2, raw1);                                    // rows of the next step (their indices were loaded a step ago)
        advance(j2, st2);
        issue_idx(item_ring[j2 & 3], st2, idx2);                   // indices of the step after it

        // ---- pre-layer in fp32 -> bf16 hi/lo -> TMEM planes ------------------------------------------------------------
        for (int cb = 0; cb < a.pre_cout; cb += 32) {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int u = 0; u < 8; ++u) {                          // four outputs at a time, weights as broadcast 16-byte loads
                float4 acc = *reinterpret_cast<const float4 *>(ps + cb + 4 * u);
#pragma unroll
                for (int i = 0; i < CIN; ++i) {
                    const float4 w = *reinterpret_cast<const float4 *>(pw + i * a.pre_cout + cb + 4 * u);
                    acc.x = fmaf(in[i], w.x, acc.x); acc.y = fmaf(in[i], w.y, acc.y);
                    acc.z = fmaf(in[i], w.z, acc.z); acc.w = fmaf(in[i], w.w, acc.w);
                }
                if (a.pre_relu) { acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f); }
                if (!valid) acc = make_float4(0.f, 0.f, 0.f, 0.f);
                split_pack(acc.x, acc.y, hi[2 * u], lo[2 * u]);
                split_pack(acc.z, acc.w, hi[2 * u + 1], lo[2 * u + 1]);
            }
            sn_tmem_st16(trow + SN_AH + (uint32_t)(cb >> 1), hi);
            if (a.planes == 2) sn_tmem_st16(trow + SN_AL + (uint32_t)(cb >> 1), lo);
        }
        SN_TRACE(1);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncthreads();
        SN_TRACE(2);

        for (int l = 0; l < a.n_layers; ++l) {
            const int N = a.N[l];
            if (warp == 0) {
                // ---- MMAs of the layer: per 64-channel chunk (K <= 64: one) hi*hi + lo*hi over the k-steps, then hi*lo ---------
                tc_fence_after();
                const uint32_t idesc = umma_idesc(N);
                const uint32_t bh = w_lo32 + ((uint32_t)a.woff[l] >> 4), bl = bh + (uint32_t)N * 8;
                const int kn = a.K[l] >> 4;
                if (a.planes == 2 && kn == 2) {                    // the common case (32 input channels, hi/lo) as straight-line code
                    sn_umma_ts(leader, tmem + SN_D, tmem + SN_AH, bh, idesc, 0);
                    sn_umma_ts(leader, tmem + SN_D, tmem + SN_AL, bh, idesc, 1);
                    sn_umma_ts(leader, tmem + SN_D, tmem + SN_AH + 8, bh + 2, idesc, 1);
                    sn_umma_ts(leader, tmem + SN_D, tmem + SN_AL + 8, bh + 2, idesc, 1);
                    sn_umma_ts(leader, tmem + SN_D, tmem + SN_AH, bl, idesc, 1);
                    sn_umma_ts(leader, tmem + SN_D, tmem + SN_AH + 8, bl + 2, idesc, 1);
                } else {
                    for (int ks = 0; ks < kn; ++ks) {
                        sn_umma_ts(leader, tmem + SN_D, tmem + SN_AH + ks * 8, bh + 2 * ks, idesc, ks != 0);
                        if (a.planes == 2) sn_umma_ts(leader, tmem + SN_D, tmem + SN_AL + ks * 8, bh + 2 * ks, idesc, 1);
                    }
                    if (a.planes == 2)
                        for (int ks = 0; ks < kn; ++ks) sn_umma_ts(leader, tmem + SN_D, tmem + SN_AH + ks * 8, bl + 2 * ks, idesc, 1);
                }
                umma_commit_if(leader, bar);
                __syncwarp();
                SN_TRACE(3 + 2 * (l != 0));
            }
            mbar_wait(bar, ph);
            ph ^= 1;
            tc_fence_after();
            SN_TRACE(4 + 2 * (l != 0));
            if (l < a.n_layers - 1) {
                const float *sh = ctab + a.soff[l];
                for (int cb = 0; cb < N; cb += 32) {
                    float v[32];
                    tmem_ld32(trow + SN_D + (uint32_t)cb, v);
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const float4 s4 = *reinterpret_cast<const float4 *>(sh + cb + 4 * u);
                        v[4 * u] += s4.x; v[4 * u + 1] += s4.y; v[4 * u + 2] += s4.z; v[4 * u + 3] += s4.w;
                    }
                    if (a.relu[l]) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
                    }
                    uint32_t hi[16], lo[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) split_pack(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
                    sn_tmem_st16(trow + SN_AH + (uint32_t)(cb >> 1), hi);
                    if (a.planes == 2) sn_tmem_st16(trow + SN_AL + (uint32_t)(cb >> 1), lo);
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncthreads();
            } else {
                float v[32];
                tmem_ld32(trow + SN_D, v);
#pragma unroll
                for (int i = 0; i < 32; ++i) m0[i] = fmaxf(m0[i], v[i]);
                if (N > 32) {
                    tmem_ld32(trow + SN_D + 32u, v);
#pragma unroll
                    for (int i = 0; i < 32; ++i) m1[i] = fmaxf(m1[i], v[i]);
                }
            }
        }
        tc_fence_before();                                         // this step's accumulator reads precede the next step's MMAs
        if (st0 == 0 && tid == 0 && a.steps > 1) item_ring[(j0 + 3) & 3] = drawn < (unsigned)a.nitems ? (int)drawn : -1;
        SN_TRACE(7);
        ++nstep;

        if (st0 == a.steps - 1) {
            // ---- end of the item: max over the S slots of each centre, shift, ReLU, coalesced stores ------------------------
            const int l = a.n_layers - 1, N = a.N[l];
            const float *sh = ctab + a.soff[l];
            for (int h = 0; h < N; h += 32) {
                float *srow = stg + tid * SN_PITCH;
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    *reinterpret_cast<float4 *>(srow + 4 * u) = h == 0 ? make_float4(m0[4 * u], m0[4 * u + 1], m0[4 * u + 2], m0[4 * u + 3])
                                                                       : make_float4(m1[4 * u], m1[4 * u + 1], m1[4 * u + 2], m1[4 * u + 3]);
                __syncthreads();
                const int col = tid & 31;
                const float shc = sh[h + col];
                const bool relu = a.relu[l] != 0;
                float *orow = a.out + ((long)item * C) * N + h + col;
#pragma unroll
                for (int i = 0; i < C / 4; ++i) {                  // warp w takes centres w, w + 4, ...: 128-byte rows out
                    const int c2 = (tid >> 5) + 4 * i;
                    float mx = stg[c2 * SN_PITCH + col];
#pragma unroll
                    for (int j = 1; j < S; ++j) mx = fmaxf(mx, stg[(j * C + c2) * SN_PITCH + col]);
                    mx += shc;
                    if ((long)item * C + c2 < a.rows) orow[(long)c2 * N] = relu ? fmaxf(mx, 0.f) : mx;
                }
                __syncthreads();
            }
        }
        advance(j0, st0);
        item = item_ring[j0 & 3];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(SN_TMEM_COLS) : "memory");
    }
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(a.counter + 1, 1u) == gridDim.x - 1) {       // every CTA has drawn its last item: rearm for the next launch
            a.counter[0] = 0;
            a.counter[1] = 0;
            __threadfence();
        }
    }
}

int g_sn_carveout = 1;

template <int CIN, int SLOG>
int sn_launch(const SnArgs &a, long grid, size_t smem, cudaStream_t st) {
    static size_t configured = 0;                                  // per instantiation: the attributes are set when the need grows
    static int carve = -2;
    if (smem > configured || carve != g_sn_carveout) {
        PAB_CUDA(cudaFuncSetAttribute(sa_narrow_tc_kernel<CIN, SLOG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PAB_CUDA(cudaFuncSetAttribute(sa_narrow_tc_kernel<CIN, SLOG>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                      g_sn_carveout ? (int)cudaSharedmemCarveoutMaxShared : (int)cudaSharedmemCarveoutDefault));
        configured = smem;
        carve = g_sn_carveout;
    }
    sa_narrow_tc_kernel<CIN, SLOG><<<(unsigned)grid, SN_THREADS, smem, st>>>(a);
    PAB_LAUNCH_CHECK();
    return 0;
}

int g_sn_enabled = 1;
int g_sn_ctas_per_sm = 0;     // 0 = automatic: 3 when the launch has the GPU to itself, 2 when SMs are shared with another stream
long long *g_sn_trace = nullptr;
int g_sn_dbg = 0;

}  // namespace

extern int g_tc_max_ctas;
unsigned int *pab_tile_counter_pair(cudaStream_t st);

PAB_API void pab_tune_sa_narrow_trace(void *device_buffer) { g_sn_trace = (long long *)device_buffer; }
PAB_API void pab_tune_sa_narrow_dbg(int flags) { g_sn_dbg = flags; }

PAB_API void pab_tune_sa_narrow(int enable, int ctas_per_sm) {
    g_sn_enabled = enable & 1;
    g_sn_carveout = (enable & 2) == 0;
    if (ctas_per_sm >= 0 && ctas_per_sm <= 3) g_sn_ctas_per_sm = ctas_per_sm;
}

// pre-layer mode modules whose tensor-core layers are all <= 64 wide: 1 if this kernel takes the module
int pab_sa_narrow_eligible(const pab_layer_t *layers, int n_layers, int k) {
    if (!g_sn_enabled || n_layers < 2 || n_layers > 1 + SN_MAXL || k < 1 || k > TM) return 0;
    const pab_layer_t &pre = layers[0];
    if (pre.c_in > 8 || !(pre.c_out == 32 || pre.c_out == 64)) return 0;
    for (int l = 1; l < n_layers; ++l) {
        const pab_layer_t &L = layers[l];
        if (!L.w_hi || L.tc_k0 != 0 || L.tc_k < 64 || L.tc_k % 64) return 0;
        if ((L.w_lo != nullptr) != (layers[1].w_lo != nullptr)) return 0;
        if (!(L.c_out == 32 || L.c_out == 64) || L.c_in != layers[l - 1].c_out) return 0;
    }
    return 1;
}

int pab_sa_narrow_launch(int b, int n, int m, int k, int nbr_stride, int c, const float *xyz, const float *feat, const int *center_idx,
                         const int *nbr_idx, const pab_layer_t *layers, int n_layers, float *out, cudaStream_t st) {
    SnArgs a{};
    const pab_layer_t &pre = layers[0];
    a.n_layers = n_layers - 1;
    a.planes = layers[1].w_lo ? 2 : 1;
    const int cin_t = pre.c_in <= 3 ? 3 : (pre.c_in <= 6 ? 6 : 8);     // the kernel's compile-time input count
    int woff = 0, soff = cin_t * pre.c_out + pre.c_out;
    for (int l = 0; l < a.n_layers; ++l) {
        const pab_layer_t &L = layers[l + 1];
        a.w_hi[l] = (const uint16_t *)L.w_hi; a.w_lo[l] = (const uint16_t *)L.w_lo; a.shift[l] = L.shift;
        a.wk[l] = L.tc_k; a.N[l] = L.c_out; a.K[l] = L.c_in; a.relu[l] = L.relu;
        a.woff[l] = woff; a.soff[l] = soff;
        woff += a.planes * L.c_out * 128;
        soff += L.c_out;
    }
    if (soff * 4 > SN_CTAB_BYTES) return PAB_EINVAL;
    a.wbytes = woff;
    a.pre_cin = pre.c_in; a.pre_cout = pre.c_out; a.pre_relu = pre.relu; a.pre_wt = pre.wt; a.pre_shift = pre.shift;
    a.n = n; a.m = m; a.k = k; a.nbr_stride = nbr_stride; a.c = c; a.xyz = xyz; a.feat = feat; a.center_idx = center_idx;
    a.nbr_idx = nbr_idx; a.out = out; a.trace = g_sn_trace; a.dbg = g_sn_dbg;
    a.rows = (long)b * m;
    if (a.rows * k >= (1L << 31)) return PAB_EINVAL;
    if (a.rows == 0) return 0;
    a.slog = k >= 4 ? 2 : (k >= 2 ? 1 : 0);                        // 4 neighbour slots x 32 centres per step (k = 20: five exact steps)
    a.steps = (k + (1 << a.slog) - 1) >> a.slog;
    const long per_item = TM >> a.slog;
    a.nitems = (int)((a.rows + per_item - 1) / per_item);
    static int n_sm = 0;
    if (!n_sm) {
        int dev = 0;
        PAB_CUDA(cudaGetDevice(&dev));
        PAB_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    }
    a.counter = pab_tile_counter_pair(st);                         // self-resetting (item, finished) pair (api.cu)
    if (!a.counter) return PAB_EINVAL;                             // caller falls back to the warp-specialised kernel
    const size_t smem = (size_t)a.wbytes + SN_STG_BYTES + SN_CTAB_BYTES + 64;
    int sms = n_sm;
    if (g_tc_max_ctas > 0 && sms > g_tc_max_ctas) sms = g_tc_max_ctas;
    // Three CTAs (168 registers x 128 threads each) fill the register file: fastest alone (59.6 vs 69.7 us for SA0 of 32 clouds), but
    // in the engine's stream mode — g_tc_max_ctas set, another stream's geometry kernels co-resident — it starves those kernels
    // (29.3 k vs 35.7 k submaps/s measured); two leave a third of the registers free
    const int per_sm = g_sn_ctas_per_sm > 0 ? g_sn_ctas_per_sm : (g_tc_max_ctas > 0 ? 2 : 3);
    long grid = (long)sms * per_sm;
    if (grid > a.nitems) grid = a.nitems;
    if (a.slog == 2 && cin_t == 3) return sn_launch<3, 2>(a, grid, smem, st);
    if (a.slog == 2 && cin_t == 6) return sn_launch<6, 2>(a, grid, smem, st);
    if (a.slog == 2) return sn_launch<8, 2>(a, grid, smem, st);
    if (a.slog == 1) return sn_launch<8, 1>(a, grid, smem, st);
    return sn_launch<8, 0>(a, grid, smem, st);
}
