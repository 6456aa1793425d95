// mlp_tc.cu — fused SharedMLP on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
// Same fusion as mlp.cu (neighbour gather / 3-NN interpolation -> every MLP layer -> max over K or row store, tile
// resident on chip), but the layer GEMMs run as tcgen05.mma with fp32 accumulators in tensor memory:
//
//   * a persistent CTA owns 128-row tiles.  Every layer input is kept as TWO bf16 planes (x = hi + lo, hi = bf16(x),
//     lo = bf16(x - hi)); the folded weights are split the same way on the host.  Each 16-wide k-step issues
//         D += hi(A) * hi(W)  +  lo(A) * hi(W)  +  hi(A) * lo(W)
//     which carries ~16 mantissa bits through every product (error ~2^-17 per term, fp32 accumulation) — the
//     descriptors stay within the 1e-4 contract where a single TF32/bf16 pass does not.
//   * the layer-0 operand (gathered rows) lives in shared memory in the canonical K-major SWIZZLE_128B layout and is
//     written by dedicated LOADER warps (6 in FP, 4 in SA mode); the operands of layers >= 1 never touch shared memory: the epilogue writes
//     them straight back into TENSOR MEMORY (tcgen05.st) and the next layer's MMAs read A from TMEM (TS form).  TMEM
//     holds D (256 fp32 columns) + the hi plane (128 columns of bf16 pairs) + the lo plane (128) = 512 columns.
//     Shared memory is therefore free again as soon as the layer-0 MMAs have completed, and the loaders gather tile i+1
//     while layers 1.. of tile i run.
//   * warp 0 streams weight blocks (<= 128 output channels x 64 k of one plane) with TMA (cp.async.bulk.tensor,
//     mbarrier complete_tx) through a ring of stages, running ahead across layers and tiles; warp 1 issues the MMAs
//     (warp-uniform loop, one elected lane) and frees stages with tcgen05.commit; warps 2..9 (256 threads) read the
//     accumulators with tcgen05.ld, add the folded BN shift (+ the rank-3 xyz update of layer 0, whose K would otherwise
//     not be a multiple of 64), ReLU, split to bf16 hi/lo.  The last layer is staged as fp32 64 columns at a time and
//     either max-pooled over the K neighbours (SA) or stored as coalesced rows (FP).
#include "tc_common.cuh"

unsigned int *pab_tile_counter_pair(cudaStream_t st);   // api.cu

int g_tc_max_ctas = 0;      // 0 = one persistent CTA per SM; otherwise a cap (leaves SMs to kernels of other streams); shared with vlad_tc.cu

namespace {

using namespace tc;

constexpr int NBLK_MAX = 128;                 // max output channels per weight stage (UMMA N)
constexpr int NEPI = 256;                     // epilogue threads (warps 2..9)
constexpr int NLOAD = 192;                    // loader threads (warps 10..15); the SA gathers use the first four warps only
constexpr int NLOAD_SA = 128;
constexpr int TC_THREADS = 64 + NEPI + NLOAD;
constexpr int MAX_STAGES = 8;
constexpr int MAX_LAYERS = 3;
constexpr int D_COLS = 256;                   // accumulator columns; TMEM columns [256,384) = hi plane, [384,512) = lo plane
constexpr uint32_t AH_COL = 256, AL_COL = 384;
constexpr uint32_t XTRA_COL = 64;             // plane columns [64,80) hold layer 0's xyz extras: the epilogue of a split pass's first
                                              // n-block rewrites columns [0,64) while the second n-block's MMAs still read the extras
constexpr int STG_COLS = 64;                  // fp32 staging of the last layer: [128 rows][64 columns]
constexpr int STG_PITCH = STG_COLS + 4;        // floats per staged row: 68 = 4 mod 32 keeps row-wise 16-byte writes and column-wise reads conflict-free
constexpr int STG_BYTES = TM * STG_PITCH * 4;                      // SA: one CTA-wide tile (rows of a group span warps)
constexpr int STG_BYTES_FP = 8 * 32 * 16 * 4;                      // FP: a private [32 rows][16 cols] tile per epilogue warp

enum { TC_SA = 1, TC_FP = 2 };

struct alignas(64) TcArgs {
    CUtensorMap tm[MAX_LAYERS][2];
    const float *shift[MAX_LAYERS];
    int K[MAX_LAYERS], N[MAX_LAYERS], relu[MAX_LAYERS];
    int ksteps[MAX_LAYERS];                    // 16-wide k-steps that carry data (the rest of the last 64-chunk is zero padding)
    int n_layers, n_stages, mode;
    int stage_bytes;                           // bytes of one weight stage: 16 KB (one plane of a <=128-channel block) or, with `pair`,
    int pair;                                  // 32 KB: hi AND lo plane of a wide block behind ONE barrier round trip
    int nbuf;                                  // layer-0 operand buffers in shared memory: 2 = the loaders stage tile i+1 while the layer-0
                                               // MMAs of tile i still read theirs (narrow modules, where the region is small)
    int planes;                                // 2: bf16 hi/lo split, 3 MMAs per product (fp32 contract); 1: plain bf16 operands, 1 MMA
    int csize, iters;                          // CTAs per cluster sharing the weight stream (1 or 2); tile-loop trips (equal for all CTAs)
    int serial_epi;                            // experiment: the epilogue of a split pass waits for BOTH n-blocks (no MMA / epilogue overlap)
    int dynamic;                               // 1: CTAs draw tiles from *counter (atomic) instead of the static blockIdx + i*grid sequence
    unsigned int *counter;                     // [0] next tile, [1] CTAs finished: the last CTA to leave zeroes both for the next launch
    int coff[MAX_LAYERS];                      // offset of each layer's shift vector in the smem constant table
    int a_region;                              // bytes of the layer-0 operand region (hi plane, then lo plane)
    int gchunks;                               // 64-channel chunks of the layer-0 operand staged at a time: a wide input (FP2's
                                               // K = 768) goes through the region in several groups, accumulating in TMEM
    long rows;                                 // SA: centres, FP: points
    int ntiles;
    // layer-0 "extra" channels (the xyz part), applied as a rank-n update from the fp32 weight rows
    const float *w_extra;                      // layers[0].wt + extra_row0 * N0 (unused: the extras go through the tensor cores)
    int n_extra;                               // their weights are the EXTRA 64-column chunk at the end of layer 0's planes
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
    const int *row_order;                      // optional processing order of a cloud's points (NULL: index order)
    long order_stride;                         // ints between the orders of consecutive clouds
    float *out;
    long long *trace;                          // optional timeline of CTA 0 (pab_tune_tc_trace), [tile][phase][event] clock64 stamps
};

#define TC_TRACE(tile_it, ph, ev)                                                                  \
    do {                                                                                          \
        if (a.trace && blockIdx.x == 0 && (tile_it) < 8 && lane == 0)                             \
            a.trace[(((tile_it)*4 + (ph)) * 8) + (ev)] = clock64();                                \
    } while (0)

// A operand from tensor memory (TS form): rows = TMEM lanes, k pairs packed in 32-bit columns
__device__ __forceinline__ void umma_f16_ts_if(uint32_t issue, uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi,
                                               uint32_t idesc, uint32_t accumulate) {
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
// The MMAs of one 64-channel chunk against one weight plane, as straight-line code: operand source (shared memory / tensor
// memory), one or two operand planes and "all four k-steps" are compile-time, so nothing but the tcgen05.mma and an address
// add is left per MMA.  With those three decided by run-time branches inside the k-step loop every MMA sat behind uniform
// branches and a reconvergence pair, and the issuing thread — not the tensor pipe — set the pace of the 128-column MMAs.
// Order per chunk as before: [hi(A) hi(W), lo(A) hi(W)] per k-step.
template <bool FROM_SMEM, bool TWO, bool FULL>
__device__ __forceinline__ void issue_hi_plane(uint32_t leader, uint32_t d, uint32_t ah, uint32_t al, uint32_t sb, uint32_t idesc,
                                               uint32_t acc0, int kn) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        if (FULL || ks < kn) {
            const uint32_t acc = ks == 0 ? acc0 : 1u;
            if (FROM_SMEM) {
                umma_f16_if(leader, d, ah + 2 * ks, UMMA_DESC_HI, sb + 2 * ks, UMMA_DESC_HI, idesc, acc);
                if (TWO) umma_f16_if(leader, d, al + 2 * ks, UMMA_DESC_HI, sb + 2 * ks, UMMA_DESC_HI, idesc, 1);
            } else {
                umma_f16_ts_if(leader, d, ah + 8 * ks, sb + 2 * ks, UMMA_DESC_HI, idesc, acc);
                if (TWO) umma_f16_ts_if(leader, d, al + 8 * ks, sb + 2 * ks, UMMA_DESC_HI, idesc, 1);
            }
        }
    }
}
// hi(A) lo(W) over the chunk's k-steps
template <bool FROM_SMEM, bool FULL>
__device__ __forceinline__ void issue_lo_plane(uint32_t leader, uint32_t d, uint32_t ah, uint32_t sb, uint32_t idesc, int kn) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        if (FULL || ks < kn) {
            if (FROM_SMEM) umma_f16_if(leader, d, ah + 2 * ks, UMMA_DESC_HI, sb + 2 * ks, UMMA_DESC_HI, idesc, 1);
            else umma_f16_ts_if(leader, d, ah + 8 * ks, sb + 2 * ks, UMMA_DESC_HI, idesc, 1);
        }
    }
}

__global__ void __launch_bounds__(TC_THREADS, 1) mlp_tc_kernel(const __grid_constant__ TcArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *a1 = smem;                                            // layer-0 operand, hi plane
    uint8_t *a2 = a.planes == 2 ? a1 + (size_t)(a.a_region >> 1) : nullptr;   //            lo plane (bf16 hi/lo mode only)
    uint8_t *stages = a1 + (size_t)a.a_region * a.nbuf;
    uint8_t *stg = stages + (size_t)a.n_stages * a.stage_bytes;
    uint8_t *misc = stg + (a.mode == TC_SA ? STG_BYTES : STG_BYTES_FP);
    uint64_t *full = reinterpret_cast<uint64_t *>(misc);
    uint64_t *empty = full + MAX_STAGES;
    uint64_t *a_full = empty + MAX_STAGES;                         // loaders -> MMA: layer-0 operand of the tile staged
    uint64_t *a_empty = a_full + 1;                                // MMA -> loaders: layer-0 MMAs of the tile completed
    // MMA <-> epilogue hand-off, per accumulator PASS (<= 256 columns).  A pass made of two 128-column n-blocks is pipelined
    // at n-block granularity ("split" pass): d_ready[i] = n-block i accumulated, t_ready[i] = its columns drained and (not the
    // last layer) its half of the next operand written to the TMEM planes, p_free = the MMAs of n-block 1 no longer read the
    // plane columns n-block 0's epilogue overwrites.  Every pass completes exactly one phase of each of the five barriers.
    uint64_t *d_ready = a_full + 2;                                // MMA -> epilogue: n-block 0 (or the whole unsplit pass)
    uint64_t *t_ready = a_full + 3;                                // epilogue -> MMA: D columns [0,128) / plane columns [0,64)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(a_full + 4);
    uint64_t *tq_full = a_full + 5;                                // dynamic tile queue: entry i published (ring of 8)
    int *tile_list = reinterpret_cast<int *>(tq_full + 8);         // [8] tile index or -1 (no more tiles)
    uint64_t *d_ready1 = reinterpret_cast<uint64_t *>(misc + 272);  // n-block 1
    uint64_t *t_ready1 = d_ready1 + 1;                             // D columns [128,256) / plane columns [64,128) (+ next tile's extras)
    uint64_t *p_free = d_ready1 + 2;
    constexpr int ABUF_STEP = 21;                                  // the second operand buffer's a_full / a_empty sit 168 bytes further
                                                                   // (misc + 296 / + 304): plain pointer arithmetic, no local arrays
    float *ctab = reinterpret_cast<float *>(misc + 384);           // [shift of every layer | pre layer]

    const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;
    if (a.trace && tid == 0) {                                     // per-CTA wall clock (ns): start here, end before the exit
        long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        a.trace[512 + 2 * blockIdx.x] = t;
    }

    if (tid == 0) {
        for (int s = 0; s < a.n_stages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, a.csize); }
        for (int i = 0; i < 2; ++i) { mbar_init(a_full + i * ABUF_STEP, a.mode == TC_FP ? NLOAD : NLOAD_SA); mbar_init(a_empty + i * ABUF_STEP, 1); }
        mbar_init(d_ready, 1);
        mbar_init(t_ready, NEPI);
        mbar_init(d_ready1, 1);
        mbar_init(t_ready1, NEPI);
        mbar_init(p_free, 1);
        for (int i = 0; i < 8; ++i) mbar_init(tq_full + i, 1);
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
        if (a.pre_cout > 0) {
            for (int i = tid; i < a.pre_cin * a.pre_cout; i += TC_THREADS) ctab[a.pre_off + i] = __ldg(a.pre_wt + i);
            for (int i = tid; i < a.pre_cout; i += TC_THREADS) ctab[a.pre_off + a.pre_cin * a.pre_cout + i] = __ldg(a.pre_shift + i);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (a.csize > 1) cluster_sync_all();                          // the peer's barriers exist before anything is multicast
    tc_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    const uint32_t crank = a.csize > 1 ? cluster_ctarank() : 0;
    const uint16_t cmask_all = (uint16_t)((1u << a.csize) - 1);
    // i-th tile of this CTA, or -1 when there is none.  Static: blockIdx + i * grid for i < iters (in a cluster every CTA runs
    // the same number of trips, trailing ones on empty tiles).  Dynamic: the loader's first thread draws tiles from a global
    // counter and publishes them through an 8-entry shared-memory ring (one mbarrier per entry); every role reads the same
    // sequence, so a CTA that starts late — its SM was busy with another stream's kernel — simply takes fewer tiles.
    auto tile_at = [&](uint32_t i) -> int {
        if (!a.dynamic) return i < (uint32_t)a.iters ? (int)(blockIdx.x + i * gridDim.x) : -1;
        mbar_wait(tq_full + (i & 7), (i >> 3) & 1);
        return tile_list[i & 7];
    };

    // FP: the point a global row stands for.  With a row order (the Morton order of the level's spatial index) consecutive
    // rows are neighbours in space and share their 3-NN rows of the known cloud, so the loaders' gathers hit in L1 instead of
    // going to L2 one 1-KB row each; every point is still computed exactly once and stored at its own position.
    auto point_of = [&](long gr) -> long {
        if (!a.row_order) return gr;
        const long cloud = gr / a.n;
        return cloud * a.n + __ldg(a.row_order + cloud * a.order_stride + (gr - cloud * a.n));
    };

    if (warp == 0) {
        // ================= TMA producer: weight blocks in (tile, layer, n-block, k-chunk, plane) order ==============
        if (lane == 0) {
            uint32_t s = 0, ph = 0;
            for (uint32_t it = 0; tile_at(it) >= 0; ++it) {
                for (int l = 0; l < a.n_layers; ++l) {
                    // layer 0 with extra channels: one more 64-column chunk (k = K[0]..) holding their weights; layer 0 staged in
                    // several operand groups: group-major order (the MMA warp finishes a group for every n-block)
                    const int nkc_main = (a.ksteps[l] + 3) >> 2;
                    const int nkc = nkc_main + (l == 0 && a.n_extra > 0 ? 1 : 0);
                    const int gc = l == 0 ? a.gchunks : nkc, ngroups = l == 0 ? (nkc_main + gc - 1) / gc : 1;
                    const int nbr = min(NBLK_MAX, a.N[l]), nnb = a.N[l] / nbr;
                    for (int g = 0; g < ngroups; ++g)
                      for (int nb = 0; nb < nnb; ++nb)
                        for (int kc = g * gc; kc < (g == ngroups - 1 ? nkc : (g + 1) * gc); ++kc)
                            for (int pl = 0; pl < a.planes; ++pl) {
                                // narrow layers (<= 64 output channels): both planes share one slot and one barrier round trip
                                const bool separate = nbr > 64 && !a.pair;      // each plane of a wide block in its own stage
                                const bool first = pl == 0 || separate;
                                uint8_t *dst = stages + (size_t)s * a.stage_bytes + (first ? 0 : nbr * 128);
                                if (first) {
                                    mbar_wait(empty + s, ph ^ 1);                // released by every CTA of the cluster
                                    mbar_expect_tx(full + s, (uint32_t)nbr * 128u * (separate ? 1u : (uint32_t)a.planes));
                                }
                                if (a.csize == 1) {
                                    tma_load_2d(dst, &a.tm[l][pl], kc * KCH, nb * nbr, full + s);
                                } else {                                         // this CTA's share of the block, delivered to all
                                    const int share = nbr / a.csize;
                                    tma_load_2d_mc(dst + (size_t)crank * share * 128, &a.tm[l][pl], kc * KCH, nb * nbr + (int)crank * share,
                                                   full + s, cmask_all);
                                }
                                if (separate || pl == a.planes - 1) {
                                    if (++s == (uint32_t)a.n_stages) { s = 0; ph ^= 1; }
                                }
                            }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================= MMA issuer: all lanes run the loops, one elected lane issues ===========================
        const uint32_t leader = elect_one();
        const uint32_t a1_lo = umma_desc_lo(smem_u32(a1)), a2_lo = umma_desc_lo(smem_u32(a1) + (uint32_t)(a.a_region >> 1));
        const uint32_t st_lo = umma_desc_lo(smem_u32(stages));
        const uint32_t st_step = (uint32_t)a.stage_bytes >> 4;
        uint32_t s = 0, ph = 0, pcount = 0, tcount = 0, gcount = 0;
        bool t1_pending = false;
        const bool two = a.planes == 2;
        for (int it = 0; tile_at((uint32_t)it) >= 0; ++it, ++tcount) {
            for (int l = 0; l < a.n_layers; ++l) {
                const int ksteps = a.ksteps[l], nkc_main = (ksteps + 3) >> 2;
                const int nkc = nkc_main + (l == 0 && a.n_extra > 0 ? 1 : 0);        // + the extras' chunk (one k-step, A from TMEM)
                const int nbr = min(NBLK_MAX, a.N[l]), nnb = a.N[l] / nbr, nb_pass = D_COLS / nbr;   // n-blocks per accumulator pass
                const uint32_t idesc = umma_idesc(nbr);
                const int gc = l == 0 ? a.gchunks : nkc, ngroups = l == 0 ? (nkc_main + gc - 1) / gc : 1;
                for (int g = 0; g < ngroups; ++g) {
                const uint32_t abuf = a.nbuf == 2 ? (gcount & 1u) : 0u, ause = a.nbuf == 2 ? (gcount >> 1) : gcount;
                const uint32_t aoff = abuf * ((uint32_t)a.a_region >> 4);   // descriptor offset of this tile's operand buffer
                if (l == 0) mbar_wait(a_full + abuf * ABUF_STEP, ause & 1);   // this group of layer-0 operand chunks is staged
                const int kc_lo = g * gc, kc_hi = g == ngroups - 1 ? nkc : (g + 1) * gc;
                const bool last_g = g == ngroups - 1;
                for (int nb = 0; nb < nnb; ++nb) {
                    const int nbp = nb % nb_pass;
                    const bool split = nbr == NBLK_MAX && nb_pass == 2 && nb - nbp + 2 <= nnb;   // this pass has two 128-column n-blocks
                    if (nbp == 0 && g == 0) {                    // new pass: the epilogue has drained D[0,128) (and written planes [0,64))
                        mbar_wait(t_ready, pcount & 1);
                        tc_fence_after();
                        t1_pending = true;
                        TC_TRACE(it, l, 0);
                    }
                    if (nbp == 1 && t1_pending) {                // second n-block: D[128,256) drained
                        mbar_wait(t_ready1, pcount & 1);
                        tc_fence_after();
                        t1_pending = false;
                    }
                    const uint32_t d = tmem + (uint32_t)(nbp * nbr);
                    const int kc_free = min(kc_lo + 1, kc_hi - 1);   // after this chunk of n-block 1 plane columns [0,64) are dead
                    for (int kc = kc_lo; kc < kc_hi; ++kc) {
                        const bool xk = kc >= nkc_main;          // the extras' chunk: k-step 0 only, operand in plane columns XTRA_COL..
                        const bool from_smem = l == 0 && !xk;
                        const int kn = xk ? 1 : min(4, ksteps - 4 * kc);
                        const uint32_t ka = (uint32_t)(kc - kc_lo) * (A_CHUNK >> 4);
                        const uint32_t ta = xk ? XTRA_COL : (uint32_t)(kc * 32);
                        if (!from_smem && ta >= 64 && t1_pending) {  // operand columns the previous pass's second half wrote
                            mbar_wait(t_ready1, pcount & 1);
                            tc_fence_after();
                            t1_pending = false;
                        }
                        // hi weight plane: hi(A) * hi(W) + lo(A) * hi(W)
                        mbar_wait(full + s, ph);
                        uint32_t sb = st_lo + s * st_step;
                        const uint32_t ah = from_smem ? a1_lo + aoff + ka : tmem + AH_COL + ta;
                        const uint32_t al = from_smem ? a2_lo + aoff + ka : tmem + AL_COL + ta;
                        const uint32_t acc0 = kc != 0;
                        if (kn == 4) {
                            if (from_smem) { if (two) issue_hi_plane<true, true, true>(leader, d, ah, al, sb, idesc, acc0, 4);
                                             else issue_hi_plane<true, false, true>(leader, d, ah, al, sb, idesc, acc0, 4); }
                            else           { if (two) issue_hi_plane<false, true, true>(leader, d, ah, al, sb, idesc, acc0, 4);
                                             else issue_hi_plane<false, false, true>(leader, d, ah, al, sb, idesc, acc0, 4); }
                        } else {
                            if (from_smem) { if (two) issue_hi_plane<true, true, false>(leader, d, ah, al, sb, idesc, acc0, kn);
                                             else issue_hi_plane<true, false, false>(leader, d, ah, al, sb, idesc, acc0, kn); }
                            else           { if (two) issue_hi_plane<false, true, false>(leader, d, ah, al, sb, idesc, acc0, kn);
                                             else issue_hi_plane<false, false, false>(leader, d, ah, al, sb, idesc, acc0, kn); }
                        }
                        if (two) {
                        if (nbr > 64 && !a.pair) {               // wide block, separate stages: the lo plane sits in the next slot
                            if (a.csize == 1) umma_commit_if(leader, empty + s);
                            else umma_commit_mc_if(leader, empty + s, cmask_all);
                            if (++s == (uint32_t)a.n_stages) { s = 0; ph ^= 1; }
                            mbar_wait(full + s, ph);
                            sb = st_lo + s * st_step;
                        } else {                                 // narrow block: the lo plane follows the hi plane in the same slot
                            sb += (uint32_t)nbr * 8;
                        }
                        // lo weight plane: hi(A) * lo(W)
                        if (kn == 4) {
                            if (from_smem) issue_lo_plane<true, true>(leader, d, ah, sb, idesc, 4);
                            else issue_lo_plane<false, true>(leader, d, ah, sb, idesc, 4);
                        } else {
                            if (from_smem) issue_lo_plane<true, false>(leader, d, ah, sb, idesc, kn);
                            else issue_lo_plane<false, false>(leader, d, ah, sb, idesc, kn);
                        }
                        }
                        if (a.csize == 1) umma_commit_if(leader, empty + s);
                        else umma_commit_mc_if(leader, empty + s, cmask_all);
                        if (++s == (uint32_t)a.n_stages) { s = 0; ph ^= 1; }
                        if (split && last_g && nbp == 1 && kc == kc_free) umma_commit_if(leader, p_free);
                    }
                    if (last_g) {
                        const bool pass_end = nbp == nb_pass - 1 || nb == nnb - 1;
                        if (pass_end && t1_pending) {
                            // nothing in this pass needed the second half: consume its phase — BEFORE the commits below.  Once
                            // d_ready is committed the epilogue may finish this pass and complete the NEXT phase of t_ready1; a
                            // waiter that then still polls for this one sees the parity of an incomplete phase and never wakes
                            // (observed: one launch in a few hundred when the issuing warp was delayed after the commits).
                            mbar_wait(t_ready1, pcount & 1);
                            t1_pending = false;
                        }
                        if (split) {
                            umma_commit_if(leader, nbp == 0 ? d_ready : d_ready1);   // this n-block's accumulators complete
                        } else if (pass_end) {
                            umma_commit_if(leader, d_ready);
                            umma_commit_if(leader, d_ready1);
                            umma_commit_if(leader, p_free);
                        }
                        if (pass_end) {
                            ++pcount;
                            TC_TRACE(it, l, 1);
                        }
                    }
                }
                if (l == 0) { umma_commit_if(leader, a_empty + abuf * ABUF_STEP); ++gcount; }   // operand group consumed: loaders may stage the next one
                }
            }
        }
        __syncwarp();
    } else if (warp < 2 + NEPI / 32) {
        // ================= epilogue warps ===========================================================================
        const int et = tid - 64;                             // 0..255
        const int ewarp = warp - 2;                          // 0..7
        const int q = warp & 3;                              // TMEM lane quarter this warp may access
        const int half = ewarp >> 2;                         // column half handled by this warp
        const int row = q * 32 + lane;                       // accumulator row owned by this thread
        const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
        uint32_t pcount = 0;
        const int G = a.mode == TC_SA ? TM / a.k : 0;

        // The rank-3 part of layer 0 (xyz_j - xyz_i of a neighbour, or the raw xyz of a point) goes through the tensor cores
        // as one more k-step: the thread that owns a row writes its extras as bf16 hi/lo pairs into columns 0..15 of the
        // operand planes (free while layer 0 runs), for the first tile here and for every next tile at the end of the
        // current tile's last layer, each time BEFORE the arrival that lets the MMA warp start that tile's layer 0.
        auto load_extras = [&](int tile, float (&xe)[3]) {
            xe[0] = xe[1] = xe[2] = 0.f;
            if (a.mode == TC_SA) {
                const int g = row / a.k, sidx = row - g * a.k;
                const long ci = (long)tile * G + g;
                if (g < G && ci < a.rows) {
                    const long cloud = ci / a.m;
                    const long pc = cloud * a.n + __ldg(a.center_idx + ci);
                    const long pn = cloud * a.n + __ldg(a.nbr_idx + ci * a.nbr_stride + sidx);
#pragma unroll
                    for (int e = 0; e < 3; ++e) xe[e] = __ldg(a.xyz + pn * 3 + e) - __ldg(a.xyz + pc * 3 + e);
                }
            } else {
                const long gr = (long)tile * TM + row;
                if (gr < a.rows) {
                    const long p = point_of(gr);
                    for (int e = 0; e < a.n_extra; ++e) xe[e] = __ldg(a.skip_feat + p * a.c_skip + e);
                }
            }
        };
        auto store_extras = [&](const float (&xe)[3]) {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) hi[i] = lo[i] = 0u;
            split_pack(xe[0], xe[1], hi[0], lo[0]);
            split_pack(xe[2], 0.f, hi[1], lo[1]);
            tmem_st16(trow + AH_COL + XTRA_COL, hi);
            if (a.planes == 2) tmem_st16(trow + AL_COL + XTRA_COL, lo);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        };
        const bool xwriter = a.n_extra > 0 && half == 0;
        float xe_next[3] = {0.f, 0.f, 0.f};
        int tile = tile_at(0);
        if (xwriter && tile >= 0) {
            load_extras(tile, xe_next);
            store_extras(xe_next);
        }
        tc_fence_before();
        mbar_arrive(t_ready);                                  // completion #0: the MMA warp may start the first phase
        mbar_arrive(t_ready1);

        for (int it = 0; tile >= 0; ++it) {                        // static cluster mode: trailing tiles >= ntiles are empty
            const int tile_next = tile_at((uint32_t)it + 1);
            if (xwriter && tile_next >= 0) load_extras(tile_next, xe_next);              // consumed at the end of this tile
            int out_pt[4] = {-1, -1, -1, -1};                     // FP: points of the rows this lane stores (row r8*8 + lane/4 of the quarter)
            if (a.mode == TC_FP) {
#pragma unroll
                for (int r8 = 0; r8 < 4; ++r8) {
                    const long gr = (long)tile * TM + q * 32 + r8 * 8 + (lane >> 2);
                    out_pt[r8] = gr < a.rows ? (int)point_of(gr) : -1;
                }
            }
            for (int l = 0; l < a.n_layers; ++l) {
                const int N = a.N[l];
                const bool last = l == a.n_layers - 1;
                const bool relu = a.relu[l] != 0;
                const float *shl = ctab + a.coff[l];
                const int npass = (N + D_COLS - 1) / D_COLS;
                for (int pass = 0; pass < npass; ++pass, ++pcount) {
                    const int ncols = min(D_COLS, N - pass * D_COLS);  // accumulator columns of this pass
                    // split pass (two 128-column n-blocks): segment i = n-block i, each half of the workers takes 64 of its columns
                    // as soon as THAT n-block is accumulated; otherwise one segment, the halves share the pass's columns
                    const bool split = ncols == D_COLS;
                    const int nseg = split ? 2 : 1;
                    const int per = split ? 64 : (ncols >= 64 ? ncols / 2 : ncols);   // columns per worker half and segment (batches of 32)
                    const bool active = ncols >= 64 || half == 0;
                    for (int seg = 0; seg < nseg; ++seg) {
                    const int segbase = split ? seg * NBLK_MAX : 0;
                    mbar_wait(seg == 0 ? d_ready : d_ready1, pcount & 1);
                    if (!split || a.serial_epi) mbar_wait(d_ready1, pcount & 1);
                    tc_fence_after();
                    if (ewarp == 0) TC_TRACE(it, l, (split && seg == 0) ? 7 : 2);
                    for (int cb = 0; cb < per; cb += 32) {
                        const int dcol = segbase + (ncols >= 64 ? half * per : 0) + cb;  // first accumulator column of this batch
                        const int col = pass * D_COLS + dcol;                          // output channel
                        float v[32];
                        if (active) {
                            tmem_ld32(trow + (uint32_t)dcol, v);
#pragma unroll
                            for (int u = 0; u < 8; ++u) {
                                const float4 sh = *reinterpret_cast<const float4 *>(shl + col + 4 * u);
                                v[4 * u] += sh.x; v[4 * u + 1] += sh.y; v[4 * u + 2] += sh.z; v[4 * u + 3] += sh.w;
                            }
                            if (relu) {
#pragma unroll
                                for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
                            }
                        }
                        if (!last) {
                            uint32_t hi[16], lo[16];
                            if (active) {
#pragma unroll
                                for (int i = 0; i < 16; ++i) split_pack(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
                            }
                            if (seg == 0 && cb == 0) {           // plane columns [0,64) are still read by the second n-block's MMAs
                                mbar_wait(p_free, pcount & 1);
                                tc_fence_after();
                            }
                            if (active) {                       // next layer's operand: bf16 pairs into the TMEM planes
                                tmem_st16(trow + AH_COL + (uint32_t)(dcol >> 1), hi);
                                if (a.planes == 2) tmem_st16(trow + AL_COL + (uint32_t)(dcol >> 1), lo);
                            }
                        } else {
                            const bool final_batch = cb + 32 >= per;
                            if (final_batch) {                   // every accumulator column of the segment has been read
                                // last segment of the tile: the operand planes are free (this layer's MMAs completed before
                                // d_ready): stage the next tile's layer-0 extras there before releasing the MMA warp
                                if (xwriter && pass == npass - 1 && seg == nseg - 1 && tile_next >= 0) store_extras(xe_next);
                                tc_fence_before();
                                if (ewarp == 0) TC_TRACE(it, l, 3);
                                mbar_arrive(seg == 0 ? t_ready : t_ready1);
                                if (!split) mbar_arrive(t_ready1);
                            }
                            if (a.mode == TC_SA) {
                                // one staging round: 32 columns of each half -> [128][64] fp32 -> max over the K rows of a group
                                if (active) {
                                    const int sc0 = ncols >= 64 ? half * 32 : 0;
                                    float *srow = reinterpret_cast<float *>(stg) + row * STG_PITCH + sc0;
#pragma unroll
                                    for (int u = 0; u < 8; ++u)
                                        *reinterpret_cast<float4 *>(srow + 4 * u) = make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
                                }
                                asm volatile("bar.sync 1, %0;" ::"n"(NEPI) : "memory");   // staging complete (epilogue warps only)
                                const int rcols = ncols >= 64 ? 64 : ncols;            // staged columns this round (32 or 64)
                                for (int e = et; e < G * rcols; e += NEPI) {
                                    const int g = e / rcols, c = e - g * rcols;
                                    const long ci = (long)tile * G + g;
                                    if (ci >= a.rows) continue;
                                    const float *scol = reinterpret_cast<const float *>(stg) + (g * a.k) * STG_PITCH + c;
                                    float mx = scol[0];
                                    for (int sidx = 1; sidx < a.k; ++sidx) mx = fmaxf(mx, scol[sidx * STG_PITCH]);
                                    const int oc = pass * D_COLS + segbase + (c >> 5) * per + cb + (c & 31);
                                    a.out[ci * N + oc] = mx;
                                }
                                asm volatile("bar.sync 1, %0;" ::"n"(NEPI) : "memory");   // staging consumed before it is rewritten
                            } else if (active) {
                                // FP: this warp's 32 rows x 32 columns leave through its private [32][16] fp32 tile, 16 columns at a
                                // time: thread = row on the way in, 4 lanes per row (64 contiguous bytes) on the way out
                                float *ws = reinterpret_cast<float *>(stg) + ewarp * (32 * 16);
#pragma unroll
                                for (int sb = 0; sb < 2; ++sb) {
                                    __syncwarp();                                   // previous read-out finished
#pragma unroll
                                    for (int u = 0; u < 4; ++u)                     // 16-byte units XOR-swizzled by row
                                        *reinterpret_cast<float4 *>(ws + lane * 16 + ((u ^ ((lane >> 1) & 3)) << 2)) =
                                            make_float4(v[16 * sb + 4 * u], v[16 * sb + 4 * u + 1], v[16 * sb + 4 * u + 2], v[16 * sb + 4 * u + 3]);
                                    __syncwarp();
                                    const int u = lane & 3;
#pragma unroll
                                    for (int r8 = 0; r8 < 4; ++r8) {
                                        const int r = r8 * 8 + (lane >> 2);         // row of this warp's quarter
                                        const long p = out_pt[r8];   // rows < 2^31 (checked by the launcher)
                                        if (p >= 0)
                                            *reinterpret_cast<float4 *>(a.out + p * N + col + 16 * sb + 4 * u) =
                                                *reinterpret_cast<const float4 *>(ws + r * 16 + ((u ^ ((r >> 1) & 3)) << 2));
                                    }
                                }
                            }
                        }
                    }
                    if (!last) {
                        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                        tc_fence_before();
                        if (ewarp == 0) TC_TRACE(it, l, 3);
                        mbar_arrive(seg == 0 ? t_ready : t_ready1);
                        if (!split) mbar_arrive(t_ready1);
                    } else if (ewarp == 0) {
                        TC_TRACE(it, l, 6);                      // output of the tile stored
                    }
                    }
                }
            }
            tile = tile_next;
        }
    } else if (a.mode == TC_FP || tid < 64 + NEPI + NLOAD_SA) {
        // ================= loader warps: stage the layer-0 operand (hi/lo planes) of the next tile =================
        const int lt = tid - 64 - NEPI;                      // 0..191
        const int lwarp = lt >> 5;                           // 0..5; SA: rows [32*lwarp, 32*lwarp + 32) of warps 0..3
        const int units0 = ((a.ksteps[0] + 3) >> 2) * 8;     // 16-byte units per operand row (whole 64-chunks)
        const int gunits = a.gchunks * 8;                    // units per operand group (one group = the whole row unless K is wide)
        const int ngroups = (units0 + gunits - 1) / gunits;
        uint32_t gcount = 0;                                 // operand groups staged so far (a_full / a_empty completions)
        // operand buffer of group gcount (two buffers: alternate; the wait is for the MMAs that read THIS buffer two groups ago)
        auto a_wait_free = [&]() {
            const uint32_t buf = a.nbuf == 2 ? (gcount & 1u) : 0u, use = a.nbuf == 2 ? (gcount >> 1) : gcount;
            if (use > 0) mbar_wait(a_empty + buf * ABUF_STEP, (use - 1) & 1);
        };
        auto a_buf1 = [&]() { return a1 + (size_t)(a.nbuf == 2 ? (gcount & 1u) : 0u) * a.a_region; };
        auto a_publish = [&]() {
            mbar_arrive(a_full + (a.nbuf == 2 ? (gcount & 1u) : 0u) * ABUF_STEP);
            ++gcount;
        };
        // pre-layer mode (one thread per row): pipelined lookups, see the branch below
        float pre_in[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        long pre_pc = 0, pre_pn = 0;
        bool pre_valid = false, pre_valid2 = false;
        auto pre_hop1 = [&](int tile, long &pc, long &pn, bool &valid) {
            const int G = TM / a.k;
            const int g = lt / a.k, sidx = lt - g * a.k;
            const long ci = (long)tile * G + g;
            valid = g < G && ci < a.rows;
            pc = pn = 0;
            if (valid) {
                const long cloud = ci / a.m;
                pc = cloud * a.n + __ldg(a.center_idx + ci);
                pn = cloud * a.n + __ldg(a.nbr_idx + ci * a.nbr_stride + sidx);
            }
        };
        auto pre_hop2 = [&](long pc, long pn, bool valid, float (&in)[8]) {
#pragma unroll
            for (int i = 0; i < 8; ++i) in[i] = 0.f;
            if (valid) {
#pragma unroll
                for (int i = 0; i < 3; ++i) in[i] = __ldg(a.xyz + pn * 3 + i) - __ldg(a.xyz + pc * 3 + i);
#pragma unroll
                for (int i = 0; i < 5; ++i)
                    if (i < a.c) in[3 + i] = __ldg(a.feat + pn * a.c + i) - __ldg(a.feat + pc * a.c + i);
            }
        };
        // dynamic mode: this thread draws the CTA's tiles from the global counter and publishes them `ahead` entries beyond the
        // tile being staged (1: the epilogue prefetches the next tile's extras; 2: the pre-layer lookups run two tiles ahead)
        const bool premode = a.mode == TC_SA && a.pre_cout > 0;
        const uint32_t ahead = premode ? 2u : 1u;
        uint32_t published = 0;
        bool drained = false;
        auto publish_upto = [&](uint32_t last) {
            if (!a.dynamic || lt != 0) return;
            while (published <= last) {
                int t = -1;
                if (!drained) {
                    const unsigned v = atomicAdd(a.counter, 1u);
                    if (v < (unsigned)a.ntiles) t = (int)v;
                    else drained = true;
                }
                tile_list[published & 7] = t;
                mbar_arrive(tq_full + (published & 7));
                ++published;
            }
        };
        publish_upto(ahead);
        if (premode && tile_at(0) >= 0) {
            pre_hop1(tile_at(0), pre_pc, pre_pn, pre_valid);
            pre_hop2(pre_pc, pre_pn, pre_valid, pre_in);
            if (tile_at(1) >= 0) pre_hop1(tile_at(1), pre_pc, pre_pn, pre_valid2);
        }
        for (int it = 0;; ++it) {
            publish_upto((uint32_t)it + ahead);
            const int tile = tile_at((uint32_t)it);
            if (tile < 0) break;
            if (lwarp == 0) TC_TRACE(it, 0, 4);
            // every branch first issues the global loads that do not need the operand region (indices, weights, the pre-layer's
            // tiny input), THEN waits for the previous tile's layer-0 MMAs to release it: the lookups overlap the wait
            if (a.mode == TC_SA && a.pre_cout > 0) {
                // one thread per row: inputs [xyz_j - xyz_i ; f_j - f_i] (pre_cin <= 8 values), the pre-layer evaluated in
                // fp32 and staged (hi/lo); the rest of the 64-channel chunk is zeroed
                // Two-deep software pipeline over tiles: the row's inputs for tile t+1 (second hop: xyz / feature rows) and
                // its point indices for tile t+2 (first hop: centre / neighbour tables) are in flight while tile t is evaluated,
                // so no dependent global-load latency sits between two tiles of this thread.
                const int r = lt;
                float in[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) in[i] = pre_in[i];
                const bool valid = pre_valid;
                if (tile_at((uint32_t)it + 1) >= 0) {
                    pre_hop2(pre_pc, pre_pn, pre_valid2, pre_in);
                    pre_valid = pre_valid2;
                }
                {
                    const int t2 = tile_at((uint32_t)it + 2);
                    if (t2 >= 0) pre_hop1(t2, pre_pc, pre_pn, pre_valid2);
                }
                a_wait_free();
                uint8_t *a1 = a_buf1(), *a2 = a.planes == 2 ? a1 + (size_t)(a.a_region >> 1) : nullptr;     // this group's buffer
                const float *pw = ctab + a.pre_off, *ps = pw + a.pre_cin * a.pre_cout;
                for (int u = 0; u < a.pre_cout / 8; ++u) {
                    float v[8];
#pragma unroll
                    for (int o = 0; o < 8; ++o) {
                        float acc = ps[u * 8 + o];
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            if (i < a.pre_cin) acc = fmaf(in[i], pw[i * a.pre_cout + u * 8 + o], acc);
                        v[o] = valid ? (a.pre_relu ? fmaxf(acc, 0.f) : acc) : 0.f;
                    }
                    store_units(a1, a2, r, u, v);
                }
                // only whole 16-channel k-steps are read by the MMAs (ksteps = ceil(pre_cout / 16)): zero what is left of the
                // last one, not the rest of the 64-channel chunk
                const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                for (int j = a.pre_cout / 8; j < 2 * a.ksteps[0]; ++j) store_units(a1, a2, r, j, z);
            } else if (a.mode == TC_SA) {
                // lane r of the warp looks up the (centre, neighbour) point of row 32*lwarp + r once; rows are then staged
                // four at a time with the indices broadcast by shuffles
                const int G = TM / a.k;
                long my_pc = -1, my_pn = -1;
                {
                    const int r = lwarp * 32 + lane;
                    const int g = r / a.k, sidx = r - g * a.k;
                    const long ci = (long)tile * G + g;
                    if (g < G && ci < a.rows) {
                        const long cloud = ci / a.m;
                        my_pc = cloud * a.n + __ldg(a.center_idx + ci);
                        my_pn = cloud * a.n + __ldg(a.nbr_idx + ci * a.nbr_stride + sidx);
                    }
                }
                for (int g = 0; g < ngroups; ++g) {
                const int j_lo = g * gunits, j_hi = min(units0, j_lo + gunits);
                a_wait_free();
                uint8_t *a1 = a_buf1(), *a2 = a.planes == 2 ? a1 + (size_t)(a.a_region >> 1) : nullptr;     // this group's buffer
                // lanes -> (row of the pass, 16-byte unit): a narrow input (64 channels = 8 units per row) puts four rows on the
                // 32 lanes of one pass instead of leaving 24 lanes without loads; four passes' worth of rows are in flight together
                const int ug = j_hi - j_lo;
                const int lpr = ug >= 32 ? 32 : (ug >= 16 ? 16 : 8), rpp = 32 / lpr;     // lanes per row, rows per pass
                const int sub = lane / lpr, jl = lane - sub * lpr;
                for (int rr = 0; rr < 32; rr += 4 * rpp) {
                    long pc[4], pn[4];
#pragma unroll
                    for (int t4 = 0; t4 < 4; ++t4) {
                        pc[t4] = __shfl_sync(0xffffffffu, my_pc, rr + t4 * rpp + sub);
                        pn[t4] = __shfl_sync(0xffffffffu, my_pn, rr + t4 * rpp + sub);
                    }
                    for (int j = j_lo + jl; j < j_hi; j += lpr) {
                        float4 n0[4], n1[4], c0[4], c1[4];
#pragma unroll
                        for (int t4 = 0; t4 < 4; ++t4) {
                            n0[t4] = n1[t4] = c0[t4] = c1[t4] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (pc[t4] >= 0 && 8 * j < a.c) {
                                const float4 *fn = reinterpret_cast<const float4 *>(a.feat + pn[t4] * a.c) + 2 * j;
                                const float4 *fc = reinterpret_cast<const float4 *>(a.feat + pc[t4] * a.c) + 2 * j;
                                n0[t4] = __ldg(fn); n1[t4] = __ldg(fn + 1); c0[t4] = __ldg(fc); c1[t4] = __ldg(fc + 1);
                            }
                        }
#pragma unroll
                        for (int t4 = 0; t4 < 4; ++t4) {
                            const float v[8] = {n0[t4].x - c0[t4].x, n0[t4].y - c0[t4].y, n0[t4].z - c0[t4].z, n0[t4].w - c0[t4].w,
                                                n1[t4].x - c1[t4].x, n1[t4].y - c1[t4].y, n1[t4].z - c1[t4].z, n1[t4].w - c1[t4].w};
                            store_units(a1, a2, lwarp * 32 + rr + t4 * rpp + sub, j - j_lo, v);
                        }
                    }
                }
                if (g + 1 < ngroups) {                        // more groups of this tile follow
                    fence_proxy_async();
                    a_publish();
                }
                }
            } else {
                // FP: every lane holds the three neighbour indices / weights of one row of its warp; rows are staged RB = 2 at a time:
                // 12 x 16-byte gathers per lane, which the 128-register budget really keeps in flight together (with four rows
                // the compiler split the 24 loads into three dependent rounds); indices and weights broadcast by shuffles
                const int ku = a.c_known / 8;
                // the 64 batches of two rows are dealt round-robin to the six warps (batch = lwarp + 6*(lane/2)), so at any time the
                // warps work on 12 consecutive rows: with a spatial row order they share most of their known rows in L1.  The
                // gathers are latency-bound (registers cap the loads a warp keeps in flight), hence as many warps as the register
                // file allows: 512 threads x 128 registers
                constexpr int RB = 2, LW = NLOAD / 32;
                int my_base = -1, my_p = -1;
                int my_i[3] = {0, 0, 0};
                float my_w[3] = {0.f, 0.f, 0.f};
                {
                    const int batch = lwarp + LW * (lane / RB);
                    const long gr = (long)tile * TM + RB * batch + (lane % RB);
                    if (batch < TM / RB && gr < a.rows) {
                        const long p = point_of(gr);
                        my_p = (int)p;
                        my_base = (int)((p / a.n) * a.m);
#pragma unroll
                        for (int e = 0; e < 3; ++e) { my_i[e] = __ldg(a.idx3 + p * 3 + e); my_w[e] = __ldg(a.w3 + p * 3 + e); }
                    }
                }
                for (int g = 0; g < ngroups; ++g) {
                const int j_lo = g * gunits, j_hi = min(units0, j_lo + gunits);
                a_wait_free();
                uint8_t *a1 = a_buf1(), *a2 = a.planes == 2 ? a1 + (size_t)(a.a_region >> 1) : nullptr;     // this group's buffer
                for (int bi = 0; lwarp + LW * bi < TM / RB; ++bi) {           // batch = lwarp + LW * bi, rows RB * batch ..
                    const int rr = RB * bi, row0 = RB * (lwarp + LW * bi);       // source lane of the batch's first row, its tile row
                    const float *f0[RB], *f1[RB], *f2[RB];
                    float w0[RB], w1[RB], w2[RB];
                    bool ok[RB];
                    int pt[RB];
#pragma unroll
                    for (int t4 = 0; t4 < RB; ++t4) {
                        const int base = __shfl_sync(0xffffffffu, my_base, rr + t4);
                        pt[t4] = __shfl_sync(0xffffffffu, my_p, rr + t4);
                        const int i0 = __shfl_sync(0xffffffffu, my_i[0], rr + t4), i1 = __shfl_sync(0xffffffffu, my_i[1], rr + t4),
                                  i2 = __shfl_sync(0xffffffffu, my_i[2], rr + t4);
                        w0[t4] = __shfl_sync(0xffffffffu, my_w[0], rr + t4); w1[t4] = __shfl_sync(0xffffffffu, my_w[1], rr + t4);
                        w2[t4] = __shfl_sync(0xffffffffu, my_w[2], rr + t4);
                        ok[t4] = base >= 0;
                        const long b0 = ok[t4] ? base : 0;
                        f0[t4] = a.known_feat + (b0 + i0) * a.c_known; f1[t4] = a.known_feat + (b0 + i1) * a.c_known;
                        f2[t4] = a.known_feat + (b0 + i2) * a.c_known;
                    }
                    for (int j = j_lo + lane; j < j_hi; j += 32) {
                        float v[RB][8];
                        if (j < ku) {
                            float4 x0[RB], x1[RB], y0[RB], y1[RB], z0[RB], z1[RB];
#pragma unroll
                            for (int t4 = 0; t4 < RB; ++t4) {
                                x0[t4] = __ldg(reinterpret_cast<const float4 *>(f0[t4]) + 2 * j); x1[t4] = __ldg(reinterpret_cast<const float4 *>(f0[t4]) + 2 * j + 1);
                                y0[t4] = __ldg(reinterpret_cast<const float4 *>(f1[t4]) + 2 * j); y1[t4] = __ldg(reinterpret_cast<const float4 *>(f1[t4]) + 2 * j + 1);
                                z0[t4] = __ldg(reinterpret_cast<const float4 *>(f2[t4]) + 2 * j); z1[t4] = __ldg(reinterpret_cast<const float4 *>(f2[t4]) + 2 * j + 1);
                            }
#pragma unroll
                            for (int t4 = 0; t4 < RB; ++t4) {
                                // interpolation_forward: fma(w2,p2, fma(w0,p0, w1*p1)) (interpolation_cuda_kernel.cu:194)
                                v[t4][0] = __fmaf_rn(w2[t4], z0[t4].x, __fmaf_rn(w0[t4], x0[t4].x, __fmul_rn(w1[t4], y0[t4].x)));
                                v[t4][1] = __fmaf_rn(w2[t4], z0[t4].y, __fmaf_rn(w0[t4], x0[t4].y, __fmul_rn(w1[t4], y0[t4].y)));
                                v[t4][2] = __fmaf_rn(w2[t4], z0[t4].z, __fmaf_rn(w0[t4], x0[t4].z, __fmul_rn(w1[t4], y0[t4].z)));
                                v[t4][3] = __fmaf_rn(w2[t4], z0[t4].w, __fmaf_rn(w0[t4], x0[t4].w, __fmul_rn(w1[t4], y0[t4].w)));
                                v[t4][4] = __fmaf_rn(w2[t4], z1[t4].x, __fmaf_rn(w0[t4], x1[t4].x, __fmul_rn(w1[t4], y1[t4].x)));
                                v[t4][5] = __fmaf_rn(w2[t4], z1[t4].y, __fmaf_rn(w0[t4], x1[t4].y, __fmul_rn(w1[t4], y1[t4].y)));
                                v[t4][6] = __fmaf_rn(w2[t4], z1[t4].z, __fmaf_rn(w0[t4], x1[t4].z, __fmul_rn(w1[t4], y1[t4].z)));
                                v[t4][7] = __fmaf_rn(w2[t4], z1[t4].w, __fmaf_rn(w0[t4], x1[t4].w, __fmul_rn(w1[t4], y1[t4].w)));
                            }
                        } else {
#pragma unroll
                            for (int t4 = 0; t4 < RB; ++t4) {
                                float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
                                if (ok[t4] && 8 * (j - ku) < a.c_skip) {
                                    const float4 *sk = reinterpret_cast<const float4 *>(a.skip_feat + (long)pt[t4] * a.c_skip) + 2 * (j - ku);
                                    s0 = __ldg(sk); s1 = __ldg(sk + 1);
                                }
                                v[t4][0] = s0.x; v[t4][1] = s0.y; v[t4][2] = s0.z; v[t4][3] = s0.w;
                                v[t4][4] = s1.x; v[t4][5] = s1.y; v[t4][6] = s1.z; v[t4][7] = s1.w;
                            }
                        }
#pragma unroll
                        for (int t4 = 0; t4 < RB; ++t4) {
                            if (!ok[t4]) {
#pragma unroll
                                for (int i = 0; i < 8; ++i) v[t4][i] = 0.f;
                            }
                            store_units(a1, a2, row0 + t4, j - j_lo, v[t4]);
                        }
                    }
                }
                if (g + 1 < ngroups) {                        // more groups of this tile follow
                    fence_proxy_async();
                    a_publish();
                }
                }
            }
            fence_proxy_async();
            if (lwarp == 0) TC_TRACE(it, 0, 5);
            a_publish();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (a.csize > 1) cluster_sync_all();                          // no CTA leaves while its peer may still signal its barriers
    if (a.trace && tid == 0) {
        long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        a.trace[512 + 2 * blockIdx.x + 1] = t;
    }
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
    if (a.dynamic && tid == 0) {                                   // every CTA has drawn its last tile before it gets here
        __threadfence();
        if (atomicAdd(a.counter + 1, 1u) == gridDim.x - 1) {
            a.counter[0] = 0;
            a.counter[1] = 0;
            __threadfence();
        }
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
}  // namespace
long long *g_tc_trace = nullptr;   // device buffer of 8 tiles x 4 phases x 8 events (pab_tune_tc_trace); debugging aid; shared with vlad_tc.cu
namespace {
int g_tc_dynamic = 0;      // tiles drawn from a global counter (pab_tune_tensor_core bit 3 = 8 sets it): as fast as the static
                           // sequence + SM reservation of engine.forward_stream, without needing to know what else is running
int g_tc_cluster = 0;      // weight multicast across CTA pairs (pab_tune_tensor_core bit 2 sets it): measured slower on
                           // B200 — the modules are bound by the MMA <-> epilogue hand-offs, not by L2 -> SM weight traffic

// Shared-memory plan of one launch: layer-0 operand region, staging, weight stages, constant table.
// `layers` are the TENSOR-CORE layers only (the optional pre-layer is passed separately).
struct TcPlan { int a_region, nbuf, gchunks, n_stages, stage_bytes, pair, coff[MAX_LAYERS], pre_off; size_t misc, smem; };

int g_tc_nbuf = 1;         // layer-0 operand buffers; 2 = double-buffered where the region is small (measured: no gain, sa0 0.132 ms either way — the loaders are not what the narrow modules wait for); pab_tune_tensor_core bit 5 = 32 selects two

int g_tc_serial_epi = 0;   // experiment (pab_tune_tensor_core bit 6 = 64)
int g_tc_pair = 1;         // wide weight blocks: hi + lo plane in one 32-KB stage (pab_tune_tensor_core bit 4 = 16 clears it)

bool tc_plan(const pab_layer_t *layers, int n_layers, const pab_layer_t *pre, int mode, TcPlan *p) {
    const int planes = layers[0].w_lo ? 2 : 1;
    const long stg_bytes = mode == TC_SA ? STG_BYTES : STG_BYTES_FP;
    int ctab = 0;
    for (int l = 0; l < n_layers; ++l) {
        p->coff[l] = ctab;
        ctab += layers[l].c_out;
    }
    p->pre_off = ctab;
    if (pre) ctab += pre->c_in * pre->c_out + pre->c_out;
    // hi + lo plane of the gathered rows; inputs wider than 5 chunks go through a 4-chunk region in groups (needs the whole
    // layer-0 output in one accumulator pass and no loader-evaluated pre-layer)
    p->gchunks = layers[0].tc_k / KCH;
    if (p->gchunks > 5) {
        if (pre || layers[0].c_out > D_COLS) return false;
        p->gchunks = 4;
    }
    p->a_region = planes * p->gchunks * A_CHUNK;
    // a second operand buffer when the whole input is one group and two regions leave room for >= 4 weight stages
    const bool one_group = layers[0].tc_k / KCH <= 5;
    p->nbuf = (g_tc_nbuf == 2 && one_group && 2L * p->a_region + stg_bytes + 4 * (NBLK_MAX * 128) + 1024 + ctab * 4 <= 227L * 1024 &&
               p->a_region <= 64 * 1024) ? 2 : 1;
    p->misc = 384 + (size_t)ctab * 4 + 64;
    const long budget = 227L * 1024 - (long)p->a_region * p->nbuf - stg_bytes - (long)p->misc;
    bool wide = false;
    for (int l = 0; l < n_layers; ++l) wide = wide || layers[l].c_out > 64;
    p->pair = (g_tc_pair && wide && planes == 2 && budget / (2 * NBLK_MAX * 128) >= 2) ? 1 : 0;
    p->stage_bytes = p->pair ? 2 * NBLK_MAX * 128 : NBLK_MAX * 128;
    p->n_stages = (int)(budget / p->stage_bytes);
    if (p->n_stages > MAX_STAGES) p->n_stages = MAX_STAGES;
    p->smem = (size_t)p->a_region * p->nbuf + (size_t)p->n_stages * p->stage_bytes + stg_bytes + p->misc;
    return p->n_stages >= 2;
}

bool tc_layers_ok(const pab_layer_t *layers, int n_layers, bool first_is_module_input) {
    if (n_layers < 1 || n_layers > MAX_LAYERS) return false;
    for (int l = 0; l < n_layers; ++l) {
        const pab_layer_t &L = layers[l];
        if (!L.w_hi || L.tc_k <= 0 || L.tc_k % KCH) return false;
        if ((L.w_lo != nullptr) != (layers[0].w_lo != nullptr)) return false;     // one precision mode per module
        if (!(L.c_out == 32 || L.c_out == 64 || L.c_out % NBLK_MAX == 0) || L.c_out > 512) return false;
        if (l < n_layers - 1 && L.c_out > D_COLS) return false;       // the next operand must fit the TMEM planes
        const bool module_input = l == 0 && first_is_module_input;
        if (!module_input && (L.tc_k0 != 0 || L.tc_k < L.c_in)) return false;          // K may be zero-padded to 64
        if (l > 0 && (L.c_in != layers[l - 1].c_out || L.c_in % 16)) return false;
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
            tc_plan(layers, n_layers, nullptr, k_group > 0 ? TC_SA : TC_FP, &p))
            return 1;
    }
    if (allow_pre && n_layers >= 2 && layers[0].c_in <= 8 && layers[0].c_out % 16 == 0 && layers[0].c_out <= 64 &&
        tc_layers_ok(layers + 1, n_layers - 1, false) && tc_plan(layers + 1, n_layers - 1, &layers[0], k_group > 0 ? TC_SA : TC_FP, &p))
        return 2;
    return 0;
}

namespace {

int pab_tc_launch(int mode, long rows, int k_group, const pab_layer_t *all_layers, int n_all, int kind, TcArgs &a, cudaStream_t st) {
    const pab_layer_t *pre = kind == 2 ? &all_layers[0] : nullptr;
    const pab_layer_t *layers = kind == 2 ? all_layers + 1 : all_layers;
    const int n_layers = kind == 2 ? n_all - 1 : n_all;
    TcPlan p;
    if (!tc_plan(layers, n_layers, pre, mode, &p)) return PAB_EINVAL;
    a.serial_epi = g_tc_serial_epi;
    a.nbuf = p.nbuf; a.a_region = p.a_region; a.gchunks = p.gchunks; a.n_stages = p.n_stages; a.stage_bytes = p.stage_bytes; a.pair = p.pair;
    const long per_tile = mode == TC_SA ? (TM / k_group) : TM;
    if (rows >= (1L << 31)) return PAB_EINVAL;
    a.ntiles = (int)((rows + per_tile - 1) / per_tile);
    if (a.ntiles == 0) return 0;
    static int n_sm = 0;
    if (!n_sm) {
        int dev = 0;
        PAB_CUDA(cudaGetDevice(&dev));
        PAB_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    }
    // CTA pairs share the weight stream: each CTA fetches half of every weight block and TMA multicasts it to both, which
    // halves the L2 -> SM weight traffic (the bound of the 256-wide modules: every 128-row tile re-reads all the weights)
    a.csize = (g_tc_cluster && a.ntiles >= 4) ? 2 : 1;
    int grid = a.ntiles < n_sm ? a.ntiles : n_sm;
    if (g_tc_max_ctas > 0 && grid > g_tc_max_ctas) grid = g_tc_max_ctas;
    if (a.csize == 2) grid = grid / 2 * 2;
    a.iters = (a.ntiles + grid - 1) / grid;
    for (int l = 0; l < n_layers; ++l) {
        const pab_layer_t &L = layers[l];
        const int nbr = L.c_out < NBLK_MAX ? L.c_out : NBLK_MAX;
        // layer 0 of a module with extra (xyz) channels: the planes carry one more 64-column chunk with their weights
        const int kcols = L.tc_k + ((l == 0 && !pre && L.c_in > L.tc_k) ? KCH : 0);
        if (make_weight_map(&a.tm[l][0], L.w_hi, L.c_out, kcols, nbr / a.csize)) return PAB_EINVAL;
        if (L.w_lo && make_weight_map(&a.tm[l][1], L.w_lo, L.c_out, kcols, nbr / a.csize)) return PAB_EINVAL;
        a.shift[l] = L.shift; a.K[l] = L.tc_k; a.N[l] = L.c_out; a.relu[l] = L.relu; a.coff[l] = p.coff[l];
        // k-steps that carry data: the staged layer-0 operand spans whole 64-chunks (only the pre-layer's output is
        // narrower), the TMEM operand of later layers exactly c_in channels
        a.ksteps[l] = l == 0 ? (pre ? (pre->c_out + 15) / 16 : L.tc_k / 16) : L.c_in / 16;
    }
    a.n_layers = n_layers; a.mode = mode; a.rows = rows; a.trace = g_tc_trace;
    a.planes = layers[0].w_lo ? 2 : 1;
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
    // dynamic tile scheduling (not with CTA pairs: they run in lock-step): one counter pair per launch out of a small pool
    a.dynamic = 0; a.counter = nullptr;
    if (g_tc_dynamic && a.csize == 1 && a.ntiles > grid) {
        a.counter = pab_tile_counter_pair(st);                     // self-resetting (tile, finished) pair (api.cu); none left: static tiles
        a.dynamic = a.counter ? 1 : 0;
    }
    PAB_CUDA(cudaFuncSetAttribute(mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = p.smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = a.csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    PAB_CUDA(cudaLaunchKernelEx(&cfg, mlp_tc_kernel, a));
    PAB_LAUNCH_CHECK();
    return 0;
}

}  // namespace

PAB_API void pab_tune_tc_max_ctas(int n) { g_tc_max_ctas = n; }

PAB_API void pab_tune_tc_trace(void *device_buffer) { g_tc_trace = (long long *)device_buffer; }

PAB_API void pab_tune_tensor_core(int enable) {
    g_tc_enabled = enable & 1;
    g_tc_cluster = (enable & 4) != 0;
    g_tc_dynamic = (enable & 8) != 0;
    g_tc_pair = (enable & 16) == 0;
    g_tc_nbuf = (enable & 32) ? 2 : 1;
    g_tc_serial_epi = (enable & 64) != 0;
}

int pab_tc_sa(int kind, int b, int n, int m, int k, int nbr_stride, int c, const float *xyz, const float *feat, const int *center_idx,
              const int *nbr_idx, const pab_layer_t *layers, int n_layers, float *out, cudaStream_t st) {
    TcArgs a{};
    a.n = n; a.m = m; a.k = k; a.nbr_stride = nbr_stride; a.c = c; a.xyz = xyz; a.feat = feat; a.center_idx = center_idx;
    a.nbr_idx = nbr_idx; a.out = out;
    return pab_tc_launch(TC_SA, (long)b * m, k, layers, n_layers, kind, a, st);
}

int pab_tc_fp(int b, int n, int m, int c_known, int c_skip, const float *known_feat, const float *skip_feat, const int *idx,
              const float *weight, const int *row_order, long order_stride, const pab_layer_t *layers, int n_layers, float *out,
              cudaStream_t st) {
    TcArgs a{};
    a.row_order = row_order; a.order_stride = order_stride;
    a.n = n; a.m = m; a.c_known = c_known; a.c_skip = c_skip; a.known_feat = known_feat; a.skip_feat = skip_feat;
    a.idx3 = idx; a.w3 = weight; a.out = out;
    return pab_tc_launch(TC_FP, (long)b * n, 0, layers, n_layers, 1, a, st);
}
