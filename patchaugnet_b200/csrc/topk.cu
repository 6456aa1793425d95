// topk.cu — brute-force exact kNN in arbitrary dimension (sm_100a): the KNN_CUDA drop-in and the retrieval kNN.
//
// Reference (libs/KNN_CUDA/knn_cuda/csrc/cuda/knn.cu:232-269): kernel 1 writes the full nr x nq squared-distance
// matrix to global memory (400 MB at 10k x 10k), kernel 2 insertion-sorts every column in place with one thread
// per query, kernel 3 takes square roots.  Here nothing is materialised: a CTA of 4 warps owns 16 queries (4 per
// warp, held in shared memory), streams tiles of 32 reference points x 64 dims through shared memory, each lane
// accumulates the squared distance of ITS reference point to the warp's 4 queries with the reference's sequential
// fma chain over the dimension (ssd = fma(diff, diff, ssd), knn.cu:80-83), and the per-query sorted top-k lists
// live in registers across the warp exactly like knn.cu's lists in this repo (ballot filter + shuffle insertion).
// Ordering: ascending distance, ties to the lower reference index (the reference's insertion sort is stable,
// knn.cu:125-128, 149-152).
#include <math.h>
#include "common.cuh"

namespace {

constexpr int TK_WARPS = 4;
constexpr int TK_DCH = 64;    // dims per staged chunk
constexpr int TK_PTS = 32;    // reference points per tile
constexpr int TK_RS = 33;     // padded smem row stride of the reference tile

struct TopkArgs {
    const float *ref, *query;
    long ref_sd, ref_sp, q_sd, q_sp;   // strides (in floats) along dim / along point
    int nr, nq, dim, k;
    float *dist;                       // element (q, j) at dist[q*d_sq + j*d_sk]
    long d_sq, d_sk;
    void *ind;
    int ind64, ind_base;               // int64 (1-based for KNN_CUDA) or int32
    const unsigned *mask;              // optional (nq, mask_words) bit mask of the reference points each query may retrieve
    long mask_words;
    int nsplit, slice;                 // the reference set is cut into nsplit slices of `slice` points (multiple of 32), one per blockIdx.y:
    int raw;                           // split runs write SQUARED distances of their partial lists at (q, split, j); a merge kernel finishes
};

template <int QPW, int KPL>
__global__ void __launch_bounds__(TK_WARPS * 32) topk_kernel(const TopkArgs a) {
    extern __shared__ __align__(16) float smem[];
    constexpr int QB = TK_WARPS * QPW;           // queries per block
    float *qs = smem;                            // [dim][QB]
    float *rs = smem + (size_t)a.dim * QB;       // [TK_DCH][TK_RS]
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int q0 = blockIdx.x * QB;

    for (int e = t; e < a.dim * QB; e += TK_WARPS * 32) {
        int d, qq;
        if (a.q_sd == 1) { d = e % a.dim; qq = e / a.dim; } else { qq = e % QB; d = e / QB; }
        const int q = q0 + qq;
        qs[d * QB + qq] = q < a.nq ? __ldg(a.query + d * a.q_sd + q * a.q_sp) : 0.f;
    }

    float ld[QPW][KPL];
    int li[QPW][KPL];
    float tau[QPW];
#pragma unroll
    for (int qi = 0; qi < QPW; ++qi) {
        tau[qi] = INFINITY;
#pragma unroll
        for (int s = 0; s < KPL; ++s) { ld[qi][s] = INFINITY; li[qi][s] = 0; }
    }
    const int ks = (a.k - 1) >> 5, kl = (a.k - 1) & 31;

    const int p_begin = blockIdx.y * a.slice, p_end = min(a.nr, p_begin + a.slice);
    for (int p0 = p_begin; p0 < p_end; p0 += TK_PTS) {
        float ssd[QPW];
#pragma unroll
        for (int qi = 0; qi < QPW; ++qi) ssd[qi] = 0.f;
        for (int d0 = 0; d0 < a.dim; d0 += TK_DCH) {
            const int dn = min(TK_DCH, a.dim - d0);
            __syncthreads();
            for (int e = t; e < TK_DCH * TK_PTS; e += TK_WARPS * 32) {
                int d, pp;
                if (a.ref_sd == 1) { d = e % TK_DCH; pp = e / TK_DCH; } else { pp = e % TK_PTS; d = e / TK_PTS; }
                const int p = p0 + pp;
                if (d < dn) rs[d * TK_RS + pp] = p < a.nr ? __ldg(a.ref + (d0 + d) * a.ref_sd + p * a.ref_sp) : 0.f;
            }
            __syncthreads();
            const float *qrow = qs + (size_t)d0 * QB + warp * QPW;
#pragma unroll 4
            for (int d = 0; d < dn; ++d) {
                const float r = rs[d * TK_RS + lane];
#pragma unroll
                for (int qi = 0; qi < QPW; ++qi) {
                    const float diff = __fsub_rn(r, qrow[d * QB + qi]);   // knn.cu:81  tmp = A - B (ref - query)
                    ssd[qi] = __fmaf_rn(diff, diff, ssd[qi]);
                }
            }
        }
        const bool pvalid = p0 + lane < a.nr;
#pragma unroll
        for (int qi = 0; qi < QPW; ++qi) {
            float d = pvalid ? ssd[qi] : INFINITY;
            if (a.mask) {                // ragged candidate sets (hard-negative mining): bit p of the query's row = allowed
                const int q = q0 + warp * QPW + qi;
                const unsigned w = q < a.nq ? __ldg(a.mask + (size_t)q * a.mask_words + (p0 >> 5)) : 0u;
                if (!((w >> lane) & 1u)) d = INFINITY;
            }
            unsigned hit = __ballot_sync(0xffffffffu, d < tau[qi]);
            while (hit) {
                const int src = __ffs(hit) - 1;
                hit &= hit - 1;
                const float cd = __shfl_sync(0xffffffffu, d, src);
                if (!(cd < tau[qi])) continue;
                const int ci = p0 + src;
                int pos = 0;
#pragma unroll
                for (int s = 0; s < KPL; ++s) pos += __popc(__ballot_sync(0xffffffffu, ld[qi][s] <= cd));
#pragma unroll
                for (int s = KPL - 1; s >= 0; --s) {
                    float ud = __shfl_up_sync(0xffffffffu, ld[qi][s], 1);
                    int ui = __shfl_up_sync(0xffffffffu, li[qi][s], 1);
                    if (s > 0) {
                        const float pd = __shfl_sync(0xffffffffu, ld[qi][s - 1], 31);
                        const int pi = __shfl_sync(0xffffffffu, li[qi][s - 1], 31);
                        if (lane == 0) { ud = pd; ui = pi; }
                    }
                    const int me = s * 32 + lane;
                    if (me == pos) { ld[qi][s] = cd; li[qi][s] = ci; }
                    else if (me > pos) { ld[qi][s] = ud; li[qi][s] = ui; }
                }
#pragma unroll
                for (int s = 0; s < KPL; ++s)
                    if (s == ks) tau[qi] = __shfl_sync(0xffffffffu, ld[qi][s], kl);
            }
        }
    }
#pragma unroll
    for (int qi = 0; qi < QPW; ++qi) {
        const int q = q0 + warp * QPW + qi;
        if (q >= a.nq) continue;
#pragma unroll
        for (int s = 0; s < KPL; ++s) {
            const int j = s * 32 + lane;
            if (j < a.k) {
                const long o = q * a.d_sq + (long)blockIdx.y * a.k * a.d_sk + j * a.d_sk;
                a.dist[o] = a.raw ? ld[qi][s] : __fsqrt_rn(ld[qi][s]);
                if (a.ind64) ((long long *)a.ind)[o] = (long long)li[qi][s] + a.ind_base;
                else ((int *)a.ind)[o] = li[qi][s] + a.ind_base;
            }
        }
    }
}

template <int QPW, int KPL>
int launch_topk(const TopkArgs &a, cudaStream_t st) {
    constexpr int QB = TK_WARPS * QPW;
    const size_t smem = sizeof(float) * ((size_t)a.dim * QB + TK_DCH * TK_RS);
    if (smem > 200 * 1024) return PAB_EINVAL;
    if (smem > 48 * 1024) PAB_CUDA(cudaFuncSetAttribute(topk_kernel<QPW, KPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    topk_kernel<QPW, KPL><<<dim3(pab_divup(a.nq, QB), a.nsplit), TK_WARPS * 32, smem, st>>>(a);
    PAB_LAUNCH_CHECK();
    return 0;
}

int run_topk(TopkArgs a, cudaStream_t st) {
    if (a.nsplit < 1) { a.nsplit = 1; a.slice = (a.nr + 31) / 32 * 32; a.raw = 0; }
    if (a.nr <= 0 || a.nq < 0 || a.dim <= 0 || a.k <= 0 || a.k > 1024 || a.k > a.nr) return PAB_EINVAL;
    if (a.nq == 0) return 0;
    if (a.k <= 32) return launch_topk<4, 1>(a, st);
    if (a.k <= 64) return launch_topk<4, 2>(a, st);
    if (a.k <= 128) return launch_topk<4, 4>(a, st);
    if (a.k <= 256) return launch_topk<2, 8>(a, st);
    if (a.k <= 512) return launch_topk<1, 16>(a, st);
    return launch_topk<1, 32>(a, st);
}

// partial lists (nq, nsplit, k) of (squared distance, index), each ascending -> the k best of their union, ascending distance,
// ties to the lower index; one CTA per query sorts the nsplit*k candidates as 64-bit (distance bits, index) keys in shared memory
__global__ void __launch_bounds__(256) topk_merge_kernel(int k, int total, int P, const float *__restrict__ pd, const int *__restrict__ pi,
                                                         float *__restrict__ dist, int *__restrict__ ind) {
    extern __shared__ unsigned long long mkeys[];
    const int q = blockIdx.x, t = threadIdx.x;
    for (int i = t; i < P; i += 256) {
        unsigned long long key = ~0ull;
        if (i < total) key = ((unsigned long long)__float_as_uint(pd[(size_t)q * total + i]) << 32) | (unsigned)pi[(size_t)q * total + i];
        mkeys[i] = key;                                   // squared distances are >= 0 (or +inf): their bit patterns order like the values
    }
    __syncthreads();
    for (int size = 2; size <= P; size <<= 1)
        for (int j = size >> 1; j > 0; j >>= 1) {
            for (int i = t; i < (P >> 1); i += 256) {
                const int a0 = 2 * i - (i & (j - 1));
                const unsigned long long ka = mkeys[a0], kb = mkeys[a0 + j];
                const bool up = (a0 & size) == 0;
                if ((ka > kb) == up) { mkeys[a0] = kb; mkeys[a0 + j] = ka; }
            }
            __syncthreads();
        }
    for (int j = t; j < k; j += 256) {
        dist[(size_t)q * k + j] = __fsqrt_rn(__uint_as_float((unsigned)(mkeys[j] >> 32)));
        ind[(size_t)q * k + j] = (int)(mkeys[j] & 0xFFFFFFFFu);
    }
}

}  // namespace

PAB_API int pab_retrieval_topk(const float *db, int ndb, const float *q, int nq, int dim, int k, float *dist, int *ind, pab_stream_t s);

// Retrieval top-k with the DATABASE split over the grid as well: with few queries (a rank's 250-query shard of configs[3] is 16
// query blocks) a grid over queries alone leaves most SMs idle, so the database is cut into slices, every (query block, slice)
// CTA keeps a partial top-k, and a merge kernel takes the k best of each query's partial lists.  Same results as
// pab_retrieval_topk (exact, ascending distance, ties to the lower index).  workspace >= pab_retrieval_topk_workspace_bytes(nq, k).
PAB_API size_t pab_retrieval_topk_workspace_bytes(int nq, int k) { return (size_t)nq * 32 * (size_t)k * 8; }

PAB_API int pab_retrieval_topk_split(const float *db, int ndb, const float *q, int nq, int dim, int k, float *dist, int *ind,
                                     void *workspace, pab_stream_t s) {
    if (ndb <= 0 || nq < 0 || k <= 0 || k > 128 || k > ndb || !workspace) return PAB_EINVAL;
    if (nq == 0) return 0;
    cudaStream_t st = (cudaStream_t)s;
    const int qblocks = pab_divup(nq, TK_WARPS * 4);
    int nsplit = (2 * 148 + qblocks - 1) / qblocks;                     // about two CTAs per SM
    if (nsplit > 32) nsplit = 32;
    while (nsplit > 1 && ndb / nsplit < 4 * k) --nsplit;                // slices much longer than the lists they feed
    if (nsplit <= 1) return pab_retrieval_topk(db, ndb, q, nq, dim, k, dist, ind, s);
    TopkArgs a{};
    a.ref = db; a.query = q; a.ref_sd = 1; a.ref_sp = dim; a.q_sd = 1; a.q_sp = dim;
    a.nr = ndb; a.nq = nq; a.dim = dim; a.k = k;
    float *pd = (float *)workspace;
    int *pi = (int *)(pd + (size_t)nq * nsplit * k);
    a.dist = pd; a.ind = pi; a.d_sq = (long)nsplit * k; a.d_sk = 1; a.ind64 = 0; a.ind_base = 0;
    a.nsplit = nsplit; a.slice = ((ndb + nsplit - 1) / nsplit + 31) / 32 * 32; a.raw = 1;
    const int rc = run_topk(a, st);
    if (rc) return rc;
    const int total = nsplit * k;
    int P = 2;
    while (P < total) P <<= 1;
    topk_merge_kernel<<<nq, 256, (size_t)P * 8, st>>>(k, total, P, pd, pi, dist, ind);
    PAB_LAUNCH_CHECK();
    return 0;
}

PAB_API int pab_knn(const float *ref, int nr, const float *query, int nq, int dim, int k, float *dist, int64_t *ind, pab_stream_t s) {
    TopkArgs a{};
    a.ref = ref; a.query = query; a.ref_sd = nr; a.ref_sp = 1; a.q_sd = nq; a.q_sp = 1;
    a.nr = nr; a.nq = nq; a.dim = dim; a.k = k;
    a.dist = dist; a.d_sq = 1; a.d_sk = nq; a.ind = ind; a.ind64 = 1; a.ind_base = 1;
    return run_topk(a, (cudaStream_t)s);
}

// Same search restricted, per query, to the reference points whose bit is set in mask (nq rows of ceil(ndb/32) words): one
// launch ranks every query against its own ragged candidate set.  A query with fewer than k candidates gets +inf distances
// (and index 0) in the unused slots.
PAB_API int pab_retrieval_topk_masked(const float *db, int ndb, const float *q, int nq, int dim, int k, const unsigned *mask,
                                      float *dist, int *ind, pab_stream_t s) {
    if (!mask) return PAB_EINVAL;
    TopkArgs a{};
    a.ref = db; a.query = q; a.ref_sd = 1; a.ref_sp = dim; a.q_sd = 1; a.q_sp = dim;
    a.nr = ndb; a.nq = nq; a.dim = dim; a.k = k;
    a.dist = dist; a.d_sq = k; a.d_sk = 1; a.ind = ind; a.ind64 = 0; a.ind_base = 0;
    a.mask = mask; a.mask_words = (ndb + 31) / 32;
    return run_topk(a, (cudaStream_t)s);
}

PAB_API int pab_retrieval_topk(const float *db, int ndb, const float *q, int nq, int dim, int k, float *dist, int *ind, pab_stream_t s) {
    TopkArgs a{};
    a.ref = db; a.query = q; a.ref_sd = 1; a.ref_sp = dim; a.q_sd = 1; a.q_sp = dim;
    a.nr = ndb; a.nq = nq; a.dim = dim; a.k = k;
    a.dist = dist; a.d_sq = k; a.d_sk = 1; a.ind = ind; a.ind64 = 0; a.ind_base = 0;
    return run_topk(a, (cudaStream_t)s);
}
