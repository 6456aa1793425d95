// patch_pairs.cu — patch-feature-contrast (a2b) triplet selection on the device.
//
// Replaces the per-pair numpy loop of place_recognition/train_place_recognition.py:310-367: for every overlap entry
// (idx1, near_indices2, far list) of a (query cloud m, positive cloud n) pair
//     idx1     = first position of entry.idx1 among m's level-0 centre indices            (np.where(...)[0][0], :338)
//     pos_idx2 = ascending positions of n's centres whose point index is in near_indices2  (np.where(np.isin), :344)
//     neg_idx2 = ascending positions of n's centres whose point index is in the far list   (:360)
// entries with no idx1 / no positive / no negative are skipped (:339-340, 345-346, 361-362); every positive yields one
// triplet (idx1, pos, neg) with neg drawn with replacement from neg_idx2 (np.random.choice, :364).  The draw uses a
// counter-based generator keyed by (seed, pair, entry, j) so a host restatement reproduces it (the parity test does).
//
// One CTA per cloud pair, one warp per entry: centres of both clouds staged in shared memory, membership by a scan of
// the (short) index list per lane, ordered compaction by ballot; entry counts -> block-wide exclusive scan -> a second
// pass writes the triplets in the reference's order (entry order, then ascending position).
#include "common.cuh"

namespace {

constexpr int PP_THREADS = 256;
constexpr int PP_WARPS = PP_THREADS / 32;
constexpr int PP_MAX_M = 1024;        // centres per cloud (SAMPLING[0] of the configuration)
constexpr int PP_MAX_ENTRIES = 1024;  // entries per pair handled by one CTA (the reference samples at most 500, :331-332)

struct PpArgs {
    int n_pairs, M, max_out;
    const int *centers, *pair_m, *pair_n, *entry_ptr, *entry_idx1, *near_ptr, *near_val, *far_ptr, *far_val;
    unsigned long long seed;
    int *out_idx1, *out_pos, *out_neg, *out_count;
};

__device__ __forceinline__ unsigned long long pp_mix(unsigned long long z) {     // splitmix64 finaliser
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__device__ __forceinline__ bool pp_member(int v, const int *list, int len) {
    bool hit = false;
    for (int i = 0; i < len; ++i) hit |= (__ldg(list + i) == v);
    return hit;
}

__global__ void __launch_bounds__(PP_THREADS) patch_triplets_kernel(const PpArgs a) {
    __shared__ int cm[PP_MAX_M], cn[PP_MAX_M];
    __shared__ int cnt[PP_MAX_ENTRIES];                 // triplets per entry, then exclusive offsets
    __shared__ short negbuf[PP_WARPS][PP_MAX_M];
    __shared__ int wsum[PP_WARPS];
    const int pair = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, M = a.M;
    const int e0 = a.entry_ptr[pair], ne = a.entry_ptr[pair + 1] - e0;
    const int *gm = a.centers + (long)a.pair_m[pair] * M, *gn = a.centers + (long)a.pair_n[pair] * M;
    for (int i = tid; i < M; i += PP_THREADS) { cm[i] = __ldg(gm + i); cn[i] = __ldg(gn + i); }
    __syncthreads();

    // one entry: first position of idx1 in cm, number of positives / negatives in cn (all lanes return the same values)
    auto scan_entry = [&](int e, int &p1, int &npos, int &nneg) {
        const int idx1 = __ldg(a.entry_idx1 + e);
        const int *near = a.near_val + __ldg(a.near_ptr + e), *far = a.far_val + __ldg(a.far_ptr + e);
        const int nl = __ldg(a.near_ptr + e + 1) - __ldg(a.near_ptr + e), fl = __ldg(a.far_ptr + e + 1) - __ldg(a.far_ptr + e);
        p1 = -1; npos = 0; nneg = 0;
        for (int base = 0; base < M; base += 32) {
            const int i = base + lane;
            const bool in = i < M;
            const unsigned hit1 = __ballot_sync(0xffffffffu, in && cm[i] == idx1);
            if (p1 < 0 && hit1) p1 = base + __ffs(hit1) - 1;
            const int v = in ? cn[i] : -1;
            npos += __popc(__ballot_sync(0xffffffffu, in && pp_member(v, near, nl)));
            nneg += __popc(__ballot_sync(0xffffffffu, in && pp_member(v, far, fl)));
        }
    };

    for (int el = warp; el < ne; el += PP_WARPS) {
        int p1, npos, nneg;
        scan_entry(e0 + el, p1, npos, nneg);
        if (lane == 0) cnt[el] = (p1 >= 0 && npos > 0 && nneg > 0) ? npos : 0;
    }
    __syncthreads();
    // block-wide exclusive scan of cnt[0..ne) (ne <= 1024: four entries per thread)
    {
        int v[4], s = 0;
        for (int j = 0; j < 4; ++j) { const int i = tid * 4 + j; v[j] = i < ne ? cnt[i] : 0; s += v[j]; }
        int inc = s;
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        int woff = 0;
        for (int w = 0; w < warp; ++w) woff += wsum[w];
        int run = woff + inc - s;
        for (int j = 0; j < 4; ++j) { const int i = tid * 4 + j; if (i < ne) cnt[i] = run; run += v[j]; }
        if (tid == PP_THREADS - 1) a.out_count[pair] = run;          // total triplets of the pair (may exceed max_out: host checks)
    }
    __syncthreads();

    for (int el = warp; el < ne; el += PP_WARPS) {
        const int e = e0 + el;
        int p1, npos, nneg;
        scan_entry(e, p1, npos, nneg);
        if (!(p1 >= 0 && npos > 0 && nneg > 0)) continue;
        const int *near = a.near_val + __ldg(a.near_ptr + e), *far = a.far_val + __ldg(a.far_ptr + e);
        const int nl = __ldg(a.near_ptr + e + 1) - __ldg(a.near_ptr + e), fl = __ldg(a.far_ptr + e + 1) - __ldg(a.far_ptr + e);
        // negatives' positions, ascending, into this warp's buffer
        int nn = 0;
        for (int base = 0; base < M; base += 32) {
            const int i = base + lane;
            const bool hit = i < M && pp_member(cn[i], far, fl);
            const unsigned b = __ballot_sync(0xffffffffu, hit);
            if (hit) negbuf[warp][nn + __popc(b & ((1u << lane) - 1))] = (short)i;
            nn += __popc(b);
        }
        __syncwarp();
        const long obase = (long)pair * a.max_out;
        int j0 = 0;
        for (int base = 0; base < M; base += 32) {
            const int i = base + lane;
            const bool hit = i < M && pp_member(cn[i], near, nl);
            const unsigned b = __ballot_sync(0xffffffffu, hit);
            if (hit) {
                const int j = j0 + __popc(b & ((1u << lane) - 1));
                const int o = cnt[el] + j;
                if (o < a.max_out) {
                    const unsigned long long h = pp_mix(pp_mix(a.seed ^ ((unsigned long long)pair << 40)) ^ (((unsigned long long)el << 20) | (unsigned)j));
                    const int r = (int)(((h >> 32) * (unsigned long long)nn) >> 32);
                    a.out_idx1[obase + o] = p1;
                    a.out_pos[obase + o] = i;
                    a.out_neg[obase + o] = negbuf[warp][r];
                }
            }
            j0 += __popc(b);
        }
        __syncwarp();
    }
}

}  // namespace

PAB_API int pab_patch_triplets(int n_pairs, int M, const int *centers, const int *pair_m, const int *pair_n, const int *entry_ptr,
                               int max_entries_per_pair, const int *entry_idx1, const int *near_ptr, const int *near_val,
                               const int *far_ptr, const int *far_val, unsigned long long seed, int max_out, int *out_idx1,
                               int *out_pos, int *out_neg, int *out_count, pab_stream_t s) {
    if (n_pairs < 0 || M <= 0 || M > PP_MAX_M || max_entries_per_pair > PP_MAX_ENTRIES || max_out < 0) return PAB_EINVAL;
    if (n_pairs == 0) return 0;
    PpArgs a{n_pairs, M, max_out, centers, pair_m, pair_n, entry_ptr, entry_idx1, near_ptr, near_val, far_ptr, far_val, seed,
             out_idx1, out_pos, out_neg, out_count};
    patch_triplets_kernel<<<n_pairs, PP_THREADS, 0, (cudaStream_t)s>>>(a);
    PAB_LAUNCH_CHECK();
    return 0;
}
