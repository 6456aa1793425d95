// knn.cu — exact kNN / 3-NN / ball query over 3-D points for sm_100a.
//
// knnquery: replaces knnquery_cuda_kernel (libs/pointops/src/knnquery/knnquery_cuda_kernel.cu:6-50), which runs
// one thread per query with a 2400-byte local-memory insertion list.  Here one WARP owns a query: reference
// points are staged in shared memory (SoA, coalesced loads), each lane evaluates one reference per step,
// a warp ballot against the current k-th best distance filters candidates (after warm-up almost every step is
// rejected by a single compare), and the running k-best SET lives unsorted in registers spread across the lanes:
// an insertion overwrites the current lexicographic maximum (found with two redux.sync) and re-derives the
// threshold with one more; the set is sorted once at the end with a shuffle bitonic network on 64-bit
// (distance, index) keys.  Ordering is the reference's: ascending squared
// distance (contracted fp32 order, dx = query - ref), ties to the lower index (strict '<' on insertion, refs are
// visited in index order).
//
// nearestneighbor (3-NN): replaces nearestneighbor_cuda_kernel_fast (interpolation_cuda_kernel.cu:134-176); one
// thread per unknown point, known points in shared memory, same strict-'<' cascade.
#include <math.h>
#include "common.cuh"

namespace {

constexpr int KNN_CHUNK = 4096;   // reference points staged per pass (48 KB SoA)
constexpr int KNN_WARPS = 16;

// 64-bit sort key: squared distance bits (>= 0, so they order as unsigned) then index
__device__ __forceinline__ unsigned long long knn_key(float d, int i) {
    return ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)i;
}

// one bitonic sort of the 32*KPL (distance, index) keys held one per (slot, lane) — dead slots sort last — then the
// first k are written out in ascending order
template <int KPL>
__device__ __forceinline__ void knn_sort_store(const float (&ld)[KPL], const int (&li)[KPL], int k, int lane, int *o, float *od) {
    unsigned long long key[KPL];
#pragma unroll
    for (int s = 0; s < KPL; ++s) key[s] = (s * 32 + lane < k) ? knn_key(ld[s], li[s]) : ~0ull;
#pragma unroll
    for (int size = 2; size <= 32 * KPL; size <<= 1) {
#pragma unroll
        for (int j = size >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (int s = 0; s < KPL; ++s) {
                const int e = s * 32 + lane;
                const bool up = (e & size) == 0;                       // ascending block
                if (j >= 32) {
                    const int sp = s ^ (j >> 5);
                    if (sp > s) {                                       // handle each register pair once
                        const unsigned long long a = key[s], b = key[sp];
                        const bool swap = (a > b) == up;
                        key[s] = swap ? b : a;
                        key[sp] = swap ? a : b;
                    }
                } else {
                    const unsigned long long other = __shfl_xor_sync(0xffffffffu, key[s], j);
                    const bool lower = (lane & j) == 0;                 // this lane holds the lower position of the pair
                    const bool take_min = lower == up;
                    key[s] = ((key[s] < other) == take_min) ? key[s] : other;
                }
            }
        }
    }
#pragma unroll
    for (int s = 0; s < KPL; ++s) {
        const int me = s * 32 + lane;
        if (me < k) {
            o[me] = (int)(unsigned)(key[s] & 0xffffffffull);
            if (od) od[me] = __uint_as_float((unsigned)(key[s] >> 32));
        }
    }
}

template <int KPL>
__global__ void __launch_bounds__(KNN_WARPS * 32)
knn_kernel(int n, int m, int k, const float *__restrict__ xyz, const float *__restrict__ new_xyz,
           int *__restrict__ idx, float *__restrict__ dist2) {
    __shared__ float sm[3 * KNN_CHUNK];          // SoA: x at +0, y at +KNN_CHUNK, z at +2*KNN_CHUNK (constant offsets)
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int cloud = blockIdx.y;
    const int q = blockIdx.x * KNN_WARPS + warp;
    const bool active = q < m;
    const float *p = xyz + (size_t)cloud * n * 3;

    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (active) {
        const float *c = new_xyz + ((size_t)cloud * m + q) * 3;
        qx = __ldg(c); qy = __ldg(c + 1); qz = __ldg(c + 2);
    }
    // The running k-best SET, unsorted, one element per (slot, lane): element e = 32*s + lane is live iff e < k.
    // Live slots start at (+inf, 0) like the reference's best[]=1e40 / besti[]=0 (knnquery_cuda_kernel.cu:23-26);
    // dead slots hold -1 so they can never be the maximum.  tau = max live distance = the k-th best so far.
    float ld[KPL];
    int li[KPL];
#pragma unroll
    for (int s = 0; s < KPL; ++s) { ld[s] = (s * 32 + lane < k) ? INFINITY : -1.f; li[s] = 0; }
    float tau = INFINITY;

    // candidates flagged in `hit` (lane order = index order) against the running set
    auto absorb = [&](unsigned hit, float d, int first_index) {
        while (hit) {
            const int src = __ffs(hit) - 1;
            hit &= hit - 1;
            const float cd = __shfl_sync(0xffffffffu, d, src);
            if (!(cd < tau)) continue;      // tau shrank since the ballot (warp-uniform).  Strict '<': refs are visited in
                                            // index order, so an equal distance with a higher index never displaces
            const int ci = first_index + src;
            // evict the lexicographic maximum (distance == tau, then the largest index)
            const int tb = __float_as_int(tau);
            int mi = -1, ms = 0;
#pragma unroll
            for (int s = 0; s < KPL; ++s)
                if (__float_as_int(ld[s]) == tb && li[s] > mi) { mi = li[s]; ms = s; }
            const int top = __reduce_max_sync(0xffffffffu, mi);
            const unsigned vb = __ballot_sync(0xffffffffu, mi == top);          // unfilled (+inf, 0) slots tie: lowest lane
            if (lane == __ffs(vb) - 1) {
#pragma unroll
                for (int s = 0; s < KPL; ++s)
                    if (s == ms) { ld[s] = cd; li[s] = ci; }
            }
            float md = ld[0];
#pragma unroll
            for (int s = 1; s < KPL; ++s) md = fmaxf(md, ld[s]);
            tau = __int_as_float(__reduce_max_sync(0xffffffffu, __float_as_int(md)));
        }
    };

    for (int base0 = 0; base0 < n; base0 += KNN_CHUNK) {
        const int cnt = min(KNN_CHUNK, n - base0);
        const int cnt64 = (cnt + 63) & ~63;       // padded with +inf coordinates: their distance is +inf, never < tau
        __syncthreads();
        for (int i = t; i < cnt64; i += KNN_WARPS * 32) {
            float x = INFINITY, y = INFINITY, z = INFINITY;
            if (i < cnt) {
                const float *src = p + (size_t)(base0 + i) * 3;
                x = __ldg(src); y = __ldg(src + 1); z = __ldg(src + 2);
            }
            sm[i] = x; sm[KNN_CHUNK + i] = y; sm[2 * KNN_CHUNK + i] = z;
        }
        __syncthreads();
        if (!active) continue;
        const float *sp = sm + lane;
        for (int base = 0; base < cnt64; base += 64) {
            const float d0 = ref_sqdist(qx, qy, qz, sp[base], sp[KNN_CHUNK + base], sp[2 * KNN_CHUNK + base]);
            const float d1 = ref_sqdist(qx, qy, qz, sp[base + 32], sp[KNN_CHUNK + base + 32], sp[2 * KNN_CHUNK + base + 32]);
            const unsigned h0 = __ballot_sync(0xffffffffu, d0 < tau);
            if (h0) absorb(h0, d0, base0 + base);
            const unsigned h1 = __ballot_sync(0xffffffffu, d1 < tau);
            if (h1) absorb(h1, d1, base0 + base + 32);
        }
    }
    if (!active) return;
    knn_sort_store<KPL>(ld, li, k, lane, idx + ((size_t)cloud * m + q) * k, dist2 ? dist2 + ((size_t)cloud * m + q) * k : nullptr);
}

// ---- spatial index: Morton-sorted copy of a cloud in 64-point chunks with bounding boxes -----------------------------
//
// The brute-force kernel above evaluates every reference for every query and — worse — its running k-best set keeps
// being displaced by late arrivals (~k ln(n/k) insertions per query).  With the references sorted along a Morton curve,
// every 64-point chunk is spatially compact; a query then visits chunks in order of their box distance (a lower bound
// of every point distance inside) and stops at the first chunk whose bound exceeds the current k-th best: typically
// 10-15 of 64 chunks at n = 4096, k = 40, with the set converging after the first few.  Results are IDENTICAL to the
// brute-force scan: the set is the k smallest (distance, original index) keys either way, the distance is the same
// fp32 expression, and the bound is evaluated with the same expression on clamped differences, so by monotonicity of
// every rounding step it never exceeds the distance of a point inside the box.
//
// Index layout per cloud (floats): x[npad] y[npad] z[npad] | perm[npad] (int: original index) | box[nch][8]
// (lo.xyz, hi.xyz, 2 pad), npad = n rounded up to 64 (pads: +inf coordinates), nch = npad / 64.
constexpr int IDX_CHUNK = 64;
constexpr int IDX_MAX_N = 8192;

__host__ __device__ inline int idx_npad(int n) { return (n + IDX_CHUNK - 1) / IDX_CHUNK * IDX_CHUNK; }
__host__ __device__ inline size_t idx_stride(int n) { return (size_t)4 * idx_npad(n) + (size_t)8 * (idx_npad(n) / IDX_CHUNK); }

__device__ __forceinline__ unsigned morton_spread(unsigned x) {      // 10 bits -> every third bit
    x = (x | (x << 16)) & 0x030000FFu;
    x = (x | (x << 8)) & 0x0300F00Fu;
    x = (x | (x << 4)) & 0x030C30C3u;
    x = (x | (x << 2)) & 0x09249249u;
    return x;
}

// one CTA per cloud: bounding box -> 30-bit Morton codes -> bitonic sort of (code, index) in shared memory -> sorted SoA
// copy, permutation and chunk boxes
__global__ void __launch_bounds__(1024) knn_index_kernel(int n, int P, const float *__restrict__ xyz, float *__restrict__ index) {
    extern __shared__ unsigned long long skey[];            // P keys
    __shared__ float red[6][32];
    __shared__ float bb[6];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5, cloud = blockIdx.x;
    const float *p = xyz + (size_t)cloud * n * 3;
    const int npad = idx_npad(n), nch = npad / IDX_CHUNK;
    float *ix = index + (size_t)cloud * idx_stride(n);
    float *iy = ix + npad, *iz = iy + npad;
    int *iperm = reinterpret_cast<int *>(iz + npad);
    float *ibox = reinterpret_cast<float *>(iperm + npad);

    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = t; i < n; i += 1024)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = __ldg(p + (size_t)i * 3 + c);
            if (fabsf(v) < INFINITY) { lo[c] = fminf(lo[c], v); hi[c] = fmaxf(hi[c], v); }   // finite values only
        }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
            hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
        }
        if (lane == 0) { red[c][warp] = lo[c]; red[3 + c][warp] = hi[c]; }
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float l = red[c][lane], h = red[3 + c][lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                l = fminf(l, __shfl_xor_sync(0xffffffffu, l, o));
                h = fmaxf(h, __shfl_xor_sync(0xffffffffu, h, o));
            }
            if (lane == 0) { bb[c] = l; bb[3 + c] = h; }
        }
    }
    __syncthreads();
    float sc[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float ext = bb[3 + c] - bb[c];
        sc[c] = (ext > 0.f && ext < INFINITY) ? 1023.5f / ext : 0.f;
    }
    for (int i = t; i < P; i += 1024) {
        unsigned long long key = ~0ull;
        if (i < n) {
            unsigned code = 0;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float v = __ldg(p + (size_t)i * 3 + c);
                float f = (v - bb[c]) * sc[c];
                f = f > 0.f ? f : 0.f;                       // also maps NaN to cell 0
                const unsigned cell = f < 1023.f ? (unsigned)f : 1023u;
                code |= morton_spread(cell) << c;
            }
            key = ((unsigned long long)code << 13) | (unsigned)i;
        }
        skey[i] = key;
    }
    __syncthreads();
    for (int size = 2; size <= P; size <<= 1)
        for (int j = size >> 1; j > 0; j >>= 1) {
            for (int i = t; i < (P >> 1); i += 1024) {
                const int a = 2 * i - (i & (j - 1));         // lower element of the pair
                const unsigned long long ka = skey[a], kb = skey[a + j];
                const bool up = (a & size) == 0;
                if ((ka > kb) == up) { skey[a] = kb; skey[a + j] = ka; }
            }
            __syncthreads();
        }
    for (int i = t; i < npad; i += 1024) {
        float x = INFINITY, y = INFINITY, z = INFINITY;
        int oi = 0;
        if (i < n) {
            oi = (int)(skey[i] & 8191ull);
            x = __ldg(p + (size_t)oi * 3); y = __ldg(p + (size_t)oi * 3 + 1); z = __ldg(p + (size_t)oi * 3 + 2);
        }
        ix[i] = x; iy[i] = y; iz[i] = z; iperm[i] = oi;
    }
    // chunk boxes: one warp per chunk, two points per lane (pads excluded: an empty chunk gets lo = +inf, hi = -inf)
    for (int c = warp; c < nch; c += 32) {
        float l[3] = {INFINITY, INFINITY, INFINITY}, h[3] = {-INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int i = c * IDX_CHUNK + u * 32 + lane;
            if (i < n) {
                const int oi = (int)(skey[i] & 8191ull);
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const float v = __ldg(p + (size_t)oi * 3 + d);
                    l[d] = fminf(l[d], v); h[d] = fmaxf(h[d], v);     // NaN coordinates are ignored (their distance is NaN: never selected)
                }
            }
        }
#pragma unroll
        for (int d = 0; d < 3; ++d)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                l[d] = fminf(l[d], __shfl_xor_sync(0xffffffffu, l[d], o));
                h[d] = fmaxf(h[d], __shfl_xor_sync(0xffffffffu, h[d], o));
            }
        if (lane < 8) {
            const float v = lane == 0 ? l[0] : lane == 1 ? l[1] : lane == 2 ? l[2] : lane == 3 ? h[0] : lane == 4 ? h[1] : lane == 5 ? h[2] : 0.f;
            ibox[c * 8 + lane] = v;
        }
    }
}

// lower bound of ref_sqdist(q, p) over every p inside the box: the same expression on the clamped differences
__device__ __forceinline__ float box_sqdist(float qx, float qy, float qz, const float *box) {
    const float dx = fmaxf(fmaxf(box[0] - qx, qx - box[3]), 0.f);
    const float dy = fmaxf(fmaxf(box[1] - qy, qy - box[4]), 0.f);
    const float dz = fmaxf(fmaxf(box[2] - qz, qz - box[5]), 0.f);
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

constexpr int KNNP_WARPS = 16;
constexpr int KNNP_SLOTS = IDX_MAX_N / IDX_CHUNK / 32;     // chunk keys per lane (4)

// One CTA stages the whole index of its cloud in shared memory; each warp answers `qpc / 16` queries one after the other.
template <int KPL>
__global__ void __launch_bounds__(KNNP_WARPS * 32, 2)
knn_pruned_kernel(int n, int m, int k, int qpc, const float *__restrict__ index, const float *__restrict__ new_xyz,
                  int *__restrict__ idx, float *__restrict__ dist2) {
    extern __shared__ __align__(16) float sidx[];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5, cloud = blockIdx.y;
    const int npad = idx_npad(n), nch = npad / IDX_CHUNK;
    {
        const float4 *src = reinterpret_cast<const float4 *>(index + (size_t)cloud * idx_stride(n));
        float4 *dst = reinterpret_cast<float4 *>(sidx);
        const int n4 = (int)(idx_stride(n) / 4);
        for (int i = t; i < n4; i += KNNP_WARPS * 32) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    const float *sx = sidx, *sy = sx + npad, *sz = sy + npad;
    const int *sperm = reinterpret_cast<const int *>(sz + npad);
    const float *sbox = reinterpret_cast<const float *>(sperm + npad);
    const unsigned cmask = 127u;                             // chunk id bits inside a chunk key (nch <= 128)

    for (int qq = warp; qq < qpc; qq += KNNP_WARPS) {
        const int q = blockIdx.x * qpc + qq;
        if (q >= m) break;
        const float *c = new_xyz + ((size_t)cloud * m + q) * 3;
        const float qx = __ldg(c), qy = __ldg(c + 1), qz = __ldg(c + 2);
        // chunk keys: box distance (low 7 mantissa bits dropped: still a lower bound) | chunk id
        unsigned ck[KNNP_SLOTS];
#pragma unroll
        for (int s = 0; s < KNNP_SLOTS; ++s) {
            const int ch = s * 32 + lane;
            ck[s] = 0xffffffffu;
            if (ch < nch) ck[s] = (__float_as_uint(box_sqdist(qx, qy, qz, sbox + ch * 8)) & ~cmask) | (unsigned)ch;
        }
        float ld[KPL];
        int li[KPL];
#pragma unroll
        for (int s = 0; s < KPL; ++s) { ld[s] = (s * 32 + lane < k) ? INFINITY : -1.f; li[s] = 0; }
        float tau = INFINITY;

        // candidates flagged in `hit` (positions pos0 + lane of the sorted copy).  Chunks arrive in arbitrary index order, so
        // a candidate enters iff its (distance, original index) key is below the set's lexicographic maximum.
        auto absorb = [&](unsigned hit, float d, int pos0) {
            while (hit) {
                const int src = __ffs(hit) - 1;
                hit &= hit - 1;
                const float cd = __shfl_sync(0xffffffffu, d, src);
                if (!(cd <= tau)) continue;                  // tau shrank since the ballot (warp-uniform)
                const int tb = __float_as_int(tau);
                int mi = -1, ms = 0;
#pragma unroll
                for (int s = 0; s < KPL; ++s)
                    if (__float_as_int(ld[s]) == tb && li[s] > mi) { mi = li[s]; ms = s; }
                const int top = __reduce_max_sync(0xffffffffu, mi);
                const int ci = sperm[pos0 + src];
                if (cd == tau && ci >= top) continue;        // equal distance: the lower original index stays
                const unsigned vb = __ballot_sync(0xffffffffu, mi == top);      // unfilled (+inf, 0) slots tie: lowest lane
                if (lane == __ffs(vb) - 1) {
#pragma unroll
                    for (int s = 0; s < KPL; ++s)
                        if (s == ms) { ld[s] = cd; li[s] = ci; }
                }
                float md = ld[0];
#pragma unroll
                for (int s = 1; s < KPL; ++s) md = fmaxf(md, ld[s]);
                tau = __int_as_float(__reduce_max_sync(0xffffffffu, __float_as_int(md)));
            }
        };

        while (true) {
            unsigned mk = ck[0];
#pragma unroll
            for (int s = 1; s < KNNP_SLOTS; ++s) mk = min(mk, ck[s]);
            const unsigned g = __reduce_min_sync(0xffffffffu, mk);
            if (g == 0xffffffffu) break;                                    // every chunk visited
            if (__uint_as_float(g & ~cmask) > tau) break;                    // no remaining chunk can hold a better point
#pragma unroll
            for (int s = 0; s < KNNP_SLOTS; ++s)
                if (ck[s] == g) ck[s] = 0xffffffffu;
            const int base = (int)(g & cmask) * IDX_CHUNK + lane;
            const float d0 = ref_sqdist(qx, qy, qz, sx[base], sy[base], sz[base]);
            const float d1 = ref_sqdist(qx, qy, qz, sx[base + 32], sy[base + 32], sz[base + 32]);
            const unsigned h0 = __ballot_sync(0xffffffffu, d0 <= tau);
            if (h0) absorb(h0, d0, base - lane);
            const unsigned h1 = __ballot_sync(0xffffffffu, d1 <= tau);
            if (h1) absorb(h1, d1, base - lane + 32);
        }
        knn_sort_store<KPL>(ld, li, k, lane, idx + ((size_t)cloud * m + q) * k, dist2 ? dist2 + ((size_t)cloud * m + q) * k : nullptr);
    }
}

constexpr int NN3_CHUNK = 2048;

// WEIGHTS=false: raw reference op (dist2 squared + idx).  WEIGHTS=true: fused inverse-distance weights,
// patch_aug_net.py:350-353:  dist = sqrt(d2); r = 1/(dist+1e-8); w = r / (r0+r1+r2).
template <bool WEIGHTS>
__global__ void __launch_bounds__(256)
three_nn_kernel(int n, int m, const float *__restrict__ unknown, const float *__restrict__ known,
                float *__restrict__ out_f, int *__restrict__ idx) {
    __shared__ float4 pts[NN3_CHUNK];           // one broadcast LDS.128 per known point
    const int t = threadIdx.x, cloud = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + t;
    const bool active = j < n;
    const float *kn = known + (size_t)cloud * m * 3;
    float ux = 0.f, uy = 0.f, uz = 0.f;
    if (active) {
        const float *u = unknown + ((size_t)cloud * n + j) * 3;
        ux = __ldg(u); uy = __ldg(u + 1); uz = __ldg(u + 2);
    }
    float b1 = INFINITY, b2 = INFINITY, b3 = INFINITY;
    int i1 = 0, i2 = 0, i3 = 0;
    for (int base0 = 0; base0 < m; base0 += NN3_CHUNK) {
        const int cnt = min(NN3_CHUNK, m - base0);
        __syncthreads();
        for (int kk = t; kk < cnt; kk += blockDim.x) {
            const float *src = kn + (size_t)(base0 + kk) * 3;
            pts[kk] = make_float4(__ldg(src), __ldg(src + 1), __ldg(src + 2), 0.f);
        }
        __syncthreads();
        if (!active) continue;
#pragma unroll 4
        for (int kk = 0; kk < cnt; ++kk) {
            const float4 pk = pts[kk];
            const float d = ref_sqdist(ux, uy, uz, pk.x, pk.y, pk.z);
            if (d < b3) {                         // same strict-'<' cascade as the reference, entered only on a hit
                const int gi = base0 + kk;
                if (d < b1) { b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = gi; }
                else if (d < b2) { b3 = b2; i3 = i2; b2 = d; i2 = gi; }
                else { b3 = d; i3 = gi; }
            }
        }
    }
    if (!active) return;
    const size_t o = ((size_t)cloud * n + j) * 3;
    idx[o] = i1; idx[o + 1] = i2; idx[o + 2] = i3;
    if (WEIGHTS) {
        const float r1 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(b1), 1e-8f));
        const float r2 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(b2), 1e-8f));
        const float r3 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(b3), 1e-8f));
        const float nrm = __fadd_rn(__fadd_rn(r1, r2), r3);
        out_f[o] = __fdiv_rn(r1, nrm); out_f[o + 1] = __fdiv_rn(r2, nrm); out_f[o + 2] = __fdiv_rn(r3, nrm);
    } else {
        out_f[o] = b1; out_f[o + 1] = b2; out_f[o + 2] = b3;
    }
}

// k <= 32: the k-best set is kept SORTED, one 64-bit (distance bits, original index) key per lane, lane i = i-th smallest
// (all 32 lanes are live: the set simply tracks the 32 smallest, the first k are the answer).  A sparse batch of candidates
// is inserted one by one (ballot -> position -> shuffle-up); a dense batch (the first chunks, where most of the 32 lanes
// beat the threshold) is bitonic-sorted and merged with the set in one go.  No final sort.
constexpr int KNN_MERGE_MIN = 8;   // candidates in a 32-lane batch from which the merge network beats one-by-one insertion

__global__ void __launch_bounds__(KNNP_WARPS * 32, 2)
knn_pruned32_kernel(int n, int m, int k, int qpc, const float *__restrict__ index, const float *__restrict__ new_xyz,
                    int *__restrict__ idx, float *__restrict__ dist2) {
    extern __shared__ __align__(16) float sidx[];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5, cloud = blockIdx.y;
    const int npad = idx_npad(n), nch = npad / IDX_CHUNK;
    {
        const float4 *src = reinterpret_cast<const float4 *>(index + (size_t)cloud * idx_stride(n));
        float4 *dst = reinterpret_cast<float4 *>(sidx);
        const int n4 = (int)(idx_stride(n) / 4);
        for (int i = t; i < n4; i += KNNP_WARPS * 32) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    const float *sx = sidx, *sy = sx + npad, *sz = sy + npad;
    const int *sperm = reinterpret_cast<const int *>(sz + npad);
    const float *sbox = reinterpret_cast<const float *>(sperm + npad);
    const unsigned cmask = 127u;

    for (int qq = warp; qq < qpc; qq += KNNP_WARPS) {
        const int q = blockIdx.x * qpc + qq;
        if (q >= m) break;
        const float *c = new_xyz + ((size_t)cloud * m + q) * 3;
        const float qx = __ldg(c), qy = __ldg(c + 1), qz = __ldg(c + 2);
        unsigned ck[KNNP_SLOTS];
#pragma unroll
        for (int s = 0; s < KNNP_SLOTS; ++s) {
            const int ch = s * 32 + lane;
            ck[s] = 0xffffffffu;
            if (ch < nch) ck[s] = (__float_as_uint(box_sqdist(qx, qy, qz, sbox + ch * 8)) & ~cmask) | (unsigned)ch;
        }
        unsigned long long sk = knn_key(INFINITY, 0);        // like the reference's best[] = 1e40 / besti[] = 0
        unsigned long long tkey = sk;                         // k-th smallest key: a candidate enters iff its key is below it
        float tau = INFINITY;

        auto absorb = [&](unsigned hit, float d, int pos) {
            const unsigned long long mine = knn_key(d, sperm[pos]);
            if (__popc(hit) >= KNN_MERGE_MIN) {
                // sort the batch (non-candidates last), reverse it against the sorted set, keep the 32 smallest, re-sort
                unsigned long long cand = ((hit >> lane) & 1u) && mine < tkey ? mine : ~0ull;
#pragma unroll
                for (int size = 2; size <= 32; size <<= 1)
#pragma unroll
                    for (int j = size >> 1; j > 0; j >>= 1) {
                        const unsigned long long other = __shfl_xor_sync(0xffffffffu, cand, j);
                        const bool take_min = ((lane & j) == 0) == ((lane & size) == 0);
                        cand = ((cand < other) == take_min) ? cand : other;
                    }
                const unsigned long long rev = __shfl_sync(0xffffffffu, cand, 31 - lane);
                unsigned long long mrg = sk < rev ? sk : rev;  // bitonic: ascending set vs descending batch
#pragma unroll
                for (int j = 16; j > 0; j >>= 1) {
                    const unsigned long long other = __shfl_xor_sync(0xffffffffu, mrg, j);
                    const bool take_min = (lane & j) == 0;
                    mrg = ((mrg < other) == take_min) ? mrg : other;
                }
                sk = mrg;
                tkey = __shfl_sync(0xffffffffu, sk, k - 1);
            } else {
                while (hit) {
                    const int src = __ffs(hit) - 1;
                    hit &= hit - 1;
                    const unsigned long long cand = __shfl_sync(0xffffffffu, mine, src);
                    if (cand >= tkey) continue;              // warp-uniform; the threshold may have moved since the ballot
                    const int at = __ffs(__ballot_sync(0xffffffffu, sk >= cand)) - 1;     // first lane not below the candidate
                    const unsigned long long up = __shfl_up_sync(0xffffffffu, sk, 1);
                    if (lane > at) sk = up;
                    else if (lane == at) sk = cand;
                    tkey = __shfl_sync(0xffffffffu, sk, k - 1);
                }
            }
            tau = __uint_as_float((unsigned)(tkey >> 32));
        };

        while (true) {
            unsigned mk = ck[0];
#pragma unroll
            for (int s = 1; s < KNNP_SLOTS; ++s) mk = min(mk, ck[s]);
            const unsigned g = __reduce_min_sync(0xffffffffu, mk);
            if (g == 0xffffffffu) break;                                    // every chunk visited
            if (__uint_as_float(g & ~cmask) > tau) break;                    // no remaining chunk can hold a better point
#pragma unroll
            for (int s = 0; s < KNNP_SLOTS; ++s)
                if (ck[s] == g) ck[s] = 0xffffffffu;
            const int base = (int)(g & cmask) * IDX_CHUNK + lane;
            float d0, d1;                                      // both halves of the chunk through the packed fp32-pair distance
            ref_sqdist_x2(pack_f32x2(qx, qx), pack_f32x2(qy, qy), pack_f32x2(qz, qz), pack_f32x2(sx[base], sx[base + 32]),
                          pack_f32x2(sy[base], sy[base + 32]), pack_f32x2(sz[base], sz[base + 32]), d0, d1);
            const unsigned h0 = __ballot_sync(0xffffffffu, d0 <= tau);
            if (h0) absorb(h0, d0, base);
            const unsigned h1 = __ballot_sync(0xffffffffu, d1 <= tau);
            if (h1) absorb(h1, d1, base + 32);
        }
        if (lane < k) {
            idx[((size_t)cloud * m + q) * k + lane] = (int)(unsigned)(sk & 0xffffffffull);
            if (dist2) dist2[((size_t)cloud * m + q) * k + lane] = __uint_as_float((unsigned)(sk >> 32));
        }
    }
}

// 3-NN against the spatial index of the known points.  One thread per unknown point; the 32 points of a warp are visited in
// the Morton order of the unknown cloud's own index when one is given, so they are neighbours in space: the warp walks the
// known chunks in order of the box-to-box distance between ITS points and the chunk, every lane scans a visited chunk
// (broadcast shared-memory reads), and the walk stops when no remaining chunk can improve any lane's third-best distance.
// Same answer as three_nn_kernel: the three smallest (distance, original index) keys.
template <bool WEIGHTS>
__global__ void __launch_bounds__(256)
three_nn_pruned_kernel(int n, int m, const float *__restrict__ unknown, const float *__restrict__ uindex,
                       const float *__restrict__ kindex, float *__restrict__ out_f, int *__restrict__ idx) {
    extern __shared__ __align__(16) float4 kpts[];           // [mpad] (x, y, z, original index) then the sub-chunk boxes
    const int t = threadIdx.x, lane = t & 31, cloud = blockIdx.y;
    const int mpad = idx_npad(m);
    // small known clouds are walked in 16-point sub-chunks (still contiguous along the Morton curve): finer boxes prune more
    const int sub = mpad <= 2048 ? 16 : IDX_CHUNK, mch = mpad / sub;
    float *sbox = reinterpret_cast<float *>(kpts + mpad);
    {
        const float *kx = kindex + (size_t)cloud * idx_stride(m), *ky = kx + mpad, *kz = ky + mpad;
        const float *kp = kz + mpad;
        for (int i = t; i < mpad; i += 256) kpts[i] = make_float4(__ldg(kx + i), __ldg(ky + i), __ldg(kz + i), __ldg(kp + i));
    }
    __syncthreads();
    for (int ch = t; ch < mch; ch += 256) {                    // pads (positions >= m) stay outside every box
        float l[3] = {INFINITY, INFINITY, INFINITY}, h[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (int kk = 0; kk < sub && ch * sub + kk < m; ++kk) {
            const float4 pk = kpts[ch * sub + kk];
            l[0] = fminf(l[0], pk.x); l[1] = fminf(l[1], pk.y); l[2] = fminf(l[2], pk.z);
            h[0] = fmaxf(h[0], pk.x); h[1] = fmaxf(h[1], pk.y); h[2] = fmaxf(h[2], pk.z);
        }
#pragma unroll
        for (int d = 0; d < 3; ++d) { sbox[ch * 8 + d] = l[d]; sbox[ch * 8 + 3 + d] = h[d]; }
    }
    __syncthreads();
    const int j = blockIdx.x * 256 + t;                        // position in the (sorted) unknown cloud
    const bool active = j < n;
    int orig = j;
    float ux = INFINITY, uy = INFINITY, uz = INFINITY;         // inactive lanes do not widen the warp's box
    if (active) {
        if (uindex) {
            const int npad = idx_npad(n);
            const float *sxp = uindex + (size_t)cloud * idx_stride(n);
            ux = __ldg(sxp + j); uy = __ldg(sxp + npad + j); uz = __ldg(sxp + 2 * npad + j);
            orig = __float_as_int(__ldg(sxp + 3 * npad + j));
        } else {
            const float *u = unknown + ((size_t)cloud * n + j) * 3;
            ux = __ldg(u); uy = __ldg(u + 1); uz = __ldg(u + 2);
        }
    }
    // the warp's box (NaN coordinates are ignored by fmin/fmax: such a lane never selects anything anyway)
    float wl[3] = {ux, uy, uz}, wh[3] = {active ? ux : -INFINITY, active ? uy : -INFINITY, active ? uz : -INFINITY};
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            wl[d] = fminf(wl[d], __shfl_xor_sync(0xffffffffu, wl[d], o));
            wh[d] = fmaxf(wh[d], __shfl_xor_sync(0xffffffffu, wh[d], o));
        }
    const unsigned cmask = 127u;
    unsigned ck[KNNP_SLOTS];
#pragma unroll
    for (int s = 0; s < KNNP_SLOTS; ++s) {
        const int ch = s * 32 + lane;
        ck[s] = 0xffffffffu;
        if (ch < mch) {
            const float *b = sbox + ch * 8;
            // box-to-box gap per axis, combined like ref_sqdist: a lower bound for every (lane point, chunk point) pair
            const float dx = fmaxf(fmaxf(b[0] - wh[0], wl[0] - b[3]), 0.f);
            const float dy = fmaxf(fmaxf(b[1] - wh[1], wl[1] - b[4]), 0.f);
            const float dz = fmaxf(fmaxf(b[2] - wh[2], wl[2] - b[5]), 0.f);
            ck[s] = (__float_as_uint(__fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)))) & ~cmask) | (unsigned)ch;
        }
    }
    float b1 = INFINITY, b2 = INFINITY, b3 = INFINITY;
    int i1 = 0, i2 = 0, i3 = 0;
    while (true) {
        unsigned mk = ck[0];
#pragma unroll
        for (int s = 1; s < KNNP_SLOTS; ++s) mk = min(mk, ck[s]);
        const unsigned g = __reduce_min_sync(0xffffffffu, mk);
        if (g == 0xffffffffu) break;
        // the largest third-best distance over the warp's active lanes (inactive lanes report 0)
        const unsigned tmax = __reduce_max_sync(0xffffffffu, active ? __float_as_uint(b3) : 0u);
        if ((g & ~cmask) > tmax) break;                        // non-negative floats order like their bit patterns
#pragma unroll
        for (int s = 0; s < KNNP_SLOTS; ++s)
            if (ck[s] == g) ck[s] = 0xffffffffu;
        // the box-to-box bound let this chunk through; skip it all the same unless SOME lane's own point is close enough
        const float mylb = box_sqdist(ux, uy, uz, sbox + (int)(g & cmask) * 8);
        if (!__any_sync(0xffffffffu, active && mylb <= b3)) continue;
        const float4 *cp = kpts + (int)(g & cmask) * sub;
        // two known points per step through the packed fp32-pair distance (bit-identical to ref_sqdist, half the fp32-pipe
        // issue slots); chunks arrive in arbitrary index order, so candidates are compared as full (distance, index) keys
        auto consider = [&](float d, int gi) {
            if (d < b1 || (d == b1 && gi < i1)) { b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = gi; }
            else if (d < b2 || (d == b2 && gi < i2)) { b3 = b2; i3 = i2; b2 = d; i2 = gi; }
            else if (d < b3 || (d == b3 && gi < i3)) { b3 = d; i3 = gi; }
        };
        const unsigned long long ux2 = pack_f32x2(ux, ux), uy2 = pack_f32x2(uy, uy), uz2 = pack_f32x2(uz, uz);
#pragma unroll 2
        for (int kk = 0; kk < sub; kk += 2) {                  // sub is 16 or 64
            const float4 pa = cp[kk], pb = cp[kk + 1];
            float da, db;
            ref_sqdist_x2(ux2, uy2, uz2, pack_f32x2(pa.x, pb.x), pack_f32x2(pa.y, pb.y), pack_f32x2(pa.z, pb.z), da, db);
            if (da <= b3) consider(da, __float_as_int(pa.w));
            if (db <= b3) consider(db, __float_as_int(pb.w));
        }
    }
    if (!active) return;
    const size_t o = ((size_t)cloud * n + orig) * 3;
    idx[o] = i1; idx[o + 1] = i2; idx[o + 2] = i3;
    if (WEIGHTS) {
        const float r1 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(b1), 1e-8f));
        const float r2 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(b2), 1e-8f));
        const float r3 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(b3), 1e-8f));
        const float nrm = __fadd_rn(__fadd_rn(r1, r2), r3);
        out_f[o] = __fdiv_rn(r1, nrm); out_f[o + 1] = __fdiv_rn(r2, nrm); out_f[o + 2] = __fdiv_rn(r3, nrm);
    } else {
        out_f[o] = b1; out_f[o + 1] = b2; out_f[o + 2] = b3;
    }
}

// ballquery_cuda_kernel_fast (ballquery_cuda_kernel.cu:47-80) and labelstat_and_ballquery (labelstat_cuda_kernel.cu:6-49):
// first `nsample` refs in index order with d2 < r^2, padded with the first hit.  One thread per query; refs in smem.
template <bool LABELS>
__global__ void __launch_bounds__(256)
ballquery_kernel(int n, int m, float radius, int nsample, int nclass, const float *__restrict__ new_xyz,
                 const float *__restrict__ xyz, const int *__restrict__ label_stat, int *__restrict__ idx,
                 int *__restrict__ new_label_stat) {
    __shared__ float xs[NN3_CHUNK], ys[NN3_CHUNK], zs[NN3_CHUNK];
    const int t = threadIdx.x, cloud = blockIdx.y;
    const int q = blockIdx.x * blockDim.x + t;
    const bool active = q < m;
    const float *p = xyz + (size_t)cloud * n * 3;
    const float radius2 = __fmul_rn(radius, radius);
    float qx = 0.f, qy = 0.f, qz = 0.f;
    int *o = nullptr, *ol = nullptr;
    if (active) {
        const float *c = new_xyz + ((size_t)cloud * m + q) * 3;
        qx = __ldg(c); qy = __ldg(c + 1); qz = __ldg(c + 2);
        if (idx) o = idx + ((size_t)cloud * m + q) * nsample;
        if (LABELS) {
            ol = new_label_stat + ((size_t)cloud * m + q) * nclass;
            for (int i = 0; i < nclass; ++i) ol[i] = 0;
        }
    }
    int cnt_hit = 0;
    bool done = !active;
    for (int base0 = 0; base0 < n; base0 += NN3_CHUNK) {
        const int cnt = min(NN3_CHUNK, n - base0);
        __syncthreads();
        for (int e = t; e < cnt * 3; e += blockDim.x) {
            const float v = __ldg(p + (size_t)base0 * 3 + e);
            const int kk = e / 3, c = e - 3 * kk;
            (c == 0 ? xs : (c == 1 ? ys : zs))[kk] = v;
        }
        __syncthreads();
        if (done) continue;
        for (int kk = 0; kk < cnt; ++kk) {
            const float d2 = ref_sqdist(qx, qy, qz, xs[kk], ys[kk], zs[kk]);
            if (d2 < radius2) {
                const int gi = base0 + kk;
                if (LABELS) {
                    const int *ls = label_stat + ((size_t)cloud * n + gi) * nclass;
                    for (int i = 0; i < nclass; ++i) ol[i] += ls[i];
                }
                if (o) {
                    if (cnt_hit == 0) for (int l = 0; l < nsample; ++l) o[l] = gi;
                    o[cnt_hit] = gi;
                    ++cnt_hit;
                    if (cnt_hit >= nsample) { done = true; break; }
                }
            }
        }
    }
}

// featuredistribute_cuda_kernel (featuredistribute_cuda_kernel.cu:4-30): nearest centre, init 100000 / -1.
__global__ void __launch_bounds__(256)
featuredistribute_kernel(int n, int m, const float *__restrict__ max_xyz, const float *__restrict__ xyz, int *__restrict__ out) {
    __shared__ float xs[NN3_CHUNK], ys[NN3_CHUNK], zs[NN3_CHUNK];
    const int t = threadIdx.x, cloud = blockIdx.y;
    const int q = blockIdx.x * blockDim.x + t;
    const bool active = q < m;
    const float *p = max_xyz + (size_t)cloud * n * 3;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (active) {
        const float *c = xyz + ((size_t)cloud * m + q) * 3;
        qx = __ldg(c); qy = __ldg(c + 1); qz = __ldg(c + 2);
    }
    float best = 100000.f;
    int besti = -1;
    for (int base0 = 0; base0 < n; base0 += NN3_CHUNK) {
        const int cnt = min(NN3_CHUNK, n - base0);
        __syncthreads();
        for (int e = t; e < cnt * 3; e += blockDim.x) {
            const float v = __ldg(p + (size_t)base0 * 3 + e);
            const int kk = e / 3, c = e - 3 * kk;
            (c == 0 ? xs : (c == 1 ? ys : zs))[kk] = v;
        }
        __syncthreads();
        if (!active) continue;
        for (int kk = 0; kk < cnt; ++kk) {
            const float d2 = ref_sqdist(xs[kk], ys[kk], zs[kk], qx, qy, qz);
            if (d2 < best) { best = d2; besti = base0 + kk; }
        }
    }
    if (active) out[(size_t)cloud * m + q] = besti;
}

}  // namespace

namespace {

int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

// the pruned path needs the whole index of a cloud in one CTA's shared memory and <= 128 chunks
bool knn_index_ok(int n) { return n >= 256 && n <= IDX_MAX_N; }

template <int KPL>
int launch_knn_pruned(int b, int n, int m, int k, const float *index, const float *new_xyz, int *idx, float *dist2, cudaStream_t st) {
    static int n_sm = 0;
    if (!n_sm) {
        int dev = 0;
        PAB_CUDA(cudaGetDevice(&dev));
        PAB_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    }
    const size_t smem = idx_stride(n) * sizeof(float);
    // one wave: as many CTAs per cloud as fit concurrently (shared-memory bound), 16-query granules
    int per_sm = (int)((220 * 1024) / (smem + 1024));
    if (per_sm > 2) per_sm = 2;                             // __launch_bounds__(512, 2)
    if (per_sm < 1) per_sm = 1;
    int ctas_per_cloud = (per_sm * n_sm) / b;
    if (ctas_per_cloud < 1) ctas_per_cloud = 1;
    int qpc = pab_divup(pab_divup(m, ctas_per_cloud), KNNP_WARPS) * KNNP_WARPS;
    dim3 grid(pab_divup(m, qpc), b);
    if (KPL == 1) {
        PAB_CUDA(cudaFuncSetAttribute(knn_pruned32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        knn_pruned32_kernel<<<grid, KNNP_WARPS * 32, smem, st>>>(n, m, k, qpc, index, new_xyz, idx, dist2);
    } else {
        PAB_CUDA(cudaFuncSetAttribute(knn_pruned_kernel<KPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        knn_pruned_kernel<KPL><<<grid, KNNP_WARPS * 32, smem, st>>>(n, m, k, qpc, index, new_xyz, idx, dist2);
    }
    PAB_LAUNCH_CHECK();
    return 0;
}

}  // namespace

PAB_API size_t pab_knn_index_bytes(int b, int n) {
    if (b <= 0 || n <= 0) return 0;
    return (size_t)b * idx_stride(n) * sizeof(float);
}

// The Morton permutation inside an index built by pab_knn_build_index: *stride_ints = ints between consecutive clouds.  Only
// when the index has no padding rows (n a multiple of 64) — otherwise NULL.
PAB_API const int *pab_knn_index_order(int n, const void *index, long *stride_ints) {
    if (!index || !knn_index_ok(n) || idx_npad(n) != n) return nullptr;
    if (stride_ints) *stride_ints = (long)idx_stride(n);
    return reinterpret_cast<const int *>(index) + 3 * (size_t)idx_npad(n);
}

PAB_API int pab_knn_build_index(int b, int n, const float *xyz, void *index, pab_stream_t s) {
    if (b < 0 || !knn_index_ok(n) || !xyz || !index) return PAB_EINVAL;
    if (b == 0) return 0;
    const int P = next_pow2(n);
    const size_t smem = (size_t)P * sizeof(unsigned long long);
    if (smem > 48 * 1024) PAB_CUDA(cudaFuncSetAttribute(knn_index_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    knn_index_kernel<<<b, 1024, smem, (cudaStream_t)s>>>(n, P, xyz, (float *)index);
    PAB_LAUNCH_CHECK();
    return 0;
}

PAB_API int pab_knnquery_indexed(int b, int n, int m, int nsample, const void *index, const float *new_xyz, int *idx, float *dist2,
                                 pab_stream_t s) {
    if (b < 0 || !knn_index_ok(n) || m < 0 || nsample <= 0 || nsample > 64 || !index) return PAB_EINVAL;
    if (b == 0 || m == 0) return 0;
    if (nsample <= 32) return launch_knn_pruned<1>(b, n, m, nsample, (const float *)index, new_xyz, idx, dist2, (cudaStream_t)s);
    return launch_knn_pruned<2>(b, n, m, nsample, (const float *)index, new_xyz, idx, dist2, (cudaStream_t)s);
}

PAB_API int pab_knnquery(int b, int n, int m, int nsample, const float *xyz, const float *new_xyz, int *idx, float *dist2, pab_stream_t s) {
    if (b < 0 || n <= 0 || m < 0 || nsample <= 0 || nsample > 200) return PAB_EINVAL;
    if (b == 0 || m == 0) return 0;
    cudaStream_t st = (cudaStream_t)s;
    if (knn_index_ok(n) && nsample <= 64 && (long)m * b >= 1024) {
        // enough queries to pay for the index: build it in stream-ordered scratch memory, query, release
        void *index = nullptr;
        PAB_CUDA(cudaMallocAsync(&index, pab_knn_index_bytes(b, n), st));
        int rc = pab_knn_build_index(b, n, xyz, index, s);
        if (!rc) rc = pab_knnquery_indexed(b, n, m, nsample, index, new_xyz, idx, dist2, s);
        const cudaError_t e = cudaFreeAsync(index, st);
        if (!rc && e != cudaSuccess) rc = (int)e;
        return rc;
    }
    dim3 grid(pab_divup(m, KNN_WARPS), b), block(KNN_WARPS * 32);
    const int kpl = (nsample + 31) / 32;
    if (kpl <= 1) knn_kernel<1><<<grid, block, 0, st>>>(n, m, nsample, xyz, new_xyz, idx, dist2);
    else if (kpl <= 2) knn_kernel<2><<<grid, block, 0, st>>>(n, m, nsample, xyz, new_xyz, idx, dist2);
    else if (kpl <= 4) knn_kernel<4><<<grid, block, 0, st>>>(n, m, nsample, xyz, new_xyz, idx, dist2);
    else knn_kernel<8><<<grid, block, 0, st>>>(n, m, nsample, xyz, new_xyz, idx, dist2);
    PAB_LAUNCH_CHECK();
    return 0;
}

namespace {

template <bool WEIGHTS>
int launch_three_nn_pruned(int b, int n, int m, const float *unknown, const void *uindex, const void *kindex, float *out_f, int *idx,
                           cudaStream_t st) {
    const size_t smem = (size_t)idx_npad(m) * 16 + (size_t)(idx_npad(m) / 16) * 32;
    if (smem > 48 * 1024)
        PAB_CUDA(cudaFuncSetAttribute(three_nn_pruned_kernel<WEIGHTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(pab_divup(n, 256), b);
    three_nn_pruned_kernel<WEIGHTS><<<grid, 256, smem, st>>>(n, m, unknown, (const float *)uindex, (const float *)kindex, out_f, idx);
    PAB_LAUNCH_CHECK();
    return 0;
}

// drop-in path: temporary indices in stream-ordered scratch memory when the problem is large enough to pay for them
template <bool WEIGHTS>
int three_nn_auto(int b, int n, int m, const float *unknown, const float *known, float *out_f, int *idx, cudaStream_t st) {
    if (knn_index_ok(m) && (long)n * b >= 4096) {
        void *kindex = nullptr, *uindex = nullptr;
        PAB_CUDA(cudaMallocAsync(&kindex, pab_knn_index_bytes(b, m), st));
        int rc = pab_knn_build_index(b, m, known, kindex, (pab_stream_t)st);
        if (!rc && knn_index_ok(n)) {                          // Morton order of the queries keeps each warp compact
            PAB_CUDA(cudaMallocAsync(&uindex, pab_knn_index_bytes(b, n), st));
            rc = pab_knn_build_index(b, n, unknown, uindex, (pab_stream_t)st);
        }
        if (!rc) rc = launch_three_nn_pruned<WEIGHTS>(b, n, m, unknown, uindex, kindex, out_f, idx, st);
        if (uindex) cudaFreeAsync(uindex, st);
        cudaFreeAsync(kindex, st);
        return rc;
    }
    dim3 grid(pab_divup(n, 256), b);
    three_nn_kernel<WEIGHTS><<<grid, 256, 0, st>>>(n, m, unknown, known, out_f, idx);
    PAB_LAUNCH_CHECK();
    return 0;
}

}  // namespace

PAB_API int pab_nearestneighbor(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx, pab_stream_t s) {
    if (b < 0 || n < 0 || m <= 0) return PAB_EINVAL;
    if (b == 0 || n == 0) return 0;
    return three_nn_auto<false>(b, n, m, unknown, known, dist2, idx, (cudaStream_t)s);
}

PAB_API int pab_three_nn_weights(int b, int n, int m, const float *unknown, const float *known, int *idx, float *weight, pab_stream_t s) {
    if (b < 0 || n < 0 || m <= 0) return PAB_EINVAL;
    if (b == 0 || n == 0) return 0;
    return three_nn_auto<true>(b, n, m, unknown, known, weight, idx, (cudaStream_t)s);
}

PAB_API int pab_three_nn_weights_indexed(int b, int n, int m, const float *unknown, const void *unknown_index, const void *known_index,
                                         int *idx, float *weight, pab_stream_t s) {
    if (b < 0 || n < 0 || !knn_index_ok(m) || !known_index || (!unknown && !unknown_index)) return PAB_EINVAL;
    if (unknown_index && !knn_index_ok(n)) return PAB_EINVAL;
    if (b == 0 || n == 0) return 0;
    return launch_three_nn_pruned<true>(b, n, m, unknown, unknown_index, known_index, weight, idx, (cudaStream_t)s);
}

PAB_API int pab_ballquery(int b, int n, int m, float radius, int nsample, const float *new_xyz, const float *xyz, int *idx, pab_stream_t s) {
    if (b < 0 || n <= 0 || m < 0 || nsample <= 0) return PAB_EINVAL;
    if (b == 0 || m == 0) return 0;
    dim3 grid(pab_divup(m, 256), b);
    ballquery_kernel<false><<<grid, 256, 0, (cudaStream_t)s>>>(n, m, radius, nsample, 0, new_xyz, xyz, nullptr, idx, nullptr);
    PAB_LAUNCH_CHECK();
    return 0;
}

PAB_API int pab_labelstat_and_ballquery(int b, int n, int m, float radius, int nsample, int nclass, const float *new_xyz, const float *xyz,
                                        const int *label_stat, int *idx, int *new_label_stat, pab_stream_t s) {
    if (b < 0 || n <= 0 || m < 0 || nsample <= 0 || nclass < 0) return PAB_EINVAL;
    if (b == 0 || m == 0) return 0;
    dim3 grid(pab_divup(m, 256), b);
    ballquery_kernel<true><<<grid, 256, 0, (cudaStream_t)s>>>(n, m, radius, nsample, nclass, new_xyz, xyz, label_stat, idx, new_label_stat);
    PAB_LAUNCH_CHECK();
    return 0;
}

PAB_API int pab_labelstat_ballrange(int b, int n, int m, float radius, int nclass, const float *new_xyz, const float *xyz,
                                    const int *label_stat, int *new_label_stat, pab_stream_t s) {
    if (b < 0 || n <= 0 || m < 0 || nclass < 0) return PAB_EINVAL;
    if (b == 0 || m == 0) return 0;
    dim3 grid(pab_divup(m, 256), b);
    // same scan without an index list / early exit (labelstat_cuda_kernel.cu:74-105)
    ballquery_kernel<true><<<grid, 256, 0, (cudaStream_t)s>>>(n, m, radius, 1 << 30, nclass, new_xyz, xyz, label_stat, nullptr, new_label_stat);
    PAB_LAUNCH_CHECK();
    return 0;
}

PAB_API int pab_featuredistribute(int b, int n, int m, const float *max_xyz, const float *xyz, int *distribute_idx, pab_stream_t s) {
    if (b < 0 || n <= 0 || m < 0) return PAB_EINVAL;
    if (b == 0 || m == 0) return 0;
    dim3 grid(pab_divup(m, 256), b);
    featuredistribute_kernel<<<grid, 256, 0, (cudaStream_t)s>>>(n, m, max_xyz, xyz, distribute_idx);
    PAB_LAUNCH_CHECK();
    return 0;
}
