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
    // one bitonic sort of the 32*KPL keys (dead slots sort last), ascending (distance, index)
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
    int *o = idx + ((size_t)cloud * m + q) * k;
    float *od = dist2 ? dist2 + ((size_t)cloud * m + q) * k : nullptr;
#pragma unroll
    for (int s = 0; s < KPL; ++s) {
        const int me = s * 32 + lane;
        if (me < k) {
            o[me] = (int)(unsigned)(key[s] & 0xffffffffull);
            if (od) od[me] = __uint_as_float((unsigned)(key[s] >> 32));
        }
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

PAB_API int pab_knnquery(int b, int n, int m, int nsample, const float *xyz, const float *new_xyz, int *idx, float *dist2, pab_stream_t s) {
    if (b < 0 || n <= 0 || m < 0 || nsample <= 0 || nsample > 200) return PAB_EINVAL;
    if (b == 0 || m == 0) return 0;
    cudaStream_t st = (cudaStream_t)s;
    dim3 grid(pab_divup(m, KNN_WARPS), b), block(KNN_WARPS * 32);
    const int kpl = (nsample + 31) / 32;
    if (kpl <= 1) knn_kernel<1><<<grid, block, 0, st>>>(n, m, nsample, xyz, new_xyz, idx, dist2);
    else if (kpl <= 2) knn_kernel<2><<<grid, block, 0, st>>>(n, m, nsample, xyz, new_xyz, idx, dist2);
    else if (kpl <= 4) knn_kernel<4><<<grid, block, 0, st>>>(n, m, nsample, xyz, new_xyz, idx, dist2);
    else knn_kernel<8><<<grid, block, 0, st>>>(n, m, nsample, xyz, new_xyz, idx, dist2);
    PAB_LAUNCH_CHECK();
    return 0;
}

PAB_API int pab_nearestneighbor(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx, pab_stream_t s) {
    if (b < 0 || n < 0 || m <= 0) return PAB_EINVAL;
    if (b == 0 || n == 0) return 0;
    dim3 grid(pab_divup(n, 256), b);
    three_nn_kernel<false><<<grid, 256, 0, (cudaStream_t)s>>>(n, m, unknown, known, dist2, idx);
    PAB_LAUNCH_CHECK();
    return 0;
}

PAB_API int pab_three_nn_weights(int b, int n, int m, const float *unknown, const float *known, int *idx, float *weight, pab_stream_t s) {
    if (b < 0 || n < 0 || m <= 0) return PAB_EINVAL;
    if (b == 0 || n == 0) return 0;
    dim3 grid(pab_divup(n, 256), b);
    three_nn_kernel<true><<<grid, 256, 0, (cudaStream_t)s>>>(n, m, unknown, known, weight, idx);
    PAB_LAUNCH_CHECK();
    return 0;
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
