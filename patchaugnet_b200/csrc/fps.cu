// fps.cu — furthest point sampling for sm_100a.
//
// Replaces furthestsampling_cuda_kernel (libs/pointops/src/sampling/sampling_cuda_kernel.cu:58-168).
// One CTA per cloud.  xyz is staged once into shared memory (SoA) with coalesced loads; each thread keeps its
// points AND their running min-distances in registers for the whole m-step loop (the reference re-reads xyz
// and read-modify-writes `temp` in global memory every step and runs an 11-barrier smem tree).  The per-step
// arg-max is: per-thread FMNMX chain -> redux.sync (warp max) -> one smem slot per warp -> ONE __syncthreads ->
// every warp re-reduces the <=32 slots with redux.sync.  Slots are double-buffered by step parity.
//
// Bit-exactness.  Distances use the reference's contracted order (common.cuh ref_sqdist, dx = p[k] - p[old]).
// The reference's winner among EQUAL maxima is fixed by its reduction tree: thread tid = k mod BS scans
// k = tid, tid+BS, ... with strict '>', then __update(v2 > v1 ? i2 : i1) folds tid+s into tid for s = BS/2..1
// (sampling_cuda_kernel.cu:48-54, 84-162), BS = opt_n_threads(n) (cuda_utils.h:15-18).  That order is
// "smallest bit-reversed tid, then smallest k"; it is encoded here as a 32-bit rank
//     rank(k) = (bitrev10(k mod BS) << 22) | (k div BS)
// and ties are resolved with a min-reduction on rank, independent of this kernel's own thread layout.
#include "common.cuh"

namespace {

__device__ __forceinline__ uint32_t fps_rank(uint32_t k, int log2bs) {
    const uint32_t tid = k & ((1u << log2bs) - 1u);
    return ((__brev(tid) >> 22) << 22) | (k >> log2bs);
}
__device__ __forceinline__ uint32_t fps_unrank(uint32_t rank, int log2bs) {
    const uint32_t tid = __brev((rank >> 22) << 22);
    return ((rank & 0x3FFFFFu) << log2bs) | tid;
}

template <int PPT, int MAXT>
__global__ void __launch_bounds__(MAXT, 1)
fps_kernel(int b, int n, int m, int log2bs, int T, const float *__restrict__ xyz, float *__restrict__ temp, int *__restrict__ idx) {
    // T threads per cloud; a CTA of 2T threads runs two clouds side by side (independent halves, named barriers): the serial
    // arg-max chain of one cloud leaves most issue slots of its SM idle, so two clouds share an SM at little cost and the
    // sampler occupies half as many SMs while the dense kernels of the previous batch run on the rest
    extern __shared__ float sm[];
    const int sub = threadIdx.x / T, t = threadIdx.x - sub * T;
    const int cloud = blockIdx.x * (blockDim.x / T) + sub;
    if (cloud >= b) return;                                        // odd tail: the whole half leaves (its barrier is its own)
    float *xs = sm + (size_t)sub * 3 * n, *ys = xs + n, *zs = xs + 2 * n;
    __shared__ int2 cand_all[2][2][32];
    int2 (*cand)[32] = cand_all[sub];
    const int bar_id = 1 + sub;

    const int lane = t & 31, warp = t >> 5, nw = T >> 5;
    const float *p = xyz + (size_t)cloud * n * 3;
    idx += (size_t)cloud * m;
    if (temp) temp += (size_t)cloud * n;

    // coalesced staging of the raw (n,3) array, de-interleaved into SoA
    for (int e = t; e < n * 3; e += T) {
        const float v = __ldg(p + e);
        const int k = e / 3, c = e - 3 * k;
        (c == 0 ? xs : (c == 1 ? ys : zs))[k] = v;
    }
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(T) : "memory");

    float px[PPT], py[PPT], pz[PPT], td[PPT];
    uint32_t rk[PPT];
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
        const int k = t + T * i;
        if (k < n) {
            px[i] = xs[k]; py[i] = ys[k]; pz[i] = zs[k];
            td[i] = temp ? temp[k] : 1e10f;
            rk[i] = fps_rank((uint32_t)k, log2bs);
        } else {
            px[i] = py[i] = pz[i] = 0.f;
            td[i] = -1.f;  // never wins: every real distance is >= 0
            rk[i] = 0xFFFFFFFFu;
        }
    }
    // points kept as fp32 PAIRS for the packed distance (PPT >= 2)
    unsigned long long px2[(PPT + 1) / 2], py2[(PPT + 1) / 2], pz2[(PPT + 1) / 2];
    if (PPT >= 2) {
#pragma unroll
        for (int i = 0; i + 1 < PPT; i += 2) {
            px2[i / 2] = pack_f32x2(px[i], px[i + 1]); py2[i / 2] = pack_f32x2(py[i], py[i + 1]); pz2[i / 2] = pack_f32x2(pz[i], pz[i + 1]);
        }
    }

    int old = 0;
    if (t == 0) idx[0] = 0;
    for (int j = 1; j < m; ++j) {
        const float x1 = xs[old], y1 = ys[old], z1 = zs[old];
        float best = -1.f;
        if (PPT >= 2) {
            const unsigned long long x2 = pack_f32x2(x1, x1), y2 = pack_f32x2(y1, y1), z2 = pack_f32x2(z1, z1);
#pragma unroll
            for (int i = 0; i + 1 < PPT; i += 2) {
                float d0, d1;
                ref_sqdist_x2(px2[i / 2], py2[i / 2], pz2[i / 2], x2, y2, z2, d0, d1);
                td[i] = fminf(d0, td[i]);
                td[i + 1] = fminf(d1, td[i + 1]);
                best = fmaxf(best, fmaxf(td[i], td[i + 1]));
            }
        } else {
#pragma unroll
            for (int i = 0; i < PPT; ++i) {
                const float d = ref_sqdist(px[i], py[i], pz[i], x1, y1, z1);
                td[i] = fminf(d, td[i]);
                best = fmaxf(best, td[i]);
            }
        }
        // floats >= 0 (and the -1 sentinel) order like their bit patterns read as signed ints
        const int bbits = __float_as_int(best);
        const int wmax = __reduce_max_sync(0xffffffffu, bbits);
        uint32_t myrank = 0xFFFFFFFFu;
        if (bbits == wmax) {
#pragma unroll
            for (int i = 0; i < PPT; ++i)
                if (__float_as_int(td[i]) == wmax) myrank = min(myrank, rk[i]);
        }
        const uint32_t wrank = __reduce_min_sync(0xffffffffu, myrank);
        uint32_t grank;
        if (nw > 1) {
            if (lane == 0) cand[j & 1][warp] = make_int2(wmax, (int)wrank);
            asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(T) : "memory");
            int2 c = lane < nw ? cand[j & 1][lane] : make_int2(INT_MIN, -1);
            const int gmax = __reduce_max_sync(0xffffffffu, c.x);
            grank = __reduce_min_sync(0xffffffffu, c.x == gmax ? (uint32_t)c.y : 0xFFFFFFFFu);
        } else {
            grank = wrank;
        }
        old = (int)fps_unrank(grank, log2bs);
        if (t == 0) idx[j] = old;
    }
    if (temp) {
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
            const int k = t + T * i;
            if (k < n) temp[k] = td[i];
        }
    }
}

// Fallback for clouds too large for the register-resident kernel: min-distances live in global `temp`
// (as in the reference), xyz is read through L1/L2.  Same rank-based tie-break.
__global__ void __launch_bounds__(1024, 1)
fps_generic_kernel(int n, int m, int log2bs, const float *__restrict__ xyz, float *__restrict__ temp, int *__restrict__ idx) {
    __shared__ int2 cand[2][32];
    const int T = blockDim.x, t = threadIdx.x;
    const int lane = t & 31, warp = t >> 5, nw = T >> 5;
    const float *p = xyz + (size_t)blockIdx.x * n * 3;
    idx += (size_t)blockIdx.x * m;
    temp += (size_t)blockIdx.x * n;
    int old = 0;
    if (t == 0) idx[0] = 0;
    for (int j = 1; j < m; ++j) {
        const float x1 = p[old * 3 + 0], y1 = p[old * 3 + 1], z1 = p[old * 3 + 2];
        int bbits = __float_as_int(-1.f);
        uint32_t myrank = 0xFFFFFFFFu;
        for (int k = t; k < n; k += T) {
            const float d = ref_sqdist(p[k * 3 + 0], p[k * 3 + 1], p[k * 3 + 2], x1, y1, z1);
            const float d2 = fminf(d, temp[k]);
            temp[k] = d2;
            const int db = __float_as_int(d2);
            const uint32_t r = fps_rank((uint32_t)k, log2bs);
            if (db > bbits || (db == bbits && r < myrank)) { bbits = db; myrank = r; }
        }
        const int wmax = __reduce_max_sync(0xffffffffu, bbits);
        const uint32_t wrank = __reduce_min_sync(0xffffffffu, bbits == wmax ? myrank : 0xFFFFFFFFu);
        if (lane == 0) cand[j & 1][warp] = make_int2(wmax, (int)wrank);
        __syncthreads();
        int2 c = lane < nw ? cand[j & 1][lane] : make_int2(INT_MIN, -1);
        const int gmax = __reduce_max_sync(0xffffffffu, c.x);
        const uint32_t grank = __reduce_min_sync(0xffffffffu, c.x == gmax ? (uint32_t)c.y : 0xFFFFFFFFu);
        old = (int)fps_unrank(grank, log2bs);
        if (t == 0) idx[j] = old;
    }
}

int ref_log2_block(int n) {  // log2(opt_n_threads(n)), cuda_utils.h:15-18, without libm
    int l = 0;
    while ((2 << l) <= n && l < 10) ++l;
    return l;
}

int g_fps_threads_override = 0;

int g_fps_clouds_per_cta = 1;

template <int PPT, int MAXT>
int launch_fps(int b, int n, int m, int threads, int log2bs, const float *xyz, float *temp, int *idx, cudaStream_t st) {
    const int cpc = (g_fps_clouds_per_cta == 2 && 2 * threads <= MAXT && b > 1) ? 2 : 1;
    const size_t smem = (size_t)cpc * n * 3 * sizeof(float);
    // static smem (candidate slots) counts against the 48 KB default too: opt in whenever we are near it
    if (smem > 40 * 1024)
        PAB_CUDA(cudaFuncSetAttribute(fps_kernel<PPT, MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fps_kernel<PPT, MAXT><<<pab_divup(b, cpc), cpc * threads, smem, st>>>(b, n, m, log2bs, threads, xyz, temp, idx);
    PAB_LAUNCH_CHECK();
    return 0;
}

}  // namespace

PAB_API void pab_tune_fps_threads(int threads) { g_fps_threads_override = threads; }

PAB_API void pab_tune_fps_clouds_per_cta(int n) { g_fps_clouds_per_cta = n == 2 ? 2 : 1; }

PAB_API int pab_furthestsampling(int b, int n, int m, const float *xyz, float *temp, int *idx, pab_stream_t s) {
    if (b < 0 || n <= 0 || m < 0 || m > n) return PAB_EINVAL;
    if (b == 0 || m == 0) return 0;
    cudaStream_t st = (cudaStream_t)s;
    const int log2bs = ref_log2_block(n);
    if (n > 8192) {
        if (!temp) return PAB_EINVAL;
        fps_generic_kernel<<<b, 1024, 0, st>>>(n, m, log2bs, xyz, temp, idx);
        PAB_LAUNCH_CHECK();
        return 0;
    }
    // threads: power of two; the per-step cost is the latency of the arg-max chain, not the distance updates, so small
    // clouds get 2 points per thread (n = 1024: 512 threads, 0.044 -> 0.035 ms), 4096 points 512 x 8, larger ones 1024 x 8
    int threads = 32;
    while (threads < 512 && threads * 2 < n) threads <<= 1;
    if (n > 4096) threads = 1024;
    if (g_fps_threads_override >= 32 && g_fps_threads_override <= 1024 &&
        (g_fps_threads_override & (g_fps_threads_override - 1)) == 0 && g_fps_threads_override * 16 >= n)
        threads = g_fps_threads_override;
    const int ppt = (n + threads - 1) / threads;
    if (ppt <= 1) return launch_fps<1, 1024>(b, n, m, threads, log2bs, xyz, temp, idx, st);
    if (ppt <= 2) return launch_fps<2, 1024>(b, n, m, threads, log2bs, xyz, temp, idx, st);
    if (ppt <= 4) return launch_fps<4, 1024>(b, n, m, threads, log2bs, xyz, temp, idx, st);
    if (ppt <= 8) return launch_fps<8, 1024>(b, n, m, threads, log2bs, xyz, temp, idx, st);
    return launch_fps<16, 512>(b, n, m, threads, log2bs, xyz, temp, idx, st);     // 16 points per thread: <= 512 threads (registers)
}
