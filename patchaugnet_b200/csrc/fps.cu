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

// ---- pruned sampler ---------------------------------------------------------------------------------------------------
// Exact FPS that touches only the points a new sample can change.  The cloud is sorted along a Morton curve once (in
// shared memory) and cut into chunks of 32 consecutive points, each with a bounding box and a record
// (max running distance, rank of the point attaining it).  A step evaluates, per chunk, the reference's own fp32
// distance expression on the differences clamped to the box — by monotonicity of every rounding step a lower bound of
// the distance of any point inside — and skips the chunk when that bound is not below the chunk's maximum: then
// min(d, temp[k]) == temp[k] for all of its points, i.e. the reference would not change them either.  Only the
// remaining chunks (a handful once a few dozen samples exist) are updated, one point per lane, and re-reduced; the
// arg-max of the step is taken over the chunk records with the same (value, rank) order as everywhere else in this
// file, so indices and the final `temp` are bit-identical to the full scan.
// Four warps per cloud and 18 bytes of shared memory per point: three clouds share an SM (4096 points), so a batch of
// 32 occupies 11 SMs instead of 32 and leaves the rest to the dense kernels of the previous batch.
constexpr int FPSP_WARPS = 4;
constexpr int FPSP_T = FPSP_WARPS * 32;

__device__ __forceinline__ uint32_t fps_rank16(uint32_t k, int log2bs) {       // order-preserving 13-bit form of fps_rank
    const uint32_t tid = k & ((1u << log2bs) - 1u);
    return ((__brev(tid) >> (32 - log2bs)) << 3) | (k >> log2bs);
}
__device__ __forceinline__ uint32_t fps_unrank16(uint32_t r, int log2bs) {
    const uint32_t tid = __brev(r >> 3) >> (32 - log2bs);
    return ((r & 7u) << log2bs) | tid;
}
__device__ __forceinline__ uint32_t morton6(uint32_t v) {                      // 6 bits -> every third bit
    v = (v | (v << 8)) & 0x0000300Fu;
    v = (v | (v << 4)) & 0x000030C3u;
    v = (v | (v << 2)) & 0x00009249u;
    return v;
}
__device__ __forceinline__ float fps_box_bound(float qx, float qy, float qz, const float (&lo)[3], const float (&hi)[3]) {
    const float dx = fmaxf(fmaxf(__fsub_rn(lo[0], qx), __fsub_rn(qx, hi[0])), 0.f);
    const float dy = fmaxf(fmaxf(__fsub_rn(lo[1], qy), __fsub_rn(qy, hi[1])), 0.f);
    const float dz = fmaxf(fmaxf(__fsub_rn(lo[2], qz), __fsub_rn(qz, hi[2])), 0.f);
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

constexpr int FPSP_MAX_CPC = 3;

#define FPSP_SYNC() asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(FPSP_T) : "memory")

template <int SLOTS>
__global__ void __launch_bounds__(FPSP_T * (SLOTS == 1 ? FPSP_MAX_CPC : 1), 1)
fps_pruned_kernel(int b, int n, int m, int log2bs, const float *__restrict__ xyz, float *__restrict__ temp, int *__restrict__ idx) {
    // n: power of two, 1024 <= n <= 4096 * SLOTS.  A CTA of cpc * 128 threads samples cpc clouds side by side (independent
    // groups of four warps with their own named barrier): the packing onto SMs is then fixed by the launch, not left to
    // the block scheduler (which would spread 32 small CTAs over 32 idle SMs and block the next dense kernel's CTAs)
    extern __shared__ float sm_all[];
    const int sub = threadIdx.x / FPSP_T, t = threadIdx.x - sub * FPSP_T;
    const int cloud = blockIdx.x * (blockDim.x / FPSP_T) + sub;
    if (cloud >= b) return;                                        // the whole group leaves; its barrier is its own
    const int bar_id = 1 + sub;
    float *sm = sm_all + (size_t)sub * ((size_t)n * 18 / 4);
    float *xs = sm, *ys = xs + n, *zs = ys + n, *td = zs + n;
    uint16_t *rk = reinterpret_cast<uint16_t *>(td + n);
    uint32_t *keys = reinterpret_cast<uint32_t *>(td);            // sort keys live in the td region until td is initialised
    __shared__ float red_all[FPSP_MAX_CPC][6][FPSP_WARPS];
    __shared__ float bb_all[FPSP_MAX_CPC][6];
    __shared__ int2 cand_all[FPSP_MAX_CPC][2][FPSP_WARPS];
    __shared__ int s_pos0_all[FPSP_MAX_CPC];
    float (*red)[FPSP_WARPS] = red_all[sub];
    float *bb = bb_all[sub];
    int2 (*cand)[FPSP_WARPS] = cand_all[sub];
    int &s_pos0 = s_pos0_all[sub];

    const int lane = t & 31, warp = t >> 5;
    const int nch = n >> 5;
    const float *p = xyz + (size_t)cloud * n * 3;
    idx += (size_t)cloud * m;
    if (temp) temp += (size_t)cloud * n;

    // ---- bounding box (finite values only)
    {
        float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (int i = t; i < n; i += FPSP_T)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float v = __ldg(p + (size_t)i * 3 + c);
                if (fabsf(v) < INFINITY) { lo[c] = fminf(lo[c], v); hi[c] = fmaxf(hi[c], v); }
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
        FPSP_SYNC();
        if (t < 6) {
            float v = red[t][0];
            for (int w = 1; w < FPSP_WARPS; ++w) v = t < 3 ? fminf(v, red[t][w]) : fmaxf(v, red[t][w]);
            bb[t] = v;
        }
        FPSP_SYNC();
    }
    // ---- Morton keys (6 bits per axis) | original index, bitonic sort
    {
        float sc[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float ext = bb[3 + c] - bb[c];
            sc[c] = (ext > 0.f && ext < INFINITY) ? 63.5f / ext : 0.f;
        }
        for (int i = t; i < n; i += FPSP_T) {
            uint32_t code = 0;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float f = (__ldg(p + (size_t)i * 3 + c) - bb[c]) * sc[c];
                f = f > 0.f ? f : 0.f;                                   // also maps NaN to cell 0
                const uint32_t cell = f < 63.f ? (uint32_t)f : 63u;
                code |= morton6(cell) << c;
            }
            keys[i] = (code << 13) | (uint32_t)i;
        }
        FPSP_SYNC();
        for (int size = 2; size <= n; size <<= 1)
            for (int j = size >> 1; j > 0; j >>= 1) {
                for (int i = t; i < (n >> 1); i += FPSP_T) {
                    const int a = 2 * i - (i & (j - 1));
                    const uint32_t ka = keys[a], kb = keys[a + j];
                    const bool up = (a & size) == 0;
                    if ((ka > kb) == up) { keys[a] = kb; keys[a + j] = ka; }
                }
                FPSP_SYNC();
            }
    }
    // ---- sorted SoA; rk holds the original index until td (which overlays the keys) is written
    for (int i = t; i < n; i += FPSP_T) {
        const int oi = (int)(keys[i] & 8191u);
        xs[i] = __ldg(p + (size_t)oi * 3); ys[i] = __ldg(p + (size_t)oi * 3 + 1); zs[i] = __ldg(p + (size_t)oi * 3 + 2);
        rk[i] = (uint16_t)oi;
        if (oi == 0) s_pos0 = i;
    }
    FPSP_SYNC();
    for (int i = t; i < n; i += FPSP_T) {
        const int oi = rk[i];
        td[i] = temp ? temp[oi] : 1e10f;
        rk[i] = (uint16_t)fps_rank16((uint32_t)oi, log2bs);
    }
    FPSP_SYNC();

    // ---- chunk records: chunk c belongs to warp c & 3, lane (c >> 2) & 31, slot c >> 7
    float blo[SLOTS][3], bhi[SLOTS][3];
    int cmaxb[SLOTS];
    uint32_t ckey[SLOTS];
    bool valid[SLOTS];
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
        valid[s] = (((s * 32 + lane) << 2) | warp) < nch;
        cmaxb[s] = __float_as_int(-1.f);
        ckey[s] = 0xFFFFFFFFu;
#pragma unroll
        for (int c = 0; c < 3; ++c) { blo[s][c] = INFINITY; bhi[s][c] = -INFINITY; }
        for (int r = 0; r < 32; ++r) {
            const int c = ((s * 32 + r) << 2) | warp;
            if (c >= nch) break;                                         // warp-uniform
            const int pos = (c << 5) + lane;
            float l3[3] = {xs[pos], ys[pos], zs[pos]}, h3[3];
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                if (!(fabsf(l3[d]) < INFINITY)) { h3[d] = -INFINITY; l3[d] = INFINITY; } else h3[d] = l3[d];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    l3[d] = fminf(l3[d], __shfl_xor_sync(0xffffffffu, l3[d], o));
                    h3[d] = fmaxf(h3[d], __shfl_xor_sync(0xffffffffu, h3[d], o));
                }
            }
            const int bits = __float_as_int(td[pos]);
            const int wmax = __reduce_max_sync(0xffffffffu, bits);
            const uint32_t wkey = __reduce_min_sync(0xffffffffu, bits == wmax ? (((uint32_t)rk[pos] << 16) | (uint32_t)pos) : 0xFFFFFFFFu);
            if (lane == r) {
#pragma unroll
                for (int d = 0; d < 3; ++d) { blo[s][d] = l3[d]; bhi[s][d] = h3[d]; }
                cmaxb[s] = wmax; ckey[s] = wkey;
            }
        }
    }

    int opos = s_pos0;
    if (t == 0) idx[0] = 0;
    for (int j = 1; j < m; ++j) {
        const float x1 = xs[opos], y1 = ys[opos], z1 = zs[opos];
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
            // a chunk whose box is at least its current maximum away cannot change (NaN bounds compare false: never skipped)
            const float lb = fps_box_bound(x1, y1, z1, blo[s], bhi[s]);
            const bool act = valid[s] && !(lb >= __int_as_float(cmaxb[s]));
            uint32_t mask = __ballot_sync(0xffffffffu, act);
            while (mask) {
                // up to four chunks per round so that their load -> distance -> reduce chains overlap; missing ones repeat
                // the first (the update is idempotent)
                int r[4];
                r[0] = __ffs(mask) - 1; mask &= mask - 1;
#pragma unroll
                for (int u = 1; u < 4; ++u) {
                    r[u] = mask ? __ffs(mask) - 1 : r[0];
                    mask &= mask - 1;
                }
                int bits[4], pos[4];
                uint32_t rank[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    pos[u] = (((((s * 32 + r[u]) << 2) | warp)) << 5) + lane;
                    const float d = ref_sqdist(xs[pos[u]], ys[pos[u]], zs[pos[u]], x1, y1, z1);
                    const float v = fminf(d, td[pos[u]]);
                    td[pos[u]] = v;
                    bits[u] = __float_as_int(v);
                    rank[u] = rk[pos[u]];
                }
                int wmax[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) wmax[u] = __reduce_max_sync(0xffffffffu, bits[u]);
                uint32_t wkey[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    wkey[u] = __reduce_min_sync(0xffffffffu, bits[u] == wmax[u] ? ((rank[u] << 16) | (uint32_t)pos[u]) : 0xFFFFFFFFu);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (lane == r[u]) { cmaxb[s] = wmax[u]; ckey[s] = wkey[u]; }
            }
        }
        // arg-max over the warp's records, then over the four warps
        int bbits = cmaxb[0];
        uint32_t bkey = ckey[0];
#pragma unroll
        for (int s = 1; s < SLOTS; ++s)
            if (cmaxb[s] > bbits || (cmaxb[s] == bbits && ckey[s] < bkey)) { bbits = cmaxb[s]; bkey = ckey[s]; }
        const int wb = __reduce_max_sync(0xffffffffu, bbits);
        const uint32_t wk = __reduce_min_sync(0xffffffffu, bbits == wb ? bkey : 0xFFFFFFFFu);
        if (lane == 0) cand[j & 1][warp] = make_int2(wb, (int)wk);
        FPSP_SYNC();
        int gb = cand[j & 1][0].x;
        uint32_t gk = (uint32_t)cand[j & 1][0].y;
#pragma unroll
        for (int w = 1; w < FPSP_WARPS; ++w) {
            const int2 c = cand[j & 1][w];
            if (c.x > gb || (c.x == gb && (uint32_t)c.y < gk)) { gb = c.x; gk = (uint32_t)c.y; }
        }
        opos = (int)(gk & 0xFFFFu);
        if (t == 0) idx[j] = (int)fps_unrank16(gk >> 16, log2bs);
    }
    if (temp) {
        FPSP_SYNC();
        for (int i = t; i < n; i += FPSP_T) temp[fps_unrank16(rk[i], log2bs)] = td[i];
    }
}

int g_fps_pruned = 0;   // measured slower than the full scan at n <= 8192: off by default

int ref_log2_block(int n) {  // log2(opt_n_threads(n)), cuda_utils.h:15-18, without libm
    int l = 0;
    while ((2 << l) <= n && l < 10) ++l;
    return l;
}

int g_fps_threads_override = 0;

int g_fps_clouds_per_cta = 1;

int g_fps_exclusive = 0;   // (measured: no effect — the scheduler already spreads the sampler CTAs; kept as a hook) pad the sampler's dynamic shared memory beyond half an SM so that two sampler CTAs never share one:
                           // the block scheduler otherwise pairs them up on some SMs and both run their latency chain slower

template <int PPT, int MAXT>
int launch_fps(int b, int n, int m, int threads, int log2bs, const float *xyz, float *temp, int *idx, cudaStream_t st) {
    const int cpc = (g_fps_clouds_per_cta >= 2 && 2 * threads <= MAXT && b > 1) ? 2 : 1;
    size_t smem = (size_t)cpc * n * 3 * sizeof(float);
    if (g_fps_exclusive && smem < 116 * 1024) smem = 116 * 1024;
    // static smem (candidate slots) counts against the 48 KB default too: opt in whenever we are near it
    if (smem > 40 * 1024)
        PAB_CUDA(cudaFuncSetAttribute(fps_kernel<PPT, MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fps_kernel<PPT, MAXT><<<pab_divup(b, cpc), cpc * threads, smem, st>>>(b, n, m, log2bs, threads, xyz, temp, idx);
    PAB_LAUNCH_CHECK();
    return 0;
}

}  // namespace

PAB_API void pab_tune_fps_threads(int threads) { g_fps_threads_override = threads; }

PAB_API void pab_tune_fps_pruned(int on) { g_fps_pruned = on; }

PAB_API void pab_tune_fps_exclusive(int on) { g_fps_exclusive = on; }

PAB_API int pab_fps_clouds_per_sm(int n) {
    const bool pow2 = n > 0 && (n & (n - 1)) == 0;
    if (g_fps_pruned && pow2 && n >= 2048 && n <= 4096) {
        const int fit = (int)((227 * 1024 - 1024) / ((size_t)n * 18));
        const int cpc = g_fps_clouds_per_cta < 1 ? 1 : g_fps_clouds_per_cta;
        return cpc < fit ? (cpc < FPSP_MAX_CPC ? cpc : FPSP_MAX_CPC) : (fit < FPSP_MAX_CPC ? fit : FPSP_MAX_CPC);
    }
    return g_fps_clouds_per_cta == 2 ? 2 : 1;
}

PAB_API void pab_tune_fps_clouds_per_cta(int n) { g_fps_clouds_per_cta = n >= 1 && n <= 3 ? n : 1; }

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
    if (g_fps_pruned && (n & (n - 1)) == 0 && n >= 2048 && n <= 8192 && m > 1) {
        // pruned sampler: sorted chunks + box bounds (see fps_pruned_kernel)
        if (n <= 4096) {
            int cpc = g_fps_clouds_per_cta < 1 ? 1 : g_fps_clouds_per_cta;
            const int fit = (int)((227 * 1024 - 1024) / ((size_t)n * 18));      // clouds whose 18 B/point fit one SM
            if (cpc > fit) cpc = fit;
            if (cpc > FPSP_MAX_CPC) cpc = FPSP_MAX_CPC;
            if (cpc > b) cpc = b;
            const size_t smem = (size_t)cpc * n * 18;
            PAB_CUDA(cudaFuncSetAttribute(fps_pruned_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            fps_pruned_kernel<1><<<pab_divup(b, cpc), cpc * FPSP_T, smem, st>>>(b, n, m, log2bs, xyz, temp, idx);
        } else {
            const size_t smem = (size_t)n * 18;
            PAB_CUDA(cudaFuncSetAttribute(fps_pruned_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            fps_pruned_kernel<2><<<b, FPSP_T, smem, st>>>(b, n, m, log2bs, xyz, temp, idx);
        }
        PAB_LAUNCH_CHECK();
        return 0;
    }
    // threads: power of two; the per-step cost is the latency of the arg-max chain, not the distance updates, so small
    // clouds get 2 points per thread (n = 1024: 512 threads, 0.044 -> 0.035 ms), 4096 points 512 x 8, larger ones 1024 x 8
    int threads = 32;
    while (threads < 512 && threads * 2 < n) threads <<= 1;
    if (n > 4096) threads = 1024;
    if (g_fps_threads_override >= 32 && g_fps_threads_override <= 1024 &&
        (g_fps_threads_override & (g_fps_threads_override - 1)) == 0 && g_fps_threads_override * 32 >= n)
        threads = g_fps_threads_override;
    const int ppt = (n + threads - 1) / threads;
    if (ppt <= 1) return launch_fps<1, 1024>(b, n, m, threads, log2bs, xyz, temp, idx, st);
    if (ppt <= 2) return launch_fps<2, 1024>(b, n, m, threads, log2bs, xyz, temp, idx, st);
    if (ppt <= 4) return launch_fps<4, 1024>(b, n, m, threads, log2bs, xyz, temp, idx, st);
    if (ppt <= 8) return launch_fps<8, 1024>(b, n, m, threads, log2bs, xyz, temp, idx, st);
    if (ppt <= 16) return launch_fps<16, 512>(b, n, m, threads, log2bs, xyz, temp, idx, st);     // 16 points per thread: <= 512 threads (registers)
    return launch_fps<32, 256>(b, n, m, threads, log2bs, xyz, temp, idx, st);                    // 32 points per thread: <= 256 threads
}
