// emd.cu — approximate earth mover's distance by the auction algorithm (sm_100a).
//
// Reference: libs/emd_module/emd_cuda.cu.  Its host loop (emd_cuda.cu:256-269) launches 7 kernels per auction round
// (clear, calc_unass_cnt, calc_unass_cnt_sum, calc_unass_idx, Bid, GetMax, Assign) for `iters` rounds — 7000+
// launches for the documented eps=0.02 / 1024 rounds — all on the legacy default stream.  Here ONE persistent CTA
// per cloud runs every round back to back: the unassigned list is rebuilt with a block scan (ascending index, so it
// is deterministic, unlike the reference's atomicAdd order), bidding splits the 1024 threads over the unassigned
// points exactly like Bid (emd_cuda.cu:95-179: thread_per_unass = 1024 / unassigned), and the round's bookkeeping
// needs only __syncthreads.  The CTA stops as soon as every point is assigned (later rounds are no-ops in the
// reference too).  Arithmetic kept from the reference: squared distance in the contracted fp32 order on (p2 - p1),
// value = 3.0 - sqrtf(d) - price evaluated in DOUBLE (the literal 3.0 is a double, emd_cuda.cu:146) and rounded to
// float, increment = best - better + eps, the +-1e-6 double window of GetMax (emd_cuda.cu:188-191).
// Where the reference is order-dependent (equal bids on one object: last writer wins) this kernel picks the
// lowest bidder index.
#include <math.h>
#include "common.cuh"

namespace {

constexpr int EMD_T = 1024;
constexpr int EMD_TILE = 2048;

__global__ void __launch_bounds__(EMD_T, 1)
emd_kernel(int n, const float *__restrict__ xyz1, const float *__restrict__ xyz2, float *__restrict__ dist,
           int *__restrict__ assignment, float *__restrict__ price, int *__restrict__ assignment_inv, int *__restrict__ bid,
           float *__restrict__ bid_inc, float *__restrict__ max_inc, int *__restrict__ unass_idx, int *__restrict__ max_idx,
           int *__restrict__ unass_cnt, float eps, int iters) {
    __shared__ float bx[EMD_TILE], by[EMD_TILE], bz[EMD_TILE], bp[EMD_TILE];
    __shared__ float best_buf[EMD_T], better_buf[EMD_T];
    __shared__ int besti_buf[EMD_T];
    __shared__ int warp_cnt[32];
    __shared__ int s_total;

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const size_t off = (size_t)blockIdx.x * n;
    xyz1 += off * 3; xyz2 += off * 3;
    dist += off; assignment += off; price += off; assignment_inv += off; bid += off; bid_inc += off;
    max_inc += off; unass_idx += off; max_idx += off;

    for (int it = 0; it < iters; ++it) {
        const bool last = it == iters - 1;
        // ---- 1. unassigned list, ascending index (calc_unass_cnt / _sum / _idx, emd_cuda.cu:30-93)
        int total = 0;
        for (int base = 0; base < n; base += EMD_T) {
            const int j = base + t;
            const bool un = j < n && assignment[j] == -1;
            const unsigned bal = __ballot_sync(0xffffffffu, un);
            if (lane == 0) warp_cnt[warp] = __popc(bal);
            __syncthreads();
            int before = 0, all = 0;
            for (int w = 0; w < 32; ++w) { const int c = warp_cnt[w]; all += c; if (w < warp) before += c; }
            if (un) unass_idx[total + before + __popc(bal & ((1u << lane) - 1u))] = j;
            total += all;
            __syncthreads();
        }
        if (t == 0) unass_cnt[blockIdx.x] = total;
        if (total == 0) break;

        // ---- 2. Bid (emd_cuda.cu:95-179)
        // Threads per bidder exactly as the reference splits them (emd_cuda.cu:106-108: its n/1024 blocks take
        // unass_per_block = ceil(total / block_cnt) bidders each, thread_per_unass = 1024 / unass_per_block).  The split
        // decides which of several objects with EXACTLY equal value a bidder picks (a thread keeps the first maximum of its
        // own scan order — tile after tile —, threads merge in ascending order with a strict '>'), so it is part of the
        // result, not just of the schedule.  One pass = the bidders of one reference block.
        const int block_cnt = n / EMD_T;
        const int per_pass = (total + block_cnt - 1) / block_cnt;
        const int tpu = EMD_T / per_pass;                      // threads per bidder
        const int passes = (total + per_pass - 1) / per_pass;
        for (int pass = 0; pass < passes; ++pass) {
            const int cnt = min(per_pass, total - pass * per_pass);  // bidders in this pass
            const int slot = t / tpu, sub = t - slot * tpu;
            const bool has = slot < cnt;
            int me = -1;
            float x1 = 0.f, y1 = 0.f, z1 = 0.f;
            if (has) {
                me = unass_idx[pass * per_pass + slot];
                x1 = xyz1[me * 3]; y1 = xyz1[me * 3 + 1]; z1 = xyz1[me * 3 + 2];
            }
            float best = -1e9f, better = -1e9f;
            int best_i = -1;
            for (int k2 = 0; k2 < n; k2 += EMD_TILE) {
                const int end_k = min(n, k2 + EMD_TILE) - k2;
                __syncthreads();
                for (int j = t; j < end_k; j += EMD_T) {
                    bx[j] = xyz2[(k2 + j) * 3]; by[j] = xyz2[(k2 + j) * 3 + 1]; bz[j] = xyz2[(k2 + j) * 3 + 2];
                    bp[j] = price[k2 + j];
                }
                __syncthreads();
                if (has) {
                    const int delta = (end_k + tpu - 1) / tpu;
                    const int l = sub * delta, r = min((sub + 1) * delta, end_k);
                    for (int k = l; k < r; ++k) {
                        const float sq = ref_sqdist(bx[k], by[k], bz[k], x1, y1, z1);
                        const float d = (float)((3.0 - (double)__fsqrt_rn(sq)) - (double)bp[k]);
                        if (d > best) { better = best; best = d; best_i = k + k2; }
                        else if (d > better) better = d;
                    }
                }
            }
            best_buf[t] = best; better_buf[t] = better; besti_buf[t] = best_i;
            __syncthreads();
            if (has && sub == 0) {
                for (int j = t + 1; j < t + tpu; ++j) {
                    if (best_buf[j] > best) { better = fmaxf(best, better_buf[j]); best = best_buf[j]; best_i = besti_buf[j]; }
                    else better = fmaxf(better, best_buf[j]);
                }
                const float inc = __fadd_rn(__fsub_rn(best, better), eps);
                bid[me] = best_i;
                bid_inc[me] = inc;
                // float atomicMax via CAS on the bit pattern (emd_cuda.cu:10-20); same-CTA, so a plain loop suffices
                int *addr = (int *)(max_inc + best_i);
                int old = *addr;
                while (inc > __int_as_float(old)) {
                    const int assumed = old;
                    old = atomicCAS(addr, assumed, __float_as_int(inc));
                    if (old == assumed) break;
                }
            }
            __syncthreads();
        }
        __threadfence_block();
        __syncthreads();

        // ---- 3. GetMax (emd_cuda.cu:181-194): the bidder whose increment equals the object's maximum (+-1e-6)
        for (int u = t; u < total; u += EMD_T) max_idx[bid[unass_idx[u]]] = 0x7fffffff;
        __syncthreads();
        for (int u = t; u < total; u += EMD_T) {
            const int j = unass_idx[u];
            const int b_id = bid[j];
            const double bi = (double)bid_inc[j], mi = (double)max_inc[b_id];
            if (bi - 1e-6 <= mi && mi <= bi + 1e-6) atomicMin(max_idx + b_id, j);
        }
        __syncthreads();

        // ---- 4. Assign (emd_cuda.cu:196-215)
        for (int u = t; u < total; u += EMD_T) {
            const int j = unass_idx[u];
            const int b_id = bid[j];
            if (!last && max_idx[b_id] == j) {
                const int prev = assignment_inv[b_id];
                if (prev != -1) assignment[prev] = -1;
                assignment_inv[b_id] = j;
                assignment[j] = b_id;
                price[b_id] += bid_inc[j];
                max_inc[b_id] = -1e9f;
            }
        }
        if (last) {
            // final round: every bidder takes its bid; an object bid on by several keeps the highest bidder index
            // in assignment_inv and accumulates every increment (serialised here in ascending bidder order)
            __syncthreads();
            if (t == 0) {
                for (int u = 0; u < total; ++u) {
                    const int j = unass_idx[u];
                    const int b_id = bid[j];
                    assignment_inv[b_id] = j;
                    assignment[j] = b_id;
                    price[b_id] += bid_inc[j];
                    max_inc[b_id] = -1e9f;
                }
            }
        }
        __syncthreads();
    }
    __syncthreads();
    // ---- CalcDist (emd_cuda.cu:217-226)
    for (int j = t; j < n; j += EMD_T) {
        const int k = assignment[j];
        dist[j] = k >= 0 ? ref_sqdist(xyz1[j * 3], xyz1[j * 3 + 1], xyz1[j * 3 + 2], xyz2[k * 3], xyz2[k * 3 + 1], xyz2[k * 3 + 2]) : 0.f;
    }
}

// NmDistanceGradKernel, emd_cuda.cu:284-303: grad_xyz1[j] += 2 g[j] (p1_j - p2_{idx[j]})   (one writer per element)
__global__ void emd_grad_kernel(int n, const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                                const float *__restrict__ grad_dist, const int *__restrict__ idx, float *__restrict__ grad_xyz) {
    const int item = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const size_t o = ((size_t)item * n + j) * 3;
    const int j2 = idx[(size_t)item * n + j];
    const size_t o2 = ((size_t)item * n + j2) * 3;
    const float g = grad_dist[(size_t)item * n + j] * 2;
#pragma unroll
    for (int c = 0; c < 3; ++c) grad_xyz[o + c] += g * (xyz1[o + c] - xyz2[o2 + c]);
}

}  // namespace

PAB_API int pab_emd_forward(int b, int n, const float *xyz1, const float *xyz2, float *dist, int *assignment, float *price,
                            int *assignment_inv, int *bid, float *bid_increments, float *max_increments, int *unass_idx,
                            int *unass_cnt, int *unass_cnt_sum, int *cnt_tmp, int *max_idx, float eps, int iters, pab_stream_t s) {
    (void)unass_cnt_sum; (void)cnt_tmp;  // reference scratch; not needed by the persistent kernel
    if (b > 512 || n % 1024 != 0 || b < 0 || n <= 0) return -1;  // emd_cuda.cu:236-249
    if (b == 0) return 1;
    emd_kernel<<<b, EMD_T, 0, (cudaStream_t)s>>>(n, xyz1, xyz2, dist, assignment, price, assignment_inv, bid, bid_increments,
                                                  max_increments, unass_idx, max_idx, unass_cnt, eps, iters);
    ++g_pab_launches;
    return cudaGetLastError() == cudaSuccess ? 1 : 0;
}

PAB_API int pab_emd_backward(int b, int n, const float *xyz1, const float *xyz2, float *gradxyz, const float *graddist, const int *idx, pab_stream_t s) {
    if (b < 0 || b > 65535 || n <= 0) return 0;
    if (b == 0) return 1;
    emd_grad_kernel<<<dim3(pab_divup(n, 256), b), 256, 0, (cudaStream_t)s>>>(n, xyz1, xyz2, graddist, idx, gradxyz);
    ++g_pab_launches;
    return cudaGetLastError() == cudaSuccess ? 1 : 0;
}
