// chamfer.cu — chamfer distance forward/backward (sm_100a).
//
// Reference: libs/chamfer_dist/chamfer.cu.  Its forward uses a fixed dim3(32,16)x512 grid (chamfer.cu:159-164): on
// the training workload (49152 patch pairs of 20 points, losses/pointnetvlad_loss.py:242-247) 20 of 512 threads
// work while each block walks P/32 patches; its backward uses grid.x = 1 (chamfer.cu:215-222), so ONE block
// serialises all P patches and accumulates with atomicAdd.  Here:
//   * small clouds (n, m <= 32): one warp per patch pair, both directions in one pass, points held in shared
//     memory; the backward is a deterministic gather (each lane sums the contributions that target its own point
//     in a fixed order) instead of atomics;
//   * large clouds: one thread per point against smem tiles of the other cloud; backward with atomics like the
//     reference.
// Semantics kept: squared distance in the reference's contracted fp32 order on (p2 - p1); FIRST minimum wins
// (strict '<', chamfer.cu:47-80, tiles merged with strict '>' at :136).
#include <math.h>
#include "common.cuh"

namespace {

constexpr int CH_WARPS = 8;

__global__ void __launch_bounds__(CH_WARPS * 32)
chamfer_small_fwd_kernel(int B, int n, int m, const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                         float *__restrict__ dist1, float *__restrict__ dist2, int *__restrict__ idx1, int *__restrict__ idx2) {
    __shared__ float s1[CH_WARPS][32 * 3], s2[CH_WARPS][32 * 3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long item = (long)blockIdx.x * CH_WARPS + warp;
    if (item >= B) return;
    const float *p1 = xyz1 + item * n * 3, *p2 = xyz2 + item * m * 3;
    for (int e = lane; e < n * 3; e += 32) s1[warp][e] = __ldg(p1 + e);
    for (int e = lane; e < m * 3; e += 32) s2[warp][e] = __ldg(p2 + e);
    __syncwarp();
    if (lane < n) {
        const float x = s1[warp][lane * 3], y = s1[warp][lane * 3 + 1], z = s1[warp][lane * 3 + 2];
        float best = 0.f; int bi = 0;
        for (int k = 0; k < m; ++k) {
            const float d = ref_sqdist(s2[warp][k * 3], s2[warp][k * 3 + 1], s2[warp][k * 3 + 2], x, y, z);
            if (k == 0 || d < best) { best = d; bi = k; }
        }
        dist1[item * n + lane] = best; idx1[item * n + lane] = bi;
    }
    if (lane < m) {
        const float x = s2[warp][lane * 3], y = s2[warp][lane * 3 + 1], z = s2[warp][lane * 3 + 2];
        float best = 0.f; int bi = 0;
        for (int k = 0; k < n; ++k) {
            const float d = ref_sqdist(s1[warp][k * 3], s1[warp][k * 3 + 1], s1[warp][k * 3 + 2], x, y, z);
            if (k == 0 || d < best) { best = d; bi = k; }
        }
        dist2[item * m + lane] = best; idx2[item * m + lane] = bi;
    }
}

// grad_xyz1[j] = 2 g1[j] (p1_j - p2_{idx1[j]})  +  sum_{j': idx2[j'] == j} -2 g2[j'] (p2_{j'} - p1_j)   (and symmetrically)
__global__ void __launch_bounds__(CH_WARPS * 32)
chamfer_small_bwd_kernel(int B, int n, int m, const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                         const int *__restrict__ idx1, const int *__restrict__ idx2, const float *__restrict__ g1,
                         const float *__restrict__ g2, float *__restrict__ grad1, float *__restrict__ grad2) {
    __shared__ float s1[CH_WARPS][32 * 3], s2[CH_WARPS][32 * 3], sg1[CH_WARPS][32], sg2[CH_WARPS][32];
    __shared__ int si1[CH_WARPS][32], si2[CH_WARPS][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long item = (long)blockIdx.x * CH_WARPS + warp;
    if (item >= B) return;
    for (int e = lane; e < n * 3; e += 32) s1[warp][e] = __ldg(xyz1 + item * n * 3 + e);
    for (int e = lane; e < m * 3; e += 32) s2[warp][e] = __ldg(xyz2 + item * m * 3 + e);
    if (lane < n) { sg1[warp][lane] = __ldg(g1 + item * n + lane) * 2; si1[warp][lane] = __ldg(idx1 + item * n + lane); }
    if (lane < m) { sg2[warp][lane] = __ldg(g2 + item * m + lane) * 2; si2[warp][lane] = __ldg(idx2 + item * m + lane); }
    __syncwarp();
    if (lane < n) {
        float acc[3];
        const int j2 = si1[warp][lane];
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[c] = sg1[warp][lane] * (s1[warp][lane * 3 + c] - s2[warp][j2 * 3 + c]);
        for (int jj = 0; jj < m; ++jj)
            if (si2[warp][jj] == lane) {
#pragma unroll
                for (int c = 0; c < 3; ++c) acc[c] += -(sg2[warp][jj] * (s2[warp][jj * 3 + c] - s1[warp][lane * 3 + c]));
            }
#pragma unroll
        for (int c = 0; c < 3; ++c) grad1[(item * n + lane) * 3 + c] = acc[c];
    }
    if (lane < m) {
        float acc[3];
        const int j1 = si2[warp][lane];
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[c] = sg2[warp][lane] * (s2[warp][lane * 3 + c] - s1[warp][j1 * 3 + c]);
        for (int jj = 0; jj < n; ++jj)
            if (si1[warp][jj] == lane) {
#pragma unroll
                for (int c = 0; c < 3; ++c) acc[c] += -(sg1[warp][jj] * (s1[warp][jj * 3 + c] - s2[warp][lane * 3 + c]));
            }
#pragma unroll
        for (int c = 0; c < 3; ++c) grad2[(item * m + lane) * 3 + c] = acc[c];
    }
}

constexpr int CH_TILE = 2048;

// one direction, large clouds: thread per point of xyz1, xyz2 tiled through shared memory
__global__ void __launch_bounds__(256)
chamfer_fwd_kernel(int n, int m, const float *__restrict__ xyz1, const float *__restrict__ xyz2, float *__restrict__ dist, int *__restrict__ idx) {
    __shared__ float xs[CH_TILE], ys[CH_TILE], zs[CH_TILE];
    const int t = threadIdx.x, item = blockIdx.y;
    const int j = blockIdx.x * 256 + t;
    const bool active = j < n;
    float x = 0.f, y = 0.f, z = 0.f;
    if (active) {
        const float *p = xyz1 + ((size_t)item * n + j) * 3;
        x = __ldg(p); y = __ldg(p + 1); z = __ldg(p + 2);
    }
    float best = 0.f; int bi = 0;
    const float *q = xyz2 + (size_t)item * m * 3;
    for (int k0 = 0; k0 < m; k0 += CH_TILE) {
        const int cnt = min(CH_TILE, m - k0);
        __syncthreads();
        for (int e = t; e < cnt * 3; e += 256) {
            const float v = __ldg(q + (size_t)k0 * 3 + e);
            const int kk = e / 3, c = e - 3 * kk;
            (c == 0 ? xs : (c == 1 ? ys : zs))[kk] = v;
        }
        __syncthreads();
        if (!active) continue;
        for (int kk = 0; kk < cnt; ++kk) {
            const float d = ref_sqdist(xs[kk], ys[kk], zs[kk], x, y, z);
            if ((k0 + kk) == 0 || d < best) { best = d; bi = k0 + kk; }
        }
    }
    if (active) { dist[(size_t)item * n + j] = best; idx[(size_t)item * n + j] = bi; }
}

// chamfer_dist_grad_kernel, chamfer.cu:173-201 (one direction, accumulates with atomics)
__global__ void __launch_bounds__(256)
chamfer_bwd_kernel(int n, int m, const float *__restrict__ xyz1, const float *__restrict__ xyz2, const float *__restrict__ g,
                   const int *__restrict__ idx, float *__restrict__ grad1, float *__restrict__ grad2) {
    const int item = blockIdx.y, j = blockIdx.x * 256 + threadIdx.x;
    if (j >= n) return;
    const size_t o1 = ((size_t)item * n + j) * 3;
    const int j2 = __ldg(idx + (size_t)item * n + j);
    const size_t o2 = ((size_t)item * m + j2) * 3;
    const float gg = __ldg(g + (size_t)item * n + j) * 2;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float v = gg * (__ldg(xyz1 + o1 + c) - __ldg(xyz2 + o2 + c));
        atomicAdd(grad1 + o1 + c, v);
        atomicAdd(grad2 + o2 + c, -v);
    }
}

}  // namespace

PAB_API int pab_chamfer_forward(int B, int n, const float *xyz1, int m, const float *xyz2, float *dist1, float *dist2, int *idx1, int *idx2, pab_stream_t s) {
    if (B < 0 || n <= 0 || m <= 0) return PAB_EINVAL;
    if (B == 0) return 0;
    cudaStream_t st = (cudaStream_t)s;
    if (n <= 32 && m <= 32) {
        chamfer_small_fwd_kernel<<<pab_divup(B, CH_WARPS), CH_WARPS * 32, 0, st>>>(B, n, m, xyz1, xyz2, dist1, dist2, idx1, idx2);
        PAB_LAUNCH_CHECK();
        return 0;
    }
    if (B > 65535) return PAB_EINVAL;
    chamfer_fwd_kernel<<<dim3(pab_divup(n, 256), B), 256, 0, st>>>(n, m, xyz1, xyz2, dist1, idx1);
    PAB_LAUNCH_CHECK();
    chamfer_fwd_kernel<<<dim3(pab_divup(m, 256), B), 256, 0, st>>>(m, n, xyz2, xyz1, dist2, idx2);
    PAB_LAUNCH_CHECK();
    return 0;
}

PAB_API int pab_chamfer_backward(int B, int n, const float *xyz1, int m, const float *xyz2, const int *idx1, const int *idx2,
                                 const float *grad_dist1, const float *grad_dist2, float *grad_xyz1, float *grad_xyz2, pab_stream_t s) {
    if (B < 0 || n <= 0 || m <= 0) return PAB_EINVAL;
    if (B == 0) return 0;
    cudaStream_t st = (cudaStream_t)s;
    if (n <= 32 && m <= 32) {
        chamfer_small_bwd_kernel<<<pab_divup(B, CH_WARPS), CH_WARPS * 32, 0, st>>>(B, n, m, xyz1, xyz2, idx1, idx2, grad_dist1, grad_dist2, grad_xyz1, grad_xyz2);
        PAB_LAUNCH_CHECK();
        return 0;
    }
    if (B > 65535) return PAB_EINVAL;
    PAB_CUDA(cudaMemsetAsync(grad_xyz1, 0, sizeof(float) * (size_t)B * n * 3, st));
    PAB_CUDA(cudaMemsetAsync(grad_xyz2, 0, sizeof(float) * (size_t)B * m * 3, st));
    chamfer_bwd_kernel<<<dim3(pab_divup(n, 256), B), 256, 0, st>>>(n, m, xyz1, xyz2, grad_dist1, idx1, grad_xyz1, grad_xyz2);
    PAB_LAUNCH_CHECK();
    chamfer_bwd_kernel<<<dim3(pab_divup(m, 256), B), 256, 0, st>>>(m, n, xyz2, xyz1, grad_dist2, idx2, grad_xyz2, grad_xyz1);
    PAB_LAUNCH_CHECK();
    return 0;
}
