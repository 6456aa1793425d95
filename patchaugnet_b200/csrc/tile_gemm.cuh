// tile_gemm.cuh — fp32 SIMT tile GEMM building blocks shared by the fused SharedMLP, NetVLAD and head kernels.
//
// A CTA of 256 threads owns a tile of R rows (points).  Activations live in shared memory ROW-MAJOR
// ([row][channel], channel contiguous — the K-major operand layout, same orientation a tcgen05 A operand wants)
// with a row stride s = 4 (mod 32) floats so that the four distinct rows a warp touches per LDS.128 fall in
// disjoint bank quads.  Each thread accumulates 8 rows x 4 columns; one column pass covers CT = 8192/R columns.
// Weights are streamed global->smem in KC-row chunks with cp.async double buffering; they are stored as
// Wt[c_in][c_out] (BatchNorm scale folded in on the host) so a chunk is a plain row copy.
#pragma once
#include "common.cuh"

namespace tg {

constexpr int KC = 16;       // weight rows per staged chunk
constexpr int THREADS = 256;

__host__ __device__ inline int stride_for(int c) {  // smallest s >= c with s % 32 == 4
    return c + ((4 - (c % 32)) + 32) % 32;
}

template <int R>
struct Geo {
    static constexpr int CT = 8192 / R;   // columns per pass
    static constexpr int RG = R / 8;      // row groups (thread rows are rg + RG*i, i<8)
    static constexpr int CG = CT / 4;     // column groups (4 consecutive columns each)
    static constexpr int CGW = CG / 8;    // warps along the column dimension
    __device__ static __forceinline__ int cg() { return ((threadIdx.x >> 5) % CGW) * 8 + (threadIdx.x & 7); }
    __device__ static __forceinline__ int rg() { return ((threadIdx.x >> 5) / CGW) * 4 + ((threadIdx.x >> 3) & 3); }
};

// acc[i][j] += sum_{k<kcount} Xs[(rg + RG*i)*sx + k] * Ws[k*sw + 4*cg + j]     (kcount % 4 == 0)
template <int RG>
__device__ __forceinline__ void fma_block(float (&acc)[8][4], const float *__restrict__ Xs, int sx, int rg,
                                          const float *__restrict__ Ws, int sw, int cg, int kcount) {
    const float *xrow = Xs + rg * sx;
    const float *wcol = Ws + 4 * cg;
#pragma unroll 2
    for (int kk = 0; kk < kcount; kk += 4) {
        float4 w0 = *reinterpret_cast<const float4 *>(wcol + (kk + 0) * sw);
        float4 w1 = *reinterpret_cast<const float4 *>(wcol + (kk + 1) * sw);
        float4 w2 = *reinterpret_cast<const float4 *>(wcol + (kk + 2) * sw);
        float4 w3 = *reinterpret_cast<const float4 *>(wcol + (kk + 3) * sw);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 x = *reinterpret_cast<const float4 *>(xrow + (RG * i) * sx + kk);
            acc[i][0] = fmaf(x.x, w0.x, acc[i][0]); acc[i][1] = fmaf(x.x, w0.y, acc[i][1]);
            acc[i][2] = fmaf(x.x, w0.z, acc[i][2]); acc[i][3] = fmaf(x.x, w0.w, acc[i][3]);
            acc[i][0] = fmaf(x.y, w1.x, acc[i][0]); acc[i][1] = fmaf(x.y, w1.y, acc[i][1]);
            acc[i][2] = fmaf(x.y, w1.z, acc[i][2]); acc[i][3] = fmaf(x.y, w1.w, acc[i][3]);
            acc[i][0] = fmaf(x.z, w2.x, acc[i][0]); acc[i][1] = fmaf(x.z, w2.y, acc[i][1]);
            acc[i][2] = fmaf(x.z, w2.z, acc[i][2]); acc[i][3] = fmaf(x.z, w2.w, acc[i][3]);
            acc[i][0] = fmaf(x.w, w3.x, acc[i][0]); acc[i][1] = fmaf(x.w, w3.y, acc[i][1]);
            acc[i][2] = fmaf(x.w, w3.z, acc[i][2]); acc[i][3] = fmaf(x.w, w3.w, acc[i][3]);
        }
    }
}

// One folded layer over the tile: Ys[r][c] = act(sum_k Xs[r][k] * Wt[k][c] + shift[c]), c < c_out.
// wstage: 2 * KC * CT floats of shared memory.  Ends with the tile fully written but NOT synchronised.
template <int R>
__device__ void layer(const float *__restrict__ Xs, int sx, float *__restrict__ Ys, int sy, const pab_layer_t &L,
                      float *__restrict__ wstage) {
    using G = Geo<R>;
    constexpr int CT = G::CT;
    const int t = threadIdx.x, cg = G::cg(), rg = G::rg();
    const int cin = L.c_in_pad, cout = L.c_out;
    const int nchunks = (cin + KC - 1) / KC;
    constexpr int UNITS = KC * CT / 4;  // 16-byte units per chunk

    for (int c0 = 0; c0 < cout; c0 += CT) {
        float acc[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

        auto load_chunk = [&](int ch) {
            float *dst = wstage + (ch & 1) * KC * CT;
            const int k0 = ch * KC;
            for (int u = t; u < UNITS; u += THREADS) {
                const int row = u / (CT / 4), cu = u % (CT / 4);
                const int k = k0 + row, col = c0 + 4 * cu;
                if (k < cin && col < cout) cp_async16(dst + row * CT + 4 * cu, L.wt + (size_t)k * cout + col);
            }
            cp_async_commit();
        };
        load_chunk(0);
        for (int ch = 0; ch < nchunks; ++ch) {
            if (ch + 1 < nchunks) { load_chunk(ch + 1); cp_async_wait<1>(); }
            else cp_async_wait<0>();
            __syncthreads();
            const int kcount = min(KC, cin - ch * KC);
            fma_block<G::RG>(acc, Xs + ch * KC, sx, rg, wstage + (ch & 1) * KC * CT, CT, cg, kcount);
            __syncthreads();
        }
        const int col = c0 + 4 * cg;
        if (col < cout) {
            const float4 sh = *reinterpret_cast<const float4 *>(L.shift + col);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float4 v = make_float4(acc[i][0] + sh.x, acc[i][1] + sh.y, acc[i][2] + sh.z, acc[i][3] + sh.w);
                if (L.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                *reinterpret_cast<float4 *>(Ys + (rg + G::RG * i) * sy + col) = v;
            }
        }
    }
}

}  // namespace tg
