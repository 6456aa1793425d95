// vlad.cu — NetVLAD soft-assignment + residual aggregation and the AdaptiveFeatureAggregator head (sm_100a).
//
// NetVLADBase.forward (place_recognition/patch_aug_net/models/loupe.py:191-222) is ~10 PyTorch kernels per level:
// transpose+contiguous, matmul, BN1d, softmax, sum, mul, transpose, matmul, sub, normalize.  Here one kernel reads
// each (128-point x C) tile of the point-major feature map ONCE into shared memory and runs both contractions on
// it:  logits = x Wc + shift  ->  softmax over K (one thread per point)  ->  vlad[K,C] += act^T x, accumulated in
// registers across the CTA's tiles; per-CTA partials are combined in a fixed order by a small finalize kernel that
// also subtracts a_sum * cluster_weights2 and applies the intra-cluster L2 normalisation (deterministic, no atomics).
//
// AdaptiveFeatureAggregator (loupe.py:57-66) + MLPAttentionLayer (loupe.py:24-41):  attention logits = max over
// output channels of conv1d(v); softmax over the K clusters; y = relu(v + v*w); fc over the flattened (C*K) vector
// split along the 21504-long reduction so the 22 MB weight is streamed exactly once; bias + BN1d + L2 in finalize.
#include <math.h>
#include "tile_gemm.cuh"

namespace {

constexpr int VR = 128;          // rows (points) per tile
constexpr int VROWS_PER_CTA = 512;
constexpr int VKMAX = 64;

struct VladArgs {
    int n, c, K, sx;
    const float *x, *wc, *shift;
    float *part, *asum;   // part (b, nchunk, K, c); asum (b, nchunk, K)
    int nchunk;
};

__global__ void __launch_bounds__(tg::THREADS, 1) vlad_partial_kernel(const VladArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int sx = a.sx;                       // stride_for(c)
    constexpr int SL = 68;                     // logits row stride (>= VKMAX, = 4 mod 32)
    constexpr int ST = VR + 4;                 // act^T row stride
    float *xs = smem;                          // [VR][sx]
    float *lg = xs + (size_t)VR * sx;          // [VR][SL]
    float *actT = lg + VR * SL;                // [VKMAX][ST]
    float *wstage = actT + VKMAX * ST;         // 2*KC*64

    const int t = threadIdx.x, cloud = blockIdx.y, chunk = blockIdx.x;
    const int row_begin = chunk * VROWS_PER_CTA;
    const int row_end = min(a.n, row_begin + VROWS_PER_CTA);
    const float *xg = a.x + (size_t)cloud * a.n * a.c;

    pab_layer_t L{};
    L.wt = a.wc; L.shift = a.shift; L.c_in = a.c; L.c_in_pad = a.c; L.c_out = a.K; L.relu = 0;

    using G2 = tg::Geo<64>;                    // second contraction: 64 "rows" (clusters) x 128-column passes
    const int cg2 = G2::cg(), rg2 = G2::rg();
    float acc[2][8][4];
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[p][i][j] = 0.f;
    float asum = 0.f;

    for (int e = t; e < VKMAX * ST; e += tg::THREADS) actT[e] = 0.f;

    for (int r0 = row_begin; r0 < row_end; r0 += VR) {
        __syncthreads();  // previous tile fully consumed
        const int c4 = a.c / 4;
        for (int e = t; e < VR * c4; e += tg::THREADS) {
            const int r = e / c4, q = e - r * c4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r0 + r < row_end) v = __ldg(reinterpret_cast<const float4 *>(xg + (size_t)(r0 + r) * a.c) + q);
            *reinterpret_cast<float4 *>(xs + (size_t)r * sx + 4 * q) = v;
        }
        tg::layer<VR>(xs, sx, lg, SL, L, wstage);   // logits (bn1 folded), first sync inside orders the xs writes
        __syncthreads();
        if (t < VR) {
            const bool valid = r0 + t < row_end;
            const float *row = lg + t * SL;
            float mx = -INFINITY;
            for (int k = 0; k < a.K; ++k) mx = fmaxf(mx, row[k]);
            float sum = 0.f;
            for (int k = 0; k < a.K; ++k) sum += expf(row[k] - mx);
            const float inv = 1.f / sum;
            for (int k = 0; k < a.K; ++k) actT[k * ST + t] = valid ? expf(row[k] - mx) * inv : 0.f;
        }
        __syncthreads();
        if (t < a.K) {
            const float *row = actT + t * ST;
            float s = 0.f;
            for (int r = 0; r < VR; ++r) s += row[r];
            asum += s;
        }
#pragma unroll
        for (int p = 0; p < 2; ++p)
            if (p * 128 < a.c) tg::fma_block<G2::RG>(acc[p], actT, ST, rg2, xs + p * 128, sx, cg2, VR);
    }
    float *part = a.part + ((size_t)cloud * a.nchunk + chunk) * a.K * a.c;
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        const int col = p * 128 + 4 * cg2;
        if (col < a.c) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int k = rg2 + G2::RG * i;
                if (k < a.K)
                    *reinterpret_cast<float4 *>(part + (size_t)k * a.c + col) =
                        make_float4(acc[p][i][0], acc[p][i][1], acc[p][i][2], acc[p][i][3]);
            }
        }
    }
    if (t < a.K) a.asum[((size_t)cloud * a.nchunk + chunk) * a.K + t] = asum;
}

// One CTA per (cloud, group of 8 clusters), one warp per cluster, eight channels per lane: the per-chunk partials are
// summed in chunk order (deterministic), a = a_sum * cluster_weights2 is subtracted, every cluster is L2-normalised over
// its channels (each cluster is independent), and the (C, K) slice is written as 32-byte runs of consecutive clusters.
constexpr int VFK = 8;           // clusters per CTA
__global__ void __launch_bounds__(256) vlad_finalize_kernel(int c, int K, int nchunk, const float *__restrict__ part,
                                                           const float *__restrict__ asum, const float *__restrict__ w2,
                                                           float *__restrict__ out, long out_bstride, long out_cstride) {
    __shared__ float w2s[256][VFK];            // cluster_weights2[t][k0 + kk]
    __shared__ float res[256][VFK + 1];
    const int t = threadIdx.x, cloud = blockIdx.y, lane = t & 31, warp = t >> 5;
    const int k0 = blockIdx.x * VFK, k = k0 + warp;
    const int kn = min(VFK, K - k0);
    for (int ch = t; ch < c; ch += 256)
        for (int kk = 0; kk < kn; ++kk) w2s[ch][kk] = __ldg(w2 + (size_t)ch * K + k0 + kk);
    __syncthreads();
    if (k < K) {
        float as = 0.f;
        for (int ch = 0; ch < nchunk; ++ch) as += __ldg(asum + ((size_t)cloud * nchunk + ch) * K + k);
        float sq = 0.f;
        for (int c0 = lane * 4; c0 < c; c0 += 128) {                       // c <= 256: two float4 per lane
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8                                        // independent loads, additions in chunk order
            for (int ch = 0; ch < nchunk; ++ch) {
                const float4 p4 = __ldg(reinterpret_cast<const float4 *>(part + (((size_t)cloud * nchunk + ch) * K + k) * c + c0));
                v.x += p4.x; v.y += p4.y; v.z += p4.z; v.w += p4.w;
            }
            v.x -= as * w2s[c0][warp]; v.y -= as * w2s[c0 + 1][warp];       // vlad - a,  a = a_sum * cluster_weights2
            v.z -= as * w2s[c0 + 2][warp]; v.w -= as * w2s[c0 + 3][warp];
            sq += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
            res[c0][warp] = v.x; res[c0 + 1][warp] = v.y; res[c0 + 2][warp] = v.z; res[c0 + 3][warp] = v.w;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        const float inv = 1.f / fmaxf(sqrtf(sq), 1e-12f);                   // F.normalize(dim=1, p=2, eps=1e-12)
        for (int c0 = lane * 4; c0 < c; c0 += 128)
#pragma unroll
            for (int j = 0; j < 4; ++j) res[c0 + j][warp] *= inv;
    }
    __syncthreads();
    for (int ch = t; ch < c; ch += 256) {
        float *dst = out + (size_t)cloud * out_bstride + (size_t)ch * out_cstride + k0;
        for (int kk = 0; kk < kn; ++kk) dst[kk] = res[ch][kk];
    }
}

// ---- AdaptiveFeatureAggregator ---------------------------------------------------------------------------

constexpr int AKC = 28;  // clusters per CTA in the attention kernel

constexpr int ATT_SPLIT = 2;   // CTAs sharing the output channels of one (cloud, cluster group)

// mx[z][b][k] = max over the z-th part of c' of sum_c w_att_t[c][c'] * v[b][c][k]   (conv1d without bias, then max over channels;
// the softmax kernel takes the maximum over z)
__global__ void __launch_bounds__(128) afa_att_kernel(int b, int c, int K, const float *__restrict__ v, const float *__restrict__ w_att_t,
                                                     float *__restrict__ mx) {
    extern __shared__ __align__(16) float vs[];  // [c][AKC]
    __shared__ float red[AKC][4];
    const int t = threadIdx.x, cloud = blockIdx.y, k0 = blockIdx.x * AKC, z = blockIdx.z;
    const int kn = min(AKC, K - k0);
    for (int e = t; e < c * AKC; e += 128) {
        const int ci = e / AKC, kk = e - ci * AKC;
        vs[e] = kk < kn ? __ldg(v + ((size_t)cloud * c + ci) * K + k0 + kk) : 0.f;
    }
    __syncthreads();
    float best[AKC];
#pragma unroll
    for (int kk = 0; kk < AKC; ++kk) best[kk] = -INFINITY;
    const int cpz = (c + ATT_SPLIT - 1) / ATT_SPLIT;
    for (int co = z * cpz + t; co < min(c, (z + 1) * cpz); co += 128) {
        float acc[AKC];
#pragma unroll
        for (int kk = 0; kk < AKC; ++kk) acc[kk] = 0.f;
        // the weight column of this output channel is a chain of dependent-latency global loads (one L2 round trip per group):
        // 16 input channels per trip — sixteen loads in flight — instead of four; the accumulation order is unchanged
        for (int ci0 = 0; ci0 < c; ci0 += 16) {
            float w[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) w[u] = ci0 + u < c ? __ldg(w_att_t + (size_t)(ci0 + u) * c + co) : 0.f;
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                if (ci0 + u < c) {
                    const float4 *vr = reinterpret_cast<const float4 *>(vs + (ci0 + u) * AKC);
#pragma unroll
                    for (int q = 0; q < AKC / 4; ++q) {
                        const float4 x = vr[q];
                        acc[4 * q + 0] = fmaf(w[u], x.x, acc[4 * q + 0]); acc[4 * q + 1] = fmaf(w[u], x.y, acc[4 * q + 1]);
                        acc[4 * q + 2] = fmaf(w[u], x.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(w[u], x.w, acc[4 * q + 3]);
                    }
                }
            }
        }
#pragma unroll
        for (int kk = 0; kk < AKC; ++kk) best[kk] = fmaxf(best[kk], acc[kk]);
    }
    const int lane = t & 31, warp = t >> 5;
#pragma unroll
    for (int kk = 0; kk < AKC; ++kk) {
        float m = best[kk];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0) red[kk][warp] = m;
    }
    __syncthreads();
    if (t < kn) mx[((size_t)z * b + cloud) * K + k0 + t] = fmaxf(fmaxf(red[t][0], red[t][1]), fmaxf(red[t][2], red[t][3]));
}

constexpr int FCH = 64;    // reduction slice per CTA in the fc kernel
constexpr int FB = 32;     // clouds per register pass

// wsm[b][k] = softmax_k(mx[b][:])      (MLPAttentionLayer softmax over the K clusters, loupe.py:31)
__global__ void __launch_bounds__(128) afa_softmax_kernel(int b, int K, int nsplit, float *__restrict__ mx, float *__restrict__ wsm) {
    __shared__ float red[4];
    const int t = threadIdx.x, cloud = blockIdx.x, lane = t & 31, warp = t >> 5;
    float *m = mx + (size_t)cloud * K;
    float mm = -INFINITY;
    for (int k = t; k < K; k += 128) {                       // combine the attention kernel's channel parts (in place, part 0)
        float a = m[k];
        for (int z = 1; z < nsplit; ++z) a = fmaxf(a, mx[((size_t)z * b + cloud) * K + k]);
        m[k] = a;
        mm = fmaxf(mm, a);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mm = fmaxf(mm, __shfl_xor_sync(0xffffffffu, mm, o));
    if (lane == 0) red[warp] = mm;
    __syncthreads();
    mm = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    __syncthreads();
    float sum = 0.f;
    for (int k = t; k < K; k += 128) sum += expf(m[k] - mm);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    sum = (red[0] + red[1]) + (red[2] + red[3]);
    for (int k = t; k < K; k += 128) wsm[(size_t)cloud * K + k] = expf(m[k] - mm) / sum;
}

// part[slice][b][o] = sum_{f in slice} y[b][f] * fc_wt[f][o],  y = relu(v + v * wsm[b][k]),  f = c*K + k
__global__ void __launch_bounds__(256) afa_fc_kernel(int b, int c, int K, int c_out, const float *__restrict__ v,
                                                    const float *__restrict__ wsm, const float *__restrict__ fc_wt,
                                                    float *__restrict__ part) {
    __shared__ __align__(16) float ys[FCH][FB];
    const int t = threadIdx.x, slice = blockIdx.x;
    const int F = c * K, f0 = slice * FCH, fn = min(FCH, F - f0);
    for (int b0 = 0; b0 < b; b0 += FB) {
        const int bn = min(FB, b - b0);
        __syncthreads();
        for (int e = t; e < FB * FCH; e += 256) {
            const int bb = e / FCH, ff = e - bb * FCH;
            float y = 0.f;
            if (bb < bn && ff < fn) {
                const int cloud = b0 + bb, k = (f0 + ff) % K;
                const float x = __ldg(v + (size_t)cloud * F + f0 + ff);
                // wsm == nullptr: plain fc over the flattened vector (PPT-Net / PointNetVLAD head), no attention weighting
                y = wsm ? fmaxf(x + x * __ldg(wsm + (size_t)cloud * K + k), 0.f) : x;
            }
            ys[ff][bb] = y;
        }
        __syncthreads();
        for (int o = t; o < c_out; o += 256) {
            float acc[FB];
#pragma unroll
            for (int bb = 0; bb < FB; ++bb) acc[bb] = 0.f;
            for (int ff0 = 0; ff0 < fn; ff0 += 16) {             // sixteen weight rows in flight; rows >= fn of ys are zero
                float w[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) w[u] = ff0 + u < fn ? __ldg(fc_wt + (size_t)(f0 + ff0 + u) * c_out + o) : 0.f;
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    const float4 *yr = reinterpret_cast<const float4 *>(&ys[ff0 + u][0]);
#pragma unroll
                    for (int q = 0; q < FB / 4; ++q) {
                        const float4 y = yr[q];
                        acc[4 * q + 0] = fmaf(y.x, w[u], acc[4 * q + 0]); acc[4 * q + 1] = fmaf(y.y, w[u], acc[4 * q + 1]);
                        acc[4 * q + 2] = fmaf(y.z, w[u], acc[4 * q + 2]); acc[4 * q + 3] = fmaf(y.w, w[u], acc[4 * q + 3]);
                    }
                }
            }
#pragma unroll
            for (int bb = 0; bb < FB; ++bb)
                if (bb < bn) part[((size_t)slice * b + b0 + bb) * c_out + o] = acc[bb];
        }
    }
}

// desc[b][o] = normalize( (sum_slices part + ...) * scale[o] + shift[o] )     (fc bias and BN1d folded by the host)
// 1024 threads: four groups each sum a quarter of the slices of output o = t % 256 (c_out <= 256 per pass), combined in a
// fixed order.
__global__ void __launch_bounds__(1024) afa_finalize_kernel(int b, int c_out, int nslice, const float *__restrict__ part,
                                                           const float *__restrict__ scale, const float *__restrict__ shift,
                                                           int l2_norm, float *__restrict__ desc) {
    __shared__ float grp[4][256];
    __shared__ float red[8];
    const int t = threadIdx.x, cloud = blockIdx.x, lane = t & 31, warp = t >> 5;
    const int g = t >> 8, ot = t & 255;
    float sq = 0.f;
    for (int o0 = 0; o0 < c_out; o0 += 256) {
        const int o = o0 + ot;
        float s = 0.f;
        if (o < c_out) {
#pragma unroll 12                                       // independent loads (latency-bound), the additions stay in slice order
            for (int sl = g; sl < nslice; sl += 4) s += __ldg(part + ((size_t)sl * b + cloud) * c_out + o);
        }
        __syncthreads();
        grp[g][ot] = s;
        __syncthreads();
        if (g == 0 && o < c_out) {
            s = (grp[0][ot] + grp[1][ot]) + (grp[2][ot] + grp[3][ot]);
            s = fmaf(s, __ldg(scale + o), __ldg(shift + o));
            desc[(size_t)cloud * c_out + o] = s;
            sq += s * s;
        }
    }
    if (!l2_norm) return;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    if (warp < 8 && lane == 0) red[warp] = sq;
    __syncthreads();
    float tot = 0.f;
    for (int w = 0; w < 8; ++w) tot += red[w];
    const float inv = 1.f / fmaxf(sqrtf(tot), 1e-12f);
    if (g == 0)
        for (int o = ot; o < c_out; o += 256) desc[(size_t)cloud * c_out + o] *= inv;
}

// PPT-Net head after the split-K fc (pptnet_origin/models/loupe.py:99-105, 107-136): x = bn2(sum of the fc partials);
// gates = sigmoid(bn1(x @ G)) (or + gating_biases, folded into scale/shift by the host); out = x * gates; optional L2.
// One CTA per cloud, c_out <= 256 (one thread per output).
__global__ void __launch_bounds__(256) gated_finalize_kernel(int b, int c_out, int nslice, const float *__restrict__ part,
                                                            const float *__restrict__ scale, const float *__restrict__ shift,
                                                            const float *__restrict__ gate_wt, const float *__restrict__ gate_scale,
                                                            const float *__restrict__ gate_shift, int l2_norm, float *__restrict__ desc) {
    __shared__ float xs[256];
    __shared__ float red[8];
    const int t = threadIdx.x, cloud = blockIdx.x, lane = t & 31, warp = t >> 5;
    float x = 0.f;
    if (t < c_out) {
        float s = 0.f;
        for (int sl = 0; sl < nslice; ++sl) s += __ldg(part + ((size_t)sl * b + cloud) * c_out + t);
        x = fmaf(s, __ldg(scale + t), __ldg(shift + t));
    }
    xs[t] = x;
    __syncthreads();
    float y = 0.f;
    if (t < c_out) {
        float g = 0.f;
        if (gate_wt) {
            for (int i = 0; i < c_out; ++i) g = fmaf(xs[i], __ldg(gate_wt + (size_t)i * c_out + t), g);
            g = fmaf(g, __ldg(gate_scale + t), __ldg(gate_shift + t));
            y = x * (1.f / (1.f + expf(-g)));
        } else {
            y = x;
        }
    }
    float sq = y * y;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    if (lane == 0) red[warp] = sq;
    __syncthreads();
    float tot = 0.f;
    for (int w = 0; w < 8; ++w) tot += red[w];
    if (t < c_out) desc[(size_t)cloud * c_out + t] = l2_norm ? y / fmaxf(sqrtf(tot), 1e-12f) : y;
}

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

// tensor-core partial kernel (vlad_tc.cu)
int pab_vlad_tc_partial(int b, int n, int c, int K, const float *x, const void *wc_hi, const void *wc_lo, const float *shift,
                        float *part, float *asum, int *nchunk_out, cudaStream_t st);

PAB_API size_t pab_netvlad_workspace_bytes(int b, int n, int c, int K) {
    const size_t nchunk = (n + 255) / 256;      // the tensor-core path uses 256-row items, the SIMT path 512-row chunks
    return align256(sizeof(float) * (size_t)b * nchunk * K * c) + align256(sizeof(float) * (size_t)b * nchunk * K);
}

PAB_API int pab_netvlad_forward(int b, int n, int c, int K, const float *x, const float *wc, const float *shift, const float *w2,
                                float *out, long out_bstride, long out_cstride, void *workspace, pab_stream_t s) {
    if (b < 0 || n <= 0 || (c != 128 && c != 256) || K <= 0 || K > VKMAX || K % 4 || !workspace) return PAB_EINVAL;
    if (b == 0) return 0;
    cudaStream_t st = (cudaStream_t)s;
    const int nchunk = (n + VROWS_PER_CTA - 1) / VROWS_PER_CTA;
    VladArgs a;
    a.n = n; a.c = c; a.K = K; a.sx = tg::stride_for(c); a.x = x; a.wc = wc; a.shift = shift; a.nchunk = nchunk;
    a.part = (float *)workspace;
    a.asum = (float *)((char *)workspace + align256(sizeof(float) * (size_t)b * nchunk * K * c));
    const size_t smem = sizeof(float) * ((size_t)VR * a.sx + VR * 68 + VKMAX * (VR + 4) + 2 * tg::KC * 64);
    PAB_CUDA(cudaFuncSetAttribute(vlad_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    vlad_partial_kernel<<<dim3(nchunk, b), tg::THREADS, smem, st>>>(a);
    PAB_LAUNCH_CHECK();
    vlad_finalize_kernel<<<dim3(pab_divup(K, VFK), b), 256, 0, st>>>(c, K, nchunk, a.part, a.asum, w2, out, out_bstride, out_cstride);
    PAB_LAUNCH_CHECK();
    return 0;
}

PAB_API int pab_netvlad_forward_tc(int b, int n, int c, int K, const float *x, const void *wc_hi, const void *wc_lo, const float *shift,
                                   const float *w2, float *out, long out_bstride, long out_cstride, void *workspace, pab_stream_t s) {
    if (b < 0 || n <= 0 || c != 256 || K <= 0 || K > VKMAX || K % 4 || !workspace) return PAB_EINVAL;
    if (b == 0) return 0;
    cudaStream_t st = (cudaStream_t)s;
    const int nchunk_max = (n + 255) / 256;
    float *part = (float *)workspace;
    float *asum = (float *)((char *)workspace + align256(sizeof(float) * (size_t)b * nchunk_max * K * c));
    int nchunk = 0;
    const int rc = pab_vlad_tc_partial(b, n, c, K, x, wc_hi, wc_lo, shift, part, asum, &nchunk, st);
    if (rc) return rc;
    vlad_finalize_kernel<<<dim3(pab_divup(K, VFK), b), 256, 0, st>>>(c, K, nchunk, part, asum, w2, out, out_bstride, out_cstride);
    PAB_LAUNCH_CHECK();
    return 0;
}

PAB_API size_t pab_afa_workspace_bytes(int b, int c, int K, int c_out) {
    const size_t nslice = ((size_t)c * K + FCH - 1) / FCH;
    return (1 + ATT_SPLIT) * align256(sizeof(float) * (size_t)b * K) + align256(sizeof(float) * nslice * b * c_out);
}

PAB_API int pab_afa_forward(int b, int c, int K, int c_out, const float *v, const float *w_att_t, const float *fc_wt,
                            const float *fc_scale, const float *fc_shift, int l2_norm, float *desc, void *workspace, pab_stream_t s) {
    if (b < 0 || c <= 0 || K <= 0 || c_out <= 0 || !workspace) return PAB_EINVAL;
    if (b == 0) return 0;
    cudaStream_t st = (cudaStream_t)s;
    float *wsm = (float *)workspace;
    float *mx = (float *)((char *)workspace + align256(sizeof(float) * (size_t)b * K));          // ATT_SPLIT parts of (b, K)
    float *part = (float *)((char *)workspace + (1 + ATT_SPLIT) * align256(sizeof(float) * (size_t)b * K));
    const int nslice = (c * K + FCH - 1) / FCH;
    const size_t smem = sizeof(float) * (size_t)c * AKC;
    if (smem > 48 * 1024) PAB_CUDA(cudaFuncSetAttribute(afa_att_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    afa_att_kernel<<<dim3(pab_divup(K, AKC), b, ATT_SPLIT), 128, smem, st>>>(b, c, K, v, w_att_t, mx);
    PAB_LAUNCH_CHECK();
    afa_softmax_kernel<<<b, 128, 0, st>>>(b, K, ATT_SPLIT, mx, wsm);
    PAB_LAUNCH_CHECK();
    afa_fc_kernel<<<nslice, 256, 0, st>>>(b, c, K, c_out, v, wsm, fc_wt, part);
    PAB_LAUNCH_CHECK();
    afa_finalize_kernel<<<b, 1024, 0, st>>>(b, c_out, nslice, part, fc_scale, fc_shift, l2_norm, desc);
    PAB_LAUNCH_CHECK();
    return 0;
}

// ---- the same head with the two products on the tensor cores (afa_tc.cu) ------------------------------------------------
int pab_afa_tc_eligible(int c, int K, int c_out);
int pab_afa_tc_att_parts(int c);
int pab_afa_tc_fc_slices(int c, int K);
int pab_afa_tc_att(int b, int c, int K, const float *v, const void *w_hi, const void *w_lo, float *mx, cudaStream_t st);
int pab_afa_tc_fused_softmax(int b, int K);
int pab_afa_tc_fc(int b, int c, int K, int c_out, const float *v, const float *wsm, const float *mx, int nparts, const void *w_hi,
                  const void *w_lo, float *part, cudaStream_t st);

PAB_API int pab_afa_tc_supported(int c, int K, int c_out) { return pab_afa_tc_eligible(c, K, c_out); }

PAB_API size_t pab_afa_tc_workspace_bytes(int b, int c, int K, int c_out) {
    if (!pab_afa_tc_eligible(c, K, c_out)) return 0;
    return (1 + (size_t)pab_afa_tc_att_parts(c)) * align256(sizeof(float) * (size_t)b * K) +
           align256(sizeof(float) * (size_t)pab_afa_tc_fc_slices(c, K) * b * c_out);
}

PAB_API int pab_afa_forward_tc(int b, int c, int K, int c_out, const float *v, const void *watt_hi, const void *watt_lo,
                               const void *wfc_hi, const void *wfc_lo, const float *fc_scale, const float *fc_shift, int l2_norm,
                               float *desc, void *workspace, pab_stream_t s) {
    if (b < 0 || c <= 0 || K <= 0 || c_out <= 0 || !workspace || !watt_hi || !watt_lo || !wfc_hi || !wfc_lo) return PAB_EINVAL;
    if (!pab_afa_tc_eligible(c, K, c_out)) return PAB_EINVAL;
    if (b == 0) return 0;
    cudaStream_t st = (cudaStream_t)s;
    const int nparts = pab_afa_tc_att_parts(c), nslice = pab_afa_tc_fc_slices(c, K);
    float *wsm = (float *)workspace;
    float *mx = (float *)((char *)workspace + align256(sizeof(float) * (size_t)b * K));
    float *part = (float *)((char *)workspace + (1 + (size_t)nparts) * align256(sizeof(float) * (size_t)b * K));
    int rc = pab_afa_tc_att(b, c, K, v, watt_hi, watt_lo, mx, st);
    if (rc) return rc;
    const bool fused = pab_afa_tc_fused_softmax(b, K) != 0;                 // softmax inside the fc kernel: three launches
    if (!fused) {
        afa_softmax_kernel<<<b, 128, 0, st>>>(b, K, nparts, mx, wsm);
        PAB_LAUNCH_CHECK();
    }
    rc = pab_afa_tc_fc(b, c, K, c_out, v, wsm, fused ? mx : nullptr, nparts, wfc_hi, wfc_lo, part, st);
    if (rc) return rc;
    afa_finalize_kernel<<<b, 1024, 0, st>>>(b, c_out, nslice, part, fc_scale, fc_shift, l2_norm, desc);
    PAB_LAUNCH_CHECK();
    return 0;
}

// gated fc head (PPT-Net / PointNetVLAD) with the fc product on the tensor cores: the fc kernel of afa_tc.cu in plain mode
PAB_API int pab_gated_fc_tc_supported(int f, int c_out) { return f > 0 && f % 64 == 0 && pab_afa_tc_eligible(64, f / 64, c_out); }

PAB_API int pab_gated_fc_forward_tc(int b, int f, int c_out, const float *v, const void *wfc_hi, const void *wfc_lo, const float *fc_scale,
                                    const float *fc_shift, const float *gate_wt, const float *gate_scale, const float *gate_shift,
                                    int l2_norm, float *desc, void *workspace, pab_stream_t s) {
    if (b < 0 || f <= 0 || c_out <= 0 || !workspace || !v || !wfc_hi || !wfc_lo || !pab_gated_fc_tc_supported(f, c_out)) return PAB_EINVAL;
    if (gate_wt && (!gate_scale || !gate_shift)) return PAB_EINVAL;
    if (b == 0) return 0;
    cudaStream_t st = (cudaStream_t)s;
    float *part = (float *)workspace;                       // pab_gated_fc_workspace_bytes: f / 64 slices >= the kernel's CTAs
    const int nslice = pab_afa_tc_fc_slices(f, 1);
    int rc = pab_afa_tc_fc(b, f, 1, c_out, v, nullptr, nullptr, 0, wfc_hi, wfc_lo, part, st);
    if (rc) return rc;
    gated_finalize_kernel<<<b, 256, 0, st>>>(b, c_out, nslice, part, fc_scale, fc_shift, gate_wt, gate_scale, gate_shift, l2_norm, desc);
    PAB_LAUNCH_CHECK();
    return 0;
}

PAB_API size_t pab_gated_fc_workspace_bytes(int b, int f, int c_out) {
    const size_t nslice = ((size_t)f + FCH - 1) / FCH;
    return align256(sizeof(float) * nslice * b * c_out);
}

PAB_API int pab_gated_fc_forward(int b, int f, int c_out, const float *v, const float *fc_wt, const float *fc_scale, const float *fc_shift,
                                 const float *gate_wt, const float *gate_scale, const float *gate_shift, int l2_norm, float *desc,
                                 void *workspace, pab_stream_t s) {
    if (b < 0 || f <= 0 || c_out <= 0 || c_out > 256 || !workspace || !v || !fc_wt) return PAB_EINVAL;
    if (gate_wt && (!gate_scale || !gate_shift)) return PAB_EINVAL;
    if (b == 0) return 0;
    cudaStream_t st = (cudaStream_t)s;
    float *part = (float *)workspace;
    const int nslice = (f + FCH - 1) / FCH;
    // the split-K fc kernel of the AFA head with the attention weighting switched off (c = f, K = 1: f = c*K + k)
    afa_fc_kernel<<<nslice, 256, 0, st>>>(b, f, 1, c_out, v, nullptr, fc_wt, part);
    PAB_LAUNCH_CHECK();
    gated_finalize_kernel<<<b, 256, 0, st>>>(b, c_out, nslice, part, fc_scale, fc_shift, gate_wt, gate_scale, gate_shift, l2_norm, desc);
    PAB_LAUNCH_CHECK();
    return 0;
}
