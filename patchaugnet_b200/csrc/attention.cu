// attention.cu — PPT-Net's grouped self-attention layer (SA_Layer), fused, sm_100a.
//
// Reference: place_recognition/pptnet_origin/models/pptnet.py:246-282.  It materialises the (B, 8, N, N) per-group
// energies (1.07 GB at B=32, N=1024), sums them, softmaxes, renormalises columns, and multiplies — ~12 kernels.
// Because q_conv and k_conv share one weight (pptnet.py:254) and the per-group dot products add up to the dot product
// over all channels, energy = Q^T Q is a symmetric Gram matrix of ONE projected tensor.  With
//     P[n1][n2] = exp(E[n1][n2] - rowmax[n1]) / rowsum[n1]           (softmax over the last dim, pptnet.py:277)
//     attn      = P / (1e-9 + colsum),  colsum[n2] = sum_n1 P[n1][n2]  (pptnet.py:278)
//     x_r[:,n2] = (sum_n1 V[:,n1] P[n1][n2]) / (1e-9 + colsum[n2])     (pptnet.py:279)
// nothing of size N x N ever leaves the SM:
//   1. pointwise projections Q = x Wq_dense, V = x Wv + bv, interleaved [Q | V] per point   (mlp.cu, point-major)
//   2. attn_stats_kernel:  row max / row sum of E, one CTA per 64-row tile, streaming 128-column tiles
//   3. attn_apply_kernel:  per 64-column tile of n2, recompute E tiles (E is symmetric, so they are produced
//      directly as P^T), accumulate U = P^T-tile x V and colsum in registers, write d = x - U / (1e-9 + colsum)
//   4. pointwise trans_conv + BN + ReLU with residual:  out = x + relu(Wt d + bt)   (mlp.cu)
// fp32 SIMT tiles (tile_gemm.cuh); all tensors point-major (b, n, c).
#include <math.h>
#include "tile_gemm.cuh"

int pab_pw_tc_eligible(const pab_layer_t *L, long out_ld);
int pab_pw_tc_launch(long rows, const float *x, const pab_layer_t *L, const float *residual, float *out, long out_ld, cudaStream_t st);
int pab_pointwise_mlp_residual(int rows, const float *x, const pab_layer_t *layers, int n_layers, const float *residual,
                               float *out, long out_ld, cudaStream_t st);   // mlp.cu

int pab_attention_tc(int b, int n, int c, const float *q, const float *v, long ld, const float *x, float *rowmax, float *rowsum,
                     float *d, int precision, cudaStream_t st);   // attention_tc.cu

namespace {

constexpr int AT_R = 64;       // rows per CTA tile
constexpr int AT_C = 128;      // columns per streamed tile (= Geo<64>::CT)
constexpr int AT_KC = 64;      // channel chunk of the Gram products
constexpr int SXQ = 68;        // stride of the [64][64] Q chunk
constexpr int SPT = 132;       // stride of the [64][128] E / P^T tile

using G = tg::Geo<AT_R>;

// E tile [rows r0..r0+64) x [cols c0..c0+128) of Q Q^T into acc (thread tile 8 x 4), K = C in chunks of 64.
// xq: smem [64][SXQ]; wq: smem [64][AT_C] (chunk of Q^T for the column tile).
__device__ __forceinline__ void gram_tile(float (&acc)[8][4], const float *__restrict__ q, long ld, int n, int C, int r0, int c0,
                                          float *xq, float *wq) {
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < C; k0 += AT_KC) {
        __syncthreads();
        // rows chunk: coalesced along channels
        for (int e = t; e < AT_R * AT_KC; e += tg::THREADS) {
            const int r = e / AT_KC, k = e - r * AT_KC;
            xq[r * SXQ + k] = (r0 + r < n) ? __ldg(q + (long)(r0 + r) * ld + k0 + k) : 0.f;
        }
        // columns chunk, transposed: lane = column (conflict-free smem writes), each warp walks 8 channels
        for (int cc = lane; cc < AT_C; cc += 32) {
            const bool ok = c0 + cc < n;
            const float *src = q + (long)(c0 + cc) * ld + k0;
            for (int k = warp * 8; k < warp * 8 + 8; ++k) wq[k * AT_C + cc] = ok ? __ldg(src + k) : 0.f;
        }
        __syncthreads();
        tg::fma_block<G::RG>(acc, xq, SXQ, G::rg(), wq, AT_C, G::cg(), AT_KC);
    }
}

// rowmax[n1], rowsum[n1] of E = Q Q^T over all n2
__global__ void __launch_bounds__(tg::THREADS, 1)
attn_stats_kernel(int n, int C, const float *__restrict__ q, long ld, float *__restrict__ rowmax, float *__restrict__ rowsum) {
    extern __shared__ __align__(16) float smem[];
    float *xq = smem, *wq = xq + AT_R * SXQ, *et = wq + AT_KC * AT_C;
    const int t = threadIdx.x, cloud = blockIdx.y, r0 = blockIdx.x * AT_R;
    q += (long)cloud * n * ld;
    float m = -INFINITY, s = 0.f;
    float acc[8][4];
    for (int c0 = 0; c0 < n; c0 += AT_C) {
        gram_tile(acc, q, ld, n, C, r0, c0, xq, wq);
#pragma unroll
        for (int i = 0; i < 8; ++i)
            *reinterpret_cast<float4 *>(et + (G::rg() + G::RG * i) * SPT + 4 * G::cg()) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        __syncthreads();
        if (t < AT_R) {
            const float *row = et + t * SPT;
            const int cn = min(AT_C, n - c0);
            float tm = m;
            for (int j = 0; j < cn; ++j) tm = fmaxf(tm, row[j]);
            float ts = s * expf(m - tm);
            for (int j = 0; j < cn; ++j) ts += expf(row[j] - tm);
            m = tm; s = ts;
        }
    }
    if (t < AT_R && r0 + t < n) {
        rowmax[(long)cloud * n + r0 + t] = m;
        rowsum[(long)cloud * n + r0 + t] = s;
    }
}

// d[n2][c] = x[n2][c] - (sum_n1 P[n1][n2] V[n1][c]) / (1e-9 + sum_n1 P[n1][n2])
__global__ void __launch_bounds__(tg::THREADS, 1)
attn_apply_kernel(int n, int C, const float *__restrict__ q, const float *__restrict__ v, long ld, const float *__restrict__ x,
                  const float *__restrict__ rowmax, const float *__restrict__ rowsum, float *__restrict__ d) {
    extern __shared__ __align__(16) float smem[];
    float *xq = smem, *wq = xq + AT_R * SXQ;          // wq: [64][128] Q^T chunk, reused as the [128][128] V tile
    float *pt = wq + AT_C * AT_C, *st = pt + AT_R * SPT, *csum = st + 2 * AT_C;
    const int t = threadIdx.x, cloud = blockIdx.y, r0 = blockIdx.x * AT_R;   // r0: first n2 of this tile
    q += (long)cloud * n * ld; v += (long)cloud * n * ld;
    x += (long)cloud * n * C; d += (long)cloud * n * C;
    rowmax += (long)cloud * n; rowsum += (long)cloud * n;
    const int rg = G::rg(), cg = G::cg();

    for (int cp = 0; cp < C; cp += AT_C) {            // channel pass of the output (recomputes P for C > 128: tiny N there)
        float u[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) u[i][j] = 0.f;
        float cs = 0.f;
        for (int c0 = 0; c0 < n; c0 += AT_C) {        // n1 tile
            float acc[8][4];
            gram_tile(acc, q, ld, n, C, r0, c0, xq, wq);     // E'[n2][n1] = E[n1][n2] (symmetric)
            __syncthreads();
            for (int j = t; j < AT_C; j += tg::THREADS) {
                const bool ok = c0 + j < n;
                st[j] = ok ? rowmax[c0 + j] : 0.f;
                st[AT_C + j] = ok ? 1.f / rowsum[c0 + j] : 0.f;      // 0 kills the padded columns
            }
            // V tile [n1 (128)][c (128)] into wq (Q^T chunk no longer needed)
            const int cw = min(AT_C, C - cp);
            for (int e = t; e < AT_C * AT_C; e += tg::THREADS) {
                const int r = e / AT_C, c = e - r * AT_C;
                wq[e] = (c0 + r < n && c < cw) ? __ldg(v + (long)(c0 + r) * ld + cp + c) : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float4 p;
                const int c = 4 * cg;
                p.x = expf(acc[i][0] - st[c]) * st[AT_C + c];
                p.y = expf(acc[i][1] - st[c + 1]) * st[AT_C + c + 1];
                p.z = expf(acc[i][2] - st[c + 2]) * st[AT_C + c + 2];
                p.w = expf(acc[i][3] - st[c + 3]) * st[AT_C + c + 3];
                *reinterpret_cast<float4 *>(pt + (rg + G::RG * i) * SPT + c) = p;
            }
            __syncthreads();
            if (t < AT_R) {
                const float *row = pt + t * SPT;
                float s = 0.f;
                for (int j = 0; j < AT_C; ++j) s += row[j];
                cs += s;
            }
            tg::fma_block<G::RG>(u, pt, SPT, rg, wq, AT_C, cg, AT_C);
        }
        __syncthreads();
        if (t < AT_R) csum[t] = cs;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = rg + G::RG * i, c = cp + 4 * cg;
            if (r0 + r < n && c < C) {
                const float inv = 1.f / (1e-9f + csum[r]);
                const float4 xv = *reinterpret_cast<const float4 *>(x + (long)(r0 + r) * C + c);
                *reinterpret_cast<float4 *>(d + (long)(r0 + r) * C + c) =
                    make_float4(xv.x - u[i][0] * inv, xv.y - u[i][1] * inv, xv.z - u[i][2] * inv, xv.w - u[i][3] * inv);
            }
        }
    }
}


// ---- whole SA_Layer of a SMALL level in one kernel ------------------------------------------------------------------------
// The deep PPT-Net levels have few points and many channels (64 x 256, 16 x 512): five launches of latency-bound kernels
// (two projections, statistics, apply, trans_conv) cost 0.29 / 0.54 ms for 64 clouds.  Here one CTA owns one cloud and keeps
// x, Q, V (N x C each) and the N x N energies in shared memory: Q = x Wq, V = x Wv + bv (thread = output channel, all N rows in
// registers, weight rows streamed coalesced from L2), E = Q Q^T, row softmax, column renormalisation, U = P^T-weighted V,
// out = x + relu((x - U / (1e-9 + colsum)) Wt + bt).  fp32 SIMT throughout.
constexpr int AS_T = 256;
constexpr int AS_MAXN = 64;

template <int NR>   // rows held per thread in the projections (N <= NR)
__device__ __forceinline__ void as_project(int n, int C, int ld, const float *__restrict__ xs, const float *__restrict__ wt,
                                           const float *__restrict__ shift, float *__restrict__ dst, bool relu_residual,
                                           const float *__restrict__ res) {
    // dst[r][co] = sum_ci xs[r][ci] * wt[ci][co] + shift[co]   (optionally res + relu(.))
    for (int co = threadIdx.x; co < C; co += AS_T) {
        float acc[NR];
        const float sh = shift ? __ldg(shift + co) : 0.f;
#pragma unroll
        for (int r = 0; r < NR; ++r) acc[r] = sh;
        for (int ci0 = 0; ci0 < C; ci0 += 8) {
            float w[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) w[u] = __ldg(wt + (size_t)(ci0 + u) * C + co);      // eight weight rows in flight
            // branch-free over all NR rows (rows >= n are zero and never stored): the NR accumulator chains are independent, so the
            // compiler interleaves them; two broadcast 16-byte loads feed eight FMAs
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                const float4 a0 = *reinterpret_cast<const float4 *>(xs + r * ld + ci0);
                const float4 a1 = *reinterpret_cast<const float4 *>(xs + r * ld + ci0 + 4);
                float a = acc[r];
                a = fmaf(a0.x, w[0], a); a = fmaf(a0.y, w[1], a); a = fmaf(a0.z, w[2], a); a = fmaf(a0.w, w[3], a);
                a = fmaf(a1.x, w[4], a); a = fmaf(a1.y, w[5], a); a = fmaf(a1.z, w[6], a); a = fmaf(a1.w, w[7], a);
                acc[r] = a;
            }
        }
#pragma unroll
        for (int r = 0; r < NR; ++r)
            if (r < n) dst[r * ld + co] = relu_residual ? res[r * ld + co] + fmaxf(acc[r], 0.f) : acc[r];
    }
}

template <int NR>
__global__ void __launch_bounds__(AS_T, 1)
attn_small_kernel(int n, int C, const float *__restrict__ x, const float *__restrict__ wq, const float *__restrict__ wv,
                  const float *__restrict__ bv, const float *__restrict__ wtr, const float *__restrict__ btr, float *__restrict__ out,
                  const float *__restrict__ qv) {
    // qv != NULL ("core" mode): the projections ran on the tensor cores (pw_tc.cu): Q and V are read from the interleaved
    // (rows, 2C) buffer, and the kernel stops after d = x - x_r (written to out); trans_conv follows as another pw_tc launch
    extern __shared__ __align__(16) float sm[];
    const int ld = C + 4;                                   // row stride: rows land in different banks
    float *xs = sm, *qs = xs + NR * ld, *vs = qs + NR * ld, *es = vs + NR * ld;      // es: [NR][NR + 4] energies / probabilities
    float *csum = es + NR * (NR + 4);                       // [NR] column sums
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5, cloud = blockIdx.x;
    x += (size_t)cloud * n * C; out += (size_t)cloud * n * C;
    for (int e = t; e < 3 * NR * ld; e += AS_T) xs[e] = 0.f;              // x, Q, V: rows >= n stay zero
    for (int e = t; e < NR * (NR + 4); e += AS_T) es[e] = 0.f;            // padding columns / rows read as zeros
    __syncthreads();
    for (int e = t; e < n * C; e += AS_T) xs[(e / C) * ld + (e % C)] = __ldg(x + e);
    __syncthreads();
    if (qv) {
        qv += (size_t)cloud * n * 2 * C;
        for (int e = t; e < n * 2 * C; e += AS_T) {
            const int r = e / (2 * C), cc = e - r * 2 * C;
            (cc < C ? qs + r * ld + cc : vs + r * ld + cc - C)[0] = __ldg(qv + e);
        }
    } else {
        as_project<NR>(n, C, ld, xs, wq, nullptr, qs, false, nullptr);
        as_project<NR>(n, C, ld, xs, wv, bv, vs, false, nullptr);
    }
    __syncthreads();
    // energies E[i][j] = Q_i . Q_j
    for (int e = t; e < n * n; e += AS_T) {
        const int i = e / n, j = e - i * n;
        const float4 *qi = reinterpret_cast<const float4 *>(qs + i * ld), *qj = reinterpret_cast<const float4 *>(qs + j * ld);
        float a = 0.f;
        for (int c4 = 0; c4 < C / 4; ++c4) {
            const float4 u = qi[c4], w = qj[c4];
            a = fmaf(u.x, w.x, a); a = fmaf(u.y, w.y, a); a = fmaf(u.z, w.z, a); a = fmaf(u.w, w.w, a);
        }
        es[i * (NR + 4) + j] = a;
    }
    __syncthreads();
    // row softmax (a warp per row), then column sums
    for (int i = warp; i < n; i += AS_T / 32) {
        float m = -INFINITY;
        for (int j = lane; j < n; j += 32) m = fmaxf(m, es[i * (NR + 4) + j]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float s = 0.f;
        for (int j = lane; j < n; j += 32) { const float p = expf(es[i * (NR + 4) + j] - m); es[i * (NR + 4) + j] = p; s += p; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float inv = 1.f / s;
        for (int j = lane; j < n; j += 32) es[i * (NR + 4) + j] *= inv;
    }
    __syncthreads();
    for (int j = t; j < n; j += AS_T) {
        float s = 0.f;
        for (int i = 0; i < n; ++i) s += es[i * (NR + 4) + j];
        csum[j] = 1.f / (1e-9f + s);
    }
    __syncthreads();
    // d[j][c] = x[j][c] - (sum_i V[i][c] P[i][j]) / (1e-9 + colsum_j), written over Q (no longer needed)
    for (int c = t; c < C; c += AS_T) {
        float acc[NR];
#pragma unroll
        for (int j = 0; j < NR; ++j) acc[j] = 0.f;
        for (int i = 0; i < n; ++i) {
            const float v = vs[i * ld + c];
            const float4 *pr = reinterpret_cast<const float4 *>(es + i * (NR + 4));     // columns >= n hold zeros
#pragma unroll
            for (int j4 = 0; j4 < NR / 4; ++j4) {
                const float4 pp = pr[j4];
                acc[4 * j4] = fmaf(v, pp.x, acc[4 * j4]); acc[4 * j4 + 1] = fmaf(v, pp.y, acc[4 * j4 + 1]);
                acc[4 * j4 + 2] = fmaf(v, pp.z, acc[4 * j4 + 2]); acc[4 * j4 + 3] = fmaf(v, pp.w, acc[4 * j4 + 3]);
            }
        }
#pragma unroll
        for (int j = 0; j < NR; ++j)
            if (j < n) qs[j * ld + c] = xs[j * ld + c] - acc[j] * csum[j];
    }
    __syncthreads();
    if (qv) {
        for (int e = t; e < n * C; e += AS_T) out[e] = qs[(e / C) * ld + (e % C)];
        return;
    }
    // out = x + relu(d Wt + bt), staged in V's buffer, then stored coalesced
    as_project<NR>(n, C, ld, qs, wtr, btr, vs, true, xs);
    __syncthreads();
    for (int e = t; e < n * C; e += AS_T) out[e] = vs[(e / C) * ld + (e % C)];
}

size_t as_smem(int nr, int C) { return sizeof(float) * ((size_t)3 * nr * (C + 4) + (size_t)nr * (nr + 4) + nr); }

// returns 0 when the layer was handled, PAB_EINVAL when the shape is not eligible
int launch_attn_small(int b, int n, int c, const float *x, const pab_layer_t *q, const pab_layer_t *v, const pab_layer_t *tr, float *out,
                      cudaStream_t st, const float *qv = nullptr) {
    if (n > AS_MAXN || c % 8 || b > 65535) return PAB_EINVAL;
    const int nr = n <= 16 ? 16 : (n <= 32 ? 32 : 64);
    const size_t smem = as_smem(nr, c);
    if (smem > 220 * 1024) return PAB_EINVAL;
    // the point-wise layers carry (c_in_pad, c_out) fp32 weights with c_in_pad == c (c % 4 == 0)
    if (q->c_in_pad != c || v->c_in_pad != c || tr->c_in_pad != c) return PAB_EINVAL;
#define AS_LAUNCH(NR)                                                                                                        \
    do {                                                                                                                     \
        PAB_CUDA(cudaFuncSetAttribute(attn_small_kernel<NR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
        attn_small_kernel<NR><<<b, AS_T, smem, st>>>(n, c, x, q->wt, v->wt, v->shift, tr->wt, tr->shift, out, qv);           \
    } while (0)
    if (nr == 16) AS_LAUNCH(16); else if (nr == 32) AS_LAUNCH(32); else AS_LAUNCH(64);
#undef AS_LAUNCH
    PAB_LAUNCH_CHECK();
    return 0;
}

inline size_t al(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

int g_attn_small = 1;
PAB_API void pab_tune_attention_small(int on) { g_attn_small = on; }

PAB_API size_t pab_sa_layer_workspace_bytes(int b, int n, int c) {
    return al(sizeof(float) * (size_t)b * n * 2 * c) + al(sizeof(float) * (size_t)b * n * c) + 2 * al(sizeof(float) * (size_t)b * n);
}

PAB_API int pab_sa_layer_forward(int b, int n, int c, const float *x, const pab_layer_t *q_layer, const pab_layer_t *v_layer,
                                 const pab_layer_t *trans_layer, float *out, void *workspace, pab_stream_t s) {
    return pab_sa_layer_forward_p(b, n, c, x, q_layer, v_layer, trans_layer, out, workspace, 2, s);
}

PAB_API int pab_sa_layer_forward_p(int b, int n, int c, const float *x, const pab_layer_t *q_layer, const pab_layer_t *v_layer,
                                   const pab_layer_t *trans_layer, float *out, void *workspace, int precision, pab_stream_t s) {
    if (b < 0 || b > 65535 || n <= 0 || c <= 0 || c % 64 || !q_layer || !v_layer || !trans_layer || !workspace) return PAB_EINVAL;
    if (q_layer->c_in != c || q_layer->c_out != c || v_layer->c_in != c || v_layer->c_out != c || trans_layer->c_in != c ||
        trans_layer->c_out != c) return PAB_EINVAL;
    if (b == 0) return 0;
    cudaStream_t st = (cudaStream_t)s;
    // small deep levels: the whole layer as one kernel — unless the projections can run on the tensor cores (pw_tc.cu), which
    // beats it (g_attn_small = 2 forces the single kernel)
    const bool tc_proj = pab_pw_tc_eligible(q_layer, 2L * c) && pab_pw_tc_eligible(v_layer, 2L * c) && pab_pw_tc_eligible(trans_layer, 0);
    if (g_attn_small && (g_attn_small == 2 || !tc_proj) && launch_attn_small(b, n, c, x, q_layer, v_layer, trans_layer, out, st) == 0) return 0;
    char *w = (char *)workspace;
    float *qv = (float *)w; w += al(sizeof(float) * (size_t)b * n * 2 * c);
    float *d = (float *)w; w += al(sizeof(float) * (size_t)b * n * c);
    float *rmax = (float *)w; w += al(sizeof(float) * (size_t)b * n);
    float *rsum = (float *)w;
    const long ld = 2L * c;                                   // [Q | V] interleaved per point
    int rc;
    // the host may hand q and v as the two halves of ONE (2c, c) weight (planes and shifts contiguous): a single N = 2c product
    const bool merged = tc_proj && (const char *)v_layer->w_hi == (const char *)q_layer->w_hi + (size_t)c * q_layer->tc_k * 2 &&
                        (const char *)v_layer->w_lo == (const char *)q_layer->w_lo + (size_t)c * q_layer->tc_k * 2 &&
                        v_layer->shift == q_layer->shift + c && v_layer->tc_k == q_layer->tc_k && !q_layer->relu && !v_layer->relu;
    if (merged) {
        pab_layer_t qvl = *q_layer;
        qvl.c_out = 2 * c;
        rc = pab_pw_tc_launch((long)b * n, x, &qvl, nullptr, qv, ld, st);
        if (rc) return rc;
    } else {
        rc = pab_pointwise_mlp_residual(b * n, x, q_layer, 1, nullptr, qv, ld, st);
        if (rc) return rc;
        rc = pab_pointwise_mlp_residual(b * n, x, v_layer, 1, nullptr, qv + c, ld, st);
        if (rc) return rc;
    }
    // small levels with tensor-core projections: the attention core of the single-kernel path between them
    if (tc_proj && g_attn_small && launch_attn_small(b, n, c, x, q_layer, v_layer, trans_layer, d, st, qv) == 0)
        return pab_pointwise_mlp_residual(b * n, d, trans_layer, 1, x, out, 0, st);
    if (precision > 0 && pab_attention_tc(b, n, c, qv, qv + c, ld, x, rmax, rsum, d, precision, st) == 0)
        return pab_pointwise_mlp_residual(b * n, d, trans_layer, 1, x, out, 0, st);
    const dim3 grid(pab_divup(n, AT_R), b);
    const size_t smem_a = sizeof(float) * (AT_R * SXQ + AT_KC * AT_C + AT_R * SPT);
    const size_t smem_b = sizeof(float) * (AT_R * SXQ + AT_C * AT_C + AT_R * SPT + 2 * AT_C + AT_R);
    PAB_CUDA(cudaFuncSetAttribute(attn_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a));
    PAB_CUDA(cudaFuncSetAttribute(attn_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
    attn_stats_kernel<<<grid, tg::THREADS, smem_a, st>>>(n, c, qv, ld, rmax, rsum);
    PAB_LAUNCH_CHECK();
    attn_apply_kernel<<<grid, tg::THREADS, smem_b, st>>>(n, c, qv, qv + c, ld, x, rmax, rsum, d);
    PAB_LAUNCH_CHECK();
    return pab_pointwise_mlp_residual(b * n, d, trans_layer, 1, x, out, 0, st);
}
