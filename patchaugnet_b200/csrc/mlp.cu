// mlp.cu — fused SharedMLP kernels (set abstraction, feature propagation, plain point-wise) for sm_100a.
//
// The reference runs, per SharedMLP layer, a cuDNN 1x1 conv + BatchNorm2d + ReLU as three kernels
// (utils/model_util/pt_util.py:98-151), after materialising the grouped (B,C+3,M,K) tensor with two grouping
// launches, two subtractions and a cat (libs/pointops/functions/pointops.py:559-570), and follows with a separate
// max_pool2d over K (patch_aug_net.py:236).  Here one kernel per module does: neighbour gather + centre subtract +
// concat (or 3-NN interpolation + skip concat) straight into a shared-memory tile, every layer of the MLP with the
// tile resident in shared memory, and the max over K (or a coalesced row store).  Only (B, M, C_out) leaves the SM.
//
// Layouts: xyz (b,n,3); features POINT-MAJOR (b,n,c); outputs point-major.  fp32 throughout (tile_gemm.cuh).
#include "tile_gemm.cuh"

// tensor-core path (mlp_tc.cu)
int pab_tc_eligible(const pab_layer_t *layers, int n_layers, int k_group, int allow_pre);
int pab_sa_narrow_eligible(const pab_layer_t *layers, int n_layers, int k);
int pab_sa_narrow_launch(int b, int n, int m, int k, int nbr_stride, int c, const float *xyz, const float *feat, const int *center_idx,
                         const int *nbr_idx, const pab_layer_t *layers, int n_layers, float *out, cudaStream_t st);
int pab_tc_sa(int kind, int b, int n, int m, int k, int nbr_stride, int c, const float *xyz, const float *feat, const int *center_idx,
              const int *nbr_idx, const pab_layer_t *layers, int n_layers, float *out, cudaStream_t st);
int pab_tc_fp(int b, int n, int m, int c_known, int c_skip, const float *known_feat, const float *skip_feat, const int *idx,
              const float *weight, const int *row_order, long order_stride, const pab_layer_t *layers, int n_layers, float *out,
              cudaStream_t st);

namespace {

enum { MODE_PLAIN = 0, MODE_SA = 1, MODE_FP = 2 };

struct MlpArgs {
    pab_layer_t L[4];
    int n_layers;
    int mode;
    long rows;          // PLAIN: rows; SA: b*m centres; FP: b*n points
    int sA, sB;         // smem row strides
    // PLAIN
    const float *x;
    const float *residual;   // optional (rows, c_last): added to the last layer's output
    long out_ld;             // PLAIN: row stride of `out` in floats (0 = c_last)
    // SA
    int n, m, k, nbr_stride, c;
    const float *xyz, *feat;
    const int *center_idx, *nbr_idx;
    float *new_xyz;
    // FP
    int c_known, c_skip;
    const float *known_feat, *skip_feat;
    const int *idx3;
    const float *w3;
    float *out;
};

template <int R>
__global__ void __launch_bounds__(tg::THREADS, 1) mlp_kernel(const MlpArgs a) {
    extern __shared__ __align__(16) float smem[];
    float *bufA = smem;
    float *bufB = bufA + (size_t)R * a.sA;
    float *wstage = bufB + (size_t)R * a.sB;
    __shared__ long meta0[R];            // SA: neighbour point offset; FP: known base offset (cloud * m)
    __shared__ long meta1[R];            // SA: centre point offset;    FP: global row
    __shared__ int meta_i[R][3];         // FP: 3-NN indices
    __shared__ float meta_w[R][3];       // FP: 3-NN weights

    const int t = threadIdx.x;
    const int c0p = a.L[0].c_in_pad, c0 = a.L[0].c_in;
    const int sA = a.sA, sB = a.sB;

    // ---- stage the layer-0 input tile into bufA -------------------------------------------------------
    int groups_per_tile = 0;
    long first = 0;
    if (a.mode == MODE_SA) {
        groups_per_tile = R / a.k;
        first = (long)blockIdx.x * groups_per_tile;  // first centre of this tile
        for (int r = t; r < R; r += tg::THREADS) {
            const int g = r / a.k, s = r - g * a.k;
            const long ci = first + g;
            long pn = -1, pc = -1;
            if (g < groups_per_tile && ci < a.rows) {
                const long cloud = ci / a.m;
                pc = cloud * a.n + __ldg(a.center_idx + ci);
                pn = cloud * a.n + __ldg(a.nbr_idx + ci * a.nbr_stride + s);
            }
            meta0[r] = pn; meta1[r] = pc;
        }
        __syncthreads();
        for (int e = t; e < R * c0p; e += tg::THREADS) {
            const int r = e / c0p, ch = e - r * c0p;
            const long pn = meta0[r], pc = meta1[r];
            float v = 0.f;
            if (pn >= 0 && ch < c0) {
                if (ch < 3) v = __ldg(a.xyz + pn * 3 + ch) - __ldg(a.xyz + pc * 3 + ch);
                else v = __ldg(a.feat + pn * a.c + (ch - 3)) - __ldg(a.feat + pc * a.c + (ch - 3));
            }
            bufA[r * sA + ch] = v;
        }
        if (a.new_xyz) {
            for (int e = t; e < groups_per_tile * 3; e += tg::THREADS) {
                const int g = e / 3, ch = e - 3 * g;
                const long ci = first + g;
                if (ci < a.rows) a.new_xyz[ci * 3 + ch] = __ldg(a.xyz + meta1[g * a.k] * 3 + ch);
            }
        }
    } else if (a.mode == MODE_FP) {
        first = (long)blockIdx.x * R;
        for (int r = t; r < R; r += tg::THREADS) {
            const long j = first + r;
            if (j < a.rows) {
                const long cloud = j / a.n;
                meta0[r] = cloud * a.m; meta1[r] = j;
#pragma unroll
                for (int q = 0; q < 3; ++q) { meta_i[r][q] = __ldg(a.idx3 + j * 3 + q); meta_w[r][q] = __ldg(a.w3 + j * 3 + q); }
            } else {
                meta0[r] = -1; meta1[r] = -1;
            }
        }
        __syncthreads();
        for (int e = t; e < R * c0p; e += tg::THREADS) {
            const int r = e / c0p, ch = e - r * c0p;
            const long base = meta0[r];
            float v = 0.f;
            if (base >= 0 && ch < c0) {
                if (ch < a.c_known) {
                    // interpolation_forward: fma(w2,p2, fma(w0,p0, w1*p1)), interpolation_cuda_kernel.cu:194
                    const float p0 = __ldg(a.known_feat + (base + meta_i[r][0]) * a.c_known + ch);
                    const float p1 = __ldg(a.known_feat + (base + meta_i[r][1]) * a.c_known + ch);
                    const float p2 = __ldg(a.known_feat + (base + meta_i[r][2]) * a.c_known + ch);
                    v = __fmaf_rn(meta_w[r][2], p2, __fmaf_rn(meta_w[r][0], p0, __fmul_rn(meta_w[r][1], p1)));
                } else {
                    v = __ldg(a.skip_feat + meta1[r] * a.c_skip + (ch - a.c_known));
                }
            }
            bufA[r * sA + ch] = v;
        }
    } else {
        first = (long)blockIdx.x * R;
        for (int e = t; e < R * c0p; e += tg::THREADS) {
            const int r = e / c0p, ch = e - r * c0p;
            const long j = first + r;
            bufA[r * sA + ch] = (j < a.rows && ch < c0) ? __ldg(a.x + j * c0 + ch) : 0.f;
        }
    }
    // (the first __syncthreads inside tg::layer orders these writes before any read)

    // ---- the MLP, tile resident ----------------------------------------------------------------------
    float *X = bufA, *Y = bufB;
    int sx = sA, sy = sB;
    for (int l = 0; l < a.n_layers; ++l) {
        tg::layer<R>(X, sx, Y, sy, a.L[l], wstage);
        float *tp = X; X = Y; Y = tp;
        int ts = sx; sx = sy; sy = ts;
    }
    __syncthreads();
    // result tile is X (stride sx), c_last channels
    const int cl = a.L[a.n_layers - 1].c_out;
    if (a.mode == MODE_SA) {
        for (int e = t; e < groups_per_tile * cl; e += tg::THREADS) {
            const int g = e / cl, ch = e - g * cl;
            const long ci = first + g;
            if (ci >= a.rows) continue;
            const float *col = X + (size_t)(g * a.k) * sx + ch;
            float v = col[0];
            for (int s = 1; s < a.k; ++s) v = fmaxf(v, col[(size_t)s * sx]);
            a.out[ci * cl + ch] = v;
        }
    } else {
        for (int e = t; e < R * cl; e += tg::THREADS) {
            const int r = e / cl, ch = e - r * cl;
            const long j = first + r;
            const long old = a.out_ld ? a.out_ld : cl;
            if (j < a.rows) a.out[j * old + ch] = X[(size_t)r * sx + ch] + (a.residual ? __ldg(a.residual + j * cl + ch) : 0.f);
        }
    }
}

struct Plan { int R; int sA, sB; size_t smem; long tiles; };

// Largest tile that fits the 227 KB opt-in shared memory limit.
int make_plan(const pab_layer_t *layers, int n_layers, long rows, int k_group, Plan *p) {
    if (n_layers < 1 || n_layers > 4) return PAB_EINVAL;
    int wa = layers[0].c_in_pad, wb = 0;
    for (int l = 0; l < n_layers; ++l) {
        if (layers[l].c_out % 4 || layers[l].c_in_pad % 4 || layers[l].c_in > layers[l].c_in_pad) return PAB_EINVAL;
        if (l > 0 && layers[l].c_in != layers[l - 1].c_out) return PAB_EINVAL;
        int &w = (l % 2 == 0) ? wb : wa;  // layer l writes B when l even, A when l odd
        if (layers[l].c_out > w) w = layers[l].c_out;
    }
    const int Rs[4] = {256, 128, 64, 32};
    for (int i = 0; i < 4; ++i) {
        const int R = Rs[i];
        if (k_group > R) continue;
        const int sA = tg::stride_for(wa), sB = tg::stride_for(wb);
        const size_t smem = sizeof(float) * ((size_t)R * sA + (size_t)R * sB + 2 * tg::KC * (8192 / R));
        // static smem: meta arrays  R*(8+8+12+12) bytes
        if (smem + (size_t)R * 40 + 1024 <= 227 * 1024) {
            p->R = R; p->sA = sA; p->sB = sB; p->smem = smem;
            const long per = k_group > 0 ? R / k_group : R;
            p->tiles = (rows + per - 1) / per;
            return 0;
        }
    }
    return PAB_EINVAL;
}

template <int R>
int launch(const MlpArgs &a, const Plan &p, cudaStream_t st) {
    PAB_CUDA(cudaFuncSetAttribute(mlp_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
    mlp_kernel<R><<<(unsigned)p.tiles, tg::THREADS, p.smem, st>>>(a);
    PAB_LAUNCH_CHECK();
    return 0;
}

int run(MlpArgs &a, const pab_layer_t *layers, int n_layers, int k_group, cudaStream_t st) {
    Plan p;
    int rc = make_plan(layers, n_layers, a.rows, k_group, &p);
    if (rc) return rc;
    for (int l = 0; l < n_layers; ++l) a.L[l] = layers[l];
    a.n_layers = n_layers; a.sA = p.sA; a.sB = p.sB;
    if (a.rows == 0) return 0;
    switch (p.R) {
        case 256: return launch<256>(a, p, st);
        case 128: return launch<128>(a, p, st);
        case 64: return launch<64>(a, p, st);
        default: return launch<32>(a, p, st);
    }
}

}  // namespace

PAB_API int pab_sa_module_forward(int b, int n, int m, int k, int nbr_stride, int c, const float *xyz, const float *feat,
                                  const int *center_idx, const int *nbr_idx, const pab_layer_t *layers, int n_layers,
                                  float *out, float *new_xyz, pab_stream_t s) {
    if (b < 0 || n <= 0 || m < 0 || k <= 0 || k > 256 || nbr_stride < k || c < 0 || !layers) return PAB_EINVAL;
    if (layers[0].c_in != c + 3) return PAB_EINVAL;
    if (!new_xyz) {
        const int kind = pab_tc_eligible(layers, n_layers, k, c <= 5);
        if (kind == 2 && pab_sa_narrow_eligible(layers, n_layers, k)) {     // tiny input, layers <= 64 wide: small CTAs, several per SM
            const int rc = pab_sa_narrow_launch(b, n, m, k, nbr_stride, c, xyz, feat, center_idx, nbr_idx, layers, n_layers, out, (cudaStream_t)s);
            if (rc != PAB_EINVAL) return rc;                                // EINVAL: no tile counter left for a graph capture
        }
        if ((kind == 1 && c % 8 == 0 && layers[0].tc_k == c && layers[0].tc_k0 == 3) || kind == 2)
            return pab_tc_sa(kind, b, n, m, k, nbr_stride, c, xyz, feat, center_idx, nbr_idx, layers, n_layers, out, (cudaStream_t)s);
    }
    MlpArgs a{};
    a.mode = MODE_SA; a.rows = (long)b * m; a.n = n; a.m = m; a.k = k; a.nbr_stride = nbr_stride; a.c = c;
    a.xyz = xyz; a.feat = feat; a.center_idx = center_idx; a.nbr_idx = nbr_idx; a.new_xyz = new_xyz; a.out = out;
    return run(a, layers, n_layers, k, (cudaStream_t)s);
}

// row_order (optional): order[cloud * order_stride + i] = the point of the cloud processed as its i-th row; the tensor-core
// kernel uses it for locality only (every point is computed once and stored at its own position: results are bit-identical)
PAB_API int pab_fp_module_forward_ordered(int b, int n, int m, int c_known, int c_skip, const float *known_feat, const float *skip_feat,
                                          const int *idx, const float *weight, const int *row_order, long order_stride,
                                          const pab_layer_t *layers, int n_layers, float *out, pab_stream_t s) {
    if (b < 0 || n < 0 || m <= 0 || c_known <= 0 || c_skip < 0 || !layers) return PAB_EINVAL;
    if (layers[0].c_in != c_known + c_skip || (c_skip > 0 && !skip_feat)) return PAB_EINVAL;
    if (c_known % 8 == 0 && pab_tc_eligible(layers, n_layers, 0, 0) == 1 && layers[0].tc_k0 == 0 &&
        ((layers[0].tc_k == c_known && c_skip <= 3) || (layers[0].tc_k == c_known + c_skip && c_skip % 8 == 0)))
        return pab_tc_fp(b, n, m, c_known, c_skip, known_feat, skip_feat, idx, weight, row_order, order_stride, layers, n_layers, out,
                         (cudaStream_t)s);
    MlpArgs a{};
    a.mode = MODE_FP; a.rows = (long)b * n; a.n = n; a.m = m; a.c_known = c_known; a.c_skip = c_skip;
    a.known_feat = known_feat; a.skip_feat = skip_feat; a.idx3 = idx; a.w3 = weight; a.out = out;
    return run(a, layers, n_layers, 0, (cudaStream_t)s);
}

PAB_API int pab_fp_module_forward(int b, int n, int m, int c_known, int c_skip, const float *known_feat, const float *skip_feat,
                                  const int *idx, const float *weight, const pab_layer_t *layers, int n_layers,
                                  float *out, pab_stream_t s) {
    return pab_fp_module_forward_ordered(b, n, m, c_known, c_skip, known_feat, skip_feat, idx, weight, nullptr, 0, layers, n_layers, out, s);
}

int pab_pw_tc_eligible(const pab_layer_t *L, long out_ld);
int pab_pw_tc_launch(long rows, const float *x, const pab_layer_t *L, const float *residual, float *out, long out_ld, cudaStream_t st);

int pab_pointwise_mlp_residual(int rows, const float *x, const pab_layer_t *layers, int n_layers, const float *residual,
                               float *out, long out_ld, cudaStream_t st) {
    if (rows < 0 || !layers) return PAB_EINVAL;
    if (n_layers == 1 && pab_pw_tc_eligible(&layers[0], out_ld))       // one wide layer with bf16 hi/lo planes: tensor cores (pw_tc.cu)
        return pab_pw_tc_launch(rows, x, &layers[0], residual, out, out_ld, st);
    MlpArgs a{};
    a.mode = MODE_PLAIN; a.rows = rows; a.x = x; a.residual = residual; a.out = out; a.out_ld = out_ld;
    return run(a, layers, n_layers, 0, st);
}

PAB_API int pab_pointwise_mlp_forward(int rows, const float *x, const pab_layer_t *layers, int n_layers, float *out, pab_stream_t s) {
    return pab_pointwise_mlp_residual(rows, x, layers, n_layers, nullptr, out, 0, (cudaStream_t)s);
}
