// bn_train.cu — train-mode BatchNorm + ReLU of the SharedMLP blocks, fused, forward and backward (sm_100a).
//
// Reference: utils/model_util/pt_util.py:98-151 — every SharedMLP block is Conv2d(1x1) -> BatchNorm2d -> ReLU(inplace); in
// training the BatchNorm uses the statistics of the batch, so it cannot be folded into the convolution.  With PyTorch / cuDNN
// the pair costs a statistics pass, a normalise pass, a ReLU pass forward and a ReLU-backward pass plus two BatchNorm-backward
// passes — measured 39 % of the GPU time of a PatchAugNet training step (cuDNN bn_bw_1C11 5.9 ms on the (288,64,1024,20)
// activations of the first SA module).  These kernels are HBM-bound elementwise / reduction passes over x (B, C, S), S = the
// product of the trailing dimensions, written to move the minimum number of bytes:
//   forward   stats:  per-channel sum and sum of squares (fp32 per thread, double across threads and CTAs; fixed tree)
//             apply:  y = relu((x - mean) * invstd * gamma + beta), float4 vectorised                       (read x twice, write y)
//   backward  reduce: dbeta = sum dz, dgamma = sum dz * xhat with dz = dy * (y > 0)                          (read dy, y, x)
//             apply:  dx = gamma * invstd * (dz - dbeta / N - xhat * dgamma / N)                             (read dy, y, x; write dx)
// Reductions are deterministic (static partition, ordered combine).  Running statistics follow nn.BatchNorm (momentum update
// with the UNBIASED batch variance).
#include "common.cuh"

namespace {

constexpr int BT_T = 256;
constexpr int BT_SPLIT_MAX = 64;          // CTAs per channel

// partial[c][split][2] doubles.  Channel c of sample b is the contiguous run x[(b*C + c)*S .. +S).
template <bool BACKWARD>
__global__ void __launch_bounds__(BT_T)
bn_reduce_kernel(int B, int C, long S, int nsplit, const float *__restrict__ x, const float *__restrict__ dy, const float *__restrict__ y,
                 const float *__restrict__ mean, const float *__restrict__ invstd, double *__restrict__ partial) {
    const int c = blockIdx.x, sp = blockIdx.y;
    // the (b, s) index space of the channel is cut into nsplit contiguous slices; with S % 4 == 0 the cuts fall on float4
    // boundaries, so every row segment of a slice is a whole number of 16-byte groups
    const long total = (long)B * S;
    const bool vec = (S & 3) == 0;
    const long unit = vec ? 4 : 1, units = total / unit;
    const long lo = units * sp / nsplit * unit, hi = units * (sp + 1) / nsplit * unit;
    float m = 0.f, is = 0.f;
    if (BACKWARD) { m = __ldg(mean + c); is = __ldg(invstd + c); }
    double a0 = 0.0, a1 = 0.0;
    float f0 = 0.f, f1 = 0.f;
    int cnt = 0;
    auto take = [&](float xv, float dyv, float yv) {
        if (BACKWARD) {
            const float dz = yv > 0.f ? dyv : 0.f;
            f0 += dz;
            f1 += dz * ((xv - m) * is);
        } else {
            f0 += xv;
            f1 += xv * xv;
        }
    };
    for (long b = lo / S; b < B && b * S < hi; ++b) {
        const long s0 = lo > b * S ? lo - b * S : 0, s1 = hi - b * S < S ? hi - b * S : S;
        const long base = (b * C + c) * S;
        if (vec) {
            for (long sidx = s0 + 4 * threadIdx.x; sidx < s1; sidx += 4 * BT_T) {
                const float4 xv = __ldg(reinterpret_cast<const float4 *>(x + base + sidx));
                float4 dv = make_float4(0.f, 0.f, 0.f, 0.f), yv = dv;
                if (BACKWARD) {
                    dv = __ldg(reinterpret_cast<const float4 *>(dy + base + sidx));
                    yv = __ldg(reinterpret_cast<const float4 *>(y + base + sidx));
                }
                take(xv.x, dv.x, yv.x); take(xv.y, dv.y, yv.y); take(xv.z, dv.z, yv.z); take(xv.w, dv.w, yv.w);
                if (++cnt == 16) { a0 += f0; a1 += f1; f0 = f1 = 0.f; cnt = 0; }      // bounded fp32 runs, double beyond
            }
        } else {
            for (long sidx = s0 + threadIdx.x; sidx < s1; sidx += BT_T) {
                take(__ldg(x + base + sidx), BACKWARD ? __ldg(dy + base + sidx) : 0.f, BACKWARD ? __ldg(y + base + sidx) : 0.f);
                if (++cnt == 64) { a0 += f0; a1 += f1; f0 = f1 = 0.f; cnt = 0; }
            }
        }
    }
    a0 += f0; a1 += f1;
    __shared__ double r0[BT_T], r1[BT_T];
    r0[threadIdx.x] = a0; r1[threadIdx.x] = a1;
    __syncthreads();
    for (int o = BT_T / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) { r0[threadIdx.x] += r0[threadIdx.x + o]; r1[threadIdx.x] += r1[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        partial[((size_t)c * nsplit + sp) * 2] = r0[0];
        partial[((size_t)c * nsplit + sp) * 2 + 1] = r1[0];
    }
}

__global__ void bn_finalize_fwd_kernel(int C, int nsplit, double count, const double *__restrict__ partial, float eps, float momentum,
                                       float *__restrict__ mean, float *__restrict__ invstd, float *__restrict__ running_mean,
                                       float *__restrict__ running_var) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double s = 0.0, ss = 0.0;
    for (int i = 0; i < nsplit; ++i) { s += partial[((size_t)c * nsplit + i) * 2]; ss += partial[((size_t)c * nsplit + i) * 2 + 1]; }
    const double mu = s / count;
    double var = ss / count - mu * mu;
    if (var < 0.0) var = 0.0;
    mean[c] = (float)mu;
    invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean) {
        const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
        running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * mu);
        running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unbiased);
    }
}

__global__ void bn_finalize_bwd_kernel(int C, int nsplit, const double *__restrict__ partial, const float *__restrict__ invstd,
                                       float *__restrict__ dgamma, float *__restrict__ dbeta, float *__restrict__ sums) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double s = 0.0, sx = 0.0;
    for (int i = 0; i < nsplit; ++i) { s += partial[((size_t)c * nsplit + i) * 2]; sx += partial[((size_t)c * nsplit + i) * 2 + 1]; }
    dbeta[c] = (float)s;
    dgamma[c] = (float)sx;
    sums[2 * c] = (float)s; sums[2 * c + 1] = (float)sx;
    (void)invstd;
}

// elementwise passes: one thread per float4 (S % 4 == 0) or per element
template <bool BACKWARD, int VEC>
__global__ void __launch_bounds__(BT_T)
bn_apply_kernel(long total_vec, int C, long S, float inv_count, const float *__restrict__ x, const float *__restrict__ dy,
                const float *__restrict__ yin, const float *__restrict__ mean, const float *__restrict__ invstd,
                const float *__restrict__ gamma, const float *__restrict__ beta, const float *__restrict__ sums, float *__restrict__ out) {
    const long v = (long)blockIdx.x * BT_T + threadIdx.x;
    if (v >= total_vec) return;
    const long e = v * VEC;
    const int c = (int)((e / S) % C);
    const float m = __ldg(mean + c), is = __ldg(invstd + c), g = __ldg(gamma + c);
    float xv[VEC], o[VEC];
    if (VEC == 4) *reinterpret_cast<float4 *>(xv) = __ldg(reinterpret_cast<const float4 *>(x + e));
    else xv[0] = __ldg(x + e);
    if (!BACKWARD) {
        const float sc = is * g, sh = __ldg(beta + c) - m * sc;
#pragma unroll
        for (int i = 0; i < VEC; ++i) o[i] = fmaxf(fmaf(xv[i], sc, sh), 0.f);
    } else {
        float dv[VEC], yv[VEC];
        if (VEC == 4) {
            *reinterpret_cast<float4 *>(dv) = __ldg(reinterpret_cast<const float4 *>(dy + e));
            *reinterpret_cast<float4 *>(yv) = __ldg(reinterpret_cast<const float4 *>(yin + e));
        } else { dv[0] = __ldg(dy + e); yv[0] = __ldg(yin + e); }
        const float k0 = __ldg(sums + 2 * c) * inv_count, k1 = __ldg(sums + 2 * c + 1) * inv_count, gs = g * is;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const float dz = yv[i] > 0.f ? dv[i] : 0.f;
            const float xh = (xv[i] - m) * is;
            o[i] = gs * (dz - k0 - xh * k1);
        }
    }
    if (VEC == 4) *reinterpret_cast<float4 *>(out + e) = *reinterpret_cast<const float4 *>(o);
    else out[e] = o[0];
}

int pick_split(int B, int C, long S) {
    const long total = (long)B * S;
    long want = (148L * 8 + C - 1) / C;                     // ~8 CTAs per SM over all channels
    if (want < 1) want = 1;
    if (want > BT_SPLIT_MAX) want = BT_SPLIT_MAX;
    while (want > 1 && total / want < 4 * BT_T) --want;     // at least a few elements per thread
    return (int)want;
}

}  // namespace

PAB_API size_t pab_bn_train_workspace_bytes(int C) { return sizeof(double) * (size_t)C * BT_SPLIT_MAX * 2 + sizeof(float) * (size_t)C * 2; }

// x, y (B, C, S) contiguous; gamma, beta (C); saved mean / invstd (C) out; running_mean / running_var (C) updated in place (or NULL)
PAB_API int pab_bn_relu_train_forward(int B, int C, long S, const float *x, const float *gamma, const float *beta, float eps, float momentum,
                                      float *running_mean, float *running_var, float *mean, float *invstd, float *y, void *workspace,
                                      pab_stream_t s) {
    if (B <= 0 || C <= 0 || C > 65535 || S <= 0 || !workspace) return PAB_EINVAL;
    cudaStream_t st = (cudaStream_t)s;
    double *partial = (double *)workspace;
    const int nsplit = pick_split(B, C, S);
    bn_reduce_kernel<false><<<dim3(C, nsplit), BT_T, 0, st>>>(B, C, S, nsplit, x, nullptr, nullptr, nullptr, nullptr, partial);
    PAB_LAUNCH_CHECK();
    bn_finalize_fwd_kernel<<<pab_divup(C, 128), 128, 0, st>>>(C, nsplit, (double)B * (double)S, partial, eps, momentum, mean, invstd,
                                                               running_mean, running_var);
    PAB_LAUNCH_CHECK();
    const long total = (long)B * C * S;
    if (S % 4 == 0)
        bn_apply_kernel<false, 4><<<pab_divup(total / 4, BT_T), BT_T, 0, st>>>(total / 4, C, S, 0.f, x, nullptr, nullptr, mean, invstd, gamma, beta, nullptr, y);
    else
        bn_apply_kernel<false, 1><<<pab_divup(total, BT_T), BT_T, 0, st>>>(total, C, S, 0.f, x, nullptr, nullptr, mean, invstd, gamma, beta, nullptr, y);
    PAB_LAUNCH_CHECK();
    return 0;
}

// dy, y, x, dx (B, C, S); dgamma, dbeta (C) out
PAB_API int pab_bn_relu_train_backward(int B, int C, long S, const float *dy, const float *y, const float *x, const float *gamma,
                                       const float *mean, const float *invstd, float *dx, float *dgamma, float *dbeta, void *workspace,
                                       pab_stream_t s) {
    if (B <= 0 || C <= 0 || C > 65535 || S <= 0 || !workspace) return PAB_EINVAL;
    cudaStream_t st = (cudaStream_t)s;
    double *partial = (double *)workspace;
    float *sums = (float *)(partial + (size_t)C * BT_SPLIT_MAX * 2);
    const int nsplit = pick_split(B, C, S);
    bn_reduce_kernel<true><<<dim3(C, nsplit), BT_T, 0, st>>>(B, C, S, nsplit, x, dy, y, mean, invstd, partial);
    PAB_LAUNCH_CHECK();
    bn_finalize_bwd_kernel<<<pab_divup(C, 128), 128, 0, st>>>(C, nsplit, partial, invstd, dgamma, dbeta, sums);
    PAB_LAUNCH_CHECK();
    const long total = (long)B * C * S;
    const float inv_count = (float)(1.0 / ((double)B * (double)S));
    if (S % 4 == 0)
        bn_apply_kernel<true, 4><<<pab_divup(total / 4, BT_T), BT_T, 0, st>>>(total / 4, C, S, inv_count, x, dy, y, mean, invstd, gamma, nullptr, sums, dx);
    else
        bn_apply_kernel<true, 1><<<pab_divup(total, BT_T), BT_T, 0, st>>>(total, C, S, inv_count, x, dy, y, mean, invstd, gamma, nullptr, sums, dx);
    PAB_LAUNCH_CHECK();
    return 0;
}
