// gather.cu — index-driven copy / scatter-add ops of libs/pointops in the reference's (b,c,n) layouts.
// These are HBM/L2-bound gathers: one thread per output element, consecutive threads on the contiguous
// output dimension (coalesced stores; index loads coalesced and shared across the channel loop via L1).
#include "common.cuh"

namespace {

// gathering_forward_cuda_kernel, sampling_cuda_kernel.cu:6-19: out[b,c,j] = points[b,c,idx[b,j]]
__global__ void gather_fwd_kernel(int c, int n, int m, const float *__restrict__ points, const int *__restrict__ idx, float *__restrict__ out) {
    const int bi = blockIdx.z, l = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const int a = __ldg(idx + (size_t)bi * m + j);
    out[((size_t)bi * c + l) * m + j] = __ldg(points + ((size_t)bi * c + l) * n + a);
}
// gathering_backward_cuda_kernel, sampling_cuda_kernel.cu:23-36
__global__ void gather_bwd_kernel(int c, int n, int m, const float *__restrict__ grad_out, const int *__restrict__ idx, float *__restrict__ grad_points) {
    const int bi = blockIdx.z, l = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const int a = __ldg(idx + (size_t)bi * m + j);
    atomicAdd(grad_points + ((size_t)bi * c + l) * n + a, __ldg(grad_out + ((size_t)bi * c + l) * m + j));
}

// grouping_forward_cuda_kernel_fast, grouping_cuda_kernel.cu:60-74: out[b,c,p,s] = points[b,c,idx[b,p,s]]
template <typename T>
__global__ void group_fwd_kernel(int c, int n, int ms, const T *__restrict__ points, const int *__restrict__ idx, T *__restrict__ out) {
    const int bi = blockIdx.z, l = blockIdx.y;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;  // p*nsample + s
    if (e >= ms) return;
    const int a = __ldg(idx + (size_t)bi * ms + e);
    out[((size_t)bi * c + l) * ms + e] = __ldg(points + ((size_t)bi * c + l) * n + a);
}
// grouping_backward_cuda_kernel, grouping_cuda_kernel.cu:28-46
__global__ void group_bwd_kernel(int c, int n, int ms, const float *__restrict__ grad_out, const int *__restrict__ idx, float *__restrict__ grad_points) {
    const int bi = blockIdx.z, l = blockIdx.y;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ms) return;
    const int a = __ldg(idx + (size_t)bi * ms + e);
    atomicAdd(grad_points + ((size_t)bi * c + l) * n + a, __ldg(grad_out + ((size_t)bi * c + l) * ms + e));
}

// interpolation_forward_cuda_kernel_fast, interpolation_cuda_kernel.cu:181-195.  The reference expression
// w0*p0 + w1*p1 + w2*p2 is contracted by nvcc to fma(w2,p2, fma(w0,p0, w1*p1)).
__global__ void interp_fwd_kernel(int c, int m, int n, const float *__restrict__ points, const int *__restrict__ idx,
                                  const float *__restrict__ weight, float *__restrict__ out) {
    const int bi = blockIdx.z, l = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const size_t o = ((size_t)bi * n + j) * 3;
    const float *p = points + ((size_t)bi * c + l) * m;
    const float w0 = __ldg(weight + o), w1 = __ldg(weight + o + 1), w2 = __ldg(weight + o + 2);
    const float p0 = __ldg(p + __ldg(idx + o)), p1 = __ldg(p + __ldg(idx + o + 1)), p2 = __ldg(p + __ldg(idx + o + 2));
    out[((size_t)bi * c + l) * n + j] = __fmaf_rn(w2, p2, __fmaf_rn(w0, p0, __fmul_rn(w1, p1)));
}
// interpolation_backward_cuda_kernel, interpolation_cuda_kernel.cu:90-114
__global__ void interp_bwd_kernel(int c, int n, int m, const float *__restrict__ grad_out, const int *__restrict__ idx,
                                  const float *__restrict__ weight, float *__restrict__ grad_points) {
    const int bi = blockIdx.z, l = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const size_t o = ((size_t)bi * n + j) * 3;
    float *gp = grad_points + ((size_t)bi * c + l) * m;
    const float g = __ldg(grad_out + ((size_t)bi * c + l) * n + j);
    atomicAdd(gp + __ldg(idx + o), g * __ldg(weight + o));
    atomicAdd(gp + __ldg(idx + o + 1), g * __ldg(weight + o + 1));
    atomicAdd(gp + __ldg(idx + o + 2), g * __ldg(weight + o + 2));
}

// labelstat_idx_cuda_kernel_fast, labelstat_cuda_kernel.cu:131-151
__global__ void labelstat_idx_kernel(int n, int m, int nsample, int nclass, const int *__restrict__ label_stat,
                                     const int *__restrict__ idx, int *__restrict__ out) {
    const int bi = blockIdx.y;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= m) return;
    int *o = out + ((size_t)bi * m + q) * nclass;
    for (int i = 0; i < nclass; ++i) o[i] = 0;
    for (int k = 0; k < nsample; ++k) {
        const int *ls = label_stat + ((size_t)bi * n + idx[((size_t)bi * m + q) * nsample + k]) * nclass;
        for (int i = 0; i < nclass; ++i) o[i] += ls[i];
    }
}

// point-major row gather: out[b,j,:] = feat[b,idx[b,j],:]   (fused-path helper: new_xyz = xyz[center_idx])
__global__ void gather_rows_kernel(long total, int n, int m, int c, const float *__restrict__ feat, const int *__restrict__ idx, float *__restrict__ out) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const long row = e / c;            // b*m + j
    const int ch = (int)(e - row * c);
    const long bi = row / m;
    out[e] = __ldg(feat + (bi * n + __ldg(idx + row)) * c + ch);
}

bool grid_ok(int b, int c) { return b <= 65535 && c <= 65535; }

}  // namespace

#define PAB_CHECK_BC(b, c) \
    if ((b) < 0 || (c) < 0 || !grid_ok((b), (c))) return PAB_EINVAL;

PAB_API int pab_gathering_forward(int b, int c, int n, int m, const float *points, const int *idx, float *out, pab_stream_t s) {
    PAB_CHECK_BC(b, c);
    if (n <= 0 || m < 0) return PAB_EINVAL;
    if (!b || !c || !m) return 0;
    gather_fwd_kernel<<<dim3(pab_divup(m, 256), c, b), 256, 0, (cudaStream_t)s>>>(c, n, m, points, idx, out);
    PAB_LAUNCH_CHECK();
    return 0;
}
PAB_API int pab_gathering_backward(int b, int c, int n, int m, const float *grad_out, const int *idx, float *grad_points, pab_stream_t s) {
    PAB_CHECK_BC(b, c);
    if (n <= 0 || m < 0) return PAB_EINVAL;
    if (!b || !c || !m) return 0;
    gather_bwd_kernel<<<dim3(pab_divup(m, 256), c, b), 256, 0, (cudaStream_t)s>>>(c, n, m, grad_out, idx, grad_points);
    PAB_LAUNCH_CHECK();
    return 0;
}
PAB_API int pab_grouping_forward(int b, int c, int n, int m, int nsample, const float *points, const int *idx, float *out, pab_stream_t s) {
    PAB_CHECK_BC(b, c);
    if (n <= 0 || m < 0 || nsample < 0) return PAB_EINVAL;
    if (!b || !c || !m || !nsample) return 0;
    group_fwd_kernel<float><<<dim3(pab_divup((long)m * nsample, 256), c, b), 256, 0, (cudaStream_t)s>>>(c, n, m * nsample, points, idx, out);
    PAB_LAUNCH_CHECK();
    return 0;
}
PAB_API int pab_grouping_int_forward(int b, int c, int n, int m, int nsample, const int64_t *points, const int *idx, int64_t *out, pab_stream_t s) {
    PAB_CHECK_BC(b, c);
    if (n <= 0 || m < 0 || nsample < 0) return PAB_EINVAL;
    if (!b || !c || !m || !nsample) return 0;
    group_fwd_kernel<long long><<<dim3(pab_divup((long)m * nsample, 256), c, b), 256, 0, (cudaStream_t)s>>>(
        c, n, m * nsample, (const long long *)points, idx, (long long *)out);
    PAB_LAUNCH_CHECK();
    return 0;
}
PAB_API int pab_grouping_backward(int b, int c, int n, int m, int nsample, const float *grad_out, const int *idx, float *grad_points, pab_stream_t s) {
    PAB_CHECK_BC(b, c);
    if (n <= 0 || m < 0 || nsample < 0) return PAB_EINVAL;
    if (!b || !c || !m || !nsample) return 0;
    group_bwd_kernel<<<dim3(pab_divup((long)m * nsample, 256), c, b), 256, 0, (cudaStream_t)s>>>(c, n, m * nsample, grad_out, idx, grad_points);
    PAB_LAUNCH_CHECK();
    return 0;
}
PAB_API int pab_interpolation_forward(int b, int c, int m, int n, const float *points, const int *idx, const float *weight, float *out, pab_stream_t s) {
    PAB_CHECK_BC(b, c);
    if (n < 0 || m <= 0) return PAB_EINVAL;
    if (!b || !c || !n) return 0;
    interp_fwd_kernel<<<dim3(pab_divup(n, 256), c, b), 256, 0, (cudaStream_t)s>>>(c, m, n, points, idx, weight, out);
    PAB_LAUNCH_CHECK();
    return 0;
}
PAB_API int pab_interpolation_backward(int b, int c, int n, int m, const float *grad_out, const int *idx, const float *weight, float *grad_points, pab_stream_t s) {
    PAB_CHECK_BC(b, c);
    if (n < 0 || m <= 0) return PAB_EINVAL;
    if (!b || !c || !n) return 0;
    interp_bwd_kernel<<<dim3(pab_divup(n, 256), c, b), 256, 0, (cudaStream_t)s>>>(c, n, m, grad_out, idx, weight, grad_points);
    PAB_LAUNCH_CHECK();
    return 0;
}
// featuregather_forward/backward (featuredistribute_cuda_kernel.cu:53-65, 89-101) are the gathering ops with
// the argument order (b, n, m, c).
PAB_API int pab_featuregather_forward(int b, int n, int m, int c, const float *max_feature, const int *distribute_idx, float *distribute_feature, pab_stream_t s) {
    return pab_gathering_forward(b, c, n, m, max_feature, distribute_idx, distribute_feature, s);
}
PAB_API int pab_featuregather_backward(int b, int n, int m, int c, const float *grad_distribute_feature, const int *distribute_idx, float *grad_max_feature, pab_stream_t s) {
    return pab_gathering_backward(b, c, n, m, grad_distribute_feature, distribute_idx, grad_max_feature, s);
}
PAB_API int pab_labelstat_idx(int b, int n, int m, int nsample, int nclass, const int *label_stat, const int *idx, int *new_label_stat, pab_stream_t s) {
    if (b < 0 || b > 65535 || n <= 0 || m < 0 || nsample < 0 || nclass < 0) return PAB_EINVAL;
    if (!b || !m) return 0;
    labelstat_idx_kernel<<<dim3(pab_divup(m, 256), b), 256, 0, (cudaStream_t)s>>>(n, m, nsample, nclass, label_stat, idx, new_label_stat);
    PAB_LAUNCH_CHECK();
    return 0;
}

PAB_API int pab_gather_rows(int b, int n, int m, int c, const float *feat, const int *idx, float *out, pab_stream_t s) {
    if (b < 0 || n <= 0 || m < 0 || c <= 0) return PAB_EINVAL;
    const long total = (long)b * m * c;
    if (total == 0) return 0;
    gather_rows_kernel<<<pab_divup(total, 256), 256, 0, (cudaStream_t)s>>>(total, n, m, c, feat, idx, out);
    PAB_LAUNCH_CHECK();
    return 0;
}
