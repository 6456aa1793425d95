// scatter.cu — DETERMINISTIC backward of the index-driven pointops (gathering / grouping / interpolation), sm_100a.
//
// The reference scatters gradients with fp32 atomicAdd (sampling_cuda_kernel.cu:23-36, grouping_cuda_kernel.cu:28-46,
// interpolation_cuda_kernel.cu:90-114): the summation order, hence the low bits of every gradient, changes from run to
// run.  Here the scatter is turned into a gather: the index tensor of a cloud is inverted ONCE per backward into a CSR
// list (target point -> the entries that read it, ascending), by a bitonic sort of (target << 15 | entry) keys in shared
// memory (one CTA per cloud), and every gradient element is then the sum of its list in that fixed order — one thread
// per (cloud, channel, target), no atomics, bit-identical from run to run.  The same inverse serves all channels.
#include "common.cuh"

namespace {

constexpr int INV_MAX_L = 32768;          // entries per cloud (m * nsample, or 3 n for the interpolation)
constexpr int INV_T = 1024;

// workspace per cloud: offsets[n + 1] | entries[L]
__global__ void __launch_bounds__(INV_T) inverse_index_kernel(int n, int L, int P, const int *__restrict__ idx, int *__restrict__ ws) {
    extern __shared__ unsigned keys[];
    const int t = threadIdx.x;
    const int *id = idx + (size_t)blockIdx.x * L;
    int *offsets = ws + (size_t)blockIdx.x * (n + 1 + L), *entries = offsets + n + 1;
    for (int i = t; i < P; i += INV_T) {
        unsigned key = 0xFFFFFFFFu;
        if (i < L) {
            const int a = __ldg(id + i);
            key = (a >= 0 && a < n) ? (((unsigned)a << 15) | (unsigned)i) : 0xFFFFFFFEu;      // out-of-range targets are dropped
        }
        keys[i] = key;
    }
    __syncthreads();
    for (int size = 2; size <= P; size <<= 1)
        for (int j = size >> 1; j > 0; j >>= 1) {
            for (int i = t; i < (P >> 1); i += INV_T) {
                const int a = 2 * i - (i & (j - 1));
                const unsigned ka = keys[a], kb = keys[a + j];
                const bool up = (a & size) == 0;
                if ((ka > kb) == up) { keys[a] = kb; keys[a + j] = ka; }
            }
            __syncthreads();
        }
    for (int i = t; i < L; i += INV_T) entries[i] = (int)(keys[i] & 32767u);
    // offsets[j] = first sorted position whose target is >= j (binary search; keys of dropped entries sort last)
    for (int j = t; j <= n; j += INV_T) {
        const unsigned want = (unsigned)j << 15;
        int lo = 0, hi = L;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (keys[mid] < want && keys[mid] < 0xFFFFFFFEu) lo = mid + 1; else hi = mid;
        }
        offsets[j] = lo;
    }
}

// grad_points[b,l,j] += sum over the entries e of target j (ascending) of grad_out[b,l,e] * (w ? w[b,e] : 1)
__global__ void scatter_add_csr_kernel(int c, int n, int L, int gdiv, const float *__restrict__ grad_out, const float *__restrict__ w,
                                       const int *__restrict__ ws, float *__restrict__ grad_points) {
    const int bi = blockIdx.z, l = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int *offsets = ws + (size_t)bi * (n + 1 + L), *entries = offsets + n + 1;
    const int p0 = __ldg(offsets + j), p1 = __ldg(offsets + j + 1);
    if (p0 == p1) return;
    const float *g = grad_out + ((size_t)bi * c + l) * (L / gdiv);
    float acc = 0.f;
    for (int p = p0; p < p1; ++p) {
        const int e = __ldg(entries + p);
        const float gv = __ldg(g + (gdiv == 1 ? e : e / gdiv));
        acc += w ? gv * __ldg(w + (size_t)bi * L + e) : gv;
    }
    grad_points[((size_t)bi * c + l) * n + j] += acc;
}

}  // namespace

PAB_API size_t pab_scatter_workspace_bytes(int b, int n, int L) { return sizeof(int) * (size_t)b * ((size_t)n + 1 + (size_t)L); }

// grad_out (b, c, L / gdiv): entry e reads element e / gdiv (gdiv = 3 for the interpolation, whose three neighbours share one
// output gradient; 1 otherwise), idx (b, L) targets in [0, n), weight (b, L) or NULL, grad_points (b, c, n) accumulated into.
// Returns PAB_EINVAL when L exceeds what one CTA sorts in shared memory (the caller then uses the atomic kernels).
PAB_API int pab_scatter_add_deterministic(int b, int c, int n, int L, int gdiv, const float *grad_out, const int *idx, const float *weight,
                                          float *grad_points, void *workspace, pab_stream_t s) {
    return pab_scatter_add_deterministic_ex(b, c, n, L, gdiv, grad_out, idx, weight, grad_points, workspace, 1, s);
}

// build_index = 0: `workspace` already holds the inverse of this idx (an earlier call with the same b, n, L, idx) — the two
// grouping backward passes of an SA module (coordinates and features) share one index tensor, hence one inversion
PAB_API int pab_scatter_add_deterministic_ex(int b, int c, int n, int L, int gdiv, const float *grad_out, const int *idx,
                                             const float *weight, float *grad_points, void *workspace, int build_index, pab_stream_t s) {
    if (gdiv < 1 || L % gdiv) return PAB_EINVAL;
    if (b < 0 || c < 0 || b > 65535 || c > 65535 || n <= 0 || n > (1 << 17) || L < 0 || L > INV_MAX_L || !workspace) return PAB_EINVAL;
    if (!b || !c || !L) return 0;
    cudaStream_t st = (cudaStream_t)s;
    int P = 2;
    while (P < L) P <<= 1;
    const size_t smem = (size_t)P * sizeof(unsigned);
    PAB_CUDA(cudaFuncSetAttribute(inverse_index_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (build_index) {
        inverse_index_kernel<<<b, INV_T, smem, st>>>(n, L, P, idx, (int *)workspace);
        PAB_LAUNCH_CHECK();
    }
    scatter_add_csr_kernel<<<dim3(pab_divup(n, 256), c, b), 256, 0, st>>>(c, n, L, gdiv, grad_out, weight, (const int *)workspace, grad_points);
    PAB_LAUNCH_CHECK();
    return 0;
}
