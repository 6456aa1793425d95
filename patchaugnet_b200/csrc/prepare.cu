// prepare.cu — device side of the input pipeline in front of the descriptor path.
//
// Replaces, per batch, the host loop of SceneDataSet.get_pc (datasets/scene_dataset.py:713-740) + normalize_point_cloud
// (utils/loading_pointclouds.py:51-63): raw cloud as stored in the .bin file (float64 or float32, N x 3) - global_offset,
// then optionally centre on the mean and scale by the largest point norm, cast to the float32 (B, N, 3) batch the network
// takes.  The arithmetic stays in float64 like numpy's (the cast to float32 is the last step, as torch's .float() in
// make_descs); sums use a fixed tree, so results are deterministic — they can differ from numpy's pairwise float64 sum in
// the last float64 bit, i.e. essentially never after the cast.
//
// One CTA per cloud: coalesced read of the raw cloud (kept in registers when it fits), block reduction of the coordinate sums,
// block reduction of the largest squared norm, scaled write.
#include "common.cuh"

namespace {

constexpr int PR_THREADS = 1024;
constexpr int PR_MAX_PER_THREAD = 8;            // points per thread kept in registers (n <= 8192)

template <typename T>
__global__ void __launch_bounds__(PR_THREADS) prepare_kernel(int n, const T *__restrict__ raw, double ox, double oy, double oz,
                                                             int normalize, int zoom, float *__restrict__ out,
                                                             double *__restrict__ meta) {
    __shared__ double red[4][32];
    __shared__ double bc[4];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5, cloud = blockIdx.x;
    const T *src = raw + (size_t)cloud * n * 3;
    float *dst = out + (size_t)cloud * n * 3;
    double px[PR_MAX_PER_THREAD], py[PR_MAX_PER_THREAD], pz[PR_MAX_PER_THREAD];
    double sx = 0.0, sy = 0.0, sz = 0.0;
#pragma unroll
    for (int i = 0; i < PR_MAX_PER_THREAD; ++i) {
        const int p = t + i * PR_THREADS;
        px[i] = py[i] = pz[i] = 0.0;
        if (p < n) {
            px[i] = (double)src[3 * p] - ox; py[i] = (double)src[3 * p + 1] - oy; pz[i] = (double)src[3 * p + 2] - oz;
            sx += px[i]; sy += py[i]; sz += pz[i];
        }
    }
    auto block_reduce3 = [&](double &a, double &b, double &c, bool is_max) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double a2 = __shfl_xor_sync(0xffffffffu, a, o), b2 = __shfl_xor_sync(0xffffffffu, b, o), c2 = __shfl_xor_sync(0xffffffffu, c, o);
            if (is_max) { a = fmax(a, a2); } else { a += a2; b += b2; c += c2; }
        }
        __syncthreads();
        if (lane == 0) { red[0][warp] = a; red[1][warp] = b; red[2][warp] = c; }
        __syncthreads();
        if (warp == 0) {
            double x = red[0][lane], y = red[1][lane], z = red[2][lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double x2 = __shfl_xor_sync(0xffffffffu, x, o), y2 = __shfl_xor_sync(0xffffffffu, y, o), z2 = __shfl_xor_sync(0xffffffffu, z, o);
                if (is_max) { x = fmax(x, x2); } else { x += x2; y += y2; z += z2; }
            }
            if (lane == 0) { bc[0] = x; bc[1] = y; bc[2] = z; }
        }
        __syncthreads();
        a = bc[0]; b = bc[1]; c = bc[2];
    };
    double cx = 0.0, cy = 0.0, cz = 0.0, m = 1.0;
    if (normalize) {
        block_reduce3(sx, sy, sz, false);
        cx = sx / n; cy = sy / n; cz = sz / n;                       // np.mean(pc, axis=0)
        double q = 0.0, u0 = 0.0, u1 = 0.0;
#pragma unroll
        for (int i = 0; i < PR_MAX_PER_THREAD; ++i) {
            const int p = t + i * PR_THREADS;
            if (p < n) {
                px[i] -= cx; py[i] -= cy; pz[i] -= cz;
                q = fmax(q, px[i] * px[i] + py[i] * py[i] + pz[i] * pz[i]);   // np.sum(pc ** 2, axis=1): (x^2 + y^2) + z^2
            }
        }
        if (zoom) {
            block_reduce3(q, u0, u1, true);
            m = sqrt(q);                                             // np.max(np.sqrt(.)) == sqrt(max(.))
        }
    }
#pragma unroll
    for (int i = 0; i < PR_MAX_PER_THREAD; ++i) {
        const int p = t + i * PR_THREADS;
        if (p < n) {
            dst[3 * p] = (float)(normalize && zoom ? px[i] / m : px[i]);
            dst[3 * p + 1] = (float)(normalize && zoom ? py[i] / m : py[i]);
            dst[3 * p + 2] = (float)(normalize && zoom ? pz[i] / m : pz[i]);
        }
    }
    if (meta && t == 0) { meta[4 * cloud] = m; meta[4 * cloud + 1] = cx; meta[4 * cloud + 2] = cy; meta[4 * cloud + 3] = cz; }
}

}  // namespace

PAB_API int pab_prepare_clouds(int b, int n, const void *raw, int raw_is_f64, const double *offset, int normalize, int zoom, float *out,
                               double *meta, pab_stream_t s) {
    if (b < 0 || n <= 0 || n > PR_THREADS * PR_MAX_PER_THREAD || !raw || !out) return PAB_EINVAL;
    if (b == 0) return 0;
    const double ox = offset ? offset[0] : 0.0, oy = offset ? offset[1] : 0.0, oz = offset ? offset[2] : 0.0;
    if (raw_is_f64)
        prepare_kernel<double><<<b, PR_THREADS, 0, (cudaStream_t)s>>>(n, (const double *)raw, ox, oy, oz, normalize, zoom, out, meta);
    else
        prepare_kernel<float><<<b, PR_THREADS, 0, (cudaStream_t)s>>>(n, (const float *)raw, ox, oy, oz, normalize, zoom, out, meta);
    PAB_LAUNCH_CHECK();
    return 0;
}
