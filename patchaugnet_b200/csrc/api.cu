// api.cu — library info entry points of libpatchaug_b200.so.
#include "common.cuh"

int g_pab_launches = 0;

PAB_API int pab_version(void) { return 1; }
PAB_API int pab_num_launches(void) { return g_pab_launches; }
PAB_API void pab_reset_launch_counter(void) { g_pab_launches = 0; }

// ---- self-resetting tile counters of the persistent kernels (mlp_tc.cu dynamic tiles, sa_narrow_tc.cu) --------------------------
// A pair = {next tile, CTAs finished}; the last CTA of a launch zeroes both, so a pair can be handed to the next launch without a
// memset.  Eager launches rotate over 256 pairs per device (two launches that share a pair are 256 dynamic launches apart: never
// in flight together).  A launch that is being CAPTURED into a CUDA graph keeps its pair for as long as the graph lives and may
// replay next to any eager launch, so it gets a pair of its own from a second pool that is never recycled; when that pool is
// exhausted the function returns NULL and the caller schedules statically / rotates.  Pools are allocated on the first call per
// device, which must not happen during a capture (the engines warm up before they capture).
unsigned int *pab_tile_counter_pair(cudaStream_t st) {
    constexpr int MAX_DEV = 16, N_EAGER = 256, N_GRAPH = 4096;
    static unsigned int *eager_pool[MAX_DEV] = {}, *graph_pool[MAX_DEV] = {};
    static unsigned int eager_seq[MAX_DEV] = {}, graph_next[MAX_DEV] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEV) return nullptr;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) return nullptr;
    if (!eager_pool[dev]) {
        if (cs != cudaStreamCaptureStatusNone) return nullptr;
        unsigned int *p = nullptr;
        const size_t bytes = 2 * (size_t)(N_EAGER + N_GRAPH) * sizeof(unsigned int);
        if (cudaMalloc(&p, bytes) != cudaSuccess || cudaMemset(p, 0, bytes) != cudaSuccess) return nullptr;
        eager_pool[dev] = p;
        graph_pool[dev] = p + 2 * N_EAGER;
    }
    if (cs != cudaStreamCaptureStatusNone) {
        if (graph_next[dev] >= (unsigned)N_GRAPH) return nullptr;
        return graph_pool[dev] + 2 * graph_next[dev]++;
    }
    return eager_pool[dev] + 2 * (eager_seq[dev]++ % N_EAGER);
}
