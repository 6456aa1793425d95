// oracle/ref_shim.cu — TEST INFRASTRUCTURE.  C-ABI doorway into the reference's own CUDA code, compiled
// from /root/reference where it lies (oracle/Makefile `ref`).  Nothing here re-implements the reference:
// the pointops launchers are already extern "C" in the reference headers; this file only adds
//   * ref_knn_device       -> knn_device()            libs/KNN_CUDA/knn_cuda/csrc/cuda/knn.cu:232-269
//   * ref_sync             -> cudaDeviceSynchronize (reference launchers use the legacy default stream)
// by #including the reference translation unit (it has C++ linkage and no header).
#include <cuda_runtime.h>
#include REF_KNN

extern "C" int ref_knn_device(float *ref_dev, int ref_nb, float *query_dev, int query_nb, int dim, int k,
                              float *dist_dev, long *ind_dev) {
    knn_device(ref_dev, ref_nb, query_dev, query_nb, dim, k, dist_dev, ind_dev, 0);
    return (int)cudaDeviceSynchronize();
}

extern "C" int ref_sync(void) { return (int)cudaDeviceSynchronize(); }
