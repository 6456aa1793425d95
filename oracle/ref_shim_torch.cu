// oracle/ref_shim_torch.cu — TEST INFRASTRUCTURE.  The reference's chamfer.cu and emd_cuda.cu include
// torch headers and wrap their kernels in at::Tensor host functions; this shim #includes those translation
// units unchanged (compiled against the image's torch headers) and launches the reference __global__
// kernels on raw pointers with the grid/block shapes of the reference host code:
//   chamfer fwd  dim3(32,16,1) x 512, twice            libs/chamfer_dist/chamfer.cu:159-164
//   chamfer bwd  dim3(1,16,1) x 256, twice             libs/chamfer_dist/chamfer.cu:215-222
//   emd fwd      iters x {clear, calc_unass_cnt, calc_unass_cnt_sum, calc_unass_idx, Bid, GetMax, Assign}
//                + CalcDist                            libs/emd_module/emd_cuda.cu:256-270
//   emd bwd      NmDistanceGradKernel                  libs/emd_module/emd_cuda.cu:305
#include REF_CHAMFER
namespace refemd {
#include REF_EMD
}

extern "C" int ref_chamfer_forward(int b, int n, const float *xyz1, int m, const float *xyz2,
                                   float *dist1, float *dist2, int *idx1, int *idx2) {
    chamfer_dist_kernel<<<dim3(32, 16, 1), 512>>>(b, n, xyz1, m, xyz2, dist1, idx1);
    chamfer_dist_kernel<<<dim3(32, 16, 1), 512>>>(b, m, xyz2, n, xyz1, dist2, idx2);
    return (int)cudaDeviceSynchronize();
}

extern "C" int ref_chamfer_backward(int b, int n, const float *xyz1, int m, const float *xyz2,
                                    const int *idx1, const int *idx2, const float *g1, const float *g2,
                                    float *grad_xyz1, float *grad_xyz2) {
    chamfer_dist_grad_kernel<<<dim3(1, 16, 1), 256>>>(b, n, xyz1, m, xyz2, g1, idx1, grad_xyz1, grad_xyz2);
    chamfer_dist_grad_kernel<<<dim3(1, 16, 1), 256>>>(b, m, xyz2, n, xyz1, g2, idx2, grad_xyz2, grad_xyz1);
    return (int)cudaDeviceSynchronize();
}

extern "C" int ref_emd_forward(int b, int n, float *xyz1, float *xyz2, float *dist, int *assignment, float *price,
                               int *assignment_inv, int *bid, float *bid_increments, float *max_increments,
                               int *unass_idx, int *unass_cnt, int *unass_cnt_sum, int *cnt_tmp, int *max_idx,
                               float eps, int iters) {
    using namespace refemd;
    if (b > 512 || n % 1024 != 0) return -1;
    for (int i = 0; i < iters; i++) {
        clear<<<1, b>>>(b, cnt_tmp, unass_cnt);
        calc_unass_cnt<<<dim3(b, n / 1024, 1), 1024>>>(b, n, assignment, unass_cnt);
        calc_unass_cnt_sum<<<1, b>>>(b, unass_cnt, unass_cnt_sum);
        calc_unass_idx<<<dim3(b, n / 1024, 1), 1024>>>(b, n, assignment, unass_idx, unass_cnt, unass_cnt_sum, cnt_tmp);
        Bid<<<dim3(b, n / 1024, 1), 1024>>>(b, n, xyz1, xyz2, eps, assignment, assignment_inv, price, bid,
                                            bid_increments, max_increments, unass_cnt, unass_cnt_sum, unass_idx);
        GetMax<<<dim3(b, n / 1024, 1), 1024>>>(b, n, assignment, bid, bid_increments, max_increments, max_idx);
        Assign<<<dim3(b, n / 1024, 1), 1024>>>(b, n, assignment, assignment_inv, price, bid, bid_increments,
                                               max_increments, max_idx, i == iters - 1);
    }
    CalcDist<<<dim3(b, n / 1024, 1), 1024>>>(b, n, xyz1, xyz2, dist, assignment);
    return cudaDeviceSynchronize() == cudaSuccess ? 1 : 0;
}

extern "C" int ref_emd_backward(int b, int n, const float *xyz1, const float *xyz2, float *gradxyz,
                                const float *graddist, const int *idx) {
    refemd::NmDistanceGradKernel<<<dim3(b, n / 1024, 1), 1024>>>(b, n, xyz1, xyz2, graddist, idx, gradxyz);
    return cudaDeviceSynchronize() == cudaSuccess ? 1 : 0;
}
