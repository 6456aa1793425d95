/*
 * patchaug_b200.h — C ABI of libpatchaug_b200.so (sm_100a).
 *
 * Drop-in boundary for the descriptor-extraction-and-retrieval hot path of WHU-USI3DV/PatchAugNet.
 * Every entry point takes plain device pointers + sizes + a CUDA stream (as void*) and returns an int:
 *   0            success
 *   > 0          a cudaError_t from the launch
 *   PAB_EINVAL   argument outside the supported range (the reference would overflow / exit)
 * Nothing is retained between calls; the caller owns every buffer.  Citations are relative to the
 * reference tree (/root/reference).
 *
 * The "pab_<name>" pointops functions replace, one for one, the extern "C" launchers the reference's
 * pybind module `pointops_cuda` calls (libs/pointops/src/pointops_api.cpp:15-40), with two deliberate
 * differences: every launch goes to the caller's stream (most reference launchers use the legacy default
 * stream), and failures are returned instead of exit(-1) (e.g. knnquery_cuda_kernel.cu:66-70).
 */
#ifndef PATCHAUG_B200_H
#define PATCHAUG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PAB_EINVAL (-22)

typedef void *pab_stream_t; /* cudaStream_t */

/* ---- library info -------------------------------------------------------------------------------- */
int pab_version(void);                 /* ABI version, bumped on any signature change */
int pab_num_launches(void);            /* kernels launched by this library since load (bench.py gpu_launches) */
void pab_reset_launch_counter(void);

/* ---- libs/pointops ------------------------------------------------------------------------------- */

/* furthestsampling_cuda_launcher  sampling/sampling_cuda_kernel.h:17, kernel .cu:58-168.
 * xyz (b,n,3) f32; temp (b,n) f32 caller-initialised (1e10, pointops.py:21), updated in place;
 * idx (b,m) i32 out.  Bit-exact with the reference including its tie-break order. */
int pab_furthestsampling(int b, int n, int m, const float *xyz, float *temp, int *idx, pab_stream_t s);

/* gathering_forward/backward_cuda_launcher  sampling_cuda_kernel.h:15-16.  points (b,c,n), idx (b,m),
 * out (b,c,m); backward accumulates into grad_points (b,c,n) (caller-zeroed, pointops.py:52). */
int pab_gathering_forward(int b, int c, int n, int m, const float *points, const int *idx, float *out, pab_stream_t s);
int pab_gathering_backward(int b, int c, int n, int m, const float *grad_out, const int *idx, float *grad_points, pab_stream_t s);

/* knnquery_cuda_launcher  knnquery/knnquery_cuda_kernel.h:14, kernel .cu:6-50.  xyz (b,n,3), new_xyz (b,m,3);
 * idx (b,m,nsample) i32 ascending by distance, ties to the lower index; dist2 (b,m,nsample) f32 or NULL
 * (the reference's dist2 write is an un-offset race, .cu:44-47; here it is the true squared distance).
 * nsample <= 200 like the reference's fixed arrays (.cu:21-22), else PAB_EINVAL. */
int pab_knnquery(int b, int n, int m, int nsample, const float *xyz, const float *new_xyz, int *idx, float *dist2, pab_stream_t s);

/* The same query against a prebuilt spatial index of xyz (a Morton-sorted copy in 64-point chunks with bounding boxes):
 * chunks are visited in order of their box distance and the scan stops at the first chunk that cannot hold a better
 * point.  Results are identical to pab_knnquery (which builds a temporary index itself when 256 <= n <= 8192 and there
 * are enough queries); build the index once when several queries share xyz.  nsample <= 64.
 * index: pab_knn_index_bytes(b, n) bytes of device memory, 16-byte aligned. */
size_t pab_knn_index_bytes(int b, int n);
int pab_knn_build_index(int b, int n, const float *xyz, void *index, pab_stream_t s);
int pab_knnquery_indexed(int b, int n, int m, int nsample, const void *index, const float *new_xyz, int *idx, float *dist2,
                         pab_stream_t s);

/* ballquery_cuda_launcher_fast  ballquery/ballquery_cuda_kernel.h, kernel .cu:47-80.  idx caller-zeroed. */
int pab_ballquery(int b, int n, int m, float radius, int nsample, const float *new_xyz, const float *xyz, int *idx, pab_stream_t s);

/* grouping_forward_cuda_launcher_fast / grouping_backward_cuda_launcher  grouping/grouping_cuda_kernel.cu:60-92, 28-46 */
int pab_grouping_forward(int b, int c, int n, int m, int nsample, const float *points, const int *idx, float *out, pab_stream_t s);
int pab_grouping_backward(int b, int c, int n, int m, int nsample, const float *grad_out, const int *idx, float *grad_points, pab_stream_t s);
/* grouping_int_forward_cuda_launcher_fast  grouping_int/grouping_int_cuda_kernel.cu:33-64 (int64 payload) */
int pab_grouping_int_forward(int b, int c, int n, int m, int nsample, const int64_t *points, const int *idx, int64_t *out, pab_stream_t s);

/* nearestneighbor_cuda_launcher_fast  interpolation/interpolation_cuda_kernel.cu:134-176, 200-214.
 * unknown (b,n,3), known (b,m,3) -> dist2 (b,n,3) SQUARED f32, idx (b,n,3) i32. */
int pab_nearestneighbor(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx, pab_stream_t s);
/* interpolation_forward_cuda_launcher_fast .cu:181-195, 216-228; interpolation_backward_cuda_launcher .cu:90-129 */
int pab_interpolation_forward(int b, int c, int m, int n, const float *points, const int *idx, const float *weight, float *out, pab_stream_t s);
int pab_interpolation_backward(int b, int c, int n, int m, const float *grad_out, const int *idx, const float *weight, float *grad_points, pab_stream_t s);

/* featuredistribute/featuredistribute_cuda_kernel.cu:4-30, 53-65, 89-101 */
int pab_featuredistribute(int b, int n, int m, const float *max_xyz, const float *xyz, int *distribute_idx, pab_stream_t s);
int pab_featuregather_forward(int b, int n, int m, int c, const float *max_feature, const int *distribute_idx, float *distribute_feature, pab_stream_t s);
int pab_featuregather_backward(int b, int n, int m, int c, const float *grad_distribute_feature, const int *distribute_idx, float *grad_max_feature, pab_stream_t s);

/* labelstat/labelstat_cuda_kernel.cu:131-151, 74-105, 6-49 */
int pab_labelstat_idx(int b, int n, int m, int nsample, int nclass, const int *label_stat, const int *idx, int *new_label_stat, pab_stream_t s);
int pab_labelstat_ballrange(int b, int n, int m, float radius, int nclass, const float *new_xyz, const float *xyz, const int *label_stat, int *new_label_stat, pab_stream_t s);
int pab_labelstat_and_ballquery(int b, int n, int m, float radius, int nsample, int nclass, const float *new_xyz, const float *xyz,
                                const int *label_stat, int *idx, int *new_label_stat, pab_stream_t s);

/* ---- libs/chamfer_dist --------------------------------------------------------------------------- */
/* chamfer_cuda_forward  chamfer_cuda.cpp:22-29, chamfer.cu:147-171: both directions in one call.
 * xyz1 (B,n,3), xyz2 (B,m,3) -> dist1 (B,n), dist2 (B,m) squared f32; idx1, idx2 i32 (first minimum wins). */
int pab_chamfer_forward(int B, int n, const float *xyz1, int m, const float *xyz2, float *dist1, float *dist2, int *idx1, int *idx2, pab_stream_t s);
/* chamfer_cuda_backward  chamfer.cu:203-229: grad_xyz1 (B,n,3), grad_xyz2 (B,m,3) are overwritten
 * (deterministic segmented sums, not atomics). */
int pab_chamfer_backward(int B, int n, const float *xyz1, int m, const float *xyz2, const int *idx1, const int *idx2,
                         const float *grad_dist1, const float *grad_dist2, float *grad_xyz1, float *grad_xyz2, pab_stream_t s);

/* ---- libs/KNN_CUDA ------------------------------------------------------------------------------- */
/* knn_device  knn_cuda/csrc/cuda/knn.cu:232-269 via knn.cpp:23-56.  ref (dim,nr), query (dim,nq) row-major f32
 * -> dist (k,nq) f32 = sqrt(squared L2), ind (k,nq) int64 1-BASED (knn_cuda/__init__.py:41-44 subtracts 1).
 * No nr x nq matrix is materialised.  k <= 1024. */
int pab_knn(const float *ref, int nr, const float *query, int nq, int dim, int k, float *dist, int64_t *ind, pab_stream_t s);

/* Retrieval kNN over row-major descriptors (replaces sklearn KDTree.query at
 * datasets/place_recognition_dataset.py:60, scene_dataset.py:1052): db (ndb,dim), q (nq,dim) ->
 * dist (nq,k) f32 Euclidean ascending, ind (nq,k) i32 0-based, ties to the lower index. k <= 1024. */
int pab_retrieval_topk(const float *db, int ndb, const float *q, int nq, int dim, int k, float *dist, int *ind, pab_stream_t s);
/* The same search with the database ALSO split over the grid (partial top-k per slice + a merge kernel): identical results, and a
 * small query set (a rank's shard) still fills the machine.  k <= 128; workspace >= pab_retrieval_topk_workspace_bytes(nq, k). */
size_t pab_retrieval_topk_workspace_bytes(int nq, int k);
int pab_retrieval_topk_split(const float *db, int ndb, const float *q, int nq, int dim, int k, float *dist, int *ind,
                             void *workspace, pab_stream_t s);
/* The same search with a per-query candidate set: mask has nq rows of ceil(ndb/32) 32-bit words, bit p of row q set = query q
 * may retrieve database entry p.  One launch for a whole batch of ragged sets (hard-negative mining,
 * datasets/scene_dataset.py:1101-1113).  Unused slots of a query with fewer than k candidates hold +inf / index 0. */
int pab_retrieval_topk_masked(const float *db, int ndb, const float *q, int nq, int dim, int k, const unsigned *mask,
                              float *dist, int *ind, pab_stream_t s);

/* ---- libs/emd_module ----------------------------------------------------------------------------- */
/* emd_cuda_forward  emd.cpp:14-22, emd_cuda.cu:228-282.  Same buffers, same return convention
 * (1 ok, 0 CUDA error, -1 bad shape: n % 1024 != 0 or b > 512). */
int pab_emd_forward(int b, int n, const float *xyz1, const float *xyz2, float *dist, int *assignment, float *price,
                    int *assignment_inv, int *bid, float *bid_increments, float *max_increments, int *unass_idx,
                    int *unass_cnt, int *unass_cnt_sum, int *cnt_tmp, int *max_idx, float eps, int iters, pab_stream_t s);
/* emd_cuda_backward  emd_cuda.cu:284-316 */
int pab_emd_backward(int b, int n, const float *xyz1, const float *xyz2, float *gradxyz, const float *graddist, const int *idx, pab_stream_t s);

/* ---- fused descriptor-extraction path (no reference counterpart: replaces chains of reference ops) -- */

/* Point-major row gather out[b,j,:] = feat[b,idx[b,j],:] (feat (b,n,c), idx (b,m), out (b,m,c)); used for
 * new_xyz = xyz[center_idx] (patch_aug_net.py:222-225 does it with a transpose + gathering + transpose). */
int pab_gather_rows(int b, int n, int m, int c, const float *feat, const int *idx, float *out, pab_stream_t s);

/* Fused 3-NN + inverse-distance weights: pointops.nearestneighbor + patch_aug_net.py:350-353.
 * unknown (b,n,3), known (b,m,3) -> idx (b,n,3) i32, weight (b,n,3) f32. */
int pab_three_nn_weights(int b, int n, int m, const float *unknown, const float *known, int *idx, float *weight, pab_stream_t s);
/* The same against prebuilt spatial indices (pab_knn_build_index): known_index over the known points (required,
 * 256 <= m <= 8192); unknown_index over the unknown points (optional: queries are then processed in Morton order, which
 * keeps every warp spatially compact; results are written at the original positions).  Identical results. */
int pab_three_nn_weights_indexed(int b, int n, int m, const float *unknown, const void *unknown_index, const void *known_index,
                                 int *idx, float *weight, pab_stream_t s);

/* One folded SharedMLP layer: y = relu?(Wt^T x + shift).  Wt is (c_in_pad, c_out) row-major with the
 * BatchNorm scale folded in and rows >= c_in zero; c_in_pad = c_in rounded up to a multiple of 4. */
typedef struct {
    const float *wt;    /* (c_in_pad, c_out) */
    const float *shift; /* (c_out) */
    int c_in, c_in_pad, c_out, relu;
    /* Optional tensor-core operands (NULL -> the layer runs on the fp32 SIMT path).  The same folded weight, restricted
     * to input channels [tc_k0, tc_k0 + tc_k), split as W = hi + lo with hi = bf16(W), lo = bf16(W - hi), each stored
     * K-major bf16.  tc_k must be a multiple of 64 (zero-padded).  A layer whose c_in exceeds the channels covered that
     * way — layer 0 of a module, whose (at most 3) remaining inputs are the xyz part — stores (c_out, tc_k + 64): the
     * extra 64-column chunk holds those channels' weights in channel order (rest zero) and is fed to the tensor cores
     * as one more k-step.  Every other layer stores (c_out, tc_k). */
    const void *w_hi, *w_lo;
    int tc_k0, tc_k;
} pab_layer_t;

/* Fused set-abstraction module (patch_aug_net.py:203-243 + pointops.py:533-582, eval mode):
 * gather centres, gather k neighbours, subtract centre, concat [xyz_rel ; feat_rel], SharedMLP, max over k.
 * xyz (b,n,3); feat (b,n,c) POINT-MAJOR (may alias xyz with c=3); center_idx (b,m); nbr_idx (b,m,nbr_stride),
 * the first k entries of each row are used; out (b,m,c_out_last) point-major; new_xyz (b,m,3) out (or NULL). */
int pab_sa_module_forward(int b, int n, int m, int k, int nbr_stride, int c, const float *xyz, const float *feat,
                          const int *center_idx, const int *nbr_idx, const pab_layer_t *layers, int n_layers,
                          float *out, float *new_xyz, pab_stream_t s);

/* Fused feature-propagation module (patch_aug_net.py:331-363): 3-NN interpolation of known_feat
 * (b,m,c_known) with idx/weight (b,n,3), concat skip (b,n,c_skip) (may be NULL with c_skip 0), SharedMLP.
 * out (b,n,c_out_last) point-major. */
int pab_fp_module_forward(int b, int n, int m, int c_known, int c_skip, const float *known_feat, const float *skip_feat,
                          const int *idx, const float *weight, const pab_layer_t *layers, int n_layers,
                          float *out, pab_stream_t s);
/* Same module with a processing order of each cloud's points: row_order[cloud * order_stride + i] = the point handled as the
 * cloud's i-th row (a permutation of 0..n-1; NULL = index order).  Scheduling only — every point is computed once and stored at
 * its own position, results are bit-identical; a spatial order (pab_knn_index_order) makes consecutive rows share their 3-NN
 * rows of the known cloud in L1. */
int pab_fp_module_forward_ordered(int b, int n, int m, int c_known, int c_skip, const float *known_feat, const float *skip_feat,
                                  const int *idx, const float *weight, const int *row_order, long order_stride,
                                  const pab_layer_t *layers, int n_layers, float *out, pab_stream_t s);
/* The Morton-order permutation stored inside an index built by pab_knn_build_index (device pointer into `index`;
 * *stride_ints = ints between consecutive clouds).  NULL when n is not a multiple of 64 (the index then has padding rows). */
const int *pab_knn_index_order(int n, const void *index, long *stride_ints);


/* Deterministic backward of the index-driven ops (replaces the fp32 atomicAdd scatters of sampling_cuda_kernel.cu:23-36,
 * grouping_cuda_kernel.cu:28-46, interpolation_cuda_kernel.cu:90-114): grad_points[b,ch,idx[b,e]] += grad_out[b,ch,e] *
 * (weight ? weight[b,e] : 1) for e in [0,L), summed per target in ascending e — bit-identical from run to run.
 * grad_out (b,c,L), idx (b,L), weight (b,L) or NULL, grad_points (b,c,n); workspace >= pab_scatter_workspace_bytes(b,n,L).
 * L <= 32768 per cloud (PAB_EINVAL beyond: use the atomic entry points). */
size_t pab_scatter_workspace_bytes(int b, int n, int L);
int pab_scatter_add_deterministic(int b, int c, int n, int L, int gdiv, const float *grad_out, const int *idx, const float *weight,
                                  float *grad_points, void *workspace, pab_stream_t s);
/* build_index = 0: workspace already holds the inverse of this idx from an earlier call with the same b, n, L, idx. */
int pab_scatter_add_deterministic_ex(int b, int c, int n, int L, int gdiv, const float *grad_out, const int *idx, const float *weight,
                                     float *grad_points, void *workspace, int build_index, pab_stream_t s);

/* Train-mode BatchNorm + ReLU of a SharedMLP block (utils/model_util/pt_util.py:98-151: conv -> BatchNorm (batch statistics)
 * -> ReLU), fused: x, y, dy, dx are (B, C, S) contiguous, S = product of the trailing dimensions.  forward: batch mean / biased
 * variance per channel, y = relu((x - mean) * invstd * gamma + beta), saved mean / invstd for the backward, running statistics
 * updated like nn.BatchNorm (momentum, unbiased variance) when the pointers are not NULL.  backward: dz = dy * (y > 0),
 * dbeta = sum dz, dgamma = sum dz * xhat, dx = gamma * invstd * (dz - dbeta/N - xhat * dgamma/N).  Deterministic reductions.
 * workspace >= pab_bn_train_workspace_bytes(C). */
size_t pab_bn_train_workspace_bytes(int C);
int pab_bn_relu_train_forward(int B, int C, long S, const float *x, const float *gamma, const float *beta, float eps, float momentum,
                              float *running_mean, float *running_var, float *mean, float *invstd, float *y, void *workspace,
                              pab_stream_t s);
int pab_bn_relu_train_backward(int B, int C, long S, const float *dy, const float *y, const float *x, const float *gamma,
                               const float *mean, const float *invstd, float *dx, float *dgamma, float *dbeta, void *workspace,
                               pab_stream_t s);

/* Plain point-wise SharedMLP over rows: x (rows, c_in) -> out (rows, c_out_last). */
int pab_pointwise_mlp_forward(int rows, const float *x, const pab_layer_t *layers, int n_layers, float *out, pab_stream_t s);

/* PPT-Net SA_Layer.forward (place_recognition/pptnet_origin/models/pptnet.py:261-282), eval: grouped tied q/k projection,
 * Gram-matrix energy, row softmax, column renormalisation, V @ attn, trans_conv + BN + ReLU, residual — without ever
 * materialising an N x N tensor in HBM.  x, out (b,n,c) point-major, c % 64 == 0.  q_layer: c -> c, the grouped tied
 * q/k weight expanded to a dense block-diagonal matrix, zero shift, no ReLU; v_layer: c -> c, v_conv (shift = bias), no
 * ReLU; trans_layer: c -> c, trans_conv with after_norm folded, ReLU.  workspace >= pab_sa_layer_workspace_bytes(b,n,c). */
size_t pab_sa_layer_workspace_bytes(int b, int n, int c);
int pab_sa_layer_forward(int b, int n, int c, const float *x, const pab_layer_t *q_layer, const pab_layer_t *v_layer,
                         const pab_layer_t *trans_layer, float *out, void *workspace, pab_stream_t s);
/* Same layer with the arithmetic of the two N x N passes chosen by the caller: precision 2 (what pab_sa_layer_forward
 * uses) = tcgen05 tensor cores on bf16 hi/lo operand planes, three MMAs per product — the fp32 contract; 1 = tensor cores on
 * plain bf16 operands (BASELINE.json configs[2]); 0 = the fp32 SIMT kernels.  Shapes the tensor-core kernel does not take
 * (c not in {64, 128}, n < 64) run on the SIMT kernels whatever the precision. */
/* Tuning hook: 1 (default) = levels with n <= 64 points run the whole layer (projections, attention, trans_conv, residual) as ONE
 * kernel, one CTA per cloud, everything in shared memory; 0 = the five-launch path for every size. */
void pab_tune_attention_small(int on);
/* Tuning hook: 1 (default) = single point-wise layers whose pab_layer_t carries bf16 hi/lo planes (c_in % 64 == 0, c_out % 32 == 0) run
 * on the tcgen05 kernel of pw_tc.cu; 0 = always the fp32 tile kernel.  pab_tune_attention_small(2) forces the single-kernel
 * SA_Layer for small levels even when its projections could use the tensor cores. */
void pab_tune_pointwise_tc(int on);
/* Tuning hook: enable = 1 (default): a set-abstraction module with a tiny input (<= 8 channels) and layers <= 64 wide (SA0 of
 * both networks) runs on sa_narrow_tc.cu — 128-thread CTAs, ctas_per_sm resident per SM (1..3; 0 = default: 3, or 2 while pab_tune_tc_max_ctas reserves SMs for another stream), weights resident in
 * shared memory; 0: the warp-specialised one-CTA-per-SM kernel of mlp_tc.cu takes it.  Results are bit-identical. */
void pab_tune_sa_narrow(int enable, int ctas_per_sm);
/* Debugging aid: device buffer of 16 x 8 int64 receiving clock64 stamps of CTA 0's first 16 steps (NULL: off). */
void pab_tune_sa_narrow_trace(void *device_buffer);
/* Debugging aid (results become wrong): 1 = no gathers, 2 = pre-layer without its multiply-adds. */
void pab_tune_sa_narrow_dbg(int flags);
int pab_sa_layer_forward_p(int b, int n, int c, const float *x, const pab_layer_t *q_layer, const pab_layer_t *v_layer,
                           const pab_layer_t *trans_layer, float *out, void *workspace, int precision, pab_stream_t s);

/* NetVLADBase.forward (patch_aug_net/models/loupe.py:191-222), eval: x (b,n,c) point-major, wc (c,K) with the
 * bn1 scale folded in, shift (K), w2 (c,K) = cluster_weights2; out written at out[b*out_bstride + ch*out_cstride + k]
 * (lets the caller place levels side by side in the (B,C,sumK) concat of loupe.py:302).  c in {128,256}, K % 4 == 0,
 * K <= 64.  workspace: >= pab_netvlad_workspace_bytes(b,n,c,K) bytes. */
size_t pab_netvlad_workspace_bytes(int b, int n, int c, int K);
int pab_netvlad_forward(int b, int n, int c, int K, const float *x, const float *wc, const float *shift, const float *w2,
                        float *out, long out_bstride, long out_cstride, void *workspace, pab_stream_t s);

/* Tensor-core variant of pab_netvlad_forward (c == 256): wc_hi / wc_lo are the bf16 hi/lo planes of the folded cluster
 * weights TRANSPOSED to (Kp, c) K-major, Kp = K rounded up to 16 with zero rows.  Same outputs and workspace. */
int pab_netvlad_forward_tc(int b, int n, int c, int K, const float *x, const void *wc_hi, const void *wc_lo, const float *shift,
                           const float *w2, float *out, long out_bstride, long out_cstride, void *workspace, pab_stream_t s);

/* AdaptiveFeatureAggregator.forward (loupe.py:57-66, 24-41), eval: v (b,c,K) -> desc (b,c_out), L2-normalised.
 * w_att_t (c_in,c_out) = mlpa.mlps.0.weight[:, :, 0] TRANSPOSED; fc_wt (c*K, c_out) = fc.weight TRANSPOSED (so the
 * 22 MB matrix streams coalesced); desc = normalize((fc_wt^T y) * fc_scale + fc_shift) with fc.bias and the
 * BatchNorm1d folded into fc_scale/fc_shift.  workspace >= pab_afa_workspace_bytes(b,c,K,c_out). */
size_t pab_afa_workspace_bytes(int b, int c, int K, int c_out);
int pab_afa_forward(int b, int c, int K, int c_out, const float *v, const float *w_att_t, const float *fc_wt,
                    const float *fc_scale, const float *fc_shift, int l2_norm, float *desc, void *workspace, pab_stream_t s);
/* The same head with its two products (attention logits, fc) on the tcgen05 tensor cores: bf16 hi/lo operand planes, three MMAs
 * per product, fp32 accumulation — the fp32 contract of the other tensor-core kernels.  watt_hi / watt_lo: the attention conv
 * weight (c_out' = c rows, c inputs contiguous) split as W = hi + lo in bf16; wfc_hi / wfc_lo: the fc weight (c_out rows, c*K
 * inputs contiguous, f = channel * K + cluster) split the same way.  pab_afa_tc_supported: 1 for the shapes the kernels take
 * (c in {64,128,192,256}, c*K % 64 == 0, c_out % 32 == 0, c_out <= 256) — otherwise use pab_afa_forward.
 * workspace >= pab_afa_tc_workspace_bytes. */
int pab_afa_tc_supported(int c, int K, int c_out);
size_t pab_afa_tc_workspace_bytes(int b, int c, int K, int c_out);
int pab_afa_forward_tc(int b, int c, int K, int c_out, const float *v, const void *watt_hi, const void *watt_lo,
                       const void *wfc_hi, const void *wfc_lo, const float *fc_scale, const float *fc_shift, int l2_norm,
                       float *desc, void *workspace, pab_stream_t s);
/* Tuning hook: 0 makes pab_afa_tc_supported return 0 (the engine then calls pab_afa_forward). */
void pab_tune_afa_tc(int on);

/* Dense head of the PPT-Net / PointNetVLAD style (pptnet_origin/models/loupe.py:99-136): desc = [normalize]( x * sigmoid(
 * (x G) * gate_scale + gate_shift) ),  x = (fc_wt^T v) * fc_scale + fc_shift.  v (b, f) row-major (the flattened VLAD);
 * fc_wt (f, c_out) = hidden_weights; fc_scale/fc_shift = the folded BatchNorm1d after it; gate_wt (c_out, c_out) =
 * gating_weights with its BatchNorm1d (or gating_biases) folded into gate_scale/gate_shift, or NULL for no gating.
 * c_out <= 256.  The f x c_out weight is streamed once (split-K over f), partials combined in a fixed order.
 * workspace >= pab_gated_fc_workspace_bytes(b, f, c_out). */
size_t pab_gated_fc_workspace_bytes(int b, int f, int c_out);
int pab_gated_fc_forward(int b, int f, int c_out, const float *v, const float *fc_wt, const float *fc_scale, const float *fc_shift,
                         const float *gate_wt, const float *gate_scale, const float *gate_shift, int l2_norm, float *desc,
                         void *workspace, pab_stream_t s);
/* The same head with the fc product on the tcgen05 tensor cores (bf16 hi/lo planes of the fc weight, (c_out, f) with the inputs
 * contiguous).  pab_gated_fc_tc_supported: f % 64 == 0, c_out % 32 == 0, c_out <= 256.  Same workspace as pab_gated_fc_forward. */
int pab_gated_fc_tc_supported(int f, int c_out);
int pab_gated_fc_forward_tc(int b, int f, int c_out, const float *v, const void *wfc_hi, const void *wfc_lo, const float *fc_scale,
                            const float *fc_shift, const float *gate_wt, const float *gate_scale, const float *gate_shift,
                            int l2_norm, float *desc, void *workspace, pab_stream_t s);

/* Device side of the input pipeline (SceneDataSet.get_pc + normalize_point_cloud, datasets/scene_dataset.py:713-740,
 * utils/loading_pointclouds.py:51-63): raw (b,n,3) clouds as stored in the .bin files (float64 when raw_is_f64, else float32;
 * device memory) minus the host array offset[3] (global_offset, may be NULL), optionally centred on the mean (normalize) and
 * divided by the largest point norm (zoom), written as the float32 (b,n,3) batch the network takes.  meta (b,4) double, may be
 * NULL: {scale, trans.x, trans.y, trans.z} = the reference's norm_meta.  float64 arithmetic, n <= 8192. */
int pab_prepare_clouds(int b, int n, const void *raw, int raw_is_f64, const double *offset, int normalize, int zoom, float *out,
                       double *meta, pab_stream_t s);

/* Patch-feature-contrast (a2b) triplet selection of one training step — replaces the per-pair numpy where/isin loop and the
 * per-triplet index_select + H2D copies of place_recognition/train_place_recognition.py:320-378.
 * centers (n_clouds, M) int32 level-0 centre indices; pair p = clouds (pair_m[p], pair_n[p]) (rows of `centers`) with overlap
 * entries [entry_ptr[p], entry_ptr[p+1]) in processing order (<= max_entries_per_pair <= 1024 each; the reference samples 500);
 * entry e = (entry_idx1[e], near list near_val[near_ptr[e]..near_ptr[e+1]), far list far_val[far_ptr[e]..)).  Output, per pair
 * p at [p*max_out ..): triplets (position of idx1 among m's centres, position of a positive among n's centres, position of a
 * negative drawn with replacement by a counter-based generator keyed by seed) in the reference's order; out_count[p] = number
 * of triplets of the pair — when it exceeds max_out only the first max_out were written (call again with more room). */
int pab_patch_triplets(int n_pairs, int M, const int *centers, const int *pair_m, const int *pair_n, const int *entry_ptr,
                       int max_entries_per_pair, const int *entry_idx1, const int *near_ptr, const int *near_val,
                       const int *far_ptr, const int *far_val, unsigned long long seed, int max_out, int *out_idx1, int *out_pos,
                       int *out_neg, int *out_count, pab_stream_t s);

/* Tuning hook: bit 0 enables (1, default) / disables (0) the tcgen05 tensor-core path of the fused SharedMLP kernels;
 * bit 2 set (5) additionally shares the weight stream across CTA pairs (thread-block clusters of 2, TMA multicast;
 * off by default: measured slower on B200); bit 3 set (9) turns on dynamic tile scheduling of the persistent CTAs (tiles
 * drawn from a global counter, so CTAs that start late because another stream's kernel holds their SM take fewer). */
void pab_tune_tensor_core(int enable);

/* Tuning hook: cap the number of persistent CTAs of the tensor-core kernels (0 = one per SM, default) so that kernels of
 * other streams keep some SMs. */
void pab_tune_tc_max_ctas(int n);

/* Debugging aid: when set to a device buffer of 8*4*8 int64, CTA 0 of every fused-MLP tensor-core launch records clock64
 * stamps of its first 8 tiles ([tile][layer][event]: 0 MMA issue start, 1 MMA issue end, 2 accumulators seen by the
 * epilogue, 3 epilogue done, 4 loader start, 5 operand staged, 6 tile output stored).  NULL (default) disables it. */
void pab_tune_tc_trace(void *device_buffer);

/* Tuning hook: force the FPS CTA size (power of two, 32..1024; 0 = automatic). */
void pab_tune_fps_threads(int threads);

/* Tuning hook: 1 = clouds of 2048/4096/8192 points use the pruned sampler (Morton chunks + box bounds, exact,
 * bit-identical indices); 0 (default) = always the full-scan register-resident sampler, which is faster at these sizes
 * because a step is bound by its arg-max latency chain, not by the distance updates. */
void pab_tune_fps_pruned(int on);

/* Tuning hook: 1 = a full-scan sampler CTA claims more than half an SM's shared memory, so no two of them can share an SM;
 * 0 (default) = only the memory it needs (measured identical: 0.40 ms for 1..128 clouds of 4096 points). */
void pab_tune_fps_exclusive(int on);

/* How many clouds of n points one SM samples concurrently with the current settings (the engine sizes the persistent
 * dense kernels of the overlapping batch by it). */
int pab_fps_clouds_per_sm(int n);

/* Tuning hook: 2 = one FPS CTA samples two clouds side by side (independent halves of the CTA; identical results), so the
 * sampler occupies half as many SMs; 1 (default) = one cloud per CTA. */
void pab_tune_fps_clouds_per_cta(int n);

#ifdef __cplusplus
}
#endif
#endif /* PATCHAUG_B200_H */
