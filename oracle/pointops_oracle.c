/*
 * oracle/pointops_oracle.c — TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, single-threaded restatement of the reference's CUDA algorithms for
 * the descriptor-extraction hot path.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library;
 * the product path (patchaugnet_b200/) never does.
 *
 * Every function cites the reference file:line it restates (paths relative to
 * /root/reference).  Floating-point arithmetic is written with explicit
 * fmaf() in the order nvcc 12.9 -O2 contracts the reference source for sm_100
 * (SURVEY.md section 0):   d = fmaf(dz,dz, fmaf(dx,dx, dy*dy)).
 * Build with -ffp-contract=off so the compiler adds no contraction of its own.
 *
 * Parity status: the reference ships no golden vectors for these ops
 * (SURVEY.md section 4).  The oracle is pinned instead against the reference's
 * own kernels compiled from /root/reference for sm_100 (oracle/_ref, see
 * oracle/Makefile) on the GPU box; fixtures produced there are committed
 * under tests/golden/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORA_API __attribute__((visibility("default")))

static inline float sqdist3(float ax, float ay, float az, float bx, float by, float bz) {
    /* nvcc contraction of (ax-bx)*(ax-bx)+(ay-by)*(ay-by)+(az-bz)*(az-bz) */
    float dx = ax - bx, dy = ay - by, dz = az - bz;
    float t = dy * dy;
    t = fmaf(dx, dx, t);
    return fmaf(dz, dz, t);
}

/* libs/pointops/src/cuda_utils.h:15-18 */
ORA_API int ora_opt_n_threads(int work_size) {
    const int pow_2 = (int)(log((double)work_size) / log(2.0));
    int v = 1 << pow_2;
    if (v > 1024) v = 1024;
    if (v < 1) v = 1;
    return v;
}

/* libs/pointops/src/sampling/sampling_cuda_kernel.cu:58-168 (kernel), :48-54 (__update),
 * :170-210 (block size = opt_n_threads(n)).  temp is caller-initialised (1e10 in pointops.py:21)
 * and is updated in place like the reference. */
ORA_API void ora_furthestsampling(int b, int n, int m, const float *xyz, float *temp, int *idx) {
    if (m <= 0) return;
    const int bs = ora_opt_n_threads(n);
    /* clouds are independent (one CUDA block per cloud in the reference): one host thread per cloud */
#pragma omp parallel for schedule(dynamic, 1)
    for (int bi = 0; bi < b; ++bi) {
        float *dists = (float *)malloc(sizeof(float) * bs);
        int *dists_i = (int *)malloc(sizeof(int) * bs);
        const float *p = xyz + (size_t)bi * n * 3;
        float *t = temp + (size_t)bi * n;
        int *out = idx + (size_t)bi * m;
        int old = 0;
        out[0] = 0;
        for (int j = 1; j < m; ++j) {
            const float x1 = p[old * 3 + 0], y1 = p[old * 3 + 1], z1 = p[old * 3 + 2];
            for (int tid = 0; tid < bs; ++tid) {
                int besti = 0;
                float best = -1.f;
                for (int k = tid; k < n; k += bs) {
                    float d = sqdist3(p[k * 3 + 0], p[k * 3 + 1], p[k * 3 + 2], x1, y1, z1);
                    float d2 = fminf(d, t[k]);
                    t[k] = d2;
                    if (d2 > best) { besti = k; best = d2; }
                }
                dists[tid] = best;
                dists_i[tid] = besti;
            }
            for (int s = bs / 2; s >= 1; s >>= 1) {
                for (int tid = 0; tid < s; ++tid) {
                    float v1 = dists[tid], v2 = dists[tid + s];
                    int i1 = dists_i[tid], i2 = dists_i[tid + s];
                    dists[tid] = v1 > v2 ? v1 : v2;
                    dists_i[tid] = v2 > v1 ? i2 : i1;
                }
            }
            old = dists_i[0];
            out[j] = old;
        }
        free(dists);
        free(dists_i);
    }
}

/* sampling_cuda_kernel.cu:6-19 */
ORA_API void ora_gathering_forward(int b, int c, int n, int m, const float *points, const int *idx, float *out) {
    for (int i = 0; i < b; ++i)
        for (int l = 0; l < c; ++l)
            for (int j = 0; j < m; ++j)
                out[((size_t)i * c + l) * m + j] = points[((size_t)i * c + l) * n + idx[(size_t)i * m + j]];
}

/* sampling_cuda_kernel.cu:23-36 (atomicAdd order is unspecified; sequential here) */
ORA_API void ora_gathering_backward(int b, int c, int n, int m, const float *grad_out, const int *idx, float *grad_points) {
    for (int i = 0; i < b; ++i)
        for (int l = 0; l < c; ++l)
            for (int j = 0; j < m; ++j)
                grad_points[((size_t)i * c + l) * n + idx[(size_t)i * m + j]] += grad_out[((size_t)i * c + l) * m + j];
}

/* libs/pointops/src/knnquery/knnquery_cuda_kernel.cu:6-50.  The reference writes dist2 without a
 * batch/point offset (:44-47, a race; Python discards it, pointops.py:425-427); the oracle writes
 * dist2[b,m,k] properly so it can be checked. */
ORA_API int ora_knnquery(int b, int n, int m, int nsample, const float *xyz, const float *new_xyz, int *idx, float *dist2) {
    if (nsample > 200 || nsample < 0) return -1; /* fixed arrays best[200] at :21-22 */
    /* queries are independent (one CUDA thread per query in the reference): host threads over (cloud, query) */
#pragma omp parallel for collapse(2) schedule(static)
    for (int bi = 0; bi < b; ++bi) {
        for (int q = 0; q < m; ++q) {
            double best[200];
            int besti[200];
            const float *p = xyz + (size_t)bi * n * 3;
            const float *c = new_xyz + ((size_t)bi * m + q) * 3;
            for (int i = 0; i < nsample; ++i) { best[i] = 1e40; besti[i] = 0; }
            for (int k = 0; k < n; ++k) {
                float d2 = sqdist3(c[0], c[1], c[2], p[k * 3 + 0], p[k * 3 + 1], p[k * 3 + 2]);
                for (int j = 0; j < nsample; ++j) {
                    if ((double)d2 < best[j]) {
                        for (int i = nsample - 1; i > j; --i) { best[i] = best[i - 1]; besti[i] = besti[i - 1]; }
                        best[j] = d2;
                        besti[j] = k;
                        break;
                    }
                }
            }
            int *o = idx + ((size_t)bi * m + q) * nsample;
            for (int i = 0; i < nsample; ++i) o[i] = besti[i];
            if (dist2) {
                float *od = dist2 + ((size_t)bi * m + q) * nsample;
                for (int i = 0; i < nsample; ++i) od[i] = (float)best[i];
            }
        }
    }
    return 0;
}

/* libs/pointops/src/ballquery/ballquery_cuda_kernel.cu:47-80 (the _fast kernel bound by
 * pointops_api.cpp:16).  idx is caller-zeroed (pointops.py:189). */
ORA_API void ora_ballquery(int b, int n, int m, float radius, int nsample, const float *new_xyz, const float *xyz, int *idx) {
    const float radius2 = radius * radius;
    for (int bi = 0; bi < b; ++bi) {
        const float *p = xyz + (size_t)bi * n * 3;
        for (int q = 0; q < m; ++q) {
            const float *c = new_xyz + ((size_t)bi * m + q) * 3;
            int *o = idx + ((size_t)bi * m + q) * nsample;
            int cnt = 0;
            for (int k = 0; k < n; ++k) {
                float d2 = sqdist3(c[0], c[1], c[2], p[k * 3 + 0], p[k * 3 + 1], p[k * 3 + 2]);
                if (d2 < radius2) {
                    if (cnt == 0) for (int l = 0; l < nsample; ++l) o[l] = k;
                    o[cnt] = k;
                    ++cnt;
                    if (cnt >= nsample) break;
                }
            }
        }
    }
}

/* libs/pointops/src/grouping/grouping_cuda_kernel.cu:60-74 */
ORA_API void ora_grouping_forward(int b, int c, int n, int m, int nsample, const float *points, const int *idx, float *out) {
    for (int bi = 0; bi < b; ++bi)
        for (int l = 0; l < c; ++l)
            for (int j = 0; j < m; ++j)
                for (int s = 0; s < nsample; ++s)
                    out[(((size_t)bi * c + l) * m + j) * nsample + s] =
                        points[((size_t)bi * c + l) * n + idx[((size_t)bi * m + j) * nsample + s]];
}

/* grouping_cuda_kernel.cu:28-46 */
ORA_API void ora_grouping_backward(int b, int c, int n, int m, int nsample, const float *grad_out, const int *idx, float *grad_points) {
    for (int bi = 0; bi < b; ++bi)
        for (int l = 0; l < c; ++l)
            for (int j = 0; j < m; ++j)
                for (int s = 0; s < nsample; ++s)
                    grad_points[((size_t)bi * c + l) * n + idx[((size_t)bi * m + j) * nsample + s]] +=
                        grad_out[(((size_t)bi * c + l) * m + j) * nsample + s];
}

/* libs/pointops/src/grouping_int/grouping_int_cuda_kernel.cu:33-47 */
ORA_API void ora_grouping_int_forward(int b, int c, int n, int m, int nsample, const int64_t *points, const int *idx, int64_t *out) {
    for (int bi = 0; bi < b; ++bi)
        for (int l = 0; l < c; ++l)
            for (int j = 0; j < m; ++j)
                for (int s = 0; s < nsample; ++s)
                    out[(((size_t)bi * c + l) * m + j) * nsample + s] =
                        points[((size_t)bi * c + l) * n + idx[((size_t)bi * m + j) * nsample + s]];
}

/* libs/pointops/src/interpolation/interpolation_cuda_kernel.cu:134-176 (3-NN, _fast) */
ORA_API void ora_nearestneighbor(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int bi = 0; bi < b; ++bi) {
        for (int j = 0; j < n; ++j) {
            const float *kn = known + (size_t)bi * m * 3;
            const float *u = unknown + ((size_t)bi * n + j) * 3;
            double best1 = 1e40, best2 = 1e40, best3 = 1e40;
            int besti1 = 0, besti2 = 0, besti3 = 0;
            for (int k = 0; k < m; ++k) {
                float d = sqdist3(u[0], u[1], u[2], kn[k * 3 + 0], kn[k * 3 + 1], kn[k * 3 + 2]);
                if (d < best1) {
                    best3 = best2; besti3 = besti2;
                    best2 = best1; besti2 = besti1;
                    best1 = d; besti1 = k;
                } else if (d < best2) {
                    best3 = best2; besti3 = besti2;
                    best2 = d; besti2 = k;
                } else if (d < best3) {
                    best3 = d; besti3 = k;
                }
            }
            float *od = dist2 + ((size_t)bi * n + j) * 3;
            int *oi = idx + ((size_t)bi * n + j) * 3;
            od[0] = (float)best1; od[1] = (float)best2; od[2] = (float)best3;
            oi[0] = besti1; oi[1] = besti2; oi[2] = besti3;
        }
    }
}

/* interpolation_cuda_kernel.cu:181-195; nvcc contracts w0*p0 + w1*p1 + w2*p2 as
 * fmaf(w2,p2, fmaf(w0,p0, w1*p1)). */
ORA_API void ora_interpolation_forward(int b, int c, int m, int n, const float *points, const int *idx, const float *weight, float *out) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int bi = 0; bi < b; ++bi)
        for (int l = 0; l < c; ++l) {
            const float *p = points + ((size_t)bi * c + l) * m;
            for (int j = 0; j < n; ++j) {
                const float *w = weight + ((size_t)bi * n + j) * 3;
                const int *ii = idx + ((size_t)bi * n + j) * 3;
                float t = w[1] * p[ii[1]];
                t = fmaf(w[0], p[ii[0]], t);
                out[((size_t)bi * c + l) * n + j] = fmaf(w[2], p[ii[2]], t);
            }
        }
}

/* interpolation_cuda_kernel.cu:90-114 */
ORA_API void ora_interpolation_backward(int b, int c, int n, int m, const float *grad_out, const int *idx, const float *weight, float *grad_points) {
    for (int bi = 0; bi < b; ++bi)
        for (int l = 0; l < c; ++l) {
            float *gp = grad_points + ((size_t)bi * c + l) * m;
            for (int j = 0; j < n; ++j) {
                const float *w = weight + ((size_t)bi * n + j) * 3;
                const int *ii = idx + ((size_t)bi * n + j) * 3;
                float g = grad_out[((size_t)bi * c + l) * n + j];
                gp[ii[0]] += g * w[0];
                gp[ii[1]] += g * w[1];
                gp[ii[2]] += g * w[2];
            }
        }
}

/* libs/pointops/src/featuredistribute/featuredistribute_cuda_kernel.cu:4-30 */
ORA_API void ora_featuredistribute(int b, int n, int m, const float *max_xyz, const float *xyz, int *distribute_idx) {
    for (int bi = 0; bi < b; ++bi)
        for (int q = 0; q < m; ++q) {
            const float *c = xyz + ((size_t)bi * m + q) * 3;
            const float *p = max_xyz + (size_t)bi * n * 3;
            float min_dist2 = 100000.f;
            int min_idx = -1;
            for (int k = 0; k < n; ++k) {
                float d2 = sqdist3(p[k * 3 + 0], p[k * 3 + 1], p[k * 3 + 2], c[0], c[1], c[2]);
                if (d2 < min_dist2) { min_idx = k; min_dist2 = d2; }
            }
            distribute_idx[(size_t)bi * m + q] = min_idx;
        }
}

/* featuredistribute_cuda_kernel.cu:53-65 */
ORA_API void ora_featuregather_forward(int b, int n, int m, int c, const float *max_feature, const int *distribute_idx, float *distribute_feature) {
    for (int bi = 0; bi < b; ++bi)
        for (int l = 0; l < c; ++l)
            for (int q = 0; q < m; ++q)
                distribute_feature[((size_t)bi * c + l) * m + q] = max_feature[((size_t)bi * c + l) * n + distribute_idx[(size_t)bi * m + q]];
}

/* featuredistribute_cuda_kernel.cu:89-101 */
ORA_API void ora_featuregather_backward(int b, int n, int m, int c, const float *grad_distribute_feature, const int *distribute_idx, float *grad_max_feature) {
    for (int bi = 0; bi < b; ++bi)
        for (int l = 0; l < c; ++l)
            for (int q = 0; q < m; ++q)
                grad_max_feature[((size_t)bi * c + l) * n + distribute_idx[(size_t)bi * m + q]] += grad_distribute_feature[((size_t)bi * c + l) * m + q];
}

/* libs/pointops/src/labelstat/labelstat_cuda_kernel.cu:131-151 */
ORA_API void ora_labelstat_idx(int b, int n, int m, int nsample, int nclass, const int *label_stat, const int *idx, int *new_label_stat) {
    for (int bi = 0; bi < b; ++bi)
        for (int q = 0; q < m; ++q) {
            int *o = new_label_stat + ((size_t)bi * m + q) * nclass;
            for (int i = 0; i < nclass; ++i) o[i] = 0;
            for (int k = 0; k < nsample; ++k) {
                const int *ls = label_stat + ((size_t)bi * n + idx[((size_t)bi * m + q) * nsample + k]) * nclass;
                for (int i = 0; i < nclass; ++i) o[i] += ls[i];
            }
        }
}

/* labelstat_cuda_kernel.cu:74-105 */
ORA_API void ora_labelstat_ballrange(int b, int n, int m, float radius, int nclass, const float *new_xyz, const float *xyz, const int *label_stat, int *new_label_stat) {
    const float radius2 = radius * radius;
    for (int bi = 0; bi < b; ++bi)
        for (int q = 0; q < m; ++q) {
            const float *c = new_xyz + ((size_t)bi * m + q) * 3;
            const float *p = xyz + (size_t)bi * n * 3;
            int *o = new_label_stat + ((size_t)bi * m + q) * nclass;
            for (int i = 0; i < nclass; ++i) o[i] = 0;
            for (int k = 0; k < n; ++k) {
                float d2 = sqdist3(c[0], c[1], c[2], p[k * 3 + 0], p[k * 3 + 1], p[k * 3 + 2]);
                if (d2 < radius2) {
                    const int *ls = label_stat + ((size_t)bi * n + k) * nclass;
                    for (int i = 0; i < nclass; ++i) o[i] += ls[i];
                }
            }
        }
}

/* labelstat_cuda_kernel.cu:6-49.  idx caller-zeroed (pointops.py:337). */
ORA_API void ora_labelstat_and_ballquery(int b, int n, int m, float radius, int nsample, int nclass, const float *new_xyz, const float *xyz,
                                         const int *label_stat, int *idx, int *new_label_stat) {
    const float radius2 = radius * radius;
    for (int bi = 0; bi < b; ++bi)
        for (int q = 0; q < m; ++q) {
            const float *c = new_xyz + ((size_t)bi * m + q) * 3;
            const float *p = xyz + (size_t)bi * n * 3;
            int *o = new_label_stat + ((size_t)bi * m + q) * nclass;
            int *oi = idx + ((size_t)bi * m + q) * nsample;
            for (int i = 0; i < nclass; ++i) o[i] = 0;
            int cnt = 0;
            for (int k = 0; k < n; ++k) {
                float d2 = sqdist3(c[0], c[1], c[2], p[k * 3 + 0], p[k * 3 + 1], p[k * 3 + 2]);
                if (d2 < radius2) {
                    const int *ls = label_stat + ((size_t)bi * n + k) * nclass;
                    for (int i = 0; i < nclass; ++i) o[i] += ls[i];
                    if (cnt == 0) for (int l = 0; l < nsample; ++l) oi[l] = k;
                    oi[cnt] = k;
                    ++cnt;
                    if (cnt >= nsample) break;
                }
            }
        }
}

/* libs/chamfer_dist/chamfer.cu:15-145.  One direction: for every point of xyz1 the nearest point of
 * xyz2 (first minimum wins, strict <; tiles of 512 merged with strict > at :136).  Difference is
 * (p2 - p1) as in the source; squares make the sign irrelevant. */
ORA_API void ora_chamfer_one_direction(int batch, int n, const float *xyz1, int m, const float *xyz2, float *dist, int *indexes) {
    for (int i = 0; i < batch; ++i)
        for (int j = 0; j < n; ++j) {
            const float *a = xyz1 + ((size_t)i * n + j) * 3;
            float best = 0.f;
            int besti = 0;
            for (int k = 0; k < m; ++k) {
                const float *q = xyz2 + ((size_t)i * m + k) * 3;
                float d = sqdist3(q[0], q[1], q[2], a[0], a[1], a[2]);
                if (k == 0 || d < best) { best = d; besti = k; }
            }
            dist[(size_t)i * n + j] = best;
            indexes[(size_t)i * n + j] = besti;
        }
}

/* chamfer.cu:173-201, one launch (grad of dist1 wrt xyz1 and xyz2).  Accumulates. */
ORA_API void ora_chamfer_grad_one_direction(int b, int n, const float *xyz1, int m, const float *xyz2, const float *grad_dist1, const int *idx1,
                                            float *grad_xyz1, float *grad_xyz2) {
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < n; ++j) {
            const float *a = xyz1 + ((size_t)i * n + j) * 3;
            int j2 = idx1[(size_t)i * n + j];
            const float *q = xyz2 + ((size_t)i * m + j2) * 3;
            float g = grad_dist1[(size_t)i * n + j] * 2;
            for (int c = 0; c < 3; ++c) {
                float v = g * (a[c] - q[c]);
                grad_xyz1[((size_t)i * n + j) * 3 + c] += v;
                grad_xyz2[((size_t)i * m + j2) * 3 + c] += -v;
            }
        }
}

/* libs/KNN_CUDA/knn_cuda/csrc/cuda/knn.cu:29-93 (distance), :105-167 (insertion sort), :178-183 (sqrt),
 * host sequence :232-269.  ref is (dim, nr), query is (dim, nq), both row-major; outputs dist (k, nq)
 * = sqrt of squared distance, ind (k, nq) 1-based int64 (Python subtracts 1, knn_cuda/__init__.py:41-44).
 * ssd is a sequential fmaf chain over dim starting from 0. */
ORA_API int ora_knn_cuda(const float *ref, int nr, const float *query, int nq, int dim, int k, float *dist_out, int64_t *ind_out) {
    if (k > nr || k <= 0) return -1;
    float *col = (float *)malloc(sizeof(float) * nr);
    float *bd = (float *)malloc(sizeof(float) * k);
    int64_t *bi = (int64_t *)malloc(sizeof(int64_t) * k);
    for (int q = 0; q < nq; ++q) {
        for (int r = 0; r < nr; ++r) {
            float ssd = 0.f;
            for (int d = 0; d < dim; ++d) {
                float tmp = ref[(size_t)d * nr + r] - query[(size_t)d * nq + q];
                ssd = fmaf(tmp, tmp, ssd);
            }
            col[r] = ssd;
        }
        /* Part 1 + Part 2 of cuInsertionSort are a stable insertion keeping the k smallest */
        int cnt = 0;
        for (int l = 0; l < nr; ++l) {
            float cur = col[l];
            if (cnt < k) {
                int i = cnt;
                if (cnt > 0 && cur < bd[cnt - 1]) {
                    i = cnt - 1;
                    for (int a = 0; a < cnt - 1; ++a) if (bd[a] > cur) { i = a; break; }
                }
                for (int j = cnt; j > i; --j) { bd[j] = bd[j - 1]; bi[j] = bi[j - 1]; }
                bd[i] = cur; bi[i] = l + 1;
                ++cnt;
            } else if (cur < bd[k - 1]) {
                int i = k - 1;
                for (int a = 0; a < k - 1; ++a) if (bd[a] > cur) { i = a; break; }
                for (int j = k - 1; j > i; --j) { bd[j] = bd[j - 1]; bi[j] = bi[j - 1]; }
                bd[i] = cur; bi[i] = l + 1;
            }
        }
        for (int j = 0; j < k; ++j) {
            dist_out[(size_t)j * nq + q] = sqrtf(bd[j]);
            ind_out[(size_t)j * nq + q] = bi[j];
        }
    }
    free(col); free(bd); free(bi);
    return 0;
}
