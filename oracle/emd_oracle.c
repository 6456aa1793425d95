/* emd_oracle.c — CPU restatement of the reference's auction-algorithm EMD.  TEST INFRASTRUCTURE (see oracle/__init__.py).
 *
 * Follows /root/reference/libs/emd_module/emd_cuda.cu:
 *   host loop                      :256-269   per round: clear, calc_unass_cnt(_sum), calc_unass_idx, Bid, GetMax, Assign
 *   Bid                            :95-179    best / second-best value over all objects, increment = best - better + eps,
 *                                             float atomicMax of the increment per object (:10-20)
 *   GetMax                         :181-194   the bidder whose increment equals the object's maximum within +-1e-6 (double)
 *   Assign                         :196-215   winner takes the object, previous owner is unassigned, price += increment,
 *                                             max_increments := -1e9; in the LAST round every bidder takes its bid
 *   CalcDist                       :217-226   squared distance to the assigned object
 *
 * Arithmetic (nvcc 12.9 -O2 contraction of the reference source):
 *   sq    = fmaf(dz,dz, fmaf(dx,dx, dy*dy))  on (p2 - p1)
 *   value = (float)((3.0 - (double)sqrtf(sq)) - (double)price)        -- the literal 3.0 is a double (:146)
 *   inc   = (best - better) + eps                                     -- float
 * The per-thread split of the objects in Bid (:106-108 thread_per_unass = 1024 / ceil(unassigned / (n/1024)); :136-138
 * delta = ceil(tile / thread_per_unass) per 2048-object tile) does not change the two largest VALUES, but it decides which
 * of several objects with exactly equal value is picked: a thread keeps the first maximum of its own scan order (its
 * slice of tile 0, then its slice of tile 1, ...), threads merge in ascending order with a strict '>' (:163-170).
 * The restatement therefore walks the objects in that (thread, tile, position) order.
 *
 * Where the reference is order-dependent — two bidders on one object whose increments both lie within 1e-6 of the
 * maximum (GetMax: last writer wins), several bidders on one object in the last round (Assign: last writer of
 * assignment_inv, racy price +=) — this restatement fixes the documented rule of the new kernel
 * (patchaugnet_b200/csrc/emd.cu): the lowest bidder index wins, last-round bidders are applied in ascending order.
 * Without such ties it is the reference's result exactly.  Parity is pinned on the GPU box against the reference's own
 * compiled kernels (tests/test_refgpu.py: identical assignments whenever the run met no tie).
 */
#include <math.h>
#include <stdlib.h>

#define ORA_API __attribute__((visibility("default")))

static float sqdist(const float *p2, const float *p1) {
    const float dx = p2[0] - p1[0], dy = p2[1] - p1[1], dz = p2[2] - p1[2];
    return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
}

/* returns 1 ok, -1 bad shape (emd_cuda.cu:236-249); rounds_used[i] = rounds cloud i ran before everything was assigned;
 * ties[i] = number of GetMax decisions in cloud i that had more than one candidate (0 => identical to the reference). */
ORA_API int ora_emd_forward(int b, int n, const float *xyz1, const float *xyz2, float eps, int iters, float *dist,
                            int *assignment, float *price, int *rounds_used, int *ties) {
    if (b > 512 || n % 1024 != 0 || b < 0 || n <= 0) return -1;
    int *assignment_inv = (int *)malloc(sizeof(int) * n), *bid = (int *)malloc(sizeof(int) * n);
    int *unass = (int *)malloc(sizeof(int) * n), *max_idx = (int *)malloc(sizeof(int) * n);
    int *ncand = (int *)malloc(sizeof(int) * n);
    float *bid_inc = (float *)malloc(sizeof(float) * n), *max_inc = (float *)malloc(sizeof(float) * n);
    for (int i = 0; i < b; ++i) {
        const float *p1 = xyz1 + (size_t)i * n * 3, *p2 = xyz2 + (size_t)i * n * 3;
        int *asg = assignment + (size_t)i * n;
        float *pr = price + (size_t)i * n, *ds = dist + (size_t)i * n;
        for (int j = 0; j < n; ++j) { asg[j] = -1; assignment_inv[j] = -1; pr[j] = 0.f; max_inc[j] = 0.f; bid[j] = 0; bid_inc[j] = 0.f; }
        int used = 0, tie_count = 0;
        for (int it = 0; it < iters; ++it) {
            const int last = it == iters - 1;
            int total = 0;
            for (int j = 0; j < n; ++j) if (asg[j] == -1) unass[total++] = j;
            if (total == 0) break;
            ++used;
            const int block_cnt = n / 1024, upb = (total + block_cnt - 1) / block_cnt, tpu = 1024 / upb;
            for (int u = 0; u < total; ++u) {                              /* Bid */
                const int j = unass[u];
                float best = -1e9f, better = -1e9f;
                int best_i = -1;
                for (int sub = 0; sub < tpu; ++sub)                        /* ascending thread merge */
                    for (int k2 = 0; k2 < n; k2 += 2048) {                 /* tiles in order within a thread */
                        const int end_k = (n < k2 + 2048 ? n : k2 + 2048) - k2;
                        const int delta = (end_k + tpu - 1) / tpu;
                        const int l = sub * delta, r = (sub + 1) * delta < end_k ? (sub + 1) * delta : end_k;
                        for (int kk = l; kk < r; ++kk) {
                            const int k = k2 + kk;
                            const float d = (float)((3.0 - (double)sqrtf(sqdist(p2 + 3 * k, p1 + 3 * j))) - (double)pr[k]);
                            if (d > best) { better = best; best = d; best_i = k; }
                            else if (d > better) better = d;
                        }
                    }
                const float inc = (best - better) + eps;
                bid[j] = best_i;
                bid_inc[j] = inc;
                if (inc > max_inc[best_i]) max_inc[best_i] = inc;
            }
            for (int u = 0; u < total; ++u) { max_idx[bid[unass[u]]] = 0x7fffffff; ncand[bid[unass[u]]] = 0; }
            for (int u = 0; u < total; ++u) {                              /* GetMax, lowest bidder index on ties */
                const int j = unass[u], o = bid[j];
                const double bi = (double)bid_inc[j], mi = (double)max_inc[o];
                if (bi - 1e-6 <= mi && mi <= bi + 1e-6) {
                    if (j < max_idx[o]) max_idx[o] = j;
                    if (++ncand[o] == 2) ++tie_count;
                }
            }
            for (int u = 0; u < total; ++u) {                              /* Assign */
                const int j = unass[u], o = bid[j];
                if (last || max_idx[o] == j) {
                    const int prev = assignment_inv[o];
                    if (!last && prev != -1) asg[prev] = -1;
                    assignment_inv[o] = j;
                    asg[j] = o;
                    pr[o] += bid_inc[j];
                    max_inc[o] = -1e9f;
                }
            }
        }
        for (int j = 0; j < n; ++j) ds[j] = asg[j] >= 0 ? sqdist(p1 + 3 * j, p2 + 3 * asg[j]) : 0.f;
        if (rounds_used) rounds_used[i] = used;
        if (ties) ties[i] = tie_count;
    }
    free(assignment_inv); free(bid); free(unass); free(max_idx); free(ncand); free(bid_inc); free(max_inc);
    return 1;
}
