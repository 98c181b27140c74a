/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product path.
 *
 * CPU restatement (plain C) of the pyamg 4.x `amg_core` loops that the
 * reference's aggregation / relaxation path calls through pyamg:
 *
 *   pyamg.graph.lloyd_cluster      <- /root/reference/ns/lib/graph.py:232
 *   pyamg.graph.bellman_ford       <- /root/reference/ns/model/agg_interp.py:475,
 *                                     /root/reference/demos/matconv.py:142
 *   pyamg.relaxation.relaxation.gauss_seidel
 *                                  <- /root/reference/ns/lib/multigrid.py:175,184
 *
 * pyamg is a third-party dependency that is NOT vendored in /root/reference,
 * NOT pinned (docker/Dockerfile:5 `pip3 install pyamg torch`) and NOT
 * installable here (no network, no wheel).  API evidence (3-tuple return of
 * lloyd_cluster, 2-tuple return of bellman_ford) pins it to the 4.x line.
 * The loops below restate the published 4.x algorithm (pyamg/amg_core/graph.h,
 * relaxation.h): **PARITY UNPINNED** for these three routines — there is no
 * pyamg binary or golden vector to diff against in this container.
 *
 * Semantics restated:
 *   bellman_ford sweep : in-place (Gauss-Seidel order) pull relaxation over CSR
 *                        rows, strict `<`, label copied from the neighbour that
 *                        produced the strict improvement.
 *   lloyd_cluster      : reset; BF outward to a fixed point; mark cluster
 *                        boundary nodes (any neighbour with a different label)
 *                        distance 0; BF inward to a fixed point; move each seed
 *                        to the first node (index order) of strictly largest
 *                        inward distance.
 *   gauss_seidel       : forward/backward in-place sweep, zero diagonals skipped.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define DEFINE_ALL(T, SUF, TMAX)                                                         \
                                                                                         \
void oracle_bf_sweep_##SUF(int n, const int *Ap, const int *Aj, const T *Ax,             \
                           T *x, int *z)                                                 \
{                                                                                        \
    for (int i = 0; i < n; i++) {                                                        \
        T xi = x[i];                                                                     \
        int zi = z[i];                                                                   \
        for (int jj = Ap[i]; jj < Ap[i + 1]; jj++) {                                     \
            const int j = Aj[jj];                                                        \
            const T d = Ax[jj] + x[j];                                                   \
            if (d < xi) { xi = d; zi = z[j]; }                                           \
        }                                                                                \
        x[i] = xi;                                                                       \
        z[i] = zi;                                                                       \
    }                                                                                    \
}                                                                                        \
                                                                                         \
/* returns the number of sweeps executed (including the final no-change sweep) */        \
int oracle_bf_fixed_point_##SUF(int n, const int *Ap, const int *Aj, const T *Ax,        \
                                T *x, int *z, T *old)                                    \
{                                                                                        \
    int sweeps = 0;                                                                      \
    do {                                                                                 \
        memcpy(old, x, (size_t)n * sizeof(T));                                           \
        oracle_bf_sweep_##SUF(n, Ap, Aj, Ax, x, z);                                      \
        sweeps++;                                                                        \
    } while (memcmp(old, x, (size_t)n * sizeof(T)) != 0);                                \
    return sweeps;                                                                       \
}                                                                                        \
                                                                                         \
/* pyamg.graph.bellman_ford(G, seeds): labels are seed NODE IDS, unreachable -> -1 */    \
int oracle_bellman_ford_##SUF(int n, const int *Ap, const int *Aj, const T *Ax,          \
                              int nseeds, const int *seeds, T *dist, int *nearest)       \
{                                                                                        \
    T *old = (T *)malloc((size_t)(n > 0 ? n : 1) * sizeof(T));                           \
    for (int i = 0; i < n; i++) { dist[i] = TMAX; nearest[i] = -1; }                     \
    for (int s = 0; s < nseeds; s++) { dist[seeds[s]] = 0; nearest[seeds[s]] = seeds[s]; } \
    int sw = oracle_bf_fixed_point_##SUF(n, Ap, Aj, Ax, dist, nearest, old);             \
    free(old);                                                                           \
    return sw;                                                                           \
}                                                                                        \
                                                                                         \
/* one call of amg_core.lloyd_cluster: x=distances (out), w=clusters (out),              \
 * z=seeds (in/out).  Labels are seed INDICES 0..num_seeds-1. */                         \
void oracle_lloyd_cluster_step_##SUF(int n, const int *Ap, const int *Aj, const T *Ax,   \
                                     int num_seeds, T *x, int *w, int *z)                \
{                                                                                        \
    T *old = (T *)malloc((size_t)(n > 0 ? n : 1) * sizeof(T));                           \
    for (int i = 0; i < n; i++) { x[i] = TMAX; w[i] = -1; }                              \
    for (int i = 0; i < num_seeds; i++) { x[z[i]] = 0; w[z[i]] = i; }                    \
    oracle_bf_fixed_point_##SUF(n, Ap, Aj, Ax, x, w, old);                               \
    for (int i = 0; i < n; i++) x[i] = TMAX;                                             \
    for (int i = 0; i < n; i++) {                                                        \
        for (int jj = Ap[i]; jj < Ap[i + 1]; jj++) {                                     \
            if (w[i] != w[Aj[jj]]) { x[i] = 0; break; }                                  \
        }                                                                                \
    }                                                                                    \
    oracle_bf_fixed_point_##SUF(n, Ap, Aj, Ax, x, w, old);                               \
    for (int i = 0; i < n; i++) {                                                        \
        const int seed = w[i];                                                           \
        if (seed == -1) continue;                                                        \
        if (x[z[seed]] < x[i]) z[seed] = i;                                              \
    }                                                                                    \
    free(old);                                                                           \
}                                                                                        \
                                                                                         \
/* pyamg.graph.lloyd_cluster(G, seeds, maxiter) outer loop; returns iterations run */    \
int oracle_lloyd_cluster_##SUF(int n, const int *Ap, const int *Aj, const T *Ax,         \
                               int num_seeds, int maxiter, T *x, int *w, int *z)         \
{                                                                                        \
    int *last = (int *)malloc((size_t)(num_seeds > 0 ? num_seeds : 1) * sizeof(int));    \
    int it = 0;                                                                          \
    for (it = 0; it < maxiter; it++) {                                                   \
        memcpy(last, z, (size_t)num_seeds * sizeof(int));                                \
        oracle_lloyd_cluster_step_##SUF(n, Ap, Aj, Ax, num_seeds, x, w, z);              \
        if (memcmp(last, z, (size_t)num_seeds * sizeof(int)) == 0) { it++; break; }      \
    }                                                                                    \
    free(last);                                                                          \
    return it;                                                                           \
}                                                                                        \
                                                                                         \
/* amg_core gauss_seidel(Ap,Aj,Ax,x,b,row_start,row_stop,row_step) */                    \
void oracle_gauss_seidel_##SUF(const int *Ap, const int *Aj, const T *Ax, T *x,          \
                               const T *b, int row_start, int row_stop, int row_step)    \
{                                                                                        \
    for (int i = row_start; i != row_stop; i += row_step) {                              \
        T rsum = 0, diag = 0;                                                            \
        for (int jj = Ap[i]; jj < Ap[i + 1]; jj++) {                                     \
            const int j = Aj[jj];                                                        \
            if (i == j) diag = Ax[jj];                                                   \
            else rsum += Ax[jj] * x[j];                                                  \
        }                                                                                \
        if (diag != (T)0.0) x[i] = (b[i] - rsum) / diag;                                 \
    }                                                                                    \
}                                                                                        \
                                                                                         \
/* amg_core jacobi: x_i = (1-w) t_i + w (b_i - sum_{j!=i} a_ij t_j)/a_ii, t = old x */   \
void oracle_jacobi_##SUF(const int *Ap, const int *Aj, const T *Ax, T *x, const T *b,    \
                         T *temp, int n, T omega)                                        \
{                                                                                        \
    const T one = 1.0;                                                                   \
    memcpy(temp, x, (size_t)n * sizeof(T));                                              \
    for (int i = 0; i < n; i++) {                                                        \
        T rsum = 0, diag = 0;                                                            \
        for (int jj = Ap[i]; jj < Ap[i + 1]; jj++) {                                     \
            const int j = Aj[jj];                                                        \
            if (i == j) diag = Ax[jj];                                                   \
            else rsum += Ax[jj] * temp[j];                                               \
        }                                                                                \
        if (diag != (T)0.0) x[i] = (one - omega) * temp[i] + omega * ((b[i] - rsum) / diag); \
    }                                                                                    \
}

DEFINE_ALL(double, f64, DBL_MAX)
DEFINE_ALL(float, f32, FLT_MAX)

/*
 * Restatement of the reference's OWN pure-Python push-form Bellman-Ford,
 * /root/reference/ns/lib/graph.py:7-53 (`modified_bellman_ford`): edges are
 * visited in the coalesced (row-major) COO order; distance is a float32 torch
 * tensor initialised to +inf, nearest_center an int64 tensor initialised to 0;
 * strict `<`; repeated until a full pass changes nothing.
 * `dist[i] + w` is evaluated in the promoted type of (float32, W).
 */
#include <math.h>
#define DEFINE_MBF(W, ACC, SUF)                                                           \
int oracle_modified_bf_##SUF(int n, long long nnz, const long long *row,                  \
                             const long long *col, const W *w, int ncenters,              \
                             const long long *centers, float *dist, long long *nearest)   \
{                                                                                         \
    for (int i = 0; i < n; i++) { dist[i] = INFINITY; nearest[i] = 0; }                   \
    for (int c = 0; c < ncenters; c++) { dist[centers[c]] = 0; nearest[centers[c]] = centers[c]; } \
    int passes = 0;                                                                       \
    for (;;) {                                                                            \
        int finished = 1;                                                                 \
        for (long long e = 0; e < nnz; e++) {                                             \
            const long long i = row[e], j = col[e];                                       \
            const ACC cand = (ACC)dist[i] + (ACC)w[e];                                    \
            if (cand < (ACC)dist[j]) {                                                    \
                dist[j] = (float)cand;                                                    \
                nearest[j] = nearest[i];                                                  \
                finished = 0;                                                             \
            }                                                                             \
        }                                                                                 \
        passes++;                                                                         \
        if (finished) break;                                                              \
        /* float32 distances relaxed with float64 weights can be rounded UP on the store and */ \
        /* then "improve" for ever (the reference's loop never ends there): give up loudly   */ \
        if (passes > 4 * n + 16) return -passes;                                          \
    }                                                                                     \
    return passes;                                                                        \
}
DEFINE_MBF(float, float, f32)
DEFINE_MBF(double, double, f64)

/*
 * pyamg 4.x strength-of-connection helpers behind
 * `pyamg.strength.evolution_strength_of_connection` — called by the reference at
 * /root/reference/utils/common.py:27,30 (`'evolution'` and the default `'olson'` measure).
 * Restated from the published amg_core (`linalg.h` / `smoothed_aggregation.h` / `evolution_strength.h`);
 * PARITY UNPINNED like the loops above.
 *
 *   incomplete_mat_mult_csr : S(i,j) = <A(i,:), B(:,j)> for (i,j) in the pattern of S only; A in CSR, B in CSC,
 *                             both with sorted indices: a merge of the two index lists, products summed in
 *                             increasing inner index, starting from 0.0.
 *   apply_distance_filter   : per row, threshold = epsilon * (smallest off-diagonal value); every OFF-diagonal
 *                             entry >= threshold is set to 0.
 *   maximum_row_value       : largest |entry| of every row.
 */
void oracle_incomplete_mat_mult_csr_f64(const int *Ap, const int *Aj, const double *Ax,
                                        const int *Bp, const int *Bj, const double *Bx,
                                        const int *Sp, const int *Sj, double *Sx, int num_rows)
{
    for (int row = 0; row < num_rows; row++) {
        for (int k = Sp[row]; k < Sp[row + 1]; k++) {
            const int col = Sj[k];
            double sum = 0.0;
            int a = Ap[row], a_end = Ap[row + 1];
            int b = Bp[col], b_end = Bp[col + 1];
            while (a < a_end && b < b_end) {
                const int ac = Aj[a], br = Bj[b];
                if (ac == br) { sum += Ax[a] * Bx[b]; a++; b++; }
                else if (ac < br) a++;
                else b++;
            }
            Sx[k] = sum;
        }
    }
}

void oracle_apply_distance_filter_f64(int n_row, double epsilon, const int *Sp, const int *Sj, double *Sx)
{
    for (int i = 0; i < n_row; i++) {
        double min_offdiagonal = DBL_MAX;
        for (int jj = Sp[i]; jj < Sp[i + 1]; jj++)
            if (Sj[jj] != i && Sx[jj] < min_offdiagonal) min_offdiagonal = Sx[jj];
        const double threshold = epsilon * min_offdiagonal;
        for (int jj = Sp[i]; jj < Sp[i + 1]; jj++)
            if (Sx[jj] >= threshold && Sj[jj] != i) Sx[jj] = 0.0;
    }
}

void oracle_maximum_row_value_f64(int n_row, double *x, const int *Sp, const int *Sj, const double *Sx)
{
    (void)Sj;
    for (int i = 0; i < n_row; i++) {
        double m = DBL_MIN;          /* std::numeric_limits<double>::min() in the original */
        for (int jj = Sp[i]; jj < Sp[i + 1]; jj++) {
            const double v = fabs(Sx[jj]);
            if (v > m) m = v;
        }
        x[i] = m;
    }
}
