/*
 * mlamg.h — C ABI of the B200-native aggregation-AMG hot path (libmlamg_b200.so).
 *
 * Drop-in boundary for nicknytko/ml-amg.  The reference is pure Python and has no
 * FFI of its own; each entry point below names the reference
 * expression (file:line under /root/reference) whose third-party CPU kernel
 * (scipy sparsetools / pyamg amg_core / SuperLU) it replaces.  The Python layer
 * in ml-amg_b200/ns mirrors the reference signatures and binds these with ctypes
 * (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - index arrays are int32 (what scipy produces), values are f32 or f64
 *     selected by `dtype`; CSR = (rowptr[n+1], col[nnz], val[nnz]);
 *   - every call enqueues on `stream` (a cudaStream_t passed as void*); calls
 *     that return a size/flag to the host synchronise that stream once;
 *   - return value: 0 = OK, otherwise an MLAMG_E* code, text via
 *     mlamg_last_error();
 *   - the caller owns every buffer; handles are opaque and freed by *_destroy;
 *   - there is no CPU fallback: without a CUDA device every compute call fails
 *     with MLAMG_ECUDA.
 */
#ifndef MLAMG_H
#define MLAMG_H

#ifdef __cplusplus
extern "C" {
#endif

typedef void *mlamg_stream_t;

enum { MLAMG_F32 = 0, MLAMG_F64 = 1 };
enum {
    MLAMG_OK = 0,
    MLAMG_EINVAL = 1,    /* bad argument */
    MLAMG_ECUDA = 2,     /* CUDA runtime error */
    MLAMG_ELIMIT = 3,    /* size beyond an implementation limit */
    MLAMG_ESINGULAR = 4, /* singular coarse operator (multigrid.py:166-170 returns conv=1.0) */
    MLAMG_EKEY = 5       /* label is not a centre (graph.py:83 KeyError) */
};

const char *mlamg_last_error(void);
int mlamg_version(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
long long mlamg_launch_count(void);

/* ------------------------------------------------------------------ V-cycle apply kernels */

/* y = A x.              multigrid.py:44,181,191  MLAMG.py:145,191,194 (scipy csr_matvec) */
int mlamg_spmv_csr(int dtype, int n, int nnz, const int *rowptr, const int *col, const void *val,
                   const void *x, void *y, mlamg_stream_t stream);
/* y = A x with the rows visited in the order row_order[0..n) (a permutation; NULL = natural order).
 * Same result as mlamg_spmv_csr; the order only steers cache reuse of the gathers. */
int mlamg_spmv_csr_perm(int dtype, int n, int nnz, const int *rowptr, const int *col, const void *val,
                        const void *x, void *y, const int *row_order, mlamg_stream_t stream);
/* y += A x.             prolongation x += P e_c: multigrid.py:181, MLAMG.py:191 */
int mlamg_spmv_add_csr(int dtype, int n, int nnz, const int *rowptr, const int *col, const void *val,
                       const void *x, void *y, mlamg_stream_t stream);
/* r = b - A x; if norm2 != NULL also *norm2 (device double) = ||r||_2^2 (deterministic).
 *                       multigrid.py:181,191  MLAMG.py:191,194 */
int mlamg_residual_csr(int dtype, int n, int nnz, const int *rowptr, const int *col, const void *val,
                       const void *x, const void *b, void *r, double *norm2, mlamg_stream_t stream);
/* x_out = x_in + dw .* (b - A x_in), x_out != x_in.  One fused pass over A.
 * dw = omega/diag (weighted Jacobi, MLAMG.py:104,143-146; multigrid.py:15-55; loss.py:72,88)
 * or 1/sum_j|a_ij| (L1-Jacobi, north-star addition). */
int mlamg_jacobi_csr(int dtype, int n, int nnz, const int *rowptr, const int *col, const void *val,
                     const void *dw, const void *b, const void *x_in, void *x_out,
                     mlamg_stream_t stream);
/* x_out = dw .* b ; r = b - A x_out (+ *norm2 = ||r||^2 if norm2 != NULL): the first sweep from a zero guess
 * (MLAMG.py:143-146 with x = 0; loss.py:72) fused with the residual that follows it in the cycle (MLAMG.py:190-191,
 * multigrid.py:181) — x is never read back from HBM, the gathers evaluate
 * dw[c]*b[c] on the fly. */
int mlamg_jacobi_zero_residual_csr(int dtype, int n, int nnz, const int *rowptr, const int *col, const void *val,
                                   const void *dw, const void *b, void *x_out, void *r, double *norm2,
                                   mlamg_stream_t stream);
/* the same on column-scaled values val_scaled[j] = a_ij * dw_j (A D_w on A's pattern, built once at setup): the
 * gathers read b alone, so the pass runs at the speed of a plain residual */
int mlamg_jacobi_zero_residual_scaled_csr(int dtype, int n, int nnz, const int *rowptr, const int *col,
                                          const void *val_scaled, const void *dw, const void *b, void *x_out,
                                          void *r, double *norm2, mlamg_stream_t stream);
/* x_out = x_in + dw .* r + Q e  (x_out may alias x_in): prolongation (x += P e, MLAMG.py:191, multigrid.py:181)
 * fused with the first post-smoothing sweep (MLAMG.py:192, :143-146).
 * With r = b - A x_in known (it was computed for the restriction),  (x_in + P e) followed by one sweep
 * x + dw.*(b - A x)  equals  x_in + dw.*r + Q e  with  Q = (I - D_w A) P  built once at setup — one pass over Q
 * replaces a pass over P and a pass over A.  (rowptr, col, val) is Q. */
int mlamg_prolong_smooth_csr(int dtype, int n, int nnz, const int *rowptr, const int *col, const void *val,
                             const void *e, const void *x_in, const void *r, const void *dw, void *x_out,
                             mlamg_stream_t stream);
/* W32 layout of a short-row CSR operator (the fine-level Q): rowptr is shared with the CSR matrix, inside every window
 * of 32 consecutive rows the entries are stored slot-major (all 0-th entries of the window in row order, then all 1-st
 * entries, ...): no padding, no row permutation, and the k-th entries of a warp's rows are contiguous, so the thread-per-row
 * kernel's col / val requests are coalesced (position = one ballot + popc per slot).  mlamg_csr_to_w32 writes the
 * reordered col / val copies; mlamg_prolong_smooth_zero_w32 is mlamg_prolong_smooth_zero_csr on them (same sums, same
 * order per row: bit-identical results). */
int mlamg_csr_to_w32(int dtype, int n, const int *rowptr, const int *col, const void *val, int *col_out, void *val_out,
                     mlamg_stream_t stream);
int mlamg_prolong_smooth_zero_w32(int dtype, int n, const int *rowptr, const int *col_w32, const void *val_w32,
                                  const void *e, const void *rhs, const void *r, const void *dw, void *x_out,
                                  mlamg_stream_t stream);
/* generic row-op on the W32 copies over rows [row0, row0 + nrows) of an operator with n_total rows — op codes of
 * mlamg_rowop_csr: 0 y = A x | 2 y = b - A x | 5 y = aux + dw.*b + A x | 7 y = dw.*(aux + b) + A x (any row range: windows are
 * aligned to absolute row numbers and the rows of the first / last window outside the range only take part in the ballots) */
int mlamg_rowop_w32(int dtype, int op, int nrows, int row0, int n_total, const int *rowptr, const int *col_w32,
                    const void *val_w32, const void *x, const void *b, const void *dw, void *y, void *aux,
                    mlamg_stream_t stream);
/* r = b - A x on the W32 copies */
int mlamg_residual_w32(int dtype, int n, const int *rowptr, const int *col_w32, const void *val_w32, const void *x,
                       const void *b, void *r, mlamg_stream_t stream);
/* x_out = dw .* (rhs + r) + Q e: mlamg_prolong_smooth_csr when x_in is the zero-guess sweep dw.*rhs itself and
 * r = rhs - A x_in (e.g. from mlamg_residual_csr on the column-scaled values with x = b = rhs): x_in is never
 * materialised. */
int mlamg_prolong_smooth_zero_csr(int dtype, int n, int nnz, const int *rowptr, const int *col, const void *val,
                                  const void *e, const void *rhs, const void *r, const void *dw, void *x_out,
                                  mlamg_stream_t stream);
/* x = dw .* b  (first sweep from a zero guess: no pass over A) */
int mlamg_jacobi_zero(int dtype, int n, const void *dw, const void *b, void *x, mlamg_stream_t stream);
/* smoother diagonal: mode 0 -> omega / a_ii, mode 1 -> 1 / sum_j |a_ij| (omega ignored) */
int mlamg_smoother_diag(int dtype, int mode, double omega, int n, const int *rowptr, const int *col,
                        const void *val, void *dw, mlamg_stream_t stream);
/* SELL-32 storage of an operator for the apply kernels (rows in slices of 32, column-major inside a
 * slice, padded with col = -1): thread-per-row kernels whose every col/val load is one coalesced warp
 * transaction.  slice_ptr has ceil(n/32)+1 entries (element offsets); *padded_nnz_host sizes scol/sval. */
int mlamg_sell_slice_ptr(int n, const int *rowptr, int *slice_ptr, long long *padded_nnz_host, mlamg_stream_t stream);
int mlamg_sell_fill(int dtype, int n, const int *rowptr, const int *col, const void *val, const int *slice_ptr,
                    int *scol, void *sval, mlamg_stream_t stream);
/* op: 0 y = A x | 1 y += A x | 2 y = b - A x (+ *norm2 = ||y||^2 if norm2 != NULL) | 3 y = x + dw.*(b - A x) */
int mlamg_sell_rowop(int dtype, int op, int n, const int *slice_ptr, const int *scol, const void *sval,
                     const void *x, const void *b, const void *dw, void *y, double *norm2, mlamg_stream_t stream);
/* generic row-op over the row range [row_begin, row_begin + nrows) (row_list == NULL) or the listed rows
 * row_list[0..nrows):
 * op 0 y=Ax | 1 y+=Ax | 2 y=b-Ax (+*norm2) | 3 y=x+dw.*(b-Ax) | 4 aux=dw.*b, y=b-A(dw.*b) (x unused) |
 * 5 y=aux+dw.*b+Ax (aux may alias y) | 6 = op 4 with val holding a_ij*dw_j (gathers b alone) |
 * 7 y=dw.*(aux+b)+Ax.  aux is NULL for ops 0-3.
 * Used by the row-partitioned multi-GPU levels to run interior rows while the halo exchange of x is in
 * flight, then the boundary rows. */
int mlamg_rowop_csr(int dtype, int op, int nrows, int nnz_hint, const int *rowptr, const int *col, const void *val,
                    const void *x, const void *b, const void *dw, void *y, void *aux, const int *row_list,
                    int row_begin, double *norm2, mlamg_stream_t stream);
/* halo pack: dst[i] = src[idx[i]] */
int mlamg_gather(int dtype, int n, const int *idx, const void *src, void *dst, mlamg_stream_t stream);
/* tuning hook: force the threads-per-row of the CSR kernels (1,2,4,8,16,32; 0 = staged shared-memory
 * thread-per-row kernel), -1 = heuristic; -2 / -3 = heuristic without / with the staged kernel */
int mlamg_set_csr_lanes(int lanes);
/* tuning hook: entries each lane keeps in flight per loop trip of the CSR kernels (2, 4, 8), 0 = default */
int mlamg_set_csr_batch(int nb);
/* multi-vector forms (N x k row-major block), loss.py:72,75,85,88 */
int mlamg_spmm_csr(int dtype, int n, int k, const int *rowptr, const int *col, const void *val,
                   const void *X, void *Y, double alpha, double beta, mlamg_stream_t stream);
/* Backward pass of the multi-vector loss (loss.py:95-96 `loss` -> `.backward()`, demos/1d_poisson.py:91-95;
 * the reference gets these from torch_sparse's autograd).  All blocks row-major.
 * SDDMM on a CSR pattern: out[j] = <U[row(j), 0..k), V[col(j), 0..k)>.  Gradient of the stored values of S in
 * Y = S X (U = dL/dY, V = X) and in Y = S^T X (U = X, V = dL/dY). */
int mlamg_sddmm_csr(int dtype, int n, int k, const int *rowptr, const int *col, const void *U, const void *V,
                    void *out, mlamg_stream_t stream);
/* out[j] = dense[row(j) * ncols + col(j)]: gradient of P's stored values in A_H = P^T A P from the dense
 * N x k block A P G^T + A^T P G (loss.py:53-54 backward). */
int mlamg_csr_sample_dense(int dtype, int n, int ncols, const int *rowptr, const int *col, const void *dense,
                           void *out, mlamg_stream_t stream);
/* gradient of P_hat (values on A's pattern) in P = P_hat Agg (agg_interp.py:481-484 backward):
 * g_phat[j] = g_p[entry (row(j), labels[col(j)]) of P], 0 when labels[col(j)] < 0. */
int mlamg_agg_product_backward(int dtype, int n, const int *a_rowptr, const int *a_col, const int *labels,
                               const int *p_rowptr, const int *p_col, const void *g_p, void *g_phat,
                               mlamg_stream_t stream);

/* small BLAS-1 helpers used by the cycle / PCG drivers (deterministic reductions) */
int mlamg_axpby(int dtype, int n, double alpha, const void *x, double beta, void *y, mlamg_stream_t stream);
int mlamg_dot(int dtype, int n, const void *x, const void *y, double *result, mlamg_stream_t stream);

/* exact forward Gauss-Seidel sweep (pyamg gauss_seidel, multigrid.py:175,184), level-scheduled.
 * mlamg_gs_schedule: level[i] = dependency depth of row i; order = rows sorted by (level,row);
 * level_ptr_host[nlevels+1] offsets into order.  Caller passes level_ptr_host capacity n+1. */
int mlamg_gs_schedule(int n, const int *rowptr, const int *col, int *level, int *order,
                      int *level_ptr_host, int *nlevels_host, mlamg_stream_t stream);
int mlamg_gauss_seidel(int dtype, int n, const int *rowptr, const int *col, const void *val,
                       const void *b, void *x, const int *order, const int *level_ptr_host,
                       int nlevels, mlamg_stream_t stream);

/* ------------------------------------------------------------------ hierarchy setup kernels */

/* out[0..n] = exclusive prefix sum of in[0..n-1] (out[n] = total).  in may alias out. */
int mlamg_scan_i32(const int *in, int *out, int n, mlamg_stream_t stream);

/* labels -> Agg CSR (graph.py:234-238): one unit entry per row with label >= 0.
 * col/val capacity n.  *nnz_host = number of labelled rows. */
int mlamg_agg_from_labels(int dtype, int n, const int *labels, int *rowptr, int *col, void *val,
                          int *nnz_host, mlamg_stream_t stream);
/* nearest centre NODE ID -> column rank (graph.py:76-84).  scratch_map: int[n].
 * MLAMG_EKEY if some nearest[i] is not a centre (reference raises KeyError). */
int mlamg_center_rank_labels(int n, int k, const int *centers, const int *nearest, int *scratch_map,
                             int *labels, mlamg_stream_t stream);
/* S = I - omega D^-1 A (multigrid.py:104-106); A must store its diagonal.  S shares A's rowptr; every
 * row is emitted in scipy's stored order for `eye - omega*Dinv@A` (off-diagonals in A's order, the
 * diagonal LAST) because that order fixes the rounding of P = S @ Agg. */
int mlamg_sa_smoother(int dtype, int n, const int *rowptr, const int *col, const void *val, double omega,
                      int *scol, void *sval, mlamg_stream_t stream);

/* C = A(m x k) * B(k x n), two-phase hash SpGEMM (scipy csr_matmat at multigrid.py:107,165;
 * torch.sparse.mm at agg_interp.py:484; torch_sparse.spspmm at loss.py:54).
 * symbolic: fills c_rowptr[m+1], *nnz_host.  numeric: fills c_col (sorted per row), c_val. */
int mlamg_spgemm_symbolic(int m, int k, int n, const int *a_rowptr, const int *a_col,
                          const int *b_rowptr, const int *b_col, int *c_rowptr, long long *nnz_host,
                          mlamg_stream_t stream);
int mlamg_spgemm_numeric(int dtype, int m, int k, int n, const int *a_rowptr, const int *a_col,
                         const void *a_val, const int *b_rowptr, const int *b_col, const void *b_val,
                         const int *c_rowptr, int *c_col, void *c_val, mlamg_stream_t stream);
/* B = A^T with sorted rows (P.T at multigrid.py:165,181). */
int mlamg_csr_transpose(int dtype, int m, int n, int nnz, const int *rowptr, const int *col,
                        const void *val, int *t_rowptr, int *t_col, void *t_val, mlamg_stream_t stream);
/* scipy drops entries whose sum is exactly 0.0: count, then compact. */
int mlamg_csr_nonzero_count(int dtype, int m, const int *rowptr, const void *val, int *new_rowptr,
                            long long *nnz_host, mlamg_stream_t stream);
int mlamg_csr_nonzero_fill(int dtype, int m, const int *rowptr, const int *col, const void *val,
                           const int *new_rowptr, int *new_col, void *new_val, mlamg_stream_t stream);
/* sort every row by column index in place (scipy sort_indices) */
int mlamg_csr_sort_rows(int dtype, int m, const int *rowptr, int *col, void *val, mlamg_stream_t stream);

/* dense coarse operator: D (n x n row-major, zero-filled by the call) from CSR */
int mlamg_csr_to_dense(int dtype, int n, const int *rowptr, const int *col, const void *val, void *dense,
                       mlamg_stream_t stream);
/* in-place inverse of a dense n x n f64 matrix by LU with partial pivoting (cuSOLVER getrf/getrs —
 * library call, bottom level only; replaces spla.factorized / splu, multigrid.py:168, MLAMG.py:122).
 * work: n*n doubles.  MLAMG_ESINGULAR on an exactly singular pivot. */
int mlamg_dense_inverse_f64(int n, double *a, double *work, mlamg_stream_t stream);
/* y = M x, M dense n x n row-major */
int mlamg_gemv(int dtype, int n, const void *m, const void *x, void *y, mlamg_stream_t stream);

/* ---- evolution strength of connection (pyamg.strength.evolution_strength_of_connection with its defaults; the
 * reference's 'evolution' and DEFAULT 'olson' measures, utils/common.py:27,30,58).  One kernel per step of the pyamg
 * evaluation, arithmetic non-fused and in pyamg's order (same bits as the CPU evaluation for the same rho).
 * S = I - inv_rho * D^-1 A on A's pattern (A must store its diagonal: *flags |= 1 otherwise); dinv_a_val (may be
 * NULL) receives the values of D^-1 A. */
int mlamg_evolution_step(int dtype, int n, const int *rowptr, const int *col, const void *val, double inv_rho,
                         void *s_val, void *dinv_a_val, int *flags, mlamg_stream_t stream);
/* amg_core incomplete_mat_mult_csr: S(i,j) = <A(i,:), B(:,j)> for (i,j) in S's pattern; A in CSR, B in CSC (= the
 * CSR arrays of B^T), indices sorted; products summed in increasing inner index. */
int mlamg_incomplete_matmul_csr(int dtype, int n, const int *Ap, const int *Aj, const void *Ax, const int *Bp,
                                const int *Bj, const void *Bx, const int *Sp, const int *Sj, void *Sx,
                                mlamg_stream_t stream);
/* in place: z_ij -> |1 - z_ii / z_ij|, 0 for weak ratios (< 1e-4), obtuse angles and stored zeros, 1e-4 for
 * near-perfect connections (< sqrt(eps)) */
int mlamg_evolution_measure(int dtype, int n, const int *rowptr, const int *col, void *val, mlamg_stream_t stream);
/* amg_core apply_distance_filter, in place: off-diagonals >= epsilon * (smallest off-diagonal of the row) -> 0 */
int mlamg_distance_filter(int dtype, int n, double epsilon, const int *rowptr, const int *col, void *val,
                          mlamg_stream_t stream);
/* out (on A's pattern-symmetric pattern) = 0.5 (M + M^T) [symmetrize != 0] or M, unit diagonal; 0 where neither
 * M nor M^T stores the entry */
int mlamg_evolution_symmetrize(int dtype, int n, const int *a_rowptr, const int *a_col, const int *m_rowptr,
                               const int *m_col, const void *m_val, int symmetrize, void *out,
                               mlamg_stream_t stream);
/* in place: v -> 1/v, then every row times the reciprocal of its largest |entry| (scale_rows_by_largest_entry) */
int mlamg_invert_scale_rows(int dtype, int n, const int *rowptr, void *val, mlamg_stream_t stream);
/* out[j] = E(row(j), col(j)) + w[j] where E stores that entry, w[j] otherwise (E's pattern inside A's):
 * `evolution(A) + W` of utils/common.py:27,30 */
int mlamg_csr_pattern_add(int dtype, int n, const int *a_rowptr, const int *a_col, const void *w, const int *e_rowptr,
                          const int *e_col, const void *e_val, void *out, mlamg_stream_t stream);

/* |lambda_max(D^-1 A)| to a reported accuracy (replaces ARPACK `eigs(Dinv@A, k=1, which='LM')`, multigrid.py:105).
 * symmetric: 1 = A is symmetric, 0 = it is not, -1 = test it (two SpMVs).  Symmetric A with a positive diagonal:
 * Lanczos on D^-1/2 A D^-1/2, stopped when the eigenvalue error estimate min(res, res^2/gap) <= tol*lambda (res = exact
 * residual norm of the Ritz pair); otherwise power iteration on D^-1 A stopped by the Rayleigh residual <= tol*lambda.
 * Host outputs (resid/steps/method may be NULL): the eigenvalue, the achieved relative residual of the eigenpair, the
 * operator applications used, the method (1 Lanczos, 0 power iteration).  All scratch is stream-ordered pool memory. */
int mlamg_lambda_max(int dtype, int n, long long nnz, const int *rowptr, const int *col, const void *val, double tol,
                     int max_steps, int symmetric, double *lambda_host, double *resid_host, int *steps_host,
                     int *method_host, mlamg_stream_t stream);

/* synthetic Dirichlet Poisson stencil (5-point if nz==1, else 7-point), x fastest.
 * nnz = mlamg_poisson_nnz(nx,ny,nz). */
long long mlamg_poisson_nnz(int nx, int ny, int nz);
int mlamg_poisson_csr(int dtype, int nx, int ny, int nz, int *rowptr, int *col, void *val,
                      mlamg_stream_t stream);
/* rows of the z-slab [z0, z0+nz_local) of the global grid, GLOBAL column ids (weak-scaling generator);
 * col/val capacity 7*nx*ny*nz_local, *nnz_host = entries written */
int mlamg_poisson_csr_slab(int dtype, int nx, int ny, int nz, int z0, int nz_local, int *rowptr, int *col,
                           void *val, long long *nnz_host, mlamg_stream_t stream);

/* HOST routine (no device work): the first k entries of numpy's legacy `RandomState(seed).permutation(n)` — the Lloyd
 * seeds of ns/lib/graph.py:229-231 — bit-identical to numpy (MT19937 + reversed Fisher-Yates with masked rejection
 * sampling).  out_host: k ints of host memory. */
int mlamg_legacy_permutation_head(unsigned seed, long long n, long long k, int *out_host);

/* ------------------------------------------------------------------ aggregation */

/* pyamg.graph.bellman_ford (agg_interp.py:475): nearest = seed NODE ID, -1 unreachable; bit-exact
 * emulation of the sequential in-place sweeps incl. tie-breaking.  *sweeps_host (may be NULL)
 * receives the number of sequential sweeps emulated. */
int mlamg_bellman_ford(int dtype, int n, const int *rowptr, const int *col, const void *w, int nseeds,
                       const int *seeds, void *dist, int *nearest, int *sweeps_host, mlamg_stream_t stream);
/* profiling aid: counters of the aggregation calls of this process (passes, sweeps, rows evaluated; see aggregation.cu) */
int mlamg_agg_stats(long long *out8, int reset);
/* pyamg.graph.lloyd_cluster (graph.py:232): seeds[k] in/out, clusters = seed INDEX or -1.
 * *iters_host (may be NULL) receives the Lloyd iterations executed. */
int mlamg_lloyd_cluster(int dtype, int n, const int *rowptr, const int *col, const void *w, int k,
                        int *seeds, int maxiter, void *dist, int *clusters, int *iters_host,
                        mlamg_stream_t stream);
/* ns.lib.graph.modified_bellman_ford (graph.py:7-53): push form over row-major COO edges,
 * float32 distances (+inf init), int64 labels (0 init).  CSR input = coalesced COO. */
int mlamg_modified_bellman_ford(int n, const int *rowptr, const int *col, const float *w, int ncenters,
                                const int *centers, float *dist, long long *nearest, int *passes_host,
                                mlamg_stream_t stream);

/* ------------------------------------------------------------------ hierarchy handle + cycle drivers */

typedef struct mlamg_hierarchy *mlamg_hierarchy_t;

int mlamg_hierarchy_create(int dtype, int nlevels, mlamg_hierarchy_t *out);
/* Level l operator and smoother diagonal (non-owning device pointers; caller keeps them alive). */
int mlamg_hierarchy_set_operator(mlamg_hierarchy_t h, int level, int n, int nnz, const int *rowptr,
                                 const int *col, const void *val, const void *dw);
/* optional processing order of the rows of R between level l and l+1 (see mlamg_spmv_csr_perm) */
int mlamg_hierarchy_set_restrict_order(mlamg_hierarchy_t h, int level, const int *row_order);
/* optional SELL-32 copy of level l's operator; when present the smoother and residual kernels use it */
int mlamg_hierarchy_set_operator_sell(mlamg_hierarchy_t h, int level, const int *slice_ptr, const int *scol,
                                      const void *sval);
/* P (n_l x n_{l+1}) and R = P^T (n_{l+1} x n_l) between level l and l+1 */
int mlamg_hierarchy_set_transfer(mlamg_hierarchy_t h, int level, int p_nnz, const int *p_rowptr,
                                 const int *p_col, const void *p_val, const int *r_rowptr,
                                 const int *r_col, const void *r_val);
/* optional values of A D_w (a_ij * dw_j) on level l's pattern: the zero-guess sweep + residual pass of a V(1,*)
 * cycle then gathers b alone (mlamg_jacobi_zero_residual_scaled_csr).  NULL removes it. */
int mlamg_hierarchy_set_operator_scaled(mlamg_hierarchy_t h, int level, const void *val_scaled);
/* optional Q = (I - D_w A) P of level l (rows n_l, columns n_{l+1}; D_w = diag(dw) of that level): when set, the
 * prolongation and the first post-smoothing sweep run as one pass over Q (mlamg_prolong_smooth_csr).
 * q_rowptr == NULL removes it. */
int mlamg_hierarchy_set_post_operator(mlamg_hierarchy_t h, int level, int q_nnz, const int *q_rowptr,
                                      const int *q_col, const void *q_val);
/* optional W32 copies (mlamg_csr_to_w32) of the column-scaled operator (values a_ij*dw_j) and of the post operator Q of a
 * level: the zero-guess V(1,*) cycle then runs its residual and its prolongation + post sweep with the thread-per-row W32
 * kernels.  Either pair may be NULL.  Pays on levels with >~ 100 k rows (measured at 256^3: level 1, 453 k rows of 17-30
 * entries: 47.6 -> 37.8 us and 31.1 -> 22.7 us; level 0: 288.9 -> 285.3 us for Q, no change for A). */
int mlamg_hierarchy_set_w32(mlamg_hierarchy_t h, int level, const int *a_col, const void *a_val_scaled, const int *q_col,
                            const void *q_val);
/* dense inverse of the coarsest operator (n x n row-major, in the hierarchy dtype) */
int mlamg_hierarchy_set_coarse_inverse(mlamg_hierarchy_t h, const void *inv);
/* allocate per-level work vectors (and the pinned staging buffers of the *_host entry points) */
int mlamg_hierarchy_finalize(mlamg_hierarchy_t h, mlamg_stream_t stream);
int mlamg_hierarchy_destroy(mlamg_hierarchy_t h);
/* algorithmic bytes of one V(nu1,nu2) cycle per SURVEY.md §8(d) */
double mlamg_hierarchy_cycle_bytes(mlamg_hierarchy_t h, int nu1, int nu2, int zero_guess);
/* 1: capture the cycle in a CUDA graph and replay it (launch-bound coarse levels), 0: plain launches */
int mlamg_hierarchy_use_graph(mlamg_hierarchy_t h, int enable);

/* one V(nu1,nu2) cycle on A x = b (pyamg multilevel __solve ordering; PyAMG.py:94,119).
 * zero_guess != 0: x is treated as 0 on entry (preconditioner apply). */
int mlamg_vcycle(mlamg_hierarchy_t h, const void *b, void *x, int nu1, int nu2, int zero_guess,
                 mlamg_stream_t stream);
/* Both solvers below are device resident: scalars, residual history and the convergence flag stay in HBM, the loop is a
 * CUDA-graph WHILE node whose body is one iteration (V-cycle included) — one graph launch and one host synchronisation
 * per solve.  mlamg_solver_loop_mode: 1 = WHILE node in use, -1 = host-driven fallback (driver without conditional
 * nodes), 0 = no solve yet; MLAMG_SOLVER_HOST_LOOP=1 in the environment forces the fallback.
 * stationary iteration x <- V(x, b) until ||b-Ax||_2 <= tol_abs or maxiter.  res_host[maxiter+1]
 * (entry 0 = initial residual).  MLAMG.py:189-195 / multigrid.py:173-199 loop with Jacobi smoothing. */
int mlamg_solve(mlamg_hierarchy_t h, const void *b, void *x, int nu1, int nu2, double tol_abs,
                int maxiter, double *res_host, int *niter_host, mlamg_stream_t stream);
/* the loop of ns/lib/multigrid.py:173-199 (amg_2_v) and MLAMG.py:189-195 in full: flags select the error measure
 * (XNORM: ||x||_2, the `error_tol` mode), the mean removal of the singular mode (:186-187) and whether the initial
 * iterate is tested (the reference loops always run one iteration first). */
#define MLAMG_SOLVE_XNORM 1
#define MLAMG_SOLVE_REMOVE_MEAN 2
#define MLAMG_SOLVE_NO_INITIAL_CHECK 4
int mlamg_solve_ex(mlamg_hierarchy_t h, const void *b, void *x, int nu1, int nu2, int flags, double tol_abs,
                   int maxiter, double *res_host, int *niter_host, mlamg_stream_t stream);
/* V-cycle preconditioned CG; stops when ||r||_2 <= rtol*||b||_2.  res_host[maxiter+1]. */
int mlamg_pcg(mlamg_hierarchy_t h, const void *b, void *x, int nu1, int nu2, double rtol, int maxiter,
              double *res_host, int *niter_host, mlamg_stream_t stream);
int mlamg_solver_loop_mode(mlamg_hierarchy_t h);
/* Device-resident PCG pieces for ROW-PARTITIONED operators (one process per GPU).  `state` is mlamg_dloop_state_bytes()
 * of device memory laid out as 10 doubles (fields 0 rz, 1 pap, 2 rr, 3 bb, 4 stop, 5 alpha, 6 beta, 7 tol, 8 tmp0, 9 tmp1)
 * followed by 4 ints (it, done, maxiter, first).  mlamg_dloop_dot leaves the LOCAL x.y in a field; the caller all-reduces
 * that field in place (e.g. ncclAllReduce on the device tensor — no host synchronisation) and calls mlamg_dloop_scalar:
 * which = 0 start (fields tmp0 = b.b and rr = r.r reduced): stop, res[0], done; 1 beta (tmp0 = r.z reduced);
 * 2 alpha (pap reduced); 3 check (rr reduced): it, res[it], done.  direction: p = z + beta p; update: x += alpha p,
 * r -= alpha Ap and the local r.r into field rr.  Every piece returns at once when `done` is set. */
int mlamg_dloop_state_bytes(void);
int mlamg_dloop_init(void *state, double tol, int maxiter, mlamg_stream_t stream);
int mlamg_dloop_dot(int dtype, int n, const void *x, const void *y, void *state, int field, mlamg_stream_t stream);
int mlamg_dloop_scalar(void *state, int which, double *res_dev, mlamg_stream_t stream);
int mlamg_dloop_direction(int dtype, int n, const void *z, void *p, const void *state, mlamg_stream_t stream);
int mlamg_dloop_update(int dtype, int n, const void *p, const void *ap, void *x, void *r, void *state,
                       mlamg_stream_t stream);
/* One Krylov step of the Arnoldi process of GMRES (the `accel='gmres'` of PyAMG.py:119) in ONE call and one host
 * synchronisation: modified Gram-Schmidt of w against the j+1 basis vectors V[0..j] (contiguous, stride n):
 * h[i] = w.V_i, w -= h[i] V_i in sequence; h[j+1] = ||w||_2; v_next (may be NULL) = w / h[j+1] unless the norm is 0.
 * h_host receives the j+2 Hessenberg entries. */
int mlamg_gmres_orthogonalize(int dtype, int n, int j, const void *V, void *w, void *v_next, double *h_host,
                              mlamg_stream_t stream);

/* Preconditioner apply with HOST buffers (PETSc PC apply shape: MLAMG.py:199-212, PyAMG.py:118-120):
 * H2D(b) -> `cycles` V-cycles from a zero guess -> D2H(x), all inside the call; returns after x_host
 * is complete. */
int mlamg_vcycle_host(mlamg_hierarchy_t h, const void *b_host, void *x_host, int nu1, int nu2,
                      int cycles, mlamg_stream_t stream);

/* ------------------------------------------------------------------ multi-GPU: peer-memory halo exchange
 * (the reference has no distributed solve).  One process per GPU; each rank exports one
 * window with CUDA IPC, its peers map it and write halo values straight into it over NVLink (no staging
 * buffer, no collective call inside the cycle).  Every value travels with the channel's sequence tag in
 * the same atomic unit (f64: 16-byte slot, f32: 8-byte slot), so the consumer needs no fence and no flag;
 * sequence numbers live in device memory and a whole cycle replays from a CUDA graph. */

/* cudaMalloc a zeroed window of `bytes` and export it: ipc_handle_host receives 64 bytes to ship to peers */
int mlamg_peer_alloc(long long bytes, void **ptr, void *ipc_handle_host);
/* map a peer's window from its 64-byte handle / unmap it / free an owned window */
int mlamg_peer_open(const void *ipc_handle_host, void **ptr);
int mlamg_peer_close(void *ptr);
int mlamg_peer_free(void *ptr);

typedef struct mlamg_channel *mlamg_channel_t;
/* bytes one element occupies in a receive region (value + tag) */
int mlamg_channel_slot_bytes(int dtype);
/* One exchange step.  Send side: n_send_peers segments of the send list (counts > 0), each with the two
 * remote destination addresses (sequence parity 0/1, already offset to this rank's segment of the peer's
 * region).  Receive side: n_recv slots in each of the local regions recv_region0/1 (segments of the
 * sources back to back).  state: 4 zeroed device u64 words owned by the caller (sequence number, CTA
 * counter, spare, error word: non-zero after a 20 s spin timeout).  At most 16 peers. */
int mlamg_channel_create(int n_send_peers, const int *send_counts_host, void *const *send_dst0_host,
                         void *const *send_dst1_host, int n_recv, const void *recv_region0,
                         const void *recv_region1, void *state, mlamg_channel_t *out);
int mlamg_channel_destroy(mlamg_channel_t ch);
/* pack src[send_idx[i]] (times scale[send_idx[i]] if scale != NULL; send_idx == NULL: src[i]; negative
 * index: 0, a padding slot) into the peers' regions and advance the sequence number.  Launched once per
 * use on every rank of a live channel. */
int mlamg_channel_push(mlamg_channel_t ch, int dtype, const int *send_idx, const void *src, const void *scale,
                       mlamg_stream_t stream);
/* after this use's push: wait for every slot, dst[dst_idx[i]] = slot i (dst_idx == NULL: dst[i];
 * negative: waited for, not stored) */
int mlamg_channel_unpack(mlamg_channel_t ch, int dtype, const int *dst_idx, void *dst, mlamg_stream_t stream);
/* after this use's push: mlamg_rowop_csr whose gathers of columns >= n_own read slot (col - n_own) of the
 * channel's receive region in place, waiting only for the values that have not landed yet */
int mlamg_channel_rowop(mlamg_channel_t ch, int dtype, int op, int nrows, int nnz_hint, const int *rowptr,
                        const int *col, const void *val, const void *x, int n_own, const void *b,
                        const void *dw, void *y, void *aux, const int *row_list, int row_begin,
                        mlamg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MLAMG_H */
