// Hierarchy handle and cycle drivers: V(nu1,nu2) cycle over a level array, stationary iteration,
// V-cycle-preconditioned CG and the host-buffer preconditioner apply.
//
// Cycle ordering = pyamg multilevel `__solve` (the only multilevel cycle the reference uses,
// ns/preconditioner/PyAMG.py:94,119): presmooth -> r = b - A x -> b_c = R r -> recurse from
// x_c = 0 -> x += P x_c -> postsmooth; the coarsest level is solved exactly (dense inverse).
// Smoother = x += dw .* (b - A x) (ns/preconditioner/MLAMG.py:143-146), one fused pass over A per
// sweep, ping-ponging between two vectors; from a zero guess the first sweep is x = dw .* b and
// needs no pass over A.  Per level and V(1,1) cycle the operator is therefore read twice
// (residual + post-smooth) on coarse levels and on a zero-guess fine level.
#include "hierarchy.cuh"

namespace mlamg {

// a setter changed what the cycle runs: drop the cached cycle graph and the solver-loop graphs built on it
static void invalidate_graphs(mlamg_hierarchy *h) {
    if (h->gexec) { cudaGraphExecDestroy(h->gexec); h->gexec = nullptr; }
    solver_graphs_reset(h);
}

int check_handle(mlamg_hierarchy_t h) {
    if (!h) return set_error(MLAMG_EINVAL, "null hierarchy handle");
    return MLAMG_OK;
}

// smoother sweep / residual on level `lev`, SELL-32 when the level carries one, CSR otherwise
template <typename T>
static int level_jacobi(const LevelData &lev, const T *b, const T *xin, T *xout, cudaStream_t s) {
    const Csr &A = lev.A;
    if (lev.sell_ptr)
        return sell_rowop_t<T>(3, A.n, lev.sell_ptr, lev.sell_col, (const T *)lev.sell_val, xin, b, (const T *)lev.dw,
                               xout, nullptr, s);
    return jacobi_t<T>(A.n, A.nnz, A.rowptr, A.col, (const T *)A.val, (const T *)lev.dw, b, xin, xout, s);
}
template <typename T>
static int level_residual(const LevelData &lev, const T *b, const T *x, T *r, double *norm2, cudaStream_t s) {
    const Csr &A = lev.A;
    if (lev.sell_ptr)
        return sell_rowop_t<T>(2, A.n, lev.sell_ptr, lev.sell_col, (const T *)lev.sell_val, x, b, nullptr, r, norm2, s);
    return residual_t<T>(A.n, A.nnz, A.rowptr, A.col, (const T *)A.val, x, b, r, norm2, s);
}

// pipe != nullptr (host-buffer apply): b arrives / x leaves in row chunks, see HostPipe
template <typename T>
static int vcycle_enqueue(mlamg_hierarchy *h, const T *b, T *x, int nu1, int nu2, int zero_guess, cudaStream_t s,
                          const HostPipe *pipe = nullptr, char *x_host = nullptr) {
    const int L = (int)h->lv.size();
    if (L == 1) {   // single level: exact solve
        return gemv_t<T>(h->lv[0].A.n, (const T *)h->coarse_inv, b, x, s);
    }
    std::vector<T *> cur(L, nullptr);      // where level l's iterate currently lives
    std::vector<const T *> rhs(L, nullptr);
    std::vector<char> lazy(L, 0);          // x = dw.*b of that level is never materialised (see below)
    rhs[0] = b;
    // ---- downward leg
    for (int l = 0; l < L - 1; l++) {
        LevelData &lev = h->lv[l];
        const Csr &A = lev.A;
        const T *dw = (const T *)lev.dw;
        T *xa = (l == 0) ? x : (T *)lev.x;     // "home" buffer
        T *xb = (T *)lev.tmp;
        const bool zero = (l > 0) || zero_guess;
        // number of ping-pong sweeps this level will see during the whole cycle
        const int pre_pp = zero ? (nu1 > 0 ? nu1 - 1 : 0) : nu1;
        const int total_pp = pre_pp + nu2;
        T *c;
        bool fused = false;
        if (zero) {
            // first write goes where the final result lands in the home buffer after total_pp swaps
            c = (total_pp % 2 == 0) ? xa : xb;
            // V(1,*) from a zero guess: x = dw.*b and r = b - A x in ONE pass (x is never read back from HBM)
            // only for short rows (thread-per-row kernel): with several lanes per row the doubled gathers make the
            // kernel L1-bound (measured at 256^3, level 1, 30 entries/row: 70 us fused vs 56 us for the pair)
            fused = (nu1 == 1) && !lev.sell_ptr && (lev.val_scaled || (double)A.nnz <= 12.0 * (double)A.n);
            // with Q on the way up, x = dw.*b is only an operand of  dw.*(b + r) + Q e : the pass on the way down is a plain
            // residual on the scaled copy, r = b - (A D_w) b, that neither stores x nor reads dw
            lazy[l] = fused && lev.val_scaled && lev.has_Q && nu2 > 0;
            if (lazy[l] && l == 0 && pipe) {
                for (int k = 0; k < pipe->nchunks; k++) {      // rows of chunk k as soon as their columns have arrived
                    const int lo = pipe->row_lo[k], cnt = pipe->row_lo[k + 1] - lo;
                    MLAMG_CUDA(cudaStreamWaitEvent(s, pipe->in_ev[pipe->need[k]], 0));
                    if (lev.w32_a_col)
                        MLAMG_TRY(w32_residual_range_t<T>(cnt, lo, A.n, A.rowptr, lev.w32_a_col, (const T *)lev.w32_a_val, rhs[l], rhs[l],
                                                          (T *)lev.r, s));
                    else
                    MLAMG_TRY(residual_range_t<T>(cnt, lo, (long long)((double)A.nnz * cnt / A.n) + 1, A.rowptr, A.col,
                                                  (const T *)lev.val_scaled, rhs[l], rhs[l], (T *)lev.r, s));
                }
            } else if (lazy[l] && lev.w32_a_col) {
                MLAMG_TRY(w32_residual_range_t<T>(A.n, 0, A.n, A.rowptr, lev.w32_a_col, (const T *)lev.w32_a_val, rhs[l], rhs[l],
                                                  (T *)lev.r, s));
            } else if (lazy[l]) {
                MLAMG_TRY(residual_t<T>(A.n, A.nnz, A.rowptr, A.col, (const T *)lev.val_scaled, rhs[l], rhs[l], (T *)lev.r,
                                        nullptr, s));
            } else if (fused && lev.val_scaled && l == 0 && pipe) {
                for (int k = 0; k < pipe->nchunks; k++) {
                    const int lo = pipe->row_lo[k], cnt = pipe->row_lo[k + 1] - lo;
                    MLAMG_CUDA(cudaStreamWaitEvent(s, pipe->in_ev[pipe->need[k]], 0));
                    MLAMG_TRY(reszero_scaled_range_t<T>(cnt, lo, (long long)((double)A.nnz * cnt / A.n) + 1, A.rowptr, A.col,
                                                        (const T *)lev.val_scaled, dw, rhs[l], c, (T *)lev.r, s));
                }
            } else if (fused && lev.val_scaled)      // on the column-scaled copy A D_w the gathers read b alone (any row length)
                MLAMG_TRY(reszero_scaled_t<T>(A.n, A.nnz, A.rowptr, A.col, (const T *)lev.val_scaled, dw, rhs[l], c, (T *)lev.r,
                                              nullptr, s));
            else if (fused)
                MLAMG_TRY(reszero_t<T>(A.n, A.nnz, A.rowptr, A.col, (const T *)A.val, dw, rhs[l], c, (T *)lev.r, nullptr, s));
            else if (nu1 > 0) MLAMG_TRY(jacobi_zero_t<T>(A.n, dw, rhs[l], c, s));
            else MLAMG_CUDA(cudaMemsetAsync(c, 0, (size_t)A.n * sizeof(T), s));
        } else {
            c = xa;   // caller's iterate
        }
        for (int k = 0; k < pre_pp; k++) {
            T *o = (c == xa) ? xb : xa;
            MLAMG_TRY(level_jacobi<T>(lev, rhs[l], c, o, s));
            c = o;
        }
        cur[l] = c;
        if (!fused) MLAMG_TRY(level_residual<T>(lev, rhs[l], c, (T *)lev.r, nullptr, s));
        const Csr &R = lev.R;
        T *bc = (T *)h->lv[l + 1].b;
        MLAMG_TRY(spmv_perm_t<T>(R.n, R.nnz, R.rowptr, R.col, (const T *)R.val, (const T *)lev.r, bc, lev.r_order, s));
        rhs[l + 1] = bc;
    }
    // ---- coarsest level: exact solve with the dense inverse
    {
        LevelData &lev = h->lv[L - 1];
        MLAMG_TRY(gemv_t<T>(lev.A.n, (const T *)h->coarse_inv, rhs[L - 1], (T *)lev.x, s));
        cur[L - 1] = (T *)lev.x;
    }
    // ---- upward leg
    for (int l = L - 2; l >= 0; l--) {
        LevelData &lev = h->lv[l];
        const Csr &A = lev.A, &P = lev.P;
        T *xa = (l == 0) ? x : (T *)lev.x;
        T *xb = (T *)lev.tmp;
        T *c = cur[l];
        int k0 = 0;
        if (lev.has_Q && nu2 > 0) {
            // x + P e followed by one sweep  ==  x + dw.*r + Q e  (r = b - A x is still in lev.r from the way down):
            // one pass over Q instead of a pass over P and a pass over A
            const Csr &Q = lev.Q;
            T *o = (c == xa) ? xb : xa;
            if (l == 0 && pipe && nu2 == 1 && o == x && x_host) {
                for (int k = 0; k < pipe->nchunks; k++) {      // every finished chunk of the result goes to the D2H copy
                    const int lo = pipe->row_lo[k], cnt = pipe->row_lo[k + 1] - lo;
                    const long long hint = (long long)((double)Q.nnz * cnt / Q.n) + 1;
                    if (lazy[l] && lev.w32_q_col)
                        MLAMG_TRY(w32_psmooth0_range_t<T>(cnt, lo, Q.n, Q.rowptr, lev.w32_q_col, (const T *)lev.w32_q_val, cur[l + 1], rhs[l],
                                                          (const T *)lev.r, (const T *)lev.dw, o, s));
                    else if (lazy[l])
                        MLAMG_TRY(psmooth0_range_t<T>(cnt, lo, hint, Q.rowptr, Q.col, (const T *)Q.val, cur[l + 1], rhs[l],
                                                      (const T *)lev.r, (const T *)lev.dw, o, s));
                    else
                    MLAMG_TRY(psmooth_range_t<T>(cnt, lo, hint, Q.rowptr, Q.col,
                                                 (const T *)Q.val, cur[l + 1], c, (const T *)lev.r, (const T *)lev.dw, o, s));
                    MLAMG_CUDA(cudaEventRecord(pipe->out_ev[k], s));
                    MLAMG_CUDA(cudaStreamWaitEvent(pipe->cs, pipe->out_ev[k], 0));
                    MLAMG_CUDA(cudaMemcpyAsync(x_host + (size_t)lo * sizeof(T), o + lo, (size_t)cnt * sizeof(T),
                                               cudaMemcpyDeviceToHost, pipe->cs));
                }
                x_host = nullptr;      // delivered
            } else if (lazy[l] && lev.w32_q_col)
                MLAMG_TRY(w32_psmooth0_range_t<T>(Q.n, 0, Q.n, Q.rowptr, lev.w32_q_col, (const T *)lev.w32_q_val, cur[l + 1], rhs[l],
                                                  (const T *)lev.r, (const T *)lev.dw, o, s));
            else if (lazy[l])
                MLAMG_TRY(psmooth0_range_t<T>(Q.n, 0, Q.nnz, Q.rowptr, Q.col, (const T *)Q.val, cur[l + 1], rhs[l], (const T *)lev.r,
                                              (const T *)lev.dw, o, s));
            else
            MLAMG_TRY(psmooth_t<T>(Q.n, Q.nnz, Q.rowptr, Q.col, (const T *)Q.val, cur[l + 1], c, (const T *)lev.r,
                                   (const T *)lev.dw, o, s));
            c = o;
            k0 = 1;
        } else {
            MLAMG_TRY(spmv_add_t<T>(P.n, P.nnz, P.rowptr, P.col, (const T *)P.val, cur[l + 1], c, s));
        }
        for (int k = k0; k < nu2; k++) {
            T *o = (c == xa) ? xb : xa;
            MLAMG_TRY(level_jacobi<T>(lev, rhs[l], c, o, s));
            c = o;
        }
        if (l == 0 && c != x) {   // non-zero guess with an odd sweep count: one copy back
            MLAMG_CUDA(cudaMemcpyAsync(x, c, (size_t)A.n * sizeof(T), cudaMemcpyDeviceToDevice, s));
            c = x;
        }
        cur[l] = c;
    }
    if (pipe && x_host) {      // result not delivered chunk by chunk (other sweep counts): one copy behind the cycle
        MLAMG_CUDA(cudaEventRecord(pipe->out_ev[0], s));
        MLAMG_CUDA(cudaStreamWaitEvent(pipe->cs, pipe->out_ev[0], 0));
        MLAMG_CUDA(cudaMemcpyAsync(x_host, x, (size_t)h->lv[0].A.n * sizeof(T), cudaMemcpyDeviceToHost, pipe->cs));
    }
    return MLAMG_OK;
}

int vcycle_dispatch(mlamg_hierarchy *h, const void *b, void *x, int nu1, int nu2, int zero_guess,
                           cudaStream_t s) {
    MLAMG_DISPATCH(h->dtype, return vcycle_enqueue<T>(h, (const T *)b, (T *)x, nu1, nu2, zero_guess, s));
    return MLAMG_OK;
}

int vcycle_run(mlamg_hierarchy *h, const void *b, void *x, int nu1, int nu2, int zero_guess, cudaStream_t s) {
    if (!h->finalized) return set_error(MLAMG_EINVAL, "hierarchy not finalized");
    if (nu1 < 0 || nu2 < 0) return set_error(MLAMG_EINVAL, "negative sweep count");
    if (b == x) return set_error(MLAMG_EINVAL, "vcycle: b aliases x");
    if (!h->use_graph) return vcycle_dispatch(h, b, x, nu1, nu2, zero_guess, s);
    const bool hit = h->gexec && h->g_b == b && h->g_x == x && h->g_nu1 == nu1 && h->g_nu2 == nu2 &&
                     h->g_zero == (zero_guess ? 1 : 0);
    if (!hit) {
        if (h->gexec) { cudaGraphExecDestroy(h->gexec); h->gexec = nullptr; }
        if (!h->cap_stream) MLAMG_CUDA(cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking));
        cudaGraph_t g = nullptr;
        MLAMG_CUDA(cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeThreadLocal));
        int rc = vcycle_dispatch(h, b, x, nu1, nu2, zero_guess, h->cap_stream);
        cudaError_t e = cudaStreamEndCapture(h->cap_stream, &g);
        if (rc != MLAMG_OK) { if (g) cudaGraphDestroy(g); return rc; }
        if (e != cudaSuccess) return set_cuda_error(e, __FILE__, __LINE__);
        e = cudaGraphInstantiate(&h->gexec, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) { h->gexec = nullptr; return set_cuda_error(e, __FILE__, __LINE__); }
        h->g_b = b; h->g_x = x; h->g_nu1 = nu1; h->g_nu2 = nu2; h->g_zero = zero_guess ? 1 : 0;
    }
    MLAMG_CUDA(cudaGraphLaunch(h->gexec, s));
    return MLAMG_OK;
}

__global__ void __launch_bounds__(256) range_max_kernel(const int *__restrict__ col, int begin, int end, int *out) {
    int m = -1;
    for (long long j = begin + (long long)blockIdx.x * blockDim.x + threadIdx.x; j < end; j += (long long)gridDim.x * blockDim.x)
        m = max(m, col[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m >= 0) atomicMax(out, m);
}

// chunk boundaries of the host pipeline and, per chunk, the last chunk of b its rows gather from
static int ensure_pipe(mlamg_hierarchy *h, cudaStream_t s) {
    HostPipe &P = h->pipe;
    if (P.ready) return MLAMG_OK;
    const Csr &A = h->lv[0].A;
    const int n = A.n;
    int rows = (n + PIPE_CHUNKS - 1) / PIPE_CHUNKS;
    rows = (rows + 255) / 256 * 256;
    P.nchunks = 0;
    for (int lo = 0; lo < n; lo += rows) P.row_lo[P.nchunks++] = lo;
    P.row_lo[P.nchunks] = n;
    MLAMG_CUDA(cudaStreamCreateWithFlags(&P.cs, cudaStreamNonBlocking));
    MLAMG_CUDA(cudaEventCreateWithFlags(&P.fork, cudaEventDisableTiming));
    for (int k = 0; k < PIPE_CHUNKS; k++) {
        MLAMG_CUDA(cudaEventCreateWithFlags(&P.in_ev[k], cudaEventDisableTiming));
        MLAMG_CUDA(cudaEventCreateWithFlags(&P.out_ev[k], cudaEventDisableTiming));
    }
    int ptr_host[PIPE_CHUNKS + 1], maxcol[PIPE_CHUNKS];
    for (int k = 0; k <= P.nchunks; k++)
        MLAMG_CUDA(cudaMemcpyAsync(&ptr_host[k], A.rowptr + P.row_lo[k], sizeof(int), cudaMemcpyDeviceToHost, s));
    Scratch d(PIPE_CHUNKS * sizeof(int), s);
    MLAMG_SCRATCH_OK(d);
    MLAMG_CUDA(cudaMemsetAsync(d.p, 0xff, PIPE_CHUNKS * sizeof(int), s));      // -1
    MLAMG_CUDA(cudaStreamSynchronize(s));
    for (int k = 0; k < P.nchunks; k++) {
        if (ptr_host[k + 1] > ptr_host[k]) {
            range_max_kernel<<<148 * 4, 256, 0, s>>>(A.col, ptr_host[k], ptr_host[k + 1], d.as<int>() + k);
            MLAMG_LAUNCHED();
        }
    }
    MLAMG_CUDA(cudaMemcpyAsync(maxcol, d.p, PIPE_CHUNKS * sizeof(int), cudaMemcpyDeviceToHost, s));
    MLAMG_CUDA(cudaStreamSynchronize(s));
    for (int k = 0; k < P.nchunks; k++) {
        int need = k;                                   // own rows of b (epilogue) at least
        for (int j = 0; j < P.nchunks; j++)
            if (maxcol[k] >= P.row_lo[j]) need = j > need ? j : need;
        P.need[k] = need;
    }
    P.ready = true;
    return MLAMG_OK;
}

}  // namespace mlamg

using namespace mlamg;

extern "C" {

int mlamg_hierarchy_create(int dtype, int nlevels, mlamg_hierarchy_t *out) {
    if (!out || nlevels < 1 || nlevels > 64) return set_error(MLAMG_EINVAL, "hierarchy_create: bad arguments");
    if (dtype != MLAMG_F32 && dtype != MLAMG_F64) return set_error(MLAMG_EINVAL, "hierarchy_create: bad dtype");
    mlamg_hierarchy *h = new mlamg_hierarchy();
    h->dtype = dtype;
    h->esz = dtype == MLAMG_F32 ? 4 : 8;
    h->lv.resize(nlevels);
    *out = h;
    return MLAMG_OK;
}

int mlamg_hierarchy_set_operator(mlamg_hierarchy_t h, int level, int n, int nnz, const int *rowptr, const int *col,
                                 const void *val, const void *dw) {
    MLAMG_TRY(check_handle(h));
    if (level < 0 || level >= (int)h->lv.size() || n <= 0 || nnz < 0)
        return set_error(MLAMG_EINVAL, "set_operator: bad level/n/nnz");
    if (h->finalized) return set_error(MLAMG_EINVAL, "set_operator: hierarchy already finalized");
    LevelData &lev = h->lv[level];
    lev.A.n = n; lev.A.nnz = nnz; lev.A.rowptr = rowptr; lev.A.col = col; lev.A.val = val;
    lev.dw = dw;
    lev.has_A = true;
    return MLAMG_OK;
}

int mlamg_hierarchy_set_restrict_order(mlamg_hierarchy_t h, int level, const int *row_order) {
    MLAMG_TRY(check_handle(h));
    if (level < 0 || level + 1 >= (int)h->lv.size()) return set_error(MLAMG_EINVAL, "set_restrict_order: bad level");
    h->lv[level].r_order = row_order;
    invalidate_graphs(h);
    return MLAMG_OK;
}

int mlamg_hierarchy_set_operator_sell(mlamg_hierarchy_t h, int level, const int *slice_ptr, const int *scol,
                                      const void *sval) {
    MLAMG_TRY(check_handle(h));
    if (level < 0 || level >= (int)h->lv.size() || !h->lv[level].has_A)
        return set_error(MLAMG_EINVAL, "set_operator_sell: set the CSR operator of the level first");
    LevelData &lev = h->lv[level];
    lev.sell_ptr = slice_ptr; lev.sell_col = scol; lev.sell_val = sval;
    invalidate_graphs(h);
    return MLAMG_OK;
}

int mlamg_hierarchy_set_transfer(mlamg_hierarchy_t h, int level, int p_nnz, const int *p_rowptr, const int *p_col,
                                 const void *p_val, const int *r_rowptr, const int *r_col, const void *r_val) {
    MLAMG_TRY(check_handle(h));
    if (level < 0 || level + 1 >= (int)h->lv.size()) return set_error(MLAMG_EINVAL, "set_transfer: bad level");
    if (h->finalized) return set_error(MLAMG_EINVAL, "set_transfer: hierarchy already finalized");
    LevelData &lev = h->lv[level];
    if (!lev.has_A || !h->lv[level + 1].has_A)
        return set_error(MLAMG_EINVAL, "set_transfer: set both level operators first");
    lev.P.n = lev.A.n; lev.P.nnz = p_nnz; lev.P.rowptr = p_rowptr; lev.P.col = p_col; lev.P.val = p_val;
    lev.R.n = h->lv[level + 1].A.n; lev.R.nnz = p_nnz; lev.R.rowptr = r_rowptr; lev.R.col = r_col; lev.R.val = r_val;
    lev.has_PR = true;
    return MLAMG_OK;
}

int mlamg_hierarchy_set_operator_scaled(mlamg_hierarchy_t h, int level, const void *val_scaled) {
    MLAMG_TRY(check_handle(h));
    if (level < 0 || level >= (int)h->lv.size() || !h->lv[level].has_A)
        return set_error(MLAMG_EINVAL, "set_operator_scaled: set the CSR operator of the level first");
    h->lv[level].val_scaled = val_scaled;
    invalidate_graphs(h);
    return MLAMG_OK;
}

int mlamg_hierarchy_set_post_operator(mlamg_hierarchy_t h, int level, int q_nnz, const int *q_rowptr, const int *q_col,
                                      const void *q_val) {
    MLAMG_TRY(check_handle(h));
    if (level < 0 || level + 1 >= (int)h->lv.size()) return set_error(MLAMG_EINVAL, "set_post_operator: bad level");
    LevelData &lev = h->lv[level];
    if (!lev.has_A) return set_error(MLAMG_EINVAL, "set_post_operator: set the level operator first");
    lev.Q.n = lev.A.n; lev.Q.nnz = q_nnz; lev.Q.rowptr = q_rowptr; lev.Q.col = q_col; lev.Q.val = q_val;
    lev.has_Q = q_rowptr != nullptr;
    invalidate_graphs(h);
    return MLAMG_OK;
}

int mlamg_hierarchy_set_w32(mlamg_hierarchy_t h, int level, const int *a_col, const void *a_val_scaled, const int *q_col,
                            const void *q_val) {
    MLAMG_TRY(check_handle(h));
    if (level < 0 || level + 1 >= (int)h->lv.size()) return set_error(MLAMG_EINVAL, "set_w32: bad level");
    LevelData &lev = h->lv[level];
    if (!lev.has_A) return set_error(MLAMG_EINVAL, "set_w32: set the level operator first");
    if ((a_col == nullptr) != (a_val_scaled == nullptr) || (q_col == nullptr) != (q_val == nullptr))
        return set_error(MLAMG_EINVAL, "set_w32: col and val copies come in pairs");
    lev.w32_a_col = a_col; lev.w32_a_val = a_val_scaled;
    lev.w32_q_col = q_col; lev.w32_q_val = q_val;
    invalidate_graphs(h);
    return MLAMG_OK;
}

int mlamg_hierarchy_set_coarse_inverse(mlamg_hierarchy_t h, const void *inv) {
    MLAMG_TRY(check_handle(h));
    h->coarse_inv = inv;
    return MLAMG_OK;
}

int mlamg_hierarchy_finalize(mlamg_hierarchy_t h, mlamg_stream_t stream) {
    MLAMG_TRY(check_handle(h));
    (void)stream;
    if (h->finalized) return MLAMG_OK;
    const int L = (int)h->lv.size();
    for (int l = 0; l < L; l++) {
        if (!h->lv[l].has_A) return set_error(MLAMG_EINVAL, "finalize: level %d has no operator", l);
        if (l + 1 < L && !h->lv[l].has_PR) return set_error(MLAMG_EINVAL, "finalize: level %d has no P/R", l);
        if (l + 1 < L && !h->lv[l].dw) return set_error(MLAMG_EINVAL, "finalize: level %d has no smoother diagonal", l);
    }
    if (!h->coarse_inv) return set_error(MLAMG_EINVAL, "finalize: coarse inverse not set");
    for (int l = 0; l < L; l++) {
        LevelData &lev = h->lv[l];
        const size_t bytes = (size_t)lev.A.n * h->esz;
        if (l > 0) {
            MLAMG_CUDA(cudaMalloc(&lev.x, bytes));
            MLAMG_CUDA(cudaMalloc(&lev.b, bytes));
        }
        if (l + 1 < L) {
            MLAMG_CUDA(cudaMalloc(&lev.tmp, bytes));
            MLAMG_CUDA(cudaMalloc(&lev.r, bytes));
        }
    }
    h->finalized = true;
    return MLAMG_OK;
}

int mlamg_hierarchy_destroy(mlamg_hierarchy_t h) {
    if (!h) return MLAMG_OK;
    for (auto &lev : h->lv) {
        if (lev.x) cudaFree(lev.x);
        if (lev.b) cudaFree(lev.b);
        if (lev.tmp) cudaFree(lev.tmp);
        if (lev.r) cudaFree(lev.r);
    }
    solver_state_free(h);
    if (h->host_b) cudaFree(h->host_b);
    if (h->host_x) cudaFree(h->host_x);
    if (h->gexec) cudaGraphExecDestroy(h->gexec);
    if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
    if (h->pipe.cs) {
        cudaStreamDestroy(h->pipe.cs);
        cudaEventDestroy(h->pipe.fork);
        for (int k = 0; k < PIPE_CHUNKS; k++) { cudaEventDestroy(h->pipe.in_ev[k]); cudaEventDestroy(h->pipe.out_ev[k]); }
    }
    delete h;
    return MLAMG_OK;
}

int mlamg_hierarchy_use_graph(mlamg_hierarchy_t h, int enable) {
    MLAMG_TRY(check_handle(h));
    h->use_graph = enable != 0;
    if (!enable && h->gexec) { cudaGraphExecDestroy(h->gexec); h->gexec = nullptr; }
    return MLAMG_OK;
}

double mlamg_hierarchy_cycle_bytes(mlamg_hierarchy_t h, int nu1, int nu2, int zero_guess) {
    if (!h) return 0.0;
    const double v = (double)h->esz;
    const int L = (int)h->lv.size();
    double total = 0.0;
    for (int l = 0; l < L - 1; l++) {
        const LevelData &lev = h->lv[l];
        const double N = lev.A.n, nnz = (double)lev.A.nnz, Nc = h->lv[l + 1].A.n, pnnz = (double)lev.P.nnz;
        const double b_jac = nnz * (v + 4) + 4 * (N + 1) + 4 * v * N;
        const double b_res = nnz * (v + 4) + 4 * (N + 1) + 3 * v * N;
        const double b_restrict = pnnz * (v + 4) + 4 * (Nc + 1) + v * N + v * Nc;
        const double b_prolong = pnnz * (v + 4) + 4 * (N + 1) + v * Nc + 2 * v * N;
        const bool zero = (l > 0) || zero_guess;
        double pre = nu1 * b_jac;
        double res = b_res;
        if (zero && nu1 > 0) pre = (nu1 - 1) * b_jac + 3 * v * N;   // x = dw.*b : read dw,b write x
        if (zero && nu1 == 1 && !lev.sell_ptr && (lev.val_scaled || nnz <= 12.0 * N)) {   // fused x = dw.*b, r = b - A x: read A, b, dw; write x, r
            pre = 0.0;
            res = nnz * (v + 4) + 4 * (N + 1) + 4 * v * N;
        }
        double post = nu2 * b_jac + b_prolong;
        if (lev.has_Q && nu2 > 0)     // fused prolongation + first post sweep: read Q, e, x (or rhs), r, dw; write x
            post = (nu2 - 1) * b_jac + (double)lev.Q.nnz * (v + 4) + 4 * (N + 1) + v * Nc + 4 * v * N;
        if (zero && nu1 == 1 && !lev.sell_ptr && lev.val_scaled && lev.has_Q && nu2 > 0)
            res = nnz * (v + 4) + 4 * (N + 1) + 2 * v * N;   // lazy x: r = b - (A D_w) b reads A', b and writes r
        total += pre + post + res + b_restrict;
    }
    const double nc = h->lv[L - 1].A.n;
    total += nc * nc * v + 2 * nc * v;   // dense inverse GEMV
    return total;
}

int mlamg_vcycle(mlamg_hierarchy_t h, const void *b, void *x, int nu1, int nu2, int zero_guess, mlamg_stream_t stream) {
    MLAMG_TRY(check_handle(h));
    return vcycle_run(h, b, x, nu1, nu2, zero_guess, as_stream(stream));
}

int mlamg_vcycle_host(mlamg_hierarchy_t h, const void *b_host, void *x_host, int nu1, int nu2, int cycles,
                      mlamg_stream_t stream) {
    MLAMG_TRY(check_handle(h));
    cudaStream_t s = as_stream(stream);
    if (!h->finalized) return set_error(MLAMG_EINVAL, "hierarchy not finalized");
    if (cycles < 1) return set_error(MLAMG_EINVAL, "vcycle_host: cycles < 1");
    if (nu1 < 0 || nu2 < 0) return set_error(MLAMG_EINVAL, "negative sweep count");
    const size_t bytes = (size_t)h->lv[0].A.n * h->esz;
    if (!h->host_b) {
        MLAMG_CUDA(cudaMalloc(&h->host_b, bytes));
        MLAMG_CUDA(cudaMalloc(&h->host_x, bytes));
    }
    if (cycles > 1 || h->lv.size() < 2) {      // several cycles: the copies are a small share, keep the graph replay
        MLAMG_CUDA(cudaMemcpyAsync(h->host_b, b_host, bytes, cudaMemcpyHostToDevice, s));
        MLAMG_TRY(vcycle_run(h, h->host_b, h->host_x, nu1, nu2, 1, s));
        for (int c = 1; c < cycles; c++) MLAMG_TRY(vcycle_run(h, h->host_b, h->host_x, nu1, nu2, 0, s));
        MLAMG_CUDA(cudaMemcpyAsync(x_host, h->host_x, bytes, cudaMemcpyDeviceToHost, s));
        MLAMG_CUDA(cudaStreamSynchronize(s));
        return MLAMG_OK;
    }
    // one preconditioner apply (PETSc PCApply shape): chunked H2D -> first pass per chunk ... last pass per chunk -> chunked D2H
    MLAMG_TRY(ensure_pipe(h, s));
    HostPipe &P = h->pipe;
    MLAMG_CUDA(cudaEventRecord(P.fork, s));
    MLAMG_CUDA(cudaStreamWaitEvent(P.cs, P.fork, 0));
    for (int k = 0; k < P.nchunks; k++) {
        const size_t off = (size_t)P.row_lo[k] * h->esz, len = (size_t)(P.row_lo[k + 1] - P.row_lo[k]) * h->esz;
        MLAMG_CUDA(cudaMemcpyAsync((char *)h->host_b + off, (const char *)b_host + off, len, cudaMemcpyHostToDevice, P.cs));
        MLAMG_CUDA(cudaEventRecord(P.in_ev[k], P.cs));
    }
    const LevelData &l0 = h->lv[0];
    const bool chunk_in = nu1 == 1 && l0.val_scaled && !l0.sell_ptr;      // = the fused first pass of vcycle_enqueue
    if (!chunk_in) MLAMG_CUDA(cudaStreamWaitEvent(s, P.in_ev[P.nchunks - 1], 0));
    int rc = MLAMG_OK;
    if (chunk_in) {
        MLAMG_DISPATCH(h->dtype, rc = vcycle_enqueue<T>(h, (const T *)h->host_b, (T *)h->host_x, nu1, nu2, 1, s, &P, (char *)x_host));
    } else {
        MLAMG_DISPATCH(h->dtype, rc = vcycle_enqueue<T>(h, (const T *)h->host_b, (T *)h->host_x, nu1, nu2, 1, s, nullptr, nullptr));
        if (rc == MLAMG_OK) {
            MLAMG_CUDA(cudaEventRecord(P.out_ev[0], s));
            MLAMG_CUDA(cudaStreamWaitEvent(P.cs, P.out_ev[0], 0));
            MLAMG_CUDA(cudaMemcpyAsync(x_host, h->host_x, bytes, cudaMemcpyDeviceToHost, P.cs));
        }
    }
    cudaError_t e1 = cudaStreamSynchronize(s);
    cudaError_t e2 = cudaStreamSynchronize(P.cs);
    if (rc != MLAMG_OK) return rc;
    if (e1 != cudaSuccess) return set_cuda_error(e1, __FILE__, __LINE__);
    if (e2 != cudaSuccess) return set_cuda_error(e2, __FILE__, __LINE__);
    return MLAMG_OK;
}

}  // extern "C"
