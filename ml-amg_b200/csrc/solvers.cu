// Iterative drivers on top of the cycle: stationary V-cycle iteration (MLAMG.py:189-195 / multigrid.py:173-199 loop) and
// V-cycle preconditioned CG (the `accel='cg'` of pyamg's multilevel solve, PyAMG.py:119) — device resident.
//
// Every scalar of the iteration (alpha, beta, r.z, p.Ap, the residual norms, the iteration counter, the convergence flag)
// lives in device memory and is produced and consumed by kernels.  The loop itself is a CUDA-graph WHILE node
// (conditional graph node): its body is one iteration — including the captured V-cycle — and the last kernel of the body
// sets the loop condition with cudaGraphSetConditional.  A solve is therefore ONE graph launch and ONE host
// synchronisation, whatever the iteration count (round 1: three stream synchronisations per PCG iteration).
// If the driver refuses conditional nodes the same body is enqueued by the host with a look-ahead of two iterations
// (kernels of an iteration that starts after convergence return immediately), polling a pinned copy of the state.
#include <stdlib.h>
#include "hierarchy.cuh"

namespace mlamg {

template <typename T> int spmv_t(int, long long, const int *, const int *, const T *, const T *, T *, cudaStream_t);
template <typename T>
int residual_t(int, long long, const int *, const int *, const T *, const T *, const T *, T *, double *, cudaStream_t);

constexpr int SOL_THREADS = 256;
constexpr int SOL_MAX_BLOCKS = 148 * 8;
constexpr int LOOKAHEAD = 2;

struct LoopState {
    double rz, pap, rr, bb, stop, alpha, beta, tol, tmp0, tmp1;      // fields 0..9 (the mlamg_dloop_* calls address them by index)
    int it, done, maxiter, first;
};
static_assert(sizeof(LoopState) == 96, "LoopState layout is part of the mlamg_dloop_* contract");

struct LoopGraph {
    cudaGraphExec_t exec = nullptr;
    const void *b = nullptr;
    void *x = nullptr;
    int nu1 = -1, nu2 = -1, flags = -1;
    double *res_d = nullptr;
    void reset() {
        if (exec) cudaGraphExecDestroy(exec);
        exec = nullptr;
    }
};

struct SolverState {
    void *r = nullptr, *z = nullptr, *p = nullptr, *ap = nullptr;
    LoopState *d = nullptr, *hpin = nullptr;      // device state, pinned staging (hpin[0] = upload, hpin[1..] = polls)
    double *part = nullptr;
    double *res_d = nullptr;
    int res_cap = 0;
    LoopGraph pcg, sol;
    cudaStream_t cap = nullptr;
    cudaEvent_t poll_ev[LOOKAHEAD + 1] = {};
    int cond_state = 0;                           // 0 untested, 1 conditional nodes work, -1 they do not
};

void solver_state_free(mlamg_hierarchy *h) {
    SolverState *S = h->solver;
    if (!S) return;
    S->pcg.reset();
    S->sol.reset();
    for (void *q : {S->r, S->z, S->p, S->ap, (void *)S->d, (void *)S->part, (void *)S->res_d})
        if (q) cudaFree(q);
    if (S->hpin) cudaFreeHost(S->hpin);
    if (S->cap) cudaStreamDestroy(S->cap);
    for (auto &e : S->poll_ev)
        if (e) cudaEventDestroy(e);
    delete S;
    h->solver = nullptr;
}

void solver_graphs_reset(mlamg_hierarchy *h) {
    if (!h->solver) return;
    h->solver->pcg.reset();
    h->solver->sol.reset();
}

static int ensure_state(mlamg_hierarchy *h, int maxiter) {
    if (!h->solver) {
        SolverState *S = new SolverState();
        h->solver = S;
        const size_t bytes = (size_t)h->lv[0].A.n * h->esz;
        MLAMG_CUDA(cudaMalloc(&S->r, bytes));
        MLAMG_CUDA(cudaMalloc(&S->z, bytes));
        MLAMG_CUDA(cudaMalloc(&S->p, bytes));
        MLAMG_CUDA(cudaMalloc(&S->ap, bytes));
        MLAMG_CUDA(cudaMalloc(&S->d, sizeof(LoopState)));
        MLAMG_CUDA(cudaMallocHost(&S->hpin, (LOOKAHEAD + 2) * sizeof(LoopState)));
        MLAMG_CUDA(cudaMalloc(&S->part, 2 * SOL_MAX_BLOCKS * sizeof(double)));
        MLAMG_CUDA(cudaStreamCreateWithFlags(&S->cap, cudaStreamNonBlocking));
        for (auto &e : S->poll_ev) MLAMG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        const char *force = getenv("MLAMG_SOLVER_HOST_LOOP");
        if (force && force[0] == '1') S->cond_state = -1;
    }
    SolverState *S = h->solver;
    if (maxiter + 1 > S->res_cap) {
        if (S->res_d) cudaFree(S->res_d);
        S->res_d = nullptr;
        S->res_cap = 0;
        const int cap = maxiter + 1 < 256 ? 256 : maxiter + 1;
        MLAMG_CUDA(cudaMalloc(&S->res_d, (size_t)cap * sizeof(double)));
        S->res_cap = cap;
        S->pcg.reset();       // the graphs hold the old pointer
        S->sol.reset();
    }
    return MLAMG_OK;
}

static unsigned vec_blocks(int n) {
    unsigned b = cdiv(n, SOL_THREADS);
    return b > (unsigned)SOL_MAX_BLOCKS ? (unsigned)SOL_MAX_BLOCKS : (b ? b : 1u);
}

// ------------------------------------------------------------------------------------------------ kernels
template <typename T>
__global__ void __launch_bounds__(SOL_THREADS) sol_dot_kernel(int n, const T *__restrict__ x, const T *__restrict__ y,
                                                               const LoopState *__restrict__ st, double *__restrict__ partial) {
    __shared__ double sm[32];
    if (st && st->done) return;
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        acc += (double)x[i] * (double)y[i];
    acc = block_sum(acc, sm);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

__device__ __forceinline__ double sol_reduce(const double *__restrict__ partial, int nb, double *sm) {
    double a = 0.0;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) a += partial[i];
    return block_sum(a, sm);
}

// prologue: bb (optional), rr -> res[0], stop, done; also arms the WHILE condition
__global__ void __launch_bounds__(1024) sol_init_kernel(const double *__restrict__ part_bb, const double *__restrict__ part_rr,
                                                         int nb, LoopState *st, double *__restrict__ res, int relative,
                                                         int check_initial, cudaGraphConditionalHandle handle,
                                                         int use_handle) {
    __shared__ double sm[32];
    double bb = 0.0;
    if (relative) bb = sol_reduce(part_bb, nb, sm);
    const double rr = sol_reduce(part_rr, nb, sm);
    if (threadIdx.x == 0) {
        const double nb2 = sqrt(bb);
        st->bb = bb;
        st->stop = relative ? st->tol * (nb2 != 0.0 ? nb2 : 1.0) : st->tol;
        st->rr = rr;
        st->it = 0;
        st->first = 1;
        res[0] = sqrt(rr);
        const int done = (check_initial && res[0] <= st->stop) || st->maxiter <= 0;
        st->done = done;
        if (use_handle) cudaGraphSetConditional(handle, done ? 0u : 1u);
    }
}

// x -= mean(x): partial sums of x, then the shift (singular mode of multigrid.py:186-187)
template <typename T>
__global__ void __launch_bounds__(SOL_THREADS) sol_sum_kernel(int n, const T *__restrict__ x, const LoopState *__restrict__ st,
                                                               double *__restrict__ partial) {
    __shared__ double sm[32];
    if (st->done) return;
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        acc += (double)x[i];
    acc = block_sum(acc, sm);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}
template <typename T>
__global__ void __launch_bounds__(SOL_THREADS) sol_shift_kernel(int n, T *__restrict__ x, const LoopState *__restrict__ st,
                                                                 const double *__restrict__ partial, int nb) {
    __shared__ double sm[32];
    __shared__ double mean;
    if (st->done) return;
    const double total = sol_reduce(partial, nb, sm);
    if (threadIdx.x == 0) mean = total / (double)n;
    __syncthreads();
    const double m = mean;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        x[i] = (T)((double)x[i] - m);
}

// r.z -> beta = rz_new / rz (0 on the first iteration), rz = rz_new
__global__ void __launch_bounds__(1024) pcg_beta_kernel(const double *__restrict__ partial, int nb, LoopState *st) {
    __shared__ double sm[32];
    if (st->done) return;
    const double rz_new = sol_reduce(partial, nb, sm);
    if (threadIdx.x == 0) {
        st->beta = st->first ? 0.0 : rz_new / st->rz;
        st->rz = rz_new;
        st->first = 0;
    }
}

// p = z + beta p
template <typename T>
__global__ void __launch_bounds__(SOL_THREADS) pcg_direction_kernel(int n, const T *__restrict__ z, T *__restrict__ p,
                                                                     const LoopState *__restrict__ st) {
    if (st->done) return;
    const double beta = st->beta;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        p[i] = (T)((double)z[i] + beta * (double)p[i]);
}

__global__ void __launch_bounds__(1024) pcg_alpha_kernel(const double *__restrict__ partial, int nb, LoopState *st) {
    __shared__ double sm[32];
    if (st->done) return;
    const double pap = sol_reduce(partial, nb, sm);
    if (threadIdx.x == 0) {
        st->pap = pap;
        st->alpha = st->rz / pap;
    }
}

// x += alpha p, r -= alpha Ap, partial r.r
template <typename T>
__global__ void __launch_bounds__(SOL_THREADS) pcg_update_kernel(int n, const T *__restrict__ p, const T *__restrict__ ap,
                                                                  T *__restrict__ x, T *__restrict__ r,
                                                                  const LoopState *__restrict__ st, double *__restrict__ partial) {
    __shared__ double sm[32];
    if (st->done) return;
    const double alpha = st->alpha;
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        x[i] = (T)((double)x[i] + alpha * (double)p[i]);
        const T ri = (T)((double)r[i] - alpha * (double)ap[i]);
        r[i] = ri;
        acc += (double)ri * (double)ri;
    }
    acc = block_sum(acc, sm);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

// end of an iteration: rr -> res[++it], done, loop condition
__global__ void __launch_bounds__(1024) sol_check_kernel(const double *__restrict__ partial, int nb, LoopState *st,
                                                          double *__restrict__ res, cudaGraphConditionalHandle handle,
                                                          int use_handle) {
    __shared__ double sm[32];
    if (st->done) {
        if (use_handle && threadIdx.x == 0) cudaGraphSetConditional(handle, 0u);
        return;
    }
    const double rr = sol_reduce(partial, nb, sm);
    if (threadIdx.x == 0) {
        const int it = st->it + 1;
        st->it = it;
        st->rr = rr;
        res[it] = sqrt(rr);
        const int done = (res[it] <= st->stop) || it >= st->maxiter;
        st->done = done;
        if (use_handle) cudaGraphSetConditional(handle, done ? 0u : 1u);
    }
}

// ------------------------------------------------------------------------------------------------ loop pieces
struct LoopCtx {
    mlamg_hierarchy *h;
    const void *b;
    void *x;
    int nu1, nu2;
    int flags;                  // MLAMG_SOLVE_* bits (stationary iteration only)
    cudaGraphConditionalHandle handle;
    int use_handle;
};

template <typename T>
static int pcg_prologue(const LoopCtx &c, cudaStream_t s) {
    SolverState *S = c.h->solver;
    const Csr &A = c.h->lv[0].A;
    const int n = A.n;
    const unsigned vb = vec_blocks(n);
    double *part_bb = S->part, *part_rr = S->part + SOL_MAX_BLOCKS;
    MLAMG_TRY(residual_t<T>(n, A.nnz, A.rowptr, A.col, (const T *)A.val, (const T *)c.x, (const T *)c.b, (T *)S->r, nullptr, s));
    sol_dot_kernel<T><<<vb, SOL_THREADS, 0, s>>>(n, (const T *)c.b, (const T *)c.b, nullptr, part_bb);
    MLAMG_LAUNCHED();
    sol_dot_kernel<T><<<vb, SOL_THREADS, 0, s>>>(n, (const T *)S->r, (const T *)S->r, nullptr, part_rr);
    MLAMG_LAUNCHED();
    MLAMG_CUDA(cudaMemsetAsync(S->p, 0, (size_t)n * sizeof(T), s));
    sol_init_kernel<<<1, 1024, 0, s>>>(part_bb, part_rr, (int)vb, S->d, S->res_d, 1, 1, c.handle, c.use_handle);
    MLAMG_LAUNCHED();
    return MLAMG_OK;
}

template <typename T>
static int pcg_body(const LoopCtx &c, cudaStream_t s) {
    SolverState *S = c.h->solver;
    const Csr &A = c.h->lv[0].A;
    const int n = A.n;
    const unsigned vb = vec_blocks(n);
    T *r = (T *)S->r, *z = (T *)S->z, *p = (T *)S->p, *ap = (T *)S->ap;
    MLAMG_TRY(vcycle_dispatch(c.h, r, z, c.nu1, c.nu2, 1, s));
    sol_dot_kernel<T><<<vb, SOL_THREADS, 0, s>>>(n, r, z, S->d, S->part);
    MLAMG_LAUNCHED();
    pcg_beta_kernel<<<1, 1024, 0, s>>>(S->part, (int)vb, S->d);
    MLAMG_LAUNCHED();
    pcg_direction_kernel<T><<<vb, SOL_THREADS, 0, s>>>(n, z, p, S->d);
    MLAMG_LAUNCHED();
    MLAMG_TRY(spmv_t<T>(n, A.nnz, A.rowptr, A.col, (const T *)A.val, p, ap, s));
    sol_dot_kernel<T><<<vb, SOL_THREADS, 0, s>>>(n, p, ap, S->d, S->part);
    MLAMG_LAUNCHED();
    pcg_alpha_kernel<<<1, 1024, 0, s>>>(S->part, (int)vb, S->d);
    MLAMG_LAUNCHED();
    pcg_update_kernel<T><<<vb, SOL_THREADS, 0, s>>>(n, p, ap, (T *)c.x, r, S->d, S->part);
    MLAMG_LAUNCHED();
    sol_check_kernel<<<1, 1024, 0, s>>>(S->part, (int)vb, S->d, S->res_d, c.handle, c.use_handle);
    MLAMG_LAUNCHED();
    return MLAMG_OK;
}

// error measure of the stationary iteration: ||b - A x||_2 (default) or ||x||_2 (MLAMG_SOLVE_XNORM, the error_tol mode of
// multigrid.py:190-193 with b = 0)
template <typename T>
static int stat_measure(const LoopCtx &c, const LoopState *pred, cudaStream_t s) {
    SolverState *S = c.h->solver;
    const Csr &A = c.h->lv[0].A;
    const int n = A.n;
    const unsigned vb = vec_blocks(n);
    if (c.flags & MLAMG_SOLVE_XNORM) {
        sol_dot_kernel<T><<<vb, SOL_THREADS, 0, s>>>(n, (const T *)c.x, (const T *)c.x, pred, S->part);
        MLAMG_LAUNCHED();
        return MLAMG_OK;
    }
    MLAMG_TRY(residual_t<T>(n, A.nnz, A.rowptr, A.col, (const T *)A.val, (const T *)c.x, (const T *)c.b, (T *)S->r, nullptr, s));
    sol_dot_kernel<T><<<vb, SOL_THREADS, 0, s>>>(n, (const T *)S->r, (const T *)S->r, pred, S->part);
    MLAMG_LAUNCHED();
    return MLAMG_OK;
}

template <typename T>
static int stat_prologue(const LoopCtx &c, cudaStream_t s) {
    SolverState *S = c.h->solver;
    const unsigned vb = vec_blocks(c.h->lv[0].A.n);
    MLAMG_TRY(stat_measure<T>(c, nullptr, s));
    sol_init_kernel<<<1, 1024, 0, s>>>(S->part, S->part, (int)vb, S->d, S->res_d, 0, (c.flags & MLAMG_SOLVE_NO_INITIAL_CHECK) ? 0 : 1,
                                       c.handle, c.use_handle);
    MLAMG_LAUNCHED();
    return MLAMG_OK;
}

template <typename T>
static int stat_body(const LoopCtx &c, cudaStream_t s) {
    SolverState *S = c.h->solver;
    const int n = c.h->lv[0].A.n;
    const unsigned vb = vec_blocks(n);
    MLAMG_TRY(vcycle_dispatch(c.h, c.b, c.x, c.nu1, c.nu2, 0, s));
    if (c.flags & MLAMG_SOLVE_REMOVE_MEAN) {
        sol_sum_kernel<T><<<vb, SOL_THREADS, 0, s>>>(n, (const T *)c.x, S->d, S->part + SOL_MAX_BLOCKS);
        MLAMG_LAUNCHED();
        sol_shift_kernel<T><<<vb, SOL_THREADS, 0, s>>>(n, (T *)c.x, S->d, S->part + SOL_MAX_BLOCKS, (int)vb);
        MLAMG_LAUNCHED();
    }
    MLAMG_TRY(stat_measure<T>(c, S->d, s));
    sol_check_kernel<<<1, 1024, 0, s>>>(S->part, (int)vb, S->d, S->res_d, c.handle, c.use_handle);
    MLAMG_LAUNCHED();
    return MLAMG_OK;
}

typedef int (*piece_fn)(const LoopCtx &, cudaStream_t);

// prologue -> WHILE(body) as one executable graph
static int build_loop_graph(LoopCtx c, piece_fn prologue, piece_fn body, cudaGraphExec_t *out) {
    SolverState *S = c.h->solver;
    cudaStream_t cs = S->cap;
    cudaGraph_t g = nullptr;
    cudaStreamCaptureStatus status;
    const cudaGraphNode_t *deps = nullptr;
    size_t ndeps = 0;
    MLAMG_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
    int rc = MLAMG_OK;
    cudaError_t e = cudaStreamGetCaptureInfo_v2(cs, &status, nullptr, &g, &deps, &ndeps);
    if (e == cudaSuccess) e = cudaGraphConditionalHandleCreate(&c.handle, g, 0, cudaGraphCondAssignDefault);
    c.use_handle = 1;
    cudaGraph_t body_graph = nullptr;
    if (e == cudaSuccess) {
        rc = prologue(c, cs);
        if (rc == MLAMG_OK) {
            e = cudaStreamGetCaptureInfo_v2(cs, &status, nullptr, &g, &deps, &ndeps);
            cudaGraphNode_t node;
            cudaGraphNodeParams params = {};
            params.type = cudaGraphNodeTypeConditional;
            params.conditional.handle = c.handle;
            params.conditional.type = cudaGraphCondTypeWhile;
            params.conditional.size = 1;
            if (e == cudaSuccess) e = cudaGraphAddNode(&node, g, deps, ndeps, &params);
            if (e == cudaSuccess) {
                body_graph = params.conditional.phGraph_out[0];
                e = cudaStreamUpdateCaptureDependencies(cs, &node, 1, cudaStreamSetCaptureDependencies);
            }
        }
    }
    cudaGraph_t g_end = nullptr;
    cudaError_t e2 = cudaStreamEndCapture(cs, &g_end);
    if (rc != MLAMG_OK || e != cudaSuccess || e2 != cudaSuccess || !body_graph) {
        if (g_end) cudaGraphDestroy(g_end);
        cudaGetLastError();
        if (rc != MLAMG_OK) return rc;
        return set_cuda_error(e != cudaSuccess ? e : (e2 != cudaSuccess ? e2 : cudaErrorUnknown), __FILE__, __LINE__);
    }
    e = cudaStreamBeginCaptureToGraph(cs, body_graph, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) { cudaGraphDestroy(g_end); cudaGetLastError(); return set_cuda_error(e, __FILE__, __LINE__); }
    rc = body(c, cs);
    cudaGraph_t dummy = nullptr;
    e = cudaStreamEndCapture(cs, &dummy);
    if (rc == MLAMG_OK && e == cudaSuccess) e = cudaGraphInstantiate(out, g_end, 0);
    cudaGraphDestroy(g_end);
    if (rc != MLAMG_OK) return rc;
    if (e != cudaSuccess) { cudaGetLastError(); *out = nullptr; return set_cuda_error(e, __FILE__, __LINE__); }
    return MLAMG_OK;
}

// lookahead: how many iterations the host-driven fallback may run ahead of the convergence flag (0 when an iteration that
// starts after convergence would still change x)
static int run_loop(mlamg_hierarchy *h, LoopGraph &G, piece_fn prologue, piece_fn body, int lookahead, const void *b, void *x,
                    int nu1, int nu2, int flags, double tol, int maxiter, double *res_host, int *niter_host, cudaStream_t s) {
    MLAMG_TRY(ensure_state(h, maxiter));
    SolverState *S = h->solver;
    LoopState &up = S->hpin[0];
    up = LoopState();
    up.tol = tol;
    up.maxiter = maxiter;
    MLAMG_CUDA(cudaMemcpyAsync(S->d, &up, sizeof(LoopState), cudaMemcpyHostToDevice, s));
    LoopCtx c{h, b, x, nu1, nu2, flags, 0, 0};
    const bool hit = G.exec && G.b == b && G.x == x && G.nu1 == nu1 && G.nu2 == nu2 && G.flags == flags && G.res_d == S->res_d;
    if (!hit && S->cond_state >= 0) {
        G.reset();
        int rc = build_loop_graph(c, prologue, body, &G.exec);
        if (rc == MLAMG_OK) {
            S->cond_state = 1;
            G.b = b; G.x = x; G.nu1 = nu1; G.nu2 = nu2; G.flags = flags; G.res_d = S->res_d;
        } else if (S->cond_state == 0) {
            S->cond_state = -1;        // conditional nodes unavailable: host-driven loop from now on
            G.exec = nullptr;
        } else {
            return rc;
        }
    }
    LoopState fin;
    if (S->cond_state == 1) {
        MLAMG_CUDA(cudaGraphLaunch(G.exec, s));
    } else {
        // host-driven: iteration k is enqueued once the state after iteration k - LOOKAHEAD is known to be "not done"
        MLAMG_TRY(prologue(c, s));
        MLAMG_CUDA(cudaMemcpyAsync(&S->hpin[1], S->d, sizeof(LoopState), cudaMemcpyDeviceToHost, s));
        MLAMG_CUDA(cudaStreamSynchronize(s));
        bool stop = S->hpin[1].done != 0;
        for (int k = 0; !stop && k < maxiter; k++) {
            MLAMG_TRY(body(c, s));
            const int slot = k % (LOOKAHEAD + 1);
            MLAMG_CUDA(cudaMemcpyAsync(&S->hpin[1 + slot], S->d, sizeof(LoopState), cudaMemcpyDeviceToHost, s));
            MLAMG_CUDA(cudaEventRecord(S->poll_ev[slot], s));
            if (k >= lookahead) {
                const int old = (k - lookahead) % (LOOKAHEAD + 1);
                MLAMG_CUDA(cudaEventSynchronize(S->poll_ev[old]));
                stop = S->hpin[1 + old].done != 0;
            }
        }
    }
    MLAMG_CUDA(cudaMemcpyAsync(&S->hpin[1], S->d, sizeof(LoopState), cudaMemcpyDeviceToHost, s));
    MLAMG_CUDA(cudaStreamSynchronize(s));
    fin = S->hpin[1];
    MLAMG_CUDA(cudaMemcpy(res_host, S->res_d, (size_t)(fin.it + 1) * sizeof(double), cudaMemcpyDeviceToHost));
    if (niter_host) *niter_host = fin.it;
    return MLAMG_OK;
}

// ------------------------------------------------------------------------------------------------ GMRES helper
// w -= h * v with h read from device memory (modified Gram-Schmidt step)
template <typename T>
__global__ void __launch_bounds__(SOL_THREADS) mgs_axpy_kernel(int n, const T *__restrict__ v, T *__restrict__ w,
                                                                const double *__restrict__ h) {
    const double a = *h;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        w[i] = (T)((double)w[i] - a * (double)v[i]);
}
__global__ void __launch_bounds__(1024) mgs_reduce_kernel(const double *__restrict__ partial, int nb, double *__restrict__ out,
                                                           int take_sqrt) {
    __shared__ double sm[32];
    const double a = sol_reduce(partial, nb, sm);
    if (threadIdx.x == 0) *out = take_sqrt ? sqrt(a) : a;
}
template <typename T>
__global__ void __launch_bounds__(SOL_THREADS) mgs_scale_kernel(int n, const T *__restrict__ w, T *__restrict__ out,
                                                                 const double *__restrict__ nrm) {
    const double a = *nrm;
    if (a == 0.0) return;
    const double inv = 1.0 / a;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = (T)((double)w[i] * inv);
}

// ------------------------------------------------------------------------------------------------ distributed loop pieces
// The same device-resident PCG for ROW-PARTITIONED operators (mlamg/distributed.py): every rank keeps a LoopState in its
// HBM; the local dot products land in a field of the state, the caller all-reduces that field in place (NCCL on the
// device tensor, no host synchronisation) and the scalar kernels below advance the state identically on every rank.
__global__ void dloop_init_kernel(LoopState *st, double tol, int maxiter) {
    LoopState z = {};
    z.tol = tol;
    z.maxiter = maxiter;
    z.first = 1;
    *st = z;
}

// which: 0 start (tmp0 = b.b, rr = r.r all-reduced): stop, res[0], done | 1 beta (tmp0 = r.z all-reduced)
//        2 alpha (pap all-reduced) | 3 check (rr all-reduced): it, res[it], done
__global__ void dloop_scalar_kernel(LoopState *st, int which, double *__restrict__ res) {
    if (which == 0) {
        const double nb = sqrt(st->tmp0);
        st->bb = st->tmp0;
        st->stop = st->tol * (nb != 0.0 ? nb : 1.0);
        st->it = 0;
        st->first = 1;
        res[0] = sqrt(st->rr);
        st->done = (res[0] <= st->stop) || st->maxiter <= 0;
        return;
    }
    if (st->done) return;
    if (which == 1) {
        st->beta = st->first ? 0.0 : st->tmp0 / st->rz;
        st->rz = st->tmp0;
        st->first = 0;
    } else if (which == 2) {
        st->alpha = st->rz / st->pap;
    } else {
        const int it = st->it + 1;
        st->it = it;
        res[it] = sqrt(st->rr);
        st->done = (res[it] <= st->stop) || it >= st->maxiter;
    }
}

template <typename T>
static int dloop_dot_t(int n, const T *x, const T *y, LoopState *st, int field, cudaStream_t s) {
    const unsigned vb = vec_blocks(n);
    Scratch part((size_t)vb * sizeof(double), s);
    MLAMG_SCRATCH_OK(part);
    sol_dot_kernel<T><<<vb, SOL_THREADS, 0, s>>>(n, x, y, nullptr, part.as<double>());
    MLAMG_LAUNCHED();
    mgs_reduce_kernel<<<1, 1024, 0, s>>>(part.as<double>(), (int)vb, reinterpret_cast<double *>(st) + field, 0);
    MLAMG_LAUNCHED();
    return MLAMG_OK;
}

template <typename T>
static int dloop_update_t(int n, const T *p, const T *ap, T *x, T *r, LoopState *st, cudaStream_t s) {
    const unsigned vb = vec_blocks(n);
    Scratch part((size_t)vb * sizeof(double), s);
    MLAMG_SCRATCH_OK(part);
    MLAMG_CUDA(cudaMemsetAsync(part.p, 0, (size_t)vb * sizeof(double), s));      // a finished loop leaves the partials untouched
    pcg_update_kernel<T><<<vb, SOL_THREADS, 0, s>>>(n, p, ap, x, r, st, part.as<double>());
    MLAMG_LAUNCHED();
    mgs_reduce_kernel<<<1, 1024, 0, s>>>(part.as<double>(), (int)vb, &st->rr, 0);
    MLAMG_LAUNCHED();
    return MLAMG_OK;
}

template <typename T>
static int gmres_orthogonalize_t(int n, int j, const T *V, T *w, T *v_next, double *h_host, cudaStream_t s) {
    const unsigned vb = vec_blocks(n);
    Scratch part((size_t)vb * sizeof(double), s), dh((size_t)(j + 2) * sizeof(double), s);
    MLAMG_SCRATCH_OK(part);
    MLAMG_SCRATCH_OK(dh);
    double *h = dh.as<double>();
    for (int i = 0; i <= j; i++) {
        const T *vi = V + (size_t)i * (size_t)n;
        sol_dot_kernel<T><<<vb, SOL_THREADS, 0, s>>>(n, w, vi, nullptr, part.as<double>());
        MLAMG_LAUNCHED();
        mgs_reduce_kernel<<<1, 1024, 0, s>>>(part.as<double>(), (int)vb, h + i, 0);
        MLAMG_LAUNCHED();
        mgs_axpy_kernel<T><<<vb, SOL_THREADS, 0, s>>>(n, vi, w, h + i);
        MLAMG_LAUNCHED();
    }
    sol_dot_kernel<T><<<vb, SOL_THREADS, 0, s>>>(n, w, w, nullptr, part.as<double>());
    MLAMG_LAUNCHED();
    mgs_reduce_kernel<<<1, 1024, 0, s>>>(part.as<double>(), (int)vb, h + j + 1, 1);
    MLAMG_LAUNCHED();
    if (v_next) {
        mgs_scale_kernel<T><<<vb, SOL_THREADS, 0, s>>>(n, w, v_next, h + j + 1);
        MLAMG_LAUNCHED();
    }
    MLAMG_CUDA(cudaMemcpyAsync(h_host, h, (size_t)(j + 2) * sizeof(double), cudaMemcpyDeviceToHost, s));
    MLAMG_CUDA(cudaStreamSynchronize(s));
    return MLAMG_OK;
}

}  // namespace mlamg

using namespace mlamg;

extern "C" {

int mlamg_solve_ex(mlamg_hierarchy_t h, const void *b, void *x, int nu1, int nu2, int flags, double tol_abs, int maxiter,
                   double *res_host, int *niter_host, mlamg_stream_t stream) {
    MLAMG_TRY(check_handle(h));
    if (!h->finalized) return set_error(MLAMG_EINVAL, "hierarchy not finalized");
    if (maxiter < 0 || !res_host) return set_error(MLAMG_EINVAL, "solve: bad maxiter/res_host");
    if (nu1 < 0 || nu2 < 0) return set_error(MLAMG_EINVAL, "negative sweep count");
    if (b == x) return set_error(MLAMG_EINVAL, "solve: b aliases x");
    MLAMG_TRY(ensure_state(h, maxiter));
    MLAMG_DISPATCH(h->dtype, return run_loop(h, h->solver->sol, stat_prologue<T>, stat_body<T>, 0, b, x, nu1, nu2, flags, tol_abs, maxiter,
                                             res_host, niter_host, as_stream(stream)));
    return MLAMG_OK;
}

int mlamg_solve(mlamg_hierarchy_t h, const void *b, void *x, int nu1, int nu2, double tol_abs, int maxiter,
                double *res_host, int *niter_host, mlamg_stream_t stream) {
    return mlamg_solve_ex(h, b, x, nu1, nu2, 0, tol_abs, maxiter, res_host, niter_host, stream);
}

int mlamg_pcg(mlamg_hierarchy_t h, const void *b, void *x, int nu1, int nu2, double rtol, int maxiter, double *res_host,
              int *niter_host, mlamg_stream_t stream) {
    MLAMG_TRY(check_handle(h));
    if (!h->finalized) return set_error(MLAMG_EINVAL, "hierarchy not finalized");
    if (maxiter < 0 || !res_host) return set_error(MLAMG_EINVAL, "pcg: bad maxiter/res_host");
    if (nu1 < 0 || nu2 < 0) return set_error(MLAMG_EINVAL, "negative sweep count");
    MLAMG_TRY(ensure_state(h, maxiter));
    MLAMG_DISPATCH(h->dtype, return run_loop(h, h->solver->pcg, pcg_prologue<T>, pcg_body<T>, LOOKAHEAD, b, x, nu1, nu2, 0, rtol, maxiter,
                                             res_host, niter_host, as_stream(stream)));
    return MLAMG_OK;
}

int mlamg_gmres_orthogonalize(int dtype, int n, int j, const void *V, void *w, void *v_next, double *h_host,
                              mlamg_stream_t stream) {
    if (n <= 0 || j < 0 || !V || !w || !h_host) return set_error(MLAMG_EINVAL, "gmres_orthogonalize: bad arguments");
    MLAMG_DISPATCH(dtype, return gmres_orthogonalize_t<T>(n, j, (const T *)V, (T *)w, (T *)v_next, h_host, as_stream(stream)));
    return MLAMG_OK;
}

int mlamg_dloop_state_bytes(void) { return (int)sizeof(LoopState); }

int mlamg_dloop_init(void *state, double tol, int maxiter, mlamg_stream_t stream) {
    if (!state || maxiter < 0) return set_error(MLAMG_EINVAL, "dloop_init: bad arguments");
    dloop_init_kernel<<<1, 1, 0, as_stream(stream)>>>((LoopState *)state, tol, maxiter);
    MLAMG_LAUNCHED();
    return MLAMG_OK;
}

int mlamg_dloop_dot(int dtype, int n, const void *x, const void *y, void *state, int field, mlamg_stream_t stream) {
    if (!state || field < 0 || field > 9 || n < 0) return set_error(MLAMG_EINVAL, "dloop_dot: bad arguments");
    MLAMG_DISPATCH(dtype, return dloop_dot_t<T>(n, (const T *)x, (const T *)y, (LoopState *)state, field, as_stream(stream)));
    return MLAMG_OK;
}

int mlamg_dloop_scalar(void *state, int which, double *res_dev, mlamg_stream_t stream) {
    if (!state || which < 0 || which > 3 || !res_dev) return set_error(MLAMG_EINVAL, "dloop_scalar: bad arguments");
    dloop_scalar_kernel<<<1, 1, 0, as_stream(stream)>>>((LoopState *)state, which, res_dev);
    MLAMG_LAUNCHED();
    return MLAMG_OK;
}

int mlamg_dloop_direction(int dtype, int n, const void *z, void *p, const void *state, mlamg_stream_t stream) {
    if (!state || n < 0) return set_error(MLAMG_EINVAL, "dloop_direction: bad arguments");
    MLAMG_DISPATCH(dtype, (pcg_direction_kernel<T><<<vec_blocks(n), SOL_THREADS, 0, as_stream(stream)>>>(n, (const T *)z, (T *)p,
                                                                                                      (const LoopState *)state)));
    MLAMG_LAUNCHED();
    return MLAMG_OK;
}

int mlamg_dloop_update(int dtype, int n, const void *p, const void *ap, void *x, void *r, void *state, mlamg_stream_t stream) {
    if (!state || n < 0) return set_error(MLAMG_EINVAL, "dloop_update: bad arguments");
    MLAMG_DISPATCH(dtype, return dloop_update_t<T>(n, (const T *)p, (const T *)ap, (T *)x, (T *)r, (LoopState *)state, as_stream(stream)));
    return MLAMG_OK;
}

/* 1 if the solver loops of this handle run as a CUDA-graph WHILE node, -1 if the host-driven fallback is in use,
 * 0 before the first solve */
int mlamg_solver_loop_mode(mlamg_hierarchy_t h) {
    if (!h || !h->solver) return 0;
    return h->solver->cond_state;
}

}  // extern "C"
