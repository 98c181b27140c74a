// Hierarchy handle shared by the cycle driver (hierarchy.cu) and the iterative solvers (solvers.cu).
#pragma once
#include <math.h>
#include <vector>
#include "common.cuh"

namespace mlamg {

template <typename T> int spmv_t(int, long long, const int *, const int *, const T *, const T *, T *, cudaStream_t);
template <typename T>
int spmv_perm_t(int, long long, const int *, const int *, const T *, const T *, T *, const int *, cudaStream_t);
template <typename T> int spmv_add_t(int, long long, const int *, const int *, const T *, const T *, T *, cudaStream_t);
template <typename T>
int residual_t(int, long long, const int *, const int *, const T *, const T *, const T *, T *, double *, cudaStream_t);
template <typename T>
int jacobi_t(int, long long, const int *, const int *, const T *, const T *, const T *, const T *, T *, cudaStream_t);
template <typename T> int jacobi_zero_t(int, const T *, const T *, T *, cudaStream_t);
template <typename T>
int reszero_t(int, long long, const int *, const int *, const T *, const T *, const T *, T *, T *, double *, cudaStream_t);
template <typename T>
int reszero_scaled_t(int, long long, const int *, const int *, const T *, const T *, const T *, T *, T *, double *, cudaStream_t);
template <typename T>
int psmooth0_range_t(int, int, long long, const int *, const int *, const T *, const T *, const T *, const T *, const T *, T *,
                     cudaStream_t);
template <typename T>
int residual_range_t(int, int, long long, const int *, const int *, const T *, const T *, const T *, T *, cudaStream_t);
template <typename T>
int psmooth_range_t(int, int, long long, const int *, const int *, const T *, const T *, const T *, const T *, const T *, T *,
                    cudaStream_t);
template <typename T>
int reszero_scaled_range_t(int, int, long long, const int *, const int *, const T *, const T *, const T *, T *, T *, cudaStream_t);
template <typename T>
int psmooth_t(int, long long, const int *, const int *, const T *, const T *, const T *, const T *, const T *, T *, cudaStream_t);
template <typename T>
int sell_rowop_t(int, int, const int *, const int *, const T *, const T *, const T *, const T *, T *, double *,
                 cudaStream_t);
template <typename T> int gemv_t(int, const T *, const T *, T *, cudaStream_t);
template <typename T>
int w32_psmooth0_range_t(int, int, int, const int *, const int *, const T *, const T *, const T *, const T *, const T *, T *, cudaStream_t);
template <typename T>
int w32_residual_range_t(int, int, int, const int *, const int *, const T *, const T *, const T *, T *, cudaStream_t);

struct Csr {
    int n = 0;          // rows
    long long nnz = 0;
    const int *rowptr = nullptr;
    const int *col = nullptr;
    const void *val = nullptr;
};

struct LevelData {
    Csr A, P, R;
    Csr Q;                                                  // optional (I - D_w A) P: prolongation fused with the first post sweep
    bool has_Q = false;
    const void *val_scaled = nullptr;                       // optional values of A D_w (a_ij * dw_j) on A's pattern
    // optional W32 copies (slot-major inside 32-row windows, apply.cu) of the scaled operator and of Q: thread-per-row
    // kernels with a coalesced operator stream — used on levels with enough rows to fill the GPU with one thread per row
    const int *w32_a_col = nullptr, *w32_q_col = nullptr;
    const void *w32_a_val = nullptr, *w32_q_val = nullptr;
    const void *dw = nullptr;
    const int *r_order = nullptr;                           // optional processing order of the rows of R
    const int *sell_ptr = nullptr, *sell_col = nullptr;   // optional SELL-32 copy of A
    const void *sell_val = nullptr;
    bool has_A = false, has_PR = false;
    void *x = nullptr, *tmp = nullptr, *b = nullptr, *r = nullptr;   // owned work vectors
};

}  // namespace mlamg

namespace mlamg {
// Host-buffer apply (mlamg_vcycle_host): the right-hand side arrives and the result leaves in row chunks on a copy
// stream; the first fine-level pass runs chunk by chunk as soon as the columns it gathers have arrived, the last one
// hands every finished chunk to the D2H copy — the two PCIe transfers overlap the two largest kernels.
constexpr int PIPE_CHUNKS = 8;
struct HostPipe {
    bool ready = false;
    cudaStream_t cs = nullptr;
    cudaEvent_t in_ev[PIPE_CHUNKS] = {}, out_ev[PIPE_CHUNKS] = {}, fork = nullptr;
    int row_lo[PIPE_CHUNKS + 1] = {};
    int need[PIPE_CHUNKS] = {};          // chunk whose arrival completes the columns gathered by the rows of chunk k
    int nchunks = 0;
};
}  // namespace mlamg

namespace mlamg { struct SolverState; }

struct mlamg_hierarchy {
    int dtype = MLAMG_F64;
    size_t esz = 8;
    std::vector<mlamg::LevelData> lv;
    const void *coarse_inv = nullptr;
    bool finalized = false;
    bool use_graph = false;
    mlamg::SolverState *solver = nullptr;        // work vectors, device scalars and loop graphs of solvers.cu (lazy)
    void *host_b = nullptr, *host_x = nullptr;   // device staging of the *_host entry points
    mlamg::HostPipe pipe;
    // CUDA graph cache of one V-cycle
    cudaStream_t cap_stream = nullptr;
    cudaGraphExec_t gexec = nullptr;
    const void *g_b = nullptr;
    void *g_x = nullptr;
    int g_nu1 = -1, g_nu2 = -1, g_zero = -1;
};


namespace mlamg {
int check_handle(mlamg_hierarchy_t h);
// enqueue one V(nu1,nu2) cycle on `s` (plain launches; capturable)
int vcycle_dispatch(mlamg_hierarchy *h, const void *b, void *x, int nu1, int nu2, int zero_guess, cudaStream_t s);
// the same through the handle's CUDA-graph cache when use_graph is set
int vcycle_run(mlamg_hierarchy *h, const void *b, void *x, int nu1, int nu2, int zero_guess, cudaStream_t s);
void solver_state_free(mlamg_hierarchy *h);
void solver_graphs_reset(mlamg_hierarchy *h);      // drop the cached PCG / stationary loop graphs (operators changed)
}  // namespace mlamg
