// V-cycle apply kernels: CSR SpMV family (SpMV, residual(+norm), fused Jacobi sweep, prolong-add),
// smoother diagonals, multi-vector SpMM, dense GEMV for the coarsest level.
//
// Roofline: every kernel here is HBM-bound (SpMV intensity ~2 flop / 12 B).  One pass over the
// operator per kernel; x gathers are served by L1/L2 (banded operators) — algorithmic bytes per
// SURVEY.md §8(d):  B_spmv = nnz(v+4) + 4(N+1) + 2vN,  B_jacobi = nnz(v+4) + 4(N+1) + 4vN.
#include "common.cuh"

namespace mlamg {

enum { OP_SPMV = 0, OP_SPMV_ADD = 1, OP_RESIDUAL = 2, OP_JACOBI = 3 };

constexpr int ROW_THREADS = 256;

// LANES threads cooperate on one row (LANES = 2..32, power of two): coalesced reads of col/val across
// the warp because consecutive rows are contiguous in CSR; segmented shuffle reduction.
template <typename T, int LANES, int OP, bool NORM>
__global__ void __launch_bounds__(ROW_THREADS)
csr_rowop_kernel(int n, const int *__restrict__ rowptr, const int *__restrict__ col,
                 const T *__restrict__ val, const T *__restrict__ x, const T *__restrict__ b,
                 const T *__restrict__ dw, T *__restrict__ y, double *__restrict__ partial) {
    const long long gtid = (long long)blockIdx.x * ROW_THREADS + threadIdx.x;
    const long long row = gtid / LANES;
    const int lane = threadIdx.x & (LANES - 1);
    T sum = (T)0;
    if (row < n) {
        const int start = rowptr[row];
        const int end = rowptr[row + 1];
        for (int j = start + lane; j < end; j += LANES) sum += val[j] * x[col[j]];
    }
    sum = group_sum<LANES>(sum);
    double rr = 0.0;
    if (row < n && lane == 0) {
        if (OP == OP_SPMV) {
            y[row] = sum;
        } else if (OP == OP_SPMV_ADD) {
            y[row] += sum;
        } else if (OP == OP_RESIDUAL) {
            const T r = b[row] - sum;
            y[row] = r;
            if (NORM) rr = (double)r * (double)r;
        } else {  // OP_JACOBI
            y[row] = x[row] + dw[row] * (b[row] - sum);
        }
    }
    if (NORM) {
        __shared__ double sm[32];
        rr = block_sum(rr, sm);
        if (threadIdx.x == 0) partial[blockIdx.x] = rr;
    }
}

static int pick_lanes(int n, long long nnz) {
    const double mean = n > 0 ? (double)nnz / (double)n : 0.0;
    if (mean <= 3.0) return 2;
    if (mean <= 6.0) return 4;
    if (mean <= 12.0) return 8;
    if (mean <= 24.0) return 16;
    return 32;
}

template <typename T, int OP, bool NORM>
static int launch_rowop(int n, long long nnz_hint, const int *rowptr, const int *col, const T *val, const T *x,
                        const T *b, const T *dw, T *y, double *norm2, cudaStream_t s) {
    if (n <= 0) {
        if (NORM && norm2) MLAMG_CUDA(cudaMemsetAsync(norm2, 0, sizeof(double), s));
        return MLAMG_OK;
    }
    const int lanes = pick_lanes(n, nnz_hint);
    const unsigned blocks = cdiv((long long)n * lanes, ROW_THREADS);
    double *partial = nullptr;
    Scratch part(NORM ? (size_t)blocks * sizeof(double) : 16, s);
    if (NORM) {
        MLAMG_SCRATCH_OK(part);
        partial = part.as<double>();
    }
#define LAUNCH(L)                                                                                       \
    csr_rowop_kernel<T, L, OP, NORM><<<blocks, ROW_THREADS, 0, s>>>(n, rowptr, col, val, x, b, dw, y, partial)
    switch (lanes) {
        case 2: LAUNCH(2); break;
        case 4: LAUNCH(4); break;
        case 8: LAUNCH(8); break;
        case 16: LAUNCH(16); break;
        default: LAUNCH(32); break;
    }
#undef LAUNCH
    MLAMG_LAUNCHED();
    if (NORM) return reduce_partials(partial, (int)blocks, norm2, s);
    return MLAMG_OK;
}

// ---- internal entry points used by hierarchy.cu (nnz known, no host sync) -------------------
template <typename T>
int spmv_t(int n, long long nnz, const int *rowptr, const int *col, const T *val, const T *x, T *y, cudaStream_t s) {
    return launch_rowop<T, OP_SPMV, false>(n, nnz, rowptr, col, val, x, nullptr, nullptr, y, nullptr, s);
}
template <typename T>
int spmv_add_t(int n, long long nnz, const int *rowptr, const int *col, const T *val, const T *x, T *y,
               cudaStream_t s) {
    return launch_rowop<T, OP_SPMV_ADD, false>(n, nnz, rowptr, col, val, x, nullptr, nullptr, y, nullptr, s);
}
template <typename T>
int residual_t(int n, long long nnz, const int *rowptr, const int *col, const T *val, const T *x, const T *b, T *r,
               double *norm2, cudaStream_t s) {
    if (norm2) return launch_rowop<T, OP_RESIDUAL, true>(n, nnz, rowptr, col, val, x, b, nullptr, r, norm2, s);
    return launch_rowop<T, OP_RESIDUAL, false>(n, nnz, rowptr, col, val, x, b, nullptr, r, nullptr, s);
}
template <typename T>
int jacobi_t(int n, long long nnz, const int *rowptr, const int *col, const T *val, const T *dw, const T *b,
             const T *x_in, T *x_out, cudaStream_t s) {
    return launch_rowop<T, OP_JACOBI, false>(n, nnz, rowptr, col, val, x_in, b, dw, x_out, nullptr, s);
}

template int spmv_t<float>(int, long long, const int *, const int *, const float *, const float *, float *, cudaStream_t);
template int spmv_t<double>(int, long long, const int *, const int *, const double *, const double *, double *, cudaStream_t);
template int spmv_add_t<float>(int, long long, const int *, const int *, const float *, const float *, float *, cudaStream_t);
template int spmv_add_t<double>(int, long long, const int *, const int *, const double *, const double *, double *, cudaStream_t);
template int residual_t<float>(int, long long, const int *, const int *, const float *, const float *, const float *, float *, double *, cudaStream_t);
template int residual_t<double>(int, long long, const int *, const int *, const double *, const double *, const double *, double *, double *, cudaStream_t);
template int jacobi_t<float>(int, long long, const int *, const int *, const float *, const float *, const float *, const float *, float *, cudaStream_t);
template int jacobi_t<double>(int, long long, const int *, const int *, const double *, const double *, const double *, const double *, double *, cudaStream_t);

// ---- elementwise -------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) jacobi_zero_kernel(int n, const T *__restrict__ dw, const T *__restrict__ b,
                                                          T *__restrict__ x) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        x[i] = dw[i] * b[i];
}

template <typename T>
int jacobi_zero_t(int n, const T *dw, const T *b, T *x, cudaStream_t s) {
    if (n <= 0) return MLAMG_OK;
    unsigned blocks = cdiv(n, 256);
    if (blocks > 148u * 16u) blocks = 148u * 16u;
    jacobi_zero_kernel<T><<<blocks, 256, 0, s>>>(n, dw, b, x);
    MLAMG_LAUNCHED();
    return MLAMG_OK;
}
template int jacobi_zero_t<float>(int, const float *, const float *, float *, cudaStream_t);
template int jacobi_zero_t<double>(int, const double *, const double *, double *, cudaStream_t);

// warp per row: dw = omega/a_ii or 1/sum|a_ij|
template <typename T>
__global__ void __launch_bounds__(256) smoother_diag_kernel(int mode, T omega, int n, const int *__restrict__ rowptr,
                                                            const int *__restrict__ col, const T *__restrict__ val,
                                                            T *__restrict__ dw) {
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    T d = (T)0, l1 = (T)0;
    if (row < n) {
        for (int j = rowptr[row] + lane; j < rowptr[row + 1]; j += 32) {
            const T v = val[j];
            if (col[j] == row) d += v;
            l1 += (v < (T)0 ? -v : v);
        }
    }
    d = warp_sum(d);
    l1 = warp_sum(l1);
    if (row < n && lane == 0) dw[row] = (mode == 0) ? omega / d : (T)1 / l1;
}

// ---- SpMM: Y = alpha * A X + beta * Y, X/Y row-major N x k (k small) --------------------------
// one warp per row; lanes span the k columns (k <= 32 per pass), coalesced reads of X rows.
template <typename T>
__global__ void __launch_bounds__(256) spmm_kernel(int n, int k, const int *__restrict__ rowptr,
                                                   const int *__restrict__ col, const T *__restrict__ val,
                                                   const T *__restrict__ X, T *__restrict__ Y, T alpha, T beta) {
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    const int start = rowptr[row], end = rowptr[row + 1];
    for (int c0 = 0; c0 < k; c0 += 32) {
        const int c = c0 + lane;
        T acc = (T)0;
        if (c < k) {
            for (int j = start; j < end; j++) acc += val[j] * X[(long long)col[j] * k + c];
            T *yp = Y + row * k + c;
            *yp = alpha * acc + ((beta == (T)0) ? (T)0 : beta * (*yp));
        }
    }
}

// ---- dense GEMV y = M x (coarsest level, M = explicit inverse), warp per row -------------------
template <typename T>
__global__ void __launch_bounds__(256) gemv_kernel(int n, const T *__restrict__ M, const T *__restrict__ x,
                                                   T *__restrict__ y) {
    const int row = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    T acc = (T)0;
    if (row < n) {
        const T *mr = M + (long long)row * n;
        for (int j = lane; j < n; j += 32) acc += mr[j] * x[j];
    }
    acc = warp_sum(acc);
    if (row < n && lane == 0) y[row] = acc;
}

template <typename T>
int gemv_t(int n, const T *M, const T *x, T *y, cudaStream_t s) {
    if (n <= 0) return MLAMG_OK;
    gemv_kernel<T><<<cdiv((long long)n * 32, 256), 256, 0, s>>>(n, M, x, y);
    MLAMG_LAUNCHED();
    return MLAMG_OK;
}
template int gemv_t<float>(int, const float *, const float *, float *, cudaStream_t);
template int gemv_t<double>(int, const double *, const double *, double *, cudaStream_t);

}  // namespace mlamg

using namespace mlamg;

extern "C" {

int mlamg_spmv_csr(int dtype, int n, int nnz, const int *rowptr, const int *col, const void *val, const void *x, void *y,
                   mlamg_stream_t stream) {
    cudaStream_t s = as_stream(stream);
    if (n < 0) return set_error(MLAMG_EINVAL, "spmv: n < 0");
    if (x == y) return set_error(MLAMG_EINVAL, "spmv: x aliases y");
    MLAMG_DISPATCH(dtype, return spmv_t<T>(n, nnz, rowptr, col, (const T *)val, (const T *)x, (T *)y, s));
    return MLAMG_OK;
}

int mlamg_spmv_add_csr(int dtype, int n, int nnz, const int *rowptr, const int *col, const void *val, const void *x, void *y,
                       mlamg_stream_t stream) {
    cudaStream_t s = as_stream(stream);
    if (n < 0) return set_error(MLAMG_EINVAL, "spmv_add: n < 0");
    if (x == y) return set_error(MLAMG_EINVAL, "spmv_add: x aliases y");
    MLAMG_DISPATCH(dtype, return spmv_add_t<T>(n, nnz, rowptr, col, (const T *)val, (const T *)x, (T *)y, s));
    return MLAMG_OK;
}

int mlamg_residual_csr(int dtype, int n, int nnz, const int *rowptr, const int *col, const void *val, const void *x,
                       const void *b, void *r, double *norm2, mlamg_stream_t stream) {
    cudaStream_t s = as_stream(stream);
    if (n < 0) return set_error(MLAMG_EINVAL, "residual: n < 0");
    if (x == r) return set_error(MLAMG_EINVAL, "residual: x aliases r");
    MLAMG_DISPATCH(dtype, return residual_t<T>(n, nnz, rowptr, col, (const T *)val, (const T *)x, (const T *)b,
                                               (T *)r, norm2, s));
    return MLAMG_OK;
}

int mlamg_jacobi_csr(int dtype, int n, int nnz, const int *rowptr, const int *col, const void *val, const void *dw,
                     const void *b, const void *x_in, void *x_out, mlamg_stream_t stream) {
    cudaStream_t s = as_stream(stream);
    if (n < 0) return set_error(MLAMG_EINVAL, "jacobi: n < 0");
    if (x_in == x_out) return set_error(MLAMG_EINVAL, "jacobi: x_in aliases x_out (ping-pong required)");
    MLAMG_DISPATCH(dtype, return jacobi_t<T>(n, nnz, rowptr, col, (const T *)val, (const T *)dw, (const T *)b,
                                             (const T *)x_in, (T *)x_out, s));
    return MLAMG_OK;
}

int mlamg_jacobi_zero(int dtype, int n, const void *dw, const void *b, void *x, mlamg_stream_t stream) {
    MLAMG_DISPATCH(dtype, return jacobi_zero_t<T>(n, (const T *)dw, (const T *)b, (T *)x, as_stream(stream)));
    return MLAMG_OK;
}

int mlamg_smoother_diag(int dtype, int mode, double omega, int n, const int *rowptr, const int *col,
                        const void *val, void *dw, mlamg_stream_t stream) {
    cudaStream_t s = as_stream(stream);
    if (n < 0 || (mode != 0 && mode != 1)) return set_error(MLAMG_EINVAL, "smoother_diag: bad n/mode");
    if (n == 0) return MLAMG_OK;
    MLAMG_DISPATCH(dtype, (smoother_diag_kernel<T><<<cdiv((long long)n * 32, 256), 256, 0, s>>>(
                              mode, (T)omega, n, rowptr, col, (const T *)val, (T *)dw)));
    MLAMG_LAUNCHED();
    return MLAMG_OK;
}

int mlamg_spmm_csr(int dtype, int n, int k, const int *rowptr, const int *col, const void *val, const void *X,
                   void *Y, double alpha, double beta, mlamg_stream_t stream) {
    cudaStream_t s = as_stream(stream);
    if (n < 0 || k < 0) return set_error(MLAMG_EINVAL, "spmm: bad n/k");
    if (n == 0 || k == 0) return MLAMG_OK;
    if (X == Y) return set_error(MLAMG_EINVAL, "spmm: X aliases Y");
    MLAMG_DISPATCH(dtype, (spmm_kernel<T><<<cdiv((long long)n * 32, 256), 256, 0, s>>>(
                              n, k, rowptr, col, (const T *)val, (const T *)X, (T *)Y, (T)alpha, (T)beta)));
    MLAMG_LAUNCHED();
    return MLAMG_OK;
}

int mlamg_gemv(int dtype, int n, const void *m, const void *x, void *y, mlamg_stream_t stream) {
    MLAMG_DISPATCH(dtype, return gemv_t<T>(n, (const T *)m, (const T *)x, (T *)y, as_stream(stream)));
    return MLAMG_OK;
}

}  // extern "C"
