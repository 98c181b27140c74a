// Backward-pass kernels of the multi-vector two-grid loss (ns/model/loss.py:32-96) and of the learned
// prolongator P = P_hat Agg (ns/model/agg_interp.py:481-484).  The reference gets these gradients from
// torch_sparse's autograd (spmm / spspmm backward); here they are three small kernels on the CSR pattern:
//   SDDMM       g_S[j]      = <U[row(j), :], V[col(j), :]>          (gradient of S in Y = S X and Y = S^T X)
//   sample      g_S[j]      = D[row(j), col(j)]                     (gradient of P in P^T A P: D = A P G^T + A^T P G)
//   agg product g_Phat[j]   = g_P[row(j), labels[col(j)]]           (gradient of P_hat in P = P_hat Agg)
// All three are HBM-bound gathers; a warp owns a row, reductions are fixed-order warp shuffles (deterministic).
#include "common.cuh"

namespace mlamg {

template <typename T>
__global__ void __launch_bounds__(256) sddmm_kernel(int n, int k, const int *__restrict__ rowptr,
                                                    const int *__restrict__ col, const T *__restrict__ U,
                                                    const T *__restrict__ V, T *__restrict__ out) {
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;     // warp-uniform
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    const int start = rowptr[row], end = rowptr[row + 1];
    const T *u = U + row * k;
    for (int j = start; j < end; j++) {
        const T *v = V + (long long)col[j] * k;
        T acc = (T)0;
        for (int c = lane; c < k; c += 32) acc += u[c] * v[c];
        acc = warp_sum(acc);
        if (lane == 0) out[j] = acc;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) sample_dense_kernel(int n, int ncols, const int *__restrict__ rowptr,
                                                           const int *__restrict__ col, const T *__restrict__ D,
                                                           T *__restrict__ out) {
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    const int start = rowptr[row], end = rowptr[row + 1];
    for (int j = start + lane; j < end; j += 32) out[j] = D[row * ncols + col[j]];
}

template <typename T>
__global__ void __launch_bounds__(256) agg_product_backward_kernel(int n, const int *__restrict__ a_rowptr,
                                                                   const int *__restrict__ a_col,
                                                                   const int *__restrict__ labels,
                                                                   const int *__restrict__ p_rowptr,
                                                                   const int *__restrict__ p_col,
                                                                   const T *__restrict__ g_p, T *__restrict__ g_phat) {
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    const int start = a_rowptr[row], end = a_rowptr[row + 1];
    const int ps = p_rowptr[row], pe = p_rowptr[row + 1];
    for (int j = start + lane; j < end; j += 32) {
        const int c = labels[a_col[j]];
        T g = (T)0;
        if (c >= 0)                                   // rows of P are short (one entry per neighbouring aggregate)
            for (int q = ps; q < pe; q++)
                if (p_col[q] == c) { g = g_p[q]; break; }
        g_phat[j] = g;
    }
}

}  // namespace mlamg

using namespace mlamg;

extern "C" {

int mlamg_sddmm_csr(int dtype, int n, int k, const int *rowptr, const int *col, const void *U, const void *V,
                    void *out, mlamg_stream_t stream) {
    cudaStream_t s = as_stream(stream);
    if (n < 0 || k < 0) return set_error(MLAMG_EINVAL, "sddmm: bad n/k");
    if (n == 0) return MLAMG_OK;
    MLAMG_DISPATCH(dtype, (sddmm_kernel<T><<<cdiv((long long)n * 32, 256), 256, 0, s>>>(
                              n, k, rowptr, col, (const T *)U, (const T *)V, (T *)out)));
    MLAMG_LAUNCHED();
    return MLAMG_OK;
}

int mlamg_csr_sample_dense(int dtype, int n, int ncols, const int *rowptr, const int *col, const void *dense,
                           void *out, mlamg_stream_t stream) {
    cudaStream_t s = as_stream(stream);
    if (n < 0 || ncols < 0) return set_error(MLAMG_EINVAL, "sample_dense: bad n/ncols");
    if (n == 0) return MLAMG_OK;
    MLAMG_DISPATCH(dtype, (sample_dense_kernel<T><<<cdiv((long long)n * 32, 256), 256, 0, s>>>(
                              n, ncols, rowptr, col, (const T *)dense, (T *)out)));
    MLAMG_LAUNCHED();
    return MLAMG_OK;
}

int mlamg_agg_product_backward(int dtype, int n, const int *a_rowptr, const int *a_col, const int *labels,
                               const int *p_rowptr, const int *p_col, const void *g_p, void *g_phat,
                               mlamg_stream_t stream) {
    cudaStream_t s = as_stream(stream);
    if (n < 0) return set_error(MLAMG_EINVAL, "agg_product_backward: bad n");
    if (n == 0) return MLAMG_OK;
    MLAMG_DISPATCH(dtype, (agg_product_backward_kernel<T><<<cdiv((long long)n * 32, 256), 256, 0, s>>>(
                              n, a_rowptr, a_col, labels, p_rowptr, p_col, (const T *)g_p, (T *)g_phat)));
    MLAMG_LAUNCHED();
    return MLAMG_OK;
}

}  // extern "C"
