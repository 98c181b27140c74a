// Dense inverse of the coarsest operator (LU with partial pivoting through cuSOLVER getrf/getrs).
// This is the one library call on the path: it replaces SuperLU (`spla.factorized`,
// ns/lib/multigrid.py:168; `splu`, ns/preconditioner/MLAMG.py:122) for the bottom level only and
// runs once per setup; the per-cycle coarse solve is the hand-written GEMV in apply.cu.
#include <cusolverDn.h>
#include "common.cuh"

namespace mlamg {

__global__ void __launch_bounds__(256) identity_kernel(int n, double *__restrict__ m) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (long long)n * n) m[i] = (i / n == i % n) ? 1.0 : 0.0;
}

}  // namespace mlamg

using namespace mlamg;

extern "C" int mlamg_dense_inverse_f64(int n, double *a, double *work, mlamg_stream_t stream) {
    cudaStream_t s = as_stream(stream);
    if (n <= 0) return set_error(MLAMG_EINVAL, "dense_inverse: n <= 0");
    // one cuSOLVER handle per (thread, device), created on first use: cusolverDnCreate/Destroy cost 30-300 ms per call
    static thread_local cusolverDnHandle_t g_hd = nullptr;
    static thread_local int g_hd_device = -1;
    int dev_now = 0;
    MLAMG_CUDA(cudaGetDevice(&dev_now));
    if (g_hd == nullptr || g_hd_device != dev_now) {
        g_hd = nullptr;
        if (cusolverDnCreate(&g_hd) != CUSOLVER_STATUS_SUCCESS) return set_error(MLAMG_ECUDA, "cusolverDnCreate failed");
        g_hd_device = dev_now;
    }
    cusolverDnHandle_t hd = g_hd;
    int rc = MLAMG_OK;
    int lwork = 0;
    double *buf = nullptr;
    int *ipiv = nullptr, *info = nullptr;
    int h_info[2] = {0, 0};
    do {
        if (cusolverDnSetStream(hd, s) != CUSOLVER_STATUS_SUCCESS) { rc = set_error(MLAMG_ECUDA, "cusolverDnSetStream failed"); break; }
        if (cusolverDnDgetrf_bufferSize(hd, n, n, a, n, &lwork) != CUSOLVER_STATUS_SUCCESS) { rc = set_error(MLAMG_ECUDA, "getrf_bufferSize failed"); break; }
        if (cudaMalloc(&buf, (size_t)(lwork > 0 ? lwork : 1) * sizeof(double)) != cudaSuccess ||
            cudaMalloc(&ipiv, (size_t)n * sizeof(int)) != cudaSuccess || cudaMalloc(&info, 2 * sizeof(int)) != cudaSuccess) {
            rc = set_error(MLAMG_ECUDA, "dense_inverse: out of device memory");
            break;
        }
        // The row-major matrix is the column-major transpose; inverting the transpose in column-major
        // storage yields the row-major inverse, so no explicit transposition is needed.
        if (cusolverDnDgetrf(hd, n, n, a, n, buf, ipiv, info) != CUSOLVER_STATUS_SUCCESS) { rc = set_error(MLAMG_ECUDA, "getrf failed"); break; }
        identity_kernel<<<cdiv((long long)n * n, 256), 256, 0, s>>>(n, work);
        count_launch();
        if (cusolverDnDgetrs(hd, CUBLAS_OP_N, n, n, a, n, ipiv, work, n, info + 1) != CUSOLVER_STATUS_SUCCESS) { rc = set_error(MLAMG_ECUDA, "getrs failed"); break; }
        if (cudaMemcpyAsync(h_info, info, 2 * sizeof(int), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
            cudaMemcpyAsync(a, work, (size_t)n * n * sizeof(double), cudaMemcpyDeviceToDevice, s) != cudaSuccess ||
            cudaStreamSynchronize(s) != cudaSuccess) {
            rc = set_cuda_error(cudaGetLastError(), __FILE__, __LINE__);
            break;
        }
        if (h_info[0] > 0) rc = set_error(MLAMG_ESINGULAR, "coarse operator is singular (zero pivot %d)", h_info[0]);
        else if (h_info[0] < 0 || h_info[1] != 0) rc = set_error(MLAMG_EINVAL, "getrf/getrs info %d/%d", h_info[0], h_info[1]);
    } while (0);
    if (buf) cudaFree(buf);
    if (ipiv) cudaFree(ipiv);
    if (info) cudaFree(info);
    return rc;
}
