// Exact forward Gauss-Seidel sweep (pyamg `gauss_seidel`, the smoother the reference's amg_2_v uses:
// ns/lib/multigrid.py:175,184), level-scheduled.  Row i depends on rows j < i of its pattern; rows of
// one dependency level are independent.  Upper neighbours (j > i) are read from a snapshot taken
// before the sweep, so the schedule needs only the true (lower-triangular) dependencies and is exact
// for non-symmetric patterns too.  Each row is summed by ONE thread in CSR order with explicit
// non-fused multiply/add, so the result is bit-identical to the sequential C loop.
#include <vector>
#include "common.cuh"

namespace mlamg {

__global__ void __launch_bounds__(256) gs_level_relax_kernel(int n, const int *__restrict__ rowptr,
                                                             const int *__restrict__ col, int *level,
                                                             int *__restrict__ changed) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int own = __ldcg(&level[i]);
    int m = own;
    for (int jj = rowptr[i]; jj < rowptr[i + 1]; jj++) {
        const int j = col[jj];
        if (j < i) {
            const int l = __ldcg(&level[j]) + 1;
            if (l > m) m = l;
        }
    }
    if (m > own) { level[i] = m; *changed = 1; }
}

__global__ void __launch_bounds__(256) gs_level_hist_kernel(int n, const int *__restrict__ level, int *__restrict__ hist,
                                                            int *__restrict__ maxlevel) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int l = level[i];
    atomicAdd(&hist[l], 1);
    atomicMax(maxlevel, l);
}

__global__ void __launch_bounds__(256) gs_level_fill_kernel(int n, const int *__restrict__ level, int *__restrict__ cursor,
                                                            int *__restrict__ order) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    order[atomicAdd(&cursor[level[i]], 1)] = (int)i;
}


template <typename T>
__global__ void __launch_bounds__(128) gs_rows_kernel(int count, const int *__restrict__ rows,
                                                      const int *__restrict__ rowptr, const int *__restrict__ col,
                                                      const T *__restrict__ val, const T *__restrict__ b,
                                                      const T *__restrict__ xold, T *x) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const int i = rows[t];
    T rsum = (T)0, diag = (T)0;
    for (int jj = rowptr[i]; jj < rowptr[i + 1]; jj++) {
        const int j = col[jj];
        const T a = val[jj];
        if (j == i) diag = a;
        else rsum = add_rn(rsum, mul_rn(a, j < i ? __ldcg(&x[j]) : xold[j]));
    }
    if (diag != (T)0) x[i] = div_rn(add_rn(b[i], -rsum), diag);
}

}  // namespace mlamg

using namespace mlamg;

extern "C" {

int mlamg_gs_schedule(int n, const int *rowptr, const int *col, int *level, int *order, int *level_ptr_host,
                      int *nlevels_host, mlamg_stream_t stream) {
    cudaStream_t s = as_stream(stream);
    if (n <= 0 || !level_ptr_host || !nlevels_host) return set_error(MLAMG_EINVAL, "gs_schedule: bad arguments");
    Scratch flag(2 * sizeof(int), s), hist((size_t)(n + 1) * sizeof(int), s);
    MLAMG_SCRATCH_OK(flag);
    MLAMG_SCRATCH_OK(hist);
    MLAMG_CUDA(cudaMemsetAsync(level, 0, (size_t)n * sizeof(int), s));
    const unsigned eb = cdiv(n, 256);
    int h[2] = {1, 0};
    while (h[0]) {
        MLAMG_CUDA(cudaMemsetAsync(flag.p, 0, 2 * sizeof(int), s));
        // a few passes per host check: the dependency depth of a grid operator is O(side length)
        for (int rep = 0; rep < 8; rep++) {
            gs_level_relax_kernel<<<eb, 256, 0, s>>>(n, rowptr, col, level, flag.as<int>());
            MLAMG_LAUNCHED();
        }
        MLAMG_CUDA(cudaMemcpyAsync(h, flag.p, sizeof(int), cudaMemcpyDeviceToHost, s));
        MLAMG_CUDA(cudaStreamSynchronize(s));
    }
    MLAMG_CUDA(cudaMemsetAsync(hist.p, 0, (size_t)(n + 1) * sizeof(int), s));
    MLAMG_CUDA(cudaMemsetAsync(flag.p, 0, 2 * sizeof(int), s));
    gs_level_hist_kernel<<<eb, 256, 0, s>>>(n, level, hist.as<int>(), flag.as<int>() + 1);
    MLAMG_LAUNCHED();
    MLAMG_CUDA(cudaMemcpyAsync(h, flag.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
    MLAMG_CUDA(cudaStreamSynchronize(s));
    const int nlev = h[1] + 1;
    MLAMG_TRY(exclusive_scan_i32(hist.as<int>(), hist.as<int>(), nlev, s));
    MLAMG_CUDA(cudaMemcpyAsync(level_ptr_host, hist.p, (size_t)(nlev + 1) * sizeof(int), cudaMemcpyDeviceToHost, s));
    gs_level_fill_kernel<<<eb, 256, 0, s>>>(n, level, hist.as<int>(), order);
    MLAMG_LAUNCHED();
    MLAMG_CUDA(cudaStreamSynchronize(s));
    *nlevels_host = nlev;
    return MLAMG_OK;
}

int mlamg_gauss_seidel(int dtype, int n, const int *rowptr, const int *col, const void *val, const void *b, void *x,
                       const int *order, const int *level_ptr_host, int nlevels, mlamg_stream_t stream) {
    cudaStream_t s = as_stream(stream);
    if (n <= 0 || nlevels <= 0) return set_error(MLAMG_EINVAL, "gauss_seidel: bad n/nlevels");
    const size_t esz = dtype == MLAMG_F32 ? 4 : 8;
    Scratch xold((size_t)n * esz, s);
    MLAMG_SCRATCH_OK(xold);
    MLAMG_CUDA(cudaMemcpyAsync(xold.p, x, (size_t)n * esz, cudaMemcpyDeviceToDevice, s));
    for (int l = 0; l < nlevels; l++) {
        const int start = level_ptr_host[l], count = level_ptr_host[l + 1] - start;
        if (count <= 0) continue;
        MLAMG_DISPATCH(dtype, (gs_rows_kernel<T><<<cdiv(count, 128), 128, 0, s>>>(
                                  count, order + start, rowptr, col, (const T *)val, (const T *)b,
                                  (const T *)xold.p, (T *)x)));
        MLAMG_LAUNCHED();
    }
    return MLAMG_OK;
}

}  // extern "C"
