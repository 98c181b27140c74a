// Hierarchy-setup kernels other than the SpGEMM: row sort, transpose, exact-zero compaction,
// Agg from labels, centre-rank labels, SA smoother values, CSR->dense, Poisson stencil generator,
// power-iteration estimate of lambda_max(D^-1 A).
#include "common.cuh"

namespace mlamg {

// ------------------------------------------------------------------ row sort (by column index)
// Rank sort: keys inside a row are distinct after SpGEMM / transpose, duplicates (user input) are
// ordered by their original position, so the result is deterministic.
constexpr int SORT_SMEM_A = 2048;    // CTA bin A: rows up to 2048 entries (256 threads)
constexpr int SORT_SMEM_B = 16384;   // CTA bin B: rows up to 16384 entries (1024 threads, 192 KB)

__global__ void __launch_bounds__(256) sort_binid_kernel(int m, const int *__restrict__ rowptr,
                                                         int *__restrict__ binid) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int len = rowptr[i + 1] - rowptr[i];
    int b;
    if (len <= 1) b = -1;
    else if (len <= 32) b = 0;
    else if (len <= SORT_SMEM_A) b = 1;
    else if (len <= SORT_SMEM_B) b = 2;
    else b = 3;
    binid[i] = b;
}

template <typename T>
__global__ void __launch_bounds__(256) sort_warp_kernel(int nrows, const int *__restrict__ rows,
                                                        const int *__restrict__ rowptr, int *__restrict__ col,
                                                        T *__restrict__ val) {
    const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= nrows) return;
    const int row = rows[r];
    const int start = rowptr[row], len = rowptr[row + 1] - start;
    int key = 0x7fffffff;
    T v = (T)0;
    if (lane < len) { key = col[start + lane]; v = val[start + lane]; }
    int rank = 0;
    for (int f = 0; f < len; f++) {
        const int kf = __shfl_sync(0xffffffffu, key, f);
        rank += (kf < key) || (kf == key && f < lane);
    }
    __syncwarp();
    if (lane < len) { col[start + rank] = key; val[start + rank] = v; }
}

template <typename T>
__global__ void sort_cta_kernel(int nrows, const int *__restrict__ rows, int cap, const int *__restrict__ rowptr,
                                int *__restrict__ col, T *__restrict__ val) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int *keys = reinterpret_cast<int *>(smem_raw);
    T *vals = reinterpret_cast<T *>(smem_raw + (size_t)cap * sizeof(int));
    const int row = rows[blockIdx.x];
    const int start = rowptr[row], len = rowptr[row + 1] - start;
    for (int t = threadIdx.x; t < len; t += blockDim.x) { keys[t] = col[start + t]; vals[t] = val[start + t]; }
    __syncthreads();
    for (int t = threadIdx.x; t < len; t += blockDim.x) {
        const int key = keys[t];
        int rank = 0;
        for (int f = 0; f < len; f++) {
            const int kf = keys[f];
            rank += (kf < key) || (kf == key && f < t);
        }
        col[start + rank] = key;
        val[start + rank] = vals[t];
    }
}

template <typename T>
static int sort_rows_t(int m, const int *rowptr, int *col, T *val, cudaStream_t s) {
    if (m <= 0) return MLAMG_OK;
    Scratch ids((size_t)m * sizeof(int), s), rows((size_t)m * sizeof(int), s);
    MLAMG_SCRATCH_OK(ids);
    MLAMG_SCRATCH_OK(rows);
    sort_binid_kernel<<<cdiv(m, 256), 256, 0, s>>>(m, rowptr, ids.as<int>());
    MLAMG_LAUNCHED();
    Bins bins;
    MLAMG_TRY(partition_rows_by_bin(m, ids.as<int>(), rows.as<int>(), &bins, s));
    if (bins.counts[3] > 0) return set_error(MLAMG_ELIMIT, "sort_rows: %d rows longer than %d", bins.counts[3], SORT_SMEM_B);
    const int *rl = rows.as<int>();
    if (bins.counts[0] > 0) {
        sort_warp_kernel<T><<<cdiv((long long)bins.counts[0] * 32, 256), 256, 0, s>>>(bins.counts[0], rl + bins.offsets[0],
                                                                                   rowptr, col, val);
        MLAMG_LAUNCHED();
    }
    if (bins.counts[1] > 0) {
        const size_t smem = (size_t)SORT_SMEM_A * (sizeof(int) + sizeof(T));
        sort_cta_kernel<T><<<bins.counts[1], 256, smem, s>>>(bins.counts[1], rl + bins.offsets[1], SORT_SMEM_A, rowptr,
                                                            col, val);
        MLAMG_LAUNCHED();
    }
    if (bins.counts[2] > 0) {
        const size_t smem = (size_t)SORT_SMEM_B * (sizeof(int) + sizeof(T));
        MLAMG_CUDA(cudaFuncSetAttribute(sort_cta_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sort_cta_kernel<T><<<bins.counts[2], 1024, smem, s>>>(bins.counts[2], rl + bins.offsets[2], SORT_SMEM_B, rowptr,
                                                             col, val);
        MLAMG_LAUNCHED();
    }
    return MLAMG_OK;
}

int sort_rows_impl(int dtype, int m, const int *rowptr, int *col, void *val, cudaStream_t s) {
    MLAMG_DISPATCH(dtype, return sort_rows_t<T>(m, rowptr, col, (T *)val, s));
    return MLAMG_OK;
}

// ------------------------------------------------------------------ transpose
__global__ void __launch_bounds__(256) col_count_kernel(int nnz, const int *__restrict__ col, int *__restrict__ cnt) {
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < nnz;
         j += (long long)gridDim.x * blockDim.x)
        atomicAdd(&cnt[col[j]], 1);
}

// LANES-free: warp per source row so the row id is known without a search
template <typename T>
__global__ void __launch_bounds__(256) transpose_fill_kernel(int m, const int *__restrict__ rowptr,
                                                             const int *__restrict__ col, const T *__restrict__ val,
                                                             int *__restrict__ cursor, int *__restrict__ t_col,
                                                             T *__restrict__ t_val) {
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;  // 8 lanes per row
    const int lane = threadIdx.x & 7;
    if (row >= m) return;
    for (int j = rowptr[row] + lane; j < rowptr[row + 1]; j += 8) {
        const int pos = atomicAdd(&cursor[col[j]], 1);
        t_col[pos] = (int)row;
        t_val[pos] = val[j];
    }
}

template <typename T>
static int transpose_t(int m, int n, int nnz, const int *rowptr, const int *col, const T *val, int *t_rowptr,
                       int *t_col, T *t_val, cudaStream_t s) {
    if (m < 0 || n < 0 || nnz < 0) return set_error(MLAMG_EINVAL, "transpose: negative size");
    MLAMG_CUDA(cudaMemsetAsync(t_rowptr, 0, (size_t)(n + 1) * sizeof(int), s));
    if (nnz > 0) {
        unsigned blocks = cdiv(nnz, 256);
        if (blocks > 148u * 32u) blocks = 148u * 32u;
        col_count_kernel<<<blocks, 256, 0, s>>>(nnz, col, t_rowptr);
        MLAMG_LAUNCHED();
    }
    MLAMG_TRY(exclusive_scan_i32(t_rowptr, t_rowptr, n, s));
    if (nnz == 0 || m == 0) return MLAMG_OK;
    Scratch cur((size_t)(n > 0 ? n : 1) * sizeof(int), s);
    MLAMG_SCRATCH_OK(cur);
    MLAMG_CUDA(cudaMemcpyAsync(cur.p, t_rowptr, (size_t)n * sizeof(int), cudaMemcpyDeviceToDevice, s));
    transpose_fill_kernel<T><<<cdiv((long long)m * 8, 256), 256, 0, s>>>(m, rowptr, col, val, cur.as<int>(), t_col, t_val);
    MLAMG_LAUNCHED();
    return sort_rows_t<T>(n, t_rowptr, t_col, t_val, s);
}

// ------------------------------------------------------------------ exact-zero compaction
template <typename T>
__global__ void __launch_bounds__(256) nonzero_count_kernel(int m, const int *__restrict__ rowptr,
                                                            const T *__restrict__ val, int *__restrict__ cnt) {
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const int lane = threadIdx.x & 7;
    int c = 0;
    if (row < m)
        for (int j = rowptr[row] + lane; j < rowptr[row + 1]; j += 8) c += (val[j] != (T)0);
    c = group_sum<8>(c);
    if (row < m && lane == 0) cnt[row] = c;
}

// one thread per row keeps the surviving entries in their original (sorted) order
template <typename T>
__global__ void __launch_bounds__(256) nonzero_fill_kernel(int m, const int *__restrict__ rowptr,
                                                           const int *__restrict__ col, const T *__restrict__ val,
                                                           const int *__restrict__ new_rowptr, int *__restrict__ new_col,
                                                           T *__restrict__ new_val) {
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= m) return;
    int o = new_rowptr[row];
    for (int j = rowptr[row]; j < rowptr[row + 1]; j++) {
        const T v = val[j];
        if (v != (T)0) { new_col[o] = col[j]; new_val[o] = v; o++; }
    }
}

// ------------------------------------------------------------------ Agg / labels
__global__ void __launch_bounds__(256) label_flag_kernel(int n, const int *__restrict__ labels, int *__restrict__ flag) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = labels[i] >= 0 ? 1 : 0;
}

template <typename T>
__global__ void __launch_bounds__(256) agg_fill_kernel(int n, const int *__restrict__ labels,
                                                       const int *__restrict__ rowptr, int *__restrict__ col,
                                                       T *__restrict__ val) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int l = labels[i];
    if (l >= 0) { const int p = rowptr[i]; col[p] = l; val[p] = (T)1; }
}

__global__ void __launch_bounds__(256) fill_i32_kernel(int n, int v, int *__restrict__ a) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = v;
}

__global__ void __launch_bounds__(256) center_map_kernel(int k, const int *__restrict__ centers, int *__restrict__ map) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < k) map[centers[i]] = (int)i;   // duplicate centres: last writer wins, like the dict at graph.py:76-78
}

__global__ void __launch_bounds__(256) center_rank_kernel(int n, const int *__restrict__ nearest,
                                                          const int *__restrict__ map, int *__restrict__ labels,
                                                          int *__restrict__ bad) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = nearest[i];
    int l = -1;
    if (c >= 0 && c < n) l = map[c];
    if (l < 0) atomicExch(bad, 1);
    labels[i] = l;
}

// ------------------------------------------------------------------ S = I - omega D^-1 A
// warp per row; values follow scipy's expression order at multigrid.py:104-106:
//   t = (omega * (1/a_ii)) * a_ij ;  s_ij = (i==j) ? 1 - t : -t
// and the row is stored as scipy stores `eye - omega*Dinv@A`: off-diagonals in A's order, diagonal last.
template <typename T>
__global__ void __launch_bounds__(256) sa_smoother_kernel(int n, const int *__restrict__ rowptr,
                                                          const int *__restrict__ col, const T *__restrict__ val,
                                                          T omega, int *__restrict__ scol, T *__restrict__ sval,
                                                          int *__restrict__ bad) {
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    const int start = rowptr[row], end = rowptr[row + 1];
    T d = (T)0;
    int pd = 0x7fffffff;
    for (int j = start + lane; j < end; j += 32)
        if (col[j] == row) { d = val[j]; pd = min(pd, j); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pd = min(pd, __shfl_xor_sync(0xffffffffu, pd, o));
    if (pd == 0x7fffffff) {
        if (lane == 0) atomicExch(bad, 1);
        return;
    }
    d = __shfl_sync(0xffffffffu, d, (pd - start) & 31);   // the lane that read position pd
    const T w = mul_rn(omega, div_rn((T)1, d));
    for (int j = start + lane; j < end; j += 32) {
        const T t = mul_rn(w, val[j]);
        const int c = col[j];
        if (j == pd) { scol[end - 1] = c; sval[end - 1] = add_rn((T)1, -t); }
        else { const int q = j < pd ? j : j - 1; scol[q] = c; sval[q] = -t; }
    }
}

// ------------------------------------------------------------------ CSR -> dense
template <typename T>
__global__ void __launch_bounds__(256) csr_to_dense_kernel(int n, const int *__restrict__ rowptr,
                                                           const int *__restrict__ col, const T *__restrict__ val,
                                                           T *__restrict__ dense) {
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    for (int j = rowptr[row] + lane; j < rowptr[row + 1]; j += 32)
        atomicAdd(&dense[row * n + col[j]], val[j]);   // duplicates (if any) are summed
}

// ------------------------------------------------------------------ Poisson stencil
__device__ __forceinline__ int stencil_count(int x, int y, int z, int nx, int ny, int nz) {
    return 1 + (x > 0) + (x < nx - 1) + (y > 0) + (y < ny - 1) + (z > 0) + (z < nz - 1);
}

// rows of the z-slab [z0, z0+nzl) of the global nx x ny x nz grid; i is the LOCAL row index
__global__ void __launch_bounds__(256) poisson_count_kernel(int nx, int ny, int nz, int z0, int nzl, int *__restrict__ cnt) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long N = (long long)nx * ny * nzl;
    if (i >= N) return;
    const int x = (int)(i % nx), y = (int)((i / nx) % ny), z = z0 + (int)(i / ((long long)nx * ny));
    cnt[i] = stencil_count(x, y, z, nx, ny, nz);
}

template <typename T>
__global__ void __launch_bounds__(256) poisson_fill_kernel(int nx, int ny, int nz, int z0, int nzl,
                                                           const int *__restrict__ rowptr, int *__restrict__ col,
                                                           T *__restrict__ val) {
    const long long il = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long N = (long long)nx * ny * nzl;
    if (il >= N) return;
    const long long sxy = (long long)nx * ny;
    const long long i = il + (long long)z0 * sxy;     // GLOBAL row = global column of the diagonal
    const int x = (int)(i % nx), y = (int)((i / nx) % ny), z = (int)(i / sxy);
    const T diag = (T)(2 * ((nx > 1) + (ny > 1) + (nz > 1)));
    int p = rowptr[il];
    if (z > 0) { col[p] = (int)(i - sxy); val[p++] = (T)-1; }
    if (y > 0) { col[p] = (int)(i - nx); val[p++] = (T)-1; }
    if (x > 0) { col[p] = (int)(i - 1); val[p++] = (T)-1; }
    col[p] = (int)i; val[p++] = diag;
    if (x < nx - 1) { col[p] = (int)(i + 1); val[p++] = (T)-1; }
    if (y < ny - 1) { col[p] = (int)(i + nx); val[p++] = (T)-1; }
    if (z < nz - 1) { col[p] = (int)(i + sxy); val[p++] = (T)-1; }
}

}  // namespace mlamg

using namespace mlamg;

extern "C" {

int mlamg_csr_sort_rows(int dtype, int m, const int *rowptr, int *col, void *val, mlamg_stream_t stream) {
    return sort_rows_impl(dtype, m, rowptr, col, val, as_stream(stream));
}

int mlamg_csr_transpose(int dtype, int m, int n, int nnz, const int *rowptr, const int *col, const void *val,
                        int *t_rowptr, int *t_col, void *t_val, mlamg_stream_t stream) {
    MLAMG_DISPATCH(dtype, return transpose_t<T>(m, n, nnz, rowptr, col, (const T *)val, t_rowptr, t_col, (T *)t_val,
                                                as_stream(stream)));
    return MLAMG_OK;
}

int mlamg_csr_nonzero_count(int dtype, int m, const int *rowptr, const void *val, int *new_rowptr,
                            long long *nnz_host, mlamg_stream_t stream) {
    cudaStream_t s = as_stream(stream);
    if (m < 0) return set_error(MLAMG_EINVAL, "nonzero_count: m < 0");
    if (m > 0) {
        MLAMG_DISPATCH(dtype, (nonzero_count_kernel<T><<<cdiv((long long)m * 8, 256), 256, 0, s>>>(
                                  m, rowptr, (const T *)val, new_rowptr)));
        MLAMG_LAUNCHED();
    }
    MLAMG_TRY(exclusive_scan_i32(new_rowptr, new_rowptr, m, s));
    int h = 0;
    MLAMG_CUDA(cudaMemcpyAsync(&h, new_rowptr + m, sizeof(int), cudaMemcpyDeviceToHost, s));
    MLAMG_CUDA(cudaStreamSynchronize(s));
    if (nnz_host) *nnz_host = h;
    return MLAMG_OK;
}

int mlamg_csr_nonzero_fill(int dtype, int m, const int *rowptr, const int *col, const void *val,
                           const int *new_rowptr, int *new_col, void *new_val, mlamg_stream_t stream) {
    cudaStream_t s = as_stream(stream);
    if (m <= 0) return m == 0 ? MLAMG_OK : set_error(MLAMG_EINVAL, "nonzero_fill: m < 0");
    MLAMG_DISPATCH(dtype, (nonzero_fill_kernel<T><<<cdiv(m, 256), 256, 0, s>>>(m, rowptr, col, (const T *)val, new_rowptr,
                                                                            new_col, (T *)new_val)));
    MLAMG_LAUNCHED();
    return MLAMG_OK;
}

int mlamg_agg_from_labels(int dtype, int n, const int *labels, int *rowptr, int *col, void *val, int *nnz_host,
                          mlamg_stream_t stream) {
    cudaStream_t s = as_stream(stream);
    if (n < 0) return set_error(MLAMG_EINVAL, "agg_from_labels: n < 0");
    if (n > 0) {
        label_flag_kernel<<<cdiv(n, 256), 256, 0, s>>>(n, labels, rowptr);
        MLAMG_LAUNCHED();
    }
    MLAMG_TRY(exclusive_scan_i32(rowptr, rowptr, n, s));
    if (n > 0) {
        MLAMG_DISPATCH(dtype, (agg_fill_kernel<T><<<cdiv(n, 256), 256, 0, s>>>(n, labels, rowptr, col, (T *)val)));
        MLAMG_LAUNCHED();
    }
    int h = 0;
    MLAMG_CUDA(cudaMemcpyAsync(&h, rowptr + n, sizeof(int), cudaMemcpyDeviceToHost, s));
    MLAMG_CUDA(cudaStreamSynchronize(s));
    if (nnz_host) *nnz_host = h;
    return MLAMG_OK;
}

int mlamg_center_rank_labels(int n, int k, const int *centers, const int *nearest, int *scratch_map, int *labels,
                             mlamg_stream_t stream) {
    cudaStream_t s = as_stream(stream);
    if (n < 0 || k < 0) return set_error(MLAMG_EINVAL, "center_rank_labels: bad n/k");
    if (n == 0) return MLAMG_OK;
    Scratch bad(sizeof(int), s);
    MLAMG_SCRATCH_OK(bad);
    MLAMG_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), s));
    fill_i32_kernel<<<cdiv(n, 256), 256, 0, s>>>(n, -1, scratch_map);
    MLAMG_LAUNCHED();
    if (k > 0) {
        center_map_kernel<<<cdiv(k, 256), 256, 0, s>>>(k, centers, scratch_map);
        MLAMG_LAUNCHED();
    }
    center_rank_kernel<<<cdiv(n, 256), 256, 0, s>>>(n, nearest, scratch_map, labels, bad.as<int>());
    MLAMG_LAUNCHED();
    int h = 0;
    MLAMG_CUDA(cudaMemcpyAsync(&h, bad.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    MLAMG_CUDA(cudaStreamSynchronize(s));
    if (h) return set_error(MLAMG_EKEY, "nearest_center holds a node that is not a centre (unreachable node?)");
    return MLAMG_OK;
}

int mlamg_sa_smoother(int dtype, int n, const int *rowptr, const int *col, const void *val, double omega, int *scol,
                      void *sval, mlamg_stream_t stream) {
    cudaStream_t s = as_stream(stream);
    if (n < 0) return set_error(MLAMG_EINVAL, "sa_smoother: n < 0");
    if (n == 0) return MLAMG_OK;
    Scratch bad(sizeof(int), s);
    MLAMG_SCRATCH_OK(bad);
    MLAMG_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), s));
    MLAMG_DISPATCH(dtype, (sa_smoother_kernel<T><<<cdiv((long long)n * 32, 256), 256, 0, s>>>(
                              n, rowptr, col, (const T *)val, (T)omega, scol, (T *)sval, bad.as<int>())));
    MLAMG_LAUNCHED();
    int h = 0;
    MLAMG_CUDA(cudaMemcpyAsync(&h, bad.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    MLAMG_CUDA(cudaStreamSynchronize(s));
    if (h) return set_error(MLAMG_EINVAL, "sa_smoother: a row stores no diagonal entry");
    return MLAMG_OK;
}

int mlamg_csr_to_dense(int dtype, int n, const int *rowptr, const int *col, const void *val, void *dense,
                       mlamg_stream_t stream) {
    cudaStream_t s = as_stream(stream);
    if (n < 0) return set_error(MLAMG_EINVAL, "csr_to_dense: n < 0");
    if (n == 0) return MLAMG_OK;
    const size_t esz = dtype == MLAMG_F32 ? 4 : 8;
    MLAMG_CUDA(cudaMemsetAsync(dense, 0, (size_t)n * n * esz, s));
    MLAMG_DISPATCH(dtype, (csr_to_dense_kernel<T><<<cdiv((long long)n * 32, 256), 256, 0, s>>>(
                              n, rowptr, col, (const T *)val, (T *)dense)));
    MLAMG_LAUNCHED();
    return MLAMG_OK;
}

long long mlamg_poisson_nnz(int nx, int ny, int nz) {
    const long long X = nx, Y = ny, Z = nz;
    return X * Y * Z + 2 * ((X - 1) * Y * Z + X * (Y - 1) * Z + X * Y * (Z - 1));
}

int mlamg_poisson_csr_slab(int dtype, int nx, int ny, int nz, int z0, int nz_local, int *rowptr, int *col, void *val,
                           long long *nnz_host, mlamg_stream_t stream) {
    cudaStream_t s = as_stream(stream);
    if (nx < 1 || ny < 1 || nz < 1 || z0 < 0 || nz_local < 1 || z0 + nz_local > nz)
        return set_error(MLAMG_EINVAL, "poisson: bad shape / slab");
    const long long N = (long long)nx * ny * nz_local;
    if ((long long)nx * ny * nz > 0x7fffffffLL || 7 * N > 0x7fffffffLL)
        return set_error(MLAMG_ELIMIT, "poisson: int32 index overflow");
    poisson_count_kernel<<<cdiv(N, 256), 256, 0, s>>>(nx, ny, nz, z0, nz_local, rowptr);
    MLAMG_LAUNCHED();
    MLAMG_TRY(exclusive_scan_i32(rowptr, rowptr, (int)N, s));
    MLAMG_DISPATCH(dtype, (poisson_fill_kernel<T><<<cdiv(N, 256), 256, 0, s>>>(nx, ny, nz, z0, nz_local, rowptr, col,
                                                                            (T *)val)));
    MLAMG_LAUNCHED();
    if (nnz_host) {
        int h = 0;
        MLAMG_CUDA(cudaMemcpyAsync(&h, rowptr + N, sizeof(int), cudaMemcpyDeviceToHost, s));
        MLAMG_CUDA(cudaStreamSynchronize(s));
        *nnz_host = h;
    }
    return MLAMG_OK;
}

int mlamg_poisson_csr(int dtype, int nx, int ny, int nz, int *rowptr, int *col, void *val, mlamg_stream_t stream) {
    return mlamg_poisson_csr_slab(dtype, nx, ny, nz, 0, nz, rowptr, col, val, nullptr, stream);
}


}  // extern "C"
