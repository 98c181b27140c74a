// Two-phase (symbolic / numeric) hash SpGEMM  C = A * B  in CSR.
//
// Replaces scipy `csr_matmat` (SMMP) at ns/lib/multigrid.py:107 ((I-wD^-1A)@Agg) and :165 (P.T@A@P),
// torch.sparse.mm at ns/model/agg_interp.py:484 and torch_sparse.spspmm at ns/model/loss.py:53-54.
//
// Rows are binned by their product upper bound ub_i = sum_{k in A_i} nnz(B_k):
//   ub <= 128            : one warp per row, 256-slot table in shared memory (per warp)
//   larger               : one CTA per row, 1024/4096/16384-slot shared-memory table picked from the
//                          exact row length found by the symbolic phase (32768 key slots in the
//                          symbolic phase itself)
// Accumulation: linear-probing insert with atomicCAS on the key and a shared-memory atomicAdd on
// the value (fp64 shared atomics are native on sm_100a).  Rows are emitted unsorted and then sorted
// by mlamg_csr_sort_rows (sort.cu) so the result is the canonical CSR scipy's sort_indices gives.
// Rows with more than 8192 distinct columns are rejected (MLAMG_ELIMIT): a Galerkin operator that
// dense belongs to the dense coarse solver, not to a sparse level.
#include "common.cuh"

namespace mlamg {

int sort_rows_impl(int dtype, int m, const int *rowptr, int *col, void *val, cudaStream_t s);

constexpr int WARP_TS = 256;          // table slots of the warp-per-row bin
constexpr int WARP_UB = 128;          // product upper bound handled by that bin
constexpr int WARPS_PER_CTA = 4;
constexpr int SYM_BIG_TS = 32768;     // key-only table of the big symbolic bin (128 KB)
constexpr int MAX_ROW_NNZ = 8192;

__device__ __forceinline__ unsigned hash_slot(int key, unsigned mask) { return ((unsigned)key * 107u) & mask; }

// returns 1 if `key` was newly inserted, 0 if present, -1 if the table is full
__device__ __forceinline__ int table_insert(int *keys, unsigned mask, int key, unsigned *slot_out) {
    unsigned h = hash_slot(key, mask);
    for (unsigned probes = 0; probes <= mask; probes++) {
        int cur = ((volatile int *)keys)[h];
        if (cur == key) { *slot_out = h; return 0; }
        if (cur == -1) {
            cur = atomicCAS(&keys[h], -1, key);
            if (cur == -1) { *slot_out = h; return 1; }
            if (cur == key) { *slot_out = h; return 0; }
        }
        h = (h + 1) & mask;
    }
    return -1;
}

// ---------------------------------------------------------------- binning
__global__ void __launch_bounds__(256) product_bound_kernel(int m, const int *__restrict__ a_rowptr,
                                                            const int *__restrict__ a_col,
                                                            const int *__restrict__ b_rowptr,
                                                            int *__restrict__ ub) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    long long s = 0;
    for (int j = a_rowptr[i]; j < a_rowptr[i + 1]; j++) {
        const int k = a_col[j];
        s += b_rowptr[k + 1] - b_rowptr[k];
    }
    ub[i] = s > 0x7fffffffLL ? 0x7fffffff : (int)s;
}

// bin id per row from (ub, nnz)
__device__ __forceinline__ int bin_of(int ub, int nnz, int symbolic) {
    if (ub == 0) return -1;
    if (ub <= WARP_UB) return 0;
    (void)nnz;
    (void)symbolic;
    return ub <= 2048 ? 1 : 2;
}

__global__ void __launch_bounds__(256) spgemm_binid_kernel(int m, const int *__restrict__ ub,
                                                           const int *__restrict__ c_rowptr, int symbolic,
                                                           int *__restrict__ binid) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int nnz = symbolic ? 0 : c_rowptr[i + 1] - c_rowptr[i];
    binid[i] = bin_of(ub[i], nnz, symbolic);
}

// ---------------------------------------------------------------- symbolic phase: warp-per-row kernel (ub <= 128)
// 4 sub-groups of 8 lanes: a sub-group takes one A entry, its lanes stride over that B row; distinct columns are counted
// through a shared-memory hash table (the numeric phase is the ORDERED kernels further down).
__global__ void __launch_bounds__(32 * WARPS_PER_CTA)
spgemm_symbolic_warp_kernel(int nrows, const int *__restrict__ rows, const int *__restrict__ a_rowptr,
                            const int *__restrict__ a_col, const int *__restrict__ b_rowptr,
                            const int *__restrict__ b_col, int *__restrict__ c_rowptr) {
    __shared__ int s_keys[WARPS_PER_CTA][WARP_TS];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long r = (long long)blockIdx.x * WARPS_PER_CTA + wid;
    if (r >= nrows) return;  // whole warp exits together
    const int row = rows[r];
    int *keys = s_keys[wid];
    for (int t = lane; t < WARP_TS; t += 32) keys[t] = -1;
    __syncwarp();
    const int sub = lane >> 3, sl = lane & 7;
    int fresh = 0;
    for (int ja = a_rowptr[row] + sub; ja < a_rowptr[row + 1]; ja += 4) {
        const int k = a_col[ja];
        for (int jb = b_rowptr[k] + sl; jb < b_rowptr[k + 1]; jb += 8) {
            unsigned slot;
            if (table_insert(keys, WARP_TS - 1, b_col[jb], &slot) > 0) fresh++;
        }
    }
    __syncwarp();
    fresh = warp_sum(fresh);
    if (lane == 0) c_rowptr[row] = fresh;  // counts; scanned by the caller
}

// ---------------------------------------------------------------- symbolic phase: CTA-per-row kernel
// dynamic shared memory: int keys[ts].  Warps take A entries, lanes stride over the B row.
__global__ void spgemm_symbolic_cta_kernel(int nrows, const int *__restrict__ rows, int ts,
                                           const int *__restrict__ a_rowptr, const int *__restrict__ a_col,
                                           const int *__restrict__ b_rowptr, const int *__restrict__ b_col,
                                           int *__restrict__ c_rowptr, int *__restrict__ overflow) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_cnt;
    __shared__ int s_full;
    int *keys = reinterpret_cast<int *>(smem_raw);
    const int row = rows[blockIdx.x];
    for (int t = threadIdx.x; t < ts; t += blockDim.x) keys[t] = -1;
    if (threadIdx.x == 0) { s_cnt = 0; s_full = 0; }
    __syncthreads();
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const unsigned mask = (unsigned)ts - 1u;
    int fresh = 0;
    for (int ja = a_rowptr[row] + wid; ja < a_rowptr[row + 1]; ja += nw) {
        const int k = a_col[ja];
        for (int jb = b_rowptr[k] + lane; jb < b_rowptr[k + 1]; jb += 32) {
            unsigned slot = 0;
            const int ins = table_insert(keys, mask, b_col[jb], &slot);
            if (ins < 0) { s_full = 1; break; }
            if (ins > 0) fresh++;
        }
    }
    fresh = warp_sum(fresh);
    if (lane == 0 && fresh) atomicAdd(&s_cnt, fresh);
    __syncthreads();
    if (threadIdx.x == 0) {
        c_rowptr[row] = s_cnt;
        if (s_full || s_cnt > MAX_ROW_NNZ) atomicExch(overflow, 1);
    }
}

// builds the permutation `rows` (size m) partitioned by bin; counts/offsets returned on the host
static int build_bins(int m, const int *ub, const int *c_rowptr, int symbolic, int *rows, Bins *bins, cudaStream_t s) {
    Scratch ids((size_t)m * sizeof(int), s);
    MLAMG_SCRATCH_OK(ids);
    spgemm_binid_kernel<<<cdiv(m, 256), 256, 0, s>>>(m, ub, c_rowptr, symbolic, ids.as<int>());
    MLAMG_LAUNCHED();
    return partition_rows_by_bin(m, ids.as<int>(), rows, bins, s);
}

static int launch_symbolic_cta(int nrows, const int *rows, int ts, int threads, const int *a_rowptr, const int *a_col,
                               const int *b_rowptr, const int *b_col, int *c_rowptr, int *overflow, cudaStream_t s) {
    if (nrows <= 0) return MLAMG_OK;
    const size_t smem = (size_t)ts * sizeof(int);
    MLAMG_CUDA(cudaFuncSetAttribute(spgemm_symbolic_cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    spgemm_symbolic_cta_kernel<<<nrows, threads, smem, s>>>(nrows, rows, ts, a_rowptr, a_col, b_rowptr, b_col, c_rowptr, overflow);
    MLAMG_LAUNCHED();
    return MLAMG_OK;
}

static int spgemm_symbolic_impl(int m, int k, int n, const int *a_rowptr, const int *a_col, const int *b_rowptr,
                                const int *b_col, int *c_rowptr, long long *nnz_host, cudaStream_t s) {
    if (m < 0 || k < 0 || n < 0) return set_error(MLAMG_EINVAL, "spgemm: negative dimension");
    if (m == 0) {
        MLAMG_TRY(exclusive_scan_i32(c_rowptr, c_rowptr, 0, s));
        if (nnz_host) *nnz_host = 0;
        return MLAMG_OK;
    }
    Scratch ubs((size_t)m * sizeof(int), s), rows((size_t)m * sizeof(int), s), ovf(sizeof(int), s);
    MLAMG_SCRATCH_OK(ubs);
    MLAMG_SCRATCH_OK(rows);
    MLAMG_SCRATCH_OK(ovf);
    MLAMG_CUDA(cudaMemsetAsync(ovf.p, 0, sizeof(int), s));
    MLAMG_CUDA(cudaMemsetAsync(c_rowptr, 0, (size_t)(m + 1) * sizeof(int), s));
    product_bound_kernel<<<cdiv(m, 256), 256, 0, s>>>(m, a_rowptr, a_col, b_rowptr, ubs.as<int>());
    MLAMG_LAUNCHED();
    Bins bins;
    MLAMG_TRY(build_bins(m, ubs.as<int>(), nullptr, 1, rows.as<int>(), &bins, s));
    const int *rl = rows.as<int>();
    if (bins.counts[0] > 0) {
        spgemm_symbolic_warp_kernel<<<cdiv(bins.counts[0], WARPS_PER_CTA), 32 * WARPS_PER_CTA, 0, s>>>(
            bins.counts[0], rl + bins.offsets[0], a_rowptr, a_col, b_rowptr, b_col, c_rowptr);
        MLAMG_LAUNCHED();
    }
    MLAMG_TRY(launch_symbolic_cta(bins.counts[1], rl + bins.offsets[1], 4096, 256, a_rowptr, a_col, b_rowptr, b_col, c_rowptr,
                                  ovf.as<int>(), s));
    MLAMG_TRY(launch_symbolic_cta(bins.counts[2], rl + bins.offsets[2], SYM_BIG_TS, 512, a_rowptr, a_col, b_rowptr, b_col,
                                  c_rowptr, ovf.as<int>(), s));
    MLAMG_TRY(exclusive_scan_i32(c_rowptr, c_rowptr, m, s));
    int h_ovf = 0, h_nnz = 0;
    MLAMG_CUDA(cudaMemcpyAsync(&h_ovf, ovf.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    MLAMG_CUDA(cudaMemcpyAsync(&h_nnz, c_rowptr + m, sizeof(int), cudaMemcpyDeviceToHost, s));
    MLAMG_CUDA(cudaStreamSynchronize(s));
    if (h_ovf) return set_error(MLAMG_ELIMIT, "spgemm: a result row has more than %d distinct columns", MAX_ROW_NNZ);
    if (h_nnz < 0) return set_error(MLAMG_ELIMIT, "spgemm: nnz overflows int32");
    if (nnz_host) *nnz_host = h_nnz;
    return MLAMG_OK;
}

// ---------------------------------------------------------------- ordered numeric kernels
// scipy's SMMP (csr_matmat) adds the products of one output entry in the order of the OUTER loop over
// the stored entries of the A row.  Inside one A entry the B row hits distinct output columns
// (canonical CSR), so those updates are independent: the kernels walk the A row sequentially, spread
// one B row over the lanes, update the hash-table value with a plain (non-atomic, non-fused)
// read-add-write and synchronise between A entries.  Result: every C_ik is bit-identical to the
// sequential loop (and identical run to run) — exact-zero decisions included.
__global__ void __launch_bounds__(256) numeric_binid_kernel(int m, const int *__restrict__ c_rowptr,
                                                            int *__restrict__ binid) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int nnz = c_rowptr[i + 1] - c_rowptr[i];
    int b;
    if (nnz == 0) b = -1;
    else if (nnz <= WARP_TS / 2) b = 0;
    else if (nnz <= 2048) b = 1;
    else if (nnz <= MAX_ROW_NNZ) b = 2;
    else b = 4;
    binid[i] = b;
}

template <typename T>
__global__ void __launch_bounds__(32 * WARPS_PER_CTA)
spgemm_warp_ordered_kernel(int nrows, const int *__restrict__ rows, const int *__restrict__ a_rowptr,
                           const int *__restrict__ a_col, const T *__restrict__ a_val,
                           const int *__restrict__ b_rowptr, const int *__restrict__ b_col,
                           const T *__restrict__ b_val, const int *__restrict__ c_rowptr, int *__restrict__ c_col,
                           T *__restrict__ c_val) {
    __shared__ int s_keys[WARPS_PER_CTA][WARP_TS];
    __shared__ T s_vals[WARPS_PER_CTA][WARP_TS];
    __shared__ int s_cnt[WARPS_PER_CTA];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long r = (long long)blockIdx.x * WARPS_PER_CTA + wid;
    if (r >= nrows) return;  // whole warp exits together
    const int row = rows[r];
    int *keys = s_keys[wid];
    T *vals = s_vals[wid];
    for (int t = lane; t < WARP_TS; t += 32) { keys[t] = -1; vals[t] = (T)0; }
    if (lane == 0) s_cnt[wid] = 0;
    __syncwarp();
    const int a0 = a_rowptr[row], a1 = a_rowptr[row + 1];
    for (int base = a0; base < a1; base += 32) {
        // every lane prefetches one A entry and the extent of its B row
        int bs = 0, be = 0;
        T av = (T)0;
        if (base + lane < a1) {
            const int k = a_col[base + lane];
            av = a_val[base + lane];
            bs = b_rowptr[k];
            be = b_rowptr[k + 1];
        }
        const int cnt = min(32, a1 - base);
        for (int e = 0; e < cnt; e++) {       // sequential over the A entries: scipy's accumulation order
            const int ebs = __shfl_sync(0xffffffffu, bs, e);
            const int ebe = __shfl_sync(0xffffffffu, be, e);
            const T eav = __shfl_sync(0xffffffffu, av, e);
            for (int jb = ebs + lane; jb < ebe; jb += 32) {
                unsigned slot = 0;
                table_insert(keys, WARP_TS - 1, b_col[jb], &slot);
                vals[slot] = add_rn(vals[slot], mul_rn(eav, b_val[jb]));
            }
            __syncwarp();
        }
    }
    const int base = c_rowptr[row];
    for (int t = lane; t < WARP_TS; t += 32) {
        const int key = keys[t];
        if (key != -1) {
            const int pos = atomicAdd(&s_cnt[wid], 1);
            c_col[base + pos] = key;
            c_val[base + pos] = vals[t];
        }
    }
}

template <typename T>
__global__ void spgemm_cta_ordered_kernel(int nrows, const int *__restrict__ rows, int ts,
                                          const int *__restrict__ a_rowptr, const int *__restrict__ a_col,
                                          const T *__restrict__ a_val, const int *__restrict__ b_rowptr,
                                          const int *__restrict__ b_col, const T *__restrict__ b_val,
                                          const int *__restrict__ c_rowptr, int *__restrict__ c_col,
                                          T *__restrict__ c_val, int *__restrict__ overflow) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_cnt;
    int *keys = reinterpret_cast<int *>(smem_raw);
    T *vals = reinterpret_cast<T *>(smem_raw + (size_t)ts * sizeof(int));
    const int row = rows[blockIdx.x];
    for (int t = threadIdx.x; t < ts; t += blockDim.x) { keys[t] = -1; vals[t] = (T)0; }
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const unsigned mask = (unsigned)ts - 1u;
    const int a0 = a_rowptr[row], a1 = a_rowptr[row + 1];
    int bs = 0, be = 0;
    T av = (T)0;
    if (a0 < a1) { const int k = a_col[a0]; av = a_val[a0]; bs = b_rowptr[k]; be = b_rowptr[k + 1]; }
    bool full = false;
    for (int ja = a0; ja < a1; ja++) {
        int nbs = 0, nbe = 0;
        T nav = (T)0;
        if (ja + 1 < a1) {                    // software prefetch of the next A entry
            const int k = a_col[ja + 1];
            nav = a_val[ja + 1];
            nbs = b_rowptr[k];
            nbe = b_rowptr[k + 1];
        }
        for (int jb = bs + threadIdx.x; jb < be; jb += blockDim.x) {
            unsigned slot = 0;
            if (table_insert(keys, mask, b_col[jb], &slot) < 0) { full = true; break; }
            vals[slot] = add_rn(vals[slot], mul_rn(av, b_val[jb]));
        }
        __syncthreads();
        bs = nbs; be = nbe; av = nav;
    }
    if (full) atomicExch(overflow, 1);
    const int base = c_rowptr[row];
    for (int t = threadIdx.x; t < ts; t += blockDim.x) {
        const int key = keys[t];
        if (key != -1) {
            const int pos = atomicAdd(&s_cnt, 1);
            c_col[base + pos] = key;
            c_val[base + pos] = vals[t];
        }
    }
}

template <typename T>
static int launch_cta_ordered(int nrows, const int *rows, int ts, int threads, const int *a_rowptr, const int *a_col,
                              const T *a_val, const int *b_rowptr, const int *b_col, const T *b_val,
                              const int *c_rowptr, int *c_col, T *c_val, int *overflow, cudaStream_t s) {
    if (nrows <= 0) return MLAMG_OK;
    const size_t smem = (size_t)ts * (sizeof(int) + sizeof(T));
    auto kern = spgemm_cta_ordered_kernel<T>;
    MLAMG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<nrows, threads, smem, s>>>(nrows, rows, ts, a_rowptr, a_col, a_val, b_rowptr, b_col, b_val, c_rowptr, c_col,
                                      c_val, overflow);
    MLAMG_LAUNCHED();
    return MLAMG_OK;
}

template <typename T>
static int spgemm_numeric_impl(int m, int k, int n, const int *a_rowptr, const int *a_col, const T *a_val,
                               const int *b_rowptr, const int *b_col, const T *b_val, const int *c_rowptr,
                               int *c_col, T *c_val, cudaStream_t s) {
    if (m <= 0) return m == 0 ? MLAMG_OK : set_error(MLAMG_EINVAL, "spgemm: m < 0");
    Scratch ids((size_t)m * sizeof(int), s), rows((size_t)m * sizeof(int), s), ovf(sizeof(int), s);
    MLAMG_SCRATCH_OK(ids);
    MLAMG_SCRATCH_OK(rows);
    MLAMG_SCRATCH_OK(ovf);
    MLAMG_CUDA(cudaMemsetAsync(ovf.p, 0, sizeof(int), s));
    numeric_binid_kernel<<<cdiv(m, 256), 256, 0, s>>>(m, c_rowptr, ids.as<int>());
    MLAMG_LAUNCHED();
    Bins bins;
    MLAMG_TRY(partition_rows_by_bin(m, ids.as<int>(), rows.as<int>(), &bins, s));
    if (bins.counts[4] > 0)
        return set_error(MLAMG_ELIMIT, "spgemm: %d result rows exceed %d entries", bins.counts[4], MAX_ROW_NNZ);
    const int *rl = rows.as<int>();
    if (bins.counts[0] > 0) {
        spgemm_warp_ordered_kernel<T><<<cdiv(bins.counts[0], WARPS_PER_CTA), 32 * WARPS_PER_CTA, 0, s>>>(
            bins.counts[0], rl + bins.offsets[0], a_rowptr, a_col, a_val, b_rowptr, b_col, b_val, c_rowptr, c_col, c_val);
        MLAMG_LAUNCHED();
    }
    MLAMG_TRY((launch_cta_ordered<T>(bins.counts[1], rl + bins.offsets[1], 4096, 256, a_rowptr, a_col, a_val, b_rowptr,
                                     b_col, b_val, c_rowptr, c_col, c_val, ovf.as<int>(), s)));
    MLAMG_TRY((launch_cta_ordered<T>(bins.counts[2], rl + bins.offsets[2], 16384, 512, a_rowptr, a_col, a_val, b_rowptr,
                                     b_col, b_val, c_rowptr, c_col, c_val, ovf.as<int>(), s)));
    MLAMG_TRY(sort_rows_impl(sizeof(T) == 4 ? MLAMG_F32 : MLAMG_F64, m, c_rowptr, c_col, c_val, s));
    int h_ovf = 0;
    MLAMG_CUDA(cudaMemcpyAsync(&h_ovf, ovf.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    MLAMG_CUDA(cudaStreamSynchronize(s));
    if (h_ovf) return set_error(MLAMG_EINVAL, "spgemm numeric: hash table overflow (c_rowptr not from the symbolic phase?)");
    return MLAMG_OK;
}

}  // namespace mlamg

using namespace mlamg;

extern "C" {

int mlamg_spgemm_symbolic(int m, int k, int n, const int *a_rowptr, const int *a_col, const int *b_rowptr,
                          const int *b_col, int *c_rowptr, long long *nnz_host, mlamg_stream_t stream) {
    return spgemm_symbolic_impl(m, k, n, a_rowptr, a_col, b_rowptr, b_col, c_rowptr, nnz_host, as_stream(stream));
}

int mlamg_spgemm_numeric(int dtype, int m, int k, int n, const int *a_rowptr, const int *a_col, const void *a_val,
                         const int *b_rowptr, const int *b_col, const void *b_val, const int *c_rowptr, int *c_col,
                         void *c_val, mlamg_stream_t stream) {
    MLAMG_DISPATCH(dtype, return spgemm_numeric_impl<T>(m, k, n, a_rowptr, a_col, (const T *)a_val, b_rowptr, b_col,
                                                        (const T *)b_val, c_rowptr, c_col, (T *)c_val,
                                                        as_stream(stream)));
    return MLAMG_OK;
}

}  // extern "C"
