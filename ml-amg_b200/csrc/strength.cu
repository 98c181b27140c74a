// Evolution strength of connection (pyamg.strength.evolution_strength_of_connection with its defaults), the
// measure behind the reference's 'evolution' and default 'olson' strength functions
// (/root/reference/utils/common.py:27,30 -> Lloyd's distance matrix, utils/common.py:58).
//
// pyamg evaluates it with scipy sparse arithmetic and three amg_core loops; every step is elementwise on a CSR
// pattern except one sparse product restricted to the pattern of A.  Here each step is one kernel, a warp per row,
// lanes over the row's entries; all arithmetic is explicitly NON-fused and in pyamg's operation order, so that —
// for the same rho — the measure has the bits of the CPU evaluation (the Lloyd aggregation that consumes it breaks
// ties between equal distances, which structured grids are full of).  HBM-bound: every kernel reads the pattern and
// one or two value arrays once.
//
//   evolution_step      S = I - (1/rho) D^-1 A on A's pattern           (scale_rows, scalar *, eye - X)
//   incomplete_matmul   Z(i,j) = <T(i,:), B(:,j)> on a given pattern     (amg_core incomplete_mat_mult_csr: sorted merge,
//                                                                          products summed in increasing inner index)
//   evolution_measure   m_ij = |1 - z_ii / z_ij| with the weak-ratio / angle / near-perfect rules (NullDim == 1 shortcut)
//   distance_filter     off-diagonals >= epsilon * (row's smallest off-diagonal) -> 0   (amg_core apply_distance_filter)
//   symmetrize          0.5 (M + M^T) + unit diagonal, evaluated on A's (symmetric) pattern
//   invert_scale_rows   m <- 1/m, rows scaled by the reciprocal of their largest entry (scale_rows_by_largest_entry)
//   pattern_add         C = E + W for W on A's pattern, E on a sub-pattern       ('olson': W = 1/|A|, 'evolution': W = 0.1)
#include "common.cuh"

namespace mlamg {

constexpr unsigned FULL = 0xffffffffu;

// value of the stored diagonal entry of `row` (0 when absent), same on every lane; *present = whether it is stored
template <typename T>
__device__ __forceinline__ T row_diagonal(int row, int start, int end, const int *__restrict__ col,
                                          const T *__restrict__ val, int lane, bool *present) {
    T d = (T)0;
    bool has = false;
    for (int j = start + lane; j < end; j += 32)
        if (col[j] == row) { d = val[j]; has = true; }
    const unsigned m = __ballot_sync(FULL, has);
    *present = m != 0u;
    if (m == 0u) return (T)0;
    return __shfl_sync(FULL, d, __ffs(m) - 1);
}

template <typename T>
__device__ __forceinline__ T row_lookup(const int *__restrict__ rowptr, const int *__restrict__ col,
                                        const T *__restrict__ val, int r, int c, bool *found) {
    for (int q = rowptr[r]; q < rowptr[r + 1]; q++)
        if (col[q] == c) { *found = true; return val[q]; }
    *found = false;
    return (T)0;
}

template <typename T>
__global__ void __launch_bounds__(256) evolution_step_kernel(int n, const int *__restrict__ rowptr,
                                                             const int *__restrict__ col, const T *__restrict__ val,
                                                             T inv_rho, T *__restrict__ s_val,
                                                             T *__restrict__ dinv_a_val, int *__restrict__ flags) {
    const long long rowl = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;     // warp-uniform
    const int lane = threadIdx.x & 31;
    if (rowl >= n) return;
    const int row = (int)rowl;
    const int start = rowptr[row], end = rowptr[row + 1];
    bool present;
    const T d = row_diagonal(row, start, end, col, val, lane, &present);
    if (!present && lane == 0) atomicExch(flags, 1);          // eye - X would add an entry: caller must store the diagonal
    const T dinv = (d != (T)0) ? div_rn((T)1, d) : (T)1;      // Dinv[D == 0] = 1.0
    for (int j = start + lane; j < end; j += 32) {
        const T t = mul_rn(val[j], dinv);                      // scale_rows(A, Dinv)
        const T s = mul_rn(inv_rho, t);                        // (1.0 / rho) * Dinv_A
        s_val[j] = (col[j] == row) ? add_rn((T)1, -s) : -s;    // eye - X
        if (dinv_a_val) dinv_a_val[j] = t;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) incomplete_matmul_kernel(int n, const int *__restrict__ Ap,
                                                                const int *__restrict__ Aj, const T *__restrict__ Ax,
                                                                const int *__restrict__ Bp, const int *__restrict__ Bj,
                                                                const T *__restrict__ Bx, const int *__restrict__ Sp,
                                                                const int *__restrict__ Sj, T *__restrict__ Sx) {
    const long long rowl = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (rowl >= n) return;
    const int row = (int)rowl;
    const int a0 = Ap[row], a1 = Ap[row + 1];
    for (int k = Sp[row] + lane; k < Sp[row + 1]; k += 32) {
        const int c = Sj[k];
        T sum = (T)0;
        int a = a0, b = Bp[c];
        const int b1 = Bp[c + 1];
        while (a < a1 && b < b1) {
            const int ac = Aj[a], br = Bj[b];
            if (ac == br) { sum = add_rn(sum, mul_rn(Ax[a], Bx[b])); a++; b++; }
            else if (ac < br) a++;
            else b++;
        }
        Sx[k] = sum;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) evolution_measure_kernel(int n, const int *__restrict__ rowptr,
                                                                const int *__restrict__ col, T *__restrict__ val) {
    const long long rowl = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (rowl >= n) return;
    const int row = (int)rowl;
    const int start = rowptr[row], end = rowptr[row + 1];
    bool present;
    const T d = row_diagonal(row, start, end, col, val, lane, &present);      // DAtilde / B, B = 1
    __syncwarp();                                                              // every lane has read the diagonal before it is overwritten
    const T near_perfect = (T)1.4901161193847656e-08;                          // sqrt(eps of double), as pyamg writes it
    for (int j = start + lane; j < end; j += 32) {
        const T z = val[j];
        T m = (T)0;
        if (z != (T)0) {                                                       // explicit zeros were eliminated before this step
            const bool angle = mul_rn(d, z) < (T)0;
            const T ratio = div_rn(d, z);
            const bool weak = fabs(ratio) < (T)1e-4;
            m = fabs(add_rn((T)1, -ratio));
            if (weak || angle) m = (T)0;
            if (m != (T)0 && m < near_perfect) m = (T)1e-4;
        }
        val[j] = m;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) distance_filter_kernel(int n, T epsilon, const int *__restrict__ rowptr,
                                                              const int *__restrict__ col, T *__restrict__ val) {
    const long long rowl = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (rowl >= n) return;
    const int row = (int)rowl;
    const int start = rowptr[row], end = rowptr[row + 1];
    T mn = Limits<T>::max();
    for (int j = start + lane; j < end; j += 32)
        if (col[j] != row && val[j] < mn) mn = val[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const T other = __shfl_xor_sync(FULL, mn, o);
        if (other < mn) mn = other;
    }
    const T threshold = mul_rn(epsilon, mn);
    for (int j = start + lane; j < end; j += 32)
        if (val[j] >= threshold && col[j] != row) val[j] = (T)0;
}

template <typename T>
__global__ void __launch_bounds__(256) evolution_symmetrize_kernel(int n, const int *__restrict__ a_rowptr,
                                                                   const int *__restrict__ a_col,
                                                                   const int *__restrict__ m_rowptr,
                                                                   const int *__restrict__ m_col,
                                                                   const T *__restrict__ m_val, int symmetrize,
                                                                   T *__restrict__ out) {
    const long long rowl = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (rowl >= n) return;
    const int row = (int)rowl;
    for (int j = a_rowptr[row] + lane; j < a_rowptr[row + 1]; j += 32) {
        const int c = a_col[j];
        T v;
        if (c == row) {
            v = (T)1;                                          // Atilde + (I - diag(Atilde)): the measure has no diagonal left
        } else {
            bool fa, fb = false;
            const T a = row_lookup(m_rowptr, m_col, m_val, row, c, &fa);
            T b = (T)0;
            if (symmetrize) b = row_lookup(m_rowptr, m_col, m_val, c, row, &fb);
            if (!symmetrize) v = a;
            else if (fa && fb) v = mul_rn((T)0.5, add_rn(a, b));
            else v = mul_rn((T)0.5, fa ? a : b);               // present on one side only (0 when on neither: dropped later)
        }
        out[j] = v;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) invert_scale_rows_kernel(int n, const int *__restrict__ rowptr,
                                                                T *__restrict__ val) {
    const long long rowl = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (rowl >= n) return;
    const int row = (int)rowl;
    const int start = rowptr[row], end = rowptr[row + 1];
    T mx = sizeof(T) == 8 ? (T)2.2250738585072014e-308 : (T)1.17549435e-38f;   // numeric_limits<T>::min(), as amg_core starts
    for (int j = start + lane; j < end; j += 32) {
        const T inv = fabs(div_rn((T)1, val[j]));
        if (inv > mx) mx = inv;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const T other = __shfl_xor_sync(FULL, mx, o);
        if (other > mx) mx = other;
    }
    const T r = (mx != (T)0) ? div_rn((T)1, mx) : mx;
    for (int j = start + lane; j < end; j += 32) val[j] = mul_rn(div_rn((T)1, val[j]), r);
}

template <typename T>
__global__ void __launch_bounds__(256) pattern_add_kernel(int n, const int *__restrict__ a_rowptr,
                                                          const int *__restrict__ a_col, const T *__restrict__ w,
                                                          const int *__restrict__ e_rowptr,
                                                          const int *__restrict__ e_col, const T *__restrict__ e_val,
                                                          T *__restrict__ out) {
    const long long rowl = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (rowl >= n) return;
    const int row = (int)rowl;
    for (int j = a_rowptr[row] + lane; j < a_rowptr[row + 1]; j += 32) {
        bool found;
        const T e = row_lookup(e_rowptr, e_col, e_val, row, a_col[j], &found);
        out[j] = found ? add_rn(e, w[j]) : w[j];
    }
}

}  // namespace mlamg

using namespace mlamg;

#define STRENGTH_GRID(n) cdiv((long long)(n) * 32, 256), 256, 0, s

extern "C" {

int mlamg_evolution_step(int dtype, int n, const int *rowptr, const int *col, const void *val, double inv_rho,
                         void *s_val, void *dinv_a_val, int *flags, mlamg_stream_t stream) {
    cudaStream_t s = as_stream(stream);
    if (n < 0) return set_error(MLAMG_EINVAL, "evolution_step: bad n");
    if (n == 0) return MLAMG_OK;
    MLAMG_DISPATCH(dtype, (evolution_step_kernel<T><<<STRENGTH_GRID(n)>>>(n, rowptr, col, (const T *)val, (T)inv_rho,
                                                                           (T *)s_val, (T *)dinv_a_val, flags)));
    MLAMG_LAUNCHED();
    return MLAMG_OK;
}

int mlamg_incomplete_matmul_csr(int dtype, int n, const int *Ap, const int *Aj, const void *Ax, const int *Bp,
                                const int *Bj, const void *Bx, const int *Sp, const int *Sj, void *Sx,
                                mlamg_stream_t stream) {
    cudaStream_t s = as_stream(stream);
    if (n < 0) return set_error(MLAMG_EINVAL, "incomplete_matmul: bad n");
    if (n == 0) return MLAMG_OK;
    MLAMG_DISPATCH(dtype, (incomplete_matmul_kernel<T><<<STRENGTH_GRID(n)>>>(n, Ap, Aj, (const T *)Ax, Bp, Bj,
                                                                              (const T *)Bx, Sp, Sj, (T *)Sx)));
    MLAMG_LAUNCHED();
    return MLAMG_OK;
}

int mlamg_evolution_measure(int dtype, int n, const int *rowptr, const int *col, void *val, mlamg_stream_t stream) {
    cudaStream_t s = as_stream(stream);
    if (n < 0) return set_error(MLAMG_EINVAL, "evolution_measure: bad n");
    if (n == 0) return MLAMG_OK;
    MLAMG_DISPATCH(dtype, (evolution_measure_kernel<T><<<STRENGTH_GRID(n)>>>(n, rowptr, col, (T *)val)));
    MLAMG_LAUNCHED();
    return MLAMG_OK;
}

int mlamg_distance_filter(int dtype, int n, double epsilon, const int *rowptr, const int *col, void *val,
                          mlamg_stream_t stream) {
    cudaStream_t s = as_stream(stream);
    if (n < 0) return set_error(MLAMG_EINVAL, "distance_filter: bad n");
    if (n == 0) return MLAMG_OK;
    MLAMG_DISPATCH(dtype, (distance_filter_kernel<T><<<STRENGTH_GRID(n)>>>(n, (T)epsilon, rowptr, col, (T *)val)));
    MLAMG_LAUNCHED();
    return MLAMG_OK;
}

int mlamg_evolution_symmetrize(int dtype, int n, const int *a_rowptr, const int *a_col, const int *m_rowptr,
                               const int *m_col, const void *m_val, int symmetrize, void *out,
                               mlamg_stream_t stream) {
    cudaStream_t s = as_stream(stream);
    if (n < 0) return set_error(MLAMG_EINVAL, "evolution_symmetrize: bad n");
    if (n == 0) return MLAMG_OK;
    MLAMG_DISPATCH(dtype, (evolution_symmetrize_kernel<T><<<STRENGTH_GRID(n)>>>(n, a_rowptr, a_col, m_rowptr, m_col,
                                                                                 (const T *)m_val, symmetrize, (T *)out)));
    MLAMG_LAUNCHED();
    return MLAMG_OK;
}

int mlamg_invert_scale_rows(int dtype, int n, const int *rowptr, void *val, mlamg_stream_t stream) {
    cudaStream_t s = as_stream(stream);
    if (n < 0) return set_error(MLAMG_EINVAL, "invert_scale_rows: bad n");
    if (n == 0) return MLAMG_OK;
    MLAMG_DISPATCH(dtype, (invert_scale_rows_kernel<T><<<STRENGTH_GRID(n)>>>(n, rowptr, (T *)val)));
    MLAMG_LAUNCHED();
    return MLAMG_OK;
}

int mlamg_csr_pattern_add(int dtype, int n, const int *a_rowptr, const int *a_col, const void *w, const int *e_rowptr,
                          const int *e_col, const void *e_val, void *out, mlamg_stream_t stream) {
    cudaStream_t s = as_stream(stream);
    if (n < 0) return set_error(MLAMG_EINVAL, "pattern_add: bad n");
    if (n == 0) return MLAMG_OK;
    MLAMG_DISPATCH(dtype, (pattern_add_kernel<T><<<STRENGTH_GRID(n)>>>(n, a_rowptr, a_col, (const T *)w, e_rowptr, e_col,
                                                                        (const T *)e_val, (T *)out)));
    MLAMG_LAUNCHED();
    return MLAMG_OK;
}

}  // extern "C"
