// V-cycle apply kernels: CSR SpMV family (SpMV, residual(+norm), fused Jacobi sweep, prolong-add),
// smoother diagonals, multi-vector SpMM, dense GEMV for the coarsest level.
//
// Roofline: every kernel here is HBM-bound (SpMV intensity ~2 flop / 12 B).  One pass over the
// operator per kernel; x gathers are served by L1/L2 (banded operators) — algorithmic bytes per
// SURVEY.md §8(d):  B_spmv = nnz(v+4) + 4(N+1) + 2vN,  B_jacobi = nnz(v+4) + 4(N+1) + 4vN.
#include <string.h>
#include "common.cuh"
#include "peer.cuh"

namespace mlamg {

// OP_RESZERO: first pre-smoothing sweep from a zero guess fused with the residual: x = dw.*b is never read back
// from HBM — the gathers evaluate dw[c]*b[c] on the fly, the epilogue stores x[row] and r[row] = b[row] - (A x)[row]
// OP_PSMOOTH: prolongation fused with the first post-smoothing sweep.  With r = b - A x already known (it was
// computed for the restriction), x + P e followed by one sweep equals  x + dw.*r + Q e,  Q = (I - D_w A) P
// precomputed at setup: one pass over Q instead of a pass over P and a pass over A.
// OP_RESZERO_S: OP_RESZERO on the column-scaled operator A D_w (values a_ij * dw_j, built once at setup): the
// gathers read b[c] alone — r = b - (A D_w) b — so the pass runs at the speed of a plain residual.
// OP_PSMOOTH0: OP_PSMOOTH when the iterate before the correction is the zero-guess sweep x = dw.*b itself:
// y = dw.*(b + r) + Q e.  The pass on the way down then is a plain residual on the scaled copy, r = b - (A D_w) b,
// that neither stores x nor reads dw (two vectors less traffic per level).
enum { OP_SPMV = 0, OP_SPMV_ADD = 1, OP_RESIDUAL = 2, OP_JACOBI = 3, OP_RESZERO = 4, OP_PSMOOTH = 5, OP_RESZERO_S = 6,
       OP_PSMOOTH0 = 7 };

constexpr int ROW_THREADS = 256;

template <typename T, int OP, bool NORM>
__device__ __forceinline__ double row_epilogue(long long row, T sum, const T *__restrict__ x, const T *__restrict__ b,
                                               const T *__restrict__ dw, T *y, T *y2) {
    double rr = 0.0;
    if (OP == OP_SPMV) {
        y[row] = sum;
    } else if (OP == OP_SPMV_ADD) {
        y[row] += sum;
    } else if (OP == OP_RESIDUAL) {
        const T r = b[row] - sum;
        y[row] = r;
        if (NORM) rr = (double)r * (double)r;
    } else if (OP == OP_RESZERO || OP == OP_RESZERO_S) {
        const T br = b[row];
        y2[row] = dw[row] * br;
        const T r = br - sum;
        y[row] = r;
        if (NORM) rr = (double)r * (double)r;
    } else if (OP == OP_PSMOOTH) {   // y2 = x before the correction (may alias y), b = residual
        y[row] = y2[row] + dw[row] * b[row] + sum;
    } else if (OP == OP_PSMOOTH0) {  // y2 = right-hand side, b = residual of x = dw.*rhs
        y[row] = dw[row] * (y2[row] + b[row]) + sum;
    } else {  // OP_JACOBI
        y[row] = x[row] + dw[row] * (b[row] - sum);
    }
    return rr;
}

// CSR: LANES threads cooperate on one row (LANES = 1..32, power of two).  Consecutive rows are
// contiguous in CSR, so a warp always reads one contiguous span of col/val; the inner loop is
// unrolled 4x so every thread keeps 4 (col,val) pairs and 4 gathers of x in flight (HBM latency is
// hidden by memory-level parallelism, not by occupancy alone).  Segmented shuffle reduction.
// HALO: columns >= halo.n_own are read in place from a peer channel's receive region (values written by the
// neighbouring GPUs over NVLink, each carrying a sequence tag; the load spins on the few that have not landed).
template <typename T, int LANES, int OP, bool NORM, bool HALO, int NBT>
// 32 registers / 8 CTAs per SM for every non-halo variant with <= 4 entries in flight per lane (none spills): the row ops
// are latency-bound gathers, occupancy is what hides them (restriction 136 -> 126 us with 8 instead of 6 CTAs per SM).  The
// halo-gathering variants stay at 6 CTAs per SM: at 32 registers they spill inside the entry loop (ptxas: 12-76 bytes), and the
// boundary rows they process run beside the interior kernel, where every wasted instruction is taken from it (2 GPUs,
// sustained: 0.908 vs 0.918-0.928 ms per cycle)
__global__ void __launch_bounds__(ROW_THREADS, (!HALO && NBT <= 4) ? 8 : (NBT <= 4 ? 6 : 4))
csr_rowop_kernel(int n, const int *__restrict__ rowptr, const int *__restrict__ col,
                 const T *__restrict__ val, const T *__restrict__ x, const T *__restrict__ b,
                 const T *__restrict__ dw, T *y, double *__restrict__ partial,
                 const int *__restrict__ row_order, int row0, const HaloLL halo, T *y2) {
    const long long gtid = (long long)blockIdx.x * ROW_THREADS + threadIdx.x;
    long long row = gtid / LANES;
    const int lane = threadIdx.x & (LANES - 1);
    T sum = (T)0;
    unsigned tag = 0;
    const T *hreg = nullptr;
    if (HALO) {
        const unsigned long long seq = *(volatile unsigned long long *)halo.state;   // advanced by this use's push
        tag = ll_tag(seq);
        hreg = reinterpret_cast<const T *>((seq & 1ull) ? halo.region[1] : halo.region[0]);
    }
#define XOWN(c) (OP == OP_RESZERO ? dw[(c)] * b[(c)] : (OP == OP_RESZERO_S ? b[(c)] : x[(c)]))
#define XLOAD(c) ((HALO && (c) >= halo.n_own) ? ll_load(hreg, (c) - halo.n_own, tag, halo.state) : XOWN(c))
    // row_order: optional list of the n rows to process (a permutation of all rows, or a subset).  Used by
    // the restriction, whose rows (aggregates) are numbered randomly by the reference's seeding — visiting
    // them in spatial order lets neighbouring aggregates share the fine-vector sectors they gather through
    // L2 — and by the multi-GPU levels (interior rows while the halo is in flight, then boundary rows).
    const bool valid = row < n;
    if (valid) row = row_order ? row_order[row] : row + row0;   // listed rows, or the range [row0, row0 + n)
    if (valid) {
        const int start = rowptr[row];
        const int end = rowptr[row + 1];
        // batches of NB predicated entries per lane: all col/val loads, then all x gathers, then the FMAs
        // (the fused zero-guess op gathers two vectors per entry: 2-entry batches keep it at 32 registers)
        constexpr int NB = NBT;
        for (int j = start + lane; j < end; j += NB * LANES) {
            bool p[NB];
            int c[NB];
            T v[NB], xv[NB];
#pragma unroll
            for (int k = 0; k < NB; k++) p[k] = (k == 0) || (j + k * LANES < end);
#pragma unroll
            for (int k = 0; k < NB; k++) c[k] = p[k] ? col[j + k * LANES] : 0;
#pragma unroll
            for (int k = 0; k < NB; k++) v[k] = p[k] ? val[j + k * LANES] : (T)0;
#pragma unroll
            for (int k = 0; k < NB; k++) xv[k] = p[k] ? XLOAD(c[k]) : (T)0;
#pragma unroll
            for (int k = 0; k < NB; k++) sum += v[k] * xv[k];
        }
    }
#undef XLOAD
#undef XOWN
    if (LANES > 1) sum = group_sum<LANES>(sum);
    double rr = 0.0;
    if (valid && lane == 0) rr = row_epilogue<T, OP, NORM>(row, sum, x, b, dw, y, y2);
    if (NORM) {
        __shared__ double sm[32];
        rr = block_sum(rr, sm);
        if (threadIdx.x == 0) partial[blockIdx.x] = rr;
    }
}

// TMA-staged thread-per-row kernel for short rows (the fine-level Q, A: mean <= 12 entries).  A CTA owns 256 consecutive
// rows, i.e. ONE contiguous span of col and of val.  An elected thread arms an mbarrier with the span's byte count and
// issues two 1-D bulk copies (cp.async.bulk global -> shared, completion signalled on the mbarrier: SASS UBLKCP); the
// copies go through the TMA path, not through the LSU/L1, so L1 is left to the gathers of x and the per-row vectors —
// the plain thread-per-row kernel asks L1 for every 128-byte line of the span once per entry of a row and is bound by
// that request rate (ncu: 83 % L1 request rate on OP_PSMOOTH0, DRAM traffic already algorithmic).  While the copy is in
// flight the threads load their row bounds and per-row operands; with 8 CTAs per SM other CTAs compute meanwhile.
// The previous staged variant (plain coalesced loads + two CTA barriers) lost to the plain kernel, 436 vs 290 us: its
// barriers serialised load and gather phases; here there is no CTA barrier after the mbarrier initialisation.
// Bulk copies need 16-byte aligned addresses and sizes: the span start is aligned down (at most 3 ints / 1 value of the
// previous rows), only whole 16-byte units are copied and the elected thread loads the < 16-byte tail itself, so nothing
// is read outside [0, end); spans larger than the shared-memory capacity take the plain loads.
constexpr int STAGE_CAP = 2048;      // entries per CTA: 8 KB + 16 KB of shared memory (f64), 8 CTAs per SM

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

template <typename T, int OP, bool NORM>
__global__ void __launch_bounds__(ROW_THREADS, 8)
csr_tma_rowop_kernel(int n, const int *__restrict__ rowptr, const int *__restrict__ col,
                     const T *__restrict__ val, const T *__restrict__ x, const T *__restrict__ b,
                     const T *__restrict__ dw, T *y, double *__restrict__ partial, int row0, T *y2) {
    __shared__ __align__(16) int s_col[STAGE_CAP + 8];
    __shared__ __align__(16) T s_val[STAGE_CAP + 4];
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ int s_span[2];
    const int t = threadIdx.x;
    const long long first = (long long)blockIdx.x * ROW_THREADS;        // first row of this CTA, relative to row0
    const int nr = (int)min((long long)ROW_THREADS, (long long)n - first);
    const long long row = first + t + row0;
    const bool valid = t < nr;
    if (t == 0) {
        const int beg = rowptr[first + row0], end = rowptr[first + row0 + nr];
        s_span[0] = beg;
        s_span[1] = end;
        constexpr int CA = 16 / sizeof(int), VA = 16 / sizeof(T);
        const int cb = beg & ~(CA - 1), vb = beg & ~(VA - 1);              // aligned starts (inside the arrays)
        // whole 16-byte units only: the copies never read past `end`; the < 16-byte tails are loaded below
        const unsigned cbytes = ((unsigned)(end - cb) * (unsigned)sizeof(int)) & ~15u;
        const unsigned vbytes = ((unsigned)(end - vb) * (unsigned)sizeof(T)) & ~15u;
        const bool staged = end > beg && end - cb <= STAGE_CAP + 8 && end - vb <= STAGE_CAP + 4 && cbytes > 0 && vbytes > 0 &&
                            (reinterpret_cast<unsigned long long>(col) & 15ull) == 0 && (reinterpret_cast<unsigned long long>(val) & 15ull) == 0;
        if (staged) {
            const unsigned bar = smem_u32(&s_bar);
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(cbytes + vbytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(s_col)),
                         "l"(col + cb), "r"(cbytes), "r"(bar)
                         : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(s_val)),
                         "l"(val + vb), "r"(vbytes), "r"(bar)
                         : "memory");
            for (int j = cb + (int)(cbytes / sizeof(int)); j < end; j++) s_col[j - cb] = col[j];
            for (int j = vb + (int)(vbytes / sizeof(T)); j < end; j++) s_val[j - vb] = val[j];
        } else {
            s_span[0] = -1 - beg;       // negative: not staged
        }
    }
    // per-row operands while the bulk copies are in flight
    int my_s = 0, my_e = 0;
    if (valid) {
        my_s = rowptr[row];
        my_e = rowptr[row + 1];
    }
    __syncthreads();                    // s_span and the initialised barrier are visible
    const bool staged = s_span[0] >= 0;
    T sum = (T)0;
    constexpr int NB = (OP == OP_RESZERO) ? 2 : 4;
#define XOWN(c) (OP == OP_RESZERO ? dw[(c)] * b[(c)] : (OP == OP_RESZERO_S ? b[(c)] : x[(c)]))
    if (staged) {
        const int beg = s_span[0];
        constexpr int CA = 16 / sizeof(int), VA = 16 / sizeof(T);
        const int coff = beg & ~(CA - 1), voff = beg & ~(VA - 1);
        const unsigned bar = smem_u32(&s_bar);
        unsigned done = 0;
        while (!done) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                         : "=r"(done)
                         : "r"(bar), "r"(0u)
                         : "memory");
        }
        for (int j = my_s; j < my_e; j += NB) {
            bool p[NB];
            int c[NB];
            T xv[NB];
#pragma unroll
            for (int k = 0; k < NB; k++) p[k] = (k == 0) || (j + k < my_e);
#pragma unroll
            for (int k = 0; k < NB; k++) c[k] = p[k] ? s_col[j + k - coff] : 0;
#pragma unroll
            for (int k = 0; k < NB; k++) xv[k] = p[k] ? XOWN(c[k]) : (T)0;
#pragma unroll
            for (int k = 0; k < NB; k++) sum += (p[k] ? s_val[j + k - voff] : (T)0) * xv[k];
        }
    } else {
        for (int j = my_s; j < my_e; j += NB) {
            bool p[NB];
            int c[NB];
            T v[NB], xv[NB];
#pragma unroll
            for (int k = 0; k < NB; k++) p[k] = (k == 0) || (j + k < my_e);
#pragma unroll
            for (int k = 0; k < NB; k++) c[k] = p[k] ? col[j + k] : 0;
#pragma unroll
            for (int k = 0; k < NB; k++) v[k] = p[k] ? val[j + k] : (T)0;
#pragma unroll
            for (int k = 0; k < NB; k++) xv[k] = p[k] ? XOWN(c[k]) : (T)0;
#pragma unroll
            for (int k = 0; k < NB; k++) sum += v[k] * xv[k];
        }
    }
#undef XOWN
    double rr = 0.0;
    if (valid) rr = row_epilogue<T, OP, NORM>(row, sum, x, b, dw, y, y2);
    if (NORM) {
        __shared__ double sm[32];
        rr = block_sum(rr, sm);
        if (threadIdx.x == 0) partial[blockIdx.x] = rr;
    }
}

// ---------------------------------------------------------------------------------------------------- W32 layout
// "CSR interleaved by warp": the rowptr of the CSR matrix is kept, but inside every window of 32 consecutive rows the
// entries are stored SLOT-major: first the 0-th entries of all rows of the window that have one (in row order), then the
// 1-st entries, ...  No padding (the array has exactly nnz entries), no row permutation (so the per-row vectors and the
// gathers of a warp keep their locality), and the k-th entries of a warp's rows are CONTIGUOUS: the thread-per-row kernel
// finds its entry at  base + (entries of earlier slots) + (rank of its lane among the lanes that still have an entry),
// both computed with one ballot + popc per slot — nothing but the row lengths is read.  With plain CSR a warp's k-th
// entries lie 5.5 entries apart (Q) and every col / val request touches 22 / 32 sectors (ncu source counters:
// 160 M of the kernel's 225 M L1 sectors, 37 M ideal); here they are 4 / 8.
// Rows [row_begin, row_end) of an operator with n_total rows; windows are aligned to absolute row numbers, so the launch
// starts at the window containing row_begin and rows of the first / last window outside the range only take part in the
// ballots (their entries occupy positions of the window too).
template <typename T, int OP>
__global__ void __launch_bounds__(ROW_THREADS, 8)
csr_w32_rowop_kernel(int row_begin, int row_end, int n_total, const int *__restrict__ rowptr, const int *__restrict__ col,
                     const T *__restrict__ val, const T *__restrict__ x, const T *__restrict__ b, const T *__restrict__ dw, T *y,
                     T *y2) {
    const long long row = (long long)(row_begin & ~31) + (long long)blockIdx.x * ROW_THREADS + threadIdx.x;
    const int last_window_end = min(n_total, (row_end + 31) & ~31);
    const bool present = row < last_window_end;              // a row of a window this launch touches
    const bool active = row >= row_begin && row < row_end;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    int start = 0, len = 0;
    if (present) {
        start = rowptr[row];
        len = rowptr[row + 1] - start;
    }
    int off = __shfl_sync(0xffffffffu, start, 0);            // first entry of the window
    const int mylen = active ? len : 0;                      // rows outside the range load nothing
    T sum = (T)0;
    constexpr int NB = 4;
    for (int k = 0;; k += NB) {
        unsigned m[NB];
#pragma unroll
        for (int j = 0; j < NB; j++) m[j] = __ballot_sync(0xffffffffu, k + j < len);
        if (m[0] == 0u) break;
        int idx[NB];
        int c[NB];
        T v[NB], xv[NB];
        int o = off;
#pragma unroll
        for (int j = 0; j < NB; j++) {
            idx[j] = o + __popc(m[j] & lt);
            o += __popc(m[j]);
        }
        off = o;
#pragma unroll
        for (int j = 0; j < NB; j++) c[j] = (k + j < mylen) ? col[idx[j]] : 0;
#pragma unroll
        for (int j = 0; j < NB; j++) v[j] = (k + j < mylen) ? val[idx[j]] : (T)0;
#pragma unroll
        for (int j = 0; j < NB; j++) xv[j] = (k + j < mylen) ? x[c[j]] : (T)0;
#pragma unroll
        for (int j = 0; j < NB; j++) sum += v[j] * xv[j];
    }
    if (active) row_epilogue<T, OP, false>(row, sum, x, b, dw, y, y2);
}

// CSR -> W32: same index arithmetic, entries scattered to their slot-major position
template <typename T>
__global__ void __launch_bounds__(ROW_THREADS)
csr_to_w32_kernel(int n, const int *__restrict__ rowptr, const int *__restrict__ col, const T *__restrict__ val,
                  int *__restrict__ col_out, T *__restrict__ val_out) {
    const long long gtid = (long long)blockIdx.x * ROW_THREADS + threadIdx.x;
    const bool valid = gtid < n;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    int start = 0, len = 0;
    if (valid) {
        start = rowptr[gtid];
        len = rowptr[gtid + 1] - start;
    }
    int off = __shfl_sync(0xffffffffu, start, 0);
    for (int k = 0;; k++) {
        const unsigned m = __ballot_sync(0xffffffffu, k < len);
        if (m == 0u) break;
        if (k < len) {
            const int idx = off + __popc(m & lt);
            col_out[idx] = col[start + k];
            val_out[idx] = val[start + k];
        }
        off += __popc(m);
    }
}

static inline unsigned w32_blocks(int row_begin, int row_end, int n_total) {
    const int first = row_begin & ~31;
    int last = (row_end + 31) & ~31;
    if (last > n_total) last = n_total;
    return cdiv(last - first, ROW_THREADS);
}

// generic W32 row-op over rows [row0, row0 + nrows) of an operator with n_total rows (ops: residual, psmooth, psmooth0, spmv)
template <typename T>
int w32_rowop_t(int op, int nrows, int row0, int n_total, const int *rowptr, const int *col_w32, const T *val_w32, const T *x,
                const T *b, const T *dw, T *y, T *aux, cudaStream_t s) {
    if (nrows <= 0) return MLAMG_OK;
    if (row0 < 0 || row0 + nrows > n_total) return set_error(MLAMG_EINVAL, "w32 row-op: row range outside the operator");
    const unsigned blocks = w32_blocks(row0, row0 + nrows, n_total);
#define W32_LAUNCH(OPC) \
    csr_w32_rowop_kernel<T, OPC><<<blocks, ROW_THREADS, 0, s>>>(row0, row0 + nrows, n_total, rowptr, col_w32, val_w32, x, b, dw, y, aux)
    switch (op) {
        case OP_SPMV: W32_LAUNCH(OP_SPMV); break;
        case OP_RESIDUAL: W32_LAUNCH(OP_RESIDUAL); break;
        case OP_PSMOOTH: W32_LAUNCH(OP_PSMOOTH); break;
        case OP_PSMOOTH0: W32_LAUNCH(OP_PSMOOTH0); break;
        default: return set_error(MLAMG_EINVAL, "w32 row-op: op %d is not built for this layout", op);
    }
#undef W32_LAUNCH
    MLAMG_LAUNCHED();
    return MLAMG_OK;
}
template int w32_rowop_t<float>(int, int, int, int, const int *, const int *, const float *, const float *, const float *, const float *, float *, float *, cudaStream_t);
template int w32_rowop_t<double>(int, int, int, int, const int *, const int *, const double *, const double *, const double *, const double *, double *, double *, cudaStream_t);

template <typename T>
int w32_psmooth0_range_t(int nrows, int row0, int n_total, const int *rowptr, const int *col_w32, const T *val_w32, const T *e,
                         const T *rhs, const T *r, const T *dw, T *x_out, cudaStream_t s) {
    return w32_rowop_t<T>(OP_PSMOOTH0, nrows, row0, n_total, rowptr, col_w32, val_w32, e, r, dw, x_out, const_cast<T *>(rhs), s);
}
// r = b - A x on the W32 copies (the scaled-copy residual of the cycle: x = b)
template <typename T>
int w32_residual_range_t(int nrows, int row0, int n_total, const int *rowptr, const int *col_w32, const T *val_w32, const T *x,
                         const T *b, T *r, cudaStream_t s) {
    return w32_rowop_t<T>(OP_RESIDUAL, nrows, row0, n_total, rowptr, col_w32, val_w32, x, b, nullptr, r, nullptr, s);
}
template int w32_residual_range_t<float>(int, int, int, const int *, const int *, const float *, const float *, const float *, float *, cudaStream_t);
template int w32_residual_range_t<double>(int, int, int, const int *, const int *, const double *, const double *, const double *, double *, cudaStream_t);
template int w32_psmooth0_range_t<float>(int, int, int, const int *, const int *, const float *, const float *, const float *, const float *, const float *, float *, cudaStream_t);
template int w32_psmooth0_range_t<double>(int, int, int, const int *, const int *, const double *, const double *, const double *, const double *, const double *, double *, cudaStream_t);

// SELL-32: rows are grouped in slices of 32; a slice stores width = max row length columns, column-
// major inside the slice (entry k of lane l at slice_ptr[s] + 32 k + l), padded with col = -1.  One
// thread per row, every load of col/val is a fully coalesced 128/256-byte warp transaction and the
// loop carries no reduction across lanes.
template <typename T, int OP, bool NORM>
__global__ void __launch_bounds__(ROW_THREADS)
sell_rowop_kernel(int n, const int *__restrict__ slice_ptr, const int *__restrict__ col, const T *__restrict__ val,
                  const T *__restrict__ x, const T *__restrict__ b, const T *__restrict__ dw, T *__restrict__ y,
                  double *__restrict__ partial) {
    const long long row = (long long)blockIdx.x * ROW_THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31;
    T sum = (T)0;
    if (row < n) {
        const long long slice = row >> 5;
        const int base = slice_ptr[slice];
        const int width = (slice_ptr[slice + 1] - base) >> 5;
        const int *c = col + base + lane;
        const T *v = val + base + lane;
        for (int k = 0; k < width; k += 4) {
            const bool p1 = k + 1 < width, p2 = k + 2 < width, p3 = k + 3 < width;
            const int c0 = c[32 * k];
            const int c1 = p1 ? c[32 * (k + 1)] : -1;
            const int c2 = p2 ? c[32 * (k + 2)] : -1;
            const int c3 = p3 ? c[32 * (k + 3)] : -1;
            const T v0 = v[32 * k];
            const T v1 = p1 ? v[32 * (k + 1)] : (T)0;
            const T v2 = p2 ? v[32 * (k + 2)] : (T)0;
            const T v3 = p3 ? v[32 * (k + 3)] : (T)0;
            const T x0 = c0 >= 0 ? x[c0] : (T)0;
            const T x1 = c1 >= 0 ? x[c1] : (T)0;
            const T x2 = c2 >= 0 ? x[c2] : (T)0;
            const T x3 = c3 >= 0 ? x[c3] : (T)0;
            sum += v0 * x0;
            sum += v1 * x1;
            sum += v2 * x2;
            sum += v3 * x3;
        }
    }
    double rr = 0.0;
    if (row < n) rr = row_epilogue<T, OP, NORM>(row, sum, x, b, dw, y, (T *)nullptr);
    if (NORM) {
        __shared__ double sm[32];
        rr = block_sum(rr, sm);
        if (threadIdx.x == 0) partial[blockIdx.x] = rr;
    }
}

static int g_force_lanes = -1;   // test / tuning hook (mlamg_set_csr_lanes), -1 = heuristic, 0 = TMA-staged kernel
static int g_force_batch = 0;    // tuning hook (mlamg_set_csr_batch): entries per lane and loop trip (2, 4, 8), 0 = default
// the heuristic never picks the staged kernel: measured at 256^3 it is slower than the plain thread-per-row
// kernel (fine Jacobi sweep 436 vs 290 us, prolongation 252 vs 181 us — the two CTA barriers serialise the load
// and gather phases and cost more memory-level parallelism than the single-touch loads save); kept as variant 0
static int g_staged = 0;

static int pick_lanes(int n, long long nnz) {
    if (g_force_lanes >= 0) return g_force_lanes;
    // ~8-12 entries per lane: enough independent loads per thread to cover HBM latency
    const double mean = n > 0 ? (double)nnz / (double)n : 0.0;
    int lanes = 32;
    if (mean <= 12.0) lanes = (g_staged && (long long)n >= 148LL * 1024) ? 0 : 1;
    else if (mean <= 24.0) lanes = 2;
    else if (mean <= 48.0) lanes = 4;
    else if (mean <= 128.0) lanes = 8;
    else if (mean <= 384.0) lanes = 16;
    // small levels: not enough rows to fill 148 SMs -> spend more lanes per row (up to the row length)
    while (lanes > 0 && lanes < 32 && (long long)n * lanes < 148LL * 1024 && 2.0 * lanes <= mean) lanes *= 2;
    return lanes;
}

template <typename T, int OP, bool NORM>
static int launch_rowop(int n, long long nnz_hint, const int *rowptr, const int *col, const T *val, const T *x,
                        const T *b, const T *dw, T *y, double *norm2, cudaStream_t s, const int *row_order = nullptr,
                        int row0 = 0, const HaloLL *halo = nullptr, T *y2 = nullptr) {
    if (n <= 0) {
        if (NORM && norm2) MLAMG_CUDA(cudaMemsetAsync(norm2, 0, sizeof(double), s));
        return MLAMG_OK;
    }
    int lanes = pick_lanes(n, nnz_hint);
    if (lanes == 0 && (row_order || halo)) lanes = 1;      // the staged kernel walks a row RANGE without halo columns
    const unsigned blocks = cdiv((long long)n * (lanes ? lanes : 1), ROW_THREADS);
    double *partial = nullptr;
    Scratch part(NORM ? (size_t)blocks * sizeof(double) : (size_t)-1, s);
    if (NORM) {
        MLAMG_SCRATCH_OK(part);
        partial = part.as<double>();
    }
    HaloLL hl;
    memset(&hl, 0, sizeof(hl));
    if (halo) hl = *halo;
    if (halo && NORM) return set_error(MLAMG_EINVAL, "rowop: no norm on the in-place halo variant");
    constexpr int NB_DEFAULT = (OP == OP_RESZERO) ? 2 : 4;
    const int nb = (g_force_batch > 0 && !halo) ? g_force_batch : NB_DEFAULT;
#define LAUNCH_NB(L, B)                                                                                              \
    csr_rowop_kernel<T, L, OP, NORM, false, B><<<blocks, ROW_THREADS, 0, s>>>(n, rowptr, col, val, x, b, dw, y,      \
                                                                              partial, row_order, row0, hl, y2)
#define LAUNCH(L)                                                                                                    \
    do {                                                                                                             \
        if (halo)                                                                                                    \
            csr_rowop_kernel<T, L, OP, false, true, NB_DEFAULT><<<blocks, ROW_THREADS, 0, s>>>(                      \
                n, rowptr, col, val, x, b, dw, y, nullptr, row_order, row0, hl, y2);                                 \
        else if (nb == 2) LAUNCH_NB(L, 2);                                                                           \
        else if (nb == 8) LAUNCH_NB(L, 8);                                                                           \
        else LAUNCH_NB(L, 4);                                                                                        \
    } while (0)
    switch (lanes) {
        case 0:
            csr_tma_rowop_kernel<T, OP, NORM><<<blocks, ROW_THREADS, 0, s>>>(n, rowptr, col, val, x, b, dw, y, partial, row0, y2);
            break;
        case 1: LAUNCH(1); break;
        case 2: LAUNCH(2); break;
        case 4: LAUNCH(4); break;
        case 8: LAUNCH(8); break;
        case 16: LAUNCH(16); break;
        default: LAUNCH(32); break;
    }
#undef LAUNCH
#undef LAUNCH_NB
    MLAMG_LAUNCHED();
    if (NORM) return reduce_partials(partial, (int)blocks, norm2, s);
    return MLAMG_OK;
}

template <typename T, int OP, bool NORM>
static int launch_sell(int n, const int *slice_ptr, const int *col, const T *val, const T *x, const T *b, const T *dw,
                       T *y, double *norm2, cudaStream_t s) {
    if (n <= 0) {
        if (NORM && norm2) MLAMG_CUDA(cudaMemsetAsync(norm2, 0, sizeof(double), s));
        return MLAMG_OK;
    }
    const unsigned blocks = cdiv(n, ROW_THREADS);
    double *partial = nullptr;
    Scratch part(NORM ? (size_t)blocks * sizeof(double) : (size_t)-1, s);
    if (NORM) {
        MLAMG_SCRATCH_OK(part);
        partial = part.as<double>();
    }
    sell_rowop_kernel<T, OP, NORM><<<blocks, ROW_THREADS, 0, s>>>(n, slice_ptr, col, val, x, b, dw, y, partial);
    MLAMG_LAUNCHED();
    if (NORM) return reduce_partials(partial, (int)blocks, norm2, s);
    return MLAMG_OK;
}

// op: 0 spmv, 1 spmv_add, 2 residual (+norm2), 3 jacobi
template <typename T>
int sell_rowop_t(int op, int n, const int *slice_ptr, const int *col, const T *val, const T *x, const T *b,
                 const T *dw, T *y, double *norm2, cudaStream_t s) {
    switch (op) {
        case OP_SPMV: return launch_sell<T, OP_SPMV, false>(n, slice_ptr, col, val, x, b, dw, y, nullptr, s);
        case OP_SPMV_ADD: return launch_sell<T, OP_SPMV_ADD, false>(n, slice_ptr, col, val, x, b, dw, y, nullptr, s);
        case OP_RESIDUAL:
            if (norm2) return launch_sell<T, OP_RESIDUAL, true>(n, slice_ptr, col, val, x, b, dw, y, norm2, s);
            return launch_sell<T, OP_RESIDUAL, false>(n, slice_ptr, col, val, x, b, dw, y, nullptr, s);
        case OP_JACOBI: return launch_sell<T, OP_JACOBI, false>(n, slice_ptr, col, val, x, b, dw, y, nullptr, s);
        default: return set_error(MLAMG_EINVAL, "sell_rowop: bad op %d", op);
    }
}
template int sell_rowop_t<float>(int, int, const int *, const int *, const float *, const float *, const float *,
                                 const float *, float *, double *, cudaStream_t);
template int sell_rowop_t<double>(int, int, const int *, const int *, const double *, const double *, const double *,
                                  const double *, double *, double *, cudaStream_t);

// ---- CSR -> SELL-32 conversion ------------------------------------------------------------------
__global__ void __launch_bounds__(256) sell_width_kernel(int n, const int *__restrict__ rowptr, int *__restrict__ sizes) {
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int len = 0;
    if (row < n) len = rowptr[row + 1] - rowptr[row];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
    if ((threadIdx.x & 31) == 0 && row < n) sizes[row >> 5] = len * 32;
}

template <typename T>
__global__ void __launch_bounds__(256) sell_fill_kernel(int n, const int *__restrict__ rowptr, const int *__restrict__ col,
                                                        const T *__restrict__ val, const int *__restrict__ slice_ptr,
                                                        int *__restrict__ scol, T *__restrict__ sval) {
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const long long slice = row >> 5;
    if (slice * 32 >= n) return;   // whole warp
    const int base = slice_ptr[slice];
    const int width = (slice_ptr[slice + 1] - base) >> 5;
    int start = 0, len = 0;
    if (row < n) { start = rowptr[row]; len = rowptr[row + 1] - start; }
    for (int k = 0; k < width; k++) {
        const bool live = k < len;
        scol[base + 32 * k + lane] = live ? col[start + k] : -1;
        sval[base + 32 * k + lane] = live ? val[start + k] : (T)0;
    }
}

// ---- internal entry points used by hierarchy.cu (nnz known, no host sync) -------------------
template <typename T>
int spmv_t(int n, long long nnz, const int *rowptr, const int *col, const T *val, const T *x, T *y, cudaStream_t s) {
    return launch_rowop<T, OP_SPMV, false>(n, nnz, rowptr, col, val, x, nullptr, nullptr, y, nullptr, s);
}
template <typename T>
int spmv_perm_t(int n, long long nnz, const int *rowptr, const int *col, const T *val, const T *x, T *y,
                const int *row_order, cudaStream_t s) {
    return launch_rowop<T, OP_SPMV, false>(n, nnz, rowptr, col, val, x, nullptr, nullptr, y, nullptr, s, row_order);
}
template int spmv_perm_t<float>(int, long long, const int *, const int *, const float *, const float *, float *, const int *, cudaStream_t);
template int spmv_perm_t<double>(int, long long, const int *, const int *, const double *, const double *, double *, const int *, cudaStream_t);
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

// x_out = dw .* b ; r = b - A x_out (+ *norm2 = ||r||^2): zero-guess first sweep fused with the residual
template <typename T>
int reszero_t(int n, long long nnz, const int *rowptr, const int *col, const T *val, const T *dw, const T *b, T *x_out,
              T *r, double *norm2, cudaStream_t s) {
    if (norm2)
        return launch_rowop<T, OP_RESZERO, true>(n, nnz, rowptr, col, val, nullptr, b, dw, r, norm2, s, nullptr, 0, nullptr, x_out);
    return launch_rowop<T, OP_RESZERO, false>(n, nnz, rowptr, col, val, nullptr, b, dw, r, nullptr, s, nullptr, 0, nullptr, x_out);
}
// x_out = x_in + dw .* r + Q e   (x_out may alias x_in)
template <typename T>
int psmooth_t(int n, long long nnz, const int *rowptr, const int *col, const T *val, const T *e, const T *x_in, const T *r,
              const T *dw, T *x_out, cudaStream_t s) {
    return launch_rowop<T, OP_PSMOOTH, false>(n, nnz, rowptr, col, val, e, r, dw, x_out, nullptr, s, nullptr, 0, nullptr,
                                              const_cast<T *>(x_in));
}
// x_out = dw .* (rhs + r) + Q e
template <typename T>
int psmooth0_range_t(int nrows, int row0, long long nnz_hint, const int *rowptr, const int *col, const T *val, const T *e,
                     const T *rhs, const T *r, const T *dw, T *x_out, cudaStream_t s) {
    return launch_rowop<T, OP_PSMOOTH0, false>(nrows, nnz_hint, rowptr, col, val, e, r, dw, x_out, nullptr, s, nullptr, row0,
                                               nullptr, const_cast<T *>(rhs));
}
template int psmooth0_range_t<float>(int, int, long long, const int *, const int *, const float *, const float *, const float *, const float *, const float *, float *, cudaStream_t);
template int psmooth0_range_t<double>(int, int, long long, const int *, const int *, const double *, const double *, const double *, const double *, const double *, double *, cudaStream_t);
// r = b - A x over the row range (the scaled-copy residual of the chunked host pipeline)
template <typename T>
int residual_range_t(int nrows, int row0, long long nnz_hint, const int *rowptr, const int *col, const T *val, const T *x,
                     const T *b, T *r, cudaStream_t s) {
    return launch_rowop<T, OP_RESIDUAL, false>(nrows, nnz_hint, rowptr, col, val, x, b, nullptr, r, nullptr, s, nullptr, row0);
}
template int residual_range_t<float>(int, int, long long, const int *, const int *, const float *, const float *, const float *, float *, cudaStream_t);
template int residual_range_t<double>(int, int, long long, const int *, const int *, const double *, const double *, const double *, double *, cudaStream_t);

// row-range forms (rows [row0, row0 + nrows)) used by the host-buffer entry point to pipeline the first and the last
// fine-level pass with the PCIe copies
template <typename T>
int psmooth_range_t(int nrows, int row0, long long nnz_hint, const int *rowptr, const int *col, const T *val, const T *e,
                    const T *x_in, const T *r, const T *dw, T *x_out, cudaStream_t s) {
    return launch_rowop<T, OP_PSMOOTH, false>(nrows, nnz_hint, rowptr, col, val, e, r, dw, x_out, nullptr, s, nullptr, row0,
                                              nullptr, const_cast<T *>(x_in));
}
template <typename T>
int reszero_scaled_range_t(int nrows, int row0, long long nnz_hint, const int *rowptr, const int *col, const T *val_scaled,
                           const T *dw, const T *b, T *x_out, T *r, cudaStream_t s) {
    return launch_rowop<T, OP_RESZERO_S, false>(nrows, nnz_hint, rowptr, col, val_scaled, nullptr, b, dw, r, nullptr, s, nullptr,
                                                row0, nullptr, x_out);
}
template int psmooth_range_t<float>(int, int, long long, const int *, const int *, const float *, const float *, const float *, const float *, const float *, float *, cudaStream_t);
template int psmooth_range_t<double>(int, int, long long, const int *, const int *, const double *, const double *, const double *, const double *, const double *, double *, cudaStream_t);
template int reszero_scaled_range_t<float>(int, int, long long, const int *, const int *, const float *, const float *, const float *, float *, float *, cudaStream_t);
template int reszero_scaled_range_t<double>(int, int, long long, const int *, const int *, const double *, const double *, const double *, double *, double *, cudaStream_t);

template int psmooth_t<float>(int, long long, const int *, const int *, const float *, const float *, const float *, const float *, const float *, float *, cudaStream_t);
template int psmooth_t<double>(int, long long, const int *, const int *, const double *, const double *, const double *, const double *, const double *, double *, cudaStream_t);

// the same on the column-scaled values a_ij * dw_j (single gather of b)
template <typename T>
int reszero_scaled_t(int n, long long nnz, const int *rowptr, const int *col, const T *val_scaled, const T *dw, const T *b,
                     T *x_out, T *r, double *norm2, cudaStream_t s) {
    if (norm2)
        return launch_rowop<T, OP_RESZERO_S, true>(n, nnz, rowptr, col, val_scaled, nullptr, b, dw, r, norm2, s, nullptr, 0,
                                                   nullptr, x_out);
    return launch_rowop<T, OP_RESZERO_S, false>(n, nnz, rowptr, col, val_scaled, nullptr, b, dw, r, nullptr, s, nullptr, 0,
                                                nullptr, x_out);
}
template int reszero_scaled_t<float>(int, long long, const int *, const int *, const float *, const float *, const float *, float *, float *, double *, cudaStream_t);
template int reszero_scaled_t<double>(int, long long, const int *, const int *, const double *, const double *, const double *, double *, double *, double *, cudaStream_t);

template int reszero_t<float>(int, long long, const int *, const int *, const float *, const float *, const float *, float *, float *, double *, cudaStream_t);
template int reszero_t<double>(int, long long, const int *, const int *, const double *, const double *, const double *, double *, double *, double *, cudaStream_t);

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

// halo pack: dst[i] = src[idx[i]]
template <typename T>
__global__ void __launch_bounds__(256) gather_kernel(int n, const int *__restrict__ idx, const T *__restrict__ src,
                                                     T *__restrict__ dst) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
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

int mlamg_spmv_csr_perm(int dtype, int n, int nnz, const int *rowptr, const int *col, const void *val, const void *x,
                        void *y, const int *row_order, mlamg_stream_t stream) {
    cudaStream_t s = as_stream(stream);
    if (n < 0) return set_error(MLAMG_EINVAL, "spmv_perm: n < 0");
    if (x == y) return set_error(MLAMG_EINVAL, "spmv_perm: x aliases y");
    MLAMG_DISPATCH(dtype, return spmv_perm_t<T>(n, nnz, rowptr, col, (const T *)val, (const T *)x, (T *)y, row_order, s));
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

int mlamg_jacobi_zero_residual_csr(int dtype, int n, int nnz, const int *rowptr, const int *col, const void *val,
                                   const void *dw, const void *b, void *x_out, void *r, double *norm2, mlamg_stream_t stream) {
    cudaStream_t s = as_stream(stream);
    if (n < 0) return set_error(MLAMG_EINVAL, "jacobi_zero_residual: n < 0");
    if (x_out == r || b == x_out || b == r) return set_error(MLAMG_EINVAL, "jacobi_zero_residual: aliased arguments");
    MLAMG_DISPATCH(dtype, return reszero_t<T>(n, nnz, rowptr, col, (const T *)val, (const T *)dw, (const T *)b, (T *)x_out,
                                              (T *)r, norm2, s));
    return MLAMG_OK;
}

int mlamg_jacobi_zero_residual_scaled_csr(int dtype, int n, int nnz, const int *rowptr, const int *col,
                                          const void *val_scaled, const void *dw, const void *b, void *x_out, void *r,
                                          double *norm2, mlamg_stream_t stream) {
    cudaStream_t s = as_stream(stream);
    if (n < 0) return set_error(MLAMG_EINVAL, "jacobi_zero_residual_scaled: n < 0");
    if (x_out == r || b == x_out || b == r) return set_error(MLAMG_EINVAL, "jacobi_zero_residual_scaled: aliased arguments");
    MLAMG_DISPATCH(dtype, return reszero_scaled_t<T>(n, nnz, rowptr, col, (const T *)val_scaled, (const T *)dw, (const T *)b,
                                                     (T *)x_out, (T *)r, norm2, s));
    return MLAMG_OK;
}

int mlamg_prolong_smooth_zero_csr(int dtype, int n, int nnz, const int *rowptr, const int *col, const void *val,
                                  const void *e, const void *rhs, const void *r, const void *dw, void *x_out,
                                  mlamg_stream_t stream) {
    cudaStream_t s = as_stream(stream);
    if (n < 0) return set_error(MLAMG_EINVAL, "prolong_smooth_zero: n < 0");
    if (e == x_out || r == x_out || rhs == x_out) return set_error(MLAMG_EINVAL, "prolong_smooth_zero: aliased arguments");
    MLAMG_DISPATCH(dtype, return psmooth0_range_t<T>(n, 0, nnz, rowptr, col, (const T *)val, (const T *)e, (const T *)rhs,
                                                     (const T *)r, (const T *)dw, (T *)x_out, s));
    return MLAMG_OK;
}

int mlamg_csr_to_w32(int dtype, int n, const int *rowptr, const int *col, const void *val, int *col_out, void *val_out,
                     mlamg_stream_t stream) {
    cudaStream_t s = as_stream(stream);
    if (n < 0) return set_error(MLAMG_EINVAL, "csr_to_w32: n < 0");
    if (n == 0) return MLAMG_OK;
    if (col == col_out || val == val_out) return set_error(MLAMG_EINVAL, "csr_to_w32: in-place conversion is not supported");
    MLAMG_DISPATCH(dtype, (csr_to_w32_kernel<T><<<cdiv(n, ROW_THREADS), ROW_THREADS, 0, s>>>(n, rowptr, col, (const T *)val, col_out,
                                                                                            (T *)val_out)));
    MLAMG_LAUNCHED();
    return MLAMG_OK;
}

int mlamg_prolong_smooth_zero_w32(int dtype, int n, const int *rowptr, const int *col_w32, const void *val_w32, const void *e,
                                  const void *rhs, const void *r, const void *dw, void *x_out, mlamg_stream_t stream) {
    if (n < 0) return set_error(MLAMG_EINVAL, "prolong_smooth_zero_w32: n < 0");
    if (e == x_out || r == x_out || rhs == x_out) return set_error(MLAMG_EINVAL, "prolong_smooth_zero_w32: aliased arguments");
    MLAMG_DISPATCH(dtype, return w32_psmooth0_range_t<T>(n, 0, n, rowptr, col_w32, (const T *)val_w32, (const T *)e, (const T *)rhs,
                                                         (const T *)r, (const T *)dw, (T *)x_out, as_stream(stream)));
    return MLAMG_OK;
}

int mlamg_rowop_w32(int dtype, int op, int nrows, int row0, int n_total, const int *rowptr, const int *col_w32, const void *val_w32,
                    const void *x, const void *b, const void *dw, void *y, void *aux, mlamg_stream_t stream) {
    MLAMG_DISPATCH(dtype, return w32_rowop_t<T>(op, nrows, row0, n_total, rowptr, col_w32, (const T *)val_w32, (const T *)x,
                                                (const T *)b, (const T *)dw, (T *)y, (T *)aux, as_stream(stream)));
    return MLAMG_OK;
}

int mlamg_residual_w32(int dtype, int n, const int *rowptr, const int *col_w32, const void *val_w32, const void *x, const void *b,
                       void *r, mlamg_stream_t stream) {
    if (n < 0) return set_error(MLAMG_EINVAL, "residual_w32: n < 0");
    if (x == r || b == r) return set_error(MLAMG_EINVAL, "residual_w32: aliased arguments");
    MLAMG_DISPATCH(dtype, return w32_residual_range_t<T>(n, 0, n, rowptr, col_w32, (const T *)val_w32, (const T *)x, (const T *)b,
                                                         (T *)r, as_stream(stream)));
    return MLAMG_OK;
}

int mlamg_prolong_smooth_csr(int dtype, int n, int nnz, const int *rowptr, const int *col, const void *val, const void *e,
                             const void *x_in, const void *r, const void *dw, void *x_out, mlamg_stream_t stream) {
    cudaStream_t s = as_stream(stream);
    if (n < 0) return set_error(MLAMG_EINVAL, "prolong_smooth: n < 0");
    if (e == x_out || r == x_out) return set_error(MLAMG_EINVAL, "prolong_smooth: aliased arguments");
    MLAMG_DISPATCH(dtype, return psmooth_t<T>(n, nnz, rowptr, col, (const T *)val, (const T *)e, (const T *)x_in, (const T *)r,
                                              (const T *)dw, (T *)x_out, s));
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

// generic row-op entry: op 0 spmv | 1 spmv_add | 2 residual(+norm2) | 3 jacobi, over all rows
// (row_list == NULL, nrows = n) or over the subset row_list[0..nrows) (interior / boundary splits of
// the row-partitioned multi-GPU levels: interior rows run while the halo exchange is in flight).
int mlamg_rowop_csr(int dtype, int op, int nrows, int nnz_hint, const int *rowptr, const int *col, const void *val,
                    const void *x, const void *b, const void *dw, void *y, void *aux, const int *row_list, int row_begin,
                    double *norm2, mlamg_stream_t stream) {
    cudaStream_t s = as_stream(stream);
    if (nrows < 0) return set_error(MLAMG_EINVAL, "rowop: nrows < 0");
    if (x == y && x) return set_error(MLAMG_EINVAL, "rowop: x aliases y");
#define ROWOP_CASE(OPC, NRM) \
    MLAMG_DISPATCH(dtype, return (launch_rowop<T, OPC, NRM>(nrows, nnz_hint, rowptr, col, (const T *)val, (const T *)x, \
                                                             (const T *)b, (const T *)dw, (T *)y, norm2, s, row_list, row_begin)))
    switch (op) {
        case OP_SPMV: ROWOP_CASE(OP_SPMV, false); break;
        case OP_SPMV_ADD: ROWOP_CASE(OP_SPMV_ADD, false); break;
        case OP_RESIDUAL:
            if (norm2) { ROWOP_CASE(OP_RESIDUAL, true); } else { ROWOP_CASE(OP_RESIDUAL, false); }
            break;
        case OP_JACOBI: ROWOP_CASE(OP_JACOBI, false); break;
        case OP_RESZERO:       // aux = x_out (x = dw.*b is produced, not read)
            if (!aux || aux == y) return set_error(MLAMG_EINVAL, "rowop: op 4 needs aux = x_out");
            MLAMG_DISPATCH(dtype, return (launch_rowop<T, OP_RESZERO, false>(nrows, nnz_hint, rowptr, col, (const T *)val, nullptr,
                                                                             (const T *)b, (const T *)dw, (T *)y, nullptr, s, row_list,
                                                                             row_begin, nullptr, (T *)aux)));
            break;
        case OP_RESZERO_S:     // op 4 on column-scaled values a_ij*dw_j: gathers b alone
            if (!aux || aux == y) return set_error(MLAMG_EINVAL, "rowop: op 6 needs aux = x_out");
            MLAMG_DISPATCH(dtype, return (launch_rowop<T, OP_RESZERO_S, false>(nrows, nnz_hint, rowptr, col, (const T *)val, nullptr,
                                                                               (const T *)b, (const T *)dw, (T *)y, nullptr, s, row_list,
                                                                               row_begin, nullptr, (T *)aux)));
            break;
        case OP_PSMOOTH0:      // aux = right-hand side; x = coarse correction, b = residual of dw.*rhs
            if (!aux || aux == y) return set_error(MLAMG_EINVAL, "rowop: op 7 needs aux = rhs");
            MLAMG_DISPATCH(dtype, return (launch_rowop<T, OP_PSMOOTH0, false>(nrows, nnz_hint, rowptr, col, (const T *)val, (const T *)x,
                                                                              (const T *)b, (const T *)dw, (T *)y, nullptr, s, row_list,
                                                                              row_begin, nullptr, (T *)aux)));
            break;
        case OP_PSMOOTH:       // aux = x_in (may alias y); x = coarse correction, b = residual
            if (!aux) return set_error(MLAMG_EINVAL, "rowop: op 5 needs aux = x_in");
            MLAMG_DISPATCH(dtype, return (launch_rowop<T, OP_PSMOOTH, false>(nrows, nnz_hint, rowptr, col, (const T *)val, (const T *)x,
                                                                             (const T *)b, (const T *)dw, (T *)y, nullptr, s, row_list,
                                                                             row_begin, nullptr, (T *)aux)));
            break;
        default: return set_error(MLAMG_EINVAL, "rowop: bad op %d", op);
    }
#undef ROWOP_CASE
    return MLAMG_OK;
}

// row-op whose gathers of columns >= n_own read the channel's receive region in place (csrc/peer.cu)
int mlamg_channel_rowop(mlamg_channel_t ch, int dtype, int op, int nrows, int nnz_hint, const int *rowptr, const int *col,
                        const void *val, const void *x, int n_own, const void *b, const void *dw, void *y, void *aux,
                        const int *row_list, int row_begin, mlamg_stream_t stream) {
    cudaStream_t s = as_stream(stream);
    if (!ch) return set_error(MLAMG_EINVAL, "channel_rowop: null channel");
    if (nrows < 0 || n_own < 0) return set_error(MLAMG_EINVAL, "channel_rowop: bad nrows/n_own");
    if (x == y) return set_error(MLAMG_EINVAL, "channel_rowop: x aliases y");
    HaloLL hl;
    hl.n_own = n_own;
    hl.region[0] = ch->dev.recv_region[0];
    hl.region[1] = ch->dev.recv_region[1];
    hl.state = ch->dev.state;
#define ROWOP_CASE(OPC) \
    MLAMG_DISPATCH(dtype, return (launch_rowop<T, OPC, false>(nrows, nnz_hint, rowptr, col, (const T *)val, (const T *)x, \
                                                               (const T *)b, (const T *)dw, (T *)y, nullptr, s, row_list, row_begin, &hl)))
    switch (op) {
        case OP_SPMV: ROWOP_CASE(OP_SPMV); break;
        case OP_SPMV_ADD: ROWOP_CASE(OP_SPMV_ADD); break;
        case OP_RESIDUAL: ROWOP_CASE(OP_RESIDUAL); break;
        case OP_JACOBI: ROWOP_CASE(OP_JACOBI); break;
        case OP_RESZERO:       // aux = x_out; halo columns carry the neighbours' dw.*b
            if (!aux || aux == y) return set_error(MLAMG_EINVAL, "channel_rowop: op 4 needs aux = x_out");
            MLAMG_DISPATCH(dtype, return (launch_rowop<T, OP_RESZERO, false>(nrows, nnz_hint, rowptr, col, (const T *)val, nullptr,
                                                                             (const T *)b, (const T *)dw, (T *)y, nullptr, s, row_list,
                                                                             row_begin, &hl, (T *)aux)));
            break;
        case OP_RESZERO_S:     // halo columns carry the neighbours' b
            if (!aux || aux == y) return set_error(MLAMG_EINVAL, "channel_rowop: op 6 needs aux = x_out");
            MLAMG_DISPATCH(dtype, return (launch_rowop<T, OP_RESZERO_S, false>(nrows, nnz_hint, rowptr, col, (const T *)val, nullptr,
                                                                               (const T *)b, (const T *)dw, (T *)y, nullptr, s, row_list,
                                                                               row_begin, &hl, (T *)aux)));
            break;
        case OP_PSMOOTH0:      // aux = right-hand side
            if (!aux || aux == y) return set_error(MLAMG_EINVAL, "channel_rowop: op 7 needs aux = rhs");
            MLAMG_DISPATCH(dtype, return (launch_rowop<T, OP_PSMOOTH0, false>(nrows, nnz_hint, rowptr, col, (const T *)val, (const T *)x,
                                                                              (const T *)b, (const T *)dw, (T *)y, nullptr, s, row_list,
                                                                              row_begin, &hl, (T *)aux)));
            break;
        case OP_PSMOOTH:       // aux = x_in (may alias y)
            if (!aux) return set_error(MLAMG_EINVAL, "channel_rowop: op 5 needs aux = x_in");
            MLAMG_DISPATCH(dtype, return (launch_rowop<T, OP_PSMOOTH, false>(nrows, nnz_hint, rowptr, col, (const T *)val, (const T *)x,
                                                                             (const T *)b, (const T *)dw, (T *)y, nullptr, s, row_list,
                                                                             row_begin, &hl, (T *)aux)));
            break;
        default: return set_error(MLAMG_EINVAL, "channel_rowop: bad op %d", op);
    }
#undef ROWOP_CASE
    return MLAMG_OK;
}

int mlamg_gather(int dtype, int n, const int *idx, const void *src, void *dst, mlamg_stream_t stream) {
    cudaStream_t s = as_stream(stream);
    if (n <= 0) return n == 0 ? MLAMG_OK : set_error(MLAMG_EINVAL, "gather: n < 0");
    unsigned blocks = cdiv(n, 256);
    MLAMG_DISPATCH(dtype, (gather_kernel<T><<<blocks, 256, 0, s>>>(n, idx, (const T *)src, (T *)dst)));
    MLAMG_LAUNCHED();
    return MLAMG_OK;
}

int mlamg_set_csr_batch(int nb) {
    if (nb != 0 && nb != 2 && nb != 4 && nb != 8) return set_error(MLAMG_EINVAL, "set_csr_batch: 0, 2, 4 or 8");
    g_force_batch = nb;
    return MLAMG_OK;
}

int mlamg_set_csr_lanes(int lanes) {
    if (lanes == -2 || lanes == -3) {      // -2 / -3: heuristic without / with the staged short-row kernel
        g_force_lanes = -1;
        g_staged = lanes == -3;
        return MLAMG_OK;
    }
    if (lanes != -1 && lanes != 0 && lanes != 1 && lanes != 2 && lanes != 4 && lanes != 8 && lanes != 16 && lanes != 32)
        return set_error(MLAMG_EINVAL, "set_csr_lanes: lanes must be -1 (heuristic), 0 (staged) or a power of two <= 32");
    g_force_lanes = lanes;
    return MLAMG_OK;
}

int mlamg_sell_slice_ptr(int n, const int *rowptr, int *slice_ptr, long long *padded_nnz_host, mlamg_stream_t stream) {
    cudaStream_t s = as_stream(stream);
    if (n < 0) return set_error(MLAMG_EINVAL, "sell_slice_ptr: n < 0");
    const int nslices = (n + 31) / 32;
    if (n > 0) {
        sell_width_kernel<<<cdiv((long long)nslices * 32, 256), 256, 0, s>>>(n, rowptr, slice_ptr);
        MLAMG_LAUNCHED();
    }
    MLAMG_TRY(exclusive_scan_i32(slice_ptr, slice_ptr, nslices, s));
    int h = 0;
    MLAMG_CUDA(cudaMemcpyAsync(&h, slice_ptr + nslices, sizeof(int), cudaMemcpyDeviceToHost, s));
    MLAMG_CUDA(cudaStreamSynchronize(s));
    if (h < 0) return set_error(MLAMG_ELIMIT, "sell: padded size overflows int32");
    if (padded_nnz_host) *padded_nnz_host = h;
    return MLAMG_OK;
}

int mlamg_sell_fill(int dtype, int n, const int *rowptr, const int *col, const void *val, const int *slice_ptr,
                    int *scol, void *sval, mlamg_stream_t stream) {
    cudaStream_t s = as_stream(stream);
    if (n <= 0) return n == 0 ? MLAMG_OK : set_error(MLAMG_EINVAL, "sell_fill: n < 0");
    const int nslices = (n + 31) / 32;
    MLAMG_DISPATCH(dtype, (sell_fill_kernel<T><<<cdiv((long long)nslices * 32, 256), 256, 0, s>>>(
                              n, rowptr, col, (const T *)val, slice_ptr, scol, (T *)sval)));
    MLAMG_LAUNCHED();
    return MLAMG_OK;
}

int mlamg_sell_rowop(int dtype, int op, int n, const int *slice_ptr, const int *scol, const void *sval, const void *x,
                     const void *b, const void *dw, void *y, double *norm2, mlamg_stream_t stream) {
    if (n < 0) return set_error(MLAMG_EINVAL, "sell_rowop: n < 0");
    if (x == y) return set_error(MLAMG_EINVAL, "sell_rowop: x aliases y");
    MLAMG_DISPATCH(dtype, return sell_rowop_t<T>(op, n, slice_ptr, scol, (const T *)sval, (const T *)x, (const T *)b,
                                                 (const T *)dw, (T *)y, norm2, as_stream(stream)));
    return MLAMG_OK;
}

int mlamg_gemv(int dtype, int n, const void *m, const void *x, void *y, mlamg_stream_t stream) {
    MLAMG_DISPATCH(dtype, return gemv_t<T>(n, (const T *)m, (const T *)x, (T *)y, as_stream(stream)));
    return MLAMG_OK;
}

}  // extern "C"
