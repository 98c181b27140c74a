// Lloyd / Bellman-Ford aggregation with the sequential reference's tie-breaking, bit-exact.
//
// Reference: pyamg 4.x amg_core `bellman_ford` / `lloyd_cluster` (called at ns/lib/graph.py:232 and
// ns/model/agg_interp.py:475) are in-place Gauss-Seidel pull sweeps in row order with strict `<`:
//     for i in rows: for jj in row i: d = w[jj] + x[col[jj]]; if d < x_i: x_i = d; z_i = z[col[jj]]
// repeated until a sweep changes no distance.  Distances at the fixed point do not depend on the
// sweep order; the labels of tied nodes do.  The kernels below emulate sweep s exactly:
//
//   X_s(i) = min( X_{s-1}(i), min_{j<i} fl(w_ij + X_s(j)), min_{j>=i} fl(w_ij + X_{s-1}(j)) )
//
// is a fixed-point problem on the DAG of "lower" edges (j<i); it has a unique solution, reached by
// in-place parallel relaxations (monotone, from above) repeated until a pass changes nothing.  The
// label of a node whose distance dropped in sweep s is the label, at the time of the sequential row
// scan, of the FIRST CSR-order neighbour whose candidate equals the new distance: Z_{s-1}(j*) when
// j* >= i or j* did not change in sweep s, else Z_s(j*) — a chain through strictly smaller indices
// that is resolved by pointer jumping.  Nothing here depends on thread scheduling, so results are
// identical run to run and identical to the sequential loops.
#include <stdlib.h>
#include <time.h>
#include "common.cuh"

namespace mlamg {

constexpr int AGG_LANES = 8;
constexpr int AGG_THREADS = 256;

template <typename T>
__device__ __forceinline__ T group_min8(T v) {
#pragma unroll
    for (int o = AGG_LANES / 2; o > 0; o >>= 1) {
        const T u = __shfl_xor_sync(0xffffffffu, v, o, AGG_LANES);
        v = u < v ? u : v;
    }
    return v;
}

__device__ __forceinline__ int group_min8i(int v) {
#pragma unroll
    for (int o = AGG_LANES / 2; o > 0; o >>= 1) {
        const int u = __shfl_xor_sync(0xffffffffu, v, o, AGG_LANES);
        v = u < v ? u : v;
    }
    return v;
}

// One relaxation pass.  ORDERED: neighbours j < i read the in-progress array xcur, j >= i read xprev
// (sequential-sweep emulation).  !ORDERED: every neighbour reads xcur (order-free fixed point).
template <typename T, bool ORDERED>
__global__ void __launch_bounds__(AGG_THREADS)
bf_relax_kernel(int n, const int *__restrict__ rowptr, const int *__restrict__ col, const T *__restrict__ w,
                const T *__restrict__ xprev, T *xcur, int *__restrict__ changed) {
    const long long gt = (long long)blockIdx.x * AGG_THREADS + threadIdx.x;
    const long long i = gt / AGG_LANES;
    const int lane = threadIdx.x & (AGG_LANES - 1);
    T own = Limits<T>::max();
    T m = Limits<T>::max();
    if (i < n) {
        own = __ldcg(&xcur[i]);
        m = own;
        for (int jj = rowptr[i] + lane; jj < rowptr[i + 1]; jj += AGG_LANES) {
            const int j = col[jj];
            const T xj = (!ORDERED || j < i) ? __ldcg(&xcur[j]) : xprev[j];
            const T d = w[jj] + xj;
            if (d < m) m = d;
        }
    }
    m = group_min8(m);
    if (i < n && lane == 0 && m < own) {
        xcur[i] = m;
        *changed = 1;
    }
}

// Frontier form of the same pass (pattern-symmetric graphs): a row is evaluated only when one of its inputs changed since
// its last evaluation.  Three byte-flag arrays: `fin` rows to evaluate in this pass (compacted into a row list and cleared
// by frontier_compact_kernel), `fout` rows for the next pass of this sweep, `fnext` rows for the first pass of the NEXT
// sweep.  A row j that lowers its value marks its readers (= its own neighbours, the pattern being
// symmetric): ORDERED — readers i > j read xcur[j] in this very sweep -> fout; readers i < j read xprev[j], i.e. see the
// change one sequential sweep later -> fnext.  !ORDERED — every reader reads xcur -> fout.  The relaxation is monotone
// and every reader of a changed value is re-evaluated afterwards, so the fixed point (hence every distance, label and
// sweep count) is the one of the full passes; only the work differs: after the first passes of a sweep a few percent
// of the rows are active.
// fin -> list of flagged rows (order arbitrary), fin cleared; *count = rows listed.  Each thread scans 16 flags with one
// 128-bit load (the flag arrays are 256-byte aligned and padded): a pass over the byte flags costs n bytes of traffic
// and n/16 threads, where a grid of 8 lanes per row would cost n*8 threads of launch work per pass even when a handful
// of rows are active.
__global__ void __launch_bounds__(AGG_THREADS)
frontier_compact_kernel(int n, unsigned char *__restrict__ fin, int *__restrict__ list, int *__restrict__ count) {
    const long long t = (long long)blockIdx.x * AGG_THREADS + threadIdx.x;
    const long long i0 = t * 16;
    uint4 f = make_uint4(0u, 0u, 0u, 0u);
    if (i0 < n) f = *reinterpret_cast<const uint4 *>(fin + i0);         // bytes past n inside the padded array are zero
    const unsigned wv[4] = {f.x, f.y, f.z, f.w};
    int mine = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) mine += __popc(wv[k] & 0x01010101u);     // flags are 0 / 1
    // warp-level exclusive scan of the per-thread counts, one global atomic per warp
    const int lane = threadIdx.x & 31;
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) return;
    int base = 0;
    if (lane == 0) base = atomicAdd(count, total);
    base = __shfl_sync(0xffffffffu, base, 0) + incl - mine;
    if (mine) {
#pragma unroll
        for (int k = 0; k < 4; k++)
#pragma unroll
            for (int bb = 0; bb < 4; bb++)
                if ((wv[k] >> (8 * bb)) & 1u) list[base++] = (int)(i0 + 4 * k + bb);
        *reinterpret_cast<uint4 *>(fin + i0) = make_uint4(0u, 0u, 0u, 0u);
    }
}

template <typename T, bool ORDERED>
__global__ void __launch_bounds__(AGG_THREADS)
bf_relax_frontier_kernel(const int *__restrict__ list, const int *__restrict__ count, const int *__restrict__ rowptr,
                         const int *__restrict__ col, const T *__restrict__ w, const T *__restrict__ xprev, T *xcur,
                         unsigned char *fout, unsigned char *fnext, int *__restrict__ changed) {
    const int lane = threadIdx.x & (AGG_LANES - 1);
    const int nact = *count;
    const long long ngroups = (long long)gridDim.x * (AGG_THREADS / AGG_LANES);
    // uniform trip count per warp: every lane takes part in the group shuffles
    const long long g0 = ((long long)blockIdx.x * AGG_THREADS + threadIdx.x) / AGG_LANES;
    const long long w0 = g0 - ((threadIdx.x & 31) / AGG_LANES);          // first group of this warp
    for (long long base = w0; base < nact; base += ngroups) {
        const long long g = base + ((threadIdx.x & 31) / AGG_LANES);
        const bool act = g < nact;
        T own = Limits<T>::max();
        T m = Limits<T>::max();
        int start = 0, end = 0, i = 0;
        if (act) {
            i = list[g];
            start = rowptr[i];
            end = rowptr[i + 1];
            own = __ldcg(&xcur[i]);
            m = own;
            for (int jj = start + lane; jj < end; jj += AGG_LANES) {
                const int j = col[jj];
                const T xj = (!ORDERED || j < i) ? __ldcg(&xcur[j]) : xprev[j];
                const T d = w[jj] + xj;
                if (d < m) m = d;
            }
        }
        m = group_min8(m);
        if (act && m < own) {
            if (lane == 0) {
                xcur[i] = m;
                *changed = 1;
            }
            for (int jj = start + lane; jj < end; jj += AGG_LANES) {
                const int j = col[jj];
                if (j == i) continue;
                if (!ORDERED || j > i) fout[j] = 1;
                else fnext[j] = 1;
            }
        }
    }
}

// *asym = 1 unless every stored entry (i, j) has a stored partner (j, i)
__global__ void __launch_bounds__(AGG_THREADS)
pattern_symmetric_kernel(int n, const int *__restrict__ rowptr, const int *__restrict__ col, int *__restrict__ asym) {
    const long long gt = (long long)blockIdx.x * AGG_THREADS + threadIdx.x;
    const long long i = gt / AGG_LANES;
    const int lane = threadIdx.x & (AGG_LANES - 1);
    if (i >= n) return;
    for (int jj = rowptr[i] + lane; jj < rowptr[i + 1]; jj += AGG_LANES) {
        const int j = col[jj];
        if (j == i) continue;
        bool found = false;
        if (j >= 0 && j < n)
            for (int kk = rowptr[j]; kk < rowptr[j + 1]; kk++)
                if (col[kk] == (int)i) { found = true; break; }
        if (!found) { *asym = 1; return; }
    }
}

// rows next to a seed are the only ones the first pass can change
__global__ void __launch_bounds__(AGG_THREADS)
seed_mark_kernel(int k, const int *__restrict__ seeds, const int *__restrict__ rowptr, const int *__restrict__ col,
                 unsigned char *__restrict__ flags) {
    const long long gt = (long long)blockIdx.x * AGG_THREADS + threadIdx.x;
    const long long sidx = gt / AGG_LANES;
    const int lane = threadIdx.x & (AGG_LANES - 1);
    if (sidx >= k) return;
    const int node = seeds[sidx];
    for (int jj = rowptr[node] + lane; jj < rowptr[node + 1]; jj += AGG_LANES) flags[col[jj]] = 1;
}

// Labels of sweep s (see file header).  link[i] = j* when Z_s(i) must be taken from Z_s(j*) which is
// itself produced in this sweep, -1 when zcur[i] is final.
// stage 1 (one thread per row): rows whose distance did not drop in this sweep keep their label; the others are listed
template <typename T>
__global__ void __launch_bounds__(AGG_THREADS)
bf_label_scan_kernel(int n, const T *__restrict__ xprev, const T *__restrict__ xcur, const int *__restrict__ zprev,
                     int *__restrict__ zcur, int *__restrict__ link, int *__restrict__ list, int *__restrict__ count) {
    const long long i = (long long)blockIdx.x * AGG_THREADS + threadIdx.x;
    const bool updated = i < n && xcur[i] < xprev[i];
    if (i < n && !updated) {
        zcur[i] = zprev[i];
        link[i] = -1;
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, updated);
    if (ballot == 0u) return;
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0) base = atomicAdd(count, __popc(ballot));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (updated) list[base + __popc(ballot & ((1u << lane) - 1u))] = (int)i;
}

// stage 2 (8 lanes per listed row): the first CSR-order neighbour attaining the new distance decides the label
template <typename T>
__global__ void __launch_bounds__(AGG_THREADS)
bf_label_kernel(const int *__restrict__ list, const int *__restrict__ count, const int *__restrict__ rowptr,
                const int *__restrict__ col, const T *__restrict__ w, const T *__restrict__ xprev,
                const T *__restrict__ xcur, const int *__restrict__ zprev, int *__restrict__ zcur, int *__restrict__ link,
                int *__restrict__ pending, int *__restrict__ error) {
    const int lane = threadIdx.x & (AGG_LANES - 1);
    const int nact = *count;
    const long long ngroups = (long long)gridDim.x * (AGG_THREADS / AGG_LANES);
    const long long g0 = ((long long)blockIdx.x * AGG_THREADS + threadIdx.x) / AGG_LANES;
    const long long w0 = g0 - ((threadIdx.x & 31) / AGG_LANES);
    for (long long base = w0; base < nact; base += ngroups) {
        const long long g = base + ((threadIdx.x & 31) / AGG_LANES);
        const bool act = g < nact;
        int first = 0x7fffffff;
        int i = 0;
        if (act) {
            i = list[g];
            const T xi = xcur[i];
            for (int jj = rowptr[i] + lane; jj < rowptr[i + 1]; jj += AGG_LANES) {
                const int j = col[jj];
                const T xj = (j < i) ? xcur[j] : xprev[j];
                const T d = w[jj] + xj;
                if (d == xi) { first = jj; break; }
            }
        }
        first = group_min8i(first);
        if (act && lane == 0) {
            if (first == 0x7fffffff) {
                *error = 1;
                zcur[i] = zprev[i];
                link[i] = -1;
            } else {
                const int j = col[first];
                if (j < i && xcur[j] < xprev[j]) {
                    link[i] = j;
                    *pending = 1;
                } else {
                    zcur[i] = zprev[j];
                    link[i] = -1;
                }
            }
        }
    }
}

// pointer jumping, Jacobi style: reads link_in (state before this launch), writes link_out
__global__ void __launch_bounds__(AGG_THREADS)
bf_chain_kernel(int n, const int *__restrict__ link_in, int *__restrict__ link_out, int *zcur,
                int *__restrict__ pending) {
    const long long i = (long long)blockIdx.x * AGG_THREADS + threadIdx.x;
    if (i >= n) return;
    const int l = link_in[i];
    if (l < 0) { link_out[i] = -1; return; }
    const int ll = link_in[l];
    if (ll < 0) {
        zcur[i] = __ldcg(&zcur[l]);   // final since a previous launch
        link_out[i] = -1;
    } else {
        link_out[i] = ll;
        *pending = 1;
    }
}

template <typename T>
__global__ void __launch_bounds__(AGG_THREADS) fill_kernel(int n, T v, T *__restrict__ a) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        a[i] = v;
}

// seeds: dist 0; label = seed node id (LABEL_IS_INDEX = false) or seed index (true)
template <typename T, bool LABEL_IS_INDEX>
__global__ void __launch_bounds__(AGG_THREADS) seed_init_kernel(int k, const int *__restrict__ seeds,
                                                                T *__restrict__ x, int *__restrict__ z) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= k) return;
    const int node = seeds[s];
    x[node] = (T)0;
    z[node] = LABEL_IS_INDEX ? (int)s : node;
}

static unsigned ew_blocks(int n) {
    unsigned b = cdiv(n > 0 ? n : 1, AGG_THREADS);
    return b > 148u * 16u ? 148u * 16u : b;
}

// counters of the last aggregation calls (debug / profiling aid, read by mlamg_agg_stats)
static long long g_stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // ordered passes, free passes, sweeps, chain passes, active rows (ordered), active rows (free), lloyd iterations, unused

static inline long long now_ns() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (long long)ts.tv_sec * 1000000000LL + ts.tv_nsec;
}

// grid of the frontier relaxation (grid-stride over the active-row list whose length lives on the device)
static unsigned frontier_grid(int n) {
    const unsigned need = cdiv((long long)n * AGG_LANES, AGG_THREADS);
    return need < 148u * 32u ? need : 148u * 32u;
}

// Pinned host mirror of a few device flags.  The buffers are allocated once per (thread, device) and reused by every
// aggregation call: cudaMallocHost / cudaFreeHost per call cost tens to hundreds of milliseconds on some hosts and made
// the Lloyd time erratic (0.17 - 0.58 s for identical device work).
struct FlagBuffers {
    int device = -1;
    int *dev = nullptr;
    int *host = nullptr;
};
static thread_local FlagBuffers g_flagbuf;

struct Flags {
    int *dev = nullptr;
    int *host = nullptr;
    int init() {
        int d = 0;
        MLAMG_CUDA(cudaGetDevice(&d));
        if (g_flagbuf.device != d) {          // first use on this device (buffers of another device stay allocated)
            g_flagbuf.dev = nullptr;
            g_flagbuf.host = nullptr;
            MLAMG_CUDA(cudaMalloc(&g_flagbuf.dev, 4 * sizeof(int)));
            MLAMG_CUDA(cudaMallocHost(&g_flagbuf.host, 4 * sizeof(int)));
            g_flagbuf.device = d;
        }
        dev = g_flagbuf.dev;
        host = g_flagbuf.host;
        return MLAMG_OK;
    }
    int clear(cudaStream_t s) { MLAMG_CUDA(cudaMemsetAsync(dev, 0, 4 * sizeof(int), s)); return MLAMG_OK; }
    int fetch(cudaStream_t s);
};

int Flags::fetch(cudaStream_t s) {
    const long long t0 = now_ns();
    MLAMG_CUDA(cudaMemcpyAsync(host, dev, 4 * sizeof(int), cudaMemcpyDeviceToHost, s));
    MLAMG_CUDA(cudaStreamSynchronize(s));
    g_stats[7] += now_ns() - t0;          // host time spent waiting for the device in the flag round trips
    return MLAMG_OK;
}


// Work arrays of one Bellman-Ford run
template <typename T>
struct BfWork {
    T *xa, *xb;
    int *za, *zb, *la, *lb;
    unsigned char *f0 = nullptr, *f1 = nullptr, *f2 = nullptr;   // frontier flags (all zero between runs); null = full passes
    int *list = nullptr;                                          // frontier: active rows of the current pass
};

// Emulates the sequential sweeps to their fixed point.  On entry x/z hold the initial state; on exit
// they hold the final distances / labels.  Returns the number of sequential sweeps (incl. the final
// no-change sweep) in *sweeps.
// `first_flags`: frontier runs only — the rows to evaluate in the first pass of the first sweep (consumed, left zero).
template <typename T>
static int bf_ordered_fixed_point(int n, const int *rowptr, const int *col, const T *w, T *x, int *z,
                                  BfWork<T> wk, Flags &fl, int *sweeps, cudaStream_t s) {
    const unsigned gb = cdiv((long long)n * AGG_LANES, AGG_THREADS);
    const unsigned eb = cdiv(n, AGG_THREADS);
    T *xprev = wk.xa, *xcur = wk.xb;
    int *zprev = wk.za, *zcur = wk.zb;
    const bool frontier = wk.f0 != nullptr;
    // frontier: fnext carries the initial rows (set by the caller in wk.f2); fin / fout are zero
    unsigned char *fin = wk.f0, *fout = wk.f1, *fnext = wk.f2;
    MLAMG_CUDA(cudaMemcpyAsync(xprev, x, (size_t)n * sizeof(T), cudaMemcpyDeviceToDevice, s));
    MLAMG_CUDA(cudaMemcpyAsync(zprev, z, (size_t)n * sizeof(int), cudaMemcpyDeviceToDevice, s));
    int nsweeps = 0;
    for (;;) {
        nsweeps++;
        g_stats[2]++;
        MLAMG_CUDA(cudaMemcpyAsync(xcur, xprev, (size_t)n * sizeof(T), cudaMemcpyDeviceToDevice, s));
        if (frontier) {      // rows whose higher neighbours changed during the previous sweep; fin and fout are all zero here
            unsigned char *t = fin; fin = fnext; fnext = t;
        }
        bool any = false;
        for (;;) {   // fixed point of sweep `nsweeps`
            MLAMG_TRY(fl.clear(s));
            if (frontier) {
                frontier_compact_kernel<<<cdiv(cdiv(n, 16), AGG_THREADS), AGG_THREADS, 0, s>>>(n, fin, wk.list, fl.dev + 1);
                MLAMG_LAUNCHED();
                bf_relax_frontier_kernel<T, true><<<frontier_grid(n), AGG_THREADS, 0, s>>>(wk.list, fl.dev + 1, rowptr, col, w, xprev,
                                                                                          xcur, fout, fnext, fl.dev);
            } else
                bf_relax_kernel<T, true><<<gb, AGG_THREADS, 0, s>>>(n, rowptr, col, w, xprev, xcur, fl.dev);
            MLAMG_LAUNCHED();
            MLAMG_TRY(fl.fetch(s));
            g_stats[0]++;
            g_stats[4] += frontier ? fl.host[1] : n;
            if (frontier) { unsigned char *t = fin; fin = fout; fout = t; }
            if (!fl.host[0]) break;
            any = true;
        }
        if (!any) break;   // this sweep changed nothing: the sequential loop stops here
        MLAMG_TRY(fl.clear(s));
        bf_label_scan_kernel<T><<<eb, AGG_THREADS, 0, s>>>(n, xprev, xcur, zprev, zcur, wk.la, wk.list, fl.dev + 3);
        MLAMG_LAUNCHED();
        bf_label_kernel<T><<<frontier_grid(n), AGG_THREADS, 0, s>>>(wk.list, fl.dev + 3, rowptr, col, w, xprev, xcur, zprev, zcur,
                                                                    wk.la, fl.dev + 1, fl.dev + 2);
        MLAMG_LAUNCHED();
        MLAMG_TRY(fl.fetch(s));
        if (fl.host[2]) return set_error(MLAMG_EINVAL, "bellman_ford: inconsistent relaxation (NaN or negative weight?)");
        int *lin = wk.la, *lout = wk.lb;
        while (fl.host[1]) {
            MLAMG_TRY(fl.clear(s));
            bf_chain_kernel<<<eb, AGG_THREADS, 0, s>>>(n, lin, lout, zcur, fl.dev + 1);
            MLAMG_LAUNCHED();
            MLAMG_TRY(fl.fetch(s));
            g_stats[3]++;
            int *t = lin; lin = lout; lout = t;
        }
        { T *t = xprev; xprev = xcur; xcur = t; }
        { int *t = zprev; zprev = zcur; zcur = t; }
    }
    MLAMG_CUDA(cudaMemcpyAsync(x, xprev, (size_t)n * sizeof(T), cudaMemcpyDeviceToDevice, s));
    MLAMG_CUDA(cudaMemcpyAsync(z, zprev, (size_t)n * sizeof(int), cudaMemcpyDeviceToDevice, s));
    if (sweeps) *sweeps = nsweeps;
    return MLAMG_OK;
}

// order-free fixed point (distances only), in place on x.  Frontier runs: wk.f0 holds the rows of the first pass.
template <typename T>
static int bf_free_fixed_point(int n, const int *rowptr, const int *col, const T *w, T *x, BfWork<T> wk, Flags &fl,
                               cudaStream_t s) {
    const unsigned gb = cdiv((long long)n * AGG_LANES, AGG_THREADS);
    unsigned char *fin = wk.f0, *fout = wk.f1;
    for (;;) {
        MLAMG_TRY(fl.clear(s));
        if (fin) {
            frontier_compact_kernel<<<cdiv(cdiv(n, 16), AGG_THREADS), AGG_THREADS, 0, s>>>(n, fin, wk.list, fl.dev + 1);
            MLAMG_LAUNCHED();
            bf_relax_frontier_kernel<T, false><<<frontier_grid(n), AGG_THREADS, 0, s>>>(wk.list, fl.dev + 1, rowptr, col, w, x, x, fout,
                                                                                       fout, fl.dev);
        } else
            bf_relax_kernel<T, false><<<gb, AGG_THREADS, 0, s>>>(n, rowptr, col, w, x, x, fl.dev);
        MLAMG_LAUNCHED();
        MLAMG_TRY(fl.fetch(s));
        g_stats[1]++;
        g_stats[5] += fin ? fl.host[1] : n;
        if (fin) { unsigned char *t = fin; fin = fout; fout = t; }
        if (!fl.host[0]) break;
    }
    // the last pass changed nothing, so it marked nothing: fin / fout are zero again
    return MLAMG_OK;
}

template <typename T>
static int alloc_work(int n, cudaStream_t s, Scratch &buf, BfWork<T> *wk) {
    MLAMG_SCRATCH_OK(buf);
    unsigned char *p = buf.as<unsigned char>();
    const size_t nx = ((size_t)n * sizeof(T) + 255) & ~(size_t)255, ni = ((size_t)n * sizeof(int) + 255) & ~(size_t)255;
    const size_t nf = ((size_t)n + 255) & ~(size_t)255;
    wk->xa = (T *)p; p += nx;
    wk->xb = (T *)p; p += nx;
    wk->za = (int *)p; p += ni;
    wk->zb = (int *)p; p += ni;
    wk->la = (int *)p; p += ni;
    wk->lb = (int *)p; p += ni;
    wk->f0 = p; p += nf;
    wk->f1 = p; p += nf;
    wk->f2 = p; p += nf;
    wk->list = (int *)p;
    MLAMG_CUDA(cudaMemsetAsync(wk->f0, 0, 3 * nf, s));
    return MLAMG_OK;
}
static size_t work_bytes(int n, size_t tsize) {
    const size_t nx = ((size_t)n * tsize + 255) & ~(size_t)255, ni = ((size_t)n * sizeof(int) + 255) & ~(size_t)255;
    const size_t nf = ((size_t)n + 255) & ~(size_t)255;
    return 2 * nx + 5 * ni + 3 * nf + 256;
}

// frontier passes need a symmetric pattern (a row's neighbours are its readers); MLAMG_AGG_FRONTIER=0 forces full passes
static int frontier_usable(int n, const int *rowptr, const int *col, Flags &fl, bool *ok, cudaStream_t s) {
    const char *env = getenv("MLAMG_AGG_FRONTIER");
    if (env && env[0] == '0') { *ok = false; return MLAMG_OK; }
    MLAMG_TRY(fl.clear(s));
    pattern_symmetric_kernel<<<cdiv((long long)n * AGG_LANES, AGG_THREADS), AGG_THREADS, 0, s>>>(n, rowptr, col, fl.dev + 3);
    MLAMG_LAUNCHED();
    MLAMG_TRY(fl.fetch(s));
    *ok = fl.host[3] == 0;
    return MLAMG_OK;
}

template <typename T>
static int bellman_ford_t(int n, const int *rowptr, const int *col, const T *w, int nseeds, const int *seeds, T *dist,
                          int *nearest, int *sweeps_host, cudaStream_t s) {
    if (n < 0 || nseeds < 0) return set_error(MLAMG_EINVAL, "bellman_ford: bad n/nseeds");
    if (n == 0) { if (sweeps_host) *sweeps_host = 0; return MLAMG_OK; }
    Flags fl;
    MLAMG_TRY(fl.init());
    Scratch buf(work_bytes(n, sizeof(T)), s);
    BfWork<T> wk;
    MLAMG_TRY(alloc_work<T>(n, s, buf, &wk));
    fill_kernel<T><<<ew_blocks(n), AGG_THREADS, 0, s>>>(n, Limits<T>::max(), dist);
    MLAMG_LAUNCHED();
    fill_kernel<int><<<ew_blocks(n), AGG_THREADS, 0, s>>>(n, -1, nearest);
    MLAMG_LAUNCHED();
    if (nseeds > 0) {
        seed_init_kernel<T, false><<<cdiv(nseeds, AGG_THREADS), AGG_THREADS, 0, s>>>(nseeds, seeds, dist, nearest);
        MLAMG_LAUNCHED();
    }
    bool frontier = false;
    MLAMG_TRY(frontier_usable(n, rowptr, col, fl, &frontier, s));
    if (!frontier) wk.f0 = wk.f1 = wk.f2 = nullptr;
    else if (nseeds > 0) {
        seed_mark_kernel<<<cdiv((long long)nseeds * AGG_LANES, AGG_THREADS), AGG_THREADS, 0, s>>>(nseeds, seeds, rowptr, col, wk.f2);
        MLAMG_LAUNCHED();
    }
    int rc = bf_ordered_fixed_point<T>(n, rowptr, col, w, dist, nearest, wk, fl, sweeps_host, s);
    MLAMG_CUDA(cudaStreamSynchronize(s));
    return rc;
}

// ------------------------------------------------------------------ Lloyd
__global__ void __launch_bounds__(AGG_THREADS)
lloyd_boundary_kernel(int n, const int *__restrict__ rowptr, const int *__restrict__ col, const int *__restrict__ z,
                      int *__restrict__ is_boundary) {
    const long long gt = (long long)blockIdx.x * AGG_THREADS + threadIdx.x;
    const long long i = gt / AGG_LANES;
    const int lane = threadIdx.x & (AGG_LANES - 1);
    int b = 0;
    if (i < n) {
        const int zi = z[i];
        for (int jj = rowptr[i] + lane; jj < rowptr[i + 1]; jj += AGG_LANES)
            if (z[col[jj]] != zi) { b = 1; break; }
    }
#pragma unroll
    for (int o = AGG_LANES / 2; o > 0; o >>= 1) b |= __shfl_xor_sync(0xffffffffu, b, o, AGG_LANES);
    if (i < n && lane == 0) is_boundary[i] = b;
}

// flags (optional): every row is evaluated in the first inward pass
template <typename T>
__global__ void __launch_bounds__(AGG_THREADS) lloyd_inward_init_kernel(int n, const int *__restrict__ is_boundary,
                                                                        T *__restrict__ x, unsigned char *__restrict__ flags) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        x[i] = is_boundary[i] ? (T)0 : Limits<T>::max();
        if (flags) flags[i] = is_boundary[i] ? 0 : 1;      // a boundary node sits at distance 0: it can never improve
    }
}

// per-cluster maximum of the inward distance (order-preserving key), then the first index attaining it
template <typename T>
__global__ void __launch_bounds__(AGG_THREADS) lloyd_max_kernel(int n, const int *__restrict__ z, const T *__restrict__ x,
                                                                unsigned long long *__restrict__ best) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = z[i];
    if (c >= 0) atomicMax(&best[c], f2key(x[i]));
}
template <typename T>
__global__ void __launch_bounds__(AGG_THREADS) lloyd_argmax_kernel(int n, const int *__restrict__ z,
                                                                   const T *__restrict__ x,
                                                                   const unsigned long long *__restrict__ best,
                                                                   int *__restrict__ arg) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = z[i];
    if (c >= 0 && f2key(x[i]) == best[c]) atomicMin(&arg[c], (int)i);
}
// z[seed] moves to the first node of strictly larger inward distance (graph.h seed update loop)
template <typename T>
__global__ void __launch_bounds__(AGG_THREADS) lloyd_move_kernel(int k, const T *__restrict__ x,
                                                                 const int *__restrict__ arg, int *__restrict__ seeds,
                                                                 int *__restrict__ moved) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= k) return;
    const int old = seeds[c];
    const int a = arg[c];
    if (a != 0x7fffffff && x[old] < x[a]) {
        seeds[c] = a;
        *moved = 1;
    }
}

template <typename T>
static int lloyd_cluster_t(int n, const int *rowptr, const int *col, const T *w, int k, int *seeds, int maxiter, T *dist,
                           int *clusters, int *iters_host, cudaStream_t s) {
    if (n <= 0 || k <= 0) return set_error(MLAMG_EINVAL, "lloyd_cluster: need n > 0 and at least one seed");
    Flags fl;
    MLAMG_TRY(fl.init());
    Scratch buf(work_bytes(n, sizeof(T)), s), bnd((size_t)n * sizeof(int), s),
        best((size_t)k * sizeof(unsigned long long), s), arg((size_t)k * sizeof(int), s);
    BfWork<T> wk;
    MLAMG_TRY(alloc_work<T>(n, s, buf, &wk));
    MLAMG_SCRATCH_OK(bnd);
    MLAMG_SCRATCH_OK(best);
    MLAMG_SCRATCH_OK(arg);
    const unsigned gb = cdiv((long long)n * AGG_LANES, AGG_THREADS), eb = cdiv(n, AGG_THREADS), kb = cdiv(k, AGG_THREADS);
    bool frontier = false;
    MLAMG_TRY(frontier_usable(n, rowptr, col, fl, &frontier, s));
    if (!frontier) wk.f0 = wk.f1 = wk.f2 = nullptr;
    int it = 0;
    for (it = 0; it < maxiter;) {
        // reset + seeds
        fill_kernel<T><<<ew_blocks(n), AGG_THREADS, 0, s>>>(n, Limits<T>::max(), dist);
        MLAMG_LAUNCHED();
        fill_kernel<int><<<ew_blocks(n), AGG_THREADS, 0, s>>>(n, -1, clusters);
        MLAMG_LAUNCHED();
        seed_init_kernel<T, true><<<kb, AGG_THREADS, 0, s>>>(k, seeds, dist, clusters);
        MLAMG_LAUNCHED();
        if (frontier) {
            seed_mark_kernel<<<cdiv((long long)k * AGG_LANES, AGG_THREADS), AGG_THREADS, 0, s>>>(k, seeds, rowptr, col, wk.f2);
            MLAMG_LAUNCHED();
        }
        // outward propagation with sequential tie-breaking
        MLAMG_TRY(bf_ordered_fixed_point<T>(n, rowptr, col, w, dist, clusters, wk, fl, nullptr, s));
        // cluster boundaries -> distance 0, interior -> max
        lloyd_boundary_kernel<<<gb, AGG_THREADS, 0, s>>>(n, rowptr, col, clusters, bnd.as<int>());
        MLAMG_LAUNCHED();
        lloyd_inward_init_kernel<T><<<eb, AGG_THREADS, 0, s>>>(n, bnd.as<int>(), dist, wk.f0);
        MLAMG_LAUNCHED();
        // inward propagation: labels of interior nodes cannot change (all their neighbours carry the
        // same label), so only the order-independent distances are needed
        MLAMG_TRY(bf_free_fixed_point<T>(n, rowptr, col, w, dist, wk, fl, s));
        // seed update
        MLAMG_CUDA(cudaMemsetAsync(best.p, 0, (size_t)k * sizeof(unsigned long long), s));
        fill_kernel<int><<<ew_blocks(k), AGG_THREADS, 0, s>>>(k, 0x7fffffff, arg.as<int>());
        MLAMG_LAUNCHED();
        lloyd_max_kernel<T><<<eb, AGG_THREADS, 0, s>>>(n, clusters, dist, best.as<unsigned long long>());
        MLAMG_LAUNCHED();
        lloyd_argmax_kernel<T><<<eb, AGG_THREADS, 0, s>>>(n, clusters, dist, best.as<unsigned long long>(), arg.as<int>());
        MLAMG_LAUNCHED();
        MLAMG_TRY(fl.clear(s));
        lloyd_move_kernel<T><<<kb, AGG_THREADS, 0, s>>>(k, dist, arg.as<int>(), seeds, fl.dev);
        MLAMG_LAUNCHED();
        MLAMG_TRY(fl.fetch(s));
        it++;
        g_stats[6]++;
        if (!fl.host[0]) break;   // seeds unchanged
    }
    if (iters_host) *iters_host = it;
    MLAMG_CUDA(cudaStreamSynchronize(s));
    return MLAMG_OK;
}

// ------------------------------------------------------------------ reference-owned push Bellman-Ford
// ns/lib/graph.py:7-53.  Pass p over the row-major edge list, with U(i) = dist[i] at the time row i is
// visited and F(i) = dist[i] at the end of the pass:
//   U(i) = min( D_{p-1}(i), min_{i'<i, i'->i} fl(U(i') + w) )        (DAG fixed point)
//   F(i) = min( U(i),       min_{i'>i, i'->i} fl(U(i') + w) )
// the label follows the FIRST source (ascending i') that attains the minimum, with the label that
// source had when its row was visited.  Needs in-edges: tr_* is the transpose CSR (sources sorted).
__global__ void __launch_bounds__(AGG_THREADS)
mbf_u_relax_kernel(int n, const int *__restrict__ tr_rowptr, const int *__restrict__ tr_col,
                   const float *__restrict__ tr_w, float *u, int *__restrict__ changed) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float own = __ldcg(&u[i]);
    float m = own;
    for (int jj = tr_rowptr[i]; jj < tr_rowptr[i + 1]; jj++) {
        const int src = tr_col[jj];
        if (src >= i) break;   // sources are sorted ascending
        const float d = __ldcg(&u[src]) + tr_w[jj];
        if (d < m) m = d;
    }
    if (m < own) { u[i] = m; *changed = 1; }
}

__global__ void __launch_bounds__(AGG_THREADS)
mbf_label_u_kernel(int n, const int *__restrict__ tr_rowptr, const int *__restrict__ tr_col,
                   const float *__restrict__ tr_w, const float *__restrict__ dprev, const float *__restrict__ u,
                   const long long *__restrict__ lprev, long long *__restrict__ lu, int *__restrict__ link,
                   int *__restrict__ pending) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float ui = u[i];
    if (!(ui < dprev[i])) { lu[i] = lprev[i]; link[i] = -1; return; }
    for (int jj = tr_rowptr[i]; jj < tr_rowptr[i + 1]; jj++) {
        const int src = tr_col[jj];
        if (src >= i) break;
        if (u[src] + tr_w[jj] == ui) {
            if (u[src] < dprev[src]) { link[i] = src; *pending = 1; }
            else { lu[i] = lprev[src]; link[i] = -1; }
            return;
        }
    }
    lu[i] = lprev[i];
    link[i] = -1;
}

__global__ void __launch_bounds__(AGG_THREADS)
mbf_chain_kernel(int n, const int *__restrict__ link_in, int *__restrict__ link_out, long long *lu,
                 int *__restrict__ pending) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int l = link_in[i];
    if (l < 0) { link_out[i] = -1; return; }
    const int ll = link_in[l];
    if (ll < 0) { lu[i] = __ldcg(&lu[l]); link_out[i] = -1; }
    else { link_out[i] = ll; *pending = 1; }
}

__global__ void __launch_bounds__(AGG_THREADS)
mbf_final_kernel(int n, const int *__restrict__ tr_rowptr, const int *__restrict__ tr_col,
                 const float *__restrict__ tr_w, const float *__restrict__ dprev, const float *__restrict__ u,
                 const long long *__restrict__ lu, float *__restrict__ dnew, long long *__restrict__ lnew,
                 int *__restrict__ changed) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float m = u[i];
    int arg = -1;
    for (int jj = tr_rowptr[i]; jj < tr_rowptr[i + 1]; jj++) {
        const int src = tr_col[jj];
        if (src <= i) continue;
        const float d = u[src] + tr_w[jj];
        if (d < m) { m = d; arg = src; }   // strict: first (ascending) source attaining the minimum
    }
    dnew[i] = m;
    lnew[i] = arg >= 0 ? lu[arg] : lu[i];
    if (m < dprev[i]) *changed = 1;
}

__global__ void __launch_bounds__(AGG_THREADS) mbf_init_kernel(int n, float *__restrict__ d, long long *__restrict__ l) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { d[i] = __int_as_float(0x7f800000); l[i] = 0; }
}
__global__ void __launch_bounds__(AGG_THREADS) mbf_seed_kernel(int k, const int *__restrict__ centers,
                                                               float *__restrict__ d, long long *__restrict__ l) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < k) { d[centers[c]] = 0.f; l[centers[c]] = centers[c]; }
}

}  // namespace mlamg

using namespace mlamg;

extern "C" {

/* profiling aid: counters accumulated by the aggregation calls of this process since the last reset
 * [0] ordered relaxation passes [1] order-free passes [2] sequential sweeps emulated [3] pointer-jumping passes
 * [4] rows evaluated by ordered passes [5] rows evaluated by order-free passes [6] Lloyd iterations [7] unused */
int mlamg_agg_stats(long long *out8, int reset) {
    for (int i = 0; i < 8; i++) {
        if (out8) out8[i] = g_stats[i];
        if (reset) g_stats[i] = 0;
    }
    return MLAMG_OK;
}

int mlamg_bellman_ford(int dtype, int n, const int *rowptr, const int *col, const void *w, int nseeds,
                       const int *seeds, void *dist, int *nearest, int *sweeps_host, mlamg_stream_t stream) {
    MLAMG_DISPATCH(dtype, return bellman_ford_t<T>(n, rowptr, col, (const T *)w, nseeds, seeds, (T *)dist, nearest,
                                                   sweeps_host, as_stream(stream)));
    return MLAMG_OK;
}

int mlamg_lloyd_cluster(int dtype, int n, const int *rowptr, const int *col, const void *w, int k, int *seeds,
                        int maxiter, void *dist, int *clusters, int *iters_host, mlamg_stream_t stream) {
    MLAMG_DISPATCH(dtype, return lloyd_cluster_t<T>(n, rowptr, col, (const T *)w, k, seeds, maxiter, (T *)dist, clusters,
                                                    iters_host, as_stream(stream)));
    return MLAMG_OK;
}

int mlamg_modified_bellman_ford(int n, const int *rowptr, const int *col, const float *w, int ncenters,
                                const int *centers, float *dist, long long *nearest, int *passes_host,
                                mlamg_stream_t stream) {
    cudaStream_t s = as_stream(stream);
    if (n < 0 || ncenters < 0) return set_error(MLAMG_EINVAL, "modified_bellman_ford: bad n/ncenters");
    if (n == 0) { if (passes_host) *passes_host = 1; return MLAMG_OK; }
    int nnz = 0;
    MLAMG_CUDA(cudaMemcpyAsync(&nnz, rowptr + n, sizeof(int), cudaMemcpyDeviceToHost, s));
    MLAMG_CUDA(cudaStreamSynchronize(s));
    Flags fl;
    MLAMG_TRY(fl.init());
    const size_t nn = (size_t)n;
    Scratch trp((nn + 1) * sizeof(int), s), trc((size_t)(nnz > 0 ? nnz : 1) * sizeof(int), s),
        trw((size_t)(nnz > 0 ? nnz : 1) * sizeof(float), s), ub(nn * sizeof(float), s), d2(nn * sizeof(float), s),
        lub(nn * sizeof(long long), s), l2(nn * sizeof(long long), s), la(nn * sizeof(int), s), lb(nn * sizeof(int), s);
    MLAMG_SCRATCH_OK(trp); MLAMG_SCRATCH_OK(trc); MLAMG_SCRATCH_OK(trw); MLAMG_SCRATCH_OK(ub); MLAMG_SCRATCH_OK(d2);
    MLAMG_SCRATCH_OK(lub); MLAMG_SCRATCH_OK(l2); MLAMG_SCRATCH_OK(la); MLAMG_SCRATCH_OK(lb);
    MLAMG_TRY(mlamg_csr_transpose(MLAMG_F32, n, n, nnz, rowptr, col, w, trp.as<int>(), trc.as<int>(), trw.as<float>(),
                                  stream));
    const unsigned eb = cdiv(n, AGG_THREADS);
    mbf_init_kernel<<<eb, AGG_THREADS, 0, s>>>(n, dist, nearest);
    MLAMG_LAUNCHED();
    if (ncenters > 0) {
        mbf_seed_kernel<<<cdiv(ncenters, AGG_THREADS), AGG_THREADS, 0, s>>>(ncenters, centers, dist, nearest);
        MLAMG_LAUNCHED();
    }
    float *dprev = dist, *dnew = d2.as<float>();
    long long *lprev = nearest, *lnew = l2.as<long long>();
    float *u = ub.as<float>();
    int passes = 0;
    for (;;) {
        passes++;
        MLAMG_CUDA(cudaMemcpyAsync(u, dprev, nn * sizeof(float), cudaMemcpyDeviceToDevice, s));
        for (;;) {
            MLAMG_TRY(fl.clear(s));
            mbf_u_relax_kernel<<<eb, AGG_THREADS, 0, s>>>(n, trp.as<int>(), trc.as<int>(), trw.as<float>(), u, fl.dev);
            MLAMG_LAUNCHED();
            MLAMG_TRY(fl.fetch(s));
            if (!fl.host[0]) break;
        }
        MLAMG_TRY(fl.clear(s));
        mbf_label_u_kernel<<<eb, AGG_THREADS, 0, s>>>(n, trp.as<int>(), trc.as<int>(), trw.as<float>(), dprev, u, lprev,
                                                      lub.as<long long>(), la.as<int>(), fl.dev + 1);
        MLAMG_LAUNCHED();
        MLAMG_TRY(fl.fetch(s));
        int *lin = la.as<int>(), *lout = lb.as<int>();
        while (fl.host[1]) {
            MLAMG_TRY(fl.clear(s));
            mbf_chain_kernel<<<eb, AGG_THREADS, 0, s>>>(n, lin, lout, lub.as<long long>(), fl.dev + 1);
            MLAMG_LAUNCHED();
            MLAMG_TRY(fl.fetch(s));
            int *t = lin; lin = lout; lout = t;
        }
        MLAMG_TRY(fl.clear(s));
        mbf_final_kernel<<<eb, AGG_THREADS, 0, s>>>(n, trp.as<int>(), trc.as<int>(), trw.as<float>(), dprev, u,
                                                    lub.as<long long>(), dnew, lnew, fl.dev);
        MLAMG_LAUNCHED();
        MLAMG_TRY(fl.fetch(s));
        { float *t = dprev; dprev = dnew; dnew = t; }
        { long long *t = lprev; lprev = lnew; lnew = t; }
        if (!fl.host[0]) break;
    }
    if (dprev != dist) {
        MLAMG_CUDA(cudaMemcpyAsync(dist, dprev, nn * sizeof(float), cudaMemcpyDeviceToDevice, s));
        MLAMG_CUDA(cudaMemcpyAsync(nearest, lprev, nn * sizeof(long long), cudaMemcpyDeviceToDevice, s));
    }
    MLAMG_CUDA(cudaStreamSynchronize(s));
    if (passes_host) *passes_host = passes;
    return MLAMG_OK;
}

}  // extern "C"
