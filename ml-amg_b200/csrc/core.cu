// Error plumbing, prefix scan, deterministic reductions and BLAS-1 helpers.
#include <stdarg.h>
#include <string.h>
#include <atomic>
#include "common.cuh"

namespace mlamg {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

int set_error(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int set_cuda_error(cudaError_t e, const char *file, int line) {
    const char *base = strrchr(file, '/');
    snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) at %s:%d", (int)e, cudaGetErrorString(e),
             base ? base + 1 : file, line);
    return MLAMG_ECUDA;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

void keep_pool_memory() {
    static std::atomic<unsigned long long> done{0};      // bit per device ordinal (< 64)
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return;
    const unsigned long long bit = 1ull << dev;
    if (done.load(std::memory_order_relaxed) & bit) return;
    cudaMemPool_t pool = nullptr;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess && pool) {
        unsigned long long threshold = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
    }
    done.fetch_or(bit, std::memory_order_relaxed);
}

// ------------------------------------------------------------------ exclusive scan (int32)
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_sums(const int *__restrict__ in, int n,
                                                               int *__restrict__ tile_sums) {
    __shared__ int ws[SCAN_THREADS / 32];
    const long long base = (long long)blockIdx.x * SCAN_TILE;
    int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        long long i = base + (long long)k * SCAN_THREADS + threadIdx.x;
        if (i < n) s += in[i];
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < SCAN_THREADS / 32; w++) t += ws[w];
        tile_sums[blockIdx.x] = t;
    }
}

// single block: exclusive scan of tile_sums[0..nt) in place
__global__ void __launch_bounds__(1024) scan_tile_offsets(int *tile_sums, int nt) {
    __shared__ int ws[32];
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < nt; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = (i < nt) ? tile_sums[i] : 0;
        // inclusive warp scan
        int x = v;
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) ws[wid] = x;
        __syncthreads();
        if (wid == 0) {
            int w = ws[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += y;
            }
            ws[lane] = w;  // inclusive scan of warp totals
        }
        __syncthreads();
        const int warp_off = (wid == 0) ? 0 : ws[wid - 1];
        const int carry = carry_s;
        if (i < nt) tile_sums[i] = carry + warp_off + x - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + warp_off + x;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_tiles(const int *in, int *out, int n,
                                                           const int *__restrict__ tile_offsets) {
    __shared__ int ws[SCAN_THREADS / 32];
    const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        long long i = base + k;
        v[k] = (i < n) ? in[i] : 0;
        s += v[k];
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int x = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) ws[wid] = x;
    __syncthreads();  // also orders every read of `in` before the writes of `out` (in may alias out)
    int warp_off = 0;
    for (int w = 0; w < wid; w++) warp_off += ws[w];
    int run = tile_offsets[blockIdx.x] + warp_off + x - s;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        long long i = base + k;
        if (i < n) out[i] = run;
        run += v[k];
        if (i == (long long)n - 1) out[n] = run;
    }
}

__global__ void scan_empty(int *out) { out[0] = 0; }

int exclusive_scan_i32(const int *in, int *out, int n, cudaStream_t s) {
    if (n < 0) return set_error(MLAMG_EINVAL, "scan: n < 0");
    if (n == 0) {
        scan_empty<<<1, 1, 0, s>>>(out);
        MLAMG_LAUNCHED();
        return MLAMG_OK;
    }
    const int nt = (int)cdiv(n, SCAN_TILE);
    Scratch tiles((size_t)nt * sizeof(int), s);
    MLAMG_SCRATCH_OK(tiles);
    scan_tile_sums<<<nt, SCAN_THREADS, 0, s>>>(in, n, tiles.as<int>());
    MLAMG_LAUNCHED();
    scan_tile_offsets<<<1, 1024, 0, s>>>(tiles.as<int>(), nt);
    MLAMG_LAUNCHED();
    scan_tiles<<<nt, SCAN_THREADS, 0, s>>>(in, out, n, tiles.as<int>());
    MLAMG_LAUNCHED();
    return MLAMG_OK;
}

// ------------------------------------------------------------------ row binning
__global__ void __launch_bounds__(256) bin_count_kernel(int m, const int *__restrict__ binid,
                                                        int *__restrict__ counts) {
    __shared__ int sc[NB];
    if (threadIdx.x < NB) sc[threadIdx.x] = 0;
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) {
        const int b = binid[i];
        if (b >= 0) atomicAdd(&sc[b], 1);
    }
    __syncthreads();
    if (threadIdx.x < NB && sc[threadIdx.x]) atomicAdd(&counts[threadIdx.x], sc[threadIdx.x]);
}

// one global atomic per (CTA, bin) instead of one per row: with 16.7 M rows falling into one or two bins the per-row
// atomics serialised on five addresses (ncu: 2.4 ms per call, 25 % of the whole setup's kernel time)
__global__ void __launch_bounds__(256) bin_fill_kernel(int m, const int *__restrict__ binid,
                                                       int *__restrict__ cursors, int *__restrict__ rows) {
    __shared__ int sc[NB], sbase[NB];
    if (threadIdx.x < NB) sc[threadIdx.x] = 0;
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = i < m ? binid[i] : -1;
    int local = 0;
    if (b >= 0) local = atomicAdd(&sc[b], 1);          // shared-memory atomic: rank inside the CTA (order is arbitrary)
    __syncthreads();
    if (threadIdx.x < NB && sc[threadIdx.x]) sbase[threadIdx.x] = atomicAdd(&cursors[threadIdx.x], sc[threadIdx.x]);
    __syncthreads();
    if (b >= 0) rows[sbase[b] + local] = (int)i;
}

int partition_rows_by_bin(int m, const int *binid, int *rows, Bins *bins, cudaStream_t s) {
    Scratch cnt(2 * NB * sizeof(int), s);
    MLAMG_SCRATCH_OK(cnt);
    int *counts = cnt.as<int>();
    int *cursors = counts + NB;
    MLAMG_CUDA(cudaMemsetAsync(counts, 0, 2 * NB * sizeof(int), s));
    bin_count_kernel<<<cdiv(m, 256), 256, 0, s>>>(m, binid, counts);
    MLAMG_LAUNCHED();
    MLAMG_CUDA(cudaMemcpyAsync(bins->counts, counts, NB * sizeof(int), cudaMemcpyDeviceToHost, s));
    MLAMG_CUDA(cudaStreamSynchronize(s));
    bins->offsets[0] = 0;
    for (int b = 0; b < NB; b++) bins->offsets[b + 1] = bins->offsets[b] + bins->counts[b];
    MLAMG_CUDA(cudaMemcpyAsync(cursors, bins->offsets, NB * sizeof(int), cudaMemcpyHostToDevice, s));
    bin_fill_kernel<<<cdiv(m, 256), 256, 0, s>>>(m, binid, cursors, rows);
    MLAMG_LAUNCHED();
    return MLAMG_OK;
}

// ------------------------------------------------------------------ deterministic reductions
__global__ void __launch_bounds__(1024) reduce_partials_kernel(const double *__restrict__ partial, int n,
                                                               double *__restrict__ result) {
    __shared__ double sm[32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += 1024) acc += partial[i];
    acc = block_sum(acc, sm);
    if (threadIdx.x == 0) *result = acc;
}

int reduce_partials(const double *partial, int n, double *result, cudaStream_t s) {
    reduce_partials_kernel<<<1, 1024, 0, s>>>(partial, n, result);
    MLAMG_LAUNCHED();
    return MLAMG_OK;
}

constexpr int BLAS_THREADS = 256;
constexpr int BLAS_MAX_BLOCKS = 148 * 8;

template <typename T>
__global__ void __launch_bounds__(BLAS_THREADS) dot_kernel(int n, const T *__restrict__ x,
                                                           const T *__restrict__ y,
                                                           double *__restrict__ partial) {
    __shared__ double sm[32];
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        acc += (double)x[i] * (double)y[i];
    acc = block_sum(acc, sm);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

template <typename T>
__global__ void __launch_bounds__(BLAS_THREADS) axpby_kernel(int n, T alpha, const T *__restrict__ x, T beta,
                                                             T *__restrict__ y) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const T yi = (beta == (T)0) ? (T)0 : beta * y[i];
        y[i] = alpha * x[i] + yi;
    }
}

}  // namespace mlamg

using namespace mlamg;

extern "C" {

const char *mlamg_last_error(void) { return g_err; }
int mlamg_version(void) { return 100; }
long long mlamg_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int mlamg_scan_i32(const int *in, int *out, int n, mlamg_stream_t stream) {
    return exclusive_scan_i32(in, out, n, as_stream(stream));
}

int mlamg_dot(int dtype, int n, const void *x, const void *y, double *result, mlamg_stream_t stream) {
    cudaStream_t s = as_stream(stream);
    if (n < 0) return set_error(MLAMG_EINVAL, "dot: n < 0");
    int blocks = (int)cdiv(n > 0 ? n : 1, BLAS_THREADS);
    if (blocks > BLAS_MAX_BLOCKS) blocks = BLAS_MAX_BLOCKS;
    Scratch part((size_t)blocks * sizeof(double), s);
    MLAMG_SCRATCH_OK(part);
    MLAMG_DISPATCH(dtype, (dot_kernel<T><<<blocks, BLAS_THREADS, 0, s>>>(n, (const T *)x, (const T *)y,
                                                                         part.as<double>())));
    MLAMG_LAUNCHED();
    return reduce_partials(part.as<double>(), blocks, result, s);
}

int mlamg_axpby(int dtype, int n, double alpha, const void *x, double beta, void *y, mlamg_stream_t stream) {
    cudaStream_t s = as_stream(stream);
    if (n <= 0) return n == 0 ? MLAMG_OK : set_error(MLAMG_EINVAL, "axpby: n < 0");
    int blocks = (int)cdiv(n, BLAS_THREADS);
    if (blocks > BLAS_MAX_BLOCKS) blocks = BLAS_MAX_BLOCKS;
    MLAMG_DISPATCH(dtype, (axpby_kernel<T><<<blocks, BLAS_THREADS, 0, s>>>(n, (T)alpha, (const T *)x, (T)beta,
                                                                           (T *)y)));
    MLAMG_LAUNCHED();
    return MLAMG_OK;
}

}  // extern "C"
