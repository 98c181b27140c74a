// Shared helpers for the mlamg sm_100a kernels (error plumbing, warp/block reductions,
// stream-ordered scratch buffers, the int32 prefix scan every two-phase setup kernel needs).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/mlamg.h"

namespace mlamg {

int set_error(int code, const char *fmt, ...);
int set_cuda_error(cudaError_t e, const char *file, int line);
void count_launch(int n = 1);

#define MLAMG_CUDA(call)                                                           \
    do {                                                                           \
        cudaError_t e__ = (call);                                                  \
        if (e__ != cudaSuccess) return ::mlamg::set_cuda_error(e__, __FILE__, __LINE__); \
    } while (0)

// after a kernel launch
#define MLAMG_LAUNCHED()                                                           \
    do {                                                                           \
        ::mlamg::count_launch();                                                   \
        cudaError_t e__ = cudaGetLastError();                                      \
        if (e__ != cudaSuccess) return ::mlamg::set_cuda_error(e__, __FILE__, __LINE__); \
    } while (0)

#define MLAMG_TRY(expr)                    \
    do {                                   \
        int rc__ = (expr);                 \
        if (rc__ != MLAMG_OK) return rc__; \
    } while (0)

#define MLAMG_DISPATCH(dtype, ...)                                       \
    switch (dtype) {                                                     \
        case MLAMG_F32: { using T = float; __VA_ARGS__; } break;         \
        case MLAMG_F64: { using T = double; __VA_ARGS__; } break;        \
        default: return ::mlamg::set_error(MLAMG_EINVAL, "bad dtype %d", (int)(dtype)); \
    }

static inline cudaStream_t as_stream(mlamg_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
static inline unsigned cdiv(long long a, long long b) { return (unsigned)((a + b - 1) / b); }

// The default memory pool returns everything to the driver at every synchronisation point unless its release
// threshold is raised: a PCG iteration (three host syncs for its dot products) then re-maps its scratch buffers from
// scratch each time (measured: 68 ms per iteration at 128^3 instead of well under 1 ms).  Called once per device.
void keep_pool_memory();

// Stream-ordered scratch allocation (cudaMallocAsync pool: no device sync after warm-up).
struct Scratch {
    void *p = nullptr;
    cudaStream_t s = nullptr;
    cudaError_t err = cudaSuccess;
    // bytes == SIZE_MAX: no allocation at all (lets a call site skip the alloc/free pair, which would
    // otherwise become two extra nodes per kernel inside a captured CUDA graph)
    Scratch(size_t bytes, cudaStream_t stream) : s(stream) {
        if (bytes == (size_t)-1) return;
        keep_pool_memory();
        err = cudaMallocAsync(&p, bytes ? bytes : 16, s);
        if (err != cudaSuccess) p = nullptr;
    }
    ~Scratch() { if (p) cudaFreeAsync(p, s); }
    template <typename U> U *as() const { return reinterpret_cast<U *>(p); }
    Scratch(const Scratch &) = delete;
    Scratch &operator=(const Scratch &) = delete;
};
#define MLAMG_SCRATCH_OK(sc) \
    do { if (!(sc).p) return ::mlamg::set_cuda_error((sc).err, __FILE__, __LINE__); } while (0)

int exclusive_scan_i32(const int *in, int *out, int n, cudaStream_t s);

// Row binning shared by the SpGEMM and row-sort kernels: binid[i] in [-1, NB) (-1 = skip);
// rows[] receives the row ids grouped by bin (order inside a bin is arbitrary).
constexpr int NB = 5;
struct Bins {
    int counts[NB];
    int offsets[NB + 1];
};
int partition_rows_by_bin(int m, const int *binid, int *rows, Bins *bins, cudaStream_t s);
// *result (device double) = sum of partial[0..n) in fixed order
int reduce_partials(const double *partial, int n, double *result, cudaStream_t s);

// ---------------------------------------------------------------- device helpers
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int LANES, typename T>
__device__ __forceinline__ T group_sum(T v) {
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o, LANES);
    return v;
}

// block-wide sum of doubles in a fixed order; result valid on thread 0.  blockDim.x <= 1024.
__device__ __forceinline__ double block_sum(double v, double *smem32) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum(v);
    if (lane == 0) smem32[wid] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    double t = 0.0;
    if (wid == 0) {
        t = (lane < nw) ? smem32[lane] : 0.0;
        t = warp_sum(t);
    }
    __syncthreads();
    return t;
}

// order-preserving unsigned keys for non-negative / general floats (used by atomic max/min)
__device__ __forceinline__ unsigned long long f2key(double v) {
    unsigned long long u = (unsigned long long)__double_as_longlong(v);
    return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ unsigned long long f2key(float v) {
    unsigned u = (unsigned)__float_as_int(v);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return (unsigned long long)u;
}

// explicitly non-fused arithmetic: the sequential CPU loops of the reference (x86-64, no FMA) round
// every product and every sum separately; kernels that must be bit-identical use these
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }

template <typename T> struct Limits;
template <> struct Limits<float> { static __host__ __device__ float max() { return 3.402823466e+38f; } };
template <> struct Limits<double> { static __host__ __device__ double max() { return 1.7976931348623157e+308; } };

}  // namespace mlamg
