// |lambda_max(D^-1 A)| with a reported accuracy — the device replacement for the ARPACK call of
// ns/lib/multigrid.py:105 (`eigs(Dinv @ A, k=1, which='LM')`, 60-160 s on the host at 128^3).
//
//   symmetric A, positive diagonal : Lanczos on B = D^-1/2 A D^-1/2 (same spectrum as D^-1 A).  B's values are formed
//       once on A's pattern (b_ij = a_ij s_i s_j), every step is one fast CSR SpMV plus two fused vector passes; the
//       Lanczos scalars alpha_j, beta_j never leave the device.  Every `check` steps the tridiagonal T_j is brought
//       to the host (16 j bytes) and its extreme Ritz pair computed by implicit QL; the run stops when the
//       eigenvalue error estimate  min(res, res^2/gap)  (res = beta_j |s_j|, the exact residual norm of the Ritz
//       pair; gap = distance to the next Ritz value)  drops below tol * theta.
//   anything else : power iteration on D^-1 A, stopped by the Rayleigh residual ||D^-1 A x - theta x|| <= tol theta.
//
// Returns the eigenvalue, the achieved relative residual and the number of operator applications.
#include <math.h>
#include <vector>
#include <algorithm>
#include "common.cuh"
#include "tridiag.h"

namespace mlamg {

template <typename T> int spmv_t(int, long long, const int *, const int *, const T *, const T *, T *, cudaStream_t);

constexpr int EIG_THREADS = 256;

// s_i = 1/sqrt(a_ii); flag[0] counts rows whose diagonal is missing or not positive
template <typename T>
__global__ void __launch_bounds__(EIG_THREADS) eig_diag_scale_kernel(int n, const int *__restrict__ rowptr,
                                                                      const int *__restrict__ col, const T *__restrict__ val,
                                                                      T *__restrict__ s, T *__restrict__ dinv,
                                                                      int *__restrict__ flag) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double d = 0.0;
    for (int j = rowptr[i]; j < rowptr[i + 1]; j++)
        if (col[j] == i) d += (double)val[j];
    if (!(d > 0.0)) atomicAdd(flag, 1);
    s[i] = d > 0.0 ? (T)rsqrt(d) : (T)0;
    dinv[i] = d != 0.0 ? (T)(1.0 / d) : (T)0;
}

// out_ij = a_ij * r_i * c_j on A's pattern (B = S A S, or D^-1 A with c = 1)
template <typename T>
__global__ void __launch_bounds__(EIG_THREADS) eig_scale_values_kernel(int n, const int *__restrict__ rowptr,
                                                                        const int *__restrict__ col, const T *__restrict__ val,
                                                                        const T *__restrict__ r, const T *__restrict__ c,
                                                                        T *__restrict__ out) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long row = t >> 3;
    const int lane = (int)(t & 7);
    if (row >= n) return;
    const T ri = r[row];
    for (int j = rowptr[row] + lane; j < rowptr[row + 1]; j += 8)
        out[j] = val[j] * ri * (c ? c[col[j]] : (T)1);
}

template <typename T>
__global__ void __launch_bounds__(EIG_THREADS) eig_init_kernel(int n, T *__restrict__ x, unsigned salt,
                                                                double *__restrict__ partial) {
    __shared__ double sm[32];
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        unsigned h = ((unsigned)i + salt) * 2654435761u;     // fixed pseudo-random start: deterministic, components on
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;        // every eigenvector
        const double v = (0.5 + (double)(h & 0xffffu) / 65536.0) * ((h & 0x10000u) ? -1.0 : 1.0);
        x[i] = (T)v;
        acc += (double)(T)v * (double)(T)v;
    }
    acc = block_sum(acc, sm);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

// partial[b] = sum over the block's slice of x.y
template <typename T>
__global__ void __launch_bounds__(EIG_THREADS) eig_dot_kernel(int n, const T *__restrict__ x, const T *__restrict__ y,
                                                               double *__restrict__ partial) {
    __shared__ double sm[32];
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        acc += (double)x[i] * (double)y[i];
    acc = block_sum(acc, sm);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

// x *= rsqrt(*ss)
template <typename T>
__global__ void __launch_bounds__(EIG_THREADS) eig_normalize_kernel(int n, T *__restrict__ x, const double *__restrict__ ss) {
    const double inv = rsqrt(*ss);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        x[i] = (T)((double)x[i] * inv);
}

// Lanczos step, second half: w <- w - alpha v - beta v_prev, partial ||w||^2  (alpha = *alpha_p, beta = *beta_p or 0)
template <typename T>
__global__ void __launch_bounds__(EIG_THREADS) eig_lanczos_update_kernel(int n, T *__restrict__ w, const T *__restrict__ v,
                                                                          const T *__restrict__ vprev,
                                                                          const double *__restrict__ alpha_p,
                                                                          const double *__restrict__ beta_p,
                                                                          double *__restrict__ partial) {
    __shared__ double sm[32];
    const double alpha = *alpha_p;
    const double beta = beta_p ? *beta_p : 0.0;
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double wi = (double)w[i] - alpha * (double)v[i];
        if (beta_p) wi -= beta * (double)vprev[i];
        const T wt = (T)wi;
        w[i] = wt;
        acc += (double)wt * (double)wt;
    }
    acc = block_sum(acc, sm);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

// *out = sum(partial) (fixed order); sqrt_out (optional) = sqrt of it
__global__ void __launch_bounds__(1024) eig_reduce_kernel(const double *__restrict__ partial, int nb, double *__restrict__ out,
                                                           double *__restrict__ sqrt_out) {
    __shared__ double sm[32];
    double a = 0.0;
    for (int i = threadIdx.x; i < nb; i += 1024) a += partial[i];
    a = block_sum(a, sm);
    if (threadIdx.x == 0) {
        *out = a;
        if (sqrt_out) *sqrt_out = sqrt(a);
    }
}

// v_next = w / *beta   (written over vprev's storage by the caller's pointer rotation)
template <typename T>
__global__ void __launch_bounds__(EIG_THREADS) eig_scale_into_kernel(int n, const T *__restrict__ w, const double *__restrict__ beta,
                                                                      T *__restrict__ out) {
    const double b = *beta;
    const double inv = b > 0.0 ? 1.0 / b : 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = (T)((double)w[i] * inv);
}

template <typename T>
static int lambda_max_lanczos_t(int n, long long nnz, const int *rowptr, const int *col, const T *val, double tol,
                                int max_steps, int symmetric, double *lambda_host, double *resid_host, int *steps_host,
                                int *method_host, cudaStream_t s) {
    if (n <= 0 || max_steps <= 0 || !(tol >= 0.0)) return set_error(MLAMG_EINVAL, "lambda_max: bad n / max_steps / tol");
    unsigned eb = cdiv(n, EIG_THREADS);
    if (eb > 148u * 8u) eb = 148u * 8u;
    Scratch sc((size_t)n * sizeof(T), s), dinv((size_t)n * sizeof(T), s), bval((size_t)std::max<long long>(nnz, 1) * sizeof(T), s);
    Scratch v0((size_t)n * sizeof(T), s), v1((size_t)n * sizeof(T), s), w((size_t)n * sizeof(T), s);
    Scratch part((size_t)eb * sizeof(double), s), scal((size_t)(2 * (max_steps + 2) + 8) * sizeof(double), s), flag(sizeof(int), s);
    MLAMG_SCRATCH_OK(sc); MLAMG_SCRATCH_OK(dinv); MLAMG_SCRATCH_OK(bval); MLAMG_SCRATCH_OK(v0); MLAMG_SCRATCH_OK(v1);
    MLAMG_SCRATCH_OK(w); MLAMG_SCRATCH_OK(part); MLAMG_SCRATCH_OK(scal); MLAMG_SCRATCH_OK(flag);
    double *alpha = scal.as<double>();                 // alpha[j], j < max_steps
    double *beta = alpha + (max_steps + 2);            // beta[j] = ||w_j|| after step j
    double *tmp = beta + (max_steps + 2);              // tmp[0..8)
    MLAMG_CUDA(cudaMemsetAsync(flag.p, 0, sizeof(int), s));
    eig_diag_scale_kernel<T><<<cdiv(n, EIG_THREADS), EIG_THREADS, 0, s>>>(n, rowptr, col, val, sc.as<T>(), dinv.as<T>(), flag.as<int>());
    MLAMG_LAUNCHED();
    int bad_diag = 0;
    MLAMG_CUDA(cudaMemcpyAsync(&bad_diag, flag.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    MLAMG_CUDA(cudaStreamSynchronize(s));
    T *x = v0.as<T>(), *y = v1.as<T>(), *wv = w.as<T>();
    const unsigned vb = cdiv((long long)n * 8, EIG_THREADS);

    if (symmetric != 0 && bad_diag == 0) {
        eig_scale_values_kernel<T><<<vb, EIG_THREADS, 0, s>>>(n, rowptr, col, val, sc.as<T>(), sc.as<T>(), bval.as<T>());
        MLAMG_LAUNCHED();
        if (symmetric < 0) {
            // probabilistic symmetry test: x.(B y) == y.(B x) for two fixed pseudo-random vectors
            eig_init_kernel<T><<<eb, EIG_THREADS, 0, s>>>(n, x, 12345u, part.as<double>());
            MLAMG_LAUNCHED();
            eig_init_kernel<T><<<eb, EIG_THREADS, 0, s>>>(n, y, 987654321u, part.as<double>());
            MLAMG_LAUNCHED();
            MLAMG_TRY(spmv_t<T>(n, nnz, rowptr, col, bval.as<T>(), y, wv, s));
            eig_dot_kernel<T><<<eb, EIG_THREADS, 0, s>>>(n, x, wv, part.as<double>());
            MLAMG_LAUNCHED();
            eig_reduce_kernel<<<1, 1024, 0, s>>>(part.as<double>(), (int)eb, tmp + 0, nullptr);
            MLAMG_LAUNCHED();
            MLAMG_TRY(spmv_t<T>(n, nnz, rowptr, col, bval.as<T>(), x, wv, s));
            eig_dot_kernel<T><<<eb, EIG_THREADS, 0, s>>>(n, y, wv, part.as<double>());
            MLAMG_LAUNCHED();
            eig_reduce_kernel<<<1, 1024, 0, s>>>(part.as<double>(), (int)eb, tmp + 1, nullptr);
            MLAMG_LAUNCHED();
            eig_dot_kernel<T><<<eb, EIG_THREADS, 0, s>>>(n, wv, wv, part.as<double>());
            MLAMG_LAUNCHED();
            eig_reduce_kernel<<<1, 1024, 0, s>>>(part.as<double>(), (int)eb, tmp + 2, nullptr);
            MLAMG_LAUNCHED();
            double h[3];
            MLAMG_CUDA(cudaMemcpyAsync(h, tmp, 3 * sizeof(double), cudaMemcpyDeviceToHost, s));
            MLAMG_CUDA(cudaStreamSynchronize(s));
            const double eps = sizeof(T) == 8 ? 1e-12 : 1e-4;
            // |x.By - y.Bx| against ||x|| ||Bx|| (||x||^2 ~ n/3 for the start vectors)
            symmetric = fabs(h[0] - h[1]) <= eps * sqrt(h[2] * (double)n) ? 1 : 0;
        }
    } else {
        symmetric = 0;
    }

    if (symmetric == 1) {
        // ---- Lanczos on B
        T *vprev = y, *v = x;
        eig_init_kernel<T><<<eb, EIG_THREADS, 0, s>>>(n, v, 0u, part.as<double>());
        MLAMG_LAUNCHED();
        eig_reduce_kernel<<<1, 1024, 0, s>>>(part.as<double>(), (int)eb, tmp + 0, nullptr);
        MLAMG_LAUNCHED();
        eig_normalize_kernel<T><<<eb, EIG_THREADS, 0, s>>>(n, v, tmp + 0);
        MLAMG_LAUNCHED();
        std::vector<double> ha, hb, d, e, z;
        int done = 0, next_check = std::min(max_steps, 20);
        double theta = 0.0, res_rel = 1.0;
        for (int j = 0; j < max_steps; j++) {
            MLAMG_TRY(spmv_t<T>(n, nnz, rowptr, col, bval.as<T>(), v, wv, s));
            eig_dot_kernel<T><<<eb, EIG_THREADS, 0, s>>>(n, v, wv, part.as<double>());
            MLAMG_LAUNCHED();
            eig_reduce_kernel<<<1, 1024, 0, s>>>(part.as<double>(), (int)eb, alpha + j, nullptr);
            MLAMG_LAUNCHED();
            eig_lanczos_update_kernel<T><<<eb, EIG_THREADS, 0, s>>>(n, wv, v, vprev, alpha + j, j > 0 ? beta + (j - 1) : nullptr,
                                                                    part.as<double>());
            MLAMG_LAUNCHED();
            eig_reduce_kernel<<<1, 1024, 0, s>>>(part.as<double>(), (int)eb, tmp + 1, beta + j);
            MLAMG_LAUNCHED();
            done = j + 1;
            if (done == next_check || done == max_steps) {
                ha.resize(done); hb.resize(done);
                MLAMG_CUDA(cudaMemcpyAsync(ha.data(), alpha, done * sizeof(double), cudaMemcpyDeviceToHost, s));
                MLAMG_CUDA(cudaMemcpyAsync(hb.data(), beta, done * sizeof(double), cudaMemcpyDeviceToHost, s));
                MLAMG_CUDA(cudaStreamSynchronize(s));
                d = ha;
                e.assign(hb.begin(), hb.end() - 1);
                if (!tridiag_ql_last_row(d, e, z)) return set_error(MLAMG_ECUDA, "lambda_max: tridiagonal QL did not converge");
                int k1 = 0;
                for (int k = 1; k < done; k++) if (fabs(d[k]) > fabs(d[k1])) k1 = k;
                double gap = 1e300;
                for (int k = 0; k < done; k++) if (k != k1) gap = std::min(gap, fabs(fabs(d[k1]) - fabs(d[k])));
                theta = fabs(d[k1]);
                const double res = hb[done - 1] * fabs(z[k1]);
                const double est = std::min(res, gap > 0.0 ? res * res / gap : res);
                res_rel = theta > 0.0 ? res / theta : 0.0;
                if (est <= tol * theta || hb[done - 1] <= 1e-300) break;          // converged, or an invariant subspace
                next_check = std::min(max_steps, done + std::max(10, done / 4));
            }
            // v_next = w / beta_j into vprev's storage, rotate
            eig_scale_into_kernel<T><<<eb, EIG_THREADS, 0, s>>>(n, wv, beta + j, vprev);
            MLAMG_LAUNCHED();
            std::swap(v, vprev);
        }
        *lambda_host = theta;
        if (resid_host) *resid_host = res_rel;
        if (steps_host) *steps_host = done;
        if (method_host) *method_host = 1;
        return MLAMG_OK;
    }

    // ---- monitored power iteration on D^-1 A (non-symmetric operator or a diagonal that is not positive)
    eig_scale_values_kernel<T><<<vb, EIG_THREADS, 0, s>>>(n, rowptr, col, val, dinv.as<T>(), (const T *)nullptr, bval.as<T>());
    MLAMG_LAUNCHED();
    eig_init_kernel<T><<<eb, EIG_THREADS, 0, s>>>(n, x, 0u, part.as<double>());
    MLAMG_LAUNCHED();
    eig_reduce_kernel<<<1, 1024, 0, s>>>(part.as<double>(), (int)eb, tmp + 0, nullptr);
    MLAMG_LAUNCHED();
    eig_normalize_kernel<T><<<eb, EIG_THREADS, 0, s>>>(n, x, tmp + 0);
    MLAMG_LAUNCHED();
    double theta = 0.0, res_rel = 1.0;
    int done = 0, next_check = std::min(max_steps, 20);
    for (int it = 0; it < max_steps; it++) {
        MLAMG_TRY(spmv_t<T>(n, nnz, rowptr, col, bval.as<T>(), x, y, s));
        eig_dot_kernel<T><<<eb, EIG_THREADS, 0, s>>>(n, x, y, part.as<double>());
        MLAMG_LAUNCHED();
        eig_reduce_kernel<<<1, 1024, 0, s>>>(part.as<double>(), (int)eb, tmp + 1, nullptr);     // x.y (||x|| = 1)
        MLAMG_LAUNCHED();
        eig_dot_kernel<T><<<eb, EIG_THREADS, 0, s>>>(n, y, y, part.as<double>());
        MLAMG_LAUNCHED();
        eig_reduce_kernel<<<1, 1024, 0, s>>>(part.as<double>(), (int)eb, tmp + 2, nullptr);     // y.y
        MLAMG_LAUNCHED();
        done = it + 1;
        if (done == next_check || done == max_steps) {
            double h[2];
            MLAMG_CUDA(cudaMemcpyAsync(h, tmp + 1, 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
            MLAMG_CUDA(cudaStreamSynchronize(s));
            theta = fabs(h[0]);
            const double r2 = std::max(0.0, h[1] - h[0] * h[0]);        // ||y - theta x||^2 with ||x|| = 1
            res_rel = theta > 0.0 ? sqrt(r2) / theta : 0.0;
            // for a dominant complex pair the Rayleigh quotient does not settle: sqrt(y.y) still bounds |lambda|
            if (res_rel <= tol) break;
            if (done == max_steps) theta = std::max(theta, sqrt(h[1]));
            next_check = std::min(max_steps, done + std::max(10, done / 4));
        }
        eig_normalize_kernel<T><<<eb, EIG_THREADS, 0, s>>>(n, y, tmp + 2);
        MLAMG_LAUNCHED();
        std::swap(x, y);
    }
    *lambda_host = theta;
    if (resid_host) *resid_host = res_rel;
    if (steps_host) *steps_host = done;
    if (method_host) *method_host = 0;
    return MLAMG_OK;
}

}  // namespace mlamg

using namespace mlamg;

extern "C" {

int mlamg_lambda_max(int dtype, int n, long long nnz, const int *rowptr, const int *col, const void *val, double tol,
                         int max_steps, int symmetric, double *lambda_host, double *resid_host, int *steps_host,
                         int *method_host, mlamg_stream_t stream) {
    MLAMG_DISPATCH(dtype, return lambda_max_lanczos_t<T>(n, nnz, rowptr, col, (const T *)val, tol, max_steps, symmetric,
                                                         lambda_host, resid_host, steps_host, method_host,
                                                         as_stream(stream)));
    return MLAMG_OK;
}

}  // extern "C"
