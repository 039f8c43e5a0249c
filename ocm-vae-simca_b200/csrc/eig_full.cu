// ALL eigenvalues of a symmetric fp64 matrix: Householder tridiagonalisation + bisection on the Sturm sequence.
//
// The reference keeps every eigenvalue of the class covariance under `_model[cls]["eigs_all"]`
// (/root/reference/utils/SIMCA.py:66, :89: explained_variance_ of a full PCA).  The fit itself only needs the leading
// pairs and three power sums of the rest (csrc/eig_sub.cu), so this runs lazily, the first time somebody reads the whole
// array.  Round 1 used cuSOLVER (syevd through torch) here; this is the repo's own solver:
//   * tridiagonalisation, unblocked Householder: per column k one reflector v (1 CTA), p = beta A v over the trailing
//     block (one warp per row, coalesced), and the symmetric rank-2 update A -= v w^T + w v^T -- three launches per
//     column, the p x p matrix stays in L2 (32 MB at p = 2048);
//   * eigenvalues of the tridiagonal (d, e) by bisection: thread i brackets the i-th smallest eigenvalue with the Sturm
//     count (LAPACK dstebz's recurrence with its pivmin safeguard); every eigenvalue is independent and good to
//     a few ulp of |T|, which is what the limits need (they use sums of the tail).
#include "ocm_common.cuh"

namespace ocm {
namespace efull {

// ---- step k, part 1: reflector of x = A[k][k+1 .. p-1] (row k == column k, contiguous) -------------------------
// v = x - alpha e1, alpha = -sign(x0) |x|, beta = 2 / |v|^2 (0 when x is already zero below the first entry).
// d[k] = A[k][k], e[k] = alpha.  One CTA.
__global__ void __launch_bounds__(1024)
house_vector_kernel(const double* __restrict__ A, int p, int k, double* __restrict__ v, double* __restrict__ beta_out,
                    double* __restrict__ d, double* __restrict__ e) {
    __shared__ double red[32];
    __shared__ double s_alpha, s_beta;
    const int m = p - k - 1;
    const double* x = A + (size_t)k * p + k + 1;
    double ss = 0.0;
    for (int i = threadIdx.x; i < m; i += blockDim.x) ss = fma(x[i], x[i], ss);
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x < 32) {
        double t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
        t = warp_sum(t);
        if (threadIdx.x == 0) {
            const double x0 = x[0];
            const double tail = t - x0 * x0;                  // |x|^2 without its first entry
            double alpha, beta;
            if (tail <= 0.0) {                                // nothing to annihilate
                alpha = x0; beta = 0.0;
            } else {
                const double nrm = sqrt(t);
                alpha = x0 >= 0.0 ? -nrm : nrm;
                const double v0 = x0 - alpha;
                beta = 2.0 / (v0 * v0 + tail);
            }
            s_alpha = alpha; s_beta = beta;
            d[k] = A[(size_t)k * p + k];
            e[k] = alpha;
            *beta_out = beta;
        }
    }
    __syncthreads();
    const double alpha = s_alpha;
    for (int i = threadIdx.x; i < m; i += blockDim.x) v[i] = (i == 0) ? x[0] - alpha : x[i];
}

// ---- part 2: pv = beta * A22 v  (A22 = trailing (m x m) block), one warp per row ----------------------------------
__global__ void __launch_bounds__(256)
symv_kernel(const double* __restrict__ A, int p, int k, const double* __restrict__ v, const double* __restrict__ beta_in,
            double* __restrict__ pv) {
    const int m = p - k - 1;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= m) return;
    const int lane = threadIdx.x & 31;
    const double* a = A + (size_t)(k + 1 + row) * p + k + 1;
    double s = 0.0;
    for (int j = lane; j < m; j += 32) s = fma(a[j], v[j], s);
    s = warp_sum(s);
    if (lane == 0) pv[row] = *beta_in * s;
}

// ---- part 3: w = pv - (beta/2)(pv . v) v ; A22 -= v w^T + w v^T --------------------------------------------------
// every CTA recomputes the scalar (a dot product of length m) instead of paying for another launch
__global__ void __launch_bounds__(256)
syr2_kernel(double* __restrict__ A, int p, int k, const double* __restrict__ v, const double* __restrict__ pv,
            const double* __restrict__ beta_in) {
    __shared__ double red[8];
    __shared__ double s_K;
    const int m = p - k - 1;
    const double beta = *beta_in;
    if (beta == 0.0) return;
    double dot = 0.0;
    for (int i = threadIdx.x; i < m; i += blockDim.x) dot = fma(pv[i], v[i], dot);
    dot = warp_sum(dot);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < 8; ++i) t += red[i];
        s_K = 0.5 * beta * t;
    }
    __syncthreads();
    const double K = s_K;
    // tile: 32 rows x all columns handled by the CTA's 8 warps (warp -> 4 rows), lanes stride columns
    const int r0 = blockIdx.x * 32 + (threadIdx.x >> 5) * 4;
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
        const int i = r0 + rr;
        if (i >= m) break;
        const double vi = v[i], wi = pv[i] - K * vi;
        double* a = A + (size_t)(k + 1 + i) * p + k + 1;
        for (int j = lane; j < m; j += 32) {
            const double vj = v[j], wj = pv[j] - K * vj;
            a[j] -= vi * wj + wi * vj;
        }
    }
}

__global__ void tridiag_tail_kernel(const double* __restrict__ A, int p, double* __restrict__ d, double* __restrict__ e) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        d[p - 1] = A[(size_t)(p - 1) * p + p - 1];
        e[p - 1] = 0.0;
    }
}

// ---- bisection: thread i -> i-th smallest eigenvalue of the symmetric tridiagonal (d, e[0..p-2]) -------------------
__global__ void __launch_bounds__(128)
bisect_kernel(const double* __restrict__ d, const double* __restrict__ e, int p, double* __restrict__ w_desc) {
    extern __shared__ double sm_b[];
    double* sd = sm_b;
    double* se2 = sm_b + p;
    __shared__ double s_lo, s_hi, s_piv;
    for (int i = threadIdx.x; i < p; i += blockDim.x) {
        sd[i] = d[i];
        se2[i] = i + 1 < p ? e[i] * e[i] : 0.0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double lo = 1e300, hi = -1e300, emax = 0.0;
        for (int i = 0; i < p; ++i) {
            const double r = (i > 0 ? sqrt(se2[i - 1]) : 0.0) + (i + 1 < p ? sqrt(se2[i]) : 0.0);
            lo = fmin(lo, sd[i] - r);
            hi = fmax(hi, sd[i] + r);
            emax = fmax(emax, se2[i]);
        }
        const double tn = fmax(fabs(lo), fabs(hi));
        s_lo = lo - 2.2e-16 * tn * p - 1e-300;
        s_hi = hi + 2.2e-16 * tn * p + 1e-300;
        s_piv = fmax(2.2250738585072014e-308 * fmax(emax, 1.0), 1e-300);
    }
    __syncthreads();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;      // eigenvalue index, ascending
    if (idx >= p) return;
    const double pivmin = s_piv;
    double lo = s_lo, hi = s_hi;
    for (int it = 0; it < 200; ++it) {
        const double mid = 0.5 * (lo + hi);
        if (mid <= lo || mid >= hi) break;                       // the interval is one ulp wide
        // Sturm count: number of eigenvalues < mid
        int cnt = 0;
        double q = sd[0] - mid;
        if (fabs(q) < pivmin) q = -pivmin;
        cnt += q < 0.0;
        for (int i = 1; i < p; ++i) {
            q = sd[i] - mid - se2[i - 1] / q;
            if (fabs(q) < pivmin) q = -pivmin;
            cnt += q < 0.0;
        }
        if (cnt > idx) hi = mid; else lo = mid;
    }
    w_desc[p - 1 - idx] = 0.5 * (lo + hi);
}

}  // namespace efull
}  // namespace ocm

// A [p][p] symmetric fp64 is OVERWRITTEN (pass a copy); work: >= 2 p + 1 doubles of device scratch followed by
// nothing else; d, e [p] receive the tridiagonal; w [p] the eigenvalues in DESCENDING order.
extern "C" int ocm_eigvals_sym_f64(double* A, int32_t p, double* work, double* d, double* e, double* w, void* stream) {
    using namespace ocm;
    if (p == 0) return 0;
    OCM_REQUIRE(A && work && d && e && w && p > 0, OCM_EINVAL, "NULL buffer or bad shape");
    OCM_REQUIRE(p <= 12000, OCM_ELIMIT, "matrix too large for the shared-memory bisection (p <= 12000)");
    cudaStream_t st = (cudaStream_t)stream;
    double* v = work;
    double* pv = work + p;
    double* beta = work + 2 * (size_t)p;
    for (int k = 0; k + 1 < p; ++k) {
        const int m = p - k - 1;
        efull::house_vector_kernel<<<1, 1024, 0, st>>>(A, p, k, v, beta, d, e);
        if (m > 1) {       // a 1 x 1 trailing block: the reflector is the identity (beta = 0) and nothing is left to update
            efull::symv_kernel<<<(m + 7) / 8, 256, 0, st>>>(A, p, k, v, beta, pv);
            efull::syr2_kernel<<<(m + 31) / 32, 256, 0, st>>>(A, p, k, v, pv, beta);
        }
    }
    int rc = check_launch("tridiagonalisation kernels");
    if (rc) return rc;
    count_launch(3 * (p > 1 ? p - 1 : 0) - 1);
    efull::tridiag_tail_kernel<<<1, 32, 0, st>>>(A, p, d, e);
    const size_t smem = sizeof(double) * 2 * (size_t)p;
    if (smem > 48 * 1024) OCM_CUDA(cudaFuncSetAttribute(efull::bisect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    efull::bisect_kernel<<<(p + 127) / 128, 128, smem, st>>>(d, e, p, w);
    return check_launch("bisect_kernel");
}
