// Batched symmetric eigensolver for SMALL fp64 matrices (m <= 64): parallel cyclic Jacobi, one CTA per matrix.
//
// Used for the Rayleigh-Ritz step of the block subspace iteration that fits all CV folds at once
// (engine.eig_topk_subspace): the projected matrices are 32..64 wide, and the library batched eigensolver costs
// several milliseconds per call at that size, which dominated the CV grid after the full per-fold eigensolves
// were removed.  Two-sided Jacobi with the round-robin (chess tournament) ordering: m/2 disjoint rotations per
// step, m-1 steps per sweep, rotations applied to rows, then columns, then the eigenvector matrix, all in shared
// memory; sweeps until the off-diagonal mass is below 1e-30 of the diagonal mass (quadratic convergence: 6-9 sweeps).
#include "ocm_common.cuh"

namespace ocm {

constexpr int JMAX = 64;

__global__ void __launch_bounds__(256)
jacobi_eigh_kernel(const double* __restrict__ Hin, int m, double* __restrict__ evals, double* __restrict__ evecs) {
    extern __shared__ double sm_j[];                       // A and V, m x (m + 1) each (padded rows)
    const int ldj = m + 1;
    double* A_ = sm_j;
    double* V_ = sm_j + (size_t)m * ldj;
#define A_AT(r, c) A_[(r) * ldj + (c)]
#define V_AT(r, c) V_[(r) * ldj + (c)]
    __shared__ double cs[JMAX / 2][2];
    __shared__ int pp[JMAX / 2], qq[JMAX / 2];
    __shared__ int order[JMAX];
    __shared__ double s_off, s_diag;
    __shared__ double red[8][2];
    const int tid = threadIdx.x, nt = blockDim.x;
    const double* H = Hin + (size_t)blockIdx.x * m * m;
    const int me = (m + 1) & ~1;               // even size for the tournament (a dummy index when m is odd)
    for (int i = tid; i < m * m; i += nt) {
        const int r = i / m, c = i % m;
        A_AT(r, c) = 0.5 * (H[r * m + c] + H[c * m + r]);
        V_AT(r, c) = r == c ? 1.0 : 0.0;
    }
    __syncthreads();
    for (int sweep = 0; sweep < 30; ++sweep) {
        {
            // off-diagonal and diagonal mass, all threads (one thread walking the m^2 / 2 entries cost ~10 us per sweep,
            // a quarter of the kernel at m = 32)
            double off = 0.0, dg = 0.0;
            for (int i = tid; i < m * m; i += nt) {
                const int r = i / m, c = i % m;
                const double v = A_AT(r, c);
                if (c > r) off = fma(v, v, off);
                else if (c == r) dg = fma(v, v, dg);
            }
            off = warp_sum(off); dg = warp_sum(dg);
            if ((tid & 31) == 0) { red[tid >> 5][0] = off; red[tid >> 5][1] = dg; }
            __syncthreads();
            if (tid == 0) {
                double o = 0.0, d = 0.0;
                for (int w = 0; w < (nt >> 5); ++w) { o += red[w][0]; d += red[w][1]; }
                s_off = o; s_diag = d;
            }
        }
        __syncthreads();
        if (s_off <= 1e-30 * s_diag || s_off == 0.0) break;
        for (int step = 0; step < me - 1; ++step) {
            // round-robin pairing: position 0 is fixed, the others rotate
            if (tid < me / 2) {
                auto player = [&](int pos) { return pos == 0 ? 0 : 1 + (pos - 1 + step) % (me - 1); };
                int a = player(tid), b = player(me - 1 - tid);
                if (a > b) { const int t = a; a = b; b = t; }
                pp[tid] = a; qq[tid] = b;
                double c = 1.0, s = 0.0;
                if (b < m) {
                    const double apq = A_AT(a, b);
                    if (apq != 0.0) {
                        const double tau = (A_AT(b, b) - A_AT(a, a)) / (2.0 * apq);
                        const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                        c = 1.0 / sqrt(1.0 + t * t);
                        s = t * c;
                    }
                }
                cs[tid][0] = c; cs[tid][1] = s;
            }
            __syncthreads();
            // A <- J^T A J block by block: the 2 x 2 block (rows of pair k) x (columns of pair l) only needs the rotations
            // of its own two pairs, so the row and the column rotations go in ONE pass (same operations in the same order
            // as rows-then-columns, one barrier and one round trip through shared memory less per step)
            const int np = me / 2;
            for (int i = tid; i < np * np; i += nt) {
                const int k = i / np, l = i % np;
                const int a = pp[k], b = qq[k], e = pp[l], f = qq[l];
                const bool rk = b < m, cl = f < m;                      // a dummy partner (odd m): that rotation is the identity
                const double ck = cs[k][0], sk = cs[k][1], c2 = cs[l][0], s2 = cs[l][1];
                double x00 = A_AT(a, e), x01 = cl ? A_AT(a, f) : 0.0;
                double x10 = rk ? A_AT(b, e) : 0.0, x11 = (rk && cl) ? A_AT(b, f) : 0.0;
                if (rk) {                                                // rows: (a, b) <- (c a - s b, s a + c b)
                    const double t0 = ck * x00 - sk * x10, t1 = sk * x00 + ck * x10; x00 = t0; x10 = t1;
                    const double u0 = ck * x01 - sk * x11, u1 = sk * x01 + ck * x11; x01 = u0; x11 = u1;
                }
                if (cl) {                                                // columns
                    const double t0 = c2 * x00 - s2 * x01, t1 = s2 * x00 + c2 * x01; x00 = t0; x01 = t1;
                    const double u0 = c2 * x10 - s2 * x11, u1 = s2 * x10 + c2 * x11; x10 = u0; x11 = u1;
                }
                A_AT(a, e) = x00;
                if (cl) A_AT(a, f) = x01;
                if (rk) A_AT(b, e) = x10;
                if (rk && cl) A_AT(b, f) = x11;
            }
            // V <- V J
            for (int i = tid; i < np * m; i += nt) {
                const int k = i / m, row = i % m;
                const int a = pp[k], b = qq[k];
                if (b < m) {
                    const double c = cs[k][0], s = cs[k][1];
                    const double vx = V_AT(row, a), vy = V_AT(row, b);
                    V_AT(row, a) = c * vx - s * vy;
                    V_AT(row, b) = s * vx + c * vy;
                }
            }
            __syncthreads();
        }
    }
    // descending order by counting: the rank of eigenvalue i is the number of larger ones (ties: the lower index first)
    if (tid < m) {
        const double di = A_AT(tid, tid);
        int rank = 0;
        const bool ni = di != di;
        for (int j = 0; j < m; ++j) {
            const double dj = A_AT(j, j);
            const bool nj = dj != dj;
            // a total order (NaNs last, ties by index): the ranks are a permutation whatever the matrix held
            const bool before = (nj || ni) ? ((!nj && ni) || (nj == ni && j < tid)) : (dj > di || (dj == di && j < tid));
            rank += before ? 1 : 0;
        }
        order[rank] = tid;
    }
    __syncthreads();
    double* w = evals + (size_t)blockIdx.x * m;
    double* U = evecs + (size_t)blockIdx.x * m * m;
    for (int i = tid; i < m; i += nt) w[i] = A_AT(order[i], order[i]);
    for (int i = tid; i < m * m; i += nt) {
        const int r = i / m, c = i % m;
        U[r * m + c] = V_AT(r, order[c]);           // column c = eigenvector of the c-th largest eigenvalue
    }
}

#undef A_AT
#undef V_AT

}  // namespace ocm

extern "C" int ocm_eigh_small_f64(const double* H, int32_t m, int32_t batch, double* evals, double* evecs, void* stream) {
    using namespace ocm;
    if (batch == 0) return 0;
    OCM_REQUIRE(H && evals && evecs && batch > 0, OCM_EINVAL, "NULL buffer");
    OCM_REQUIRE(m >= 1 && m <= JMAX, OCM_ELIMIT, "matrix size must be 1..64");
    const size_t smem = sizeof(double) * 2 * (size_t)m * (m + 1);
    if (smem > 48 * 1024)
        OCM_CUDA(cudaFuncSetAttribute(jacobi_eigh_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    jacobi_eigh_kernel<<<(unsigned)batch, 256, smem, (cudaStream_t)stream>>>(H, m, evals, evecs);
    return check_launch("jacobi_eigh_kernel");
}
