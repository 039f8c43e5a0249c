// Building blocks of the batched block subspace iteration that produces the PCA loadings: the
// k leading eigenpairs of a BATCH of symmetric fp64 covariance matrices (one per CV fold, or one for a plain fit) and
// the power sums of the rest of the spectrum.  This stands in for the LAPACK SVD that sklearn's PCA runs at
// /root/reference/utils/SIMCA.py:64 and :75 (its S^2/(n-1) are the eigenvalues of this covariance, its Vt rows the
// eigenvectors) and for the eigenvalue tail that /root/reference/utils/SIMCA.py:89,:192-205 feeds to the Q limits.
// Round 1 ran these steps through torch (cuBLAS batched GEMM, cuSOLVER potrf / trsm / geqrf); they are the repo's own
// kernels now.  All matrices are row-major fp64: C [F][p][p], blocks [F][p][kb] (kb <= 64).
//
//   symm_mul        Z = C Q                      the one O(p^2 kb) step: FP64 tensor pipe (mma.sync m8n8k4, DMMA)
//   tn_small        H = A^T B  (kb x kb)         reduction over p, split over CTAs, fp64 atomics on a tiny output
//   chol_solve      L = chol(G), Y <- Y L^-T     CholeskyQR step (in-CTA Cholesky of a kb x kb matrix + row-wise solve)
//   mul_small       Y = Z U    (p x kb)(kb x kb) Ritz rotation
//   ritz_residual   |Z U - X diag(w)| per column
//   deflate         R = C - X_k diag(w_k) X_k^T  then tail sums tr R, |R|_F^2, tr R^3 = sum_ij (R^2)_ij R_ij with R^2
//                   from the Gram kernel (R symmetric: R^T R = R^2), all without cancellation.
#include "ocm_common.cuh"

namespace ocm {
namespace esub {

constexpr int KBMAX = 64;

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// ---------------------------------------------------------------- Z = C Q ------------------------------
// CTA: 64 rows of Z x all kb columns; K (columns of C) in chunks of 32; 8 warps, warp w owns rows 8w..8w+7.
constexpr int SM_ROWS = 64, SM_KC = 32, SM_CP = SM_KC + 4;     // padded pitch of the C tile (doubles)

template <int NT8>   // kb / 8 (kb padded to a multiple of 8)
__global__ void __launch_bounds__(256)
symm_mul_kernel(const double* __restrict__ C, const double* __restrict__ Q, double* __restrict__ Z, int p, int kb) {
    __shared__ double Cs[SM_ROWS][SM_CP];
    __shared__ double Qs[SM_KC][NT8 * 8 + 1];
    const int f = blockIdx.y, i0 = blockIdx.x * SM_ROWS;
    const double* Cf = C + (size_t)f * p * p;
    const double* Qf = Q + (size_t)f * p * kb;
    double* Zf = Z + (size_t)f * p * kb;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tq = lane & 3;
    double acc[NT8][2];
#pragma unroll
    for (int n = 0; n < NT8; ++n) acc[n][0] = acc[n][1] = 0.0;
    // The next chunk's global loads are issued into registers BEFORE the DMMAs of the current one (the first version
    // loaded, synchronised, multiplied, synchronised: every chunk paid the full L2 latency; the solve of a CV grid spends a
    // quarter of its time in this kernel).  C tile: 64 x 32, 8 doubles per thread (row = tid / 4, 8 consecutive columns);
    // Q tile: 32 x (8 NT8), NT8 doubles per thread.
    const int cr = tid >> 2, cc8 = (tid & 3) * 8, ci = i0 + cr;
    constexpr int QPT = SM_KC * NT8 * 8 / 256;
    double creg[8], qreg[QPT];
    auto fetch = [&](int j0) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int j = j0 + cc8 + e;
            creg[e] = (ci < p && j < p) ? Cf[(size_t)ci * p + j] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < QPT; ++u) {
            const int idx = tid + 256 * u, r = idx / (NT8 * 8), c = idx % (NT8 * 8);
            const int j = j0 + r;
            qreg[u] = (j < p && c < kb) ? Qf[(size_t)j * kb + c] : 0.0;
        }
    };
    fetch(0);
    for (int j0 = 0; j0 < p; j0 += SM_KC) {
#pragma unroll
        for (int e = 0; e < 8; ++e) Cs[cr][cc8 + e] = creg[e];
#pragma unroll
        for (int u = 0; u < QPT; ++u) { const int idx = tid + 256 * u; Qs[idx / (NT8 * 8)][idx % (NT8 * 8)] = qreg[u]; }
        __syncthreads();
        if (j0 + SM_KC < p) fetch(j0 + SM_KC);
#pragma unroll
        for (int k4 = 0; k4 < SM_KC / 4; ++k4) {
            const double a = Cs[8 * warp + g][4 * k4 + tq];
#pragma unroll
            for (int n = 0; n < NT8; ++n) dmma884(acc[n][0], acc[n][1], a, Qs[4 * k4 + tq][8 * n + g]);
        }
        __syncthreads();
    }
    const int i = i0 + 8 * warp + g;
    if (i < p) {
#pragma unroll
        for (int n = 0; n < NT8; ++n)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = 8 * n + 2 * tq + e;
                if (c < kb) Zf[(size_t)i * kb + c] = acc[n][e];
            }
    }
}

// ---------------------------------------------------------------- H = A^T B ----------------------------
// grid (chunks of rows, F); each CTA reduces its rows into a ka x kb tile in registers (thread -> up to 16 entries),
// then fp64 atomics into H (zeroed by the caller wrapper).
__global__ void __launch_bounds__(256)
tn_small_kernel(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ H, int p, int ka, int kb,
                int rows_per_cta) {
    __shared__ double As[32][KBMAX + 1];
    __shared__ double Bs[32][KBMAX + 1];
    const int f = blockIdx.y;
    const double* Af = A + (size_t)f * p * ka;
    const double* Bf = B + (size_t)f * p * kb;
    double* Hf = H + (size_t)f * ka * kb;
    const int r_begin = blockIdx.x * rows_per_cta, r_end = min(p, r_begin + rows_per_cta);
    const int tid = threadIdx.x;
    // thread -> entries (a, b) with a = ty + 16 u, b = tx + 16 v, u, v in 0..3
    const int tx = tid & 15, ty = tid >> 4;
    double acc[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = 0.0;
    for (int r0 = r_begin; r0 < r_end; r0 += 32) {
        for (int idx = tid; idx < 32 * KBMAX; idx += 256) {
            const int r = idx / KBMAX, c = idx % KBMAX;
            const int row = r0 + r;
            As[r][c] = (row < r_end && c < ka) ? Af[(size_t)row * ka + c] : 0.0;
            Bs[r][c] = (row < r_end && c < kb) ? Bf[(size_t)row * kb + c] : 0.0;
        }
        __syncthreads();
#pragma unroll 4
        for (int r = 0; r < 32; ++r) {
            double a[4], b[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { a[u] = As[r][ty + 16 * u]; b[u] = Bs[r][tx + 16 * u]; }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < 4; ++v) acc[u][v] = fma(a[u], b[v], acc[u][v]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const int a = ty + 16 * u, b = tx + 16 * v;
            if (a < ka && b < kb && acc[u][v] != 0.0) atomicAdd(Hf + a * kb + b, acc[u][v]);
        }
}

// ---------------------------------------------------------------- Cholesky of kb x kb (one CTA per matrix) -----
// Lower factor L (row-major, zeros above the diagonal) of the symmetrised G; info[f] = 0 or the 1-based index of the
// first non-positive pivot (then L is not usable and the caller falls back).
__global__ void __launch_bounds__(64)
chol_small_kernel(const double* __restrict__ G, int kb, double* __restrict__ L, int* __restrict__ info) {
    __shared__ double Ls[KBMAX][KBMAX + 1];
    __shared__ int bad;
    const int f = blockIdx.x, t = threadIdx.x;
    const double* Gf = G + (size_t)f * kb * kb;
    for (int idx = t; idx < kb * kb; idx += 64) {
        const int r = idx / kb, c = idx % kb;
        Ls[r][c] = 0.5 * (Gf[r * kb + c] + Gf[c * kb + r]);
    }
    if (t == 0) bad = 0;
    __syncthreads();
    for (int j = 0; j < kb; ++j) {
        // column j: L[j][j] = sqrt(G[j][j] - sum_k L[j][k]^2), L[i][j] = (G[i][j] - sum_k L[i][k] L[j][k]) / L[j][j]
        if (t == 0) {
            const double gjj = Ls[j][j];
            double s = gjj;
            for (int k = 0; k < j; ++k) s -= Ls[j][k] * Ls[j][k];
            // a pivot that is rounding noise of the diagonal entry means a (numerically) rank-deficient block
            const bool ok = s > 1e-15 * gjj && gjj > 0.0;   // (a block of condition 1e6 still has pivots ~1e-12 gjj)
            if (!ok && bad == 0) bad = j + 1;
            Ls[j][j] = ok ? sqrt(s) : 1.0;
        }
        __syncthreads();
        const int i = j + 1 + t;
        if (i < kb) {
            double s = Ls[i][j];
            for (int k = 0; k < j; ++k) s -= Ls[i][k] * Ls[j][k];
            Ls[i][j] = s / Ls[j][j];
        }
        __syncthreads();
    }
    double* Lf = L + (size_t)f * kb * kb;
    for (int idx = t; idx < kb * kb; idx += 64) {
        const int r = idx / kb, c = idx % kb;
        Lf[idx] = c <= r ? Ls[r][c] : 0.0;
    }
    if (t == 0) info[f] = bad;
}

// Y <- Y L^-T: every row y solves x L^T = y, i.e. forward substitution x_c = (y_c - sum_{k<c} x_k L[c][k]) / L[c][c]
__global__ void __launch_bounds__(128)
trsm_rows_kernel(double* __restrict__ Y, const double* __restrict__ L, int p, int kb) {
    __shared__ double Ls[KBMAX][KBMAX + 1];
    const int f = blockIdx.y;
    const double* Lf = L + (size_t)f * kb * kb;
    for (int idx = threadIdx.x; idx < kb * kb; idx += blockDim.x) Ls[idx / kb][idx % kb] = Lf[idx];
    __syncthreads();
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= p) return;
    double* y = Y + ((size_t)f * p + row) * kb;
    double x[KBMAX];
#pragma unroll 1
    for (int c = 0; c < kb; ++c) {
        double s = y[c];
        for (int k = 0; k < c; ++k) s -= x[k] * Ls[c][k];
        x[c] = s / Ls[c][c];
    }
    for (int c = 0; c < kb; ++c) y[c] = x[c];
}

// ---------------------------------------------------------------- Y = Z U (row-wise small product) ------------
__global__ void __launch_bounds__(128)
mul_small_kernel(const double* __restrict__ Zm, const double* __restrict__ U, double* __restrict__ Y, int p, int kb, int ku) {
    __shared__ double Us[KBMAX][KBMAX + 1];
    const int f = blockIdx.y;
    const double* Uf = U + (size_t)f * kb * ku;
    for (int idx = threadIdx.x; idx < kb * ku; idx += blockDim.x) Us[idx / ku][idx % ku] = Uf[idx];
    __syncthreads();
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= p) return;
    const double* z = Zm + ((size_t)f * p + row) * kb;
    double* y = Y + ((size_t)f * p + row) * ku;
    double zr[KBMAX];
    for (int k = 0; k < kb; ++k) zr[k] = z[k];
    for (int c = 0; c < ku; ++c) {
        double s = 0.0;
        for (int k = 0; k < kb; ++k) s = fma(zr[k], Us[k][c], s);
        y[c] = s;
    }
}

// ---------------------------------------------------------------- residual norms ------------------------------
// res2[f][c] += sum_rows (ZU[row][c] - X[row][c] w[c])^2   (the caller takes the square root)
__global__ void __launch_bounds__(256)
ritz_residual_kernel(const double* __restrict__ ZU, const double* __restrict__ X, const double* __restrict__ w,
                     double* __restrict__ res2, int p, int kb, int rows_per_cta) {
    const int f = blockIdx.y;
    const int c = threadIdx.x % KBMAX, lane_row = threadIdx.x / KBMAX;      // 4 row lanes x 64 columns
    const int r_begin = blockIdx.x * rows_per_cta, r_end = min(p, r_begin + rows_per_cta);
    double s = 0.0;
    if (c < kb) {
        const double wc = w[(size_t)f * kb + c];
        for (int r = r_begin + lane_row; r < r_end; r += 4) {
            const size_t o = ((size_t)f * p + r) * kb + c;
            const double d = ZU[o] - X[o] * wc;
            s = fma(d, d, s);
        }
    }
    __shared__ double red[4][KBMAX];
    red[lane_row][c] = s;
    __syncthreads();
    if (lane_row == 0 && c < kb) {
        const double t = red[0][c] + red[1][c] + red[2][c] + red[3][c];
        if (t != 0.0) atomicAdd(res2 + (size_t)f * kb + c, t);
    }
}

// ---------------------------------------------------------------- deflation + tail sums -----------------------
// R = C - sum_{a<k} w_a x_a x_a^T   (X [F][p][kb], the first k columns are used)
__global__ void __launch_bounds__(256)
deflate_kernel(const double* __restrict__ C, const double* __restrict__ X, const double* __restrict__ w, double* __restrict__ R,
               int p, int kb, int k) {
    __shared__ double Xi[16][KBMAX + 1], Xj[16][KBMAX + 1], ws[KBMAX];
    const int f = blockIdx.z, i0 = blockIdx.y * 16, j0 = blockIdx.x * 16;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const double* Xf = X + (size_t)f * p * kb;
    for (int idx = threadIdx.x; idx < 16 * k; idx += 256) {
        const int r = idx / k, a = idx % k;
        Xi[r][a] = (i0 + r < p) ? Xf[(size_t)(i0 + r) * kb + a] : 0.0;
        Xj[r][a] = (j0 + r < p) ? Xf[(size_t)(j0 + r) * kb + a] : 0.0;
    }
    if (threadIdx.x < k) ws[threadIdx.x] = w[(size_t)f * kb + threadIdx.x];
    __syncthreads();
    const int i = i0 + ty, j = j0 + tx;
    if (i >= p || j >= p) return;
    double s = 0.0;
    for (int a = 0; a < k; ++a) s = fma(ws[a] * Xi[ty][a], Xj[tx][a], s);
    const size_t o = (size_t)f * p * p + (size_t)i * p + j;
    R[o] = C[o] - s;
}

// tails[f] = {tr R, |R|_F^2, sum_ij (R^2)_ij R_ij}; R2 holds the UPPER triangle of R^2 (Gram kernel output)
__global__ void __launch_bounds__(256)
tail_sums_kernel(const double* __restrict__ R, const double* __restrict__ R2, double* __restrict__ tails, int p) {
    const int f = blockIdx.y;
    const double* Rf = R + (size_t)f * p * p;
    const double* R2f = R2 + (size_t)f * p * p;
    double t1 = 0.0, t2 = 0.0, t3 = 0.0;
    const int64_t total = (int64_t)p * p;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx / p), j = (int)(idx % p);
        const double r = Rf[idx];
        t2 = fma(r, r, t2);
        if (i == j) { t1 += r; t3 = fma(R2f[idx], r, t3); }
        else if (j > i) t3 = fma(2.0 * R2f[idx], r, t3);
    }
    t1 = warp_sum(t1); t2 = warp_sum(t2); t3 = warp_sum(t3);
    __shared__ double red[8][3];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { red[w][0] = t1; red[w][1] = t2; red[w][2] = t3; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double s = 0.0;
        for (int k = 0; k < 8; ++k) s += red[k][threadIdx.x];
        atomicAdd(tails + (size_t)f * 3 + threadIdx.x, s);
    }
}

}  // namespace esub
}  // namespace ocm

using namespace ocm;

extern "C" int ocm_eig_symm_mul_f64(const double* C, const double* Q, double* Z, int32_t p, int32_t kb, int32_t batch, void* stream) {
    if (batch == 0) return 0;
    OCM_REQUIRE(C && Q && Z && p > 0 && batch > 0, OCM_EINVAL, "NULL buffer or bad shape");
    OCM_REQUIRE(kb >= 1 && kb <= esub::KBMAX, OCM_ELIMIT, "block width must be 1..64");
    dim3 grid((p + esub::SM_ROWS - 1) / esub::SM_ROWS, batch);
    const int nt8 = (kb + 7) / 8;
    cudaStream_t st = (cudaStream_t)stream;
    switch (nt8) {
        case 1: esub::symm_mul_kernel<1><<<grid, 256, 0, st>>>(C, Q, Z, p, kb); break;
        case 2: esub::symm_mul_kernel<2><<<grid, 256, 0, st>>>(C, Q, Z, p, kb); break;
        case 3: esub::symm_mul_kernel<3><<<grid, 256, 0, st>>>(C, Q, Z, p, kb); break;
        case 4: esub::symm_mul_kernel<4><<<grid, 256, 0, st>>>(C, Q, Z, p, kb); break;
        case 5: esub::symm_mul_kernel<5><<<grid, 256, 0, st>>>(C, Q, Z, p, kb); break;
        case 6: esub::symm_mul_kernel<6><<<grid, 256, 0, st>>>(C, Q, Z, p, kb); break;
        case 7: esub::symm_mul_kernel<7><<<grid, 256, 0, st>>>(C, Q, Z, p, kb); break;
        default: esub::symm_mul_kernel<8><<<grid, 256, 0, st>>>(C, Q, Z, p, kb); break;
    }
    return check_launch("symm_mul_kernel");
}

extern "C" int ocm_eig_tn_small_f64(const double* A, const double* B, double* H, int32_t p, int32_t ka, int32_t kb,
                                    int32_t batch, void* stream) {
    if (batch == 0) return 0;
    OCM_REQUIRE(A && B && H && p > 0 && batch > 0, OCM_EINVAL, "NULL buffer or bad shape");
    OCM_REQUIRE(ka >= 1 && ka <= esub::KBMAX && kb >= 1 && kb <= esub::KBMAX, OCM_ELIMIT, "block widths must be 1..64");
    cudaStream_t st = (cudaStream_t)stream;
    OCM_CUDA(cudaMemsetAsync(H, 0, sizeof(double) * (size_t)batch * ka * kb, st));
    int chunks = (sm_count() * 2 + batch - 1) / batch;
    if (chunks > (p + 63) / 64) chunks = (p + 63) / 64;
    if (chunks < 1) chunks = 1;
    const int rpc = ((p + chunks - 1) / chunks + 31) / 32 * 32;
    dim3 grid((p + rpc - 1) / rpc, batch);
    esub::tn_small_kernel<<<grid, 256, 0, st>>>(A, B, H, p, ka, kb, rpc);
    return check_launch("tn_small_kernel");
}

extern "C" int ocm_eig_chol_solve_f64(const double* G, double* Y, double* L, int32_t* info, int32_t p, int32_t kb,
                                      int32_t batch, void* stream) {
    if (batch == 0) return 0;
    OCM_REQUIRE(G && Y && L && info && p > 0 && batch > 0, OCM_EINVAL, "NULL buffer or bad shape");
    OCM_REQUIRE(kb >= 1 && kb <= esub::KBMAX, OCM_ELIMIT, "block width must be 1..64");
    cudaStream_t st = (cudaStream_t)stream;
    esub::chol_small_kernel<<<batch, 64, 0, st>>>(G, kb, L, info);
    int rc = check_launch("chol_small_kernel");
    if (rc) return rc;
    dim3 grid((p + 127) / 128, batch);
    esub::trsm_rows_kernel<<<grid, 128, 0, st>>>(Y, L, p, kb);
    return check_launch("trsm_rows_kernel");
}

extern "C" int ocm_eig_mul_small_f64(const double* Zm, const double* U, double* Y, int32_t p, int32_t kb, int32_t ku,
                                     int32_t batch, void* stream) {
    if (batch == 0) return 0;
    OCM_REQUIRE(Zm && U && Y && p > 0 && batch > 0, OCM_EINVAL, "NULL buffer or bad shape");
    OCM_REQUIRE(kb >= 1 && kb <= esub::KBMAX && ku >= 1 && ku <= esub::KBMAX, OCM_ELIMIT, "block widths must be 1..64");
    dim3 grid((p + 127) / 128, batch);
    esub::mul_small_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(Zm, U, Y, p, kb, ku);
    return check_launch("mul_small_kernel");
}

extern "C" int ocm_eig_ritz_residual_f64(const double* ZU, const double* X, const double* w, double* res2, int32_t p,
                                         int32_t kb, int32_t batch, void* stream) {
    if (batch == 0) return 0;
    OCM_REQUIRE(ZU && X && w && res2 && p > 0 && batch > 0, OCM_EINVAL, "NULL buffer or bad shape");
    OCM_REQUIRE(kb >= 1 && kb <= esub::KBMAX, OCM_ELIMIT, "block width must be 1..64");
    cudaStream_t st = (cudaStream_t)stream;
    OCM_CUDA(cudaMemsetAsync(res2, 0, sizeof(double) * (size_t)batch * kb, st));
    const int rpc = 256;
    dim3 grid((p + rpc - 1) / rpc, batch);
    esub::ritz_residual_kernel<<<grid, 256, 0, st>>>(ZU, X, w, res2, p, kb, rpc);
    return check_launch("ritz_residual_kernel");
}

extern "C" int ocm_gram_f64acc_f64(const double* X, int64_t ldx, const int64_t* row_idx, int64_t n, int32_t p,
                                   const double* shift, double* G, void* stream);

extern "C" int ocm_eig_tail_sums_f64(const double* C, const double* X, const double* w, int32_t p, int32_t kb, int32_t k,
                                     int32_t batch, double* R, double* R2, double* tails, void* stream) {
    if (batch == 0) return 0;
    OCM_REQUIRE(C && X && w && R && R2 && tails && p > 0 && batch > 0, OCM_EINVAL, "NULL buffer or bad shape");
    OCM_REQUIRE(kb >= 1 && kb <= esub::KBMAX && k >= 0 && k <= kb, OCM_ELIMIT, "block width must be 1..64, k <= kb");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((p + 15) / 16, (p + 15) / 16, batch);
    esub::deflate_kernel<<<grid, 256, 0, st>>>(C, X, w, R, p, kb, k);
    int rc = check_launch("deflate_kernel");
    if (rc) return rc;
    OCM_CUDA(cudaMemsetAsync(R2, 0, sizeof(double) * (size_t)batch * p * p, st));
    OCM_CUDA(cudaMemsetAsync(tails, 0, sizeof(double) * (size_t)batch * 3, st));
    for (int f = 0; f < batch; ++f) {      // R^T R = R^2 (R symmetric): the FP64 tensor-pipe Gram kernel, upper triangle
        rc = ocm_gram_f64acc_f64(R + (size_t)f * p * p, p, nullptr, p, p, nullptr, R2 + (size_t)f * p * p, stream);
        if (rc) return rc;
    }
    dim3 g2(sm_count() * 2 / batch > 0 ? sm_count() * 2 / batch : 1, batch);
    esub::tail_sums_kernel<<<g2, 256, 0, st>>>(R, R2, tails, p);
    return check_launch("tail_sums_kernel");
}
