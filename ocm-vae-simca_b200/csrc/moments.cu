// PCA-fit moments: column sums / extrema and the X^T X Gram of selected rows,
// accumulated in fp64 (FP64 CUDA cores for tiny problems, the FP64 tensor pipe -- DMMA -- otherwise),
// plus the combination of per-partition moments into a covariance.  This replaces the LAPACK SVD that sklearn's PCA runs inside
// utils/SIMCA.py:64-66 / :75 -- the eigen-decomposition of the covariance gives the same
// loadings and explained variances (S^2/(n-1)) without ever forming U.
//
// This is the path for fits that live in L2 (CV folds, config 1/2) and the checker of the
// bulk config-4 path, the integer tcgen05 Gram in gram_tc.cu.  Rows are selected through an optional int64 gather index so CV
// folds never materialise copies of X.
#include "ocm_common.cuh"

#include <stdlib.h>

namespace ocm {

// ---------------------------------------------------------------- column sums -----------
// grid: (ceil(p/256), row_blocks); each thread owns one column, walks a strided set of rows.
template <typename T>
__global__ void __launch_bounds__(256)
colstats_kernel(const T* __restrict__ X, int64_t ldx, const int64_t* __restrict__ row_idx,
                int64_t n, int p, const double* __restrict__ shift,
                double* __restrict__ colsum, double* __restrict__ colsumsq) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= p) return;
    const double sh = shift ? shift[j] : 0.0;
    double s = 0.0, ss = 0.0;
    for (int64_t r = blockIdx.y; r < n; r += gridDim.y) {
        const int64_t row = row_idx ? row_idx[r] : r;
        const double v = (double)X[row * ldx + j] - sh;
        s += v;
        ss = fma(v, v, ss);
    }
    atomicAdd(colsum + j, s);
    if (colsumsq) atomicAdd(colsumsq + j, ss);
}

// ---------------------------------------------------------------- column sums + extrema --
// One pass for the bulk fit: sum (fp64), min and max (fp32 bit patterns, see atomic_min/max_float) of every column.
// Thread -> 4 consecutive columns (128-bit loads when aligned), rows strided over gridDim.y, 4 rows in flight.
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
    if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_min_float(float* addr, float v) {
    if (v >= 0.f) atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ float to_float_up(float v) { return v; }
__device__ __forceinline__ float to_float_dn(float v) { return v; }
__device__ __forceinline__ float to_float_up(double v) { return __double2float_ru(v); }
__device__ __forceinline__ float to_float_dn(double v) { return __double2float_rd(v); }

template <typename T>
__global__ void __launch_bounds__(256)
colstats_minmax_kernel(const T* __restrict__ X, int64_t ldx, const int64_t* __restrict__ row_idx, int64_t n, int p,
                       double* __restrict__ colsum, float* __restrict__ colmin, float* __restrict__ colmax, int vec_ok) {
    const int j0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (j0 >= p) return;
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    T mn[4], mx[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) { mn[e] = (T)INFINITY; mx[e] = (T)-INFINITY; }
    const bool full = j0 + 3 < p;
    auto load = [&](int64_t r, T* v) {
        const int64_t row = row_idx ? row_idx[r] : r;
        const T* xr = X + row * ldx + j0;
        if constexpr (sizeof(T) == 4) {
            if (vec_ok && full) {
                const float4 t = ldg_stream4(reinterpret_cast<const float*>(xr));
                v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
                return;
            }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = (j0 + e < p) ? xr[e] : T(0);
    };
    const int64_t step = gridDim.y;
    int64_t r = blockIdx.y;
    for (; r + 3 * step < n; r += 4 * step) {
        T v[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u) load(r + u * step, v[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                s[e] += (double)v[u][e];
                mn[e] = v[u][e] < mn[e] ? v[u][e] : mn[e];
                mx[e] = v[u][e] > mx[e] ? v[u][e] : mx[e];
            }
    }
    for (; r < n; r += step) {
        T v[4];
        load(r, v);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            s[e] += (double)v[e];
            mn[e] = v[e] < mn[e] ? v[e] : mn[e];
            mx[e] = v[e] > mx[e] ? v[e] : mx[e];
        }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        if (j0 + e >= p) break;
        atomicAdd(colsum + j0 + e, s[e]);
        if (colmin) atomic_min_float(colmin + j0 + e, to_float_dn(mn[e]));
        if (colmax) atomic_max_float(colmax + j0 + e, to_float_up(mx[e]));
    }
}

template <typename T>
int colstats_minmax_impl(const T* X, int64_t ldx, const int64_t* row_idx, int64_t n, int32_t p, double* colsum,
                         float* colmin, float* colmax, void* stream) {
    if (n == 0) return 0;
    OCM_REQUIRE(X && colsum && n > 0 && p > 0 && ldx >= p, OCM_EINVAL, "bad shape or NULL buffer");
    const int threads = 256, cols_per_block = threads * 4;
    const int gx = (p + cols_per_block - 1) / cols_per_block;
    int64_t gy = (int64_t)sm_count() * 8 / gx;
    if (gy > (n + 7) / 8) gy = (n + 7) / 8;
    if (gy < 1) gy = 1;
    if (gy > 65535) gy = 65535;
    const int vec_ok = (sizeof(T) == 4 && ldx % 4 == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0) ? 1 : 0;
    dim3 grid(gx, (unsigned)gy);
    colstats_minmax_kernel<T><<<grid, threads, 0, (cudaStream_t)stream>>>(X, ldx, row_idx, n, p, colsum, colmin, colmax, vec_ok);
    return check_launch("colstats_minmax_kernel");
}

// ---------------------------------------------------------------- fp64 Gram -------------
// 64x64 output tile per CTA (upper-triangle tiles only), 4x4 doubles per thread, K (rows)
// split over gridDim.z with one fp64 atomicAdd per output element per split.
constexpr int GT = 64;   // tile edge
constexpr int GK = 16;   // rows per shared-memory stage

template <typename T>
__global__ void __launch_bounds__(256)
gram_f64_kernel(const T* __restrict__ X, int64_t ldx, const int64_t* __restrict__ row_idx,
                int64_t n, int p, const double* __restrict__ shift, double* __restrict__ G,
                int64_t rows_per_split) {
    // map blockIdx.x to an upper-triangle tile (ti <= tj)
    const int nt = (p + GT - 1) / GT;
    int t = blockIdx.x, ti = 0;
    while (t >= nt - ti) { t -= nt - ti; ++ti; }
    const int tj = ti + t;
    const int i0 = ti * GT, j0 = tj * GT;

    __shared__ double As[GK][GT + 1];
    __shared__ double Bs[GK][GT + 1];

    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;

    const int64_t r_begin = (int64_t)blockIdx.z * rows_per_split;
    const int64_t r_end = min(n, r_begin + rows_per_split);

    // loader mapping: 256 threads cover 16 rows x 64 cols in 4 steps of 4 rows
    const int lc = threadIdx.x & 63, lr = threadIdx.x >> 6;
    const double shA = (shift && i0 + lc < p) ? shift[i0 + lc] : 0.0;
    const double shB = (shift && j0 + lc < p) ? shift[j0 + lc] : 0.0;

    for (int64_t r0 = r_begin; r0 < r_end; r0 += GK) {
#pragma unroll
        for (int s = 0; s < GK / 4; ++s) {
            const int kk = lr + 4 * s;
            const int64_t r = r0 + kk;
            double a = 0.0, b = 0.0;
            if (r < r_end) {
                const int64_t row = row_idx ? row_idx[r] : r;
                const T* xr = X + row * ldx;
                if (i0 + lc < p) a = (double)xr[i0 + lc] - shA;
                if (j0 + lc < p) b = (double)xr[j0 + lc] - shB;
            }
            As[kk][lc] = a;
            Bs[kk][lc] = b;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < GK; ++kk) {
            double a[4], b[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { a[u] = As[kk][ty * 4 + u]; b[u] = Bs[kk][tx * 4 + u]; }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < 4; ++v) acc[u][v] = fma(a[u], b[v], acc[u][v]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int i = i0 + ty * 4 + u;
        if (i >= p) continue;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const int j = j0 + tx * 4 + v;
            if (j < p && j >= i) atomicAdd(G + (int64_t)i * p + j, acc[u][v]);
        }
    }
}

// ---------------------------------------------------------------- fp64 tensor-core Gram ---
// The same contraction on the FP64 tensor pipe (DMMA, mma.sync m8n8k4 f64): 128x128 output tile per CTA
// (upper-triangle tiles only), 8 warps of 64x32, 32- (fp32) or 16-row (fp64) K stages double-buffered in shared memory, the next
// stage prefetched into registers while the current one is multiplied.  Why fp64 and not the tcgen05 low-precision
// kinds: the 5th-gen tensor core adds into its fp32 accumulator with truncation, which biases a chain of k
// accumulations by ~k * 2^-25 relative (measured in score_tc.cu: 384 chained MMAs -> -1.1e-5).  On strongly
// correlated spectra that bias is structured (sign-consistent products) and of the size of the small
// eigenvalues a SIMCA model needs (lambda_20 / lambda_1 ~ 2e-4), and an on-chip fp64 tile (128x256 doubles =
// 256 KB) does not exist to flush short chains into.  DMMA keeps every product and sum in fp64.
constexpr int DT = 128;        // tile edge
template <typename T> struct GramStage { static constexpr int DK = sizeof(T) == 4 ? 32 : 16; };   // rows per stage (fp32 rows are half the prefetch registers)
constexpr int DPITCH = DT + 4; // doubles; +32 B per row keeps the fragment LDS.64 conflict free

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <typename T>
__global__ void __launch_bounds__(256, 1)
gram_dmma_kernel(const T* __restrict__ X, int64_t ldx, const int64_t* __restrict__ row_idx,
                 int64_t n, int p, const double* __restrict__ shift, double* __restrict__ G,
                 int64_t rows_per_split) {
    constexpr int DK = GramStage<T>::DK, NH = DK / 8;
    extern __shared__ __align__(16) double sm_d[];
    double (*As)[DK][DPITCH] = reinterpret_cast<double (*)[DK][DPITCH]>(sm_d);
    double (*Bs)[DK][DPITCH] = reinterpret_cast<double (*)[DK][DPITCH]>(sm_d + 2 * DK * DPITCH);

    const int nt = (p + DT - 1) / DT;
    int t = blockIdx.x, ti = 0;
    while (t >= nt - ti) { t -= nt - ti; ++ti; }
    const int tj = ti + t;
    const int i0 = ti * DT, j0 = tj * DT;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 2, wn = warp & 3, g = lane >> 2, tq = lane & 3;

    double acc[8][4][2];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

    const int64_t r_begin = (int64_t)blockIdx.z * rows_per_split;
    const int64_t r_end = min(n, r_begin + rows_per_split);

    // loader: 2 x (4 consecutive columns) of the A block and of the B block per thread and stage
    const int lk0 = tid >> 5, lc = (tid & 31) * 4;           // rows lk0 and lk0 + 8, columns lc .. lc+3
    double shA[4], shB[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        shA[e] = (shift && i0 + lc + e < p) ? shift[i0 + lc + e] : 0.0;
        shB[e] = (shift && j0 + lc + e < p) ? shift[j0 + lc + e] : 0.0;
    }
    // The loads of the NEXT stage are issued before the MMAs of the current one and must stay raw until the stash: the
    // first version converted and shifted right at the fetch, which put the wait for the loads in front of the MMAs
    // (ncu: 27 % of all warp samples on that line).
    T ra[NH][4], rb[NH][4];
    bool rok[NH];
    bool cokA[4], cokB[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) { cokA[e] = i0 + lc + e < p; cokB[e] = j0 + lc + e < p; }
    auto fetch = [&](int64_t r0) {
#pragma unroll
        for (int h = 0; h < NH; ++h) {
            const int64_t r = r0 + lk0 + 8 * h;
            rok[h] = r < r_end;
            const int64_t row = rok[h] ? (row_idx ? row_idx[r] : r) : 0;
            const T* xr = X + row * ldx;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                ra[h][e] = (rok[h] && cokA[e]) ? xr[i0 + lc + e] : T(0);
                rb[h][e] = (rok[h] && cokB[e]) ? xr[j0 + lc + e] : T(0);
            }
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int h = 0; h < NH; ++h) {
            double va[4], vb[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                va[e] = (rok[h] && cokA[e]) ? (double)ra[h][e] - shA[e] : 0.0;
                vb[e] = (rok[h] && cokB[e]) ? (double)rb[h][e] - shB[e] : 0.0;
            }
            double2* da = reinterpret_cast<double2*>(&As[buf][lk0 + 8 * h][lc]);
            double2* db = reinterpret_cast<double2*>(&Bs[buf][lk0 + 8 * h][lc]);
            da[0] = make_double2(va[0], va[1]); da[1] = make_double2(va[2], va[3]);
            db[0] = make_double2(vb[0], vb[1]); db[1] = make_double2(vb[2], vb[3]);
        }
    };

    fetch(r_begin);
    stash(0);
    __syncthreads();
    int buf = 0;
    for (int64_t r0 = r_begin; r0 < r_end; r0 += DK) {
        const bool more = r0 + DK < r_end;
        if (more) fetch(r0 + DK);                             // global loads in flight under the MMAs
#pragma unroll
        for (int k4 = 0; k4 < DK / 4; ++k4) {
            double a[8], b[4];
#pragma unroll
            for (int mt = 0; mt < 8; ++mt) a[mt] = As[buf][4 * k4 + tq][64 * wm + 8 * mt + g];
#pragma unroll
            for (int ntl = 0; ntl < 4; ++ntl) b[ntl] = Bs[buf][4 * k4 + tq][32 * wn + 8 * ntl + g];
#pragma unroll
            for (int mt = 0; mt < 8; ++mt)
#pragma unroll
                for (int ntl = 0; ntl < 4; ++ntl) dmma884(acc[mt][ntl][0], acc[mt][ntl][1], a[mt], b[ntl]);
        }
        if (more) stash(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }
#pragma unroll
    for (int mt = 0; mt < 8; ++mt) {
        const int i = i0 + 64 * wm + 8 * mt + g;
        if (i >= p) continue;
#pragma unroll
        for (int ntl = 0; ntl < 4; ++ntl)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int j = j0 + 32 * wn + 8 * ntl + 2 * tq + e;
                if (j < p && j >= i) atomicAdd(G + (int64_t)i * p + j, acc[mt][ntl][e]);
            }
    }
}

// ---------------------------------------------------------------- combine ---------------
struct PartsArgs {
    int ngroups;
    double weights[64];
    double N;
};

__global__ void __launch_bounds__(256)
mean_from_parts_kernel(const double* __restrict__ sparts, PartsArgs a, int p,
                       const double* __restrict__ shift, double* __restrict__ delta,
                       double* __restrict__ mean) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= p) return;
    double s = 0.0;
    for (int g = 0; g < a.ngroups; ++g)
        if (a.weights[g] != 0.0) s += a.weights[g] * sparts[(int64_t)g * p + j];
    const double d = s / a.N;
    delta[j] = d;
    if (mean) mean[j] = (shift ? shift[j] : 0.0) + d;
}

__global__ void __launch_bounds__(256)
cov_from_parts_kernel(const double* __restrict__ Gparts, PartsArgs a, int p,
                      const double* __restrict__ delta, double* __restrict__ C) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= p || j < i) return;
    const int64_t pp = (int64_t)p * p;
    double g = 0.0;
    for (int k = 0; k < a.ngroups; ++k)
        if (a.weights[k] != 0.0) g += a.weights[k] * Gparts[k * pp + (int64_t)i * p + j];
    const double c = (g - a.N * delta[i] * delta[j]) / (a.N - 1.0);
    C[(int64_t)i * p + j] = c;
    C[(int64_t)j * p + i] = c;
}

template <typename T>
int colstats_impl(const T* X, int64_t ldx, const int64_t* row_idx, int64_t n, int32_t p,
                  const double* shift, double* colsum, double* colsumsq, void* stream) {
    if (n == 0) return 0;
    OCM_REQUIRE(X && colsum && n > 0 && p > 0 && ldx >= p, OCM_EINVAL, "bad shape or NULL buffer");
    int64_t rb = (n + 63) / 64;
    int64_t cap = (int64_t)sm_count() * 8 / ((p + 255) / 256) + 1;
    if (rb > cap) rb = cap;
    if (rb > 65535) rb = 65535;
    dim3 grid((p + 255) / 256, (unsigned)rb);
    colstats_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(X, ldx, row_idx, n, p, shift, colsum, colsumsq);
    return check_launch("colstats_kernel");
}

template <typename T>
int gram_impl(const T* X, int64_t ldx, const int64_t* row_idx, int64_t n, int32_t p,
              const double* shift, double* G, void* stream) {
    if (n == 0) return 0;
    OCM_REQUIRE(X && G && n > 0 && p > 0 && ldx >= p, OCM_EINVAL, "bad shape or NULL buffer");
    static const bool simple = [] { const char* e = getenv("OCM_GRAM_SIMPLE"); return e && e[0] == '1'; }();
    if (!simple && p >= 96 && n >= 64) {
        // FP64 tensor-pipe kernel: 128x128 tiles, one CTA per SM, K split to fill the machine
        const int ntd = (p + DT - 1) / DT;
        const int tiles = ntd * (ntd + 1) / 2;
        // K splits: the CTAs take equal time, so pick the split count whose tiles * splits fills whole waves of SMs
        // best (p = 2048: 136 tiles -> 37 splits = exactly 34 waves of 148 instead of 136 CTAs on 148 SMs), with at
        // least 4096 rows per CTA so the 128 x 128 atomic epilogue stays negligible
        const int sms = sm_count();
        int64_t max_splits = n / 4096;
        if (max_splits < 1) max_splits = sms / tiles > 0 && n >= 512 ? sms / tiles : 1;
        if (max_splits > 64) max_splits = 64;
        const int64_t rows_cap = (n + 255) / 256;
        if (max_splits > rows_cap) max_splits = rows_cap;
        int64_t splits = 1;
        double best = 0.0;
        for (int64_t sp = 1; sp <= max_splits; ++sp) {
            const int64_t ctas = (int64_t)tiles * sp, waves = (ctas + sms - 1) / sms;
            const double eff = (double)ctas / (double)(waves * sms);
            if (eff > best + 0.01) { best = eff; splits = sp; }
        }
        int64_t rps = (n + splits - 1) / splits;
        constexpr int DK = GramStage<T>::DK;
        rps = (rps + DK - 1) / DK * DK;
        splits = (n + rps - 1) / rps;
        const size_t smem = sizeof(double) * 4 * DK * DPITCH;
        auto kern = gram_dmma_kernel<T>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
        dim3 grid(tiles, 1, (unsigned)splits);
        kern<<<grid, 256, smem, (cudaStream_t)stream>>>(X, ldx, row_idx, n, p, shift, G, rps);
        return check_launch("gram_dmma_kernel");
    }
    const int nt = (p + GT - 1) / GT;
    const int ntiles = nt * (nt + 1) / 2;
    // enough K-splits to fill the machine ~2 waves, at least 64 rows per split
    int64_t want = ((int64_t)sm_count() * 4 + ntiles - 1) / ntiles;
    int64_t max_splits = (n + 63) / 64;
    int64_t splits = want < 1 ? 1 : want;
    if (splits > max_splits) splits = max_splits;
    if (splits > 65535) splits = 65535;
    int64_t rows_per_split = (n + splits - 1) / splits;
    rows_per_split = (rows_per_split + GK - 1) / GK * GK;
    splits = (n + rows_per_split - 1) / rows_per_split;
    dim3 grid(ntiles, 1, (unsigned)splits);
    gram_f64_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(X, ldx, row_idx, n, p, shift, G, rows_per_split);
    return check_launch("gram_f64_kernel");
}

}  // namespace ocm

extern "C" int ocm_colstats_f32(const float* X, int64_t ldx, const int64_t* row_idx, int64_t n, int32_t p,
                                const double* shift, double* colsum, double* colsumsq, void* stream) {
    return ocm::colstats_impl<float>(X, ldx, row_idx, n, p, shift, colsum, colsumsq, stream);
}
extern "C" int ocm_colstats_f64(const double* X, int64_t ldx, const int64_t* row_idx, int64_t n, int32_t p,
                                const double* shift, double* colsum, double* colsumsq, void* stream) {
    return ocm::colstats_impl<double>(X, ldx, row_idx, n, p, shift, colsum, colsumsq, stream);
}
extern "C" int ocm_colstats_minmax_f32(const float* X, int64_t ldx, const int64_t* row_idx, int64_t n, int32_t p,
                                       double* colsum, float* colmin, float* colmax, void* stream) {
    return ocm::colstats_minmax_impl<float>(X, ldx, row_idx, n, p, colsum, colmin, colmax, stream);
}
extern "C" int ocm_colstats_minmax_f64(const double* X, int64_t ldx, const int64_t* row_idx, int64_t n, int32_t p,
                                       double* colsum, float* colmin, float* colmax, void* stream) {
    return ocm::colstats_minmax_impl<double>(X, ldx, row_idx, n, p, colsum, colmin, colmax, stream);
}
extern "C" int ocm_gram_f64acc_f32(const float* X, int64_t ldx, const int64_t* row_idx, int64_t n, int32_t p,
                                   const double* shift, double* G, void* stream) {
    return ocm::gram_impl<float>(X, ldx, row_idx, n, p, shift, G, stream);
}
extern "C" int ocm_gram_f64acc_f64(const double* X, int64_t ldx, const int64_t* row_idx, int64_t n, int32_t p,
                                   const double* shift, double* G, void* stream) {
    return ocm::gram_impl<double>(X, ldx, row_idx, n, p, shift, G, stream);
}

extern "C" int ocm_cov_from_parts_f64(const double* Gparts, const double* sparts, const int64_t* h_counts,
                                      const double* h_weights, int32_t ngroups, int32_t p,
                                      const double* shift, double* C, double* mean, double* delta,
                                      void* stream) {
    using namespace ocm;
    OCM_REQUIRE(Gparts && sparts && h_counts && C && delta && p > 0, OCM_EINVAL, "NULL buffer");
    OCM_REQUIRE(ngroups >= 1 && ngroups <= 64, OCM_ELIMIT, "1..64 partitions supported");
    PartsArgs a;
    a.ngroups = ngroups;
    a.N = 0.0;
    for (int g = 0; g < ngroups; ++g) {
        a.weights[g] = h_weights ? h_weights[g] : 1.0;
        a.N += a.weights[g] * (double)h_counts[g];
    }
    OCM_REQUIRE(a.N >= 2.0, OCM_EINVAL, "need at least two rows");
    mean_from_parts_kernel<<<(p + 255) / 256, 256, 0, (cudaStream_t)stream>>>(sparts, a, p, shift, delta, mean);
    int rc = check_launch("mean_from_parts_kernel");
    if (rc) return rc;
    dim3 grid((p + 255) / 256, p);
    cov_from_parts_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(Gparts, a, p, delta, C);
    return check_launch("cov_from_parts_kernel");
}
