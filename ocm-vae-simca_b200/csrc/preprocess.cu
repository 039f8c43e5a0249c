// Streaming preprocessing: row SNV, column centring, column autoscaling.
//
// Reference behaviour restated:
//   SNV        (x - mean_row) / (std_row + eps), population std   simca_nuts.py:48-50, utils/data_utils.py:57
//   centre     x - mu_col                                         inside sklearn PCA (SIMCA.py:64,:127)
//   autoscale  (x - mu_col) / sigma_col                           vae_model.py:124
//
// HBM-bound: 8*p bytes per spectrum (read + write).  One warp per row; the row stays in
// registers between the statistics pass and the write when p <= 32*4*kRegChunks, otherwise
// it is re-read (it is still in L1/L2).
#include "ocm_common.cuh"

namespace ocm {

constexpr int kRegChunks = 10;  // rows up to 1280 floats stay in registers (vector path)

template <bool VEC>
__global__ void __launch_bounds__(256)
preprocess_kernel(const float* __restrict__ X, int64_t n, int p, int64_t ldx, int mode,
                  const float* __restrict__ mu, const float* __restrict__ sigma, float eps,
                  float* __restrict__ Y, int64_t ldy) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const bool need_stats = (mode == OCM_PRE_SNV || mode == OCM_PRE_SNV_CENTER);
    const bool use_mu = (mode == OCM_PRE_CENTER || mode == OCM_PRE_AUTOSCALE || mode == OCM_PRE_SNV_CENTER);

    for (int64_t row = warp0; row < n; row += nwarps) {
        const float* x = X + row * ldx;
        float* y = Y + row * ldy;
        if constexpr (VEC) {
            const int nchunk = (p + 127) >> 7;  // float4 per lane per chunk
            float4 v[kRegChunks];
            const bool in_regs = nchunk <= kRegChunks;
            float m = 0.f, inv = 1.f;
            if (need_stats) {
                float s = 0.f;
                if (in_regs) {
#pragma unroll
                    for (int c = 0; c < kRegChunks; ++c) {
                        int j = (c << 7) + (lane << 2);
                        v[c] = (c < nchunk && j < p) ? ldg_stream4(x + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                        s += (v[c].x + v[c].y) + (v[c].z + v[c].w);
                    }
                } else {
                    for (int j = lane << 2; j < p; j += 128) {
                        float4 t = *reinterpret_cast<const float4*>(x + j);
                        s += (t.x + t.y) + (t.z + t.w);
                    }
                }
                m = warp_sum(s) / (float)p;
                float ss = 0.f;
                if (in_regs) {
#pragma unroll
                    for (int c = 0; c < kRegChunks; ++c) {
                        int j = (c << 7) + (lane << 2);
                        if (c < nchunk && j < p) {
                            float a = v[c].x - m, b = v[c].y - m, cc = v[c].z - m, d = v[c].w - m;
                            ss += (a * a + b * b) + (cc * cc + d * d);
                        }
                    }
                } else {
                    for (int j = lane << 2; j < p; j += 128) {
                        float4 t = *reinterpret_cast<const float4*>(x + j);
                        float a = t.x - m, b = t.y - m, cc = t.z - m, d = t.w - m;
                        ss += (a * a + b * b) + (cc * cc + d * d);
                    }
                }
                float sd = sqrtf(warp_sum(ss) / (float)p);
                inv = 1.f / (sd + eps);
            }
            auto emit = [&](int j, float4 t) {
                if (need_stats) {
                    t.x = (t.x - m) * inv; t.y = (t.y - m) * inv; t.z = (t.z - m) * inv; t.w = (t.w - m) * inv;
                }
                if (use_mu) {
                    float4 u = *reinterpret_cast<const float4*>(mu + j);
                    t.x -= u.x; t.y -= u.y; t.z -= u.z; t.w -= u.w;
                }
                if (mode == OCM_PRE_AUTOSCALE) {
                    float4 sg = *reinterpret_cast<const float4*>(sigma + j);
                    t.x /= sg.x; t.y /= sg.y; t.z /= sg.z; t.w /= sg.w;
                }
                *reinterpret_cast<float4*>(y + j) = t;
            };
            if (need_stats && in_regs) {
#pragma unroll
                for (int c = 0; c < kRegChunks; ++c) {
                    int j = (c << 7) + (lane << 2);
                    if (c < nchunk && j < p) emit(j, v[c]);
                }
            } else {
                for (int j = lane << 2; j < p; j += 128) emit(j, *reinterpret_cast<const float4*>(x + j));
            }
        } else {
            float m = 0.f, inv = 1.f;
            if (need_stats) {
                float s = 0.f;
                for (int j = lane; j < p; j += 32) s += x[j];
                m = warp_sum(s) / (float)p;
                float ss = 0.f;
                for (int j = lane; j < p; j += 32) { float a = x[j] - m; ss += a * a; }
                inv = 1.f / (sqrtf(warp_sum(ss) / (float)p) + eps);
            }
            for (int j = lane; j < p; j += 32) {
                float t = x[j];
                if (need_stats) t = (t - m) * inv;
                if (use_mu) t -= mu[j];
                if (mode == OCM_PRE_AUTOSCALE) t /= sigma[j];
                y[j] = t;
            }
        }
    }
}

}  // namespace ocm

extern "C" int ocm_preprocess_f32(const float* X, int64_t n, int32_t p, int64_t ldx, int32_t mode,
                                  const float* mu, const float* sigma, float eps,
                                  float* Y, int64_t ldy, void* stream) {
    using namespace ocm;
    if (n == 0) return 0;   // empty input: nothing to do (pointers of empty buffers may be NULL)
    OCM_REQUIRE(X && Y && n > 0 && p > 0 && ldx >= p && ldy >= p, OCM_EINVAL, "bad shape or NULL buffer");
    OCM_REQUIRE(mode >= OCM_PRE_SNV && mode <= OCM_PRE_SNV_CENTER, OCM_EINVAL, "unknown mode");
    OCM_REQUIRE(mode == OCM_PRE_SNV || mu, OCM_EINVAL, "mode needs mu");
    OCM_REQUIRE(mode != OCM_PRE_AUTOSCALE || sigma, OCM_EINVAL, "autoscale needs sigma");
    if (n == 0) return 0;
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    const bool vec = (p % 4 == 0) && (ldx % 4 == 0) && (ldy % 4 == 0) && al16(X) && al16(Y) &&
                     (!mu || al16(mu)) && (!sigma || al16(sigma));
    const int warps_per_block = 8;
    int64_t blocks = (n + warps_per_block - 1) / warps_per_block;
    int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    cudaStream_t st = (cudaStream_t)stream;
    if (vec)
        preprocess_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(X, n, p, ldx, mode, mu, sigma, eps, Y, ldy);
    else
        preprocess_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(X, n, p, ldx, mode, mu, sigma, eps, Y, ldy);
    return check_launch("preprocess_kernel");
}

// ---------------------------------------------------------------------------------------------------------
// Savitzky-Golay filter (optionally after SNV) -- SURVEY.md §8(f) f1: the step between SNV and SIMCA in every
// reference driver (simca_nuts.py:52, utils/data_utils.py:59: window 5, polyorder 2, deriv 1, scipy mode 'interp').
// A p-point FIR once the row is on chip: one warp per row, row staged in shared memory (after SNV if requested),
// interior points y[i] = sum_d g[d] x[i+d-h], the h points at each end from the polynomial fitted to the first /
// last `window` samples, given as two small matrices computed on the host from the same scipy routine.
// HBM-bound: 8*p bytes per spectrum.
namespace ocm {

constexpr int SG_MAXW = 33;

struct SavgolArgs {
    int window, half;
    float g[SG_MAXW];
};

__global__ void __launch_bounds__(128)
savgol_kernel(const float* __restrict__ X, int64_t n, int p, int64_t ldx, int snv, float eps, SavgolArgs sg,
              const float* __restrict__ edge_l, const float* __restrict__ edge_r, float* __restrict__ Y, int64_t ldy) {
    extern __shared__ float s_rows[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float* row = s_rows + (size_t)w * p;
    const int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + w;
    const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int h = sg.half, win = sg.window;
    for (int64_t r = wid; r < n; r += nw) {
        const float* x = X + r * ldx;
        float s = 0.f;
        for (int j = lane; j < p; j += 32) { const float v = x[j]; row[j] = v; s += v; }
        if (snv) {
            const float m = warp_sum(s) / (float)p;
            float ss = 0.f;
            for (int j = lane; j < p; j += 32) { const float d = row[j] - m; ss = fmaf(d, d, ss); }
            const float inv = 1.f / (sqrtf(warp_sum(ss) / (float)p) + eps);
            for (int j = lane; j < p; j += 32) row[j] = (row[j] - m) * inv;
        }
        __syncwarp();
        float* y = Y + r * ldy;
        for (int j = lane; j < p; j += 32) {
            float acc = 0.f;
            if (j >= h && j < p - h) {
#pragma unroll 5
                for (int d = 0; d < win; ++d) acc = fmaf(sg.g[d], row[j + d - h], acc);
            } else if (j < h) {
                for (int d = 0; d < win; ++d) acc = fmaf(edge_l[j * win + d], row[d], acc);
            } else {
                const int e = j - (p - h);
                for (int d = 0; d < win; ++d) acc = fmaf(edge_r[e * win + d], row[p - win + d], acc);
            }
            y[j] = acc;
        }
        __syncwarp();
    }
}

// Vector path: 128-bit loads / stores, window length known at compile time so the taps sit in the constant bank and
// the window lives in registers; SNV is applied as (x - m) on the way into the window and * inv on the way out, so the
// staged row is written once.  Per 4 outputs: 1 STS.128 + 1 LDS.128 (variance pass) + (1 + 2*ceil(h/4)) LDS.128.
template <int W>
__global__ void __launch_bounds__(128)
savgol_vec_kernel(const float* __restrict__ X, int64_t n, int p, int64_t ldx, int snv, float eps, SavgolArgs sg,
                  const float* __restrict__ edge_l, const float* __restrict__ edge_r, float* __restrict__ Y, int64_t ldy) {
    constexpr int H = W / 2, HP4 = (H + 3) / 4, HP = HP4 * 4, NV4 = 1 + 2 * HP4;
    extern __shared__ float4 s_rows4[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int nch = p >> 2;
    float4* row4 = s_rows4 + (size_t)w * nch;
    const float* row = reinterpret_cast<const float*>(row4);
    const int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + w;
    const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = wid; r < n; r += nw) {
        const float* x = X + r * ldx;
        float s = 0.f;
        for (int c = lane; c < nch; c += 32) {
            const float4 v = ldg_stream4(x + 4 * c);
            row4[c] = v;
            s += (v.x + v.y) + (v.z + v.w);
        }
        float m = 0.f, inv = 1.f;
        if (snv) {
            m = warp_sum(s) / (float)p;
            float ss = 0.f;
            for (int c = lane; c < nch; c += 32) {
                const float4 v = row4[c];
                const float a = v.x - m, b = v.y - m, cc = v.z - m, d = v.w - m;
                ss = fmaf(a, a, ss); ss = fmaf(b, b, ss); ss = fmaf(cc, cc, ss); ss = fmaf(d, d, ss);
            }
            inv = 1.f / (sqrtf(warp_sum(ss) / (float)p) + eps);
        }
        __syncwarp();
        float* y = Y + r * ldy;
        for (int c = lane; c < nch; c += 32) {
            float v[4 * NV4];
#pragma unroll
            for (int k = 0; k < NV4; ++k) {
                const int cc = c - HP4 + k;
                float4 t = make_float4(m, m, m, m);
                if (cc >= 0 && cc < nch) t = row4[cc];
                v[4 * k + 0] = t.x - m; v[4 * k + 1] = t.y - m; v[4 * k + 2] = t.z - m; v[4 * k + 3] = t.w - m;
            }
            float o[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float acc = 0.f;
#pragma unroll
                for (int d = 0; d < W; ++d) acc = fmaf(sg.g[d], v[HP - H + q + d], acc);
                o[q] = acc;
            }
            const int j0 = 4 * c;
            if (j0 < H || j0 + 3 >= p - H) {              // the polynomial-fit ends (mode 'interp')
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int j = j0 + q;
                    if (j < H) {
                        float acc = 0.f;
                        for (int d = 0; d < W; ++d) acc = fmaf(edge_l[j * W + d], row[d] - m, acc);
                        o[q] = acc;
                    } else if (j >= p - H) {
                        float acc = 0.f;
                        for (int d = 0; d < W; ++d) acc = fmaf(edge_r[(j - (p - H)) * W + d], row[p - W + d] - m, acc);
                        o[q] = acc;
                    }
                }
            }
            __stcs(reinterpret_cast<float4*>(y) + c, make_float4(o[0] * inv, o[1] * inv, o[2] * inv, o[3] * inv));
        }
        __syncwarp();
    }
}

// Same filter for rows of at most 128 * NV floats: the row also lives in registers (variance without a second trip
// through shared memory) and the NEXT row's loads are issued before the filter pass of the current one, so every warp
// keeps a row of HBM traffic in flight all the time (the version above loads, waits, computes: 81-84 % of the HBM peak).
template <int W, int NV>
__global__ void __launch_bounds__(128)
savgol_pf_kernel(const float* __restrict__ X, int64_t n, int p, int64_t ldx, int snv, float eps, SavgolArgs sg,
                 const float* __restrict__ edge_l, const float* __restrict__ edge_r, float* __restrict__ Y, int64_t ldy) {
    constexpr int H = W / 2, HP4 = (H + 3) / 4, HP = HP4 * 4, NV4 = 1 + 2 * HP4;
    extern __shared__ float4 s_rows4[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int nch = p >> 2;
    float4* row4 = s_rows4 + (size_t)w * nch;
    const float* row = reinterpret_cast<const float*>(row4);
    const int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + w;
    const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
    float4 nxt[NV];
    auto fetch = [&](int64_t r) {
        const float* x = X + r * ldx;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int c = lane + 32 * i;
            nxt[i] = (r < n && c < nch) ? ldg_stream4(x + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    fetch(wid);
    for (int64_t r = wid; r < n; r += nw) {
        float4 cur[NV];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            cur[i] = nxt[i];
            const int c = lane + 32 * i;
            if (c < nch) row4[c] = cur[i];
            s += (cur[i].x + cur[i].y) + (cur[i].z + cur[i].w);
        }
        fetch(r + nw);                                   // in flight during everything below
        float m = 0.f, inv = 1.f;
        if (snv) {
            m = warp_sum(s) / (float)p;
            float ss = 0.f;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                if (lane + 32 * i < nch) {
                    const float a = cur[i].x - m, b = cur[i].y - m, cc = cur[i].z - m, d = cur[i].w - m;
                    ss = fmaf(a, a, ss); ss = fmaf(b, b, ss); ss = fmaf(cc, cc, ss); ss = fmaf(d, d, ss);
                }
            }
            inv = 1.f / (sqrtf(warp_sum(ss) / (float)p) + eps);
        }
        __syncwarp();
        float* y = Y + r * ldy;
#pragma unroll 2
        for (int c = lane; c < nch; c += 32) {
            float v[4 * NV4];
#pragma unroll
            for (int k = 0; k < NV4; ++k) {
                const int cc = c - HP4 + k;
                float4 t = make_float4(m, m, m, m);
                if (cc >= 0 && cc < nch) t = row4[cc];
                v[4 * k + 0] = t.x - m; v[4 * k + 1] = t.y - m; v[4 * k + 2] = t.z - m; v[4 * k + 3] = t.w - m;
            }
            float o[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float acc = 0.f;
#pragma unroll
                for (int d = 0; d < W; ++d) acc = fmaf(sg.g[d], v[HP - H + q + d], acc);
                o[q] = acc;
            }
            const int j0 = 4 * c;
            if (j0 < H || j0 + 3 >= p - H) {              // the polynomial-fit ends (mode 'interp')
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int j = j0 + q;
                    if (j < H) {
                        float acc = 0.f;
                        for (int d = 0; d < W; ++d) acc = fmaf(edge_l[j * W + d], row[d] - m, acc);
                        o[q] = acc;
                    } else if (j >= p - H) {
                        float acc = 0.f;
                        for (int d = 0; d < W; ++d) acc = fmaf(edge_r[(j - (p - H)) * W + d], row[p - W + d] - m, acc);
                        o[q] = acc;
                    }
                }
            }
            __stcs(reinterpret_cast<float4*>(y) + c, make_float4(o[0] * inv, o[1] * inv, o[2] * inv, o[3] * inv));
        }
        __syncwarp();
    }
}

template <int W>
static cudaError_t launch_savgol_vec(unsigned blocks, size_t smem, cudaStream_t st, const float* X, int64_t n, int p, int64_t ldx,
                                     int snv, float eps, const SavgolArgs& sg, const float* el, const float* er, float* Y, int64_t ldy) {
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(savgol_vec_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    if (p <= 512) {
        auto k = savgol_pf_kernel<W, 4>;
        if (smem > 48 * 1024) { cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); if (e != cudaSuccess) return e; }
        k<<<blocks, 128, smem, st>>>(X, n, p, ldx, snv, eps, sg, el, er, Y, ldy);
    } else if (p <= 1024) {
        auto k = savgol_pf_kernel<W, 8>;
        if (smem > 48 * 1024) { cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); if (e != cudaSuccess) return e; }
        k<<<blocks, 128, smem, st>>>(X, n, p, ldx, snv, eps, sg, el, er, Y, ldy);
    } else {
        savgol_vec_kernel<W><<<blocks, 128, smem, st>>>(X, n, p, ldx, snv, eps, sg, el, er, Y, ldy);
    }
    return cudaSuccess;
}

}  // namespace ocm

extern "C" int ocm_savgol_f32(const float* X, int64_t n, int32_t p, int64_t ldx, int32_t snv, float eps,
                              const float* h_coef, int32_t window, const float* edge_l, const float* edge_r,
                              float* Y, int64_t ldy, void* stream) {
    using namespace ocm;
    if (n == 0) return 0;
    OCM_REQUIRE(X && Y && h_coef && edge_l && edge_r && n > 0 && p > 0 && ldx >= p && ldy >= p, OCM_EINVAL, "bad shape or NULL buffer");
    OCM_REQUIRE(window >= 3 && (window & 1) && window <= SG_MAXW && window <= p, OCM_ELIMIT, "window must be odd, 3..33 and <= p");
    OCM_REQUIRE(X != Y, OCM_EINVAL, "in-place filtering is not supported");
    OCM_REQUIRE((size_t)p * 4 * 4 <= 200 * 1024, OCM_ELIMIT, "row too long for the shared-memory stage (p <= 12800)");
    SavgolArgs sg;
    sg.window = window; sg.half = window / 2;
    for (int i = 0; i < window; ++i) sg.g[i] = h_coef[i];
    const size_t smem = (size_t)4 * p * sizeof(float);
    if (smem > 48 * 1024) OCM_CUDA(cudaFuncSetAttribute(savgol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t blocks = (n + 3) / 4;
    // as many resident CTAs per SM as shared memory allows (each warp loads a whole row and then waits: the bytes in
    // flight per SM are what hides the HBM latency), at most 16
    int64_t per_sm = (int64_t)(200 * 1024) / (int64_t)(smem + 1024);
    if (per_sm > 16) per_sm = 16;
    if (per_sm < 1) per_sm = 1;
    const int64_t cap = (int64_t)sm_count() * per_sm;
    if (blocks > cap) blocks = cap;
    const bool vec = p % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0 && ((uintptr_t)X & 15) == 0 && ((uintptr_t)Y & 15) == 0;
    if (vec && window >= 5 && window <= 15) {
        cudaError_t e = cudaSuccess;
        const cudaStream_t st = (cudaStream_t)stream;
        switch (window) {
#define OCM_SG_CASE(W_) case W_: e = launch_savgol_vec<W_>((unsigned)blocks, smem, st, X, n, p, ldx, snv, eps, sg, edge_l, edge_r, Y, ldy); break;
            OCM_SG_CASE(5) OCM_SG_CASE(7) OCM_SG_CASE(9) OCM_SG_CASE(11) OCM_SG_CASE(13) OCM_SG_CASE(15)
#undef OCM_SG_CASE
        }
        OCM_CUDA(e);
        return check_launch("savgol_vec_kernel");
    }
    savgol_kernel<<<(unsigned)blocks, 128, smem, (cudaStream_t)stream>>>(X, n, p, ldx, snv, eps, sg, edge_l, edge_r, Y, ldy);
    return check_launch("savgol_kernel");
}
