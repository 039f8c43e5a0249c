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
