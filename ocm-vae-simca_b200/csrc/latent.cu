// VAE-SIMCA scoring epilogue: latent Mahalanobis distance and squared
// reconstruction residual in one pass, so that the per-batch device->host round trips of
// utils/final_vaesimca.py:479-492 / VAE_SIMCA.py:359-365 disappear.
//   d2 = (z - m)^T Sinv (z - m)        (Sinv full symmetric: the latent covariance is not diagonal)
//   q  = sum_j (x_j - xrec_j)^2        (xrec fp32 or bf16)
// HBM-bound: 4p (x) + 4p|2p (xrec) + 4d (z) + 8 B out per spectrum.  One warp per row.
#include "ocm_common.cuh"

namespace ocm {

template <bool BF16, bool VEC>
__global__ void __launch_bounds__(256)
latent_score_kernel(const float* __restrict__ z, int64_t n, int d, int64_t ldz,
                    const float* __restrict__ m, const float* __restrict__ Sinv,
                    const float* __restrict__ x, const void* __restrict__ xrec_, int p,
                    int64_t ldx, int64_t ldxr, float* __restrict__ d2, float* __restrict__ q) {
    extern __shared__ float sm[];
    float* s_S = sm;                 // d*d
    float* s_m = sm + d * d;         // d
    float* s_diff = s_m + d;         // warps * d
    const bool do_lat = z != nullptr;
    if (do_lat) {
        for (int i = threadIdx.x; i < d * d; i += blockDim.x) s_S[i] = Sinv[i];
        for (int i = threadIdx.x; i < d; i += blockDim.x) s_m[i] = m[i];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float* diff = s_diff + w * d;
    const int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + w;
    const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t row = wid; row < n; row += nw) {
        if (do_lat) {
            for (int i = lane; i < d; i += 32) diff[i] = z[row * ldz + i] - s_m[i];
            __syncwarp();
            float part = 0.f;
            for (int i = lane; i < d; i += 32) {
                float s = 0.f;
                for (int k = 0; k < d; ++k) s = fmaf(s_S[i * d + k], diff[k], s);
                part = fmaf(s, diff[i], part);
            }
            part = warp_sum(part);
            if (lane == 0) d2[row] = part;
            __syncwarp();
        }
        if (x != nullptr) {
            const float* xr = x + row * ldx;
            float acc = 0.f;
            if constexpr (VEC) {
                for (int j = lane << 2; j < p; j += 128) {
                    const float4 a = ldg_stream4(xr + j);
                    float4 b;
                    if constexpr (BF16) {
                        const __nv_bfloat16* r = reinterpret_cast<const __nv_bfloat16*>(xrec_) + row * ldxr + j;
                        const uint2 raw = *reinterpret_cast<const uint2*>(r);
                        const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&raw.x);
                        const __nv_bfloat162 hi = *reinterpret_cast<const __nv_bfloat162*>(&raw.y);
                        b = make_float4(__low2float(lo), __high2float(lo), __low2float(hi), __high2float(hi));
                    } else {
                        b = ldg_stream4(reinterpret_cast<const float*>(xrec_) + row * ldxr + j);
                    }
                    const float e0 = a.x - b.x, e1 = a.y - b.y, e2 = a.z - b.z, e3 = a.w - b.w;
                    acc += (e0 * e0 + e1 * e1) + (e2 * e2 + e3 * e3);
                }
            } else {
                for (int j = lane; j < p; j += 32) {
                    float b;
                    if constexpr (BF16) b = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(xrec_)[row * ldxr + j]);
                    else b = reinterpret_cast<const float*>(xrec_)[row * ldxr + j];
                    const float e = xr[j] - b;
                    acc = fmaf(e, e, acc);
                }
            }
            acc = warp_sum(acc);
            if (lane == 0) q[row] = acc;
        }
    }
}

}  // namespace ocm

extern "C" int ocm_latent_score_f32(const float* z, int64_t n, int32_t d, int64_t ldz, const float* m,
                                    const float* Sinv, const float* x, const void* xrec,
                                    int32_t xrec_is_bf16, int32_t p, int64_t ldx, int64_t ldxrec,
                                    float* d2, float* q, void* stream) {
    using namespace ocm;
    if (n == 0) return 0;
    OCM_REQUIRE(n > 0, OCM_EINVAL, "bad n");
    OCM_REQUIRE(z || x, OCM_EINVAL, "nothing to score");
    if (z) {
        OCM_REQUIRE(m && Sinv && d2 && d >= 1 && ldz >= d, OCM_EINVAL, "latent part needs m, Sinv, d2");
        OCM_REQUIRE(d <= 96, OCM_ELIMIT, "latent dim must be <= 96");
    }
    if (x) OCM_REQUIRE(xrec && q && p >= 1 && ldx >= p && ldxrec >= p, OCM_EINVAL, "recon part needs xrec, q");
    if (n == 0) return 0;
    const int dd = z ? d : 0;
    const size_t smem = sizeof(float) * ((size_t)dd * dd + dd + 8 * (size_t)dd);
    auto al = [](const void* q_, size_t a) { return (reinterpret_cast<uintptr_t>(q_) & (a - 1)) == 0; };
    bool vec = x && (p % 4 == 0) && (ldx % 4 == 0) && (ldxrec % 4 == 0) && al(x, 16) &&
               al(xrec, xrec_is_bf16 ? 8 : 16);
    int64_t blocks = (n + 7) / 8;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    cudaStream_t st = (cudaStream_t)stream;
#define OCM_LAUNCH_LAT(B, V) \
    latent_score_kernel<B, V><<<(unsigned)blocks, 256, smem, st>>>(z, n, dd, ldz, m, Sinv, x, xrec, p, ldx, ldxrec, d2, q)
    if (xrec_is_bf16) { if (vec) OCM_LAUNCH_LAT(true, true); else OCM_LAUNCH_LAT(true, false); }
    else              { if (vec) OCM_LAUNCH_LAT(false, true); else OCM_LAUNCH_LAT(false, false); }
#undef OCM_LAUNCH_LAT
    return check_launch("latent_score_kernel");
}
