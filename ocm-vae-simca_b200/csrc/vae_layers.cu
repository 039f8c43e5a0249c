// Two elementwise-bound pieces of the ConvVAE1D scoring pass (config 5; vae_model.py:6-129 run by
// utils/final_vaesimca.py:471-533), written for channels-last bf16 activations [n][L][C]:
//
//  * vae_conv_in_kernel: the FIRST encoder block in one pass -- standardise the fp32 spectrum ((x - mean) / std),
//    Conv1d(1 -> C, K taps, stride 1, zero padding K/2) with eval-mode BatchNorm folded into weights and bias, activation,
//    bf16 channels-last output.  A 1-channel convolution has no GEMM in it (K = 9 products per output): cuDNN ran it as
//    an fp32 -> tf32 implicit GEMM with two layout conversions, 1.6 ms per 8192 x 1024 batch; this kernel is bound by
//    its 64 KB of output per spectrum.
//  * bias_act_bf16_kernel: y <- act(y + bias[c]) in place.  cuDNN's fused bias + activation epilogue knows ReLU only,
//    so PyTorch runs a broadcast add_ and an ELU as two passes over every activation tensor (28.7 % + 19.9 % of the
//    scoring pass, torch.profiler); this is one pass at the stream rate.
// The GEMM-shaped convolutions in between stay cuDNN's (north-star item 5 keeps the network in PyTorch).
#include "ocm_common.cuh"

namespace ocm {
namespace vae {

constexpr int ACT_NONE = 0, ACT_ELU = 1, ACT_GELU = 2;

__device__ __forceinline__ float activate(float v, int act) {
    if (act == ACT_ELU) return v > 0.f ? v : __expf(v) - 1.f;                        // nn.ELU(alpha = 1)
    if (act == ACT_GELU) return 0.5f * v * (1.f + erff(v * 0.70710678118654752f));   // nn.GELU(approximate = 'none')
    return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}

constexpr int TILE_L = 256;      // positions per block (one per thread)
constexpr int K_MAX = 15;

template <int C>
__global__ void __launch_bounds__(TILE_L)
vae_conv_in_kernel(const float* __restrict__ X, int64_t n, int L, int64_t ldx, const float* __restrict__ mean,
                   const float* __restrict__ stdv, const float* __restrict__ w, const float* __restrict__ bias, int K, int act,
                   __nv_bfloat16* __restrict__ out, int tiles_per_row) {
    __shared__ __align__(16) float s_w[K_MAX * C];      // [k][c]: the C weights of one tap are contiguous (broadcast LDS.128)
    __shared__ __align__(16) float s_b[C];
    __shared__ float s_x[TILE_L + K_MAX - 1];
    const int pad = K >> 1;
    for (int i = threadIdx.x; i < K * C; i += TILE_L) { const int k = i / C, c = i % C; s_w[i] = w[c * K + k]; }
    for (int i = threadIdx.x; i < C; i += TILE_L) s_b[i] = bias[i];
    const int64_t ntiles = n * tiles_per_row;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t row = tile / tiles_per_row;
        const int l0 = (int)(tile % tiles_per_row) * TILE_L;
        __syncthreads();                                  // previous tile's s_x fully consumed (and the weights are in place)
        for (int i = threadIdx.x; i < TILE_L + K - 1; i += TILE_L) {
            const int l = l0 + i - pad;
            s_x[i] = (l >= 0 && l < L) ? (X[row * ldx + l] - mean[l]) / stdv[l] : 0.f;      // zero padding of the standardised row
        }
        __syncthreads();
        const int l = l0 + threadIdx.x;
        if (l >= L) continue;
        float acc[C];
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] = s_b[c];
        for (int k = 0; k < K; ++k) {
            const float xv = s_x[threadIdx.x + k];
            const float4* wk = reinterpret_cast<const float4*>(s_w + k * C);
#pragma unroll
            for (int c4 = 0; c4 < C / 4; ++c4) {
                const float4 ww = wk[c4];
                acc[4 * c4] = fmaf(ww.x, xv, acc[4 * c4]);         acc[4 * c4 + 1] = fmaf(ww.y, xv, acc[4 * c4 + 1]);
                acc[4 * c4 + 2] = fmaf(ww.z, xv, acc[4 * c4 + 2]); acc[4 * c4 + 3] = fmaf(ww.w, xv, acc[4 * c4 + 3]);
            }
        }
        uint4* o = reinterpret_cast<uint4*>(out + (row * L + l) * C);
#pragma unroll
        for (int c8 = 0; c8 < C / 8; ++c8) {
            uint4 v;
            v.x = pack_bf16x2(activate(acc[8 * c8], act), activate(acc[8 * c8 + 1], act));
            v.y = pack_bf16x2(activate(acc[8 * c8 + 2], act), activate(acc[8 * c8 + 3], act));
            v.z = pack_bf16x2(activate(acc[8 * c8 + 4], act), activate(acc[8 * c8 + 5], act));
            v.w = pack_bf16x2(activate(acc[8 * c8 + 6], act), activate(acc[8 * c8 + 7], act));
            o[c8] = v;
        }
    }
}

// y[r][c] <- act(y[r][c] + bias[c]) for a dense [rows][C] bf16 array, C % 8 == 0; 16 bytes per thread and step
__global__ void __launch_bounds__(256)
bias_act_bf16_kernel(__nv_bfloat16* __restrict__ Y, int64_t chunks, int C, const float* __restrict__ bias, int act) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < chunks; i += stride) {
        uint4* p = reinterpret_cast<uint4*>(Y) + i;
        const int c0 = (int)((i * 8) % C);
        const float4 b0 = *reinterpret_cast<const float4*>(bias + c0), b1 = *reinterpret_cast<const float4*>(bias + c0 + 4);
        uint4 v = *p;
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
        uint4 r;
        r.x = pack_bf16x2(activate(__low2float(h[0]) + b0.x, act), activate(__high2float(h[0]) + b0.y, act));
        r.y = pack_bf16x2(activate(__low2float(h[1]) + b0.z, act), activate(__high2float(h[1]) + b0.w, act));
        r.z = pack_bf16x2(activate(__low2float(h[2]) + b1.x, act), activate(__high2float(h[2]) + b1.y, act));
        r.w = pack_bf16x2(activate(__low2float(h[3]) + b1.z, act), activate(__high2float(h[3]) + b1.w, act));
        *p = r;
    }
}

// Tail of the decoder in one pass: y [n][Ly][C] is the LAST transposed convolution's raw output (bias not yet added);
// a = act(y + bias), s = sum_c a_c w_out[c] + b_out (the 1x1 output convolution), x_rec = s std_l + mean_l (positions
// l >= Ly: s = 0, what F.pad gives), q = sum_l (x_l - x_rec_l)^2.  Neither the activated tensor nor the reconstruction
// goes back to HBM (x_rec optionally, as bf16).  One warp per spectrum, a lane per position: 2 C bytes each, contiguous.
template <int C>
__global__ void __launch_bounds__(256)
vae_tail_q_kernel(const __nv_bfloat16* __restrict__ Y, int64_t n, int Ly, const float* __restrict__ bias, int act,
                  const float* __restrict__ w_out, float b_out, const float* __restrict__ X, int64_t ldx, int L,
                  const float* __restrict__ mean, const float* __restrict__ stdv, float* __restrict__ q,
                  __nv_bfloat16* __restrict__ xrec) {
    __shared__ __align__(16) float s_b[C], s_w[C];
    for (int i = threadIdx.x; i < C; i += blockDim.x) { s_b[i] = bias[i]; s_w[i] = w_out[i]; }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t row = wid; row < n; row += nw) {
        float acc = 0.f;
        for (int l = lane; l < L; l += 32) {
            float s = 0.f;
            if (l < Ly) {
                const uint4* yp = reinterpret_cast<const uint4*>(Y + (row * Ly + l) * C);
                s = b_out;
#pragma unroll
                for (int c8 = 0; c8 < C / 8; ++c8) {
                    const uint4 v = yp[c8];
                    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int c = 8 * c8 + 2 * k;
                        s = fmaf(activate(__low2float(h[k]) + s_b[c], act), s_w[c], s);
                        s = fmaf(activate(__high2float(h[k]) + s_b[c + 1], act), s_w[c + 1], s);
                    }
                }
            }
            const float xr = fmaf(s, stdv[l], mean[l]);
            if (xrec) xrec[row * L + l] = __float2bfloat16(xr);
            const float e = X[row * ldx + l] - xr;
            acc = fmaf(e, e, acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) q[row] = acc;
    }
}

}  // namespace vae
}  // namespace ocm

extern "C" int ocm_vae_tail_q_bf16(const void* Y, int64_t n, int32_t Ly, int32_t C, const float* bias, int32_t act,
                                   const float* w_out, float b_out, const float* X, int64_t ldx, int32_t L,
                                   const float* spec_mean, const float* spec_std, float* q, void* xrec, void* stream) {
    using namespace ocm;
    using namespace ocm::vae;
    if (n == 0) return 0;
    OCM_REQUIRE(Y && bias && w_out && X && spec_mean && spec_std && q && n > 0 && Ly > 0 && L > 0 && ldx >= L, OCM_EINVAL,
                "bad shape or NULL buffer");
    OCM_REQUIRE(C == 16 || C == 32 || C == 64, OCM_ELIMIT, "last block: 16 / 32 / 64 filters");
    OCM_REQUIRE(act >= ACT_NONE && act <= ACT_GELU, OCM_EINVAL, "unknown activation");
    OCM_REQUIRE((reinterpret_cast<uintptr_t>(Y) & 15) == 0, OCM_EALIGN, "Y must be 16-byte aligned");
    int64_t blocks = (n + 7) / 8;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    cudaStream_t st = (cudaStream_t)stream;
    const __nv_bfloat16* y = static_cast<const __nv_bfloat16*>(Y);
    __nv_bfloat16* xr = static_cast<__nv_bfloat16*>(xrec);
#define OCM_TAIL(C_) vae_tail_q_kernel<C_><<<(unsigned)blocks, 256, 0, st>>>(y, n, Ly, bias, act, w_out, b_out, X, ldx, L, spec_mean, spec_std, q, xr)
    if (C == 16) OCM_TAIL(16); else if (C == 32) OCM_TAIL(32); else OCM_TAIL(64);
#undef OCM_TAIL
    return check_launch("vae_tail_q_kernel");
}

extern "C" int ocm_vae_conv_in_bf16(const float* X, int64_t n, int32_t L, int64_t ldx, const float* spec_mean,
                                    const float* spec_std, const float* w, const float* bias, int32_t C, int32_t K,
                                    int32_t act, void* out, void* stream) {
    using namespace ocm;
    using namespace ocm::vae;
    if (n == 0) return 0;
    OCM_REQUIRE(X && spec_mean && spec_std && w && bias && out && n > 0 && L > 0 && ldx >= L, OCM_EINVAL, "bad shape or NULL buffer");
    OCM_REQUIRE((C == 16 || C == 32 || C == 64) && K >= 1 && K <= K_MAX && (K & 1), OCM_ELIMIT,
                "first block: 16 / 32 / 64 filters, odd kernel size <= 15");
    OCM_REQUIRE(act >= ACT_NONE && act <= ACT_GELU, OCM_EINVAL, "unknown activation");
    OCM_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, OCM_EALIGN, "output must be 16-byte aligned");
    const int tiles_per_row = (L + TILE_L - 1) / TILE_L;
    int64_t blocks = n * tiles_per_row;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    cudaStream_t st = (cudaStream_t)stream;
    __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out);
    if (C == 16) vae_conv_in_kernel<16><<<(unsigned)blocks, TILE_L, 0, st>>>(X, n, L, ldx, spec_mean, spec_std, w, bias, K, act, o, tiles_per_row);
    else if (C == 32) vae_conv_in_kernel<32><<<(unsigned)blocks, TILE_L, 0, st>>>(X, n, L, ldx, spec_mean, spec_std, w, bias, K, act, o, tiles_per_row);
    else vae_conv_in_kernel<64><<<(unsigned)blocks, TILE_L, 0, st>>>(X, n, L, ldx, spec_mean, spec_std, w, bias, K, act, o, tiles_per_row);
    return check_launch("vae_conv_in_kernel");
}

extern "C" int ocm_bias_act_bf16(void* Y, int64_t rows, int32_t C, const float* bias, int32_t act, void* stream) {
    using namespace ocm;
    using namespace ocm::vae;
    if (rows == 0) return 0;
    OCM_REQUIRE(Y && bias && rows > 0 && C > 0, OCM_EINVAL, "bad shape or NULL buffer");
    OCM_REQUIRE(C % 8 == 0, OCM_ELIMIT, "channel count must be a multiple of 8");
    OCM_REQUIRE(act >= ACT_NONE && act <= ACT_GELU, OCM_EINVAL, "unknown activation");
    OCM_REQUIRE((reinterpret_cast<uintptr_t>(Y) & 15) == 0 && (reinterpret_cast<uintptr_t>(bias) & 15) == 0, OCM_EALIGN,
                "Y and bias must be 16-byte aligned");
    const int64_t chunks = rows * (C / 8);
    int64_t blocks = (chunks + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    bias_act_bf16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(static_cast<__nv_bfloat16*>(Y), chunks, C, bias, act);
    return check_launch("bias_act_bf16_kernel");
}
