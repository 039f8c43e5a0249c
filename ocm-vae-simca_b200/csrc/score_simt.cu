// Fused SIMCA scoring on the FP32 pipe -- the general path:
// any p, any leading dimension, fp32 or fp64 spectra, A up to 64, optional fused SNV,
// single model or all nested models A = 1..Amax (CV grid).
//
// Replaces, per row, pca.transform + pca.inverse_transform + the two reductions + the
// decision of utils/SIMCA.py:127-145 with one read of the spectrum:
//     xc = pre(x) - mu ; t = P xc ; r = xc - P^T t ; T2 = t^T Sinv t ; Q = |r|^2 ;
//     accept = rule(T2*ct, Q*cq) < dlim
// Q keeps the EXPLICIT residual: |xc|^2 - |t|^2 loses 1e-5 parity to cancellation
//  while errors of t enter |r|^2 only at second order.
//
// Mapping: one warp owns R rows at a time, lanes stride the wavelength axis (coalesced
// 128 B per row per step), P and mu live in shared memory, every P element read from
// shared memory feeds R FMAs.  The projection needs the whole row before the residual can
// start, so the row is read twice; the second read hits L1/L2 (R*4p bytes per warp).
// The tensor-core path (score_tc.cu) removes the FP32-pipe ceiling for the bulk config.
#include "ocm_common.cuh"

namespace ocm {

template <int AMAX> struct RowsPerWarp { static constexpr int value = AMAX <= 12 ? 4 : (AMAX <= 32 ? 2 : 1); };

struct ScoreArgs {
    int64_t n;
    int p;
    int64_t ldx;
    int A;          // components actually used (<= AMAX)
    int rule;
    int snv;
    float snv_eps;
    float ct, cq, dlim;                 // single-model scales
    const float* ctv; const float* cqv; const float* dlimv;   // per-A scales (MULTI)
    const float* P;                     // [A][p]
    const float* W;                     // Sinv [A][A] (single) or inv_eig [A] (MULTI)
    float* T2; float* Q; uint8_t* accept; float* scores;
    int p_in_smem;
};

template <typename TIn, int AMAX, bool MULTI>
__global__ void __launch_bounds__(256)
score_simt_kernel(const TIn* __restrict__ X, const TIn* __restrict__ mu, ScoreArgs a) {
    constexpr int R = RowsPerWarp<AMAX>::value;
    extern __shared__ __align__(16) float smem[];
    // layout: [mu_f32 (p, fp32 input only)] [W (A*A or A)] [P (AMAX * p) if it fits]
    float* s_mu = smem;
    const int mu_floats = (sizeof(TIn) == 4) ? a.p : 0;
    float* s_W = smem + mu_floats;
    const int w_floats = MULTI ? AMAX : a.A * a.A;
    float* s_P = s_W + ((w_floats + 3) & ~3);

    for (int i = threadIdx.x; i < mu_floats; i += blockDim.x) s_mu[i] = (float)mu[i];
    for (int i = threadIdx.x; i < w_floats; i += blockDim.x)
        s_W[i] = MULTI ? (i < a.A ? a.W[i] : 0.f) : a.W[i];
    if (a.p_in_smem) {
        for (int i = threadIdx.x; i < AMAX * a.p; i += blockDim.x) {
            const int c = i / a.p;
            s_P[i] = c < a.A ? a.P[i] : 0.f;
        }
    }
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int p = a.p;

    auto Pval = [&](int c, int j) -> float {
        if (a.p_in_smem) return s_P[c * p + j];
        return c < a.A ? __ldg(a.P + (int64_t)c * p + j) : 0.f;
    };

    for (int64_t g = wid * R; g < a.n; g += nw * R) {
        const TIn* xr[R];
        bool live[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            live[r] = g + r < a.n;
            xr[r] = X + (live[r] ? g + r : g) * a.ldx;   // dead rows alias row g (results dropped)
        }
        // ---- optional SNV statistics (population std, eps on the std) -------------------
        float rm[R], rinv[R];
#pragma unroll
        for (int r = 0; r < R; ++r) { rm[r] = 0.f; rinv[r] = 1.f; }
        if (a.snv) {
            float s[R];
#pragma unroll
            for (int r = 0; r < R; ++r) s[r] = 0.f;
            for (int j = lane; j < p; j += 32)
#pragma unroll
                for (int r = 0; r < R; ++r) s[r] += (float)xr[r][j];
#pragma unroll
            for (int r = 0; r < R; ++r) { rm[r] = warp_sum(s[r]) / (float)p; s[r] = 0.f; }
            for (int j = lane; j < p; j += 32)
#pragma unroll
                for (int r = 0; r < R; ++r) { float d = (float)xr[r][j] - rm[r]; s[r] = fmaf(d, d, s[r]); }
#pragma unroll
            for (int r = 0; r < R; ++r) rinv[r] = 1.f / (sqrtf(warp_sum(s[r]) / (float)p) + a.snv_eps);
        }
        auto centred = [&](int r, int j) -> float {
            if (a.snv) {
                const float v = ((float)xr[r][j] - rm[r]) * rinv[r];
                return v - (sizeof(TIn) == 4 ? s_mu[j] : (float)mu[j]);
            }
            if constexpr (sizeof(TIn) == 4) return (float)xr[r][j] - s_mu[j];
            else return (float)(xr[r][j] - mu[j]);    // centre in fp64, then round once
        };

        // ---- pass 1: scores t = P xc ------------------------------------------------------
        float t[R][AMAX];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int c = 0; c < AMAX; ++c) t[r][c] = 0.f;
        for (int j = lane; j < p; j += 32) {
            float xc[R];
#pragma unroll
            for (int r = 0; r < R; ++r) xc[r] = centred(r, j);
#pragma unroll
            for (int c = 0; c < AMAX; ++c) {
                const float pv = Pval(c, j);
#pragma unroll
                for (int r = 0; r < R; ++r) t[r][c] = fmaf(xc[r], pv, t[r][c]);
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int c = 0; c < AMAX; ++c) t[r][c] = warp_sum(t[r][c]);

        // ---- pass 2: explicit residual ----------------------------------------------------
        float q[R];
#pragma unroll
        for (int r = 0; r < R; ++r) q[r] = 0.f;
        for (int j = lane; j < p; j += 32) {
            float rr[R];
#pragma unroll
            for (int r = 0; r < R; ++r) rr[r] = centred(r, j);
#pragma unroll
            for (int c = 0; c < AMAX; ++c) {
                const float pv = Pval(c, j);
#pragma unroll
                for (int r = 0; r < R; ++r) rr[r] = fmaf(-t[r][c], pv, rr[r]);
            }
#pragma unroll
            for (int r = 0; r < R; ++r) q[r] = fmaf(rr[r], rr[r], q[r]);
        }
#pragma unroll
        for (int r = 0; r < R; ++r) q[r] = warp_sum(q[r]);

        // ---- distances + decision ---------------------------------------------------------
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (!live[r]) continue;
            const int64_t row = g + r;
            if constexpr (MULTI) {
                // lane c (and c+32) owns model A = c+1: prefix T2, suffix-corrected Q
                for (int c0 = 0; c0 < AMAX; c0 += 32) {
                    const int c = c0 + lane;
                    float t2 = 0.f, qq = q[r];
#pragma unroll
                    for (int k = 0; k < AMAX; ++k) {
                        const float tk2 = t[r][k] * t[r][k];
                        if (k <= c) t2 = fmaf(tk2, s_W[k], t2);
                        else if (k < a.A) qq += tk2;
                    }
                    if (c < a.A) {
                        const int64_t o = row * a.A + c;
                        if (a.T2) a.T2[o] = t2;
                        if (a.Q) a.Q[o] = qq;
                        if (a.accept)
                            a.accept[o] = rule_distance(a.rule, t2 * a.ctv[c], qq * a.cqv[c]) < a.dlimv[c] ? 1 : 0;
                    }
                }
            } else {
                // T2 = t^T W t : lane c accumulates row c of W
                float part = 0.f;
                for (int c = lane; c < a.A; c += 32) {
                    float s = 0.f;
#pragma unroll
                    for (int k = 0; k < AMAX; ++k)
                        if (k < a.A) s = fmaf(s_W[c * a.A + k], t[r][k], s);
                    // t[r][c] with a runtime c: select without dynamic register indexing
                    float tc = 0.f;
#pragma unroll
                    for (int k = 0; k < AMAX; ++k) tc = (k == c) ? t[r][k] : tc;
                    part = fmaf(s, tc, part);
                }
                const float t2 = warp_sum(part);
                if (a.scores) {
                    for (int c = lane; c < a.A; c += 32) {
                        float tc = 0.f;
#pragma unroll
                        for (int k = 0; k < AMAX; ++k) tc = (k == c) ? t[r][k] : tc;
                        a.scores[row * a.A + c] = tc;
                    }
                }
                if (lane == 0) {
                    if (a.T2) a.T2[row] = t2;
                    if (a.Q) a.Q[row] = q[r];
                    if (a.accept) a.accept[row] = rule_distance(a.rule, t2 * a.ct, q[r] * a.cq) < a.dlim ? 1 : 0;
                }
            }
        }
    }
}

template <typename TIn, int AMAX, bool MULTI>
int launch_score(const TIn* X, const TIn* mu, ScoreArgs& a, cudaStream_t st) {
    constexpr int R = RowsPerWarp<AMAX>::value;
    auto kern = score_simt_kernel<TIn, AMAX, MULTI>;
    const size_t mu_floats = (sizeof(TIn) == 4) ? a.p : 0;
    const size_t w_floats = MULTI ? AMAX : (size_t)a.A * a.A;
    const size_t base = (mu_floats + ((w_floats + 3) & ~(size_t)3)) * sizeof(float);
    const size_t with_p = base + (size_t)AMAX * a.p * sizeof(float);
    size_t smem = with_p;
    a.p_in_smem = 1;
    if (with_p > 200 * 1024) { smem = base; a.p_in_smem = 0; }
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    }
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, smem);
    if (occ < 1) occ = 1;
    int64_t groups = (a.n + R - 1) / R;
    int64_t blocks = (groups + 7) / 8;
    const int64_t cap = (int64_t)sm_count() * occ;
    if (blocks > cap) blocks = cap;
    kern<<<(unsigned)blocks, 256, smem, st>>>(X, mu, a);
    return check_launch("score_simt_kernel");
}

template <typename TIn, bool MULTI>
int dispatch_score(const TIn* X, const TIn* mu, ScoreArgs& a, cudaStream_t st) {
    const int A = a.A;
    if (A <= 4)  return launch_score<TIn, 4, MULTI>(X, mu, a, st);
    if (A <= 8)  return launch_score<TIn, 8, MULTI>(X, mu, a, st);
    if (A <= 12) return launch_score<TIn, 12, MULTI>(X, mu, a, st);
    if (A <= 16) return launch_score<TIn, 16, MULTI>(X, mu, a, st);
    if (A <= 20) return launch_score<TIn, 20, MULTI>(X, mu, a, st);
    if (A <= 24) return launch_score<TIn, 24, MULTI>(X, mu, a, st);
    if (A <= 32) return launch_score<TIn, 32, MULTI>(X, mu, a, st);
    return launch_score<TIn, 64, MULTI>(X, mu, a, st);
}

template <typename TIn>
int score_single(const TIn* X, int64_t n, int32_t p, int64_t ldx, const TIn* mu, const float* P,
                 const float* Sinv, const ocm_score_params* hp, float* T2, float* Q, uint8_t* accept,
                 float* scores, void* stream) {
    if (n == 0) return 0;
    OCM_REQUIRE(X && mu && P && Sinv && hp, OCM_EINVAL, "NULL buffer");
    OCM_REQUIRE(n > 0 && p > 0 && ldx >= p, OCM_EINVAL, "bad shape");
    OCM_REQUIRE(hp->A >= 1 && hp->A <= 64, OCM_ELIMIT, "n_components must be 1..64");
    OCM_REQUIRE(hp->rule >= OCM_RULE_SIM && hp->rule <= OCM_RULE_DD, OCM_EINVAL, "unknown rule");
    if (n == 0) return 0;
    ScoreArgs a{};
    a.n = n; a.p = p; a.ldx = ldx; a.A = hp->A; a.rule = hp->rule; a.snv = hp->snv; a.snv_eps = hp->snv_eps;
    a.ct = hp->ct; a.cq = hp->cq; a.dlim = hp->dlim; a.P = P; a.W = Sinv; a.T2 = T2; a.Q = Q; a.accept = accept; a.scores = scores;
    return dispatch_score<TIn, false>(X, mu, a, (cudaStream_t)stream);
}

template <typename TIn>
int score_multi(const TIn* X, int64_t n, int32_t p, int64_t ldx, const TIn* mu, const float* P,
                const float* inv_eig, int32_t Amax, int32_t rule, int32_t snv, float snv_eps,
                const float* ct, const float* cq, const float* dlim, float* T2, float* Q,
                uint8_t* accept, void* stream) {
    if (n == 0) return 0;
    OCM_REQUIRE(X && mu && P && inv_eig, OCM_EINVAL, "NULL buffer");
    OCM_REQUIRE(n > 0 && p > 0 && ldx >= p, OCM_EINVAL, "bad shape");
    OCM_REQUIRE(Amax >= 1 && Amax <= 64, OCM_ELIMIT, "Amax must be 1..64");
    OCM_REQUIRE(rule >= OCM_RULE_SIM && rule <= OCM_RULE_DD, OCM_EINVAL, "unknown rule");
    OCM_REQUIRE(!accept || (ct && cq && dlim), OCM_EINVAL, "accept needs ct/cq/dlim");
    if (n == 0) return 0;
    ScoreArgs a{};
    a.n = n; a.p = p; a.ldx = ldx; a.A = Amax; a.rule = rule; a.snv = snv; a.snv_eps = snv_eps;
    a.ctv = ct; a.cqv = cq; a.dlimv = dlim; a.P = P; a.W = inv_eig; a.T2 = T2; a.Q = Q; a.accept = accept;
    return dispatch_score<TIn, true>(X, mu, a, (cudaStream_t)stream);
}

// ---------------------------------------------------------------- conformity counts -----
__global__ void __launch_bounds__(256)
conformity_kernel(const uint8_t* __restrict__ accept, const uint8_t* __restrict__ is_target,
                  int64_t n, int m, unsigned long long* __restrict__ counts) {
    const int c = blockIdx.y;
    unsigned long long tp = 0, tn = 0, fp = 0, fn = 0;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        const bool acc = accept[r * m + c] != 0, tgt = is_target[r] != 0;
        tp += acc && tgt; tn += !acc && !tgt; fp += acc && !tgt; fn += !acc && tgt;
    }
    tp = warp_sum(tp); tn = warp_sum(tn); fp = warp_sum(fp); fn = warp_sum(fn);
    if ((threadIdx.x & 31) == 0) {
        if (tp) atomicAdd(counts + 4 * c + 0, tp);
        if (tn) atomicAdd(counts + 4 * c + 1, tn);
        if (fp) atomicAdd(counts + 4 * c + 2, fp);
        if (fn) atomicAdd(counts + 4 * c + 3, fn);
    }
}

}  // namespace ocm

extern "C" int ocm_simca_score_simt_f32(const float* X, int64_t n, int32_t p, int64_t ldx, const float* mu,
                                        const float* P, const float* Sinv, const ocm_score_params* hp,
                                        float* T2, float* Q, uint8_t* accept, float* scores, void* stream) {
    return ocm::score_single<float>(X, n, p, ldx, mu, P, Sinv, hp, T2, Q, accept, scores, stream);
}
extern "C" int ocm_simca_score_simt_f64in(const double* X, int64_t n, int32_t p, int64_t ldx, const double* mu,
                                          const float* P, const float* Sinv, const ocm_score_params* hp,
                                          float* T2, float* Q, uint8_t* accept, float* scores, void* stream) {
    return ocm::score_single<double>(X, n, p, ldx, mu, P, Sinv, hp, T2, Q, accept, scores, stream);
}
extern "C" int ocm_simca_score_multiA_f32(const float* X, int64_t n, int32_t p, int64_t ldx, const float* mu,
                                          const float* P, const float* inv_eig, int32_t Amax, int32_t rule,
                                          int32_t snv, float snv_eps, const float* ct, const float* cq,
                                          const float* dlim, float* T2, float* Q, uint8_t* accept, void* stream) {
    return ocm::score_multi<float>(X, n, p, ldx, mu, P, inv_eig, Amax, rule, snv, snv_eps, ct, cq, dlim, T2, Q, accept, stream);
}
extern "C" int ocm_simca_score_multiA_f64in(const double* X, int64_t n, int32_t p, int64_t ldx, const double* mu,
                                            const float* P, const float* inv_eig, int32_t Amax, int32_t rule,
                                            int32_t snv, float snv_eps, const float* ct, const float* cq,
                                            const float* dlim, float* T2, float* Q, uint8_t* accept, void* stream) {
    return ocm::score_multi<double>(X, n, p, ldx, mu, P, inv_eig, Amax, rule, snv, snv_eps, ct, cq, dlim, T2, Q, accept, stream);
}

extern "C" int ocm_conformity_counts(const uint8_t* accept, const uint8_t* is_target, int64_t n, int32_t m,
                                     int64_t* counts, void* stream) {
    using namespace ocm;
    if (n == 0) return 0;
    OCM_REQUIRE(accept && is_target && counts && n > 0 && m >= 1, OCM_EINVAL, "bad arguments");
    OCM_REQUIRE(m <= 65535, OCM_ELIMIT, "too many decision columns");
    if (n == 0) return 0;
    int64_t bx = (n + 255) / 256;
    if (bx > 1024) bx = 1024;
    dim3 grid((unsigned)bx, (unsigned)m);
    conformity_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(accept, is_target, n, m,
                                                             reinterpret_cast<unsigned long long*>(counts));
    return check_launch("conformity_kernel");
}
