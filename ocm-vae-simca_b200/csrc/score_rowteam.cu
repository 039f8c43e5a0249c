// Fused SIMCA scoring, register-resident "row team" kernel for the bulk shapes
// (fp32 rows, p = NT*CPL, A <= 12): the FP32-pipe path that carries config 3 until/unless
// the tcgen05 kernel (score_tc.cu) takes the shape.
//
// Why this mapping.  Per spectrum the work is 2*p*A FMAs (project + reconstruct, the explicit
// residual is kept for parity -- SURVEY.md H2) against 4*p bytes, i.e. the FP32 pipe
// (128 FMA/clk/SM) and HBM are in balance at A = 10, so every non-FMA instruction in the
// inner loops costs throughput.  The generic kernel pays an LDS per 4 FMAs for the loadings and
// re-reads the row; here a team of NT threads owns a row group and every thread owns the SAME
// CPL columns for the whole kernel, so
//   * its slice of the loadings P[A][CPL] and of the mean lives in REGISTERS (no LDS in the loops),
//   * its slice of the R rows stays in registers between projection and residual (one HBM read),
//   * the next row group is prefetched into registers a full iteration ahead (8 x 16 B per lane),
//   * the only cross-thread traffic is R*A partial scores + R partial Q per group: a halving
//     (reduce-scatter) shuffle tree inside each warp, then one small shared-memory exchange.
// Inner loops are pure FFMA: 2*R*A*CPL per lane per group.
#include "ocm_common.cuh"

namespace ocm {

constexpr int RT_R = 4;   // rows per group

// Reduce-scatter over the warp: on entry every lane holds C partial values; halving steps while C is
// even (each step sends the half the lane does not own), then a butterfly over the remaining lane
// bits.  On exit v[0..rs_count(C,16)) hold warp totals of value indices base+0.. (base depends on
// the lane); lanes that differ only in the bits of rs_dup_mask(C,16) hold identical results.
__host__ __device__ constexpr int rs_count(int C, int O) { return (O >= 1 && C % 2 == 0) ? rs_count(C / 2, O / 2) : C; }
__host__ __device__ constexpr int rs_dup_mask(int C, int O) { return O == 0 ? 0 : (C % 2 == 0 ? rs_dup_mask(C / 2, O / 2) : 2 * O - 1); }

template <int C, int O>
__device__ __forceinline__ void warp_reduce_scatter(float* v, int lane, int& base) {
    if constexpr (O == 0) {
        return;
    } else if constexpr (C % 2 == 0) {
        const bool upper = (lane & O) != 0;
#pragma unroll
        for (int i = 0; i < C / 2; ++i) {
            const float send = upper ? v[i] : v[i + C / 2];
            const float keep = upper ? v[i + C / 2] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, O);
        }
        base += upper ? C / 2 : 0;
        warp_reduce_scatter<C / 2, O / 2>(v, lane, base);
    } else {
#pragma unroll
        for (int o = O; o >= 1; o >>= 1)
#pragma unroll
            for (int i = 0; i < C; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
    }
}

struct RowTeamArgs {
    const float* X; int64_t n; int64_t ldx;
    const float* mu; const float* P; const float* Sinv;
    int rule; float ct, cq, dlim;
    float* T2; float* Q; uint8_t* accept;
};

template <int A, int CPL, int NT>
__global__ void __launch_bounds__(NT, NT == 128 ? 2 : 1)
score_rowteam_kernel(RowTeamArgs a) {
    constexpr int R = RT_R, NV = CPL / 4, NW = NT / 32, N = R * A, p = NT * CPL;
    constexpr int RS_CR = rs_count(N, 16), RS_DUP = rs_dup_mask(N, 16);
    constexpr int QS_CR = rs_count(R, 16), QS_DUP = rs_dup_mask(R, 16);
    __shared__ __align__(16) float s_red[NW][N];
    __shared__ __align__(16) float s_t[N];
    __shared__ float s_q[NW][R];
    __shared__ float s_S[A * A];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < A * A; i += NT) s_S[i] = a.Sinv[i];

    // this thread's columns: float4 slots tid + NT*k
    float Pr[A][CPL], mur[CPL];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int col = 4 * (tid + NT * k);
        const float4 m4 = *reinterpret_cast<const float4*>(a.mu + col);
        mur[4 * k + 0] = m4.x; mur[4 * k + 1] = m4.y; mur[4 * k + 2] = m4.z; mur[4 * k + 3] = m4.w;
#pragma unroll
        for (int c = 0; c < A; ++c) {
            const float4 p4 = *reinterpret_cast<const float4*>(a.P + (int64_t)c * p + col);
            Pr[c][4 * k + 0] = p4.x; Pr[c][4 * k + 1] = p4.y; Pr[c][4 * k + 2] = p4.z; Pr[c][4 * k + 3] = p4.w;
        }
    }

    const int64_t ngroups = (a.n + R - 1) / R;
    float xn[R][CPL];
    auto load_group = [&](int64_t g) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            int64_t row = g * R + r;
            row = row < a.n ? row : a.n - 1;
            const float* xr = a.X + row * a.ldx;
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                const float4 v = ldg_stream4(xr + 4 * (tid + NT * k));
                xn[r][4 * k + 0] = v.x; xn[r][4 * k + 1] = v.y; xn[r][4 * k + 2] = v.z; xn[r][4 * k + 3] = v.w;
            }
        }
    };

    int64_t g = blockIdx.x;
    if (g < ngroups) load_group(g);
    __syncthreads();
    for (; g < ngroups; g += gridDim.x) {
        float x[R][CPL];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int j = 0; j < CPL; ++j) x[r][j] = xn[r][j] - mur[j];
        if (g + gridDim.x < ngroups) load_group(g + gridDim.x);     // prefetch a full iteration ahead

        // ---- projection partials ----------------------------------------------------------
        float acc[N];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int c = 0; c < A; ++c) {
                float s = 0.f;
#pragma unroll
                for (int j = 0; j < CPL; ++j) s = fmaf(x[r][j], Pr[c][j], s);
                acc[r * A + c] = s;
            }
        int base = 0;
        warp_reduce_scatter<N, 16>(acc, lane, base);
        if ((lane & RS_DUP) == 0) {
#pragma unroll
            for (int i = 0; i < RS_CR; ++i) s_red[warp][base + i] = acc[i];
        }
        __syncthreads();
        if (tid < N) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < NW; ++w) s += s_red[w][tid];
            s_t[tid] = s;
        }
        __syncthreads();

        // ---- explicit residual ------------------------------------------------------------
        float q[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float t[A];
#pragma unroll
            for (int c4 = 0; c4 < (A + 3) / 4; ++c4) {
                // A is even; rows start at multiples of A floats -> 8-byte aligned pairs
                if constexpr (A % 4 == 0) {
                    const float4 v = *reinterpret_cast<const float4*>(&s_t[r * A + 4 * c4]);
                    t[4 * c4] = v.x; t[4 * c4 + 1] = v.y; t[4 * c4 + 2] = v.z; t[4 * c4 + 3] = v.w;
                } else {
#pragma unroll
                    for (int e = 0; e < 4; e += 2)
                        if (4 * c4 + e < A) {
                            const float2 v = *reinterpret_cast<const float2*>(&s_t[r * A + 4 * c4 + e]);
                            t[4 * c4 + e] = v.x; t[4 * c4 + e + 1] = v.y;
                        }
                }
            }
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                float rr = x[r][j];
#pragma unroll
                for (int c = 0; c < A; ++c) rr = fmaf(-t[c], Pr[c][j], rr);
                s = fmaf(rr, rr, s);
            }
            q[r] = s;
        }
        int qbase = 0;
        warp_reduce_scatter<R, 16>(q, lane, qbase);
        if ((lane & QS_DUP) == 0) {
#pragma unroll
            for (int i = 0; i < QS_CR; ++i) s_q[warp][qbase + i] = q[i];
        }
        __syncthreads();

        // ---- distances + decision: one thread per row ---------------------------------------
        if (tid < R) {
            const int64_t row = g * R + tid;
            if (row < a.n) {
                float qq = 0.f;
#pragma unroll
                for (int w = 0; w < NW; ++w) qq += s_q[w][tid];
                float t2 = 0.f;
#pragma unroll
                for (int c = 0; c < A; ++c) {
                    float s = 0.f;
#pragma unroll
                    for (int k = 0; k < A; ++k) s = fmaf(s_S[c * A + k], s_t[tid * A + k], s);
                    t2 = fmaf(s, s_t[tid * A + c], t2);
                }
                if (a.T2) a.T2[row] = t2;
                if (a.Q) a.Q[row] = qq;
                if (a.accept) a.accept[row] = rule_distance(a.rule, t2 * a.ct, qq * a.cq) < a.dlim ? 1 : 0;
            }
        }
        // s_red / s_t / s_q are rewritten only after the next iteration's barriers (see header)
    }
}

template <int A, int CPL, int NT>
int launch_rowteam(const RowTeamArgs& a, cudaStream_t st) {
    auto kern = score_rowteam_kernel<A, CPL, NT>;
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, 0);
    if (occ < 1) occ = 1;
    const int64_t ngroups = (a.n + RT_R - 1) / RT_R;
    int64_t blocks = (int64_t)sm_count() * occ;
    if (blocks > ngroups) blocks = ngroups;
    kern<<<(unsigned)blocks, NT, 0, st>>>(a);
    return check_launch("score_rowteam_kernel");
}

template <int CPL, int NT>
int dispatch_rowteam(int A, const RowTeamArgs& a, cudaStream_t st) {
    switch (A) {
        case 2: return launch_rowteam<2, CPL, NT>(a, st);
        case 4: return launch_rowteam<4, CPL, NT>(a, st);
        case 6: return launch_rowteam<6, CPL, NT>(a, st);
        case 8: return launch_rowteam<8, CPL, NT>(a, st);
        case 10: return launch_rowteam<10, CPL, NT>(a, st);
        case 12: return launch_rowteam<12, CPL, NT>(a, st);
    }
    set_error("score_rowteam: unsupported A=%d", A);
    return OCM_ELIMIT;
}

}  // namespace ocm

// shape gate: fp32 rows of 512 / 1024 / 2048 wavelengths, 16-byte aligned, EVEN n_components <= 12, no fused SNV
extern "C" int ocm_simca_score_rowteam_supported(int64_t n, int32_t p, int64_t ldx, int32_t A, int32_t snv,
                                                 const void* X) {
    if (n < 1 || snv) return 0;
    if (!(p == 512 || p == 1024 || p == 2048)) return 0;
    if (A < 2 || A > 12 || (A & 1)) return 0;
    if (ldx % 4 != 0 || (reinterpret_cast<uintptr_t>(X) & 15) != 0) return 0;
    return 1;
}

extern "C" int ocm_simca_score_rowteam_f32(const float* X, int64_t n, int32_t p, int64_t ldx, const float* mu,
                                           const float* P, const float* Sinv, const ocm_score_params* hp,
                                           float* T2, float* Q, uint8_t* accept, void* stream) {
    using namespace ocm;
    if (n == 0) return 0;
    OCM_REQUIRE(X && mu && P && Sinv && hp, OCM_EINVAL, "NULL buffer");
    OCM_REQUIRE(ocm_simca_score_rowteam_supported(n, p, ldx, hp->A, hp->snv, X), OCM_ELIMIT,
                "shape not supported by the row-team kernel");
    OCM_REQUIRE((reinterpret_cast<uintptr_t>(mu) & 15) == 0 && (reinterpret_cast<uintptr_t>(P) & 15) == 0,
                OCM_EALIGN, "mu / P must be 16-byte aligned");
    RowTeamArgs a{X, n, ldx, mu, P, Sinv, hp->rule, hp->ct, hp->cq, hp->dlim, T2, Q, accept};
    cudaStream_t st = (cudaStream_t)stream;
    if (p == 512) return dispatch_rowteam<4, 128>(hp->A, a, st);
    if (p == 1024) return dispatch_rowteam<8, 128>(hp->A, a, st);
    return dispatch_rowteam<8, 256>(hp->A, a, st);
}
