// Fused SIMCA scoring, register-resident "row team" kernel for the bulk shapes
// (fp32 rows, p = NT*CPL, even A <= 12): the FP32-pipe path that carries config 3 until/unless
// the tcgen05 kernel (score_tc.cu) takes the shape.
//
// Why this mapping.  Per spectrum the work is 2*p*A FMAs (project + reconstruct; the explicit
// residual is kept for parity -- SURVEY.md H2) against 4*p bytes, i.e. the FP32 pipe
// (128 FMA/clk/SM) and HBM are in balance at A = 10, so every non-FMA instruction in the
// inner loops costs throughput.  The generic kernel pays an LDS per 4 FMAs for the loadings and
// re-reads the row; here a team of NT threads owns a row group and every thread owns the SAME
// CPL columns for the whole kernel, so
//   * its slice of the loadings P[A][CPL] lives in REGISTERS (no LDS for P in the loops),
//   * its slice of the R rows stays in registers between projection and residual (one HBM read),
//   * the next row groups stream HBM -> shared memory with cp.async two groups ahead; each thread
//     copies exactly the 16-byte slots it will read back itself, so no barrier guards the stage,
//   * all inner-loop arithmetic is packed FFMA2 (fma.rn.f32x2, new on sm_100): two FMAs per issue
//     slot, pairs = adjacent wavelengths, so the issue port is not the limiter,
//   * the only cross-thread traffic is R*A partial scores + R partial Q per group: a halving
//     (reduce-scatter) shuffle tree inside each warp, then one small shared-memory exchange.
#include "ocm_common.cuh"

namespace ocm {

constexpr int RT_R = 4;        // rows per group
constexpr int RT_STAGES = 3;   // cp.async ring depth (groups in flight)

// Reduce-scatter over the warp: on entry every lane holds C partial values; halving steps while C is
// even (each step sends the half the lane does not own), then a butterfly over the remaining lane
// bits.  On exit v[0..rs_count(C,16)) hold warp totals of value indices base+0.. (base depends on
// the lane); lanes that differ only in the bits of rs_dup_mask(C,16) hold identical results.
__host__ __device__ constexpr int rs_count(int C, int O) { return (O >= 1 && C % 2 == 0) ? rs_count(C / 2, O / 2) : C; }
__host__ __device__ constexpr int rs_dup_mask(int C, int O) {
    return O == 0 ? 0 : (C % 2 == 0 ? rs_dup_mask(C / 2, O / 2) : 2 * O - 1);
}

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

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Packed-pair helpers.  The FFMA2s are issued through `asm volatile` on purpose: volatile asm
// statements keep their program order, which pins the schedule to "R*A (resp. R*CPL/2) independent
// accumulator chains in flight" -- left to itself the compiler shortens live ranges by running the
// chains almost back to back, and with two warps per scheduler the FMA latency is then exposed
// (measured: 25 % of samples in stall_wait, FMA pipe 47 % busy).
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float2 unpack2(f32x2 v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ void ffma2_acc(f32x2& d, f32x2 a, f32x2 b) {
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}

struct RowTeamArgs {
    const float* X; int64_t n; int64_t ldx;
    const float* mu; const float* P; const float* Sinv;
    int rule; float ct, cq, dlim;
    float* T2; float* Q; uint8_t* accept;
};

template <int A, int CPL, int NT>
__global__ void __launch_bounds__(NT, NT == 128 ? 3 : 1)
score_rowteam_kernel(RowTeamArgs a) {
    constexpr int R = RT_R, NV = CPL / 4, NP = CPL / 2, NW = NT / 32, N = R * A, p = NT * CPL, S = RT_STAGES;
    constexpr int RS_CR = rs_count(N, 16), RS_DUP = rs_dup_mask(N, 16);
    constexpr int QS_CR = rs_count(R, 16), QS_DUP = rs_dup_mask(R, 16);
    extern __shared__ __align__(16) float s_stage[];          // [S][R][p]
    __shared__ __align__(16) float s_red[NW][N];
    __shared__ __align__(16) float2 s_tneg[N];                // (-t, -t): ready-made FFMA2 operand
    __shared__ float s_q[NW][R];
    __shared__ float s_S[A * A];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < A * A; i += NT) s_S[i] = a.Sinv[i];

    // this thread's columns: float4 slots tid + NT*k, i.e. pairs (2 per slot)
    f32x2 Pr[A][NP];
    float* s_mu = s_stage + (size_t)S * R * p;                 // [p]; each thread reads back its own slots
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int col = 4 * (tid + NT * k);
        *reinterpret_cast<float4*>(s_mu + col) = *reinterpret_cast<const float4*>(a.mu + col);
#pragma unroll
        for (int c = 0; c < A; ++c) {
            const float4 p4 = *reinterpret_cast<const float4*>(a.P + (int64_t)c * p + col);
            Pr[c][2 * k] = pack2(p4.x, p4.y); Pr[c][2 * k + 1] = pack2(p4.z, p4.w);
        }
    }

    const int64_t ngroups = (a.n + R - 1) / R;
    auto stage_group = [&](int64_t g, int slot) {
        if (g < ngroups) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                int64_t row = g * R + r;
                row = row < a.n ? row : a.n - 1;
                const float* xr = a.X + row * a.ldx;
                float* dst = s_stage + ((size_t)slot * R + r) * p;
#pragma unroll
                for (int k = 0; k < NV; ++k) cp_async16(dst + 4 * (tid + NT * k), xr + 4 * (tid + NT * k));
            }
        }
        cp_async_commit();     // always commit (possibly empty) so the group arithmetic stays uniform
    };

    int64_t g = blockIdx.x;
#pragma unroll
    for (int s = 0; s < S - 1; ++s) stage_group(g + (int64_t)s * gridDim.x, s);
    __syncthreads();           // s_S visible
    int slot = 0;
    for (; g < ngroups; g += gridDim.x) {
        stage_group(g + (int64_t)(S - 1) * gridDim.x, (slot + S - 1) % S);
        cp_async_wait<S - 1>();                                   // this thread's copies of group g landed

        f32x2 x[R][NP];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float* src = s_stage + ((size_t)slot * R + r) * p;
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                const float4 v = *reinterpret_cast<const float4*>(src + 4 * (tid + NT * k));
                const float4 m = *reinterpret_cast<const float4*>(s_mu + 4 * (tid + NT * k));
                x[r][2 * k] = pack2(v.x - m.x, v.y - m.y);
                x[r][2 * k + 1] = pack2(v.z - m.z, v.w - m.w);
            }
        }

        // ---- projection partials: R*A independent chains, pairs = (even, odd) wavelength --------
        float acc[N];
#pragma unroll
        for (int h = 0; h < R; h += 2) {                      // two rows at a time: 2*A live pair accumulators
            f32x2 acc2[2][A];
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int c = 0; c < A; ++c) acc2[r][c] = 0ull;
#pragma unroll
            for (int j = 0; j < NP; ++j)
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int c = 0; c < A; ++c) ffma2_acc(acc2[r][c], x[h + r][j], Pr[c][j]);
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int c = 0; c < A; ++c) {
                    const float2 v = unpack2(acc2[r][c]);
                    acc[(h + r) * A + c] = v.x + v.y;
                }
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
            // transposed to [c][r] so the residual loop reads two rows' operands per LDS.128
            s_tneg[(tid % A) * R + tid / A] = make_float2(-s, -s);
        }
        __syncthreads();

        // ---- explicit residual r = xc - P^T t (R*CPL/2 independent chains), Q = |r|^2 -----------
#pragma unroll
        for (int c = 0; c < A; ++c) {
            f32x2 tn[R];
#pragma unroll
            for (int r = 0; r < R; r += 2) {
                const ulonglong2 tt = *reinterpret_cast<const ulonglong2*>(&s_tneg[c * R + r]);
                tn[r] = tt.x; tn[r + 1] = tt.y;
            }
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int j = 0; j < NP; ++j) ffma2_acc(x[r][j], tn[r], Pr[c][j]);
        }
        float q[R];
        {
            f32x2 s2[R];
#pragma unroll
            for (int r = 0; r < R; ++r) s2[r] = 0ull;
#pragma unroll
            for (int j = 0; j < NP; ++j)
#pragma unroll
                for (int r = 0; r < R; ++r) ffma2_acc(s2[r], x[r][j], x[r][j]);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float2 v = unpack2(s2[r]);
                q[r] = v.x + v.y;
            }
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
                    for (int k = 0; k < A; ++k) s = fmaf(s_S[c * A + k], s_tneg[k * R + tid].x, s);
                    t2 = fmaf(s, s_tneg[c * R + tid].x, t2);      // signs cancel in the quadratic form
                }
                if (a.T2) a.T2[row] = t2;
                if (a.Q) a.Q[row] = qq;
                if (a.accept) a.accept[row] = rule_distance(a.rule, t2 * a.ct, qq * a.cq) < a.dlim ? 1 : 0;
            }
        }
        // s_red / s_tneg / s_q are rewritten only after the NEXT iteration's barriers: s_red after
        // everyone passed barriers 2+3 of this one, s_tneg after barrier 1 of the next (the finalising
        // threads reach it only after their reads above), s_q after barrier 2 of the next.
        slot = (slot + 1) % S;
    }
    cp_async_wait<0>();
}

template <int A, int CPL, int NT>
int launch_rowteam(const RowTeamArgs& a, cudaStream_t st) {
    auto kern = score_rowteam_kernel<A, CPL, NT>;
    const size_t smem = sizeof(float) * (RT_STAGES * RT_R + 1) * (size_t)NT * CPL;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem);
    if (occ < 1) occ = 1;
    const int64_t ngroups = (a.n + RT_R - 1) / RT_R;
    int64_t blocks = (int64_t)sm_count() * occ;
    if (blocks > ngroups) blocks = ngroups;
    kern<<<(unsigned)blocks, NT, smem, st>>>(a);
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
