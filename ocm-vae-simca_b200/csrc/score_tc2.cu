// Fused SIMCA scoring on tcgen05, second generation: the two passes of a tile run CONCURRENTLY in specialised warps,
// and a tile's wavelengths are split over the CTAs of a thread-block cluster.
//
// What the first kernel (score_tc.cu) taught us (ncu, profiles/r01_score_tc_v11.md): per tile it alternates
// a pass that streams the tile from HBM (~11 us per 128 x 1024 tile: the HBM floor itself when all SMs are in it) with
// a pass that re-reads the tile from L2 and forms the residual (~5 us, HBM idle).  The passes load different
// resources, so the way up is to overlap them -- but interleaving two 512 KB tiles per SM doubled the L2 working set to
// 152 MB (> 126 MB), and the second pass started missing (measured: DRAM traffic +50 %).
//
// Here a cluster of C CTAs owns one 128-spectrum tile and CTA r owns the wavelengths [r W, (r+1) W), W <= 512:
//   * per SM a tile is 128 x 512 floats = 256 KB; the HBM stream is held to LEAD stages ahead of pass 2 (default 1.5
//     tiles), so that what waits in L2 for its second read stays near 60 MB over 148 SMs (C = 2, p = 1024);
//   * the loadings slice is 48 KB instead of 96 KB, which pays for a second TMA ring;
//   * the partial scores of the C slices are exchanged through distributed shared memory (8 KB per CTA and tile,
//     st.shared::cluster + remote mbarrier arrive), summed in rank order by every CTA (identical bits everywhere), and
//     the partial Q of the slices is sent to rank 0, which writes T2, Q and the decision.
// Roles of one CTA (24 warps):
//   TMA1      streams the slice of the tiles ahead from HBM (L2 evict_last) into ring 1 (3 stages of 128 x 64 floats)
//   C1  x8    pass 1: LDS -> centre -> exact 3-way bf16 split -> tcgen05.st into an A-operand buffer (thread = spectrum x 32 columns)
//   MMA-P     projection: T_partial[128 x 16] += A * P^T, 3 MMAs per k-step against the stacked terms of P (see score_tc.cu)
//   CT  x4    per tile: partial scores back (tcgen05.ld), cluster exchange, T2, split scores -> TMEM (thread = spectrum)
//   MMA-R     reconstruction D[128 x 64] = T * P[:, chunk], six MMAs per 64 columns, two TMEM buffers
//   TMA2      re-streams the slice from L2 (evict_first) into ring 2 (3 boxes of 128 x 32 floats)
//   C2  x8    pass 2: LDS -> centre -> minus the reconstruction read back from TMEM -> Q (thread = spectrum x 16 columns of
//             every box)
// All hand-offs are mbarriers with bounded waits (trap on expiry, see score_tc_common.cuh).  Each barrier has waiters
// whose own progress holds its producers to one phase ahead (noted where it is not the classic ring argument), so a
// parity wait can never alias.
// Numerics are those of score_tc.cu: exact 3-term bf16 splits of both operands, six products, fp32 accumulation in TMEM
// with the large products rotating over accumulators (<= 16 chained accumulations), explicit residual.
#include "ocm_common.cuh"
#include "score_tc_common.cuh"

#include <cuda.h>
#include <stdlib.h>

namespace ocm {
namespace tc2 {
using namespace tcs;

constexpr int NCOMP = 16;
constexpr int TILE_M = 128;
constexpr int STAGE_COLS = 64;            // pass-1 stage = one 128-byte bf16 swizzle atom of P = two TMA boxes
constexpr int BOX_COLS = 32;              // TMA box: 128 rows x 32 floats (128 B rows, SWIZZLE_128B)
constexpr int NS_MAX = 8;                 // stages per CTA slice (512 wavelengths)
constexpr int R1 = 3;                     // ring 1: stages
constexpr int R2MAX = 4;                  // ring 2: boxes (64 KB in flight: three boxes left pass 2 waiting for L2 46 % of its time);
                                          // three for clusters of four, whose exchange buffers take the room
constexpr int NC1 = 256, NC2 = 256, NCT = 128;       // pass-1 converters, pass-2 converters (two groups), score threads
constexpr int W_C2 = NC1 / 32, W_CT = W_C2 + NC2 / 32, W_TMA1 = W_CT + NCT / 32, W_TMA2 = W_TMA1 + 1, W_MMAP = W_TMA1 + 2,
              W_MMAR = W_TMA1 + 3;
constexpr int NTHREADS = 32 * (W_MMAR + 1);
static_assert(NTHREADS == 768 && W_C2 % 4 == 0 && W_CT % 4 == 0, "converter warps must start at a TMEM lane quadrant boundary");

// TMEM columns
constexpr uint32_t TM_A = 0;              // 2 A-operand buffers x (3 terms x 32 columns)
constexpr uint32_t TM_R = 192;            // 2 reconstruction buffers x 64 columns
constexpr uint32_t TM_D1 = 320;           // A term 1 x [P1;P2;P3] (N = 48), rotating over 2 accumulators
constexpr uint32_t TM_D2 = 416;           // A term 2 x [P1;P2]    (N = 32)
constexpr uint32_t TM_D3 = TM_D2;         // A term 3 x [P1]       (N = 16): accumulates into the (2,1) columns, it is their sum that is needed
constexpr uint32_t TM_TS = 448;           // split scores, one set per tile parity (24 columns each): term i at +8 i
static_assert(TM_TS + 48 <= 512, "TMEM column budget");

constexpr size_t SM_PATOM = 3 * NCOMP * 128;                   // 6144: 64 columns x (3 terms x 16 components)
constexpr size_t SM_BOX = (size_t)TILE_M * BOX_COLS * 4;       // 16384
constexpr size_t SM_STAGE = 2 * SM_BOX;
constexpr size_t XSLOT = (size_t)TILE_M * NCOMP * 4;            // 8192: partial scores of one peer, [4 comps][row] float4
constexpr size_t SM_OFF_P = 0;
constexpr size_t SM_OFF_R1 = NS_MAX * SM_PATOM;                 // 49152 (1024-aligned)
constexpr size_t SM_OFF_MU = SM_OFF_R1 + R1 * SM_STAGE;
constexpr size_t SM_OFF_S = SM_OFF_MU + NS_MAX * STAGE_COLS * 4;
constexpr size_t SM_OFF_T2 = SM_OFF_S + NCOMP * NCOMP * 4;      // [2 parities][128] T2 of the tile, CT -> C2
constexpr size_t SM_OFF_Q = SM_OFF_T2 + 2 * TILE_M * 4;         // [2 parities][128] partial Q of C2 group 1
constexpr size_t SM_OFF_BAR = SM_OFF_Q + 2 * TILE_M * 4;
constexpr size_t SM_OFF_X = SM_OFF_BAR + 320;                   // [C-1 peers] XSLOT (one tile at a time, credit: XFREE), then [2][C-1][128] floats of Q
template <int C> __host__ __device__ constexpr int r2_slots() { return C >= 4 ? 3 : R2MAX; }
template <int C> __host__ __device__ constexpr size_t off_r2() { return (SM_OFF_X + (size_t)(C - 1) * (XSLOT + 2 * TILE_M * 4) + 1023) / 1024 * 1024; }   // ring 2 (1024-byte aligned boxes)
template <int C> __host__ __device__ constexpr size_t sm_total() { return off_r2<C>() + (size_t)r2_slots<C>() * SM_BOX; }
static_assert(SM_OFF_R1 % 1024 == 0 && sm_total<1>() <= 232448 && sm_total<2>() <= 232448 && sm_total<4>() <= 232448, "shared memory budget");

enum { B_FULL1 = 0, B_EMPTY1 = B_FULL1 + R1, B_FULL2 = B_EMPTY1 + R1, B_EMPTY2 = B_FULL2 + R2MAX, B_AFULL = B_EMPTY2 + R2MAX,
       B_AFREE = B_AFULL + 2, B_TREADY = B_AFREE + 2, B_TFREE, B_THFULL, B_C2DONE = B_THFULL + 2, B_DREADY = B_C2DONE + 2,
       B_DFREE = B_DREADY + 2, B_XFULL = B_DFREE + 2, B_XFREE, B_QFULL, B_T2RDY = B_QFULL + 2, B_COUNT = B_T2RDY + 2 };
// words after the barriers: TMEM base, pass-2 progress (stages consumed by C2 group 0), tiles whose pass 1 has begun
static_assert(8 * B_COUNT + 16 <= 320, "barrier block");

struct Args {
    int64_t n;
    int p;            // wavelengths (multiple of 4)
    int ns;           // stages of 64 columns per CTA slice (<= NS_MAX); CTA r owns columns [64 ns r, 64 ns (r + 1))
    const float* mu; const float* P; const float* Sinv;
    int A; int rule; float ct, cq, dlim;
    float* T2; float* Q; uint8_t* accept;
    unsigned int* status;
    uint32_t max_polls;
    int inject;       // test hook: the HBM producer of CTA 0 drops one stage of its first tile
    int lead;         // stages the HBM stream may run ahead of pass 2 (> ns)
    int pol;          // L2 policy of the two rings, see make_policy: bits 0-1 pass 1, bits 2-3 pass 2
    uint32_t zero;    // always 0: makes the ring-slot releases depend on the loaded data (mbar_arrive_at_after)
};

// ---- cluster-scope PTX ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_id() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_count() { uint32_t r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_rank(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float a) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(a) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote_after(uint32_t cluster_addr, uint32_t loaded_bits, uint32_t zero) {
    const uint32_t cnt = 1u + (loaded_bits & zero);          // see mbar_arrive_at_after
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_addr), "r"(cnt) : "memory");
}
// bounded wait that records WHICH hand-off expired (status word = 100 + barrier slot, 200 + counter) before it traps
static __device__ __noinline__ void wait_timed_out_code(unsigned int* status, uint32_t code) {
    if (status) { atomicExch(status, code); __threadfence_system(); }
    __trap();
}
__device__ __forceinline__ void mbar_wait_c(uint32_t bar, uint32_t parity, const WaitCtx& w, uint32_t bar0) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > w.max_polls) wait_timed_out_code(w.status, 100u + ((bar - bar0) >> 3));
    }
}
// the same for roles that are off the critical path (they wait most of the time): sleep between polls, so that their
// polling does not take issue slots -- and power -- from the pass-1 converters (ncu: 38 % of all warp instructions were polls)
template <int SLEEP_NS>
__device__ __forceinline__ void mbar_wait_s(uint32_t bar, uint32_t parity, const WaitCtx& w, uint32_t bar0) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(SLEEP_NS);
        if (++spins > w.max_polls) wait_timed_out_code(w.status, 100u + ((bar - bar0) >> 3));
    }
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity, const WaitCtx& w) {
    uint32_t spins = 0;
    for (;;) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
        if (++spins > w.max_polls) wait_timed_out_code(w.status, 190u);
    }
}
__device__ __forceinline__ uint64_t make_policy(int kind) {     // 0 evict_last, 1 evict_normal, 2 evict_first, 3 evict_unchanged
    uint64_t pol;
    if (kind == 0) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    else if (kind == 1) asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    else if (kind == 2) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    else asm volatile("createpolicy.fractional.L2::evict_unchanged.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr) : "memory");
    return r;
}
__device__ __forceinline__ float lds_f1(uint32_t addr) {
    float r;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"(addr) : "memory");
    return r;
}
__device__ __forceinline__ void sts_f1(uint32_t addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u32_volatile(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u32_volatile(uint32_t addr, uint32_t v) {
    asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
// a single producer thread publishes a monotonic counter in shared memory, a single consumer thread waits for it
__device__ __forceinline__ void wait_counter(uint32_t addr, uint32_t at_least, const WaitCtx& w) {
    uint32_t spins = 0;
    while ((int32_t)(lds_u32_volatile(addr) - at_least) < 0) {
        __nanosleep(64);
        if (++spins > w.max_polls) wait_timed_out_code(w.status, 200u + (addr & 15u));
    }
}

// bring-up trace (a.pol bit 8): the roles of CTA 0 leave their position in words 1.. of the mapped status block
#ifdef OCM_TC2_TRACE
#define TC2_TRACE(role, val)                                                                                   \
    do {                                                                                                       \
        if ((a.pol & 256) && (blockIdx.x == 0 || (a.pol & 512)) && a.status) *reinterpret_cast<volatile unsigned int*>(a.status + (role)) = (unsigned int)(val); \
    } while (0)
#else
#define TC2_TRACE(role, val) do { } while (0)
#endif

template <int C>
__global__ void __launch_bounds__(NTHREADS, 1)
score_tc2_kernel(const __grid_constant__ CUtensorMap tmap, Args a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    float* s_mu = reinterpret_cast<float*>(smem + SM_OFF_MU);
    float* s_S = reinterpret_cast<float*>(smem + SM_OFF_S);
    uint32_t* s_words = reinterpret_cast<uint32_t*>(smem + SM_OFF_BAR + 8 * B_COUNT);      // [0] TMEM base, [1] progress, [2] begun
    const uint32_t sb = smem_u32(smem);
    const uint32_t bar0 = sb + (uint32_t)SM_OFF_BAR;
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    const uint32_t w_prog = bar0 + 8u * (uint32_t)B_COUNT + 4u, w_begun = w_prog + 4u;
    const WaitCtx wctx{a.status, a.max_polls};

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = C > 1 ? cluster_rank() : 0u;
    const int64_t ntiles = (a.n + TILE_M - 1) / TILE_M;
    const int64_t cid = C > 1 ? (int64_t)cluster_id() : (int64_t)blockIdx.x;
    const int64_t ncl = C > 1 ? (int64_t)cluster_count() : (int64_t)gridDim.x;
    const int64_t my_tiles = cid < ntiles ? (ntiles - cid + ncl - 1) / ncl : 0;      // the same for every CTA of a cluster
    const int ns = a.ns;
    const int col0 = (int)rank * ns * STAGE_COLS;               // first wavelength of my slice

    // ---- one-time setup ---------------------------------------------------------------------
    if (tid == 0) {
        for (int s = 0; s < R1; ++s) { mbar_init(BAR(B_FULL1 + s), 1); mbar_init(BAR(B_EMPTY1 + s), NC1); }
        for (int s = 0; s < R2MAX; ++s) { mbar_init(BAR(B_FULL2 + s), 1); mbar_init(BAR(B_EMPTY2 + s), NC2); }
        for (int b = 0; b < 2; ++b) {
            mbar_init(BAR(B_AFULL + b), NC1); mbar_init(BAR(B_AFREE + b), 1);
            mbar_init(BAR(B_DREADY + b), 1); mbar_init(BAR(B_DFREE + b), NC2);
            mbar_init(BAR(B_THFULL + b), NCT); mbar_init(BAR(B_C2DONE + b), NC2); mbar_init(BAR(B_T2RDY + b), NCT);
            mbar_init(BAR(B_QFULL + b), C > 1 ? (C - 1) * (NC2 / 2) : 1);
        }
        mbar_init(BAR(B_TREADY), 1); mbar_init(BAR(B_TFREE), NCT);
        mbar_init(BAR(B_XFULL), C > 1 ? (C - 1) * NCT : 1); mbar_init(BAR(B_XFREE), C > 1 ? (C - 1) * NCT : 1);
        s_words[1] = 0u; s_words[2] = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == W_MMAP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_words)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // my slice of the loadings -> smem, three bf16 terms stacked along the rows of the 128B-swizzled atom (layout and
    // its double use as K-major / MN-major operand: score_tc.cu)
    const int nchunk16 = ns * 8;
    for (int idx = tid; idx < NCOMP * nchunk16; idx += NTHREADS) {
        const int c = idx / nchunk16, j8 = idx % nchunk16;
        const int gcol = col0 + 8 * j8;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = 0.f;
        if (c < a.A) {
            if (gcol + 4 <= a.p) {
                const float4 lo = *reinterpret_cast<const float4*>(a.P + (size_t)c * a.p + gcol);
                v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w;
            }
            if (gcol + 8 <= a.p) {
                const float4 hi = *reinterpret_cast<const float4*>(a.P + (size_t)c * a.p + gcol + 4);
                v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
            }
        }
        uint32_t w1[4], w2[4], w3[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) split3_pack(v[2 * e], v[2 * e + 1], w1[e], w2[e], w3[e]);
        const size_t off = (size_t)(j8 >> 3) * SM_PATOM + (size_t)c * 128 + (size_t)((j8 & 7) ^ (c & 7)) * 16;
        *reinterpret_cast<uint4*>(smem + SM_OFF_P + off) = make_uint4(w1[0], w1[1], w1[2], w1[3]);
        *reinterpret_cast<uint4*>(smem + SM_OFF_P + off + 2048) = make_uint4(w2[0], w2[1], w2[2], w2[3]);
        *reinterpret_cast<uint4*>(smem + SM_OFF_P + off + 4096) = make_uint4(w3[0], w3[1], w3[2], w3[3]);
    }
    for (int i = tid; i < ns * STAGE_COLS; i += NTHREADS) s_mu[i] = (col0 + i) < a.p ? a.mu[col0 + i] : 0.f;
    for (int i = tid; i < NCOMP * NCOMP; i += NTHREADS) {
        const int r = i / NCOMP, c = i % NCOMP;
        s_S[i] = (r < a.A && c < a.A) ? a.Sinv[r * a.A + c] : 0.f;
    }
    int offdiag = 0;
    for (int i = tid; i < a.A * a.A; i += NTHREADS) offdiag |= (i / a.A != i % a.A && a.Sinv[i] != 0.f) ? 1 : 0;
    const bool s_diag = __syncthreads_or(offdiag) == 0;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (s_words[0] != 0u) {         // the CTA is alone on its SM and owns all 512 columns: base 0, addresses are immediates
        if (a.status) { atomicExch(a.status, 2u); __threadfence_system(); }
        __trap();
    }
    constexpr uint32_t tmem = 0u;
    if (C > 1) cluster_sync();      // every CTA's barriers exist before the first remote arrive

    if (warp < W_C2) {
        // =====================================================================================
        // C1: pass-1 converters.  thread = (spectrum row, box g of the stage), 16 columns at a time
        // =====================================================================================
        const int row = 32 * (warp & 3) + lane;
        const int g = warp >> 2;
        const uint32_t lane_addr = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
        const uint32_t ring0 = sb + (uint32_t)SM_OFF_R1 + (uint32_t)g * (uint32_t)SM_BOX + (uint32_t)row * 128u;
        const uint32_t rsw = (uint32_t)(row & 7);
        const uint32_t mu0 = sb + (uint32_t)SM_OFF_MU + 4u * (uint32_t)(BOX_COLS * g);
        const uint32_t bar_full0 = BAR(B_FULL1), bar_full_end = BAR(B_FULL1 + R1);
        uint32_t st = ring0, bar_full = bar_full0, ring_phase = 0;
        uint32_t ab = 0, aph = 0;               // A buffer of the next stage, parity of its current use
        const uint32_t abase0 = lane_addr + TM_A + (uint32_t)(16 * g);

        for (int64_t k = 0; k < my_tiles; ++k) {
            uint32_t mu_addr = mu0;
#pragma unroll 1
            for (int s = 0; s < ns; ++s) {
                f32x2 xa[8], xb[8];
                mbar_wait_c(bar_full, ring_phase, wctx, bar0);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const ulonglong2 v = lds128(st + (((uint32_t)i ^ rsw) << 4));
                    const ulonglong2 m = lds128(mu_addr + 16u * (uint32_t)i);
                    xa[2 * i] = sub2(v.x, m.x);
                    xa[2 * i + 1] = sub2(v.y, m.y);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const ulonglong2 v = lds128(st + (((uint32_t)(4 + i) ^ rsw) << 4));
                    const ulonglong2 m = lds128(mu_addr + 64u + 16u * (uint32_t)i);
                    xb[2 * i] = sub2(v.x, m.x);
                    xb[2 * i + 1] = sub2(v.y, m.y);
                }
                {
                    uint32_t l0, h0, l1, h1;
                    upk2(xa[7], l0, h0); upk2(xb[7], l1, h1);          // the last chunk of each half (loads return in order)
                    mbar_arrive_at_after<8 * (B_EMPTY1 - B_FULL1)>(bar_full, h0 | h1, a.zero);
                }
                st += (uint32_t)SM_STAGE; bar_full += 8u;
                if (bar_full == bar_full_end) { st = ring0; bar_full = bar_full0; ring_phase ^= 1u; }
                mu_addr += 4u * STAGE_COLS;

                mbar_wait_c(BAR(B_AFREE) + 8u * ab, aph ^ 1u, wctx, bar0);          // MMAs of the buffer's previous stage retired
                tc_fence_after();
                const uint32_t abase = abase0 + ab * 96u;
                {
                    uint32_t w1[8], w2[8], w3[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) split3_pack2(xa[i], w1[i], w2[i], w3[i]);
                    tc_st8(abase, w1); tc_st8(abase + 32, w2); tc_st8(abase + 64, w3);
                }
                {
                    uint32_t w1[8], w2[8], w3[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) split3_pack2(xb[i], w1[i], w2[i], w3[i]);
                    tc_st8(abase + 8, w1); tc_st8(abase + 40, w2); tc_st8(abase + 72, w3);
                }
                tc_wait_st();
                tc_fence_before();
                mbar_arrive(BAR(B_AFULL) + 8u * ab);
                if (tid == 0) TC2_TRACE(1, (k << 8) | (s + 1));
                aph ^= ab; ab ^= 1u;                                        // the parity flips after buffer 1
            }
        }
    } else if (warp < W_CT) {
        // =====================================================================================
        // C2: pass 2.  thread = (spectrum row, column half h of every box).  Both groups of four warps read EVERY box:
        // a ring slot whose successive uses were consumed by alternating groups let a group that never waited for use
        // u - 1 mistake "use u - 2 complete" for "use u complete" (same parity) -- found as whole warps of wrong Q.
        // =====================================================================================
        const int row = 32 * (warp & 3) + lane;
        const uint32_t h = (uint32_t)(warp - W_C2) >> 2;
        const uint32_t lane_addr = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
        const uint32_t ring0 = sb + (uint32_t)off_r2<C>() + (uint32_t)row * 128u;
        uint32_t xo[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) xo[i] = (uint32_t)(((4 * (int)h + i) ^ (row & 7)) << 4);
        const uint32_t mu0 = sb + (uint32_t)SM_OFF_MU + 4u * 16u * h;
        const uint32_t bar_full0 = BAR(B_FULL2), bar_full_end = BAR(B_FULL2 + r2_slots<C>());
        uint32_t st = ring0, bar_full = bar_full0, ring_phase = 0;
        uint32_t db = 0, dph = 0;               // reconstruction buffer of the next chunk, parity of its current use
        const uint32_t qbuf0 = sb + (uint32_t)SM_OFF_X + (uint32_t)(C - 1) * (uint32_t)XSLOT;              // [par][slot][row] float
        uint32_t prog = 0;

        for (int64_t k = 0; k < my_tiles; ++k) {
            const uint32_t par = (uint32_t)(k & 1), xph = (uint32_t)((k >> 1) & 1);
            f32x2 q2 = pk2(0u, 0u);
            uint32_t mu_addr = mu0;
#pragma unroll 1
            for (int c2 = 0; c2 < ns; ++c2) {
                const uint32_t dbase = lane_addr + TM_R + db * 64u + 16u * h;
#pragma unroll
                for (int sub = 0; sub < 2; ++sub) {
                    f32x2 xc[8];
                    mbar_wait_s<32>(bar_full, ring_phase, wctx, bar0);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const ulonglong2 v = lds128(st + xo[i]);
                        const ulonglong2 m = lds128(mu_addr + 16u * (uint32_t)i);
                        xc[2 * i] = sub2(v.x, m.x);
                        xc[2 * i + 1] = sub2(v.y, m.y);
                    }
                    {
                        uint32_t l0, h0, l1, h1;
                        upk2(xc[5], l0, h0); upk2(xc[7], l1, h1);          // the last two chunks (loads return in order)
                        mbar_arrive_at_after<8 * (B_EMPTY2 - B_FULL2)>(bar_full, h0 | h1, a.zero);
                    }
                    st += (uint32_t)SM_BOX; bar_full += 8u;
                    if (bar_full == bar_full_end) { st = ring0; bar_full = bar_full0; ring_phase ^= 1u; }
                    mu_addr += 4u * BOX_COLS;
                    if (sub == 0) {
                        mbar_wait_c(BAR(B_DREADY) + 8u * db, dph, wctx, bar0);      // the chunk's six MMAs have retired
                        tc_fence_after();
                    }
                    uint32_t v[16];
                    tc_ld16(dbase + (uint32_t)(sub * BOX_COLS), v);
                    tc_wait_ld16(v);
                    if (sub == 1) {
                        tc_fence_before();
                        mbar_arrive(BAR(B_DFREE) + 8u * db);
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        f32x2 r = sub2(xc[i], pk2(v[2 * i], v[2 * i + 1]));
                        if (a.pol & 8192) r = xc[i];                              // bring-up: Q <- |xc|^2
                        if (a.pol & 16384) r = pk2(v[2 * i], v[2 * i + 1]);       // bring-up: Q <- |reconstruction|^2
                        q2 = fma2(r, r, q2);
                    }
                }
                dph ^= db; db ^= 1u;
                if (a.pol & 4096) named_bar_sync(1, NC2);                  // experiment: no skew between the two column groups
                if (tid == NC1) sts_u32_volatile(w_prog, ++prog);          // the HBM producer may move on (see TMA1)
                if (lane == 0 && (warp & 3) == 0) TC2_TRACE(2 + h, (k << 8) | (c2 + 1));
                __syncwarp();                                              // the tcgen05 loads of the next chunk are .aligned
            }
            uint32_t qa, qb;
            upk2(q2, qa, qb);
            float qq = __uint_as_float(qa) + __uint_as_float(qb);
            // ---- the two column groups of a spectrum meet in shared memory, the slices' partial Q at rank 0 --------
            const uint32_t sq = sb + (uint32_t)SM_OFF_Q + (par * (uint32_t)TILE_M + (uint32_t)row) * 4u;
            if (h == 1) sts_f1(sq, qq);
            named_bar_sync(1, NC2);
            if (h == 0) {
                qq += lds_f1(sq);
                if (rank == 0) mbar_wait_c(BAR(B_T2RDY) + 8u * par, xph, wctx, bar0);
                const float t2 = lds_f1(sb + (uint32_t)SM_OFF_T2 + (par * (uint32_t)TILE_M + (uint32_t)row) * 4u);
                if (C > 1) {
                    if (rank != 0) {
                        st_cluster_f32(map_to_rank(qbuf0 + ((par * (uint32_t)(C - 1) + rank - 1u) * (uint32_t)TILE_M + (uint32_t)row) * 4u, 0u), qq);
                        mbar_arrive_remote(map_to_rank(BAR(B_QFULL) + 8u * par, 0u));
                    } else {
                        mbar_wait_cluster(BAR(B_QFULL) + 8u * par, xph, wctx);
                        float q_own = qq;
#pragma unroll
                        for (int s = 0; s < C - 1; ++s) qq += lds_f1(qbuf0 + ((par * (uint32_t)(C - 1) + (uint32_t)s) * (uint32_t)TILE_M + (uint32_t)row) * 4u);
                        if (a.pol & 1024) {          // bring-up: Q <- rank 0's partial, T2 <- the peers' partials
                            const int64_t grow = (cid + k * ncl) * TILE_M + row;
                            if (grow < a.n) { a.Q[grow] = q_own; a.T2[grow] = qq - q_own; }
                            mbar_arrive(BAR(B_C2DONE) + 8u * par);
                            continue;
                        }
                    }
                }
                if (rank == 0) {
                    const int64_t grow = (cid + k * ncl) * TILE_M + row;
                    if (grow < a.n) {
                        if (a.T2) a.T2[grow] = t2;
                        if (a.Q) a.Q[grow] = qq;
                        if (a.accept) a.accept[grow] = rule_distance(a.rule, t2 * a.ct, qq * a.cq) < a.dlim ? 1 : 0;
                    }
                }
            }
            mbar_arrive(BAR(B_C2DONE) + 8u * par);             // this parity's scores / T2 / Q slots may be rewritten
            if (lane == 0 && (warp & 3) == 0) TC2_TRACE(2 + h, (k << 8) | 0xff);
        }
    } else if (warp < W_TMA1) {
        // =====================================================================================
        // CT: per tile, partial scores of my slice -> cluster exchange -> T2 -> split scores into TMEM.  thread = spectrum
        // =====================================================================================
        const int row = 32 * (warp & 3) + lane;
        const uint32_t lane_addr = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
        const uint32_t xbuf0 = sb + (uint32_t)SM_OFF_X;                                  // [par][slot][q][row] float4
        for (int64_t k = 0; k < my_tiles; ++k) {
            const uint32_t par = (uint32_t)(k & 1), xph = (uint32_t)((k >> 1) & 1);
            if (tid == NC1 + NC2) TC2_TRACE(4, (k << 8) | 1);
            mbar_wait_s<32>(BAR(B_TREADY), par, wctx, bar0);
            tc_fence_after();
            float t[16];
            {
                float cross[16];
                uint32_t c0[16], c1[16];
                tc_ld16(lane_addr + TM_D1 + 16u, c0); tc_ld16(lane_addr + TM_D1 + 32u, c1);          // (1,2) (1,3), accumulator 0
                tc_wait_ld16x2(c0, c1);
#pragma unroll
                for (int i = 0; i < 16; ++i) cross[i] = __uint_as_float(c0[i]) + __uint_as_float(c1[i]);
                tc_ld16(lane_addr + TM_D1 + 48u + 16u, c0); tc_ld16(lane_addr + TM_D1 + 48u + 32u, c1);   // accumulator 1
                tc_wait_ld16x2(c0, c1);
#pragma unroll
                for (int i = 0; i < 16; ++i) cross[i] += __uint_as_float(c0[i]) + __uint_as_float(c1[i]);
                tc_ld16(lane_addr + TM_D2, c0); tc_ld16(lane_addr + TM_D2 + 16u, c1);                  // (2,1) (2,2)
                tc_wait_ld16x2(c0, c1);
#pragma unroll
                for (int i = 0; i < 16; ++i) cross[i] += __uint_as_float(c0[i]) + __uint_as_float(c1[i]);
                tc_ld16(lane_addr + TM_D1, c0); tc_ld16(lane_addr + TM_D1 + 48u, c1);                  // (1,1), both accumulators
                tc_wait_ld16x2(c0, c1);
#pragma unroll
                for (int i = 0; i < 16; ++i) t[i] = (__uint_as_float(c0[i]) + __uint_as_float(c1[i])) + cross[i];
            }
            tc_fence_before();
            mbar_arrive(BAR(B_TFREE));                         // the projection of the next tile may overwrite the accumulators
            // everything this parity's slots held two tiles ago has been consumed (here AND, through the exchange that
            // follows, before any peer can finish the tile after this one)
            if (tid == NC1 + NC2) TC2_TRACE(4, (k << 8) | 2);
            mbar_wait_s<128>(BAR(B_C2DONE) + 8u * par, xph ^ 1u, wctx, bar0);
            if (tid == NC1 + NC2) TC2_TRACE(4, (k << 8) | 3);
            if (C > 1) {
                // all-to-all of the partial scores, then the same sum in rank order on every CTA.  One slot per peer:
                // a peer's slot may be rewritten once that peer has read the previous tile's partials (XFREE, remote).
                mbar_wait_cluster(BAR(B_XFREE), par ^ 1u, wctx);
#pragma unroll
                for (int d = 1; d < C; ++d) {
                    const uint32_t peer = (rank + (uint32_t)d) % (uint32_t)C;
                    const uint32_t slot = rank < peer ? rank : rank - 1u;
                    const uint32_t dst = map_to_rank(xbuf0 + slot * (uint32_t)XSLOT + (uint32_t)row * 16u, peer);
#pragma unroll
                    for (int q = 0; q < 4; ++q) st_cluster_v4(dst + (uint32_t)q * 2048u, t[4 * q], t[4 * q + 1], t[4 * q + 2], t[4 * q + 3]);
                    mbar_arrive_remote(map_to_rank(BAR(B_XFULL), peer));
                }
                mbar_wait_cluster(BAR(B_XFULL), par, wctx);
                float sum[16];
#pragma unroll
                for (int s = 0; s < C; ++s) {
                    float v[16];
                    if ((uint32_t)s == rank) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = t[i];
                    } else {
                        const uint32_t slot = (uint32_t)s < rank ? (uint32_t)s : (uint32_t)(s - 1);
                        const uint32_t src = xbuf0 + slot * (uint32_t)XSLOT + (uint32_t)row * 16u;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 f = lds_f4(src + (uint32_t)q * 2048u);
                            v[4 * q] = f.x; v[4 * q + 1] = f.y; v[4 * q + 2] = f.z; v[4 * q + 3] = f.w;
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) sum[i] = s == 0 ? v[i] : sum[i] + v[i];
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) t[i] = sum[i];
#pragma unroll
                for (int d = 1; d < C; ++d)      // my slots are read (the sum depends on every load): every peer may send its next tile
                    mbar_arrive_remote_after(map_to_rank(BAR(B_XFREE), (rank + (uint32_t)d) % (uint32_t)C),
                                             __float_as_uint(t[0]) | __float_as_uint(t[4]) | __float_as_uint(t[8]) | __float_as_uint(t[12]), a.zero);
            }
            uint32_t s1[8], s2[8], s3[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) split3_pack(t[2 * i], t[2 * i + 1], s1[i], s2[i], s3[i]);
            const uint32_t ts = lane_addr + TM_TS + 24u * par;
            tc_st8(ts, s1); tc_st8(ts + 8, s2); tc_st8(ts + 16, s3);
            tc_wait_st();
            tc_fence_before();
            mbar_arrive(BAR(B_THFULL) + 8u * par);             // release: the T2 slot reaches C2 through MMA-R's commit -> DREADY chain
            if (rank == 0) {
                float t2 = 0.f;
                if (s_diag) {
#pragma unroll
                    for (int r = 0; r < NCOMP; ++r) t2 = fmaf(s_S[r * (NCOMP + 1)] * t[r], t[r], t2);
                } else {
#pragma unroll
                    for (int r = 0; r < NCOMP; ++r) {
                        float s = 0.f;
#pragma unroll
                        for (int c = 0; c < NCOMP; ++c) s = fmaf(s_S[r * NCOMP + c], t[c], s);
                        t2 = fmaf(s, t[r], t2);
                    }
                }
                sts_f1(sb + (uint32_t)SM_OFF_T2 + (par * (uint32_t)TILE_M + (uint32_t)row) * 4u, t2);      // read by C2 after its pass 2 of this tile
                mbar_arrive(BAR(B_T2RDY) + 8u * par);          // off the critical path: the reconstruction is already running
            }
            if (tid == NC1 + NC2) TC2_TRACE(4, (k << 8) | 9);
        }
    } else if (warp == W_TMA1) {
        // =====================================================================================
        // TMA1: HBM -> ring 1, at most a.lead stages ahead of what pass 2 has consumed, so that the lines waiting in
        // L2 for their second read stay within what L2 keeps (the tile stays in L2 for pass 2: evict_last)
        // =====================================================================================
        if (lane == 0) {
            const uint64_t pol_keep = make_policy(a.pol & 3);
            int slot = 0;
            uint32_t ring_phase = 0, issued = 0;
            bool stop = false;
            for (int64_t k = 0; k < my_tiles && !stop; ++k) {
                const int row0 = (int)((cid + k * ncl) * TILE_M);
                for (int s = 0; s < ns; ++s) {
                    if (a.inject && blockIdx.x == 0 && k == 0 && s == 3) { stop = true; break; }      // test hook
                    if ((int)issued >= a.lead) wait_counter(w_prog, issued + 1u - (uint32_t)a.lead, wctx);
                    mbar_wait_c(BAR(B_EMPTY1 + slot), ring_phase ^ 1u, wctx, bar0);
                    mbar_expect_tx(BAR(B_FULL1 + slot), (uint32_t)SM_STAGE);
                    const uint32_t dst = sb + (uint32_t)SM_OFF_R1 + (uint32_t)slot * (uint32_t)SM_STAGE;
                    tma_load_2d(dst, &tmap, col0 + s * STAGE_COLS, row0, BAR(B_FULL1 + slot), pol_keep);
                    tma_load_2d(dst + (uint32_t)SM_BOX, &tmap, col0 + s * STAGE_COLS + BOX_COLS, row0, BAR(B_FULL1 + slot), pol_keep);
                    if (++slot == R1) { slot = 0; ring_phase ^= 1u; }
                    ++issued;
                    TC2_TRACE(5, (k << 8) | (s + 1));
                    if (s == (ns > 1 ? 1 : 0)) sts_u32_volatile(w_begun, (uint32_t)k + 1u);    // pass 2 of this tile may prefetch
                }
            }
        }
        __syncwarp();
    } else if (warp == W_TMA2) {
        // =====================================================================================
        // TMA2: L2 -> ring 2.  It may fetch a tile's first boxes as soon as pass 1 has REQUESTED them (a second
        // request for a line in flight merges in L2), long before C2 is ready for the tile.
        // =====================================================================================
        if (lane == 0) {
            const int k2 = (a.pol >> 2) & 3;
            const uint64_t pol_stream = make_policy(k2 == 0 ? 2 : k2 == 2 ? 0 : k2);
            int slot = 0;
            uint32_t ring_phase = 0;
            for (int64_t k = 0; k < my_tiles; ++k) {
                const int row0 = (int)((cid + k * ncl) * TILE_M);
                wait_counter(w_begun, (uint32_t)k + 1u, wctx);
                if (a.pol & 2048) wait_counter(w_prog, (uint32_t)k * (uint32_t)ns, wctx);      // experiment: no prefetch across tiles
                for (int j = 0; j < 2 * ns; ++j) {
                    mbar_wait_s<64>(BAR(B_EMPTY2 + slot), ring_phase ^ 1u, wctx, bar0);
                    mbar_expect_tx(BAR(B_FULL2 + slot), (uint32_t)SM_BOX);
                    tma_load_2d(sb + (uint32_t)off_r2<C>() + (uint32_t)slot * (uint32_t)SM_BOX, &tmap, col0 + j * BOX_COLS, row0,
                                BAR(B_FULL2 + slot), pol_stream);
                    if (++slot == r2_slots<C>()) { slot = 0; ring_phase ^= 1u; }
                    TC2_TRACE(6, (k << 8) | (j + 1));
                }
            }
        }
        __syncwarp();
    } else if (warp == W_MMAP) {
        // =====================================================================================
        // MMA-P: projection.  Per k-step (16 wavelengths) three MMAs: A1 x [P1;P2;P3] -> D1 (rotating over two
        // accumulators: 16 chained accumulations each for a 512-column slice), A2 x [P1;P2] -> D2, A3 x [P1] -> D3.
        // =====================================================================================
        constexpr uint32_t IDESC_N48 = make_idesc(TILE_M, 48, false, false);
        constexpr uint32_t IDESC_N32 = make_idesc(TILE_M, 32, false, false);
        constexpr uint32_t IDESC_N16 = make_idesc(TILE_M, 16, false, false);
        const uint32_t sP = sb + (uint32_t)SM_OFF_P;
        uint32_t ab = 0, aph = 0;
        for (int64_t k = 0; k < my_tiles; ++k) {
            mbar_wait_c(BAR(B_TFREE), (uint32_t)((k & 1) ^ 1), wctx, bar0);     // CT has read the previous tile's accumulators (tile 0: passes)
            tc_fence_after();
            uint32_t patom = sP;
#pragma unroll 1
            for (int s = 0; s < ns; ++s) {
                mbar_wait_c(BAR(B_AFULL) + 8u * ab, aph, wctx, bar0);
                tc_fence_after();
                const uint64_t d0 = make_desc(patom, 16, 1024);
                patom += (uint32_t)SM_PATOM;
                const uint32_t nz = s != 0 ? 1u : 0u;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint64_t d_b = d0 + (uint64_t)(((uint32_t)ks * 32u) >> 4);
                    const uint32_t a0 = tmem + TM_A + ab * 96u + (uint32_t)(8 * ks);
                    tc_mma_ts(tmem + TM_D1 + 48u * (uint32_t)(ks & 1), a0, d_b, IDESC_N48, (nz | (uint32_t)(ks >> 1)) != 0 ? 1u : 0u);
                    tc_mma_ts(tmem + TM_D2, a0 + 32, d_b, IDESC_N32, (nz | (uint32_t)ks) != 0 ? 1u : 0u);
                    tc_mma_ts(tmem + TM_D3, a0 + 64, d_b, IDESC_N16, 1u);
                }
                tc_commit_elect(BAR(B_AFREE) + 8u * ab);
                aph ^= ab; ab ^= 1u;
                if (lane == 0) TC2_TRACE(7, (k << 8) | (s + 1));
            }
            tc_commit_elect(BAR(B_TREADY));
        }
    } else {
        // =====================================================================================
        // MMA-R: reconstruction of my slice, 64 columns per chunk (one atom of P read MN-major), six products.
        // THFULL is per tile parity: CT can only arrive for tile k+2 after C2 -- hence this warp -- finished tile k.
        // =====================================================================================
        constexpr uint32_t IDESC2 = make_idesc(TILE_M, STAGE_COLS, false, true);
        const uint32_t sP = sb + (uint32_t)SM_OFF_P;
        const int pa[6] = {0, 0, 1, 1, 0, 2}, pb[6] = {0, 1, 0, 1, 2, 0};
        uint32_t db = 0, dph = 0;
        for (int64_t k = 0; k < my_tiles; ++k) {
            const uint32_t par = (uint32_t)(k & 1);
            mbar_wait_s<32>(BAR(B_THFULL) + 8u * par, (uint32_t)((k >> 1) & 1), wctx, bar0);
            tc_fence_after();
            const uint32_t ts = tmem + TM_TS + 24u * par;
#pragma unroll 1
            for (int c2 = 0; c2 < ns; ++c2) {
                mbar_wait_c(BAR(B_DFREE) + 8u * db, dph ^ 1u, wctx, bar0);          // C2 has read the buffer's previous chunk
                tc_fence_after();
                const uint64_t d0 = make_desc(sP + (uint32_t)c2 * (uint32_t)SM_PATOM, (uint32_t)SM_PATOM, 1024u);
                const uint32_t dst = tmem + TM_R + db * 64u;
#pragma unroll
                for (int m = 0; m < 6; ++m)
                    tc_mma_ts(dst, ts + 8u * (uint32_t)pa[m], d0 + (uint64_t)(((uint32_t)pb[m] * 2048u) >> 4), IDESC2, m != 0 ? 1u : 0u);
                tc_commit_elect(BAR(B_DREADY) + 8u * db);
                dph ^= db; db ^= 1u;
                if (lane == 0) TC2_TRACE(8, (k << 8) | (c2 + 1));
            }
        }
    }

    // ---- teardown: nobody leaves while a peer may still write into its shared memory ------------------
    tc_fence_before();
    __syncthreads();
    if (C > 1) cluster_sync();
    if (warp == W_MMAP) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

// ---- host side ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_tensor_map(CUtensorMap* map, const float* X, int64_t n, int p, int64_t ldx) {
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
            set_error("cuTensorMapEncodeTiled is not available from the driver");
            return e != cudaSuccess ? (int)e : (int)cudaErrorNotSupported;
        }
        encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    const cuuint64_t dims[2] = {(cuuint64_t)p, (cuuint64_t)n};      // boxes beyond p / n are zero filled
    const cuuint64_t strides[1] = {(cuuint64_t)ldx * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)BOX_COLS, (cuuint32_t)TILE_M};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(X), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return (int)cudaErrorInvalidValue;
    }
    return 0;
}

template <int C>
static int launch(const CUtensorMap& tmap, const Args& a, cudaStream_t stream) {
    auto kern = score_tc2_kernel<C>;
    static int max_clusters = 0;                 // per template instance: co-resident clusters of this kernel
    const int smem = (int)sm_total<C>();
    if (!max_clusters) {
        OCM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        if (C > 1) {
            cudaLaunchConfig_t q = {};
            q.gridDim = dim3((unsigned)(C * (sm_count() / C)), 1, 1);
            q.blockDim = dim3(NTHREADS, 1, 1);
            q.dynamicSmemBytes = (size_t)smem;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            q.attrs = at; q.numAttrs = 1;
            int nc = 0;
            OCM_CUDA(cudaOccupancyMaxActiveClusters(&nc, kern, &q));
            OCM_REQUIRE(nc >= 1, OCM_ELIMIT, "no cluster of this kernel fits the device");
            max_clusters = nc;
        } else {
            max_clusters = sm_count();
        }
    }
    const int64_t ntiles = (a.n + TILE_M - 1) / TILE_M;
    const int64_t ncl = ntiles < max_clusters ? ntiles : max_clusters;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(ncl * C), 1, 1);
    cfg.blockDim = dim3(NTHREADS, 1, 1);
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = C > 1 ? 1 : 0;
    OCM_CUDA(cudaLaunchKernelEx(&cfg, kern, tmap, a));
    return check_launch("score_tc2_kernel");
}

}  // namespace tc2
}  // namespace ocm

// shape gate of the cluster kernel: fp32 rows, 64 <= p <= 2048 (one, two or four CTAs of <= 512 wavelengths per tile),
// n_components <= 16, 16-byte aligned rows, no fused SNV
extern "C" int ocm_simca_score_tc2_supported(int64_t n, int32_t p, int64_t ldx, int32_t A, int32_t snv, const void* X) {
    if (snv || p < 64 || p > 4 * ocm::tc2::NS_MAX * ocm::tc2::STAGE_COLS || (p & 3) || A < 1 || A > ocm::tc2::NCOMP) return 0;
    if (ldx % 4 != 0 || (reinterpret_cast<uintptr_t>(X) & 15) != 0) return 0;
    if (n >= (1ll << 31) - ocm::tc2::TILE_M) return 0;
    return 1;
}

extern "C" int ocm_simca_score_tc2_f32(const float* X, int64_t n, int32_t p, int64_t ldx, const float* mu,
                                       const float* P, const float* Sinv, const ocm_score_params* hp,
                                       float* T2, float* Q, uint8_t* accept, int32_t inject_polls, void* stream) {
    using namespace ocm;
    if (n == 0) return 0;
    OCM_REQUIRE(X && mu && P && Sinv && hp, OCM_EINVAL, "NULL buffer");
    OCM_REQUIRE(p >= 64 && p <= 4 * tc2::NS_MAX * tc2::STAGE_COLS && (p & 3) == 0 && hp->A >= 1 && hp->A <= tc2::NCOMP && !hp->snv,
                OCM_ELIMIT, "cluster tensor-core path needs 64 <= p <= 2048 (multiple of 4), n_components <= 16, no fused SNV");
    OCM_REQUIRE(ldx % 4 == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0 && (reinterpret_cast<uintptr_t>(P) & 15) == 0,
                OCM_EALIGN, "X rows and P must be 16-byte aligned");
    OCM_REQUIRE(n < (1ll << 31) - tc2::TILE_M, OCM_ELIMIT, "row count must fit the 32-bit TMA coordinate");
    int rc = tc_status_report("ocm_simca_score_tc2_f32");          // an EARLIER launch trapped: say so before touching the driver
    if (rc) return rc;
    CUtensorMap tmap;
    rc = tc2::make_tensor_map(&tmap, X, n, p, ldx);
    if (rc) return rc;
    const int stages = (p + tc2::STAGE_COLS - 1) / tc2::STAGE_COLS;
    const int C = stages > 2 * tc2::NS_MAX ? 4 : stages > tc2::NS_MAX ? 2 : 1;
    const int ns = (stages + C - 1) / C;
    static const int env_lead = [] { const char* e = getenv("OCM_TC2_LEAD"); return e ? atoi(e) : 0; }();   // in stages (experiments)
    static const int env_pol = [] { const char* e = getenv("OCM_TC2_POL"); return e ? atoi(e) : 0; }();
    int lead = env_lead > 0 ? env_lead : ns + ns / 2;
    if (lead < ns + 1) lead = ns + 1;                 // pass 2 of a tile needs ALL of its pass 1: anything less deadlocks
    tc2::Args a{n, p, ns, mu, P, Sinv, hp->A, hp->rule, hp->ct, hp->cq, hp->dlim, T2, Q, accept,
                tc_status_device_ptr(), inject_polls > 0 ? (uint32_t)inject_polls : (1u << 24), inject_polls > 0 ? 1 : 0, lead, env_pol, 0u};
    return C == 4 ? tc2::launch<4>(tmap, a, (cudaStream_t)stream) : C == 2 ? tc2::launch<2>(tmap, a, (cudaStream_t)stream) : tc2::launch<1>(tmap, a, (cudaStream_t)stream);
}

extern "C" int ocm_simca_score_tc_debug_words(uint32_t* out16) {
    unsigned int* d = ocm::tc_status_device_ptr();      // mapped: the same address space on the host under UVA
    if (!d) return -1;
    for (int i = 0; i < 16; ++i) out16[i] = reinterpret_cast<volatile unsigned int*>(d)[i];
    return 0;
}
