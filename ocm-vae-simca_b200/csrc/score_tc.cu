// Fused SIMCA scoring on the 5th-generation tensor cores (tcgen05 + TMEM) -- the bulk path for
// config 3 (fp32 rows, 64 <= p <= 1152, A <= 16; benchmarked at p = 1024).
//
// Why tensor cores for an HBM-bound op: per spectrum 2*p*A FMAs ride on 4*p bytes, so the FP32 pipe
// and HBM balance at A = 10; FP32 operand delivery caps the FP32-pipe kernel (score_rowteam.cu) near
// 45 % of the FMA peak, i.e. ~43 % of the HBM roofline (ncu, profiles/).  tcgen05 takes its operands from
// TMEM / shared memory, not from the register file, and leaves the CUDA cores ~8 instructions per element
// (centre, split, pack, square).
//
// Precision: T2 is first-order and Q second-order sensitive to the scores t, and the
// reconstruction must be good to ~2^-22 of |x|.  Every fp32 operand is therefore split EXACTLY into three
// bf16 terms (x = b1 + b2 + b3 up to 2^-24, truncation splits, remainders exact in fp32) and each product
// is the six MMAs (1,1) (1,2) (2,1) (2,2) (1,3) (3,1) with fp32 accumulation in TMEM.  bf16 rather than
// split TF32 because the loadings must be BOTH a K-major operand (projection, K = wavelengths) and an
// MN-major operand (reconstruction, K = components): for 16-bit types the K-major and MN-major 128B-swizzle
// atoms are byte-identical, so ONE shared-memory copy of P (96 KB) serves both; kind::tf32 accepts MN-major
// only in the 32B-base swizzle, which needs a second 128 KB copy that does not fit (found on the B200:
// a plain MN-major tf32 descriptor is a silent no-op).
// The tensor core accumulates with truncation, so the large (1,1) products rotate over four accumulators
// (16 chained accumulations each, bias < 1e-6).  A tcgen05.mma that accumulates into the SAME TMEM tile as its
// predecessor waits for it (~100 cycles measured for these small N = 16 MMAs, which made the MMA warp the
// bottleneck), so the five cross products get two accumulators each as well: 14 accumulators, and no two
// consecutive MMAs of the projection touch the same one; the converter thread adds them with round-to-nearest.
// The residual is formed EXPLICITLY, r = xc - P^T t element by element (never as |xc|^2 - |t|^2, which cancels):
// the tensor core rebuilds P^T t in TMEM and the converter threads subtract it from the fp32 centred spectrum.
//
// Data flow of one CTA (persistent, one per SM, 19 warps), tile = 128 spectra x p wavelengths:
//   pass 1  stages of 128 x 64 floats stream HBM -> shared memory (TMA, two 128B-swizzled boxes per stage, 3-slot
//           mbarrier ring, L2 evict_last); each of 512 converter threads owns one TMEM lane (= spectrum) and 16
//           columns of the stage: LDS -> centre -> 3-way split -> pack -> tcgen05.st into an A-operand buffer; two
//           MMA warps issue T[128x16] += A * P^T (3 tcgen05.mma per k-step against the stacked terms of P, 12 per
//           stage, A from TMEM, P from shared memory).
//   scores  T is read back once (tcgen05.ld), T2 = t^T Sinv t stays in a register, t is split and
//           stored as the A operand of the second product.
//   pass 2  the MMA warps rebuild D[128x128] = T * P[:, chunk] chunk by chunk (6 MMAs per 128 columns, the same smem
//           copy of P read MN-major, two TMEM buffers); the same tile is streamed again -- from L2, where pass 1 left
//           it (evict_first now) -- and the converters read D back (tcgen05.ld), subtract it from the centred
//           spectrum in registers and accumulate Q = sum r^2.
// Only T2, Q and the decision go back to HBM.  All hand-offs are mbarriers; every wait is bounded.
//
// Measured on the B200 (profiles/r01_score_tc_v11.md): 8.65-8.8 ms per 10 M x 1024 spectra for a launch on its own =
// 72 % of the measured HBM peak, DRAM traffic 42.1 GB for 41.0 GB algorithmic (the second read is served by L2); in a
// long back-to-back run the board's 1000 W power cap pulls the SM clock down and the same launch takes 9.2-10.0 ms
// (tools/step_times.py) -- the kernel is instruction- and latency-bound in its converter warps, not HBM-bound.
// Tried and rejected, each A/B-timed on the same box (DESIGN.md has the numbers): a 7-slot ring of single boxes, a
// 4/5-slot ring at smaller p (ring depth is not the limiter), TMA L2 prefetch ahead of the ring, back-off sleeps in the
// polling loops, reading pass 2 straight from L2 into registers (row-strided 16 B loads: 3x slower), and interleaving
// pass 1 of tile s with pass 2 of tile s-1 (the kept working set doubles to 76 MB, L2 hit rate falls from 49 % to
// 35 %, DRAM traffic +50 %).
#include "ocm_common.cuh"
#include "score_tc_common.cuh"

#include <atomic>
#include <cuda.h>
#include <stdlib.h>

namespace ocm {
namespace tc {
using namespace tcs;

constexpr int P_MAX = 1152;             // most wavelengths supported (the loadings live in smem: 96 B per column)
constexpr int NCOMP = 16;               // MMA N of the projection / K of the reconstruction (components padded)
constexpr int TILE_M = 128;             // spectra per tile = TMEM lanes
constexpr int STAGE_COLS = 64;          // columns per stage = one 128-byte bf16 swizzle atom of P = two TMA boxes of 32 floats
constexpr int BOX_COLS = 32;            // TMA box: 128 rows x 32 floats (128 B rows, SWIZZLE_128B -> row-per-lane LDS.128 conflict free)
constexpr int NSTAGE = 3;               // ring depth (96 KB in flight per SM)
constexpr int NABUF = 2;                // A-operand buffers: 3 terms x 32 TMEM columns (64 bf16) each
constexpr int D2_COLS = 128;            // columns per residual chunk
constexpr int SUBS = D2_COLS / STAGE_COLS;                  // stages per residual chunk (2)
constexpr int NCONV = 512;              // converter threads (16 warps): TMEM lane = spectrum, 4 column quarters per stage
constexpr int NQ = NCONV / TILE_M;      // column quarters
constexpr int QCOLS = STAGE_COLS / NQ;  // 16 columns per thread per stage
constexpr int NMMAW = 2;                // MMA-issuing warps (disjoint accumulators; the issue loop of ONE warp was the bottleneck)
constexpr int NTHREADS = NCONV + 32 + 32 * NMMAW;   // + TMA producer warp + MMA warps
constexpr int NACC_MAIN = 4;            // accumulators of the A-term-1 products (one per k-step of a stage)

// TMEM column map (512 columns allocated; base address 0 because the CTA owns the whole TMEM of its SM)
constexpr uint32_t TM_ABUF = 0;         // pass 1: buffer b at [96 b, 96 b + 96): term i at +32 i ; pass 2: D2 buffers [128 b, +128)
// Projection accumulators.  The three bf16 terms of P are stacked along N in shared memory, so ONE MMA per A term
// covers several products:  A1 x [P1;P2;P3] (N=48) -> D1 = [(1,1) | (1,2) | (1,3)],  A2 x [P1;P2] (N=32) -> D2,
// A3 x [P1] (N=16) -> D3.  D1 rotates over the 4 k-steps of a stage (16 chained accumulations per tile: the
// tensor core adds into fp32 with truncation), D2 / D3 exist once per MMA warp.  [192, 256) overlaps the pass-2
// residual buffer 1, which is only live after these accumulators have been read back.
constexpr uint32_t TM_D1 = 192;         // + 48 ks
constexpr uint32_t TM_D2 = 384;         // + 32 mw
constexpr uint32_t TM_D3 = 448;         // + 16 mw
constexpr uint32_t TM_TS = 480;         // split scores: term i at +8 i (16 bf16 = 8 columns)
static_assert(TM_TS + 24 <= 512, "TMEM column budget");

constexpr size_t SM_PATOM = 3 * NCOMP * 128;                              // 6144: one 64-column atom, 48 rows = 3 terms x 16 components
constexpr size_t SM_BOX = (size_t)TILE_M * BOX_COLS * 4;                  // 16384
constexpr size_t SM_STAGE = 2 * SM_BOX;                                   // 32768
constexpr size_t SM_OFF_P = 0;
constexpr size_t SM_OFF_RING = (size_t)(P_MAX / 64) * SM_PATOM;            // 110592
constexpr size_t SM_OFF_MU = SM_OFF_RING + NSTAGE * SM_STAGE;
constexpr size_t SM_OFF_S = SM_OFF_MU + P_MAX * 4;
constexpr size_t SM_OFF_Q = SM_OFF_S + NCOMP * NCOMP * 4;
constexpr size_t SM_OFF_BAR = SM_OFF_Q + (size_t)NQ * TILE_M * 4;
constexpr size_t SM_OFF_MUS = SM_OFF_BAR + 256;                           // -2^XSHIFT * mu (fp16 A-operand mode)
constexpr size_t SM_OFF_ST = SM_OFF_MUS + P_MAX * 4;                      // fused SNV: [4 quarters][128 rows] {pivot, mean - pivot, M2, -} of a row's quarter
constexpr size_t SM_TOTAL = SM_OFF_ST + (size_t)NQ * TILE_M * 16;
static_assert(SM_TOTAL <= 232448, "shared memory budget");
// fp16 operand mode: every operand is prescaled by a power of two (exact) so that its low-order terms stay clear of the
// fp16 subnormal range, and the accumulators are scaled back (exactly) when they are read.  The spectra and the scores
// are scaled from the caller's x_rms (ocm_score_params): y = (x - mu) 2^xs with x_rms 2^xs in (8, 16] (finite up to
// ~4000 x_rms), t 2^ts with sqrt(p) x_rms 2^ts in (32, 64]; the loadings by a fixed 2^10 (|P| <= 1, promised with x_rms).
// Whatever leaves the fp16 range turns into inf / NaN distances, which the fix-up kernel below recomputes in fp32.
constexpr int PSHIFT = 10;

struct Args {
    int64_t n;
    int p;            // wavelengths (multiple of 4, <= P_MAX); processed as ns stages of 64 columns, zero padded
    int ns;           // stages per pass = 2 * ceil(p / 128)
    const float* mu; const float* P; const float* Sinv;
    int A; int rule; float ct, cq, dlim;
    float* T2; float* Q; uint8_t* accept;
    unsigned int* status;     // protocol status word (mapped host memory), see mbar_wait
    uint32_t max_polls;       // bound of every mbarrier wait
    int inject;               // test hook: CTA 0 drops one TMA stage of its first tile
    uint32_t zero;            // always 0, see mbar_arrive_at_after
    float kx, kt;             // fp16 mode: 2^xs, 2^ts (exact powers of two)
    float snv_eps;            // fused SNV: x <- (x - mean_row) / (std_row(ddof=0) + snv_eps) before the centring
    int map3d;                // the tensor map has a third dimension (32-column blocks): one TMA instruction per stage
    unsigned int* nf_flag;    // fp16 mode: receives nf_id when a row comes out non-finite (the fix-up kernel then scans)
    unsigned int nf_id;
};

// mbarrier slots
enum { B_FULL = 0, B_EMPTY = NSTAGE, B_AFULL = 2 * NSTAGE, B_AFREE = 2 * NSTAGE + NABUF, B_DFREE = 2 * NSTAGE + 2 * NABUF,
       B_DREADY = B_DFREE + 2, B_TREADY = B_DFREE + 4, B_THFULL = B_DFREE + 5, B_COUNT = B_DFREE + 6 };

static_assert(NACC_MAIN == 4 && STAGE_COLS / 16 == NACC_MAIN, "one D1 accumulator per k-step of a stage");
static_assert(NABUF == 2 && P_MAX % D2_COLS == 0, "two A buffers; residual chunks of 128 columns");
static_assert(8 * B_COUNT + 8 <= 256, "barrier block (+ TMEM base word + abort word)");

// F16A = fp16 operand mode: the A operand of the projection is a 2-term fp16 split of the prescaled centred spectrum
// (six converter instructions per pair instead of nine, two TMEM stores and two MMAs per k-step instead of three), the
// loadings and the scores are 3-term fp16 splits; false = round 1's 3-term bf16 splits everywhere.
// SNV = fused standard normal variate (simca_nuts.py:48-50): the tile is streamed a third time, FIRST, for the row mean
// and standard deviation (each converter thread over its quarter of the columns around its own pivot, the four quarters
// combined with the parallel-variance formula through shared memory); passes 1 and 2 then read the tile from L2 and fold
// the normalisation into the centring: xc = x rstd + (-m rstd - mu).
template <bool F16A, bool SNV>
__global__ void __launch_bounds__(NTHREADS, 1)
score_tc_kernel(const __grid_constant__ CUtensorMap tmap, Args a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    float* s_mu = reinterpret_cast<float*>(smem + SM_OFF_MU);
    float* s_S = reinterpret_cast<float*>(smem + SM_OFF_S);
    float* s_q = reinterpret_cast<float*>(smem + SM_OFF_Q);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + SM_OFF_BAR + 8 * B_COUNT);
    const uint32_t sb = smem_u32(smem);
    const uint32_t bar0 = sb + (uint32_t)SM_OFF_BAR;
    auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    const WaitCtx wctx{a.status, a.max_polls};

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t ntiles = (a.n + TILE_M - 1) / TILE_M;
    const int64_t my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    constexpr int W_TMA = NCONV / 32, W_MMA = NCONV / 32 + 1;   // MMA warps: W_MMA .. W_MMA + NMMAW - 1
    static_assert(NMMAW == 2, "MMA warp j owns residual buffer j");
    const int ns = a.ns, nc = a.ns / SUBS;                      // stages per pass, residual chunks per tile
    // mbarrier phase of the u-th use of a buffer is u & 1; every role keeps the phase of its next use as a running bit
    // (A buffers are used in pairs, one stage each; residual buffer b serves chunks b, b+2, ...).

    // ---- one-time setup ---------------------------------------------------------------------
    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(BAR(B_FULL + s), 1); mbar_init(BAR(B_EMPTY + s), NCONV); }
        for (int b = 0; b < NABUF; ++b) { mbar_init(BAR(B_AFULL + b), NCONV); mbar_init(BAR(B_AFREE + b), NMMAW); }
        for (int b = 0; b < 2; ++b) { mbar_init(BAR(B_DFREE + b), NCONV); mbar_init(BAR(B_DREADY + b), 1); }
        mbar_init(BAR(B_TREADY), NMMAW);
        mbar_init(BAR(B_THFULL), TILE_M);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == W_MMA) {   // the first MMA warp owns the TMEM allocation
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // loadings -> smem: three bf16 terms stacked along the row axis of the canonical 128B-swizzled atom
    //   [col/64] atoms of 48 rows (= term*16 + component) x 128 B (= 64 bf16 columns), 16-byte chunk index XOR (row % 8).
    // The SAME bytes are a K-major SW128 operand for the projection (rows = N = components, K = columns) and an
    // MN-major SW128 operand for the reconstruction (rows = K = components, MN = columns).
    const int nchunk16 = ns * 8;                                          // 16-byte chunks per row of the padded P
    for (int idx = tid; idx < NCOMP * nchunk16; idx += NTHREADS) {
        const int c = idx / nchunk16, j8 = idx % nchunk16;              // 8 columns = one 16-byte chunk
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = 0.f;
        if (c < a.A) {                                                   // columns >= p and components >= A are zero
            if (8 * j8 + 4 <= a.p) {
                const float4 lo = *reinterpret_cast<const float4*>(a.P + (size_t)c * a.p + 8 * j8);
                v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w;
            }
            if (8 * j8 + 8 <= a.p) {
                const float4 hi = *reinterpret_cast<const float4*>(a.P + (size_t)c * a.p + 8 * j8 + 4);
                v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
            }
        }
        uint32_t w1[4], w2[4], w3[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (F16A) split3_f16_pack(v[2 * e] * (float)(1 << PSHIFT), v[2 * e + 1] * (float)(1 << PSHIFT), w1[e], w2[e], w3[e]);
            else split3_pack(v[2 * e], v[2 * e + 1], w1[e], w2[e], w3[e]);
        }
        // atom (64 columns) -> row = term * 16 + component (128 B) -> 16-byte chunk XOR (row % 8)
        const size_t off = (size_t)(j8 >> 3) * SM_PATOM + (size_t)c * 128 + (size_t)((j8 & 7) ^ (c & 7)) * 16;
        *reinterpret_cast<uint4*>(smem + SM_OFF_P + off) = make_uint4(w1[0], w1[1], w1[2], w1[3]);
        *reinterpret_cast<uint4*>(smem + SM_OFF_P + off + 2048) = make_uint4(w2[0], w2[1], w2[2], w2[3]);
        *reinterpret_cast<uint4*>(smem + SM_OFF_P + off + 4096) = make_uint4(w3[0], w3[1], w3[2], w3[3]);
    }
    for (int i = tid; i < ns * STAGE_COLS; i += NTHREADS) {
        const float m = i < a.p ? a.mu[i] : 0.f;
        s_mu[i] = m;
        if (F16A) reinterpret_cast<float*>(smem + SM_OFF_MUS)[i] = -m * a.kx;
    }
    for (int i = tid; i < NCOMP * NCOMP; i += NTHREADS) {
        const int r = i / NCOMP, c = i % NCOMP;
        s_S[i] = (r < a.A && c < a.A) ? a.Sinv[r * a.A + c] : 0.f;
    }
    // cov(T) of PCA scores is diagonal, and so is the Sinv our own fit passes (diag(1/lambda)): then T2 is 16 FMAs per
    // spectrum instead of a 16 x 16 quadratic form.  Exact test (every off-diagonal entry == 0), general path otherwise.
    int offdiag = 0;
    for (int i = tid; i < a.A * a.A; i += NTHREADS) offdiag |= (i / a.A != i % a.A && a.Sinv[i] != 0.f) ? 1 : 0;
    const bool s_diag = __syncthreads_or(offdiag) == 0;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the MMA (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // The CTA is alone on its SM (shared-memory footprint) and allocates all 512 columns, so the base is 0; using
    // the constant keeps every TMEM address an immediate in the MMA issue loop.  Anything else is flagged.
    if (*s_tmem != 0u) {                                                   // uniform over the CTA: nothing below would be valid
        if (a.status) { atomicExch(a.status, 2u); __threadfence_system(); }
        __trap();
    }
    constexpr uint32_t tmem = 0u;

    if (warp < NCONV / 32) {
        // =====================================================================================
        // converter warps
        // =====================================================================================
        const int row = 32 * (warp & 3) + lane;         // TMEM lane == spectrum inside the tile
        const int qd = warp >> 2;                       // which 16 columns of a 64-column stage
        const uint32_t lane_addr = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
        // my 4 x 16 B inside the swizzled box: chunk c = 4 (qd & 1) + i lives at physical chunk c ^ (row % 8)
        const uint32_t my_off = (uint32_t)(qd >> 1) * (uint32_t)SM_BOX + (uint32_t)row * 128u;
        const int c_base = 4 * (qd & 1), rsw = row & 7;

        float t2_row = 0.f, q_acc = 0.f;

        // All per-stage state is kept incrementally in 32-bit registers (shared-space addresses, phase bits): the first
        // version recomputed slots, barrier addresses and use counts (64-bit) every stage, ~75 scalar instructions per
        // 16 elements, and those sit on the same ALU pipe as the split masks and packs (ncu: ALU pipe 55 % busy,
        // issue-limited converters).
        const uint32_t ring0 = sb + (uint32_t)SM_OFF_RING + my_off;
        uint32_t xo[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) xo[i] = (uint32_t)(((c_base + i) ^ rsw) << 4);
        const uint32_t mu0 = sb + (uint32_t)SM_OFF_MU + 4u * (uint32_t)(QCOLS * qd);
        const uint32_t bar_full0 = BAR(B_FULL), bar_full_end = BAR(B_FULL + NSTAGE);
        uint32_t st = ring0, bar_full = bar_full0, ring_phase = 0;      // the slot is implied by its FULL barrier address
        uint32_t aph = 0;      // parity of the current use of the A buffers (they advance in pairs)
        uint32_t dph = 0;      // bit b: parity of the next read of residual buffer b
        const uint32_t abase0 = lane_addr + TM_ABUF + (uint32_t)(8 * qd);
        const uint32_t dbase0 = lane_addr + TM_ABUF + (uint32_t)(QCOLS * qd);

        // fused SNV: per-row scale and offset of the normalisation (rstd, -mean rstd), set per tile by the statistics pass
        // xc = (x - ref) rstd + (-(mean - ref) rstd - mu): the mean is carried as an exact fp32 reference value plus a small
        // correction, so that neither it nor the offset it turns into is rounded at the size of the spectrum's baseline
        f32x2 snv_a = pk2(0u, 0u), snv_ak = pk2(0u, 0u), snv_b = pk2(0u, 0u), snv_bk = pk2(0u, 0u), snv_ref = pk2(0u, 0u);
        int first_col = QCOLS * qd;                   // first wavelength of my quarter in the stage being loaded (SNV: padding mask)
        // one stage: wait for the TMA, my 16 columns -> registers (centred), release the slot
        auto load_stage = [&](uint32_t mu_addr, f32x2* xc) {
            mbar_wait(bar_full, ring_phase, wctx);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const ulonglong2 v = lds128(st + xo[i]);
                const ulonglong2 m = lds128(mu_addr + 16u * (uint32_t)i);
                if (SNV) {
                    // columns beyond p are zero-filled by the TMA: after the normalisation they would not be zero
                    const bool live = first_col + 4 * i < a.p;
                    xc[2 * i] = live ? fma2(sub2(v.x, snv_ref), snv_a, sub2(snv_b, m.x)) : pk2(0u, 0u);
                    xc[2 * i + 1] = live ? fma2(sub2(v.y, snv_ref), snv_a, sub2(snv_b, m.y)) : pk2(0u, 0u);
                } else {
                    xc[2 * i] = sub2(v.x, m.x);
                    xc[2 * i + 1] = sub2(v.y, m.y);
                }
            }
            if (SNV) first_col += STAGE_COLS;
            {
                uint32_t l0, h0, l1, h1;
                upk2(xc[5], l0, h0); upk2(xc[7], l1, h1);      // the slot is released only when its loads have returned
                mbar_arrive_at_after<8 * (B_EMPTY - B_FULL)>(bar_full, h0 | h1, a.zero);
            }
            st += (uint32_t)SM_STAGE; bar_full += 8u;
            if (bar_full == bar_full_end) { st = ring0; bar_full = bar_full0; ring_phase ^= 1u; }
        };

        // same, for the fp16 A operand: y = x * 2^XSHIFT - mu * 2^XSHIFT in one packed FMA (exact scaling)
        auto load_stage_scaled = [&](uint32_t mu_addr, f32x2* y) {
            mbar_wait(bar_full, ring_phase, wctx);
            const uint32_t kx = __float_as_uint(a.kx);
            const f32x2 k2 = pk2(kx, kx);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const ulonglong2 v = lds128(st + xo[i]);
                const ulonglong2 m = lds128(mu_addr + (uint32_t)(SM_OFF_MUS - SM_OFF_MU) + 16u * (uint32_t)i);
                if (SNV) {           // y = kx ((x - mean) rstd - mu) = x (kx rstd) + (-kx mean rstd) + (-kx mu); padding stays 0 x P = 0
                    y[2 * i] = fma2(sub2(v.x, snv_ref), snv_ak, add2(snv_bk, m.x));
                    y[2 * i + 1] = fma2(sub2(v.y, snv_ref), snv_ak, add2(snv_bk, m.y));
                } else {
                    y[2 * i] = fma2(v.x, k2, m.x);
                    y[2 * i + 1] = fma2(v.y, k2, m.y);
                }
            }
            {
                uint32_t l0, h0, l1, h1;
                upk2(y[5], l0, h0); upk2(y[7], l1, h1);
                mbar_arrive_at_after<8 * (B_EMPTY - B_FULL)>(bar_full, h0 | h1, a.zero);
            }
            st += (uint32_t)SM_STAGE; bar_full += 8u;
            if (bar_full == bar_full_end) { st = ring0; bar_full = bar_full0; ring_phase ^= 1u; }
        };

        float cntq[NQ];                               // fused SNV: wavelengths (not padding) in each column quarter
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            int c = 0;
            if (SNV)
                for (int j = 0; j < ns; ++j) { const int c0 = STAGE_COLS * j + QCOLS * q; c += c0 >= a.p ? 0 : (a.p - c0 < QCOLS ? a.p - c0 : QCOLS); }
            cntq[q] = (float)c;
        }
        for (int64_t k = 0; k < my_tiles; ++k) {
            const uint32_t kpar = (uint32_t)(k & 1);
            if (SNV) {
                // ---- pass 0: row statistics.  My quarter of every stage around my own pivot (its first element), so the
                // sums stay small whatever the baseline of the spectrum; the quarters meet through shared memory.
                float piv = 0.f, s1 = 0.f, s2 = 0.f;
                int col = QCOLS * qd;
#pragma unroll 1
                for (int sgi = 0; sgi < ns; ++sgi) {
                    mbar_wait(bar_full, ring_phase, wctx);
                    float x[16];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const ulonglong2 v = lds128(st + xo[i]);
                        uint32_t a0, a1, a2, a3;
                        upk2(v.x, a0, a1); upk2(v.y, a2, a3);
                        x[4 * i] = __uint_as_float(a0); x[4 * i + 1] = __uint_as_float(a1);
                        x[4 * i + 2] = __uint_as_float(a2); x[4 * i + 3] = __uint_as_float(a3);
                    }
                    mbar_arrive_at_after<8 * (B_EMPTY - B_FULL)>(bar_full, __float_as_uint(x[11]) | __float_as_uint(x[15]), a.zero);
                    st += (uint32_t)SM_STAGE; bar_full += 8u;
                    if (bar_full == bar_full_end) { st = ring0; bar_full = bar_full0; ring_phase ^= 1u; }
                    if (sgi == 0) piv = x[0];                         // column 16 qd < 64 <= p: always a real wavelength
                    // per-stage partial sums in four lanes, then one addition into the running totals: 256 sequential fp32
                    // additions lost 1e-6 of the variance, which T2 (a ratio to small eigenvalues) shows at 2e-5
                    float p1[4] = {0.f, 0.f, 0.f, 0.f}, p2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const bool live = col + i < a.p;              // columns beyond p are zero padding
                        const float d = live ? x[i] - piv : 0.f;
                        p1[i & 3] += d; p2[i & 3] = fmaf(d, d, p2[i & 3]);
                    }
                    s1 += (p1[0] + p1[1]) + (p1[2] + p1[3]);
                    s2 += (p2[0] + p2[1]) + (p2[2] + p2[3]);
                    col += STAGE_COLS;
                }
                const float nq = cntq[qd];
                const float c_q = s1 / nq, m2_q = fmaxf(s2 - s1 * c_q, 0.f);
                float4* s_st = reinterpret_cast<float4*>(smem + SM_OFF_ST);
                s_st[qd * TILE_M + row] = make_float4(piv, c_q, m2_q, 0.f);
                named_bar_sync(2, NCONV);
                const float ref = s_st[row].x;                        // quarter 0's pivot: every quarter of the row agrees on it
                float mrel = 0.f, m2 = 0.f, ntot = 0.f;               // mean relative to ref
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    const float4 sq_ = s_st[q * TILE_M + row];
                    const float nqf = cntq[q];
                    if (nqf > 0.f) {                                  // Chan's update of (n, mean, M2)
                        const float delta = ((sq_.x - ref) + sq_.y) - mrel, nn = ntot + nqf;
                        mrel += delta * nqf / nn;
                        m2 += sq_.z + delta * delta * ntot * nqf / nn;
                        ntot = nn;
                    }
                }
                const float rstd = 1.0f / (sqrtf(m2 / ntot) + a.snv_eps);
                const float b = -mrel * rstd;
                snv_ref = pk2(__float_as_uint(ref), __float_as_uint(ref));
                snv_a = pk2(__float_as_uint(rstd), __float_as_uint(rstd));
                snv_b = pk2(__float_as_uint(b), __float_as_uint(b));
                snv_ak = pk2(__float_as_uint(rstd * a.kx), __float_as_uint(rstd * a.kx));
                snv_bk = pk2(__float_as_uint(b * a.kx), __float_as_uint(b * a.kx));
                named_bar_sync(2, NCONV);                             // the statistics buffer may be rewritten (next tile)
            }
            // ---- pass 1: centre, split, A operand of the projection (ns is even: stages come in buffer pairs) -------
            uint32_t mu_addr = mu0;
            first_col = QCOLS * qd;
#pragma unroll 1
            for (int chunk = 0; chunk < ns; chunk += 2) {
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    f32x2 xc[QCOLS / 2];
                    if (F16A) load_stage_scaled(mu_addr, xc); else load_stage(mu_addr, xc);
                    mu_addr += 4u * STAGE_COLS;
                    mbar_wait(BAR(B_AFREE + b), aph ^ 1u, wctx);           // MMAs of the buffer's previous user retired
                    tc_fence_after();
                    const uint32_t abase = abase0 + (uint32_t)(b * 96);
                    if (F16A) {
                        uint32_t w1[8], w2[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) split2_f16_pack2(xc[i], w1[i], w2[i]);
                        tc_st8(abase, w1);
                        tc_st8(abase + 32, w2);
                    } else {
                        uint32_t w1[8], w2[8], w3[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            split3_pack2(xc[i], w1[i], w2[i], w3[i]);
                        }
                        tc_st8(abase, w1);
                        tc_st8(abase + 32, w2);
                        tc_st8(abase + 64, w3);
                    }
                    tc_wait_st();
                    tc_fence_before();
                    mbar_arrive(BAR(B_AFULL + b));
                }
                aph ^= 1u;
            }
            mbar_wait(BAR(B_TREADY), kpar, wctx);
            {
            // projection complete: scores back, T2 in a register, split scores -> A operand of pass 2
            tc_fence_after();
            if (qd == 0) {
                float t[16];
                {
                    // small cross products first, then the four large (1,1) partial sums, all round-to-nearest
                    float cross[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) cross[i] = 0.f;
                    auto add_pair = [&](uint32_t c0a, uint32_t c1a) {
                        uint32_t c0[16], c1[16];
                        tc_ld16(lane_addr + c0a, c0);
                        tc_ld16(lane_addr + c1a, c1);
                        tc_wait_ld16x2(c0, c1);
#pragma unroll
                        for (int i = 0; i < 16; ++i) cross[i] += __uint_as_float(c0[i]) + __uint_as_float(c1[i]);
                    };
#pragma unroll
                    for (int ks = 0; ks < NACC_MAIN; ++ks) add_pair(TM_D1 + 48u * ks + 16u, TM_D1 + 48u * ks + 32u);   // (1,2) (1,3)
                    add_pair(TM_D2 + 0u, TM_D2 + 16u);            // (2,1) (2,2) of MMA warp 0
                    add_pair(TM_D2 + 32u, TM_D2 + 48u);           //                 MMA warp 1
                    if (!F16A) add_pair(TM_D3 + 0u, TM_D3 + 16u);   // (3,1) of both MMA warps (3-term A operand only)
                    uint32_t m0[16], m1[16];
                    tc_ld16(lane_addr + TM_D1 + 0u, m0);
                    tc_ld16(lane_addr + TM_D1 + 48u, m1);
                    tc_wait_ld16x2(m0, m1);
#pragma unroll
                    for (int i = 0; i < 16; ++i) t[i] = __uint_as_float(m0[i]) + __uint_as_float(m1[i]);
                    tc_ld16(lane_addr + TM_D1 + 96u, m0);
                    tc_ld16(lane_addr + TM_D1 + 144u, m1);
                    tc_wait_ld16x2(m0, m1);
#pragma unroll
                    for (int i = 0; i < 16; ++i) t[i] = (t[i] + (__uint_as_float(m0[i]) + __uint_as_float(m1[i]))) + cross[i];
                    if (F16A) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) t[i] *= 1.0f / (a.kx * (float)(1 << PSHIFT));      // undo the prescales (exact)
                    }
                }
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
                t2_row = t2;
                uint32_t s1[8], s2[8], s3[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (F16A) split3_f16_pack(t[2 * i] * a.kt, t[2 * i + 1] * a.kt, s1[i], s2[i], s3[i]);
                    else split3_pack(t[2 * i], t[2 * i + 1], s1[i], s2[i], s3[i]);
                }
                tc_st8(lane_addr + TM_TS, s1);
                tc_st8(lane_addr + TM_TS + 8, s2);
                tc_st8(lane_addr + TM_TS + 16, s3);
                tc_wait_st();
                tc_fence_before();
                mbar_arrive(BAR(B_THFULL));
            }
            }
            // ---- pass 2: the MMA warps rebuild T * P chunk by chunk in TMEM (they need nothing from us but a free
            // buffer); we stream the tile again, subtract the reconstruction from the centred spectrum in registers and
            // accumulate Q = sum r^2 ------------------------------------------------------------------------------
            f32x2 q2 = pk2(0u, 0u);
            mu_addr = mu0;
            first_col = QCOLS * qd;
#pragma unroll 1
            for (int c2 = 0; c2 < nc; ++c2) {
                const uint32_t b = (uint32_t)c2 & 1u;
                const uint32_t dbase = dbase0 + b * (uint32_t)D2_COLS;
#pragma unroll
                for (int sub = 0; sub < SUBS; ++sub) {
                    f32x2 xc[QCOLS / 2];
                    load_stage(mu_addr, xc);
                    mu_addr += 4u * STAGE_COLS;
                    if (sub == 0) {
                        mbar_wait(BAR(B_DREADY) + 8u * b, (dph >> b) & 1u, wctx);   // the chunk's six MMAs have retired
                        dph ^= 1u << b;
                        tc_fence_after();
                    }
                    uint32_t v[QCOLS];
                    tc_ld16(dbase + (uint32_t)(sub * STAGE_COLS), v);
                    tc_wait_ld16(v);
                    if (sub == SUBS - 1) {
                        tc_fence_before();
                        mbar_arrive(BAR(B_DFREE) + 8u * b);                    // my part of the buffer is in registers
                    }
#pragma unroll
                    for (int i = 0; i < QCOLS / 2; ++i) {
                        f32x2 r;
                        if (F16A) {          // the reconstruction was accumulated as (16 t)(1024 P): r = xc - D / (2^ts 2^10)
                            const uint32_t kd = __float_as_uint(-1.0f / (a.kt * (float)(1 << PSHIFT)));
                            r = fma2(pk2(v[2 * i], v[2 * i + 1]), pk2(kd, kd), xc[i]);
                        } else {
                            r = sub2(xc[i], pk2(v[2 * i], v[2 * i + 1]));
                        }
                        q2 = fma2(r, r, q2);
                    }
                }
            }
            {
                uint32_t qa, qb;
                upk2(q2, qa, qb);
                q_acc = __uint_as_float(qa) + __uint_as_float(qb);
            }
            {
            // combine the four column quarters of each spectrum, decide, write back
            if (qd != 0) s_q[(qd - 1) * TILE_M + row] = q_acc;
            named_bar_sync(2, NCONV);
            if (qd == 0) {
                const int64_t grow = ((int64_t)blockIdx.x + k * gridDim.x) * TILE_M + row;
                if (grow < a.n) {
                    const float qq = (q_acc + s_q[row]) + (s_q[TILE_M + row] + s_q[2 * TILE_M + row]);
                    const float tt = t2_row;
                    if (a.T2) a.T2[grow] = tt;
                    if (a.Q) a.Q[grow] = qq;
                    if (F16A && a.nf_flag && !(isfinite(tt) && isfinite(qq))) atomicMax(a.nf_flag, a.nf_id);
                    if (a.accept) a.accept[grow] = rule_distance(a.rule, tt * a.ct, qq * a.cq) < a.dlim ? 1 : 0;
                }
            }
            named_bar_sync(2, NCONV);                         // s_q reusable
            }
        }
    } else if (warp == W_TMA) {
        // =====================================================================================
        // TMA producer: one lane streams the tile twice (pass 1 keeps it in L2, pass 2 lets it go)
        // =====================================================================================
        if (lane == 0) {
            uint64_t pol_keep, pol_stream;
            asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
            int slot = 0;
            uint32_t ring_phase = 0;
            bool stop = false;
            for (int64_t k = 0; k < my_tiles && !stop; ++k) {
                const int row0 = (int)(((int64_t)blockIdx.x + k * gridDim.x) * TILE_M);
                constexpr int NPASS = SNV ? 3 : 2;                         // SNV: one more pass, first, for the row statistics
                const int nw = NPASS * ns, last0 = (NPASS - 1) * ns;
                // (no division here: an integer divide per stage on this lane's critical path -- it re-arms a ring of three
                // slots -- cost the whole kernel 3-5 %, A/B-measured)
                for (int w = 0, chunk = 0; w < nw; ++w, ++chunk) {
                    if (chunk == ns) chunk = 0;
                    const int pass = w >= last0 ? 1 : 0;                   // only the last pass lets the tile go
                    // test hook: the producer of CTA 0 dies before its 4th stage, that stage never arrives
                    if (a.inject && blockIdx.x == 0 && k == 0 && w == 3) { stop = true; break; }
                    mbar_wait(BAR(B_EMPTY + slot), ring_phase ^ 1, wctx);   // all converters drained the slot
                    mbar_expect_tx(BAR(B_FULL + slot), (uint32_t)SM_STAGE);
                    const uint32_t dst = sb + (uint32_t)SM_OFF_RING + (uint32_t)slot * (uint32_t)SM_STAGE;
                    const uint64_t pol = pass == 0 ? pol_keep : pol_stream;
                    if (a.map3d) {       // both 32-column boxes of the stage with one instruction
                        tma_load_3d(dst, &tmap, 0, row0, 2 * chunk, BAR(B_FULL + slot), pol);
                    } else {
                        tma_load_2d(dst, &tmap, chunk * STAGE_COLS, row0, BAR(B_FULL + slot), pol);
                        tma_load_2d(dst + (uint32_t)SM_BOX, &tmap, chunk * STAGE_COLS + BOX_COLS, row0, BAR(B_FULL + slot), pol);
                    }
                    if (++slot == NSTAGE) { slot = 0; ring_phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else {
        // =====================================================================================
        // MMA warps: NMMAW warps, each converged, one elected lane issues.  Warp j owns k-steps {2j, 2j+1} of every
        // projection stage (its own accumulators) and the residual chunks with buffer b == j, so every phase of a
        // buffer's barriers is waited on by ONE warp in order (a parity wait issued a phase early returns at once).
        // =====================================================================================
        const int mw = warp - W_MMA;                                        // 0 or 1
        constexpr uint32_t IDESC_N48 = make_idesc(TILE_M, 48, false, false, F16A);
        constexpr uint32_t IDESC_N32 = make_idesc(TILE_M, 32, false, false, F16A);
        constexpr uint32_t IDESC_N16 = make_idesc(TILE_M, 16, false, false, F16A);
        constexpr uint32_t IDESC2 = make_idesc(TILE_M, D2_COLS, false, true, F16A);  // B = P read MN-major
        constexpr uint32_t lbo2 = (uint32_t)SM_PATOM, sbo2 = 1024u;
        const uint32_t sP = sb + (uint32_t)SM_OFF_P;
        // the six products of a 3-term x 3-term split that matter at 2^-24: (A term, B term)
        const int pa[6] = {0, 0, 1, 1, 0, 2}, pb[6] = {0, 1, 0, 1, 2, 0};
        uint32_t aph = 0, dph = 0;          // running phase of the A buffers (pairwise) and of MY residual buffer
        for (int64_t k = 0; k < my_tiles; ++k) {
            // ---- projection: three MMAs per k-step (A term i against the stacked terms of P) ---------
            uint32_t patom = sP;
#pragma unroll 1
            for (int chunk = 0; chunk < ns; chunk += 2) {
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    mbar_wait(BAR(B_AFULL + b), aph, wctx);
                    tc_fence_after();
                    // K-major SW128: one 64-column atom per stage, K advances 32 B (16 bf16) inside the atom, 8-row groups
                    // 1024 B apart (rows = term*16 + component, so N = 48 / 32 / 16 selects [P1;P2;P3] / [P1;P2] / [P1])
                    const uint64_t d0 = make_desc(patom, 16, 1024);
                    patom += (uint32_t)SM_PATOM;
                    const uint32_t first = (chunk | b) != 0 ? 1u : 0u;
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {
                        const int ks = 2 * mw + kk;
                        const uint64_t d_b = d0 + (uint64_t)(((uint32_t)ks * 32u) >> 4);
                        const uint32_t a0 = tmem + TM_ABUF + (uint32_t)(b * 96 + 8 * ks);
                        tc_mma_ts(tmem + TM_D1 + 48u * (uint32_t)ks, a0, d_b, IDESC_N48, first);
                        tc_mma_ts(tmem + TM_D2 + 32u * (uint32_t)mw, a0 + 32, d_b, IDESC_N32, (first | (uint32_t)kk) != 0 ? 1u : 0u);
                        if (!F16A) tc_mma_ts(tmem + TM_D3 + 16u * (uint32_t)mw, a0 + 64, d_b, IDESC_N16, (first | (uint32_t)kk) != 0 ? 1u : 0u);
                    }
                    tc_commit_elect(BAR(B_AFREE + b));
                }
                aph ^= 1u;
            }
            tc_commit_elect(BAR(B_TREADY));
            // ---- reconstruction: D = T * P (chunks with buffer b == mw) -----------------------------
            mbar_wait(BAR(B_THFULL), (uint32_t)(k & 1), wctx);
            tc_fence_after();
#pragma unroll 1
            for (int c2 = mw; c2 < nc; c2 += NMMAW) {
                mbar_wait(BAR(B_DFREE + mw), dph ^ 1u, wctx);   // every converter has read the buffer's previous chunk
                dph ^= 1u;
                tc_fence_after();
                // MN-major SW128: 64-column atoms SM_PATOM apart (LBO), 8-component groups 1024 B apart (SBO); K = 16 in
                // one MMA; the B term selects the 16-row slab (2048 B) inside the atom
                const uint64_t d0 = make_desc(sP + (uint32_t)(c2 * (D2_COLS / 64)) * (uint32_t)SM_PATOM, lbo2, sbo2);
                const uint32_t dst = tmem + TM_ABUF + (uint32_t)(mw * D2_COLS);
#pragma unroll
                for (int m = 0; m < 6; ++m)
                    tc_mma_ts(dst, tmem + TM_TS + 8u * (uint32_t)pa[m], d0 + (uint64_t)(((uint32_t)pb[m] * 2048u) >> 4), IDESC2, m != 0 ? 1u : 0u);
                tc_commit_elect(BAR(B_DREADY + mw));
            }
        }
    }

    // ---- teardown -----------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}


// ---- host side: TMA descriptor of the spectra matrix ---------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_tensor_map(CUtensorMap* map, const float* X, int64_t n, int p, int64_t ldx, bool three_d) {
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
    CUresult r;
    if (three_d) {
        // (32 columns, rows, 32-column blocks): a box of {32, 128, 2} is a whole 64-column stage, laid out in shared memory
        // as the two swizzled 128 x 32 boxes the converters expect.  Only for p % 32 == 0 (a partial block would not be
        // zero filled but read on into the next row).
        const cuuint64_t dims[3] = {(cuuint64_t)BOX_COLS, (cuuint64_t)n, (cuuint64_t)(p / BOX_COLS)};
        const cuuint64_t strides[2] = {(cuuint64_t)ldx * sizeof(float), (cuuint64_t)BOX_COLS * sizeof(float)};
        const cuuint32_t box[3] = {(cuuint32_t)BOX_COLS, (cuuint32_t)TILE_M, 2};
        const cuuint32_t estr[3] = {1, 1, 1};
        r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(X), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        const cuuint64_t dims[2] = {(cuuint64_t)p, (cuuint64_t)n};      // columns >= p of a box are zero filled
        const cuuint64_t strides[1] = {(cuuint64_t)ldx * sizeof(float)};
        const cuuint32_t box[2] = {(cuuint32_t)BOX_COLS, (cuuint32_t)TILE_M};
        const cuuint32_t estr[2] = {1, 1};
        r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(X), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return (int)cudaErrorInvalidValue;
    }
    return 0;
}

}  // namespace tc
}  // namespace ocm

// ---- fix-up of the fp16 operand mode -------------------------------------------------------------------------------
// Scans T2 / Q (8 bytes per row, coalesced) and recomputes in plain fp32 every row whose distances are not finite: an
// operand that left the fp16 range, or a spectrum that held NaN / inf to begin with (which then comes out NaN again,
// rejected, like the reference).  A warp takes 32 rows at a time and works together on each flagged one.
namespace ocm {
namespace tc {
__global__ void __launch_bounds__(256) score_fixup_kernel(const float* __restrict__ X, int64_t n, int p, int64_t ldx,
                                                          const float* __restrict__ mu, const float* __restrict__ P,
                                                          const float* __restrict__ Sinv, int A, int rule, float ct, float cq,
                                                          float dlim, int snv, float snv_eps, float* T2, float* Q, uint8_t* accept,
                                                          const unsigned int* nf_flag, unsigned int nf_id) {
    // my launch (or a later one on another stream: atomicMax) has stored an id >= mine if it saw such a row; an earlier flag is < mine
    if (nf_flag && *reinterpret_cast<const volatile unsigned int*>(nf_flag) < nf_id) return;
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t base = warp * 32; base < n; base += nwarps * 32) {
        const int64_t i = base + lane;
        const bool flagged = i < n && !(isfinite(T2[i]) && isfinite(Q[i]));
        unsigned todo = __ballot_sync(0xffffffffu, flagged);
        while (todo) {
            const int l = __ffs(todo) - 1;
            todo &= todo - 1;
            const int64_t r = base + l;
            const float* x = X + r * ldx;
            float sa = 1.f, sb = 0.f;                       // fused SNV: x <- x sa + sb (two passes over the row: mean, then variance)
            if (snv) {
                float m = 0.f, v = 0.f;
                for (int j = lane; j < p; j += 32) m += x[j];
                m = warp_sum(m) / (float)p;
                for (int j = lane; j < p; j += 32) { const float d = x[j] - m; v = fmaf(d, d, v); }
                v = warp_sum(v) / (float)p;
                sa = 1.0f / (sqrtf(v) + snv_eps);
                sb = -m * sa;
            }
            float t[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) t[c] = 0.f;
            for (int j = lane; j < p; j += 32) {
                const float xc = fmaf(x[j], sa, sb) - mu[j];
#pragma unroll
                for (int c = 0; c < 16; ++c)
                    if (c < A) t[c] = fmaf(xc, P[(size_t)c * p + j], t[c]);
            }
#pragma unroll
            for (int c = 0; c < 16; ++c) t[c] = warp_sum(t[c]);
            float q = 0.f;
            for (int j = lane; j < p; j += 32) {
                float rj = fmaf(x[j], sa, sb) - mu[j];
#pragma unroll
                for (int c = 0; c < 16; ++c)
                    if (c < A) rj = fmaf(-t[c], P[(size_t)c * p + j], rj);
                q = fmaf(rj, rj, q);
            }
            q = warp_sum(q);
            float t2 = 0.f;
            for (int a = 0; a < A; ++a) {
                float sa = 0.f;
                for (int b = 0; b < A; ++b) sa = fmaf(Sinv[a * A + b], t[b], sa);
                t2 = fmaf(sa, t[a], t2);
            }
            if (lane == 0) {
                T2[r] = t2;
                Q[r] = q;
                if (accept) accept[r] = rule_distance(rule, t2 * ct, q * cq) < dlim ? 1 : 0;
            }
        }
    }
}
}  // namespace tc
}  // namespace ocm

static std::atomic<int> g_inject_polls{0};     // test hook, see ocm_simca_score_tc_inject_stall

// second-generation kernel (score_tc2.cu): concurrent passes, wavelengths split over a CTA cluster; p <= 1024
extern "C" int ocm_simca_score_tc2_supported(int64_t n, int32_t p, int64_t ldx, int32_t A, int32_t snv, const void* X);
extern "C" int ocm_simca_score_tc2_f32(const float* X, int64_t n, int32_t p, int64_t ldx, const float* mu, const float* P,
                                       const float* Sinv, const ocm_score_params* hp, float* T2, float* Q, uint8_t* accept,
                                       int32_t inject_polls, void* stream);

// Shape gate of the dispatcher: fp32 rows of 1024 wavelengths, n_components <= 16, 16-byte aligned rows, no fused
// SNV, and enough rows to fill the machine (below ~2 tiles per SM the FP32-pipe kernels have less fixed cost:
// this kernel stages 96 KB of loadings per CTA).  OCM_DISABLE_TC=1 forces the FP32-pipe kernels (A/B runs).
extern "C" int ocm_simca_score_tc_supported(int64_t n, int32_t p, int64_t ldx, int32_t A, int32_t snv, const void* X) {
    static const bool disabled = [] { const char* e = getenv("OCM_DISABLE_TC"); return e && e[0] == '1'; }();
    if (disabled) return 0;
    if (p > ocm::tc::P_MAX) return n >= 32768 ? ocm_simca_score_tc2_supported(n, p, ldx, A, snv, X) : 0;
    if (p < 64 || p > ocm::tc::P_MAX || (p & 3) || A < 1 || A > ocm::tc::NCOMP) return 0;      // fused SNV: supported
    if (ldx % 4 != 0 || (reinterpret_cast<uintptr_t>(X) & 15) != 0) return 0;
    if (n < 32768 || n >= (1ll << 31) - ocm::tc::TILE_M) return 0;
    return 1;
}

extern "C" int ocm_simca_score_tc_f32(const float* X, int64_t n, int32_t p, int64_t ldx, const float* mu,
                                      const float* P, const float* Sinv, const ocm_score_params* hp,
                                      float* T2, float* Q, uint8_t* accept, void* stream) {
    using namespace ocm;
    if (n == 0) return 0;
    OCM_REQUIRE(X && mu && P && Sinv && hp, OCM_EINVAL, "NULL buffer");
    // OCM_SCORE_TC2=1: the cluster kernel (score_tc2.cu) for p <= 1024 -- measured ~10 % behind this one, kept opt-in
    static const bool use_tc2 = [] { const char* e = getenv("OCM_SCORE_TC2"); return e && e[0] == '1'; }();
    if ((use_tc2 && p <= 1024) || p > tc::P_MAX)        // wider than this kernel's loadings buffer: clusters of four
        return ocm_simca_score_tc2_f32(X, n, p, ldx, mu, P, Sinv, hp, T2, Q, accept, g_inject_polls.exchange(0), stream);
    OCM_REQUIRE(p >= 64 && p <= tc::P_MAX && (p & 3) == 0 && hp->A >= 1 && hp->A <= tc::NCOMP, OCM_ELIMIT,
                "tensor-core path needs 64 <= p <= 1152 (multiple of 4), n_components <= 16");
    OCM_REQUIRE(ldx % 4 == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0 && (reinterpret_cast<uintptr_t>(P) & 15) == 0,
                OCM_EALIGN, "X rows and P must be 16-byte aligned");
    OCM_REQUIRE(n < (1ll << 31) - tc::TILE_M, OCM_ELIMIT, "row count must fit the 32-bit TMA coordinate");
    int rc = tc_status_report("ocm_simca_score_tc_f32");          // an EARLIER launch trapped: say so before touching the driver
    if (rc) return rc;
    CUtensorMap tmap;
    static const bool map2d_only = [] { const char* e = getenv("OCM_SCORE_TMA2D"); return e && e[0] == '1'; }();      // A/B switch
    const bool map3d = !map2d_only && p % tc::BOX_COLS == 0;
    rc = tc::make_tensor_map(&tmap, X, n, p, ldx, map3d);
    if (rc) return rc;
    const int inject = g_inject_polls.exchange(0);
    tc::Args a{n, p, 2 * ((p + tc::D2_COLS - 1) / tc::D2_COLS), mu, P, Sinv, hp->A, hp->rule, hp->ct, hp->cq, hp->dlim, T2, Q, accept,
               tc_status_device_ptr(), inject > 0 ? (uint32_t)inject : (1u << 24), inject > 0 ? 1 : 0, 0u, 1.0f, 1.0f, 1e-8f, map3d ? 1 : 0, nullptr, 0u};
    // fp16 operand mode: only when the caller states the scale of the data (x_rms) and asks for both distances (the
    // fix-up kernel finds out-of-range rows by their non-finite T2 / Q).  OCM_SCORE_BF16A=1 forces the bf16 x 3 mode.
    static const bool bf16a = [] { const char* e = getenv("OCM_SCORE_BF16A"); return e && e[0] == '1'; }();
    const bool f16 = !bf16a && hp->x_rms > 0.f && isfinite(hp->x_rms) && T2 && Q;
    if (f16) {
        int e = 0;
        frexpf(hp->x_rms, &e);                                     // x_rms = m 2^e, m in [0.5, 1)
        a.kx = ldexpf(1.0f, 4 - e);                                // x_rms kx in [8, 16)
        frexpf(hp->x_rms * sqrtf((float)p), &e);
        a.kt = ldexpf(1.0f, 6 - e);                                // |xc| of a typical row times kt in [32, 64)
    }
    a.snv_eps = hp->snv_eps;
    static std::atomic<unsigned int> launch_id{0};
    a.nf_flag = f16 ? tc_flag_device_ptr() : nullptr;          // NULL (allocation failed): the fix-up scans unconditionally
    a.nf_id = launch_id.fetch_add(1) + 1;
    auto kern = f16 ? (hp->snv ? tc::score_tc_kernel<true, true> : tc::score_tc_kernel<true, false>)
                    : (hp->snv ? tc::score_tc_kernel<false, true> : tc::score_tc_kernel<false, false>);
    OCM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::SM_TOTAL));
    const int64_t ntiles = (n + tc::TILE_M - 1) / tc::TILE_M;
    int64_t blocks = sm_count();
    if (blocks > ntiles) blocks = ntiles;
    kern<<<(unsigned)blocks, tc::NTHREADS, tc::SM_TOTAL, (cudaStream_t)stream>>>(tmap, a);
    rc = check_launch("score_tc_kernel");
    if (rc || !f16) return rc;
    tc::score_fixup_kernel<<<(unsigned)(2 * sm_count()), 256, 0, (cudaStream_t)stream>>>(X, n, p, ldx, mu, P, Sinv, hp->A, hp->rule, hp->ct,
                                                                                      hp->cq, hp->dlim, hp->snv, hp->snv_eps, T2, Q, accept, a.nf_flag, a.nf_id);
    return check_launch("score_fixup_kernel");
}

// synchronise, then read the protocol status word of the tcgen05 kernels (0 = clean)
extern "C" int ocm_simca_score_tc_status(void) {
    const bool failed = cudaDeviceSynchronize() != cudaSuccess;
    const int v = ocm::tc_status_take();
    return (failed && v == 0) ? -1 : v;
}

extern "C" int ocm_simca_score_tc_inject_stall(int32_t max_polls) {
    OCM_REQUIRE(max_polls >= 16, OCM_EINVAL, "max_polls must be >= 16");
    g_inject_polls.store(max_polls);
    return 0;
}
