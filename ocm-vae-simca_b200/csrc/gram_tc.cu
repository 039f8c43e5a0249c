// PCA-fit Gram on the 5th-generation tensor cores with EXACT accumulation (BASELINE configs[3]):
// G += (X - shift)^T (X - shift) for the bulk fit (50 M x 2048 over 8 GPUs), replacing the SVD inside sklearn's
// PCA.fit (/root/reference/utils/SIMCA.py:64-66, :75).  The FP64 tensor pipe (moments.cu, DMMA) saturates at
// ~27 TFLOP/s; tcgen05 adds into fp32 with truncation, which a Gram of correlated spectra cannot afford
// (DESIGN.md).  So the contraction is done in INTEGER arithmetic, which the tensor core performs exactly:
//
//   1. quantise   v[r][j] = rint((x[r][j] - shift[j]) * 2^sc[j])  with 2^sc[j] chosen per column from max|x - shift| so
//                 that |v| < 2^30 (31-bit fixed point: the quantum is 2^-30 of the column's largest deviation,
//                 64x finer than an fp32 ulp of that deviation), and split v into four BALANCED base-256 digits
//                 v = d0 + 256 d1 + 256^2 d2 + 256^3 d3, d_k in [-128, 127]  (u = v + 0x808080, digits = bytes(u) ^ 0x80),
//                 written as four int8 planes [rows][p] (same bytes as the fp32 input);
//   2. multiply   G_q = sum_{k,l} 256^(k+l) D_k^T D_l with tcgen05.mma.kind::i8 (int8 x int8 -> int32 in TMEM, exact):
//                 one accumulator per weight d = k + l, 128 x 128 output tile per CTA (upper-triangle tiles only),
//                 operands straight from the TMA-written 128B-swizzled boxes as MN-major descriptors (no transpose:
//                 the planes keep the row-major layout of X), 3-stage mbarrier ring of 64-row stages;
//                 a chain is at most 32 704 rows (4 products x 2^14 x rows < 2^31), then
//   3. flush      the int32 accumulators are combined in int64, converted to fp64, scaled by 2^-(sc_i + sc_j) and
//                 added into the CTA's own fp64 tile of a partial buffer (every (tile, K split) has exactly one owner, so
//                 a plain vector read-modify-write: the first version used red.global.add.f64 and spent half its time
//                 in 16 K atomics per CTA); one small kernel folds the partials into G at the end of the call.
//   The diagonal of G does not come from the MMAs but from sum_r q^2 accumulated in fp64 by the quantiser, so it is
//   exact to fp64 rounding whatever `products` is.  The quantiser of chunk c+1 runs on a side stream under the MMAs of
//   chunk c (double-buffered planes).
// The result is the exact Gram of the quantised data for `products` = 16; 13 drops the three products of weight
// < 256^2 (relative error <= 2^-37 of max_i max_j per row), 10 those below 256^3 (<= 2^-28).  Column sums of v are
// accumulated as int64 in the same pass, so cov = (G - N d d^T) / (N - 1) is the covariance OF THE QUANTISED ROWS to
// fp64 rounding.  Four accumulators of 128 columns fill the 512 TMEM columns, so 13 / 16 products run as a second
// launch over the low digit planes (weights 256^0..256^2).
#include "ocm_common.cuh"
#include "tc_ptx.cuh"

#include <mutex>
#include <stdlib.h>

namespace ocm {
namespace gtc {

using namespace ptx;

constexpr int S = 4;                        // int8 digit planes
constexpr int TILE = 128;                   // row block of the output tile = MMA M; column blocks are TILE wide too
constexpr int KSTEP = 32;                   // K of one int8 MMA
constexpr int NTHREADS = 192;               // warp 0: TMA, warp 1: MMA issue + TMEM owner, warps 2..5: flush
constexpr int MAX_CHAIN_ROWS = 32704;       // 4 products x 128 x 128 x rows < 2^31
constexpr int V_LIMIT = 0x5FFFFFFF;         // clamp of the fixed-point value (top digit stays <= 96)
constexpr int QROWS = 64;
#ifndef OCM_GRAM_DEFAULT_MODE
#define OCM_GRAM_DEFAULT_MODE 3
#endif
//                   // the quantiser pads a chunk to a multiple of this many rows
// NB = column blocks per CTA tile.  NB = 1: 128 x 128 tiles, four accumulators (one pass over the data for the ten
// leading products).  NB = 2: 128 x 256 tiles, MMA N = 256, two accumulators, two weights per pass -- at N = 128 the
// operand reads of the MMAs (8 KB per 64 cycles = 128 B/clk) plus the TMA writes of the next stage exceed the
// 128 B/clk of the shared-memory pipe (measured: 62 % of the MMA rate); at N = 256 an MMA reads 12 KB per 128 cycles.
template <int NB> struct Cfg {
    static constexpr int KB = NB == 1 ? 64 : 32;                          // rows per pipeline stage (KB / 32 MMAs per product)
    static constexpr int BOX_BYTES = TILE * KB;                           // one plane's box: KB rows of 128 B, SWIZZLE_128B
    static constexpr int STAGE_BYTES = (S + S * NB) * BOX_BYTES;          // A planes + B planes x column blocks
    static constexpr int NSTAGE = NB == 1 ? 3 : 4;
    static constexpr size_t SM_OFF_BAR = (size_t)NSTAGE * STAGE_BYTES;
    static constexpr size_t SM_TOTAL = SM_OFF_BAR + 128 + 1024;           // + barriers + slack for the 1024-byte alignment
    static constexpr int MAX_ACC = 4 / NB;
};

// A launch ("pass") accumulates the digit products (k, l) of weights 256^d, d in [DLO, DHI], one TMEM accumulator per
// weight.  The pass is a TEMPLATE parameter: with the product list in a run-time table every MMA operand went through
// R2UR and the single issuing warp needed ~100 cycles per 64-cycle MMA (all three kernel variants sat at ~60 % of the
// MMA rate whatever their shared-memory traffic); with compile-time products the descriptors are constant offsets on
// uniform registers.
template <int DLO, int DHI> struct Pass {
    static constexpr int NACC = DHI - DLO + 1;
    static constexpr int DMIN = DLO;
    static constexpr int KMIN = DLO > S - 1 ? DLO - (S - 1) : 0;       // lowest digit plane any product of the pass touches
    static constexpr int KMAX = DHI < S - 1 ? DHI : S - 1;
    static constexpr int NLOAD = KMAX - KMIN + 1;
    static constexpr int count() { int c = 0; for (int d = DLO; d <= DHI; ++d) for (int k = 0; k < S; ++k) { const int l = d - k; if (l >= 0 && l < S) ++c; } return c; }
};

struct Args {
    int p;                  // columns of G
    int nt;                 // 128-column blocks per edge
    int ntiles;             // CTA tiles (upper triangle)
    int nstages;            // stages in this chunk (ceil(rows / KB))
    int stages_per_split;
    const double* inv_scale;    // [p_pad] 2^-sc
    double* part;           // [splits][ntiles][8 NB][128][16] fp64 partial tiles, each owned by ONE CTA of a launch
    unsigned int* status;   // protocol status word (mapped host memory), see mbar_wait_bounded
};

// tile index -> (row block ti, first column block tj); NB = 2 tiles cover column blocks tj, tj + 1 with tj even
template <int NB>
__host__ __device__ inline void tile_coords(int tile, int nt, int& ti, int& tj) {
    if (NB == 1) {
        int t = tile; ti = 0;
        while (t >= nt - ti) { t -= nt - ti; ++ti; }
        tj = ti + t;
    } else {
        const int nt2 = (nt + 1) / 2;
        int t = tile; ti = 0;
        while (t >= nt2 - ti / 2) { t -= nt2 - ti / 2; ++ti; }
        tj = 2 * (ti / 2 + t);
    }
}
template <int NB>
inline int tile_count(int nt) {
    if (NB == 1) return nt * (nt + 1) / 2;
    const int nt2 = (nt + 1) / 2;
    int c = 0;
    for (int ti = 0; ti < nt; ++ti) c += nt2 - ti / 2;
    return c;
}

// ------------------------------------------------------------------------------------------ quantise
// One warp = one row x 128 columns per step (lane -> 4 consecutive columns), 8 warps walk the rows of the block's slab.
template <typename T>
__global__ void __launch_bounds__(256)
gram_quantize_kernel(const T* __restrict__ X, int64_t ldx, const int64_t* __restrict__ row_idx, int64_t r0, int rows,
                     int rows_pad, int p, int p_pad, const double* __restrict__ shift, const double* __restrict__ scale,
                     uint8_t* __restrict__ planes, int64_t plane_stride, long long* __restrict__ colsum_q,
                     double* __restrict__ colsumsq_q, unsigned long long* __restrict__ flags, int rows_per_block,
                     int vec_ok) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = blockIdx.x * TILE + 4 * lane;
    const int rb = blockIdx.y * rows_per_block;
    const int re = min(rows_pad, rb + rows_per_block);
    double sh[4], sc[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const bool ok = c0 + e < p;
        sh[e] = ok ? shift[c0 + e] : 0.0;
        sc[e] = ok ? scale[c0 + e] : 0.0;
    }
    long long cs[4] = {0, 0, 0, 0};
    double css[4] = {0.0, 0.0, 0.0, 0.0};
    unsigned int clamped = 0, nonfinite = 0;
    for (int r = rb + warp; r < re; r += 8) {
        float x[4] = {0.f, 0.f, 0.f, 0.f};
        double xd[4] = {0.0, 0.0, 0.0, 0.0};
        const bool live = r < rows;
        if (live) {
            const int64_t row = row_idx ? row_idx[r0 + r] : r0 + r;
            const T* xr = X + row * ldx;
            if constexpr (sizeof(T) == 4) {
                if (vec_ok && c0 + 3 < p) {
                    const float4 v4 = ldg_stream4(reinterpret_cast<const float*>(xr) + c0);
                    x[0] = v4.x; x[1] = v4.y; x[2] = v4.z; x[3] = v4.w;
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) if (c0 + e < p) x[e] = (float)xr[c0 + e];
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) xd[e] = (double)x[e];
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) if (c0 + e < p) xd[e] = (double)xr[c0 + e];
            }
        }
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const double y = live ? (xd[e] - sh[e]) * sc[e] : 0.0;   // scale is an exact power of two; padding rows are 0
            if (live && c0 + e < p && !(fabs(y) < 1e300)) ++nonfinite;    // NaN or inf input
            int v = __double2int_rn(y);                         // NaN -> 0, saturates
            if (v > V_LIMIT) { v = V_LIMIT; ++clamped; }
            if (v < -V_LIMIT) { v = -V_LIMIT; ++clamped; }
            cs[e] += v;
            const double vd = (double)v;
            css[e] = fma(vd, vd, css[e]);
            w[e] = ((uint32_t)v + 0x00808080u) ^ 0x00808080u;   // bytes = balanced digits d0..d3
        }
        // byte transpose: plane k gets byte k of the four columns
        const uint32_t lo01 = __byte_perm(w[0], w[1], 0x5140), hi01 = __byte_perm(w[0], w[1], 0x7362);
        const uint32_t lo23 = __byte_perm(w[2], w[3], 0x5140), hi23 = __byte_perm(w[2], w[3], 0x7362);
        const uint32_t d0 = __byte_perm(lo01, lo23, 0x5410), d1 = __byte_perm(lo01, lo23, 0x7632);
        const uint32_t d2 = __byte_perm(hi01, hi23, 0x5410), d3 = __byte_perm(hi01, hi23, 0x7632);
        uint8_t* out = planes + (int64_t)r * p_pad + c0;
        *reinterpret_cast<uint32_t*>(out) = d0;
        *reinterpret_cast<uint32_t*>(out + plane_stride) = d1;
        *reinterpret_cast<uint32_t*>(out + 2 * plane_stride) = d2;
        *reinterpret_cast<uint32_t*>(out + 3 * plane_stride) = d3;
    }
    __shared__ long long s_cs[8][TILE];
    __shared__ double s_css[8][TILE];
#pragma unroll
    for (int e = 0; e < 4; ++e) { s_cs[warp][4 * lane + e] = cs[e]; s_css[warp][4 * lane + e] = css[e]; }
    __syncthreads();
    if (threadIdx.x < TILE) {
        long long t = 0;
        double t2 = 0.0;
#pragma unroll
        for (int w8 = 0; w8 < 8; ++w8) { t += s_cs[w8][threadIdx.x]; t2 += s_css[w8][threadIdx.x]; }
        const int c = blockIdx.x * TILE + threadIdx.x;
        if (c < p && t != 0) atomicAdd(reinterpret_cast<unsigned long long*>(colsum_q + c), (unsigned long long)t);
        if (c < p && t2 != 0.0) atomicAdd(colsumsq_q + c, t2);
    }
    clamped = __reduce_add_sync(0xffffffffu, clamped);
    nonfinite = __reduce_add_sync(0xffffffffu, nonfinite);
    if (lane == 0) {
        if (clamped) atomicAdd(flags, (unsigned long long)clamped);
        if (nonfinite) atomicAdd(flags + 1, (unsigned long long)nonfinite);
    }
}

// scale[j] = 2^sc, inv_scale[j] = 2^-sc with sc = 30 - e, max|x - shift| = f 2^e (f in [0.5, 1)): |v| <= 2^30
__global__ void gram_scale_kernel(const float* __restrict__ colmin, const float* __restrict__ colmax,
                                  const double* __restrict__ shift, int p, int p_pad, double* __restrict__ scale,
                                  double* __restrict__ inv_scale, long long* __restrict__ colsum_q,
                                  double* __restrict__ colsumsq_q) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= p_pad) return;
    colsumsq_q[j] = 0.0;
    int sc = 0;
    if (j < p) {
        const double m = fmax((double)colmax[j] - shift[j], shift[j] - (double)colmin[j]) * (1.0 + 1e-6);
        if (m > 0.0 && m < 1e300) {
            int e;
            frexp(m, &e);
            sc = 30 - e;
            if (sc > 1000) sc = 1000;
            if (sc < -1000) sc = -1000;
        }
    }
    scale[j] = ldexp(1.0, sc);
    inv_scale[j] = ldexp(1.0, -sc);
    colsum_q[j] = 0;
}

// G[i][j] += sum over K splits of the partial tiles (upper triangle); the diagonal comes from the quantiser's sum q^2
template <int NB>
__global__ void __launch_bounds__(256)
gram_fold_kernel(const double* __restrict__ part, int splits, int nt, int ntiles, int p, const double* __restrict__ colsumsq_q,
                 const double* __restrict__ inv_scale, double* __restrict__ G) {
    const int tile = blockIdx.x;
    int ti, tj;
    tile_coords<NB>(tile, nt, ti, tj);
    constexpr int TW = TILE * NB, TSZ = TILE * TW;
    for (int idx = threadIdx.x; idx < TSZ; idx += blockDim.x) {
        const int r = idx / TW, c = idx % TW;
        const int i = ti * TILE + r, j = tj * TILE + c;
        if (i >= p || j >= p || j < i) continue;
        double acc;
        if (i == j) {
            acc = colsumsq_q[i] * inv_scale[i] * inv_scale[i];
        } else {
            acc = 0.0;
            const int64_t off = ((int64_t)(c >> 4) * TILE + r) * 16 + (c & 15);
            for (int sp = 0; sp < splits; ++sp) acc += part[((int64_t)sp * ntiles + tile) * (int64_t)TSZ + off];
        }
        G[(int64_t)i * p + j] += acc;
    }
}

__global__ void gram_colsum_kernel(const long long* __restrict__ colsum_q, const double* __restrict__ inv_scale, int p,
                                   double* __restrict__ colsum) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < p) colsum[j] += (double)colsum_q[j] * inv_scale[j];
}

// ------------------------------------------------------------------------------------------ multiply
template <int NB, int DLO, int DHI>
__global__ void __launch_bounds__(NTHREADS, 1)
gram_tc_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ Args a) {
    using C = Cfg<NB>;
    using PS = Pass<DLO, DHI>;
    static_assert(PS::NACC <= C::MAX_ACC, "accumulators of the pass must fit the 512 TMEM columns");
    constexpr int NSTAGE = C::NSTAGE, KB = C::KB, BOX_BYTES = C::BOX_BYTES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sb = (smem_u32(smem_raw) + 1023u) & ~1023u;      // SWIZZLE_128B atoms are 1024-byte aligned
    const uint32_t bar0 = sb + (uint32_t)C::SM_OFF_BAR;
    auto BAR_FULL = [&](int s) { return bar0 + 8u * (uint32_t)s; };
    auto BAR_EMPTY = [&](int s) { return bar0 + 8u * (uint32_t)(NSTAGE + s); };
    const uint32_t bar_acc = bar0 + 8u * (uint32_t)(2 * NSTAGE);
    const uint32_t tmem_slot = bar0 + 8u * (uint32_t)(2 * NSTAGE + 1);
    const WaitCtx wctx{tmem_slot + 4u, a.status, 1u << 22};

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // work unit: (upper-triangle tile, K split)
    const int tile = blockIdx.x % a.ntiles, split = blockIdx.x / a.ntiles;
    int ti, tj;
    tile_coords<NB>(tile, a.nt, ti, tj);
    const bool diag = NB == 1 && ti == tj;                      // B is A: load the planes once
    const int st0 = split * a.stages_per_split;
    const int nst = min(a.nstages, st0 + a.stages_per_split) - st0;
    if (nst <= 0) return;                                        // uniform over the CTA

    if (tid == 0) {
        asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(wctx.abort_addr), "r"(0u) : "memory");
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(BAR_FULL(s), 1); mbar_init(BAR_EMPTY(s), 1); }
        mbar_init(bar_acc, 1);
        fence_mbar_init();
    }
    if (warp == 0 && lane == 0) tma_prefetch_desc(&tmap);
    if (warp == 1) tmem_alloc(tmem_slot, 512u);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot) : "memory");

    if (warp == 0) {
        // ===== TMA producer: digit planes of the tile's row block (A) and column block(s) (B) ================
        if (lane == 0) {
            const uint64_t pol = policy_evict_last();          // every box is read by up to nt tiles: keep it in L2
            const uint32_t bytes = (uint32_t)((diag ? 1 : 1 + NB) * PS::NLOAD * BOX_BYTES);
            int slot = 0;
            uint32_t ph = 0;
            for (int s = 0; s < nst; ++s) {
                if (!mbar_wait_bounded(BAR_EMPTY(slot), ph ^ 1u, wctx, 11u)) break;
                mbar_expect_tx(BAR_FULL(slot), bytes);
                const uint32_t dst = sb + (uint32_t)slot * (uint32_t)C::STAGE_BYTES;
                const int row = (st0 + s) * KB;
#pragma unroll
                for (int k = PS::KMIN; k <= PS::KMAX; ++k)
                    tma_load_3d(dst + (uint32_t)(k * BOX_BYTES), &tmap, ti * TILE, row, k, BAR_FULL(slot), pol);
                if (!diag)
#pragma unroll
                    for (int k = PS::KMIN; k <= PS::KMAX; ++k)
#pragma unroll
                        for (int cb = 0; cb < NB; ++cb)        // columns beyond p_pad are zero-filled by the TMA
                            tma_load_3d(dst + (uint32_t)((S + k * NB + cb) * BOX_BYTES), &tmap, (tj + cb) * TILE, row, k,
                                        BAR_FULL(slot), pol);
                if (++slot == NSTAGE) { slot = 0; ph ^= 1u; }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issue: warp-uniform loop, one elected lane issues ============================================
        constexpr uint32_t IDESC = make_idesc_i8(TILE, TILE * NB, true, true);
        const uint32_t b_off = diag ? 0u : (uint32_t)(S * BOX_BYTES);
        int slot = 0;
        uint32_t ph = 0;
        bool fine = true;
        for (int s = 0; s < nst; ++s) {
            if (!__all_sync(0xffffffffu, mbar_wait_bounded(BAR_FULL(slot), ph, wctx, 12u))) { fine = false; break; }
            tc_fence_after();
            const uint32_t base = sb + (uint32_t)slot * (uint32_t)C::STAGE_BYTES;
            // MN-major SW128 operands: 128 columns = one 128-byte atom row, 8-row groups 1024 B apart (SBO), one K = 32
            // MMA covers four groups; the B operand of an N = 256 MMA is two atoms BOX_BYTES apart (LBO).  Descriptors
            // of the products are constant offsets (in 16-byte units, low word) from the stage's two base descriptors.
            const uint64_t dA = make_desc_sw128(base, BOX_BYTES, 1024);
            const uint64_t dB = make_desc_sw128(base + b_off, BOX_BYTES, 1024);
            const uint32_t not_first = s > 0 ? 1u : 0u;
#pragma unroll
            for (int ks = 0; ks < KB / KSTEP; ++ks) {              // the next k-step starts 32 rows = 4096 B further
#pragma unroll
                for (int d = DLO; d <= DHI; ++d) {
#pragma unroll
                    for (int k = 0; k < S; ++k) {
                        const int l = d - k;
                        if (l < 0 || l >= S) continue;
                        const bool opens = ks == 0 && k == (d > S - 1 ? d - (S - 1) : 0);    // first product into this accumulator
                        tc_mma_ss_i8(tmem + (uint32_t)((d - DLO) * TILE * NB), dA + (uint64_t)((k * BOX_BYTES + ks * 4096) >> 4),
                                     dB + (uint64_t)((l * NB * BOX_BYTES + ks * 4096) >> 4), IDESC, opens ? not_first : 1u);
                    }
                }
            }
            tc_commit_elect(BAR_EMPTY(slot));                   // the slot is free once these MMAs have read it
            if (++slot == NSTAGE) { slot = 0; ph ^= 1u; }
        }
        if (fine) tc_commit_elect(bar_acc);         // an abandoned pipeline publishes nothing: the flush marks the tile NaN
    } else {
        // ===== flush: int32 accumulators -> int64 -> fp64 -> += into this CTA's own partial tile ====================
        const int q = warp & 3;                                  // TMEM lane quadrant this warp may access
        const bool ok = __all_sync(0xffffffffu, mbar_wait_bounded(bar_acc, 0, wctx, 13u));      // false: the pipeline gave up -> the tile becomes NaN
        tc_fence_after();
        const int r = 32 * q + lane;                             // row of the tile
        const double w_i = ldexp(a.inv_scale[ti * TILE + r], 8 * PS::DMIN);
        const uint32_t lane_addr = tmem + ((uint32_t)(32 * q) << 16);
        // partial tile layout [c16][row][16]: a warp's 32 rows x 128 B are one contiguous 4 KB run per column group
        double* tile_base = a.part + ((int64_t)split * a.ntiles + tile) * (int64_t)(TILE * TILE * NB);
        const int ncol_ok = a.nt * TILE - tj * TILE;             // columns of this tile inside the padded matrix
        for (int c16 = 0; c16 < NB * TILE / 16; ++c16) {
            if (16 * c16 >= ncol_ok) break;                      // second column block of an odd edge: nothing there
            long long tsum[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) tsum[e] = 0;
#pragma unroll
            for (int ac = 0; ac < PS::NACC; ++ac) {
                uint32_t v[16];
                tc_ld16(lane_addr + (uint32_t)(ac * TILE * NB + 16 * c16), v);
                tc_wait_ld();
                tc_pin16(v);
#pragma unroll
                for (int e = 0; e < 16; ++e) tsum[e] += (long long)(int)v[e] << (8 * ac);
            }
            const double* isc = a.inv_scale + tj * TILE + 16 * c16;
            double2* dst = reinterpret_cast<double2*>(tile_base + ((int64_t)c16 * TILE + r) * 16);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                double2 cur = dst[e];
                cur.x += (double)tsum[2 * e] * (w_i * isc[2 * e]);
                cur.y += (double)tsum[2 * e + 1] * (w_i * isc[2 * e + 1]);
                if (!ok) cur.x = cur.y = __longlong_as_double(0x7ff8000000000000ll);
                dst[e] = cur;
            }
        }
        tc_fence_before();
    }
    if (lds32_volatile(wctx.abort_addr)) {     // in-flight TMA boxes / MMAs of the abandoned pipeline: let them land
        for (int i = 0; i < 100; ++i) __nanosleep(2000);
    }
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512u);
}

// ------------------------------------------------------------------------------------------ multiply, CTA pair
// cta_group::2: the two CTAs of a cluster share one MMA of M = 256 x N = 128.  Each CTA owns 128 rows of the pair tile
// (its own A planes, its own 4 x 128 accumulator columns) and holds only HALF of the B planes (64 columns, SWIZZLE_64B
// boxes), so per MMA it reads 4 KB of A + 2 KB of B from shared memory instead of 8 KB, and the TMA writes 48 KB per
// stage instead of 64 KB: the shared-memory pipe (128 B/clk) no longer caps the MMA rate of the 128 x 128 single-CTA
// kernel (~65 % measured).  Leader = cluster rank 0: it alone issues the MMAs; both CTAs' TMA loads signal ITS full
// barrier; its commits free the slot in both CTAs (multicast) and publish the accumulators to both.
struct Cfg2 {
    static constexpr int KB = 64;
    static constexpr int A_BOX = TILE * KB;                  // 8192: 128 columns x 64 rows, SWIZZLE_128B
    static constexpr int B_BOX = (TILE / 2) * KB;            // 4096:  64 columns x 64 rows, SWIZZLE_64B
    static constexpr int STAGE_BYTES = S * (A_BOX + B_BOX);  // 49152
    static constexpr int NSTAGE = 4;
    static constexpr size_t SM_OFF_BAR = (size_t)NSTAGE * STAGE_BYTES;
    static constexpr size_t SM_TOTAL = SM_OFF_BAR + 128 + 1024;
};

// pair tile -> (row pair-block tp: row blocks 2 tp, 2 tp + 1; column block tj >= 2 tp)
__host__ __device__ inline void pair_coords(int tile, int nt, int& tp, int& tj) {
    int t = tile; tp = 0;
    while (t >= nt - 2 * tp) { t -= nt - 2 * tp; ++tp; }
    tj = 2 * tp + t;
}
inline int pair_count(int nt) {
    int c = 0;
    for (int tp = 0; 2 * tp < nt; ++tp) c += nt - 2 * tp;
    return c;
}

template <int DLO, int DHI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
gram_tc2_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB, const __grid_constant__ Args a) {
    using C = Cfg2;
    using PS = Pass<DLO, DHI>;
    static_assert(PS::NACC <= 4, "accumulators of the pass must fit the 512 TMEM columns");
    constexpr int NSTAGE = C::NSTAGE, KB = C::KB;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sb = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar0 = sb + (uint32_t)C::SM_OFF_BAR;
    auto BAR_FULL = [&](int s) { return bar0 + 8u * (uint32_t)s; };
    auto BAR_EMPTY = [&](int s) { return bar0 + 8u * (uint32_t)(NSTAGE + s); };
    const uint32_t bar_acc = bar0 + 8u * (uint32_t)(2 * NSTAGE);
    const uint32_t tmem_slot = bar0 + 8u * (uint32_t)(2 * NSTAGE + 1);
    const WaitCtx wctx{tmem_slot + 4u, a.status, 1u << 22};

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rank = (int)cluster_ctarank();
    const int unit = blockIdx.x >> 1;
    const int tile = unit % a.ntiles, split = unit / a.ntiles;
    int tp, tj;
    pair_coords(tile, a.nt, tp, tj);
    const int ti = 2 * tp + rank;                                // this CTA's row block
    const int st0 = split * a.stages_per_split;
    const int nst = min(a.nstages, st0 + a.stages_per_split) - st0;
    if (nst <= 0) return;                                        // uniform over the pair

    if (tid == 0) {
        asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(wctx.abort_addr), "r"(0u) : "memory");
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(BAR_FULL(s), 1); mbar_init(BAR_EMPTY(s), 1); }
        mbar_init(bar_acc, 1);
        fence_mbar_init();
    }
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmapA); tma_prefetch_desc(&tmapB); }
    if (warp == 1) tmem_alloc_2sm(tmem_slot, 512u);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                          // the peer's barriers exist before anything signals them
    tc_fence_after();
    uint32_t tmem;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot) : "memory");

    if (warp == 0) {
        // ===== TMA producer (each CTA loads its own A planes and its half of the B planes) =====================
        if (lane == 0) {
            const uint64_t pol = policy_evict_last();
            const uint32_t bytes_pair = (uint32_t)(2 * PS::NLOAD * (C::A_BOX + C::B_BOX));
            int slot = 0;
            uint32_t ph = 0;
            for (int s = 0; s < nst; ++s) {
                if (!mbar_wait_bounded(BAR_EMPTY(slot), ph ^ 1u, wctx, 21u)) break;
                if (rank == 0) mbar_expect_tx(BAR_FULL(slot), bytes_pair);
                const uint32_t dst = sb + (uint32_t)slot * (uint32_t)C::STAGE_BYTES;
                const int row = (st0 + s) * KB;
#pragma unroll
                for (int k = PS::KMIN; k <= PS::KMAX; ++k) {
                    tma_load_3d_2sm(dst + (uint32_t)(k * C::A_BOX), &tmapA, ti * TILE, row, k, BAR_FULL(slot), pol);
                    tma_load_3d_2sm(dst + (uint32_t)(S * C::A_BOX + k * C::B_BOX), &tmapB, tj * TILE + (TILE / 2) * rank, row, k,
                                    BAR_FULL(slot), pol);
                }
                if (++slot == NSTAGE) { slot = 0; ph ^= 1u; }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issue: the leader only ==========================================================================
        if (rank == 0) {
            constexpr uint32_t IDESC = make_idesc_i8(2 * TILE, TILE, true, true);
            int slot = 0;
            uint32_t ph = 0;
            bool fine = true;
            for (int s = 0; s < nst; ++s) {
                if (!__all_sync(0xffffffffu, mbar_wait_bounded(BAR_FULL(slot), ph, wctx, 22u))) { fine = false; break; }
                tc_fence_after();
                const uint32_t base = sb + (uint32_t)slot * (uint32_t)C::STAGE_BYTES;
                // A: MN-major SW128 (128 rows of G = 128 bytes), 8-row K groups 1024 B apart, k-step = 4096 B
                // B: this CTA's 64 columns, MN-major SW64 (64-byte rows), 8-row K groups 512 B apart, k-step = 2048 B
                const uint64_t dA = make_desc_sw128(base, C::A_BOX, 1024);
                const uint64_t dB = make_desc_sw64(base + (uint32_t)(S * C::A_BOX), C::B_BOX, 512);
                const uint32_t not_first = s > 0 ? 1u : 0u;
#pragma unroll
                for (int ks = 0; ks < KB / KSTEP; ++ks) {
#pragma unroll
                    for (int d = DLO; d <= DHI; ++d) {
#pragma unroll
                        for (int k = 0; k < S; ++k) {
                            const int l = d - k;
                            if (l < 0 || l >= S) continue;
                            const bool opens = ks == 0 && k == (d > S - 1 ? d - (S - 1) : 0);
                            tc_mma_ss_i8_2sm(tmem + (uint32_t)((d - DLO) * TILE), dA + (uint64_t)((k * C::A_BOX + ks * 4096) >> 4),
                                             dB + (uint64_t)((l * C::B_BOX + ks * 2048) >> 4), IDESC, opens ? not_first : 1u);
                        }
                    }
                }
                tc_commit_2sm_elect(BAR_EMPTY(slot));           // frees the slot in BOTH CTAs
                if (++slot == NSTAGE) { slot = 0; ph ^= 1u; }
            }
            if (fine) tc_commit_2sm_elect(bar_acc);             // accumulators complete: both CTAs flush their half
        }
        __syncwarp();
    } else {
        // ===== flush: this CTA's 128 rows of the pair tile =========================================================
        const int q = warp & 3;
        const bool ok = __all_sync(0xffffffffu, mbar_wait_bounded(bar_acc, 0, wctx, 23u));
        tc_fence_after();
        const int r = 32 * q + lane;
        const double w_i = ldexp(a.inv_scale[ti * TILE + r], 8 * PS::DMIN);
        const uint32_t lane_addr = tmem + ((uint32_t)(32 * q) << 16);
        double* tile_base = a.part + (((int64_t)split * a.ntiles + tile) * 2 + rank) * (int64_t)(TILE * TILE);
        for (int c16 = 0; c16 < TILE / 16; ++c16) {
            long long tsum[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) tsum[e] = 0;
#pragma unroll
            for (int ac = 0; ac < PS::NACC; ++ac) {
                uint32_t v[16];
                tc_ld16(lane_addr + (uint32_t)(ac * TILE + 16 * c16), v);
                tc_wait_ld();
                tc_pin16(v);
#pragma unroll
                for (int e = 0; e < 16; ++e) tsum[e] += (long long)(int)v[e] << (8 * ac);
            }
            const double* isc = a.inv_scale + tj * TILE + 16 * c16;
            double2* dst = reinterpret_cast<double2*>(tile_base + ((int64_t)c16 * TILE + r) * 16);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                double2 cur = dst[e];
                cur.x += (double)tsum[2 * e] * (w_i * isc[2 * e]);
                cur.y += (double)tsum[2 * e + 1] * (w_i * isc[2 * e + 1]);
                if (!ok) cur.x = cur.y = __longlong_as_double(0x7ff8000000000000ll);
                dst[e] = cur;
            }
        }
        tc_fence_before();
    }
    if (lds32_volatile(wctx.abort_addr)) {
        for (int i = 0; i < 100; ++i) __nanosleep(2000);
    }
    __syncthreads();
    cluster_sync_all();                                          // nobody leaves while the peer may still read its shared memory
    if (warp == 1) tmem_dealloc_2sm(tmem, 512u);
}

// fold for the pair layout: partial tile (pair tile, rank) covers row block 2 tp + rank, column block tj
__global__ void __launch_bounds__(256)
gram_fold2_kernel(const double* __restrict__ part, int splits, int nt, int ntiles, int p, const double* __restrict__ colsumsq_q,
                  const double* __restrict__ inv_scale, double* __restrict__ G) {
    const int tile = blockIdx.x >> 1, rank = blockIdx.x & 1;
    int tp, tj;
    pair_coords(tile, nt, tp, tj);
    const int ti = 2 * tp + rank;
    for (int idx = threadIdx.x; idx < TILE * TILE; idx += blockDim.x) {
        const int r = idx >> 7, c = idx & 127;
        const int i = ti * TILE + r, j = tj * TILE + c;
        if (i >= p || j >= p || j < i) continue;
        double acc;
        if (i == j) {
            acc = colsumsq_q[i] * inv_scale[i] * inv_scale[i];
        } else {
            acc = 0.0;
            const int64_t off = ((int64_t)(c >> 4) * TILE + r) * 16 + (c & 15);
            for (int sp = 0; sp < splits; ++sp) acc += part[(((int64_t)sp * ntiles + tile) * 2 + rank) * (int64_t)(TILE * TILE) + off];
        }
        G[(int64_t)i * p + j] += acc;
    }
}

static int make_planes_map(CUtensorMap* map, const uint8_t* planes, int p_pad, int rows_cap, int KB, int box_cols = TILE) {
    EncodeTiledFn encode = tensor_map_encoder();
    if (!encode) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return (int)cudaErrorNotSupported; }
    const cuuint64_t dims[3] = {(cuuint64_t)p_pad, (cuuint64_t)rows_cap, (cuuint64_t)S};
    const cuuint64_t strides[2] = {(cuuint64_t)p_pad, (cuuint64_t)p_pad * (cuuint64_t)rows_cap};
    const cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)KB, 1};  // box_cols columns x KB rows of one plane
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(planes), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols == TILE ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return (int)cudaErrorInvalidValue; }
    return 0;
}

// K splits of a chunk: every (tile, split) is one CTA with its own partial tile.  Cost model in cycles: 64 NB cycles
// per MMA, ~15 k cycles of set-up + flush per CTA; whole waves of `sms` CTAs.
static int choose_splits(int ntiles, int nstages, int sms, int nprod, int nb) {
    int max_splits = ntiles >= sms ? 1 : sms / ntiles;
    if (max_splits > nstages) max_splits = nstages;
    if (max_splits < 1) max_splits = 1;
    int best_sp = 1;
    double best = 1e300;
    for (int sp = 1; sp <= max_splits; ++sp) {
        const int64_t waves = ((int64_t)ntiles * sp + sms - 1) / sms;
        const double cost = (double)waves * ((double)((nstages + sp - 1) / sp) * 64.0 * nb * nprod * (nb == 1 ? 2 : 1) + 15000.0);
        if (cost < best * 0.999) { best = cost; best_sp = sp; }
    }
    return best_sp;
}

// kernel variant: 1 = 128 x 128 tile per CTA, 3 = 256 x 128 tile per CTA PAIR (cta_group::2, the default: ~10 % faster).
// OCM_GRAM_NB=1 picks the single-CTA kernel for A/B runs.  (A 128 x 256 single-CTA variant, template NB = 2 with two
// weights per pass, measured 1.7x SLOWER -- its low-weight passes load as much as the main pass for a third of the MMAs --
// and is not instantiated.)
static int column_blocks() {
    static const int nb = [] {
        const char* e = getenv("OCM_GRAM_NB");
        if (e && e[0] == '3') return 3;
        if (e && e[0] == '1') return 1;
        return OCM_GRAM_DEFAULT_MODE;
    }();
    return nb;
}

struct Layout {
    int nb, p_pad, rows_cap, nt, ntiles, max_splits;
    size_t off_scale, off_inv, off_colsum, off_colsumsq, off_flags, off_part, off_planes, plane_buf_bytes, total;
};
static Layout workspace_layout(int64_t n, int p) {
    Layout L;
    L.nb = column_blocks();
    L.p_pad = (p + TILE - 1) / TILE * TILE;
    L.nt = L.p_pad / TILE;
    L.ntiles = L.nb == 1 ? tile_count<1>(L.nt) : pair_count(L.nt);   // mode 3: PAIR tiles
    int64_t rc = (n + QROWS - 1) / QROWS * QROWS;
    if (rc > MAX_CHAIN_ROWS) rc = MAX_CHAIN_ROWS;
    if (rc < QROWS) rc = QROWS;
    L.rows_cap = (int)rc;
    const int sms = sm_count();
    const int ctas_per_tile = L.nb == 3 ? 2 : 1;
    L.max_splits = L.ntiles * ctas_per_tile >= sms ? 1 : sms / (L.ntiles * ctas_per_tile);
    size_t o = 0;
    L.off_scale = o; o += sizeof(double) * L.p_pad;
    L.off_inv = o; o += sizeof(double) * (L.p_pad + TILE);      // + one block: the padded second column block of an odd edge
    L.off_colsum = o; o += sizeof(long long) * L.p_pad;
    L.off_colsumsq = o; o += sizeof(double) * L.p_pad;
    L.off_flags = o; o += 64;
    o = (o + 1023) / 1024 * 1024;
    L.off_part = o; o += sizeof(double) * (size_t)L.max_splits * L.ntiles * TILE * TILE * 2;      // 2: column blocks (mode 2) or CTAs (mode 3)
    o = (o + 1023) / 1024 * 1024;
    L.plane_buf_bytes = (size_t)S * L.rows_cap * L.p_pad;
    const int nbuf = n > L.rows_cap ? 2 : 1;                 // double-buffered when there is more than one chunk
    L.off_planes = o; o += (size_t)nbuf * L.plane_buf_bytes;
    L.total = o;
    return L;
}

// side stream + events for running the quantiser of chunk c+1 under the MMAs of chunk c (one set per device)
struct Side {
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, q_done[2] = {nullptr, nullptr}, g_done[2] = {nullptr, nullptr};
    std::mutex mu;
};
static Side* side_of_device(int dev) {
    static Side sides[64];
    if (dev < 0 || dev >= 64) return nullptr;
    Side* sd = &sides[dev];
    if (!sd->stream) {
        if (cudaStreamCreateWithFlags(&sd->stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        cudaEventCreateWithFlags(&sd->fork, cudaEventDisableTiming);
        for (int b = 0; b < 2; ++b) {
            cudaEventCreateWithFlags(&sd->q_done[b], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&sd->g_done[b], cudaEventDisableTiming);
        }
    }
    return sd;
}

template <int NB, int DLO, int DHI>
static int launch_gram(const CUtensorMap& tmap, const Args& a, int ctas, cudaStream_t stream) {
    auto kern = gram_tc_kernel<NB, DLO, DHI>;
    OCM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<NB>::SM_TOTAL));
    kern<<<(unsigned)ctas, NTHREADS, Cfg<NB>::SM_TOTAL, stream>>>(tmap, a);
    return check_launch("gram_tc_kernel");
}
template <int DLO, int DHI>
static int launch_gram_pair(const CUtensorMap& tmapA, const CUtensorMap& tmapB, const Args& a, int ctas, cudaStream_t stream) {
    auto kern = gram_tc2_kernel<DLO, DHI>;
    OCM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg2::SM_TOTAL));
    kern<<<(unsigned)ctas, NTHREADS, Cfg2::SM_TOTAL, stream>>>(tmapA, tmapB, a);
    return check_launch("gram_tc2_kernel");
}

// one chunk: every pass of the chosen kernel variant (mode 1: 128 x 128 tiles, 2: 128 x 256 tiles, 3: CTA pairs)
static int launch_passes(int mode, int products, const CUtensorMap& tmap, const CUtensorMap& tmapB, const Args& a, int ctas,
                         cudaStream_t st) {
    int rc = 0;
    if (mode == 1) {
        rc = launch_gram<1, 3, 6>(tmap, a, ctas, st);
        if (!rc && products == 13) rc = launch_gram<1, 2, 2>(tmap, a, ctas, st);
        if (!rc && products == 16) rc = launch_gram<1, 0, 2>(tmap, a, ctas, st);
    } else {
        rc = launch_gram_pair<3, 6>(tmap, tmapB, a, ctas, st);
        if (!rc && products == 13) rc = launch_gram_pair<2, 2>(tmap, tmapB, a, ctas, st);
        if (!rc && products == 16) rc = launch_gram_pair<0, 2>(tmap, tmapB, a, ctas, st);
    }
    return rc;
}

template <typename T>
int gram_tc_impl(const T* X, int64_t ldx, const int64_t* row_idx, int64_t n, int32_t p, const double* shift,
                 const float* colmin, const float* colmax, int32_t products, void* work, size_t work_bytes, double* G,
                 double* colsum, int64_t* flags, void* stream_) {
    if (n == 0) return 0;
    OCM_REQUIRE(X && shift && colmin && colmax && work && G && n > 0 && p > 0 && ldx >= p, OCM_EINVAL, "bad shape or NULL buffer");
    OCM_REQUIRE(products == 10 || products == 13 || products == 16, OCM_EINVAL, "products must be 10, 13 or 16");
    OCM_REQUIRE(p <= 128 * 181, OCM_ELIMIT, "p too large for the tile index");
    {
        const int st = tc_status_report("ocm_gram_tc");       // an EARLIER tensor-core launch gave up: say so first
        if (st) return st;
    }
    const Layout L = workspace_layout(n, p);
    OCM_REQUIRE(work_bytes >= L.total, OCM_EINVAL, "workspace too small (see ocm_gram_tc_workspace_bytes)");
    OCM_REQUIRE((reinterpret_cast<uintptr_t>(work) & 1023) == 0, OCM_EALIGN, "workspace must be 1024-byte aligned");
    cudaStream_t stream = (cudaStream_t)stream_;
    uint8_t* wb = static_cast<uint8_t*>(work);
    double* scale = reinterpret_cast<double*>(wb + L.off_scale);
    double* inv_scale = reinterpret_cast<double*>(wb + L.off_inv);
    long long* colsum_q = reinterpret_cast<long long*>(wb + L.off_colsum);
    double* colsumsq_q = reinterpret_cast<double*>(wb + L.off_colsumsq);
    unsigned long long* wflags = reinterpret_cast<unsigned long long*>(wb + L.off_flags);
    double* part = reinterpret_cast<double*>(wb + L.off_part);
    const int64_t plane_stride = (int64_t)L.rows_cap * L.p_pad;
    const int sms = sm_count();
    const int64_t nchunks = (n + L.rows_cap - 1) / L.rows_cap;

    const int first_pass_products = Pass<3, 6>::count();

    // the K split is fixed for the call (the partial buffer is indexed by it); the last, shorter chunk just has fewer stages
    const int KB = 64;
    const int ctas_per_tile = L.nb == 3 ? 2 : 1;
    const int nstages_full = (int)(((n < L.rows_cap ? n : L.rows_cap) + KB - 1) / KB);
    int splits = choose_splits(L.ntiles * ctas_per_tile, nstages_full, sms, first_pass_products, L.nb == 2 ? 2 : 1);
    if (splits > L.max_splits) splits = L.max_splits;

    OCM_CUDA(cudaMemsetAsync(wflags, 0, 64, stream));
    OCM_CUDA(cudaMemsetAsync(part, 0, sizeof(double) * (size_t)splits * L.ntiles * TILE * TILE * 2, stream));
    OCM_CUDA(cudaMemsetAsync(inv_scale + L.p_pad, 0, sizeof(double) * TILE, stream));
    gram_scale_kernel<<<(L.p_pad + 255) / 256, 256, 0, stream>>>(colmin, colmax, shift, p, L.p_pad, scale, inv_scale, colsum_q, colsumsq_q);
    int rc = check_launch("gram_scale_kernel");
    if (rc) return rc;

    CUtensorMap tmap[2], tmapB[2];
    for (int b = 0; b < (nchunks > 1 ? 2 : 1); ++b) {
        rc = make_planes_map(&tmap[b], wb + L.off_planes + (size_t)b * L.plane_buf_bytes, L.p_pad, L.rows_cap, KB);
        if (rc) return rc;
        if (L.nb == 3) {
            rc = make_planes_map(&tmapB[b], wb + L.off_planes + (size_t)b * L.plane_buf_bytes, L.p_pad, L.rows_cap, KB, TILE / 2);
            if (rc) return rc;
        }
    }
    const int vec_ok = (sizeof(T) == 4 && ldx % 4 == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0) ? 1 : 0;

    // several chunks: quantise chunk c+1 on the side stream while the MMAs of chunk c run on the caller's stream
    int dev = 0;
    OCM_CUDA(cudaGetDevice(&dev));
    Side* sd = nchunks > 1 ? side_of_device(dev) : nullptr;
    std::unique_lock<std::mutex> lock;
    cudaStream_t qstream = stream;
    if (sd) {
        lock = std::unique_lock<std::mutex>(sd->mu);
        qstream = sd->stream;
        OCM_CUDA(cudaEventRecord(sd->fork, stream));
        OCM_CUDA(cudaStreamWaitEvent(qstream, sd->fork, 0));
    }
    for (int64_t c = 0; c < nchunks; ++c) {
        const int b = sd ? (int)(c & 1) : 0;
        const int64_t r0 = c * L.rows_cap;
        const int rows = (int)((n - r0 < L.rows_cap) ? n - r0 : L.rows_cap);
        const int rows_pad = (rows + QROWS - 1) / QROWS * QROWS;
        uint8_t* planes = wb + L.off_planes + (size_t)b * L.plane_buf_bytes;
        const int rpb = 256;
        dim3 qgrid(L.nt, (rows_pad + rpb - 1) / rpb);
        if (sd && c >= 2) OCM_CUDA(cudaStreamWaitEvent(qstream, sd->g_done[b], 0));     // the buffer's previous MMAs are done
        gram_quantize_kernel<T><<<qgrid, 256, 0, qstream>>>(X, ldx, row_idx, r0, rows, rows_pad, p, L.p_pad, shift, scale,
                                                            planes, plane_stride, colsum_q, colsumsq_q, wflags, rpb, vec_ok);
        rc = check_launch("gram_quantize_kernel");
        if (rc) return rc;
        if (sd) {
            OCM_CUDA(cudaEventRecord(sd->q_done[b], qstream));
            OCM_CUDA(cudaStreamWaitEvent(stream, sd->q_done[b], 0));
        }
        Args a{};
        a.p = p; a.nt = L.nt; a.ntiles = L.ntiles; a.nstages = (rows + KB - 1) / KB;
        a.stages_per_split = (nstages_full + splits - 1) / splits;
        a.inv_scale = inv_scale; a.part = part; a.status = tc_status_device_ptr();
        rc = launch_passes(L.nb, products, tmap[b], tmapB[b], a, L.ntiles * ctas_per_tile * splits, stream);
        if (rc) return rc;
        if (sd) OCM_CUDA(cudaEventRecord(sd->g_done[b], stream));
    }
    if (L.nb == 3) gram_fold2_kernel<<<2 * L.ntiles, 256, 0, stream>>>(part, splits, L.nt, L.ntiles, p, colsumsq_q, inv_scale, G);
    else gram_fold_kernel<1><<<L.ntiles, 256, 0, stream>>>(part, splits, L.nt, L.ntiles, p, colsumsq_q, inv_scale, G);
    rc = check_launch("gram_fold_kernel");
    if (rc) return rc;
    if (colsum) {
        gram_colsum_kernel<<<(p + 255) / 256, 256, 0, stream>>>(colsum_q, inv_scale, p, colsum);
        rc = check_launch("gram_colsum_kernel");
        if (rc) return rc;
    }
    if (flags) OCM_CUDA(cudaMemcpyAsync(flags, wflags, 2 * sizeof(int64_t), cudaMemcpyDeviceToDevice, stream));
    return 0;
}

}  // namespace gtc
}  // namespace ocm

extern "C" int ocm_gram_tc_workspace_bytes(int64_t n, int32_t p, int64_t* bytes) {
    OCM_REQUIRE(bytes && n >= 0 && p > 0, OCM_EINVAL, "bad argument");
    *bytes = (int64_t)ocm::gtc::workspace_layout(n, p).total;
    return 0;
}
extern "C" int ocm_gram_tc_f32(const float* X, int64_t ldx, const int64_t* row_idx, int64_t n, int32_t p,
                               const double* shift, const float* colmin, const float* colmax, int32_t products, void* work,
                               int64_t work_bytes, double* G, double* colsum, int64_t* flags, void* stream) {
    return ocm::gtc::gram_tc_impl<float>(X, ldx, row_idx, n, p, shift, colmin, colmax, products, work, (size_t)work_bytes, G,
                                         colsum, flags, stream);
}
extern "C" int ocm_gram_tc_f64(const double* X, int64_t ldx, const int64_t* row_idx, int64_t n, int32_t p,
                               const double* shift, const float* colmin, const float* colmax, int32_t products, void* work,
                               int64_t work_bytes, double* G, double* colsum, int64_t* flags, void* stream) {
    return ocm::gtc::gram_tc_impl<double>(X, ldx, row_idx, n, p, shift, colmin, colmax, products, work, (size_t)work_bytes, G,
                                          colsum, flags, stream);
}
