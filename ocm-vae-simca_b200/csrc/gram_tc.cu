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
//                 added into the fp64 Gram with red.global.add.f64.
// The result is the exact Gram of the quantised data for `products` = 16; 13 drops the three products of weight
// < 256^2 (relative error <= 2^-37 of max_i max_j per row), 10 those below 256^3 (<= 2^-28).  Column sums of v are
// accumulated as int64 in the same pass, so cov = (G - N d d^T) / (N - 1) is the covariance OF THE QUANTISED ROWS to
// fp64 rounding.  Four accumulators of 128 columns fill the 512 TMEM columns, so 13 / 16 products run as a second
// launch over the low digit planes (weights 256^0..256^2).
#include "ocm_common.cuh"
#include "tc_ptx.cuh"

#include <stdlib.h>

namespace ocm {
namespace gtc {

using namespace ptx;

constexpr int S = 4;                        // int8 digit planes
constexpr int TILE = 128;                   // output tile edge = MMA M = MMA N
constexpr int KB = 64;                      // rows per pipeline stage (two K = 32 MMAs per product)
constexpr int NSTAGE = 3;
constexpr int BOX_BYTES = TILE * KB;        // 8192: one plane's box, rows of 128 B, SWIZZLE_128B
constexpr int STAGE_BYTES = 2 * S * BOX_BYTES;
constexpr int NTHREADS = 192;               // warp 0: TMA, warp 1: MMA issue + TMEM owner, warps 2..5: flush
constexpr int MAX_CHAIN_ROWS = 32704;       // 4 products x 128 x 128 x rows < 2^31, multiple of KB
constexpr int V_LIMIT = 0x5FFFFFFF;         // clamp of the fixed-point value (top digit stays <= 96)
constexpr size_t SM_OFF_BAR = (size_t)NSTAGE * STAGE_BYTES;
constexpr size_t SM_TOTAL = SM_OFF_BAR + 128 + 1024;      // + barriers + slack for the 1024-byte alignment

struct Products {
    int n;                  // number of (k, l) digit products in this launch
    int nacc;               // accumulators used (each TILE columns of TMEM)
    int nload;              // digit planes 0 .. nload-1 are loaded
    int dmin;               // weight of accumulator 0 is 256^dmin, accumulator a is 256^(dmin + a)
    uint8_t k[16], l[16], acc[16];
};

struct Args {
    int p;                  // columns of G
    int nt;                 // tiles per edge
    int nstages;            // stages in this chunk (rows_pad / KB)
    int stages_per_split;
    const double* inv_scale;    // [p_pad] 2^-sc
    double* G;              // [p][p] upper triangle accumulated into
    Products pr;
};

// ------------------------------------------------------------------------------------------ quantise
// One warp = one row x 128 columns per step (lane -> 4 consecutive columns), 8 warps walk the rows of the block's slab.
template <typename T>
__global__ void __launch_bounds__(256)
gram_quantize_kernel(const T* __restrict__ X, int64_t ldx, const int64_t* __restrict__ row_idx, int64_t r0, int rows,
                     int rows_pad, int p, int p_pad, const double* __restrict__ shift, const double* __restrict__ scale,
                     uint8_t* __restrict__ planes, int64_t plane_stride, long long* __restrict__ colsum_q,
                     unsigned long long* __restrict__ flags, int rows_per_block, int vec_ok) {
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
#pragma unroll
    for (int e = 0; e < 4; ++e) s_cs[warp][4 * lane + e] = cs[e];
    __syncthreads();
    if (threadIdx.x < TILE) {
        long long t = 0;
#pragma unroll
        for (int w8 = 0; w8 < 8; ++w8) t += s_cs[w8][threadIdx.x];
        const int c = blockIdx.x * TILE + threadIdx.x;
        if (c < p && t != 0) atomicAdd(reinterpret_cast<unsigned long long*>(colsum_q + c), (unsigned long long)t);
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
                                  double* __restrict__ inv_scale, long long* __restrict__ colsum_q) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= p_pad) return;
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

__global__ void gram_colsum_kernel(const long long* __restrict__ colsum_q, const double* __restrict__ inv_scale, int p,
                                   double* __restrict__ colsum) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < p) colsum[j] += (double)colsum_q[j] * inv_scale[j];
}

// ------------------------------------------------------------------------------------------ multiply
__global__ void __launch_bounds__(NTHREADS, 1)
gram_tc_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ Args a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sb = (smem_u32(smem_raw) + 1023u) & ~1023u;      // SWIZZLE_128B atoms are 1024-byte aligned
    const uint32_t bar0 = sb + (uint32_t)SM_OFF_BAR;
    auto BAR_FULL = [&](int s) { return bar0 + 8u * (uint32_t)s; };
    auto BAR_EMPTY = [&](int s) { return bar0 + 8u * (uint32_t)(NSTAGE + s); };
    const uint32_t bar_acc = bar0 + 8u * (uint32_t)(2 * NSTAGE);
    const uint32_t tmem_slot = bar0 + 8u * (uint32_t)(2 * NSTAGE + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // work unit: (upper-triangle tile, K split)
    const int ntiles = a.nt * (a.nt + 1) / 2;
    const int tile = blockIdx.x % ntiles, split = blockIdx.x / ntiles;
    int t = tile, ti = 0;
    while (t >= a.nt - ti) { t -= a.nt - ti; ++ti; }
    const int tj = ti + t;
    const bool diag = ti == tj;
    const int st0 = split * a.stages_per_split;
    const int nst = min(a.nstages, st0 + a.stages_per_split) - st0;
    if (nst <= 0) return;                                        // uniform over the CTA

    if (tid == 0) {
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
        // ===== TMA producer: digit planes of the tile's row block (A) and column block (B) ===================
        if (lane == 0) {
            const uint64_t pol = policy_evict_last();          // every box is read by up to nt tiles: keep it in L2
            const uint32_t bytes = (uint32_t)((diag ? 1 : 2) * a.pr.nload * BOX_BYTES);
            int slot = 0;
            uint32_t ph = 0;
            for (int s = 0; s < nst; ++s) {
                mbar_wait_or_trap(BAR_EMPTY(slot), ph ^ 1u);
                mbar_expect_tx(BAR_FULL(slot), bytes);
                const uint32_t dst = sb + (uint32_t)slot * (uint32_t)STAGE_BYTES;
                const int row = (st0 + s) * KB;
                for (int k = 0; k < a.pr.nload; ++k)
                    tma_load_3d(dst + (uint32_t)(k * BOX_BYTES), &tmap, ti * TILE, row, k, BAR_FULL(slot), pol);
                if (!diag)
                    for (int k = 0; k < a.pr.nload; ++k)
                        tma_load_3d(dst + (uint32_t)((S + k) * BOX_BYTES), &tmap, tj * TILE, row, k, BAR_FULL(slot), pol);
                if (++slot == NSTAGE) { slot = 0; ph ^= 1u; }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issue: warp-uniform loop, one elected lane issues ============================================
        constexpr uint32_t IDESC = make_idesc_i8(TILE, TILE, true, true);
        const uint32_t b_off = diag ? 0u : (uint32_t)(S * BOX_BYTES);
        uint32_t used = 0;
        int slot = 0;
        uint32_t ph = 0;
        for (int s = 0; s < nst; ++s) {
            mbar_wait_or_trap(BAR_FULL(slot), ph);
            tc_fence_after();
            const uint32_t base = sb + (uint32_t)slot * (uint32_t)STAGE_BYTES;
#pragma unroll
            for (int ks = 0; ks < KB / 32; ++ks) {
                // MN-major SW128 operand: 128 columns = one 128-byte atom row, 8-row groups 1024 B apart (SBO); one
                // K = 32 MMA covers four groups; the next k-step starts 4096 B further
                for (int m = 0; m < a.pr.n; ++m) {
                    const uint32_t acc = a.pr.acc[m];
                    const uint64_t da = make_desc_sw128(base + (uint32_t)a.pr.k[m] * BOX_BYTES + (uint32_t)ks * 4096u, BOX_BYTES, 1024);
                    const uint64_t db = make_desc_sw128(base + b_off + (uint32_t)a.pr.l[m] * BOX_BYTES + (uint32_t)ks * 4096u, BOX_BYTES, 1024);
                    tc_mma_ss_i8(tmem + acc * (uint32_t)TILE, da, db, IDESC, (used >> acc) & 1u);
                    used |= 1u << acc;
                }
            }
            tc_commit_elect(BAR_EMPTY(slot));                   // the slot is free once these MMAs have read it
            if (++slot == NSTAGE) { slot = 0; ph ^= 1u; }
        }
        tc_commit_elect(bar_acc);
    } else {
        // ===== flush: int32 accumulators -> int64 -> fp64 -> red.add into G ====================================
        const int q = warp & 3;                                  // TMEM lane quadrant this warp may access
        mbar_wait_or_trap(bar_acc, 0);
        tc_fence_after();
        const int i = ti * TILE + 32 * q + lane;
        const double w_i = i < a.p ? ldexp(a.inv_scale[i], 8 * a.pr.dmin) : 0.0;
        const uint32_t lane_addr = tmem + ((uint32_t)(32 * q) << 16);
        for (int c16 = 0; c16 < TILE / 16; ++c16) {
            long long tsum[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) tsum[e] = 0;
            for (int ac = 0; ac < a.pr.nacc; ++ac) {
                uint32_t v[16];
                tc_ld16(lane_addr + (uint32_t)(ac * TILE + 16 * c16), v);
                tc_wait_ld();
                tc_pin16(v);
#pragma unroll
                for (int e = 0; e < 16; ++e) tsum[e] += (long long)(int)v[e] << (8 * ac);
            }
            const int j0 = tj * TILE + 16 * c16;
            if (i < a.p) {
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const int j = j0 + e;
                    if (j < a.p && j >= i && tsum[e] != 0)
                        atomicAdd(a.G + (int64_t)i * a.p + j, (double)tsum[e] * (w_i * a.inv_scale[j]));
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512u);
}

static int make_planes_map(CUtensorMap* map, const uint8_t* planes, int p_pad, int rows_cap) {
    EncodeTiledFn encode = tensor_map_encoder();
    if (!encode) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return (int)cudaErrorNotSupported; }
    const cuuint64_t dims[3] = {(cuuint64_t)p_pad, (cuuint64_t)rows_cap, (cuuint64_t)S};
    const cuuint64_t strides[2] = {(cuuint64_t)p_pad, (cuuint64_t)p_pad * (cuuint64_t)rows_cap};
    const cuuint32_t box[3] = {(cuuint32_t)TILE, (cuuint32_t)KB, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(planes), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return (int)cudaErrorInvalidValue; }
    return 0;
}

// products with weight 256^d, d in [dlo, dhi], one accumulator per weight
static Products product_table(int dlo, int dhi) {
    Products pr{};
    pr.dmin = dlo;
    pr.nacc = dhi - dlo + 1;
    int kmax = 0;
    for (int d = dlo; d <= dhi; ++d)
        for (int k = 0; k < S; ++k) {
            const int l = d - k;
            if (l < 0 || l >= S) continue;
            pr.k[pr.n] = (uint8_t)k; pr.l[pr.n] = (uint8_t)l; pr.acc[pr.n] = (uint8_t)(d - dlo);
            ++pr.n;
            if (k > kmax) kmax = k;
            if (l > kmax) kmax = l;
        }
    pr.nload = kmax + 1;
    return pr;
}

struct Layout {
    int p_pad, rows_cap;
    size_t off_scale, off_inv, off_colsum, off_flags, off_planes, total;
};
static Layout workspace_layout(int64_t n, int p) {
    Layout L;
    L.p_pad = (p + TILE - 1) / TILE * TILE;
    int64_t rc = (n + KB - 1) / KB * KB;
    if (rc > MAX_CHAIN_ROWS) rc = MAX_CHAIN_ROWS;
    L.rows_cap = (int)rc;
    size_t o = 0;
    L.off_scale = o; o += sizeof(double) * L.p_pad;
    L.off_inv = o; o += sizeof(double) * L.p_pad;
    L.off_colsum = o; o += sizeof(long long) * L.p_pad;
    L.off_flags = o; o += 64;
    o = (o + 1023) / 1024 * 1024;
    L.off_planes = o; o += (size_t)S * L.rows_cap * L.p_pad;
    L.total = o;
    return L;
}

template <typename T>
int gram_tc_impl(const T* X, int64_t ldx, const int64_t* row_idx, int64_t n, int32_t p, const double* shift,
                 const float* colmin, const float* colmax, int32_t products, void* work, size_t work_bytes, double* G,
                 double* colsum, int64_t* flags, void* stream_) {
    if (n == 0) return 0;
    OCM_REQUIRE(X && shift && colmin && colmax && work && G && n > 0 && p > 0 && ldx >= p, OCM_EINVAL, "bad shape or NULL buffer");
    OCM_REQUIRE(products == 10 || products == 13 || products == 16, OCM_EINVAL, "products must be 10, 13 or 16");
    OCM_REQUIRE(p <= 128 * 181, OCM_ELIMIT, "p too large for the tile index");
    const Layout L = workspace_layout(n, p);
    OCM_REQUIRE(work_bytes >= L.total, OCM_EINVAL, "workspace too small (see ocm_gram_tc_workspace_bytes)");
    OCM_REQUIRE((reinterpret_cast<uintptr_t>(work) & 1023) == 0, OCM_EALIGN, "workspace must be 1024-byte aligned");
    cudaStream_t stream = (cudaStream_t)stream_;
    uint8_t* wb = static_cast<uint8_t*>(work);
    double* scale = reinterpret_cast<double*>(wb + L.off_scale);
    double* inv_scale = reinterpret_cast<double*>(wb + L.off_inv);
    long long* colsum_q = reinterpret_cast<long long*>(wb + L.off_colsum);
    unsigned long long* wflags = reinterpret_cast<unsigned long long*>(wb + L.off_flags);
    uint8_t* planes = wb + L.off_planes;
    const int64_t plane_stride = (int64_t)L.rows_cap * L.p_pad;

    OCM_CUDA(cudaMemsetAsync(wflags, 0, 64, stream));
    gram_scale_kernel<<<(L.p_pad + 255) / 256, 256, 0, stream>>>(colmin, colmax, shift, p, L.p_pad, scale, inv_scale, colsum_q);
    int rc = check_launch("gram_scale_kernel");
    if (rc) return rc;

    CUtensorMap tmap;
    rc = make_planes_map(&tmap, planes, L.p_pad, L.rows_cap);
    if (rc) return rc;
    auto kern = gram_tc_kernel;
    OCM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL));

    const int nt = L.p_pad / TILE, ntiles = nt * (nt + 1) / 2, sms = sm_count();
    const int vec_ok = (sizeof(T) == 4 && ldx % 4 == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0) ? 1 : 0;
    Products passes[2];
    int npass = 1;
    passes[0] = product_table(3, 6);
    if (products == 13) { passes[1] = product_table(2, 2); npass = 2; }
    if (products == 16) { passes[1] = product_table(0, 2); npass = 2; }

    for (int64_t r0 = 0; r0 < n; r0 += L.rows_cap) {
        const int rows = (int)((n - r0 < L.rows_cap) ? n - r0 : L.rows_cap);
        const int rows_pad = (rows + KB - 1) / KB * KB;
        const int rpb = 256;
        dim3 qgrid(nt, (rows_pad + rpb - 1) / rpb);
        gram_quantize_kernel<T><<<qgrid, 256, 0, stream>>>(X, ldx, row_idx, r0, rows, rows_pad, p, L.p_pad, shift, scale,
                                                           planes, plane_stride, colsum_q, wflags, rpb, vec_ok);
        rc = check_launch("gram_quantize_kernel");
        if (rc) return rc;
        const int nstages = rows_pad / KB;
        // K splits: fill the SMs with whole waves; every CTA flushes a 128 x 128 tile, so keep >= 16 stages each
        int splits = 1;
        {
            int max_splits = nstages / 16;
            if (max_splits < 1) max_splits = 1;
            double best = 0.0;
            for (int sp = 1; sp <= max_splits && sp <= 4 * sms; ++sp) {
                const int64_t ctas = (int64_t)ntiles * sp, waves = (ctas + sms - 1) / sms;
                const double eff = (double)ctas / (double)(waves * sms);
                if (eff > best + 0.01) { best = eff; splits = sp; }
            }
        }
        Args a{};
        a.p = p; a.nt = nt; a.nstages = nstages;
        a.stages_per_split = (nstages + splits - 1) / splits;
        splits = (nstages + a.stages_per_split - 1) / a.stages_per_split;
        a.inv_scale = inv_scale; a.G = G;
        for (int ps = 0; ps < npass; ++ps) {
            a.pr = passes[ps];
            kern<<<(unsigned)(ntiles * splits), NTHREADS, SM_TOTAL, stream>>>(tmap, a);
            rc = check_launch("gram_tc_kernel");
            if (rc) return rc;
        }
    }
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
