// Staged bulk scoring: inputs the tcgen05 kernels do not take directly are brought into their format chunk by chunk.
//
//  * fp64 spectra (numpy's default dtype: what utils/SIMCA.py:127 hands to sklearn's transform) are centred IN fp64 (and
//    SNV-normalised in fp64 when asked), rounded once to fp32 -- the same rounding point as the generic kernel's fp64
//    path, `(float)(x - mu)` -- and the chunk is scored by the bulk dispatcher with a zero mean.
//  * fp32 spectra with fused SNV at widths the tcgen05 kernel cannot fuse it (p > 1152: the cluster kernel) go through
//    the streaming SNV + centring kernel (preprocess.cu, mode OCM_PRE_SNV_CENTER) in the same way.
//
// 16p (fp64) / 12p (fp32) instead of 8p / 4p bytes of HBM traffic per spectrum, but at the stream rate instead of the
// generic kernel's ~10 % of it.  The scratch chunk (1 GB + at most 32 767 rows; OCM_STAGED_SCRATCH_MB, measured 12 % faster
// than 256 MB chunks: fewer partial waves of the persistent kernel) comes from a stream-ordered memory pool of the library's
// own, one per device, that keeps its memory between calls (the default pool hands it back at every synchronisation:
// re-mapping 256 MB per call cost more than the kernels).  Small / odd shapes and calls that want the score matrix stay
// on the generic kernel (score_simt.cu).
#include "ocm_common.cuh"

#include <mutex>

extern "C" int ocm_simca_score_simt_f64in(const double*, int64_t, int32_t, int64_t, const double*, const float*,
                                          const float*, const ocm_score_params*, float*, float*, uint8_t*, float*, void*);
extern "C" int ocm_simca_score_f32(const float*, int64_t, int32_t, int64_t, const float*, const float*, const float*,
                                   const ocm_score_params*, float*, float*, uint8_t*, float*, void*);
extern "C" int ocm_simca_score_tc_supported(int64_t n, int32_t p, int64_t ldx, int32_t A, int32_t snv, const void* X);
extern "C" int ocm_simca_score_simt_f32(const float*, int64_t, int32_t, int64_t, const float*, const float*, const float*,
                                        const ocm_score_params*, float*, float*, uint8_t*, float*, void*);

namespace ocm {
namespace staged {

__device__ __forceinline__ double2 ldg_stream_d2(const double* p) {
    double2 r;
    asm volatile("ld.global.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}

// One warp per spectrum; lanes own pairs of wavelengths (16-byte loads, 8-byte stores; every warp instruction moves
// 512 contiguous bytes in and 256 out).  SNV: mean, then the variance about that mean (population, eps on the standard
// deviation: the generic kernel's convention, score_simt.cu), both re-reads of the 8p-byte row hit L1/L2.
template <bool SNV>
__global__ void __launch_bounds__(256)
centre_f64_kernel(const double* __restrict__ X, int64_t n, int p, int64_t ldx, const double* __restrict__ mu,
                  double snv_eps, float* __restrict__ out, int64_t ldo) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int pairs = p >> 1;
    for (int64_t row = wid; row < n; row += nw) {
        const double* xr = X + row * ldx;
        float* o = out + row * ldo;
        double m = 0.0, rinv = 1.0;
        if (SNV) {
            double s = 0.0;
            for (int j = lane; j < pairs; j += 32) { const double2 v = *reinterpret_cast<const double2*>(xr + 2 * j); s += v.x + v.y; }
            m = warp_sum(s) / (double)p;
            s = 0.0;
            for (int j = lane; j < pairs; j += 32) {
                const double2 v = *reinterpret_cast<const double2*>(xr + 2 * j);
                const double a = v.x - m, b = v.y - m;
                s = fma(a, a, fma(b, b, s));
            }
            rinv = 1.0 / (sqrt(warp_sum(s) / (double)p) + snv_eps);
        }
#pragma unroll 4
        for (int j = lane; j < pairs; j += 32) {
            const double2 v = SNV ? *reinterpret_cast<const double2*>(xr + 2 * j) : ldg_stream_d2(xr + 2 * j);
            const double2 c = *reinterpret_cast<const double2*>(mu + 2 * j);
            float2 r;
            if (SNV) { r.x = (float)((v.x - m) * rinv - c.x); r.y = (float)((v.y - m) * rinv - c.y); }
            else     { r.x = (float)(v.x - c.x);              r.y = (float)(v.y - c.y); }
            *reinterpret_cast<float2*>(o + 2 * j) = r;
        }
    }
}

constexpr int64_t kMinBulkRows = 32768;               // the tcgen05 gate (ocm_simca_score_tc_supported)
constexpr int kNoScratch = -1000;                     // internal: no scratch chunk available, nothing has been launched
constexpr int kMaxDevices = 64;

int64_t scratch_bytes() {                             // fp32 scratch per chunk (+ up to 32 767 rows for the last one)
    const char* e = getenv("OCM_STAGED_SCRATCH_MB");           // read per call: the tests shrink it to span several chunks
    const long mb = e ? atol(e) : 0;
    return (int64_t)(mb >= 16 && mb <= 16384 ? mb : 1024) << 20;
}

cudaMemPool_t g_pool[kMaxDevices];
std::mutex g_pool_mutex;

// the library's own stream-ordered pool on the current device; never hands memory back by itself
int scratch_pool(cudaMemPool_t* out) {
    int dev = 0;
    OCM_CUDA(cudaGetDevice(&dev));
    OCM_REQUIRE(dev >= 0 && dev < kMaxDevices, OCM_ELIMIT, "device ordinal out of range");
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    if (!g_pool[dev]) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        cudaMemPool_t pool = nullptr;
        OCM_CUDA(cudaMemPoolCreate(&pool, &props));
        uint64_t keep = UINT64_MAX;
        OCM_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        g_pool[dev] = pool;
    }
    *out = g_pool[dev];
    return 0;
}

// Scores n rows chunk by chunk: `fill(first_row, rows, scratch)` launches the kernel that writes the fp32 rows [rows][p]
// of the chunk -- already centred (mu == nullptr: the bulk dispatcher then scores against a zero mean) or preprocessed
// but not centred (mu = the model mean).
template <typename Fill>
int score_chunks(int64_t n, int32_t p, const float* P, const float* Sinv, const ocm_score_params* hp,
                 float* T2, float* Q, uint8_t* accept, void* stream, Fill fill, const float* mu = nullptr) {
    cudaStream_t st = (cudaStream_t)stream;
    int64_t chunk = scratch_bytes() / (4 * (int64_t)p);
    if (chunk < kMinBulkRows) chunk = kMinBulkRows;
    const int64_t cap = (n < chunk + kMinBulkRows) ? n : chunk + kMinBulkRows;     // rows the scratch must hold
    const size_t zero_bytes = ((size_t)p * sizeof(float) + 255) & ~(size_t)255;
    cudaMemPool_t pool = nullptr;
    int rc = scratch_pool(&pool);
    if (rc) return kNoScratch;
    char* ws = nullptr;
    if (cudaMallocFromPoolAsync(reinterpret_cast<void**>(&ws), zero_bytes + (size_t)cap * p * sizeof(float), pool, st) != cudaSuccess) {
        cudaGetLastError();                             // no scratch to be had (memory, or a driver without pools):
        return kNoScratch;                              // the caller scores on the generic kernel instead
    }
    float* zero_mu = reinterpret_cast<float*>(ws);
    float* xc = reinterpret_cast<float*>(ws + zero_bytes);
    rc = (int)cudaMemsetAsync(zero_mu, 0, zero_bytes, st);
    if (rc) set_error("cudaMemsetAsync: %s", cudaGetErrorString((cudaError_t)rc));
    ocm_score_params prm = *hp;
    prm.snv = 0;                                        // applied by the staging kernel
    for (int64_t r0 = 0; r0 < n && rc == 0;) {
        const int64_t m = (n - r0 < chunk + kMinBulkRows) ? n - r0 : chunk;      // no tail below the bulk gate
        rc = fill(r0, m, xc);
        if (rc == 0)
            rc = ocm_simca_score_f32(xc, m, p, p, mu ? mu : zero_mu, P, Sinv, &prm, T2 ? T2 + r0 : nullptr, Q ? Q + r0 : nullptr,
                                     accept ? accept + r0 : nullptr, nullptr, stream);
        r0 += m;
    }
    const cudaError_t e_free = cudaFreeAsync(ws, st);
    if (rc == 0 && e_free != cudaSuccess) { set_error("cudaFreeAsync: %s", cudaGetErrorString(e_free)); rc = (int)e_free; }
    return rc;
}

}  // namespace staged
}  // namespace ocm

// gives the scratch pools' memory back to the driver (called by ocm_simca_host_release)
extern "C" int ocm_simca_staged_release(void) {
    using namespace ocm::staged;
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    for (int d = 0; d < kMaxDevices; ++d)
        if (g_pool[d]) cudaMemPoolTrimTo(g_pool[d], 0);
    return 0;
}

extern "C" int ocm_simca_score_f64in_bulk_supported(int64_t n, int32_t p, int64_t ldx, int32_t A, const void* X,
                                                    const void* mu) {
    if (n < ocm::staged::kMinBulkRows || (ldx & 1) || ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(mu)) & 15))
        return 0;
    return ocm_simca_score_tc_supported(n, p, p, A, 0, nullptr);      // the scratch rows: fp32, leading dimension p, aligned
}

extern "C" int ocm_simca_score_f64in(const double* X, int64_t n, int32_t p, int64_t ldx, const double* mu,
                                     const float* P, const float* Sinv, const ocm_score_params* hp,
                                     float* T2, float* Q, uint8_t* accept, float* scores, void* stream) {
    using namespace ocm;
    using namespace ocm::staged;
    if (n == 0) return 0;
    OCM_REQUIRE(hp, OCM_EINVAL, "NULL params");
    static const bool disabled = [] { const char* e = getenv("OCM_DISABLE_STAGED"); return e && e[0] == '1'; }();
    if (scores || disabled || !X || !mu || !ocm_simca_score_f64in_bulk_supported(n, p, ldx, hp->A, X, mu))
        return ocm_simca_score_simt_f64in(X, n, p, ldx, mu, P, Sinv, hp, T2, Q, accept, scores, stream);
    const int snv = hp->snv;
    const double eps = (double)hp->snv_eps;
    const int rc = score_chunks(n, p, P, Sinv, hp, T2, Q, accept, stream, [&](int64_t r0, int64_t m, float* xc) {
        const int64_t want = (m + 7) / 8, most = (int64_t)sm_count() * 8;
        const unsigned blocks = (unsigned)(want < most ? want : most);
        if (snv) centre_f64_kernel<true><<<blocks, 256, 0, (cudaStream_t)stream>>>(X + r0 * ldx, m, p, ldx, mu, eps, xc, p);
        else     centre_f64_kernel<false><<<blocks, 256, 0, (cudaStream_t)stream>>>(X + r0 * ldx, m, p, ldx, mu, 0.0, xc, p);
        return check_launch("centre_f64_kernel");
    });
    if (rc == kNoScratch) return ocm_simca_score_simt_f64in(X, n, p, ldx, mu, P, Sinv, hp, T2, Q, accept, scores, stream);
    return rc;
}

// fp32 rows with fused SNV at a shape whose plain (snv = 0) form the tcgen05 kernels take but whose fused form they do
// not: 1 = ocm_simca_score_f32 stages SNV + centring through the scratch instead of falling back to the generic kernel
extern "C" int ocm_simca_score_snv_staged_supported(int64_t n, int32_t p, int64_t ldx, int32_t A, const void* X, const void* mu) {
    static const bool disabled = [] { const char* e = getenv("OCM_DISABLE_STAGED"); return e && e[0] == '1'; }();
    if (disabled || n < ocm::staged::kMinBulkRows || (ldx & 3) || ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(mu)) & 15))
        return 0;
    return !ocm_simca_score_tc_supported(n, p, ldx, A, 1, X) && ocm_simca_score_tc_supported(n, p, p, A, 0, nullptr);
}

extern "C" int ocm_simca_score_snv_staged_f32(const float* X, int64_t n, int32_t p, int64_t ldx, const float* mu,
                                              const float* P, const float* Sinv, const ocm_score_params* hp,
                                              float* T2, float* Q, uint8_t* accept, void* stream) {
    using namespace ocm;
    using namespace ocm::staged;
    if (n == 0) return 0;
    OCM_REQUIRE(X && mu && P && Sinv && hp && hp->snv, OCM_EINVAL, "NULL buffer or snv == 0");
    OCM_REQUIRE(ocm_simca_score_snv_staged_supported(n, p, ldx, hp->A, X, mu), OCM_ELIMIT, "shape not supported by the staged SNV path");
    const float eps = hp->snv_eps;
    const int rc = score_chunks(n, p, P, Sinv, hp, T2, Q, accept, stream, [&](int64_t r0, int64_t m, float* xc) {
        return ocm_preprocess_f32(X + r0 * ldx, m, p, ldx, OCM_PRE_SNV_CENTER, mu, nullptr, eps, xc, p, stream);
    });
    if (rc == kNoScratch) return ocm_simca_score_simt_f32(X, n, p, ldx, mu, P, Sinv, hp, T2, Q, accept, nullptr, stream);
    return rc;
}

// Raw spectra -> decisions in one call: [SNV +] Savitzky-Golay (ocm_savgol_f32: simca_nuts.py:48-52,
// utils/data_utils.py:57-59) into the scratch chunk, then the bulk dispatcher against the model (whose mean lives in the
// preprocessed space).  What the reference's drivers do as three numpy passes and a second full-size matrix.
extern "C" int ocm_simca_score_savgol_f32(const float* X, int64_t n, int32_t p, int64_t ldx, int32_t snv, float snv_eps,
                                          const float* h_coef, int32_t window, const float* edge_l, const float* edge_r,
                                          const float* mu, const float* P, const float* Sinv, const ocm_score_params* hp,
                                          float* T2, float* Q, uint8_t* accept, void* stream) {
    using namespace ocm;
    using namespace ocm::staged;
    if (n == 0) return 0;
    OCM_REQUIRE(X && mu && P && Sinv && hp && h_coef && edge_l && edge_r, OCM_EINVAL, "NULL buffer");
    OCM_REQUIRE(hp->snv == 0, OCM_EINVAL, "row SNV goes in front of the filter (the snv argument), not behind it");
    OCM_REQUIRE(n > 0 && p > 0 && ldx >= p, OCM_EINVAL, "bad shape");
    const int rc = score_chunks(n, p, P, Sinv, hp, T2, Q, accept, stream, [&](int64_t r0, int64_t m, float* xc) {
        return ocm_savgol_f32(X + r0 * ldx, m, p, ldx, snv, snv_eps, h_coef, window, edge_l, edge_r, xc, p, stream);
    }, mu);
    if (rc == kNoScratch) { set_error("ocm_simca_score_savgol_f32: no scratch memory for the filtered chunk"); return (int)cudaErrorMemoryAllocation; }
    return rc;
}
