// Public scoring entry points: path selection and the host-buffer convenience call.
#include "ocm_common.cuh"

#include <mutex>

extern "C" int ocm_simca_score_simt_f32(const float*, int64_t, int32_t, int64_t, const float*, const float*,
                                        const float*, const ocm_score_params*, float*, float*, uint8_t*, float*, void*);
extern "C" int ocm_simca_score_tc_f32(const float*, int64_t, int32_t, int64_t, const float*, const float*,
                                      const float*, const ocm_score_params*, float*, float*, uint8_t*, void*);
extern "C" int ocm_simca_score_tc_supported(int64_t n, int32_t p, int64_t ldx, int32_t A, int32_t snv, const void* X);
extern "C" int ocm_simca_score_rowteam_supported(int64_t n, int32_t p, int64_t ldx, int32_t A, int32_t snv, const void* X);
extern "C" int ocm_simca_score_rowteam_f32(const float*, int64_t, int32_t, int64_t, const float*, const float*,
                                           const float*, const ocm_score_params*, float*, float*, uint8_t*, void*);

extern "C" int ocm_simca_score_snv_staged_supported(int64_t n, int32_t p, int64_t ldx, int32_t A, const void* X, const void* mu);
extern "C" int ocm_simca_score_snv_staged_f32(const float*, int64_t, int32_t, int64_t, const float*, const float*,
                                              const float*, const ocm_score_params*, float*, float*, uint8_t*, void*);
extern "C" int ocm_simca_staged_release(void);

extern "C" int ocm_simca_score_f32(const float* X, int64_t n, int32_t p, int64_t ldx, const float* mu,
                                   const float* P, const float* Sinv, const ocm_score_params* hp,
                                   float* T2, float* Q, uint8_t* accept, float* scores, void* stream) {
    if (n == 0) return 0;
    OCM_REQUIRE(hp, OCM_EINVAL, "NULL params");
    if (!scores && ocm_simca_score_tc_supported(n, p, ldx, hp->A, hp->snv, X))
        return ocm_simca_score_tc_f32(X, n, p, ldx, mu, P, Sinv, hp, T2, Q, accept, stream);
    // fused SNV wider than the tcgen05 kernel fuses it: SNV + centring staged through a scratch chunk, then the bulk kernel
    if (!scores && hp->snv && ocm_simca_score_snv_staged_supported(n, p, ldx, hp->A, X, mu))
        return ocm_simca_score_snv_staged_f32(X, n, p, ldx, mu, P, Sinv, hp, T2, Q, accept, stream);
    if (!scores && ocm_simca_score_rowteam_supported(n, p, ldx, hp->A, hp->snv, X) &&
        (reinterpret_cast<uintptr_t>(mu) & 15) == 0 && (reinterpret_cast<uintptr_t>(P) & 15) == 0)
        return ocm_simca_score_rowteam_f32(X, n, p, ldx, mu, P, Sinv, hp, T2, Q, accept, stream);
    return ocm_simca_score_simt_f32(X, n, p, ldx, mu, P, Sinv, hp, T2, Q, accept, scores, stream);
}

// Host buffers in, host buffers out: chunked, double-buffered over two streams so the H2D copy of chunk i+1 overlaps the
// kernel of chunk i.  This is the call bench.py times as e2e.  The staging buffers and the two streams are a
// per-device workspace that is kept between calls (grown on demand, released by ocm_simca_host_release): allocating
// and freeing ~1 GB per call costs milliseconds and a device-wide synchronisation.  Contiguous rows go as ONE 1-D copy
// per chunk (a 2-D copy of 128 Ki rows x 4 KiB ran at 19 GB/s on the B200 box where the 1-D copy reaches 55 GB/s).
namespace {
struct HostWorkspace {
    int device = -1;
    size_t x_bytes = 0, out_rows = 0, model_floats = 0;
    float* d_X[2] = {nullptr, nullptr};
    float* d_T2[2] = {nullptr, nullptr};
    float* d_Q[2] = {nullptr, nullptr};
    uint8_t* d_acc[2] = {nullptr, nullptr};
    float* d_model = nullptr;                 // mu | P | Sinv
    cudaStream_t st[2] = {nullptr, nullptr};
    void release() {
        for (int b = 0; b < 2; ++b) {
            if (st[b]) { cudaStreamSynchronize(st[b]); cudaStreamDestroy(st[b]); st[b] = nullptr; }
            cudaFree(d_X[b]); cudaFree(d_T2[b]); cudaFree(d_Q[b]); cudaFree(d_acc[b]);
            d_X[b] = d_T2[b] = d_Q[b] = nullptr; d_acc[b] = nullptr;
        }
        cudaFree(d_model); d_model = nullptr;
        x_bytes = out_rows = model_floats = 0; device = -1;
    }
};
// one workspace (and one lock) PER DEVICE: threads that drive different GPUs neither serialise on each other nor tear
// down each other's staging buffers (round-1 advice); a failing call releases its own device's entry only
constexpr int kMaxDevices = 64;
HostWorkspace g_ws[kMaxDevices];
std::mutex g_ws_mutex[kMaxDevices];
}  // namespace

extern "C" int ocm_simca_host_release(void) {
    int cur = 0;
    cudaGetDevice(&cur);
    for (int d = 0; d < kMaxDevices; ++d) {
        std::lock_guard<std::mutex> lock(g_ws_mutex[d]);
        if (g_ws[d].device < 0) continue;
        cudaSetDevice(d);
        g_ws[d].release();
    }
    cudaSetDevice(cur);
    ocm_simca_staged_release();
    return 0;
}

extern "C" int ocm_simca_score_host_f32(const float* h_X, int64_t n, int32_t p, int64_t ldx,
                                        const float* h_mu, const float* h_P, const float* h_Sinv,
                                        const ocm_score_params* hp, int64_t chunk_rows,
                                        float* h_T2, float* h_Q, uint8_t* h_accept) {
    using namespace ocm;
    OCM_REQUIRE(h_X && h_mu && h_P && h_Sinv && hp, OCM_EINVAL, "NULL buffer");
    OCM_REQUIRE(n >= 0 && p > 0 && ldx >= p, OCM_EINVAL, "bad shape");
    OCM_REQUIRE(hp->A >= 1 && hp->A <= 64, OCM_ELIMIT, "n_components must be 1..64");
    if (n == 0) return 0;
    if (chunk_rows <= 0) chunk_rows = 1 << 16;
    if (chunk_rows > n) chunk_rows = n;
    const int A = hp->A;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) {
        set_error("ocm_simca_score_host_f32: no usable current device");
        return (int)cudaErrorInvalidDevice;
    }
    std::lock_guard<std::mutex> lock(g_ws_mutex[dev]);      // one host-buffer call at a time per DEVICE
    HostWorkspace& w = g_ws[dev];
#define OCM_TRY(call)                                                                        \
    do { cudaError_t e_ = (call); if (e_ != cudaSuccess) {                                   \
        set_error("%s: %s", #call, cudaGetErrorString(e_)); w.release(); return (int)e_; } } while (0)
    w.device = dev;
    const size_t need_x = sizeof(float) * (size_t)chunk_rows * p, need_model = (size_t)p + (size_t)A * p + (size_t)A * A;
    for (int b = 0; b < 2; ++b) {
        if (!w.st[b]) OCM_TRY(cudaStreamCreateWithFlags(&w.st[b], cudaStreamNonBlocking));
        if (w.x_bytes < need_x) { OCM_TRY(cudaFree(w.d_X[b])); w.d_X[b] = nullptr; OCM_TRY(cudaMalloc(&w.d_X[b], need_x)); }
        if (w.out_rows < (size_t)chunk_rows) {
            OCM_TRY(cudaFree(w.d_T2[b])); OCM_TRY(cudaFree(w.d_Q[b])); OCM_TRY(cudaFree(w.d_acc[b]));
            w.d_T2[b] = w.d_Q[b] = nullptr; w.d_acc[b] = nullptr;
            OCM_TRY(cudaMalloc(&w.d_T2[b], sizeof(float) * chunk_rows));
            OCM_TRY(cudaMalloc(&w.d_Q[b], sizeof(float) * chunk_rows));
            OCM_TRY(cudaMalloc(&w.d_acc[b], chunk_rows));
        }
    }
    if (w.x_bytes < need_x) w.x_bytes = need_x;
    if (w.out_rows < (size_t)chunk_rows) w.out_rows = (size_t)chunk_rows;
    if (w.model_floats < need_model) {
        OCM_TRY(cudaFree(w.d_model)); w.d_model = nullptr;
        OCM_TRY(cudaMalloc(&w.d_model, sizeof(float) * need_model));
        w.model_floats = need_model;
    }
    float* d_mu = w.d_model; float* d_P = d_mu + p; float* d_S = d_P + (size_t)A * p;
    // the model goes first on BOTH streams' timeline: copy on stream 0, stream 1 waits for it
    OCM_TRY(cudaMemcpyAsync(d_mu, h_mu, sizeof(float) * p, cudaMemcpyHostToDevice, w.st[0]));
    OCM_TRY(cudaMemcpyAsync(d_P, h_P, sizeof(float) * (size_t)A * p, cudaMemcpyHostToDevice, w.st[0]));
    OCM_TRY(cudaMemcpyAsync(d_S, h_Sinv, sizeof(float) * (size_t)A * A, cudaMemcpyHostToDevice, w.st[0]));
    cudaEvent_t model_ready;
    OCM_TRY(cudaEventCreateWithFlags(&model_ready, cudaEventDisableTiming));
    cudaError_t e_rec = cudaEventRecord(model_ready, w.st[0]);
    if (e_rec == cudaSuccess) e_rec = cudaStreamWaitEvent(w.st[1], model_ready, 0);
    cudaEventDestroy(model_ready);                     // released by the runtime once the wait has been satisfied
    OCM_TRY(e_rec);
    int b = 0;
    for (int64_t r0 = 0; r0 < n; r0 += chunk_rows, b ^= 1) {
        const int64_t m = (n - r0 < chunk_rows) ? n - r0 : chunk_rows;
        if (ldx == p)
            OCM_TRY(cudaMemcpyAsync(w.d_X[b], h_X + r0 * ldx, sizeof(float) * (size_t)m * p, cudaMemcpyHostToDevice, w.st[b]));
        else
            OCM_TRY(cudaMemcpy2DAsync(w.d_X[b], sizeof(float) * p, h_X + r0 * ldx, sizeof(float) * ldx,
                                      sizeof(float) * p, (size_t)m, cudaMemcpyHostToDevice, w.st[b]));
        const int rc = ocm_simca_score_f32(w.d_X[b], m, p, p, d_mu, d_P, d_S, hp, w.d_T2[b], w.d_Q[b], w.d_acc[b], nullptr, w.st[b]);
        if (rc) { w.release(); return rc; }
        if (h_T2) OCM_TRY(cudaMemcpyAsync(h_T2 + r0, w.d_T2[b], sizeof(float) * m, cudaMemcpyDeviceToHost, w.st[b]));
        if (h_Q) OCM_TRY(cudaMemcpyAsync(h_Q + r0, w.d_Q[b], sizeof(float) * m, cudaMemcpyDeviceToHost, w.st[b]));
        if (h_accept) OCM_TRY(cudaMemcpyAsync(h_accept + r0, w.d_acc[b], (size_t)m, cudaMemcpyDeviceToHost, w.st[b]));
    }
    for (int k = 0; k < 2; ++k) OCM_TRY(cudaStreamSynchronize(w.st[k]));
#undef OCM_TRY
    return tc_status_report("ocm_simca_score_host_f32");      // everything has completed: the status word is exact here
}
