// Public scoring entry points: path selection and the host-buffer convenience call.
#include "ocm_common.cuh"

extern "C" int ocm_simca_score_simt_f32(const float*, int64_t, int32_t, int64_t, const float*, const float*,
                                        const float*, const ocm_score_params*, float*, float*, uint8_t*, float*, void*);
extern "C" int ocm_simca_score_tc_f32(const float*, int64_t, int32_t, int64_t, const float*, const float*,
                                      const float*, const ocm_score_params*, float*, float*, uint8_t*, void*);
extern "C" int ocm_simca_score_tc_supported(int64_t n, int32_t p, int64_t ldx, int32_t A, int32_t snv, const void* X);
extern "C" int ocm_simca_score_rowteam_supported(int64_t n, int32_t p, int64_t ldx, int32_t A, int32_t snv, const void* X);
extern "C" int ocm_simca_score_rowteam_f32(const float*, int64_t, int32_t, int64_t, const float*, const float*,
                                           const float*, const ocm_score_params*, float*, float*, uint8_t*, void*);

extern "C" int ocm_simca_score_f32(const float* X, int64_t n, int32_t p, int64_t ldx, const float* mu,
                                   const float* P, const float* Sinv, const ocm_score_params* hp,
                                   float* T2, float* Q, uint8_t* accept, float* scores, void* stream) {
    if (n == 0) return 0;
    OCM_REQUIRE(hp, OCM_EINVAL, "NULL params");
    if (!scores && ocm_simca_score_tc_supported(n, p, ldx, hp->A, hp->snv, X))
        return ocm_simca_score_tc_f32(X, n, p, ldx, mu, P, Sinv, hp, T2, Q, accept, stream);
    if (!scores && ocm_simca_score_rowteam_supported(n, p, ldx, hp->A, hp->snv, X) &&
        (reinterpret_cast<uintptr_t>(mu) & 15) == 0 && (reinterpret_cast<uintptr_t>(P) & 15) == 0)
        return ocm_simca_score_rowteam_f32(X, n, p, ldx, mu, P, Sinv, hp, T2, Q, accept, stream);
    return ocm_simca_score_simt_f32(X, n, p, ldx, mu, P, Sinv, hp, T2, Q, accept, scores, stream);
}

// Host buffers in, host buffers out: chunked, double-buffered over two streams so the H2D
// copy of chunk i+1 overlaps the kernel of chunk i.  This is the call bench.py times as e2e.
extern "C" int ocm_simca_score_host_f32(const float* h_X, int64_t n, int32_t p, int64_t ldx,
                                        const float* h_mu, const float* h_P, const float* h_Sinv,
                                        const ocm_score_params* hp, int64_t chunk_rows,
                                        float* h_T2, float* h_Q, uint8_t* h_accept) {
    using namespace ocm;
    OCM_REQUIRE(h_X && h_mu && h_P && h_Sinv && hp, OCM_EINVAL, "NULL buffer");
    OCM_REQUIRE(n >= 0 && p > 0 && ldx >= p, OCM_EINVAL, "bad shape");
    OCM_REQUIRE(hp->A >= 1 && hp->A <= 64, OCM_ELIMIT, "n_components must be 1..64");
    if (n == 0) return 0;
    if (chunk_rows <= 0) chunk_rows = 1 << 17;
    if (chunk_rows > n) chunk_rows = n;
    const int A = hp->A;
    float *d_mu = nullptr, *d_P = nullptr, *d_S = nullptr;
    float* d_X[2] = {nullptr, nullptr};
    float* d_T2[2] = {nullptr, nullptr};
    float* d_Q[2] = {nullptr, nullptr};
    uint8_t* d_acc[2] = {nullptr, nullptr};
    cudaStream_t st[2] = {nullptr, nullptr};
    int rc = 0;
    auto cleanup = [&]() {
        for (int b = 0; b < 2; ++b) {
            if (st[b]) { cudaStreamSynchronize(st[b]); cudaStreamDestroy(st[b]); }
            cudaFree(d_X[b]); cudaFree(d_T2[b]); cudaFree(d_Q[b]); cudaFree(d_acc[b]);
        }
        cudaFree(d_mu); cudaFree(d_P); cudaFree(d_S);
    };
#define OCM_TRY(call)                                                                        \
    do { cudaError_t e_ = (call); if (e_ != cudaSuccess) {                                   \
        set_error("%s: %s", #call, cudaGetErrorString(e_)); cleanup(); return (int)e_; } } while (0)
    OCM_TRY(cudaMalloc(&d_mu, sizeof(float) * p));
    OCM_TRY(cudaMalloc(&d_P, sizeof(float) * (size_t)A * p));
    OCM_TRY(cudaMalloc(&d_S, sizeof(float) * (size_t)A * A));
    OCM_TRY(cudaMemcpy(d_mu, h_mu, sizeof(float) * p, cudaMemcpyHostToDevice));
    OCM_TRY(cudaMemcpy(d_P, h_P, sizeof(float) * (size_t)A * p, cudaMemcpyHostToDevice));
    OCM_TRY(cudaMemcpy(d_S, h_Sinv, sizeof(float) * (size_t)A * A, cudaMemcpyHostToDevice));
    for (int b = 0; b < 2; ++b) {
        OCM_TRY(cudaStreamCreateWithFlags(&st[b], cudaStreamNonBlocking));
        OCM_TRY(cudaMalloc(&d_X[b], sizeof(float) * (size_t)chunk_rows * p));
        OCM_TRY(cudaMalloc(&d_T2[b], sizeof(float) * chunk_rows));
        OCM_TRY(cudaMalloc(&d_Q[b], sizeof(float) * chunk_rows));
        OCM_TRY(cudaMalloc(&d_acc[b], chunk_rows));
    }
    int b = 0;
    for (int64_t r0 = 0; r0 < n; r0 += chunk_rows, b ^= 1) {
        const int64_t m = (n - r0 < chunk_rows) ? n - r0 : chunk_rows;
        OCM_TRY(cudaMemcpy2DAsync(d_X[b], sizeof(float) * p, h_X + r0 * ldx, sizeof(float) * ldx,
                                  sizeof(float) * p, (size_t)m, cudaMemcpyHostToDevice, st[b]));
        rc = ocm_simca_score_f32(d_X[b], m, p, p, d_mu, d_P, d_S, hp, d_T2[b], d_Q[b], d_acc[b], nullptr, st[b]);
        if (rc) { cleanup(); return rc; }
        if (h_T2) OCM_TRY(cudaMemcpyAsync(h_T2 + r0, d_T2[b], sizeof(float) * m, cudaMemcpyDeviceToHost, st[b]));
        if (h_Q) OCM_TRY(cudaMemcpyAsync(h_Q + r0, d_Q[b], sizeof(float) * m, cudaMemcpyDeviceToHost, st[b]));
        if (h_accept) OCM_TRY(cudaMemcpyAsync(h_accept + r0, d_acc[b], (size_t)m, cudaMemcpyDeviceToHost, st[b]));
    }
    for (int k = 0; k < 2; ++k) OCM_TRY(cudaStreamSynchronize(st[k]));
#undef OCM_TRY
    cleanup();
    return 0;
}
