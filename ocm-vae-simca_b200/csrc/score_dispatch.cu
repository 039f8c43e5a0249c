// Public scoring entry points: path selection and the host-buffer convenience call.
#include "ocm_common.cuh"

#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

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
//
// Pinned (page-locked / registered) fp32 rows are copied straight from the caller's buffer.  PAGEABLE rows -- what a
// numpy array is -- and fp64 rows go through two pinned bounce buffers of the workspace that a few host threads fill in
// parallel while the previous chunk is on the link: a cudaMemcpyAsync from pageable memory is staged by the driver on
// one thread (and blocks the host), the parallel fill keeps the link busy.  fp64 rows are centred in fp64 and rounded
// once during that fill ((float)(x - mu), bit-identical to the device-side fp64 path), so only 4 bytes per element cross
// the link; with fused SNV they are rounded to fp32 and normalised by the kernel.  Results come back through pinned
// buffers as well and are handed to the caller's arrays when the slot is reused.
namespace {
struct HostWorkspace {
    int device = -1;
    size_t x_bytes = 0, out_rows = 0, model_floats = 0, bounce_bytes = 0, bounce_rows = 0;
    float* d_X[2] = {nullptr, nullptr};
    float* d_T2[2] = {nullptr, nullptr};
    float* d_Q[2] = {nullptr, nullptr};
    uint8_t* d_acc[2] = {nullptr, nullptr};
    float* d_model = nullptr;                 // mu | P | Sinv
    float* h_X[2] = {nullptr, nullptr};       // pinned bounce buffers (pageable / fp64 sources)
    float* h_out[2] = {nullptr, nullptr};     // pinned results of a chunk: T2 | Q | accept
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    cudaStream_t st[2] = {nullptr, nullptr};
    void release() {
        for (int b = 0; b < 2; ++b) {
            if (st[b]) { cudaStreamSynchronize(st[b]); cudaStreamDestroy(st[b]); st[b] = nullptr; }
            cudaFree(d_X[b]); cudaFree(d_T2[b]); cudaFree(d_Q[b]); cudaFree(d_acc[b]);
            d_X[b] = d_T2[b] = d_Q[b] = nullptr; d_acc[b] = nullptr;
            cudaFreeHost(h_X[b]); cudaFreeHost(h_out[b]); h_X[b] = h_out[b] = nullptr;
            if (ev_h2d[b]) { cudaEventDestroy(ev_h2d[b]); ev_h2d[b] = nullptr; }
            if (ev_done[b]) { cudaEventDestroy(ev_done[b]); ev_done[b] = nullptr; }
        }
        cudaFree(d_model); d_model = nullptr;
        x_bytes = out_rows = model_floats = bounce_bytes = bounce_rows = 0; device = -1;
    }
};
// one workspace (and one lock) PER DEVICE: threads that drive different GPUs neither serialise on each other nor tear
// down each other's staging buffers (round-1 advice); a failing call releases its own device's entry only
constexpr int kMaxDevices = 64;
HostWorkspace g_ws[kMaxDevices];
std::mutex g_ws_mutex[kMaxDevices];

bool is_pinned(const void* ptr) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}

int fill_threads() {
    const char* e = getenv("OCM_HOST_THREADS");
    long t = e ? atol(e) : 0;
    if (t < 1) { t = (long)std::thread::hardware_concurrency(); if (t > 16) t = 16; }
    return (int)(t < 1 ? 1 : (t > 64 ? 64 : t));
}

// work(r0, r1) over [0, m) in `nt` contiguous ranges, one per thread (the caller's thread takes the first).  A thread that
// cannot be started (resource limits) does not take the process down: its range is done inline.
template <typename Work>
void parallel_ranges(int64_t m, int nt, Work work) {
    if (nt > m) nt = (int)(m > 0 ? m : 1);
    if (nt <= 1) { work((int64_t)0, m); return; }
    std::vector<std::thread> th;
    th.reserve(nt - 1);
    const int64_t per = (m + nt - 1) / nt;
    for (int t = 1; t < nt; ++t) {
        const int64_t r0 = t * per, r1 = (r0 + per < m) ? r0 + per : m;
        if (r0 >= r1) continue;
        try { th.emplace_back(work, r0, r1); } catch (...) { work(r0, r1); }
    }
    work((int64_t)0, per < m ? per : m);
    for (auto& t : th) t.join();
}

// rows [0, m) of the caller's matrix -> dense fp32 rows [m][p] in a pinned buffer, on `nt` threads (row ranges)
template <typename TIn>
void fill_rows(float* dst, const TIn* src, int64_t m, int32_t p, int64_t ldx, const double* mu_or_null, int nt) {
    auto work = [=](int64_t r0, int64_t r1) {
        if (sizeof(TIn) == 4 && ldx == p) {          // dense rows: one large copy per thread (streaming stores inside memcpy)
            memcpy(dst + r0 * (int64_t)p, src + r0 * ldx, sizeof(float) * (size_t)(r1 - r0) * p);
            return;
        }
        for (int64_t r = r0; r < r1; ++r) {
            const TIn* x = src + r * ldx;
            float* o = dst + r * (int64_t)p;
            if constexpr (sizeof(TIn) == 4) {
                memcpy(o, x, sizeof(float) * (size_t)p);
            } else {
                if (mu_or_null) for (int32_t j = 0; j < p; ++j) o[j] = (float)(x[j] - mu_or_null[j]);      // centre in fp64, round once
                else            for (int32_t j = 0; j < p; ++j) o[j] = (float)x[j];
            }
        }
    };
    parallel_ranges(m, nt, work);
}

template <typename TIn>
int score_host_impl(const TIn* h_X, int64_t n, int32_t p, int64_t ldx, const TIn* h_mu, const float* h_P,
                    const float* h_Sinv, const ocm_score_params* hp, int64_t chunk_rows,
                    float* h_T2, float* h_Q, uint8_t* h_accept, const char* who) {
    using namespace ocm;
    constexpr bool F64 = sizeof(TIn) == 8;
    OCM_REQUIRE(h_X && h_mu && h_P && h_Sinv && hp, OCM_EINVAL, "NULL buffer");
    OCM_REQUIRE(n >= 0 && p > 0 && ldx >= p, OCM_EINVAL, "bad shape");
    OCM_REQUIRE(hp->A >= 1 && hp->A <= 64, OCM_ELIMIT, "n_components must be 1..64");
    if (n == 0) return 0;
    if (chunk_rows <= 0) chunk_rows = 1 << 16;
    if (chunk_rows > n) chunk_rows = n;
    // equal chunks (no short tail): with the default size every chunk of a call of >= 32 768 rows stays above the bulk
    // kernels' row gate, so all rows of one call are scored by the same kernel
    chunk_rows = (n + (n + chunk_rows - 1) / chunk_rows - 1) / ((n + chunk_rows - 1) / chunk_rows);
    const int A = hp->A;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) {
        set_error("%s: no usable current device", who);
        return (int)cudaErrorInvalidDevice;
    }
    std::lock_guard<std::mutex> lock(g_ws_mutex[dev]);      // one host-buffer call at a time per DEVICE
    HostWorkspace& w = g_ws[dev];
#define OCM_TRY(call)                                                                        \
    do { cudaError_t e_ = (call); if (e_ != cudaSuccess) {                                   \
        set_error("%s: %s", #call, cudaGetErrorString(e_)); w.release(); return (int)e_; } } while (0)
    w.device = dev;
    const bool bounce = F64 || !is_pinned(h_X);             // through the pinned bounce buffers (parallel fill)
    const bool out_pinned = (!h_T2 || is_pinned(h_T2)) && (!h_Q || is_pinned(h_Q)) && (!h_accept || is_pinned(h_accept));
    const bool bounce_out = !out_pinned;                    // a D2H copy into pageable memory would block the host on the kernel
    const size_t need_x = sizeof(float) * (size_t)chunk_rows * p, need_model = 2 * (size_t)p + (size_t)A * p + (size_t)A * A;
    for (int b = 0; b < 2; ++b) {
        if (!w.st[b]) OCM_TRY(cudaStreamCreateWithFlags(&w.st[b], cudaStreamNonBlocking));
        if (!w.ev_h2d[b]) OCM_TRY(cudaEventCreateWithFlags(&w.ev_h2d[b], cudaEventDisableTiming));
        if (!w.ev_done[b]) OCM_TRY(cudaEventCreateWithFlags(&w.ev_done[b], cudaEventDisableTiming));
        if (w.x_bytes < need_x) { OCM_TRY(cudaFree(w.d_X[b])); w.d_X[b] = nullptr; OCM_TRY(cudaMalloc(&w.d_X[b], need_x)); }
        if (w.out_rows < (size_t)chunk_rows) {
            OCM_TRY(cudaFree(w.d_T2[b])); OCM_TRY(cudaFree(w.d_Q[b])); OCM_TRY(cudaFree(w.d_acc[b]));
            w.d_T2[b] = w.d_Q[b] = nullptr; w.d_acc[b] = nullptr;
            OCM_TRY(cudaMalloc(&w.d_T2[b], sizeof(float) * chunk_rows));
            OCM_TRY(cudaMalloc(&w.d_Q[b], sizeof(float) * chunk_rows));
            OCM_TRY(cudaMalloc(&w.d_acc[b], chunk_rows));
        }
        if (bounce && w.bounce_bytes < need_x) { OCM_TRY(cudaFreeHost(w.h_X[b])); w.h_X[b] = nullptr; OCM_TRY(cudaMallocHost(&w.h_X[b], need_x)); }
        if (bounce_out && w.bounce_rows < (size_t)chunk_rows) {
            OCM_TRY(cudaFreeHost(w.h_out[b])); w.h_out[b] = nullptr;
            OCM_TRY(cudaMallocHost(&w.h_out[b], 9 * (size_t)chunk_rows + 16));
        }
    }
    if (w.x_bytes < need_x) w.x_bytes = need_x;
    if (w.out_rows < (size_t)chunk_rows) w.out_rows = (size_t)chunk_rows;
    if (bounce && w.bounce_bytes < need_x) w.bounce_bytes = need_x;
    if (bounce_out && w.bounce_rows < (size_t)chunk_rows) w.bounce_rows = (size_t)chunk_rows;
    if (w.model_floats < need_model) {
        OCM_TRY(cudaFree(w.d_model)); w.d_model = nullptr;
        OCM_TRY(cudaMalloc(&w.d_model, sizeof(float) * need_model));
        w.model_floats = need_model;
    }
    float* d_mu = w.d_model; float* d_P = d_mu + p; float* d_S = d_P + (size_t)A * p; float* d_zero = d_S + (size_t)A * A;
    // fp64 rows without SNV are centred during the fill: the kernel sees a zero mean.  With SNV the kernel normalises the
    // rounded rows and subtracts the fp32 mean.
    const bool centre_on_host = F64 && !hp->snv;
    std::vector<float> mu32;
    std::vector<double> mu64;
    const float* h_mu32 = nullptr;
    if constexpr (F64) {
        mu32.resize(p);
        for (int32_t j = 0; j < p; ++j) mu32[j] = (float)h_mu[j];
        h_mu32 = mu32.data();
        if (centre_on_host) mu64.assign(h_mu, h_mu + p);
    } else {
        h_mu32 = h_mu;
    }
    // the model goes first on BOTH streams' timeline: copy on stream 0, stream 1 waits for it
    OCM_TRY(cudaMemcpyAsync(d_mu, h_mu32, sizeof(float) * p, cudaMemcpyHostToDevice, w.st[0]));
    OCM_TRY(cudaMemcpyAsync(d_P, h_P, sizeof(float) * (size_t)A * p, cudaMemcpyHostToDevice, w.st[0]));
    OCM_TRY(cudaMemcpyAsync(d_S, h_Sinv, sizeof(float) * (size_t)A * A, cudaMemcpyHostToDevice, w.st[0]));
    OCM_TRY(cudaMemsetAsync(d_zero, 0, sizeof(float) * p, w.st[0]));
    OCM_TRY(cudaStreamSynchronize(w.st[0]));               // mu32 (a local) has been read; stream 1 may start
    const float* d_mean = centre_on_host ? d_zero : d_mu;
    const int nt = bounce ? fill_threads() : 1;
    struct Pending { int64_t r0 = 0, m = 0; bool live = false; } pend[2];
    auto deliver = [&](int b) -> cudaError_t {              // results of the chunk that used slot b -> the caller's arrays
        if (!pend[b].live) return cudaSuccess;
        cudaError_t e = cudaEventSynchronize(w.ev_done[b]);
        if (e != cudaSuccess) return e;
        const int64_t r0 = pend[b].r0, m = pend[b].m;
        const float* o = w.h_out[b];
        if (h_T2) memcpy(h_T2 + r0, o, sizeof(float) * (size_t)m);
        if (h_Q) memcpy(h_Q + r0, o + chunk_rows, sizeof(float) * (size_t)m);
        if (h_accept) memcpy(h_accept + r0, reinterpret_cast<const uint8_t*>(o + 2 * chunk_rows), (size_t)m);
        pend[b].live = false;
        return cudaSuccess;
    };
    int b = 0;
    for (int64_t r0 = 0; r0 < n; r0 += chunk_rows, b ^= 1) {
        const int64_t m = (n - r0 < chunk_rows) ? n - r0 : chunk_rows;
        if (bounce_out) OCM_TRY(deliver(b));                // results of the chunk that used this slot two chunks ago
        if (bounce) {
            OCM_TRY(cudaEventSynchronize(w.ev_h2d[b]));     // the copy that last read this bounce buffer has finished
            fill_rows<TIn>(w.h_X[b], h_X + r0 * ldx, m, p, ldx, centre_on_host ? mu64.data() : nullptr, nt);
            OCM_TRY(cudaMemcpyAsync(w.d_X[b], w.h_X[b], sizeof(float) * (size_t)m * p, cudaMemcpyHostToDevice, w.st[b]));
            OCM_TRY(cudaEventRecord(w.ev_h2d[b], w.st[b]));
        } else if constexpr (!F64) {
            if (ldx == p)
                OCM_TRY(cudaMemcpyAsync(w.d_X[b], h_X + r0 * ldx, sizeof(float) * (size_t)m * p, cudaMemcpyHostToDevice, w.st[b]));
            else
                OCM_TRY(cudaMemcpy2DAsync(w.d_X[b], sizeof(float) * p, h_X + r0 * ldx, sizeof(float) * ldx,
                                          sizeof(float) * p, (size_t)m, cudaMemcpyHostToDevice, w.st[b]));
        }
        const int rc = ocm_simca_score_f32(w.d_X[b], m, p, p, d_mean, d_P, d_S, hp, w.d_T2[b], w.d_Q[b], w.d_acc[b], nullptr, w.st[b]);
        if (rc) { w.release(); return rc; }
        if (bounce_out) {
            float* o = w.h_out[b];
            if (h_T2) OCM_TRY(cudaMemcpyAsync(o, w.d_T2[b], sizeof(float) * m, cudaMemcpyDeviceToHost, w.st[b]));
            if (h_Q) OCM_TRY(cudaMemcpyAsync(o + chunk_rows, w.d_Q[b], sizeof(float) * m, cudaMemcpyDeviceToHost, w.st[b]));
            if (h_accept) OCM_TRY(cudaMemcpyAsync(o + 2 * chunk_rows, w.d_acc[b], (size_t)m, cudaMemcpyDeviceToHost, w.st[b]));
            OCM_TRY(cudaEventRecord(w.ev_done[b], w.st[b]));
            pend[b].r0 = r0; pend[b].m = m; pend[b].live = true;
        } else {
            if (h_T2) OCM_TRY(cudaMemcpyAsync(h_T2 + r0, w.d_T2[b], sizeof(float) * m, cudaMemcpyDeviceToHost, w.st[b]));
            if (h_Q) OCM_TRY(cudaMemcpyAsync(h_Q + r0, w.d_Q[b], sizeof(float) * m, cudaMemcpyDeviceToHost, w.st[b]));
            if (h_accept) OCM_TRY(cudaMemcpyAsync(h_accept + r0, w.d_acc[b], (size_t)m, cudaMemcpyDeviceToHost, w.st[b]));
        }
    }
    for (int k = 0; k < 2; ++k) OCM_TRY(cudaStreamSynchronize(w.st[k]));
    for (int k = 0; k < 2; ++k) OCM_TRY(deliver(k));
#undef OCM_TRY
    return tc_status_report(who);      // everything has completed: the status word is exact here
}
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

// Pageable host rows -> device rows at the rate of the parallel bounce fill (the same workspace): what engine.to_device
// uses for a large numpy matrix in front of fit / CV / preprocessing.  Blocking; the data is complete on return.
extern "C" int ocm_upload_rows(const void* h_src, int64_t n, int64_t row_bytes, int64_t src_stride,
                               void* d_dst, int64_t dst_stride) {
    using namespace ocm;
    if (n == 0) return 0;
    OCM_REQUIRE(h_src && d_dst && n > 0 && row_bytes > 0 && src_stride >= row_bytes && dst_stride >= row_bytes, OCM_EINVAL,
                "bad shape or NULL buffer");
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) {
        set_error("ocm_upload_rows: no usable current device");
        return (int)cudaErrorInvalidDevice;
    }
    std::lock_guard<std::mutex> lock(g_ws_mutex[dev]);
    HostWorkspace& w = g_ws[dev];
#define OCM_TRY(call)                                                                        \
    do { cudaError_t e_ = (call); if (e_ != cudaSuccess) {                                   \
        set_error("%s: %s", #call, cudaGetErrorString(e_)); w.release(); return (int)e_; } } while (0)
    w.device = dev;
    const char* src = static_cast<const char*>(h_src);
    char* dst = static_cast<char*>(d_dst);
    for (int b = 0; b < 2; ++b) {
        if (!w.st[b]) OCM_TRY(cudaStreamCreateWithFlags(&w.st[b], cudaStreamNonBlocking));
        if (!w.ev_h2d[b]) OCM_TRY(cudaEventCreateWithFlags(&w.ev_h2d[b], cudaEventDisableTiming));
    }
    if (is_pinned(h_src)) {
        if (src_stride == row_bytes && dst_stride == row_bytes)
            OCM_TRY(cudaMemcpyAsync(dst, src, (size_t)n * row_bytes, cudaMemcpyHostToDevice, w.st[0]));
        else
            OCM_TRY(cudaMemcpy2DAsync(dst, (size_t)dst_stride, src, (size_t)src_stride, (size_t)row_bytes, (size_t)n,
                                      cudaMemcpyHostToDevice, w.st[0]));
        OCM_TRY(cudaStreamSynchronize(w.st[0]));
        return 0;
    }
    int64_t rows = (64ll << 20) / row_bytes;
    if (rows < 1) rows = 1;
    if (rows > n) rows = n;
    const size_t need = (size_t)rows * (size_t)row_bytes;
    if (w.bounce_bytes < need) {
        for (int b = 0; b < 2; ++b) { OCM_TRY(cudaFreeHost(w.h_X[b])); w.h_X[b] = nullptr; OCM_TRY(cudaMallocHost(&w.h_X[b], need)); }
        w.bounce_bytes = need;
    }
    const int nt = fill_threads();
    int b = 0;
    for (int64_t r0 = 0; r0 < n; r0 += rows, b ^= 1) {
        const int64_t m = (n - r0 < rows) ? n - r0 : rows;
        OCM_TRY(cudaEventSynchronize(w.ev_h2d[b]));
        char* bounce = reinterpret_cast<char*>(w.h_X[b]);
        auto work = [=](int64_t a0, int64_t a1) {
            if (src_stride == row_bytes) { memcpy(bounce + a0 * row_bytes, src + (r0 + a0) * src_stride, (size_t)(a1 - a0) * row_bytes); return; }
            for (int64_t r = a0; r < a1; ++r) memcpy(bounce + r * row_bytes, src + (r0 + r) * src_stride, (size_t)row_bytes);
        };
        parallel_ranges(m, (int64_t)m * row_bytes < (4 << 20) ? 1 : nt, work);
        if (dst_stride == row_bytes)      // dense destination: ONE 1-D copy (a 2-D copy of narrow rows runs at a third of the link)
            OCM_TRY(cudaMemcpyAsync(dst + r0 * dst_stride, bounce, (size_t)m * row_bytes, cudaMemcpyHostToDevice, w.st[b]));
        else
            OCM_TRY(cudaMemcpy2DAsync(dst + r0 * dst_stride, (size_t)dst_stride, bounce, (size_t)row_bytes, (size_t)row_bytes, (size_t)m,
                                      cudaMemcpyHostToDevice, w.st[b]));
        OCM_TRY(cudaEventRecord(w.ev_h2d[b], w.st[b]));
    }
    for (int k = 0; k < 2; ++k) OCM_TRY(cudaStreamSynchronize(w.st[k]));
#undef OCM_TRY
    return 0;
}

extern "C" int ocm_simca_score_host_f32(const float* h_X, int64_t n, int32_t p, int64_t ldx,
                                        const float* h_mu, const float* h_P, const float* h_Sinv,
                                        const ocm_score_params* hp, int64_t chunk_rows,
                                        float* h_T2, float* h_Q, uint8_t* h_accept) {
    return score_host_impl<float>(h_X, n, p, ldx, h_mu, h_P, h_Sinv, hp, chunk_rows, h_T2, h_Q, h_accept,
                                  "ocm_simca_score_host_f32");
}

extern "C" int ocm_simca_score_host_f64in(const double* h_X, int64_t n, int32_t p, int64_t ldx,
                                          const double* h_mu, const float* h_P, const float* h_Sinv,
                                          const ocm_score_params* hp, int64_t chunk_rows,
                                          float* h_T2, float* h_Q, uint8_t* h_accept) {
    return score_host_impl<double>(h_X, n, p, ldx, h_mu, h_P, h_Sinv, hp, chunk_rows, h_T2, h_Q, h_accept,
                                   "ocm_simca_score_host_f64in");
}
