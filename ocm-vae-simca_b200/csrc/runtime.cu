// Library bookkeeping: ABI version, thread-local error text, launch counter.
#include "ocm_common.cuh"

#include <atomic>
#include <stdarg.h>

namespace ocm {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int sm_count() {
    static thread_local int cached_dev = -1;
    static thread_local int cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached = n;
        cached_dev = dev;
    }
    return cached;
}

// ---- protocol status of the tcgen05 kernels ---------------------------------------------------------------
// One 32-bit word per device in MAPPED pinned host memory.  A kernel whose bounded mbarrier wait expires writes a
// non-zero code there and traps: the launch fails (CUDA reports cudaErrorLaunchFailure at the next synchronisation and
// the context is lost, as for any device-side fault).  The host reads the word without touching the device at the
// entry of every tensor-core call and returns OCM_ETIMEOUT with the cause spelled out; the word is NOT cleared, the
// process has to be restarted.  (The Gram kernel, whose waits are off the critical path, instead stops its roles,
// writes NaN tiles and finishes: codes >= 10, reported once and cleared.)
namespace {
struct StatusWord { unsigned int* host = nullptr; unsigned int* dev = nullptr; };
StatusWord g_status[64];
}

unsigned int* tc_status_device_ptr() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    StatusWord& w = g_status[dev];
    if (!w.host) {
        void* h = nullptr;
        if (cudaHostAlloc(&h, 64, cudaHostAllocMapped) != cudaSuccess) return nullptr;
        w.host = static_cast<unsigned int*>(h);
        for (int i = 0; i < 16; ++i) w.host[i] = 0;
        void* d = nullptr;
        if (cudaHostGetDevicePointer(&d, h, 0) != cudaSuccess) { cudaFreeHost(h); w.host = nullptr; return nullptr; }
        w.dev = static_cast<unsigned int*>(d);
    }
    return w.dev;
}

// One 32-bit word per device in DEVICE memory for the fp16 operand mode of the scoring kernel: a launch that produced a
// non-finite distance stores its launch id there (atomicMax), and the fix-up kernel that follows on the same stream
// scans T2 / Q only when it finds its own id -- otherwise every CTA reads one word and leaves.  Never reset (ids grow).
namespace {
unsigned int* g_flag[64];
}
unsigned int* tc_flag_device_ptr() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    if (!g_flag[dev]) {
        void* d = nullptr;
        if (cudaMalloc(&d, 64) != cudaSuccess) return nullptr;
        if (cudaMemset(d, 0, 64) != cudaSuccess) { cudaFree(d); return nullptr; }
        g_flag[dev] = static_cast<unsigned int*>(d);
    }
    return g_flag[dev];
}

int tc_status_take() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
    StatusWord& w = g_status[dev];
    if (!w.host) return 0;
    const unsigned int v = *reinterpret_cast<volatile unsigned int*>(w.host);
    if (v >= 10u && v < 100u) *reinterpret_cast<volatile unsigned int*>(w.host) = 0;      // Gram codes: the kernel finished, report once
    return (int)v;
}

int tc_status_report(const char* where) {
    const int v = tc_status_take();
    if (!v) return 0;
    if (v >= 10 && v < 100)
        set_error("%s: the tcgen05 Gram kernel of an earlier launch gave up a bounded mbarrier wait (code %d); the tiles it "
                  "owned were written as NaN", where, v);
    else
        set_error("%s: the tcgen05 scoring kernel of an earlier launch %s (code %d) and trapped; that launch failed, none of "
                  "its outputs is valid and the CUDA context of this process is lost", where,
                  v == 2 ? "found its TMEM allocation away from column 0" : "gave up a bounded mbarrier wait", v);      // >= 100: 100 + barrier slot of score_tc2.cu
    return OCM_ETIMEOUT;
}

}  // namespace ocm

extern "C" int ocm_abi_version(void) { return OCM_ABI_VERSION; }
extern "C" const char* ocm_last_error(void) { return ocm::g_err; }
extern "C" uint64_t ocm_launch_count(void) { return ocm::g_launches.load(std::memory_order_relaxed); }
