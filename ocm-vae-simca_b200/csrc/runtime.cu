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

}  // namespace ocm

extern "C" int ocm_abi_version(void) { return OCM_ABI_VERSION; }
extern "C" const char* ocm_last_error(void) { return ocm::g_err; }
extern "C" uint64_t ocm_launch_count(void) { return ocm::g_launches.load(std::memory_order_relaxed); }
