// Shared helpers for the ocm_simca CUDA translation units (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ocm_simca.h"

namespace ocm {

constexpr int kWarp = 32;

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int  sm_count();
unsigned int* tc_status_device_ptr();      // device-visible status word of the tcgen05 kernels (mapped host memory)
unsigned int* tc_flag_device_ptr();        // device word of the fp16-mode fix-up (see runtime.cu); NULL if it cannot be had
int  tc_status_take();                     // read from the host (no device access); 0 = clean
int  tc_status_report(const char* where);  // 0, or OCM_ETIMEOUT with the error text set

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    count_launch();
    return 0;
}

#define OCM_REQUIRE(cond, code, msg)                 \
    do {                                             \
        if (!(cond)) {                               \
            ::ocm::set_error("%s: %s", __func__, msg); \
            return (code);                           \
        }                                            \
    } while (0)

#define OCM_CUDA(call)                                                     \
    do {                                                                   \
        cudaError_t e_ = (call);                                           \
        if (e_ != cudaSuccess) {                                           \
            ::ocm::set_error("%s: %s", #call, cudaGetErrorString(e_));     \
            return (int)e_;                                                \
        }                                                                  \
    } while (0)

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// streaming 128-bit load that does not pollute L1 (the spectra are read exactly once)
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
    float4 r;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ float rule_distance(int rule, float a, float b) {
    // utils/SIMCA.py:131-144
    if (rule == OCM_RULE_SIM) return fmaxf(a, b);
    if (rule == OCM_RULE_ALT) return sqrtf(a * a + b * b);
    return a + b;  // CI and DD
}

}  // namespace ocm
