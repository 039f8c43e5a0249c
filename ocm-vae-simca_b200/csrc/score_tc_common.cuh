// PTX wrappers, operand splits and descriptors shared by the tcgen05 scoring kernels (score_tc.cu, score_tc2.cu).
#pragma once

#include "ocm_common.cuh"

#include <cuda.h>

namespace ocm {
namespace tcs {

// ---------------------------------------------------------------- PTX wrappers ------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
template <int OFF>
__device__ __forceinline__ void mbar_arrive_at(uint32_t bar) {      // [bar + OFF]: the offset rides in the instruction
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0+%1];" ::"r"(bar), "n"(OFF) : "memory");
}
// Release of a TMA ring slot AFTER the loads that read it have returned.  The arrive has no data dependence on those
// LDS: ptxas schedules SYNCS.ARRIVE straight behind their issue, and on the B200 the slot then does get refilled under
// a load that is still queued (measured in score_tc2.cu: 30-70 rows in 10 M whose last 16-byte chunk of a box held the
// NEXT occupant of the slot; a fence before the arrive made them disappear).  Here the arrival count is 1 + (loaded bits
// & zero) with `zero` a kernel argument that is always 0 -- a dependence the compiler cannot fold, for one LOP + IADD.
template <int OFF>
__device__ __forceinline__ void mbar_arrive_at_after(uint32_t bar, uint32_t loaded_bits, uint32_t zero) {
    const uint32_t cnt = 1u + (loaded_bits & zero);
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0+%1], %2;" ::"r"(bar), "n"(OFF), "r"(cnt) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait.  A hand-off that never completes (protocol bug, wedged TMA) must neither hang the GPU nor pass
// unnoticed: after `max_polls` polls the waiting thread writes a code into the device's status word (mapped host
// memory, runtime.cu) and TRAPS -- the launch fails, the stream reports cudaErrorLaunchFailure at its next
// synchronisation and every later call of the library returns OCM_ETIMEOUT (the host reads the status word without
// touching the device).  Nothing the launch wrote is handed out as a result.
// Round 2 first carried an abort word through every role (NaN rows for the tiles of the CTA that gave up, warp votes
// around every wait so that no role touches a barrier after an abandoned wait): A/B-timed at +13 % kernel time
// (tools/ab_score.py, same box: 8.87 -> 10.0 ms first launch), for a path that must never run.  The trap costs nothing.
// (Also A/B-tested, twice: a leaner polling loop -- PTX loop around try_wait, 5 instructions per round -- is 10 %
// SLOWER end to end than this one, try_wait suspend hints 6 % slower, a __nanosleep back-off makes no difference.)
struct WaitCtx {
    unsigned int* status;       // device-visible status word (may be NULL)
    uint32_t max_polls;
};
static __device__ __noinline__ void wait_timed_out(unsigned int* status) {
    if (status) { atomicExch(status, 1u); __threadfence_system(); }
    __trap();
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, const WaitCtx& w) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > w.max_polls) wait_timed_out(w.status);
    }
}
__device__ __forceinline__ ulonglong2 lds128(uint32_t addr) {
    ulonglong2 r;
    asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(r.x), "=l"(r.y) : "r"(addr) : "memory");
    return r;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// tcgen05.ld is asynchronous: its destination registers are valid only after tcgen05.wait::ld.  The wait has no data
// dependence on them, so the compiler is free to schedule ordinary arithmetic on those registers between the load and
// the wait.  Passing the registers through an empty volatile asm AFTER the wait pins every later use behind it.
__device__ __forceinline__ void tc_wait_ld16(uint32_t* v) {
    tc_wait_ld();
    asm volatile("" : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                      "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]) :: "memory");
}
__device__ __forceinline__ void tc_wait_ld16x2(uint32_t* a, uint32_t* b) {
    tc_wait_ld();
    asm volatile("" : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]),
                      "+r"(a[8]), "+r"(a[9]), "+r"(a[10]), "+r"(a[11]), "+r"(a[12]), "+r"(a[13]), "+r"(a[14]), "+r"(a[15]) :: "memory");
    asm volatile("" : "+r"(b[0]), "+r"(b[1]), "+r"(b[2]), "+r"(b[3]), "+r"(b[4]), "+r"(b[5]), "+r"(b[6]), "+r"(b[7]),
                      "+r"(b[8]), "+r"(b[9]), "+r"(b[10]), "+r"(b[11]), "+r"(b[12]), "+r"(b[13]), "+r"(b[14]), "+r"(b[15]) :: "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem descriptor], bf16 inputs, fp32 accumulate.  Called by the WHOLE (converged) MMA
// warp; one elected lane issues.  Keeping the control flow warp-uniform lets the compiler hold descriptors and
// addresses in uniform registers -- with an `if (lane == 0)` branch around it every operand went through R2UR
// and the issue loop ran at ~100 cycles per MMA (ncu: converters waiting on the A-buffer-free barrier).
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void tc_commit_elect(uint32_t bar) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
        ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
          "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void cp_async16_hint(uint32_t dst, const void* src, uint64_t policy) {
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "l"(policy) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// packed fp32 pair arithmetic (sm_100 FADD2/FFMA2): one issue slot for two lanes of work
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(uint32_t lo, uint32_t hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
    return r;
}
__device__ __forceinline__ void upk2(f32x2 v, uint32_t& lo, uint32_t& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
// 3-way split of a packed pair: the two remainders are computed with one packed subtract each
__device__ __forceinline__ void split3_pack2(f32x2 x, uint32_t& w1, uint32_t& w2, uint32_t& w3) {
    uint32_t u0, u1, v0, v1, s0, s1;
    upk2(x, u0, u1);
    w1 = __byte_perm(u0, u1, 0x7632);
    const f32x2 r = sub2(x, pk2(u0 & 0xFFFF0000u, u1 & 0xFFFF0000u));
    upk2(r, v0, v1);
    w2 = __byte_perm(v0, v1, 0x7632);
    const f32x2 q = sub2(r, pk2(v0 & 0xFFFF0000u, v1 & 0xFFFF0000u));
    upk2(q, s0, s1);
    w3 = __byte_perm(s0, s1, 0x7632);
}

// 2-way fp16 split of a packed pair y (|y| < 65504): hi = rn_f16(y), lo = rn_f16(y - hi); y - (hi + lo) <= 2^-22 |y|
// (2^-25 absolute in the subnormal range).  Six instructions per pair against nine for the 3-way bf16 split, two TMEM
// stores per stage instead of three and two MMAs per k-step instead of three.
__device__ __forceinline__ void split2_f16_pack2(f32x2 y, uint32_t& whi, uint32_t& wlo) {
    uint32_t y0, y1;
    upk2(y, y0, y1);
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(whi) : "r"(y1), "r"(y0));       // high half <- y1, low half <- y0
    uint32_t h0, h1;
    asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tcvt.f32.f16 %0, lo;\n\tcvt.f32.f16 %1, hi;\n\t}"
        : "=r"(h0), "=r"(h1) : "r"(whi));
    const f32x2 r = sub2(y, pk2(h0, h1));
    uint32_t r0, r1;
    upk2(r, r0, r1);
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(wlo) : "r"(r1), "r"(r0));
}

// 3-way fp16 split of two (prescaled) fp32 values: h1 = rn(x), h2 = rn(x - h1), h3 = rn(x - h1 - h2); 33 bits, or
// 2^-25 absolute once a term is subnormal -- which is why the operands are prescaled (PSHIFT / TSHIFT below)
__device__ __forceinline__ void split3_f16_pack(float x0, float x1, uint32_t& w1, uint32_t& w2, uint32_t& w3) {
    const f32x2 x = pk2(__float_as_uint(x0), __float_as_uint(x1));
    uint32_t lo;
    split2_f16_pack2(x, w1, lo);                  // w1 = rn(x), lo = rn(x - w1) -- but we need the exact remainder for term 3
    uint32_t h0, h1;
    asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tcvt.f32.f16 %0, lo;\n\tcvt.f32.f16 %1, hi;\n\t}"
        : "=r"(h0), "=r"(h1) : "r"(w1));
    const f32x2 r = sub2(x, pk2(h0, h1));         // exact
    split2_f16_pack2(r, w2, w3);                  // w2 = rn(r), w3 = rn(r - w2)
}

// exact 3-way bf16 split of two fp32 values, packed as the tensor core wants them (low half = lower K index):
// b1 = top 16 bits, r1 = x - b1 (exact), b2 = top 16 bits of r1, r2 = r1 - b2 (exact), b3 = top 16 bits of r2
__device__ __forceinline__ void split3_pack(float x0, float x1, uint32_t& w1, uint32_t& w2, uint32_t& w3) {
    const uint32_t u0 = __float_as_uint(x0), u1 = __float_as_uint(x1);
    w1 = __byte_perm(u0, u1, 0x7632);
    const float r0 = x0 - __uint_as_float(u0 & 0xFFFF0000u), r1 = x1 - __uint_as_float(u1 & 0xFFFF0000u);
    const uint32_t v0 = __float_as_uint(r0), v1 = __float_as_uint(r1);
    w2 = __byte_perm(v0, v1, 0x7632);
    const float s0 = r0 - __uint_as_float(v0 & 0xFFFF0000u), s1 = r1 - __uint_as_float(v1 & 0xFFFF0000u);
    w3 = __byte_perm(__float_as_uint(s0), __float_as_uint(s1), 0x7632);
}

// shared-memory matrix descriptor, 128B swizzle (SM100 UMMA SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout_type 2 = SWIZZLE_128B [61,64))
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor (UMMA InstrDescriptor): c=F32 [4,6), a=BF16 [7,10), b=BF16 [10,13), a_negate 13,
// a_major 15, b_major 16 (0 = K-major, 1 = MN-major), N>>3 [17,23), M>>4 [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, bool a_neg, bool b_mn_major, bool f16 = false) {
    // a format [7,10) and b format [10,13): 0 = F16, 1 = BF16.  The two fields exist separately, but an MMA with A = F16
    // from TMEM and B = BF16 from shared memory raises "illegal instruction" on the B200 (measured), so both follow `f16`.
    return (1u << 4) | ((f16 ? 0u : 1u) << 7) | ((f16 ? 0u : 1u) << 10) | ((a_neg ? 1u : 0u) << 13) | ((b_mn_major ? 1u : 0u) << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tmap, int c0, int c1, int c2, uint32_t bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(c2), "r"(bar), "l"(policy) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tmap, int c0, int c1, uint32_t bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(bar), "l"(policy) : "memory");
}


}  // namespace tcs
}  // namespace ocm
