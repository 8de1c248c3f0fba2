// Thin inline-PTX layer for the Blackwell (sm_100a) tensor-core path: mbarrier, bulk async copy (TMA engine, 1-D),
// TMEM allocation / load / store, tcgen05.mma issue + commit, and the UMMA descriptor encodings we use.
// Bit layouts follow the PTX ISA tcgen05 "matrix descriptor" / "instruction descriptor" tables.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

// suspend-time hint of mbarrier.try_wait (ns) and the number of tries before a wait gives up (~2 s in total)
#ifndef SPN_MBAR_HINT_NS
#define SPN_MBAR_HINT_NS 0   // measured: no hint is 3 - 10 % faster on every tensor-core kernel (profiles/README.md, round 2)
#endif
#define SPN_MBAR_TRIES (2000000000ull / (SPN_MBAR_HINT_NS > 1000 ? SPN_MBAR_HINT_NS : 1000))

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ------------------------------------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// try_wait with a suspend-time hint: the hardware parks the thread until the phase completes or the hint expires, so a
// waiting warp does not burn issue slots on a polling loop (18 warps share four schedulers).
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
        : "memory");
    return ok != 0;
}
// Bounded wait: gives up (and raises the CTA-wide abort flag) after ~2 s so a protocol bug can never hang the GPU.
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, volatile int* abort_flag) {
    if (mbar_try_wait(bar, parity)) return true;
    const long long t0 = clock64();
    while (true) {
        if (mbar_try_wait(bar, parity)) return true;
        if (*abort_flag) return false;
        if (clock64() - t0 > 4000000000LL) {
            *abort_flag = 1;
            return false;
        }
    }
}

// Wait as ONE opaque PTX block (spin loop, suspend-time hint, bounded): the compiler sees no loop, no branch and no
// result, so control flow around it stays provably warp-uniform and ptxas keeps loop counters and the tcgen05 descriptors
// derived from them in UNIFORM registers (with a C-level loop / bool result every later MMA cost ~14 R2UR / VOTE / ELECT
// instructions, ~56 clk).  A timeout (protocol bug; ~2 s) raises the CTA-wide abort flag in shared memory; from then on
// every wait falls through at once, the kernel drains with garbage results and the host sees the abort count.
__device__ __forceinline__ void mbar_wait_sticky(uint64_t* bar, uint32_t parity, volatile int* abort_flag) {
#if SPN_MBAR_HINT_NS > 0
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .u32 cnt, flg;\n\t"
        "mov.u32 cnt, 0;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra LAB_DONE;\n\t"
        "ld.volatile.shared.u32 flg, [%3];\n\t"
        "setp.ne.u32 p, flg, 0;\n\t"
        "@p bra LAB_DONE;\n\t"
        "add.u32 cnt, cnt, 1;\n\t"
        "setp.lt.u32 p, cnt, %4;\n\t"
        "@p bra LAB_WAIT;\n\t"
        "st.volatile.shared.u32 [%3], 1;\n\t"
        "LAB_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity), "r"((uint32_t)SPN_MBAR_HINT_NS), "r"(smem_u32((const void*)abort_flag)), "r"((uint32_t)SPN_MBAR_TRIES)
        : "memory");
#else
    // No suspend-time hint: with a hint ptxas follows a failed TRYWAIT by NANOSLEEP.SYNCS, whose wake-up granularity is far
    // coarser than the hand-off latencies of these pipelines.  The plain try_wait still parks the warp for the hardware's own
    // (short) window, so the loop does not flood the issue slots; the abort flag is polled every 64 rounds.
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .u32 cnt, flg, low;\n\t"
        "mov.u32 cnt, 0;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra LAB_DONE;\n\t"
        "add.u32 cnt, cnt, 1;\n\t"
        "and.b32 low, cnt, 63;\n\t"
        "setp.ne.u32 p, low, 0;\n\t"
        "@p bra LAB_WAIT;\n\t"
        "ld.volatile.shared.u32 flg, [%2];\n\t"
        "setp.ne.u32 p, flg, 0;\n\t"
        "@p bra LAB_DONE;\n\t"
        "setp.lt.u32 p, cnt, %3;\n\t"
        "@p bra LAB_WAIT;\n\t"
        "st.volatile.shared.u32 [%2], 1;\n\t"
        "LAB_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(smem_u32((const void*)abort_flag)), "r"(1u << 26)
        : "memory");
#endif
}

// Teardown check of the tensor-core kernels: a CTA whose abort flag is up (timed-out wait, or a TMEM allocation that did
// not start at column 0) kills the launch.  __trap() makes the error sticky for the context, so training cannot go on
// with wrong log-likelihoods or gradients.
__device__ __forceinline__ void tc_abort_trap(volatile int* abort_flag, int* counter) {
    if (*abort_flag) {
        atomicAdd(counter, 1);
        __threadfence_system();
        __trap();
    }
}

// Non-blocking probe of an mbarrier phase, returned as a warp-uniform value (lane 0's result, broadcast by a shuffle so
// that the compiler knows a branch on it is not divergent).  Used to learn EARLY -- before a long blocking instruction
// sequence -- that the next hand-off has already happened, so that its ~90-clk try_wait latency is not paid in series.
__device__ __forceinline__ bool mbar_probe_uniform(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return __shfl_sync(0xffffffffu, ok, 0) != 0;
}

// ------------------------------------------------------------------------------------------------ bulk copy (TMA, 1-D)
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// bulk copy shared -> global (TMA engine), tracked with bulk async-groups
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until at most N of this thread's bulk groups have not finished READING their shared-memory source
template <int N>
__device__ __forceinline__ void bulk_wait_group_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_group() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

// Warp-collective variants: every lane executes the statement, one elected lane issues.  Unlike `if (lane == 0) ...` these
// add no divergent branch to the surrounding code (a divergent branch anywhere in a pipeline loop makes ptxas give up the
// uniform-register code path of the whole kernel, see mbar_wait_sticky).
__device__ __forceinline__ void bulk_g2s_elect(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
        "@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t}" ::"r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void bulk_s2g_elect(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile(
        "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
        "@q cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n\t}" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_elect(uint64_t* bar, uint32_t bytes) {
    asm volatile(
        "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
        "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes)
        : "memory");
}

// make generic-proxy shared-memory writes visible to the async proxy (tensor core / TMA reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ------------------------------------------------------------------------------------------------ tcgen05 fences
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------ TMEM alloc
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ------------------------------------------------------------------------------------------------ TMEM <-> registers
// 32x32b shape: thread t of warp w touches TMEM lane 32*(w%4)+t; xN = N consecutive 32-bit columns.
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
                 "r"(r[3])
                 : "memory");
}
// store NCOLS (multiple of 4, <= 64) consecutive columns
template <int NCOLS>
__device__ __forceinline__ void tmem_store_cols(uint32_t taddr, const uint32_t* r) {
    static_assert(NCOLS % 4 == 0 && NCOLS <= 64, "unsupported column count");
    int done = 0;
    if constexpr (NCOLS >= 32) { tmem_st32(taddr, r); done = 32; }
    if constexpr (NCOLS >= 64) { tmem_st32(taddr + 32, r + 32); done = 64; }
    if constexpr ((NCOLS % 32) >= 16) { tmem_st16(taddr + done, r + done); done += 16; }
    if constexpr ((NCOLS % 16) >= 8) { tmem_st8(taddr + done, r + done); done += 8; }
    if constexpr ((NCOLS % 8) >= 4) { tmem_st4(taddr + done, r + done); done += 4; }
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
        "[%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
// load NCOLS (even, <= 32) consecutive columns of this thread's TMEM lane (asynchronous: tmem_wait_ld before use)
template <int NCOLS>
__device__ __forceinline__ void tmem_load_cols(uint32_t taddr, uint32_t* r) {
    static_assert(NCOLS % 2 == 0 && NCOLS <= 32, "unsupported column count");
    int done = 0;
    if constexpr (NCOLS == 32) { tmem_ld32(taddr, r); done = 32; }
    if constexpr (NCOLS < 32 && NCOLS >= 16) { tmem_ld16(taddr, r); done = 16; }
    if constexpr (NCOLS < 32 && (NCOLS % 16) >= 8) { tmem_ld8(taddr + done, r + done); done += 8; }
    if constexpr (NCOLS < 32 && (NCOLS % 8) >= 4) { tmem_ld4(taddr + done, r + done); done += 4; }
    if constexpr (NCOLS < 32 && (NCOLS % 4) >= 2) { tmem_ld2(taddr + done, r + done); done += 2; }
}

// ------------------------------------------------------------------------------------------------ packed fp32 (FFMA2)
// Blackwell issues two fp32 FMAs per lane and instruction on 64-bit register pairs (fma.rn.f32x2); a scalar operand is
// expressed as a pair with both halves equal, which ptxas turns into the broadcast form of FFMA2.
__device__ __forceinline__ uint64_t f2pack(float lo, float hi) {
    uint64_t v;
    asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(lo), "f"(hi));
    return v;
}
__device__ __forceinline__ uint64_t u2pack(uint32_t lo, uint32_t hi) {
    uint64_t v;
    asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "r"(lo), "r"(hi));
    return v;
}
__device__ __forceinline__ void f2unpack(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// ------------------------------------------------------------------------------------------------ descriptors
// Shared-memory matrix descriptor, no swizzle.  Operand memory is a grid of 128-byte "core matrices"
// (8 rows x 16 bytes).  For a K-major operand the rows are M/N indices and the 16 bytes are 8 K-elements:
//   LBO = byte stride between the two K-halves of one K=16 step, SBO = byte stride between 8-row groups.
// For an MN-major operand the rows are K indices and the 16 bytes are 8 consecutive M/N elements:
//   SBO = byte stride between 8-element groups along M/N, LBO = byte stride between 8-row groups along K.
// (cute/arch/mma_sm100_desc.hpp SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48),
//  layout_type [61,64) = 0 for SWIZZLE_NONE.)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// Instruction descriptor for kind::f16 with BF16 A/B and FP32 accumulate (InstrDescriptor bitfield):
// c_format=F32 [4,6)=1, a_format=BF16 [7,10)=1, b_format=BF16 [10,13)=1, a_major [15], b_major [16] (1 = MN-major),
// n_dim = N>>3 [17,23), m_dim = M>>4 [24,29).
__host__ __device__ constexpr uint32_t instr_desc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[tmem] * B[smem]    (A: 128 lanes x K=16 bf16 packed in 8 columns)
__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
// Same, for warp-uniform callers: every lane executes the call with identical operands and one elected lane issues.
// Keeping the control flow uniform lets ptxas hold the descriptors in uniform registers instead of running an
// elect/broadcast loop per instruction (measured: ~150 clk per MMA with a divergent `if (lane == 0)` issue branch).
__device__ __forceinline__ void mma_ts_uniform(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void mma_ss_uniform(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void mma_commit_uniform(uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar))
        : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void mma_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
// mbarrier arrive once all previously issued MMAs of this thread have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// ------------------------------------------------------------------------------------------------ bf16 helpers
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t bcast_lo(uint32_t v) { return __byte_perm(v, v, 0x1010); }
__device__ __forceinline__ uint32_t bcast_hi(uint32_t v) { return __byte_perm(v, v, 0x3232); }
__device__ __forceinline__ uint32_t hmul2_bf16(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }

}  // namespace tc
