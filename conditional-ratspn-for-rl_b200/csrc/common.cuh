// Shared helpers for the sm_100a kernels of libspn_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "spn_b200.h"

#define SPN_LOG_SQRT_2PI 0.91893853320467274178f
#define SPN_LOG_2 0.69314718055994530942f

static inline int spn_launch_status() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SPN_OK : -(1000 + (int)e);
}

static inline int spn_cuda_status(cudaError_t e) { return e == cudaSuccess ? SPN_OK : -(1000 + (int)e); }

static inline unsigned spn_blocks(int64_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

__device__ __forceinline__ float spn_softplus(float y) {
    // torch.nn.functional.softplus (beta = 1, threshold = 20)
    return y > 20.f ? y : log1pf(expf(y));
}

// tanh change-of-variables correction of distributions.py:220 evaluated on the pre-squash value x
__device__ __forceinline__ float spn_tanh_correction(float x) {
    return 2.f * (SPN_LOG_2 - x - spn_softplus(-2.f * x));
}

// Philox4x32-10 counter based RNG (Salmon et al. 2011); one call yields four 32-bit words.
__device__ __forceinline__ uint4 spn_philox4x32(uint64_t counter, uint64_t stream_id, uint64_t seed) {
    uint32_t c0 = (uint32_t)counter, c1 = (uint32_t)(counter >> 32);
    uint32_t c2 = (uint32_t)stream_id, c3 = (uint32_t)(stream_id >> 32);
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

// uniform in (0, 1]
__device__ __forceinline__ float spn_u01(uint32_t bits) { return ((float)(bits >> 8) + 1.0f) * (1.0f / 16777216.0f); }

// Gumbel(0,1) noise for virtual element `elem` of noise stream `stream_id`: -log(Exp(1)), Exp(1) = -log(u)
__device__ __forceinline__ float spn_gumbel(uint64_t elem, uint64_t stream_id, uint64_t seed) {
    uint4 r = spn_philox4x32(elem >> 2, stream_id, seed);
    uint32_t bits = (elem & 3) == 0 ? r.x : (elem & 3) == 1 ? r.y : (elem & 3) == 2 ? r.z : r.w;
    float e = -__logf(spn_u01(bits));
    return -__logf(fmaxf(e, 1e-30f));
}

// standard normal via Box-Muller on two words of one Philox block
__device__ __forceinline__ float spn_normal(uint64_t elem, uint64_t stream_id, uint64_t seed) {
    uint4 r = spn_philox4x32(elem >> 1, stream_id, seed);
    float u1 = spn_u01((elem & 1) ? r.z : r.x), u2 = spn_u01((elem & 1) ? r.w : r.y);
    float rad = sqrtf(-2.f * __logf(u1));
    return rad * __cosf(6.283185307179586f * u2);
}

// ------------------------------------------------------------------------------------------------ cp.async (LDGSTS)
// src_bytes < size zero-fills the remainder (src_bytes = 0: pure zero fill, the source is not read).
__device__ __forceinline__ void spn_cp_async16(void* dst_smem, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src),
                 "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void spn_cp_async4(void* dst_smem, const void* src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src),
                 "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void spn_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void spn_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
