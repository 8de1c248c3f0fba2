// Micro-benchmark: how fast can one warp issue tcgen05.mma (M=128, N=48, K=16, A in TMEM, B in smem) under different
// coding styles?  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../conditional-ratspn-for-rl_b200/csrc -I../../include
#include <cstdio>
#include "tc_common.cuh"
using namespace tc;

constexpr int NM = 1000;  // MMAs per measurement (100 stages of 10)
constexpr uint32_t IDESC = instr_desc_bf16(128, 48, 0, 1);

__device__ __forceinline__ void issue10(uint32_t a0, uint64_t d0, uint32_t s) {
    asm volatile(
        "{\n\t.reg .pred q, p;\n\t.reg .b64 d;\n\t.reg .b32 ta;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "mov.b64 d, %1;\n\tmov.b32 ta, %0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%5], [ta], d, %2, {%3, %3, %3, %3}, p;\n\t"
        "setp.ne.b32 p, 1, 0;\n\t"
        "add.s64 d, d, 80;\n\tadd.s32 ta, ta, 8;\n\t@q tcgen05.mma.cta_group::1.kind::f16 [%5], [ta], d, %2, {%3, %3, %3, %3}, p;\n\t"
        "add.s64 d, d, 80;\n\tadd.s32 ta, ta, 8;\n\t@q tcgen05.mma.cta_group::1.kind::f16 [%5], [ta], d, %2, {%3, %3, %3, %3}, p;\n\t"
        "add.s64 d, d, 80;\n\tadd.s32 ta, ta, 8;\n\t@q tcgen05.mma.cta_group::1.kind::f16 [%5], [ta], d, %2, {%3, %3, %3, %3}, p;\n\t"
        "add.s64 d, d, 80;\n\tadd.s32 ta, ta, 8;\n\t@q tcgen05.mma.cta_group::1.kind::f16 [%5], [ta], d, %2, {%3, %3, %3, %3}, p;\n\t"
        "add.s64 d, d, 80;\n\tadd.s32 ta, ta, 8;\n\t@q tcgen05.mma.cta_group::1.kind::f16 [%5], [ta], d, %2, {%3, %3, %3, %3}, p;\n\t"
        "add.s64 d, d, 80;\n\tadd.s32 ta, ta, 8;\n\t@q tcgen05.mma.cta_group::1.kind::f16 [%5], [ta], d, %2, {%3, %3, %3, %3}, p;\n\t"
        "add.s64 d, d, 80;\n\tadd.s32 ta, ta, 8;\n\t@q tcgen05.mma.cta_group::1.kind::f16 [%5], [ta], d, %2, {%3, %3, %3, %3}, p;\n\t"
        "add.s64 d, d, 80;\n\tadd.s32 ta, ta, 8;\n\t@q tcgen05.mma.cta_group::1.kind::f16 [%5], [ta], d, %2, {%3, %3, %3, %3}, p;\n\t"
        "add.s64 d, d, 80;\n\tadd.s32 ta, ta, 8;\n\t@q tcgen05.mma.cta_group::1.kind::f16 [%5], [ta], d, %2, {%3, %3, %3, %3}, p;\n\t"
        "}" ::"r"(a0), "l"(d0), "r"(IDESC), "r"(0u), "r"(s ? 1u : 0u), "r"(0u)
        : "memory");
}

template <int STYLE>
__global__ void __launch_bounds__(160, 1) bench(long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint64_t bars[6];
    __shared__ int slot2;
    __shared__ uint32_t slot;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int i = 0; i < 6; ++i) mbar_init(&bars[i], 1); slot2 = 0; mbar_fence_init(); }
    if (warp == 4) tmem_alloc(&slot, 512);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (warp == 4) {
        const uint32_t wsm = smem_u32(smem);
        long long t0 = clock64();
        if (STYLE == 0) {  // elect inside every MMA (current kernels)
#pragma unroll 1
            for (int s = 0; s < NM / 10; ++s) {
#pragma unroll
                for (int kk = 0; kk < 10; ++kk)
                    mma_ts_uniform(0, 128 + (s & 3) * 80 + kk * 8, smem_desc(wsm + (uint32_t)((s % 10) * 10 + 2 * kk) * 640, 640, 128), IDESC, (s | kk) ? 1u : 0u);
            }
        } else if (STYLE == 1) {  // one divergent branch around the whole loop
            if (lane == 0) {
#pragma unroll 1
                for (int s = 0; s < NM / 10; ++s) {
#pragma unroll
                    for (int kk = 0; kk < 10; ++kk)
                        mma_ts(0, 128 + (s & 3) * 80 + kk * 8, smem_desc(wsm + (uint32_t)((s % 10) * 10 + 2 * kk) * 640, 640, 128), IDESC, (s | kk) ? 1u : 0u);
                }
            }
            __syncwarp();
        } else if (STYLE == 2) {  // lane-0 branch per stage
#pragma unroll 1
            for (int s = 0; s < NM / 10; ++s) {
                if (lane == 0) {
#pragma unroll
                    for (int kk = 0; kk < 10; ++kk)
                        mma_ts(0, 128 + (s & 3) * 80 + kk * 8, smem_desc(wsm + (uint32_t)((s % 10) * 10 + 2 * kk) * 640, 640, 128), IDESC, (s | kk) ? 1u : 0u);
                }
                __syncwarp();
            }
        } else if (STYLE == 3) {  // one asm block per stage: one elect, descriptors advanced inside PTX
#pragma unroll 1
            for (int s = 0; s < NM / 10; ++s) {
                const uint64_t d0 = smem_desc(wsm + (uint32_t)((s % 10) * 10) * 640, 640, 128);
                const uint32_t a0 = 128 + (s & 3) * 80;
                issue10(a0, d0, (uint32_t)s);
            }
        }
        else if (STYLE >= 4) {
            // 4: + one tcgen05.commit per stage; 5: + a probe (try_wait) before every stage; 6: two commits per stage;
            // 7: commit + probe + a blocking wait on an already completed barrier
#pragma unroll 1
            for (int s = 0; s < NM / 10; ++s) {
                const uint64_t d0 = smem_desc(wsm + (uint32_t)((s % 10) * 10) * 640, 640, 128);
                const uint32_t a0 = 128 + (s & 3) * 80;
                bool rdy = false;
                if (STYLE == 5 || STYLE == 7) rdy = mbar_probe_uniform(&bars[4], 1);
                issue10(a0, d0, (uint32_t)s);
                mma_commit_uniform(&bars[s & 3]);
                if (STYLE == 6) mma_commit_uniform(&bars[5]);
                if (STYLE == 7 && !rdy) mbar_wait_sticky(&bars[4], 1, (volatile int*)&slot2);
            }
        }
        long long t1 = clock64();
        mma_commit_uniform(&bar);
        mbar_wait_sticky(&bar, 0, (volatile int*)&slot);
        long long t2 = clock64();
        if (lane == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (warp == 4) tmem_dealloc(0, 512);
}

template <int STYLE>
void run(long long* d) {
    cudaFuncSetAttribute(bench<STYLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    long long h[2];
    for (int rep = 0; rep < 2; ++rep) {
        bench<STYLE><<<1, 160, 200 * 1024>>>(d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("style %d: %s\n", STYLE, cudaGetErrorString(e)); return; }
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    }
    printf("style %d: issue %.1f clk/MMA, issue+drain %.1f clk/MMA\n", STYLE, (double)h[0] / NM, (double)h[1] / NM);
}

int main() {
    long long* d;
    cudaMalloc(&d, 16);
    run<0>(d); run<1>(d); run<2>(d); run<3>(d); run<4>(d); run<5>(d); run<6>(d); run<7>(d);
    return 0;
}
