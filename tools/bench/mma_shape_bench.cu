// Micro-benchmark: execution time of tcgen05.mma (M = 128, K = 16, A in TMEM, B in shared memory, no swizzle) as a
// function of N and of the B operand's layout (MN-major vs K-major, strides), chains of 3 K steps per accumulator as in
// the tsum kernels.  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../conditional-ratspn-for-rl_b200/csrc -I../../include
#include <cstdio>
#include "tc_common.cuh"
using namespace tc;

constexpr int NCHAIN = 300;  // chains of 3 MMAs

__global__ void __launch_bounds__(160, 1) bench(long long* out, int N, int bmn, uint32_t lbo, uint32_t sbo, uint32_t kstep_bytes, int ncols) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 4) tmem_alloc(&slot, 512);
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (warp == 4) {
        const uint32_t wsm = smem_u32(smem);
        const uint32_t idesc = instr_desc_bf16(128, N, 0, bmn);
        const int nbuf = (512 - 32) / ncols;
        long long t0 = clock64();
#pragma unroll 1
        for (int s = 0; s < NCHAIN; ++s) {
            const uint32_t acc = (uint32_t)((s % nbuf) * ncols);
            const uint32_t base = wsm + (uint32_t)((s % 4) * 25600);
#pragma unroll
            for (int kk = 0; kk < 3; ++kk)
                mma_ts_uniform(acc, 480 + kk * 8, smem_desc(base + kk * kstep_bytes, lbo, sbo), idesc, kk ? 1u : 0u);
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

void run(const char* name, long long* d, int N, int bmn, uint32_t lbo, uint32_t sbo, uint32_t kstep, int ncols) {
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    long long h[2];
    for (int rep = 0; rep < 2; ++rep) {
        bench<<<1, 160, 200 * 1024>>>(d, N, bmn, lbo, sbo, kstep, ncols);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    }
    printf("%-52s issue %.1f clk/MMA, issue+drain %.1f clk/MMA (floor %d)\n", name, (double)h[0] / (3 * NCHAIN), (double)h[1] / (3 * NCHAIN), N / 2);
}

int main() {
    long long* d;
    cudaMalloc(&d, 16);
    run("N=48  MN-major sbo=128 lbo=640 (round-1 forward)", d, 48, 1, 640, 128, 1280, 64);
    run("N=160 MN-major sbo=128 lbo=2560 (tsum FWD)", d, 160, 1, 2560, 128, 5120, 160);
    run("N=160 K-major  sbo=640 lbo=128 (tsum DX)", d, 160, 0, 128, 640, 256, 160);
    run("N=160 K-major  sbo=256 lbo=128 (dense K-major)", d, 160, 0, 128, 256, 5120, 160);
    run("N=128 MN-major sbo=128 lbo=2048", d, 128, 1, 2048, 128, 4096, 128);
    run("N=128 K-major  sbo=640 lbo=128", d, 128, 0, 128, 640, 256, 128);
    run("N=240 MN-major sbo=128 lbo=3840", d, 240, 1, 3840, 128, 7680, 240);
    run("N=240 K-major  sbo=640 lbo=128", d, 240, 0, 128, 640, 256, 240);
    run("N=80  MN-major sbo=128 lbo=1280", d, 80, 1, 1280, 128, 2560, 80);
    run("N=80  K-major  sbo=640 lbo=128", d, 80, 0, 128, 640, 256, 80);
    run("N=256 MN-major sbo=128 lbo=4096", d, 256, 1, 4096, 128, 8192, 256);
    return 0;
}
