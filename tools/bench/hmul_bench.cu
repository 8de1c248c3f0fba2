// Micro-benchmark: issue rate of packed half-precision multiplies on one SM (clk per warp instruction per SMSP).
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cstdint>

template <int KIND>
__device__ __forceinline__ uint32_t op(uint32_t a, uint32_t b) {
    uint32_t d;
    if (KIND == 0) asm volatile("mul.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    else if (KIND == 1) asm volatile("mul.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    else if (KIND == 2) { float f; asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(f) : "f"(__uint_as_float(a)), "f"(__uint_as_float(b))); d = __float_as_uint(f); }
    else if (KIND == 3) asm volatile("fma.rn.bf16x2 %0, %1, %2, %2;" : "=r"(d) : "r"(a), "r"(b));
    else if (KIND == 4) asm volatile("fma.rn.f16x2 %0, %1, %2, %2;" : "=r"(d) : "r"(a), "r"(b));
    else { asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(__uint_as_float(a)), "f"(__uint_as_float(b))); }
    return d;
}

template <int KIND>
__global__ void bench(uint32_t* out, long long* clk, int iters) {
    uint32_t v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = threadIdx.x * 7 + i;
    const uint32_t u = out[0];
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = op<KIND>(u, v[i]);
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) s ^= v[i];
    out[1 + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[0] = t1 - t0;
}

template <int KIND>
void run(const char* name, uint32_t* o, long long* c) {
    for (int warps_per_smsp = 1; warps_per_smsp <= 4; warps_per_smsp *= 2) {
        const int threads = 128 * warps_per_smsp, iters = 200;
        bench<KIND><<<1, threads>>>(o, c, iters);
        cudaDeviceSynchronize();
        long long h;
        cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        printf("%-16s %d warps/SMSP: %.2f clk per warp-instruction per SMSP\n", name, warps_per_smsp,
               (double)h / (iters * 32.0 * warps_per_smsp));
    }
}

int main() {
    uint32_t* o; long long* c;
    cudaMalloc(&o, 4096 * 4); cudaMalloc(&c, 8);
    cudaMemset(o, 0, 4096 * 4);
    run<0>("mul.bf16x2", o, c); run<1>("mul.f16x2", o, c); run<2>("mul.f32", o, c); run<3>("fma.bf16x2", o, c);
    run<4>("fma.f16x2", o, c); run<5>("cvt.bf16x2.f32", o, c);
    return 0;
}
