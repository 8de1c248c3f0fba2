// Micro-benchmarks behind the round-2 kernel decisions (one SM, clk per warp instruction per SMSP / bytes per clk):
//   * fp32 pipe: fma.rn.f32 (three register operands), fma.rn.f32x2 / mul.rn.f32x2 (packed, sm_100), mixed
//   * TMEM: tcgen05.ld / tcgen05.st 32x32b throughput with 4 / 8 / 16 warps (bytes per clk per SM)
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipe_bench pipe_bench.cu
#include <cstdint>
#include <cstdio>

template <int KIND>
__global__ void fbench(float* out, long long* clk, int iters) {
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = threadIdx.x * 0.001f + i;
    const float a = out[0], b = out[1];
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (KIND == 0) {
#pragma unroll
            for (int i = 0; i < 32; ++i) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(v[i]) : "f"(a), "f"(b));
        } else if (KIND == 1) {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
                asm volatile(
                    "{\n\t.reg .b64 x, y, z;\n\tmov.b64 x, {%0, %1};\n\tmov.b64 y, {%2, %2};\n\tmov.b64 z, {%3, %3};\n\t"
                    "fma.rn.f32x2 x, x, y, z;\n\tmov.b64 {%0, %1}, x;\n\t}"
                    : "+f"(v[i]), "+f"(v[i + 1]) : "f"(a), "f"(b));
            }
        } else if (KIND == 2) {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
                asm volatile(
                    "{\n\t.reg .b64 x, y;\n\tmov.b64 x, {%0, %1};\n\tmov.b64 y, {%2, %2};\n\t"
                    "mul.rn.f32x2 x, x, y;\n\tmov.b64 {%0, %1}, x;\n\t}"
                    : "+f"(v[i]), "+f"(v[i + 1]) : "f"(a));
            }
        } else if (KIND == 3) {   // fma with distinct registers: d = v[i] * v[j] + v[k]
#pragma unroll
            for (int i = 0; i < 32; ++i) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(v[i]) : "f"(v[(i + 7) & 31]), "f"(v[(i + 13) & 31]));
        } else if (KIND == 6) {   // fma.f32x2 with three DISTINCT register pairs: acc[i] = v[j] * v[k] + acc[i] (dX epilogue pattern)
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
                const int j = (i + 6) & 31, k = (i + 14) & 31;
                asm volatile(
                    "{\n\t.reg .b64 x, y, z;\n\tmov.b64 x, {%0, %1};\n\tmov.b64 y, {%2, %3};\n\tmov.b64 z, {%4, %5};\n\t"
                    "fma.rn.f32x2 x, y, z, x;\n\tmov.b64 {%0, %1}, x;\n\t}"
                    : "+f"(v[i]), "+f"(v[i + 1]) : "f"(v[j]), "f"(v[j + 1]), "f"(v[k]), "f"(v[k + 1]));
            }
        } else if (KIND == 7) {   // fma.f32x2 with a broadcast scalar and two distinct pairs: acc[i] = a * v[k] + acc[i] (forward epilogue)
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
                const int k = (i + 14) & 31;
                asm volatile(
                    "{\n\t.reg .b64 x, y, z;\n\tmov.b64 x, {%0, %1};\n\tmov.b64 y, {%2, %2};\n\tmov.b64 z, {%3, %4};\n\t"
                    "fma.rn.f32x2 x, y, z, x;\n\tmov.b64 {%0, %1}, x;\n\t}"
                    : "+f"(v[i]), "+f"(v[i + 1]) : "f"(a), "f"(v[k]), "f"(v[k + 1]));
            }
        } else if (KIND == 8) {   // add.f32x2, distinct pairs
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
                const int k = (i + 14) & 31;
                asm volatile(
                    "{\n\t.reg .b64 x, z;\n\tmov.b64 x, {%0, %1};\n\tmov.b64 z, {%2, %3};\n\t"
                    "add.rn.f32x2 x, x, z;\n\tmov.b64 {%0, %1}, x;\n\t}"
                    : "+f"(v[i]), "+f"(v[i + 1]) : "f"(v[k]), "f"(v[k + 1]));
            }
        } else if (KIND == 4) {
#pragma unroll
            for (int i = 0; i < 32; ++i) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(v[i]) : "f"(a));
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += v[i];
    out[2 + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[0] = t1 - t0;
}

template <int KIND>
void frun(const char* name, float* o, long long* c, int per_iter) {
    for (int wps = 1; wps <= 4; wps *= 2) {
        const int threads = 128 * wps, iters = 200;
        fbench<KIND><<<1, threads>>>(o, c, iters);
        cudaDeviceSynchronize();
        long long h;
        cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        printf("%-28s %d warps/SMSP: %.2f clk per warp-instruction per SMSP\n", name, wps, (double)h / (iters * (double)per_iter * wps));
    }
}

// ---------------------------------------------------------------------------------------------------------------- TMEM
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X, bool STORE, bool VARY>
__global__ void tbench(uint32_t* out, long long* clk, int iters) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base0 = slot + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t woff = (uint32_t)((warp >> 2) * 128);   // warps of one quadrant start in different column ranges
    uint32_t r[X];
#pragma unroll
    for (int i = 0; i < X; ++i) r[i] = threadIdx.x + i;
    __syncthreads();
    long long t0 = clock64();
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
        const uint32_t base = base0 + (VARY ? (woff + (uint32_t)it * X) % (512u - X) : woff % (512u - X));
        if (STORE) {
            if constexpr (X == 32) {
                asm volatile(
                    "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(base),
                    "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                    "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
                    "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
                    "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
                    : "memory");
            }
        } else {
            if constexpr (X == 32) {
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                      "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                      "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                      "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(base)
                    : "memory");
            } else if constexpr (X == 16) {
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                      "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                    : "r"(base)
                    : "memory");
            }
            if ((it & 3) == 3) {   // consume every 4th batch: up to 4 loads in flight per warp
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int i = 0; i < X; ++i) acc ^= r[i];
            }
        }
    }
    if (STORE) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    else asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    long long t1 = clock64();
#pragma unroll
    for (int i = 0; i < X; ++i) acc ^= r[i];
    out[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x == 0) clk[0] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
}

template <int X, bool STORE, bool VARY>
void trun(const char* name, uint32_t* o, long long* c) {
    for (int warps = 4; warps <= 16; warps *= 2) {
        const int iters = 2000;
        tbench<X, STORE, VARY><<<1, warps * 32>>>(o, c, iters);
        cudaError_t e = cudaDeviceSynchronize();
        long long h;
        cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        const double bytes = (double)iters * warps * 32 * X * 4;
        printf("%-20s x%-3d %2d warps: %.1f bytes/clk/SM  (%.1f clk per warp instruction)  [%s]\n", name, X, warps, bytes / h,
               (double)h / iters, cudaGetErrorString(e));
    }
}

int main() {
    float* o; long long* c;
    cudaMalloc(&o, 8192 * 4); cudaMalloc(&c, 8);
    cudaMemset(o, 0, 8192 * 4);
    frun<0>("fma.f32 (a, b uniform)", o, c, 32);
    frun<3>("fma.f32 (3 distinct regs)", o, c, 32);
    frun<1>("fma.f32x2", o, c, 16);
    frun<2>("mul.f32x2", o, c, 16);
    frun<6>("fma.f32x2 (3 distinct pairs)", o, c, 16);
    frun<7>("fma.f32x2 (bcast, 2 pairs)", o, c, 16);
    frun<8>("add.f32x2 (2 distinct pairs)", o, c, 16);
    frun<4>("mul.f32", o, c, 32);
    frun<5>("ex2.approx.f32", o, c, 32);
    trun<32, false, false>("tcgen05.ld same addr", (uint32_t*)o, c);
    trun<32, false, true>("tcgen05.ld moving addr", (uint32_t*)o, c);
    trun<16, false, false>("tcgen05.ld same addr", (uint32_t*)o, c);
    trun<16, false, true>("tcgen05.ld moving addr", (uint32_t*)o, c);
    trun<32, true, false>("tcgen05.st same addr", (uint32_t*)o, c);
    trun<32, true, true>("tcgen05.st moving addr", (uint32_t*)o, c);
    return 0;
}
