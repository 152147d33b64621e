// Which pipe do float<->int conversions use on sm_100a, and do they overlap with FP32 work?
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
template <int MODE>
__global__ void k(float* out, float seed, int iseed, long long* cycles) {
    float x[8]; int q[8]; float y[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = seed * (i + 1) + threadIdx.x; q[i] = iseed + i; y[i] = seed + i; }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0 || MODE == 2 || MODE == 4) asm volatile("cvt.rzi.s32.f32 %0, %1;" : "=r"(q[i]) : "f"(x[i]));
            if (MODE == 1 || MODE == 3 || MODE == 4) asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(y[i]) : "r"(q[i]));
            if (MODE == 2 || MODE == 3 || MODE == 5) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(seed), "f"(1.0f));
            if (MODE == 6) asm volatile("shl.b32 %0, %0, 1;" : "+r"(q[i]));
            if (MODE == 7) { asm volatile("shl.b32 %0, %0, 1;" : "+r"(q[i])); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(seed), "f"(1.0f)); }
        }
    }
    long long t1 = clock64();
    __syncthreads();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i] + y[i] + (float)q[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == blockDim.x - 1) cycles[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char* name, int nops) {
    long long* cyc; float* out;
    cudaMalloc(&cyc, 1024 * 8); cudaMalloc(&out, 148 * 1024 * 4);
    k<MODE><<<148, 512>>>(out, 1.0f, 3, cyc); cudaDeviceSynchronize();
    k<MODE><<<148, 512>>>(out, 1.0f, 3, cyc); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-40s cycles=%8lld  warp-instr/clk/SMSP=%.3f\n", name, h, (double)nops * 8 * ITERS * (512 / 32) / 4.0 / (double)h);
}
int main() {
    run<0>("F2I.TRUNC", 1); run<1>("I2F", 1); run<5>("FFMA", 1); run<2>("F2I + FFMA", 2); run<3>("I2F + FFMA", 2);
    run<4>("F2I + I2F", 2); run<6>("SHL", 1); run<7>("SHL + FFMA", 2);
    return 0;
}
