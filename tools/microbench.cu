// Instruction-throughput microbenchmarks that decide the shape of the mix kernel's inner loop
// (SURVEY.md §7 H0): packed FP32x2 vs scalar, float->int conversion cost, round-down add,
// shared-memory gather rate, dependent-add latency. Build: see tools/run_microbench.sh.
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define ILP 8

__device__ __forceinline__ unsigned long long pk(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float lo(unsigned long long v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a; }
__device__ __forceinline__ float hi(unsigned long long v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return b; }

template <int MODE>
__global__ void k_tp(float* out, float seed, long long* cycles) {
    float x[ILP];
    unsigned long long p[ILP];
    int q[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) { x[i] = seed + i + threadIdx.x; p[i] = pk(x[i], x[i] + 1.0f); q[i] = i; }
    const unsigned long long pc = pk(seed, seed * 0.5f), pm = pk(1.0001f, 0.9999f);
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (MODE == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(seed), "f"(1.0f));
            if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pm), "l"(pc));
            if (MODE == 2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pc));
            if (MODE == 3) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pm));
            if (MODE == 4) asm volatile("add.rm.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(seed));
            if (MODE == 5) { asm volatile("cvt.rzi.s32.f32 %0, %1;" : "=r"(q[i]) : "f"(x[i])); asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(seed)); }
            if (MODE == 6) { asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(x[i]) : "r"(q[i])); asm volatile("add.s32 %0, %0, 1;" : "+r"(q[i])); }
            if (MODE == 7) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(seed));
            if (MODE == 8) asm volatile("cvt.rzi.f32.f32 %0, %0;" : "+f"(x[i]));
            if (MODE == 9) asm volatile("add.rm.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pc));
        }
    }
    long long t1 = clock64();
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i] + lo(p[i]) + hi(p[i]) + (float)q[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// dependent-chain latency: MODE 0 scalar add, 1 packed add
template <int MODE>
__global__ void k_lat(float* out, float seed, long long* cycles) {
    float x = seed;
    unsigned long long p = pk(seed, seed + 1.0f), pc = pk(seed, 1.5f);
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            if (MODE == 0) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x) : "f"(seed));
            if (MODE == 1) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p) : "l"(pc));
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = x + lo(p) + hi(p);
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// shared-memory gather: each lane reads smem[(base + lane*stride_num/16 + j) & mask]; MODE 0 = two LDS.32
// (a, a+1), MODE 1 = one LDS.64
template <int MODE>
__global__ void k_lds(float* out, int stride16, long long* cycles) {
    __shared__ float sm[8192 + 8];
    for (int i = threadIdx.x; i < 8192 + 8; i += blockDim.x) sm[i] = (float)i;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    int idx = ((lane * stride16) >> 4) + (threadIdx.x >> 5) * 64;
    float acc = 0.0f;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            int a = (idx + i * 40) & 8191;
            if (MODE == 0) { acc += sm[a]; acc += sm[a + 1]; }
            if (MODE == 1) { float2 v = *reinterpret_cast<const float2*>(&sm[a & ~1]); acc += v.x; acc += v.y; }
        }
        idx += 3;
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <class F>
static void run(const char* name, F launch, int threads, double ops_per_thread_iter) {
    long long* cyc;
    float* out;
    cudaMalloc(&cyc, 1024 * sizeof(long long));
    cudaMalloc(&out, 1024 * 1024 * sizeof(float));
    launch(out, cyc);
    cudaDeviceSynchronize();
    launch(out, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[4];
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double ops = ops_per_thread_iter * ITERS * threads;
    printf("%-44s threads/SM=%4d  cycles=%9lld  thread-ops/clk/SM=%7.2f  (%s)\n", name, threads, h[0], ops / (double)h[0],
           cudaGetErrorString(e));
    cudaFree(cyc);
    cudaFree(out);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs, clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    const int nb = p.multiProcessorCount;  // one CTA per SM
    for (int threads : {128, 256, 512, 1024}) {
        run("FFMA scalar (independent x8)", [&](float* o, long long* c) { k_tp<0><<<nb, threads>>>(o, 1.0f, c); }, threads, ILP);
        run("FFMA2 packed (counted as 1 op)", [&](float* o, long long* c) { k_tp<1><<<nb, threads>>>(o, 1.0f, c); }, threads, ILP);
        run("FADD2 packed", [&](float* o, long long* c) { k_tp<2><<<nb, threads>>>(o, 1.0f, c); }, threads, ILP);
        run("FMUL2 packed", [&](float* o, long long* c) { k_tp<3><<<nb, threads>>>(o, 1.0f, c); }, threads, ILP);
        run("FADD.RM scalar", [&](float* o, long long* c) { k_tp<4><<<nb, threads>>>(o, 1.0f, c); }, threads, ILP);
        run("FADD2.RM packed", [&](float* o, long long* c) { k_tp<9><<<nb, threads>>>(o, 1.0f, c); }, threads, ILP);
        run("FADD scalar", [&](float* o, long long* c) { k_tp<7><<<nb, threads>>>(o, 1.0f, c); }, threads, ILP);
        run("F2I.TRUNC + FADD (2 ops)", [&](float* o, long long* c) { k_tp<5><<<nb, threads>>>(o, 1.0f, c); }, threads, 2 * ILP);
        run("I2F + IADD (2 ops)", [&](float* o, long long* c) { k_tp<6><<<nb, threads>>>(o, 1.0f, c); }, threads, 2 * ILP);
        run("FRND.TRUNC", [&](float* o, long long* c) { k_tp<8><<<nb, threads>>>(o, 1.0f, c); }, threads, ILP);
    }
    run("dependent FADD chain (latency: clk/op = 32/x)", [&](float* o, long long* c) { k_lat<0><<<nb, 32>>>(o, 1.0f, c); }, 32, 16);
    run("dependent FADD2 chain (latency: clk/op = 32/x)", [&](float* o, long long* c) { k_lat<1><<<nb, 32>>>(o, 1.0f, c); }, 32, 16);
    for (int threads : {256, 512, 1024}) {
        for (int s16 : {16, 14, 18, 22, 32, 128}) {
            char nm[96];
            snprintf(nm, sizeof nm, "LDS.32 pair gather, lane stride %.3f", s16 / 16.0);
            run(nm, [&](float* o, long long* c) { k_lds<0><<<nb, threads>>>(o, s16, c); }, threads, 16);
            snprintf(nm, sizeof nm, "LDS.64 gather, lane stride %.3f", s16 / 16.0);
            run(nm, [&](float* o, long long* c) { k_lds<1><<<nb, threads>>>(o, s16, c); }, threads, 8);
        }
    }
    return 0;
}
