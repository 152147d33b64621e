// Can a small kernel run underneath a persistent kernel that takes (almost) all shared memory and 152 regs x 384 threads?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __maxnreg__(152) big(float* out, long long spin) {
    extern __shared__ float sm[];
    sm[threadIdx.x] = threadIdx.x;
    long long t0 = clock64();
    float acc = 0;
    while (clock64() - t0 < spin) acc += sm[(threadIdx.x * 7) & 1023];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void __launch_bounds__(32) small(float* out, long long spin) {
    long long t0 = clock64();
    float acc = 0;
    while (clock64() - t0 < spin) acc += 1.0f;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main(int argc, char** argv) {
    int smem = argc > 1 ? atoi(argv[1]) : 229056;
    int carve = argc > 2 ? atoi(argv[2]) : 1;
    float *a, *b; cudaMalloc(&a, 148 * 384 * 4); cudaMalloc(&b, 2048 * 32 * 4);
    cudaFuncSetAttribute(big, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (carve) cudaFuncSetAttribute(small, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaStream_t s1, s2; cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
    cudaEvent_t e0, e1, e2; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, big); printf("big regs %d, small ", fa.numRegs); cudaFuncGetAttributes(&fa, small); printf("regs %d\n", fa.numRegs);
    for (int rep = 0; rep < 3; rep++) {
        cudaDeviceSynchronize();
        cudaEventRecord(e0, s1);
        big<<<148, 384, smem, s1>>>(a, 190000);        // ~100 us
        cudaEventRecord(e1, s1);
        cudaStreamWaitEvent(s2, e0, 0);
        small<<<2048, 32, 0, s2>>>(b, 28000);           // 2048 warps x ~15 us each
        cudaEventRecord(e2, s2);
        cudaDeviceSynchronize();
        float t1, t2; cudaEventElapsedTime(&t1, e0, e1); cudaEventElapsedTime(&t2, e0, e2);
        printf("smem %d carve %d: big done at %.1f us, small done at %.1f us  (%s)\n", smem, carve, t1 * 1e3, t2 * 1e3, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
