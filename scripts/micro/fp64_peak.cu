// Microbenchmark: FP64 FMA (SIMT) vs DMMA m8n8k4 throughput on one GPU.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double* out, int iters) {
    double a[16];
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3 + i;
    double b = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fma(a[i], b, c);
    }
    double s = 0;
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int NT>
__global__ void k_dmma(double* out, int iters) {
    double acc[NT][2];
    for (int i = 0; i < NT; ++i) { acc[i][0] = 0; acc[i][1] = 0; }
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NT; ++i) dmma(acc[i][0], acc[i][1], a, b);
    }
    double s = 0;
    for (int i = 0; i < NT; ++i) s += acc[i][0] + acc[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    double* out;
    cudaMalloc(&out, sizeof(double) * 148 * 8 * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    for (int threads : {128, 256, 512, 1024}) {
        for (int bps : {1, 2}) {
            int blocks = 148 * bps;
            int iters = 20000;
            k_dfma<<<blocks, threads>>>(out, 100);
            cudaEventRecord(e0);
            k_dfma<<<blocks, threads>>>(out, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            double flops = 2.0 * 16 * iters * (double)threads * blocks;
            printf("DFMA threads %4d blocks/SM %d: %.2f TFLOP/s\n", threads, bps, flops / ms / 1e9);
            k_dmma<8><<<blocks, threads>>>(out, 100);
            cudaEventRecord(e0);
            k_dmma<8><<<blocks, threads>>>(out, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            flops = 2.0 * 256 * 8 * iters * (double)(threads / 32) * blocks;
            printf("DMMA(8 acc) threads %4d blocks/SM %d: %.2f TFLOP/s\n", threads, bps, flops / ms / 1e9);
            k_dmma<32><<<blocks, threads>>>(out, 100);
            cudaEventRecord(e0);
            k_dmma<32><<<blocks, threads>>>(out, iters / 4);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            flops = 2.0 * 256 * 32 * (iters / 4) * (double)(threads / 32) * blocks;
            printf("DMMA(32 acc) threads %4d blocks/SM %d: %.2f TFLOP/s\n", threads, bps, flops / ms / 1e9);
        }
    }
    cudaError_t e = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
