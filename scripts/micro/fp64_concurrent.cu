// Microbenchmark: do the FP64 SIMT pipe (DFMA) and the FP64 tensor pipe (DMMA
// m8n8k4) of one SM run concurrently, or do they share execution resources?
// Also: shared-memory LDS.64 / LDS.128 throughput next to DFMA.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// mode 0: all warps DFMA; 1: all warps DMMA; 2: even warps DFMA, odd warps DMMA
__global__ void k_mix(double* out, int iters, int mode) {
    const int warp = threadIdx.x >> 5;
    const bool do_dmma = mode == 1 || (mode == 2 && (warp & 1));
    double acc[16];
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    double a = 1.0000001, b = 1e-9;
    if (do_dmma) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) dmma(acc[2 * i], acc[2 * i + 1], a, b);
        }
    } else {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
        }
    }
    double s = 0;
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// shared memory read throughput: each thread reads `iters` x 8 doubles (LDS.64 or LDS.128)
template <int VEC>
__global__ void k_lds(double* out, int iters) {
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i;
    __syncthreads();
    double s = 0;
    int base = threadIdx.x * VEC;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int idx = (base + j * 1024 * VEC / 2 + it) & 8191 & ~(VEC - 1);
            if (VEC == 1) s += sm[idx];
            else { double2 v = *reinterpret_cast<double2*>(&sm[idx]); s += v.x + v.y; }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    double* out;
    cudaMalloc(&out, sizeof(double) * 148 * 4 * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    const int threads = 512, blocks = 148 * 2, iters = 20000;
    for (int mode = 0; mode < 3; ++mode) {
        k_mix<<<blocks, threads>>>(out, 100, mode);
        cudaEventRecord(e0);
        k_mix<<<blocks, threads>>>(out, iters, mode);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        const double warps = (double)threads / 32 * blocks;
        double fl_dfma = 2.0 * 16 * 32 * iters, fl_dmma = 2.0 * 256 * 8 * iters;   // per warp
        double flops = mode == 0 ? warps * fl_dfma : mode == 1 ? warps * fl_dmma
                                                               : warps / 2 * (fl_dfma + fl_dmma);
        printf("mode %d (%s): %.3f ms, %.2f TFLOP/s total\n", mode,
               mode == 0 ? "DFMA only" : mode == 1 ? "DMMA only" : "half DFMA + half DMMA", ms,
               flops / ms / 1e9);
    }
    for (int vec = 1; vec <= 2; ++vec) {
        const int it2 = 4000;
        cudaEventRecord(e0);
        if (vec == 1) k_lds<1><<<blocks, threads, 65536>>>(out, it2);
        else k_lds<2><<<blocks, threads, 65536>>>(out, it2);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double bytes = (double)blocks * threads * it2 * 8 * 8 * vec;
        printf("LDS.%d: %.3f ms, %.1f TB/s aggregate = %.1f B/clk/SM at 1.965 GHz\n", 64 * vec, ms,
               bytes / ms / 1e9, bytes / (ms * 1e-3) / 148 / 1.965e9);
    }
    printf("status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
