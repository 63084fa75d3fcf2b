// Micro-benchmark 4: FP64 throughput of the vector pipe (DFMA) vs the tensor pipe (mma.sync m8n8k4 f64, "DMMA")
// and of both interleaved -- can a float64 FIR filter use the two at the same time on B200?
//   OP 0: 64 independent DFMA per iteration per thread
//   OP 1: 8 independent DMMA m8n8k4 per iteration per warp (each = 256 FMA = 8 FMA per thread)
//   OP 2: both (64 DFMA + 8 DMMA per iteration)
// Reports FMA-equivalents per clock per SM.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int OP>
__global__ void __launch_bounds__(256) k(double* out, int iters, double seed) {
    double acc[8], c[16];
    for (int i = 0; i < 8; ++i) acc[i] = seed * i;
    for (int i = 0; i < 16; ++i) c[i] = seed * (i + 1);
    double a = seed + threadIdx.x, b = seed * 0.5 + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
        if (OP == 0 || OP == 2) {
#pragma unroll
            for (int s = 0; s < 8; ++s)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] = fma(a, b, acc[j]);
        }
        if (OP == 1 || OP == 2) {
#pragma unroll
            for (int j = 0; j < 8; ++j) dmma(c[2 * j], c[2 * j + 1], a, b);
        }
    }
    double r = 0;
    for (int i = 0; i < 8; ++i) r += acc[i];
    for (int i = 0; i < 16; ++i) r += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int OP>
void run(const char* name, double fma_per_thread_iter) {
    const int blocks = 148 * 4, threads = 256, iters = 20000;
    double* out;
    cudaMalloc(&out, blocks * threads * sizeof(double));
    k<OP><<<blocks, threads>>>(out, 100, 1e-9);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<OP><<<blocks, threads>>>(out, iters, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    int clk_khz;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double fmas = (double)blocks * threads * iters * fma_per_thread_iter;
    printf("%-28s %8.3f ms  %7.2f TFLOP/s  %6.1f FMA/clk/SM (at %d MHz nominal)\n", name, ms, 2 * fmas / ms / 1e9,
           fmas / (ms * 1e-3) / 148 / (clk_khz * 1e3), clk_khz / 1000);
    cudaFree(out);
}

int main() {
    run<0>("DFMA only", 64);
    run<1>("DMMA m8n8k4 only", 64);   // 8 mma x 256 FMA / 32 threads
    run<2>("DFMA + DMMA interleaved", 128);
    return 0;
}
