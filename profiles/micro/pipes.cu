// Micro-benchmark: per-SM throughput of DFMA, FFMA, F2F.F64.F32, IADD3 (64-bit add), LDS.32 on this B200.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run: ./pipes
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void __launch_bounds__(1024) k(double* out, float* fin, int iters) {
    __shared__ float sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = fin[i];
    __syncthreads();
    double a0 = threadIdx.x, a1 = 1, a2 = 2, a3 = 3, a4 = 4, a5 = 5, a6 = 6, a7 = 7;
    float f0 = threadIdx.x, f1 = 1, f2 = 2, f3 = 3, f4 = 4, f5 = 5, f6 = 6, f7 = 7;
    unsigned long long u0 = threadIdx.x, u1 = 1, u2 = 2, u3 = 3;
    const double w = 1.0000001;
    const float wf = 1.0000001f;
    int idx = threadIdx.x;
    for (int i = 0; i < iters; ++i) {
        if (OP == 0) {
            a0 = fma(a0, w, 1e-9); a1 = fma(a1, w, 1e-9); a2 = fma(a2, w, 1e-9); a3 = fma(a3, w, 1e-9);
            a4 = fma(a4, w, 1e-9); a5 = fma(a5, w, 1e-9); a6 = fma(a6, w, 1e-9); a7 = fma(a7, w, 1e-9);
        } else if (OP == 1) {
            f0 = fmaf(f0, wf, 1e-9f); f1 = fmaf(f1, wf, 1e-9f); f2 = fmaf(f2, wf, 1e-9f); f3 = fmaf(f3, wf, 1e-9f);
            f4 = fmaf(f4, wf, 1e-9f); f5 = fmaf(f5, wf, 1e-9f); f6 = fmaf(f6, wf, 1e-9f); f7 = fmaf(f7, wf, 1e-9f);
        } else if (OP == 2) {  // 8 cvt f32->f64 + 8 DADD to keep them alive
            a0 += (double)f0; a1 += (double)f1; a2 += (double)f2; a3 += (double)f3;
            a4 += (double)f4; a5 += (double)f5; a6 += (double)f6; a7 += (double)f7;
            f0 += 1.f; f1 += 1.f; f2 += 1.f; f3 += 1.f; f4 += 1.f; f5 += 1.f; f6 += 1.f; f7 += 1.f;
        } else if (OP == 3) {  // 8 64-bit integer adds of 32-bit values
            unsigned v = (unsigned)i * 2654435761u;
            u0 += v; u1 += v ^ 1; u2 += v ^ 2; u3 += v ^ 3; u0 += v ^ 4; u1 += v ^ 5; u2 += v ^ 6; u3 += v ^ 7;
        } else if (OP == 4) {  // 8 LDS.32 conflict-free
            f0 += sm[(idx) & 4095]; f1 += sm[(idx + 32) & 4095]; f2 += sm[(idx + 64) & 4095]; f3 += sm[(idx + 96) & 4095];
            f4 += sm[(idx + 128) & 4095]; f5 += sm[(idx + 160) & 4095]; f6 += sm[(idx + 192) & 4095]; f7 += sm[(idx + 224) & 4095];
            idx += 256;
        } else if (OP == 5) {  // 8 DADD
            a0 += w; a1 += w; a2 += w; a3 += w; a4 += w; a5 += w; a6 += w; a7 += w;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + f0 + f1 + f2 + f3 + f4 + f5 + f6 + f7 +
                                                   (double)(u0 + u1 + u2 + u3);
}

template <int OP>
void run(const char* name, int threads) {
    double* out; float* fin;
    cudaMalloc(&out, 148 * 2 * 1024 * sizeof(double));
    cudaMalloc(&fin, 4096 * sizeof(float));
    cudaMemset(fin, 0, 4096 * sizeof(float));
    int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<148, threads>>>(out, fin, 100);
    cudaEventRecord(e0);
    k<OP><<<148, threads>>>(out, fin, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = 8.0 * iters * threads * 148;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%-10s threads/SM=%4d  %.3f ms  %.2f Tops/s  %.1f lanes/clk/SM @%d MHz(nominal max)\n", name, threads, ms,
           ops / ms / 1e9, ops / (ms * 1e-3) / 148 / (clk * 1e3), clk / 1000);
    cudaFree(out); cudaFree(fin);
}

int main() {
    for (int t : {256, 1024}) {
        if (t == 256) { run<0>("DFMA", 256); run<1>("FFMA", 256); run<2>("F2F+DADD", 256); run<3>("IADD64", 256); run<4>("LDS32", 256); run<5>("DADD", 256); }
        else { run<0>("DFMA", 1024); run<1>("FFMA", 1024); run<2>("F2F+DADD", 1024); run<3>("IADD64", 1024); run<4>("LDS32", 1024); run<5>("DADD", 1024); }
    }
    return 0;
}
