// Probe which (box, coordinate) combinations a 2-D float32 TMA box load accepts on sm_100a.
// usage: tma_probe nx rows ld box_w box_h x y [static_bytes]
#include <cstdio>
#include <cstdlib>
#include "../../topo_descriptors_b200/csrc/tma.cuh"
using namespace topo;

__global__ void probe(const __grid_constant__ CUtensorMap tmap, int x, int y, int bytes, float* out, int n) {
    extern __shared__ __align__(128) unsigned char raw[];
    __shared__ __align__(8) uint64_t bar;
    float* tile = reinterpret_cast<float*>(raw);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    __syncthreads();
    if (threadIdx.x == 0) { mbar_expect_tx(&bar, bytes); tma_load_2d(tile, &tmap, x, y, &bar); }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = tile[i];
}

int main(int argc, char** argv) {
    if (argc < 8) return 2;
    long nx = atol(argv[1]), rows = atol(argv[2]), ld = atol(argv[3]);
    int bw = atoi(argv[4]), bh = atoi(argv[5]), x = atoi(argv[6]), y = atoi(argv[7]);
    float* d; cudaMalloc(&d, ld * rows * 4);
    float* h = (float*)malloc(ld * rows * 4);
    for (long i = 0; i < ld * rows; ++i) h[i] = (float)(i % 9973) + 1.f;
    cudaMemcpy(d, h, ld * rows * 4, cudaMemcpyHostToDevice);
    float* o; cudaMalloc(&o, bw * bh * 4);
    CUtensorMap t;
    if (!make_tmap_2d_f32(&t, d, nx, rows, ld, bw, bh)) { printf("ENCODE_FAIL\n"); return 0; }
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    probe<<<1, 128, bw * bh * 4>>>(t, x, y, bw * bh * 4, o, bw * bh);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("FAIL %s\n", cudaGetErrorString(e)); return 0; }
    float* r = (float*)malloc(bw * bh * 4);
    cudaMemcpy(r, o, bw * bh * 4, cudaMemcpyDeviceToHost);
    long bad = 0;
    for (int j = 0; j < bh; ++j) for (int i = 0; i < bw; ++i) {
        long gx = x + i, gy = y + j;
        float want = (gx >= 0 && gx < nx && gy >= 0 && gy < rows) ? h[gy * ld + gx] : 0.f;
        if (r[j * bw + i] != want) ++bad;
    }
    printf(bad ? "WRONG %ld\n" : "OK\n", bad);
    return 0;
}
