// 2-D float64 FFT convolution building blocks on top of fft_smem.cuh: one line (row of a T x T plane) per CTA,
// row transform -> complex transpose -> row transform.  Spectra stay in digit-reversed order in both dimensions, so
// the product of two spectra needs no permutation.  Used by valley_fft.cu (rotated-kernel bank) and disc.cu (exact
// disc sums of the integer planes).
#pragma once

#include "fft_smem.cuh"

namespace topo {

// Forward transform of one line: load_in(n) = sample n (natural order); the digit-reversed spectrum goes to
// out[0 .. N) with coalesced stores.
template <int N, class Load>
__device__ __forceinline__ void fft2d_forward_line(double2* buf, const double2* __restrict__ tw, int tid, Load load_in,
                                                   double2* __restrict__ out) {
    constexpr int NT = FftShape<N>::NT;
    fft_forward_outer<N>(buf, tw, tid, load_in);
    for (int u = tid; u < N / 8; u += NT) {  // innermost (stride-1) stage in place
        double2 v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = buf[pad(8 * u + q)];
        dft8<false>(v);
#pragma unroll
        for (int q = 0; q < 8; ++q) buf[pad(8 * u + q)] = v[q];
    }
    __syncthreads();
    for (int i = tid; i < N; i += NT) out[i] = buf[pad(i)];
}

// Inverse transform of one digit-reversed line: load_perm(i) = element at position i; every natural-order output n is
// handed to store_out(n, value).  No 1/N.
template <int N, class LoadP, class Store>
__device__ __forceinline__ void fft2d_inverse_line(double2* buf, const double2* __restrict__ tw, int tid, LoadP load_perm,
                                                   Store store_out) {
    constexpr int NT = FftShape<N>::NT;
    for (int i = tid; i < N; i += NT) buf[pad(i)] = load_perm(i);
    __syncthreads();
    for (int u = tid; u < N / 8; u += NT) {
        double2 v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = buf[pad(8 * u + q)];
        dft8<true>(v);
#pragma unroll
        for (int q = 0; q < 8; ++q) buf[pad(8 * u + q)] = v[q];
    }
    __syncthreads();
    fft_inverse_outer<N>(buf, tw, tid, store_out);
}

// second forward pass (and the first one of complex data): [planes][N][N] natural order -> digit-reversed lines
template <int N>
static __global__ void __launch_bounds__(FftShape<N>::NT, N >= 8192 ? 1 : 3)
    fft2d_fwd_cplx_kernel(const double2* __restrict__ src, double2* __restrict__ dst, const double2* __restrict__ tw) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* buf = reinterpret_cast<double2*>(smem_raw);
    const int64_t base = ((int64_t)blockIdx.y * N + blockIdx.x) * N;
    fft2d_forward_line<N>(buf, tw, threadIdx.x, [&](int n) { return __ldg(src + base + n); }, dst + base);
}

// first inverse pass: the line is the product of two spectra (a: [planes][N][N], k: [N][N], same for every plane).
// TS = false: dst[plane][line][n] (a transpose pass follows).  TS = true: the outputs n in [n_lo, n_hi) -- the only ones
// the second pass will turn into pixels -- go straight to the transposed plane dst[plane][n][line]: 16-byte stores
// 16 N bytes apart, whose partner halves come from the CTA of line + 1 in the same wave and merge in L2, so DRAM sees
// whole sectors and the separate transpose pass (a read + a write of the plane) disappears.
template <int N, bool TS>
static __global__ void __launch_bounds__(FftShape<N>::NT, N >= 8192 ? 1 : 3)
    fft2d_inv_product_kernel(const double2* __restrict__ a, const double2* __restrict__ k, double2* __restrict__ dst,
                             const double2* __restrict__ tw, int n_lo, int n_hi) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* buf = reinterpret_cast<double2*>(smem_raw);
    const int64_t base = ((int64_t)blockIdx.y * N + blockIdx.x) * N;
    const double2* __restrict__ kk = k + (int64_t)blockIdx.x * N;
    double2* out = TS ? dst + (int64_t)blockIdx.y * N * N + blockIdx.x : dst + base;
    fft2d_inverse_line<N>(buf, tw, threadIdx.x, [&](int i) { return cmul(__ldg(a + base + i), __ldg(kk + i)); },
                          [&](int n, double2 y) {
                              if constexpr (TS) {
                                  if (n >= n_lo && n < n_hi) out[(int64_t)n * N] = y;
                              } else {
                                  out[n] = y;
                              }
                          });
}

// [planes][n][n] complex transpose, 32 x 32 tiles
static __global__ void __launch_bounds__(256) fft2d_transpose_kernel(const double2* __restrict__ in, double2* __restrict__ out, int n) {
    __shared__ double2 t[32][33];
    const int64_t base = (int64_t)blockIdx.z * n * n;
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8) t[i][threadIdx.x] = __ldg(in + base + (int64_t)(r0 + i) * n + c0 + threadIdx.x);
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) out[base + (int64_t)(c0 + i) * n + r0 + threadIdx.x] = t[threadIdx.x][i];
}

template <int N>
static int fft2d_set_smem_attributes() {
    using S = FftShape<N>;
    TOPO_CUDA(cudaFuncSetAttribute(fft2d_fwd_cplx_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::SMEM));
    TOPO_CUDA(cudaFuncSetAttribute((fft2d_inv_product_kernel<N, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::SMEM));
    TOPO_CUDA(cudaFuncSetAttribute((fft2d_inv_product_kernel<N, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::SMEM));
    return 0;
}

}  // namespace topo
