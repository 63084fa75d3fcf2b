// 2-D float64 FFT convolution building blocks on top of fft_smem.cuh: one line (row of a T x T plane) per CTA,
// row transform -> complex transpose -> row transform.  Spectra stay in digit-reversed order in both dimensions, so
// the product of two spectra needs no permutation.  Used by valley_fft.cu (rotated-kernel bank) and disc.cu (exact
// disc sums of the integer planes).
#pragma once

#include "fft_smem.cuh"

namespace topo {

// Forward transform of one line: load_in(n) = sample n (natural order); the digit-reversed spectrum goes to
// out[0 .. N) with coalesced stores.
// BY8: position i goes to out[(i & 7) * (N / 8) + (i >> 3)] -- the order in which the first inverse stage of
// fft2d_inverse_line_product wants the multiplier line (thread u owns positions 8u .. 8u+7: with this order its eight
// loads are coalesced across the threads).
template <int N, bool BY8 = false, class Load>
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
    if constexpr (BY8) {
        for (int i = tid; i < N; i += NT) out[i] = buf[pad((i % (N / 8)) * 8 + i / (N / 8))];  // out[q * N/8 + u] = position 8u + q
    } else {
        for (int i = tid; i < N; i += NT) out[i] = buf[pad(i)];
    }
}

// 16-byte asynchronous global -> shared copy (LDGSTS, L1 bypassed): a whole line is put in flight by its CTA without
// holding a register per element -- the synchronous load loop it replaces kept ONE load per thread in flight and paid the
// DRAM latency sixteen times per line (ncu: long-scoreboard stalls 10-19 per issue, issue slots 18-35 % busy).
__device__ __forceinline__ void cp_async_16(double2* smem_dst, const double2* __restrict__ gmem_src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Inverse stages of a line that already sits (digit-reversed, padded, synchronised) in `buf`; every natural-order
// output n is handed to store_out(n, value).  No 1/N.
template <int N, class Store>
__device__ __forceinline__ void fft2d_inverse_staged(double2* buf, const double2* __restrict__ tw, int tid, Store store_out) {
    constexpr int NT = FftShape<N>::NT;
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

// Inverse transform of one digit-reversed line: load_perm(i) = element at position i.
template <int N, class LoadP, class Store>
__device__ __forceinline__ void fft2d_inverse_line(double2* buf, const double2* __restrict__ tw, int tid, LoadP load_perm,
                                                   Store store_out) {
    constexpr int NT = FftShape<N>::NT;
#pragma unroll 8
    for (int i = tid; i < N; i += NT) buf[pad(i)] = load_perm(i);
    __syncthreads();
    fft2d_inverse_staged<N>(buf, tw, tid, store_out);
}

// ... of a line stored contiguously at src[0 .. N): staged with asynchronous copies, all in flight at once.
template <int N, class Store>
__device__ __forceinline__ void fft2d_inverse_line_from(double2* buf, const double2* __restrict__ tw, int tid,
                                                        const double2* __restrict__ src, Store store_out) {
    constexpr int NT = FftShape<N>::NT;
#pragma unroll
    for (int i = 0; i < N / NT; ++i) cp_async_16(buf + pad(tid + i * NT), src + tid + i * NT);
    cp_async_wait_all();
    __syncthreads();
    fft2d_inverse_staged<N>(buf, tw, tid, store_out);
}

// ... of the product a[i] * k[i] of two lines: `a` (contiguous) arrives by asynchronous copies; the multiplier line is
// stored by eights (k8[q * N/8 + u] = k[8u + q], see fft2d_forward_line<BY8>) so that every thread fetches the eight
// factors of its first butterfly with coalesced loads and applies them in registers on the way into the first stage.
template <int N, class Store>
__device__ __forceinline__ void fft2d_inverse_line_product(double2* buf, const double2* __restrict__ tw, int tid,
                                                           const double2* __restrict__ a, const double2* __restrict__ k8,
                                                           Store store_out) {
    constexpr int NT = FftShape<N>::NT, PER = N / NT;
#pragma unroll
    for (int i = 0; i < PER; ++i) cp_async_16(buf + pad(tid + i * NT), a + tid + i * NT);
    double2 kr[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) kr[q] = __ldg(k8 + q * (N / 8) + tid);
    cp_async_wait_all();
    __syncthreads();
    for (int u = tid; u < N / 8; u += NT) {
        double2 v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = cmul(buf[pad(8 * u + q)], kr[q]);
        if (u + NT < N / 8) {  // the next butterfly's factors travel while this one is computed
#pragma unroll
            for (int q = 0; q < 8; ++q) kr[q] = __ldg(k8 + q * (N / 8) + u + NT);
        }
        dft8<true>(v);
#pragma unroll
        for (int q = 0; q < 8; ++q) buf[pad(8 * u + q)] = v[q];
    }
    __syncthreads();
    fft_inverse_outer<N>(buf, tw, tid, store_out);
}

// second forward pass (and the first one of complex data): [planes][lines][N] natural order -> digit-reversed lines
// (BY8: the lines of a multiplier spectrum, in the order fft2d_inverse_line_product reads them)
template <int N, bool BY8 = false>
static __global__ void __launch_bounds__(FftShape<N>::NT, N >= 8192 ? 1 : 3)
    fft2d_fwd_cplx_kernel(const double2* __restrict__ src, double2* __restrict__ dst, const double2* __restrict__ tw) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* buf = reinterpret_cast<double2*>(smem_raw);
    const int64_t base = ((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * N;  // grid (lines, planes)
    fft2d_forward_line<N, BY8>(buf, tw, threadIdx.x, [&](int n) { return __ldg(src + base + n); }, dst + base);
}

// first inverse pass: the line is the product of two spectra (a: [planes][lines][N], k: [lines][N] with its lines
// stored by eights, same for every plane);
// dst[plane][line][n] for n in [n_lo, n_hi) -- the columns the transpose pass will carry over (multiples of 32).
// (Storing the lines already transposed -- 16-byte stores N x 16 bytes apart, one per row -- was measured: the first
// pass grows from 3.2 to 6.3 ms per 25 tiles of 4096^2, more than the 2.2 ms transpose pass it would replace.)
template <int N>
static __global__ void __launch_bounds__(FftShape<N>::NT, N >= 8192 ? 1 : 3)
    fft2d_inv_product_kernel(const double2* __restrict__ a, const double2* __restrict__ k, double2* __restrict__ dst,
                             const double2* __restrict__ tw, int n_lo, int n_hi) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* buf = reinterpret_cast<double2*>(smem_raw);
    // grid (planes, lines): the planes of one line are neighbours in launch order, so the line of k they all multiply
    // by is read from DRAM once and then served by L2 (with the planes outermost it came back from DRAM for every plane)
    const int plane = blockIdx.x, line = blockIdx.y;
    const int64_t base = ((int64_t)plane * gridDim.y + line) * N;
    double2* out = dst + base;
    fft2d_inverse_line_product<N>(buf, tw, threadIdx.x, a + base, k + (int64_t)line * N, [&](int n, double2 y) {
        if (n >= n_lo && n < n_hi) out[n] = y;
    });
}

// [planes][rows][cols] -> [planes][cols][rows] complex transpose, 32 x 32 tiles (rows, cols multiples of 32): grid
// (column tiles, row tiles, planes); blockIdx.x counts column tiles from col_tile0 (a caller that only needs some rows
// of the result transposes only those columns of the source)
static __global__ void __launch_bounds__(256) fft2d_transpose_kernel(const double2* __restrict__ in, double2* __restrict__ out, int rows,
                                                                     int cols, int col_tile0 = 0) {
    __shared__ double2 t[32][33];
    const int64_t base = (int64_t)blockIdx.z * rows * cols;
    const int c0 = (blockIdx.x + col_tile0) * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8) t[i][threadIdx.x] = __ldg(in + base + (int64_t)(r0 + i) * cols + c0 + threadIdx.x);
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) out[base + (int64_t)(c0 + i) * rows + r0 + threadIdx.x] = t[threadIdx.x][i];
}

template <int N>
static int fft2d_set_smem_attributes() {
    using S = FftShape<N>;
    TOPO_CUDA(cudaFuncSetAttribute((fft2d_fwd_cplx_kernel<N, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::SMEM));
    TOPO_CUDA(cudaFuncSetAttribute((fft2d_fwd_cplx_kernel<N, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::SMEM));
    TOPO_CUDA(cudaFuncSetAttribute(fft2d_inv_product_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::SMEM));
    return 0;
}

}  // namespace topo
