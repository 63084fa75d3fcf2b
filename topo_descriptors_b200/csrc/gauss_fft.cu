// Wide-radius Gaussian passes by overlap-save convolution with a float64 FFT held in shared memory.
// Reference semantics: scipy.ndimage.gaussian_filter -> correlate1d per axis with float64 weights and accumulation,
// float32 rounding after each axis, mode="reflect" (topo.py:80, 631-635).  The direct kernels of gauss.cu spend
// 2*lw+1 DFMA per pixel and axis on the FP64 pipe (radius 801: 35.8 ms per axis at 16384^2); an N-point transform
// costs ~5 N log2 N flops for N - 2*lw outputs of TWO lines, i.e. ~40x fewer FP64 operations at radius 801, with a
// rounding error (~1e-15 relative, measured in profiles/proto/fft_conv.py) that is as far below the float32 output
// rounding as the direct float64 sum's.
//
// Per CTA: one segment of N samples of a pair of lines, packed as real + i * imaginary (the filter is real, so one
// complex transform filters both).  Forward decimation-in-frequency stages run in place and leave the spectrum in
// digit-reversed order; the multiplier H (the same forward transform of the wrapped kernel, 1/N folded in, computed
// by a one-CTA setup launch) is applied in that order; inverse decimation-in-time stages mirror the forward ones, so
// no permutation pass exists.  Radix 8 (a leading radix 2 or 4 stage when N is not a power of 8); the first stage
// reads global memory directly, the last one writes it, and the innermost stage pair (stride 1) is fused with the
// multiplication -- N = 4096 touches shared memory 12 half-passes instead of 18.  Element i lives at i + (i >> 3):
// every stage then moves 16-byte elements without bank conflicts, including the stride-1 stage.
#include <math.h>

#include "fft_smem.cuh"

namespace topo {

struct FftConvParams {
    const float* in;
    float* out;
    int64_t ld_in, ld_out;
    int n_lines;        // lines of `in` / `out`
    int n_glob;         // global length of the filtered axis (period of the reflection)
    int in0, in_len;    // global positions present in a line of `in`
    int out0, out_len;  // global positions to produce (stored from column 0 of `out`)
    const double* w;    // half kernel, w[0] = centre .. w[lw]  (setup launch only)
    double* H;          // N multipliers, digit-reversed order
    double2* tw;        // tw[i] = exp(-2 pi i / N)
    int lw;
};

// SETUP: transform the wrapped kernel h[n] = w[min(n, N - n)] (0 beyond lw) and store Re / N as the multiplier.
template <int N, bool SETUP>
__global__ void __launch_bounds__(FftShape<N>::NT, N >= 8192 ? 1 : (N <= 2048 ? 4 : 3)) fft_conv_kernel(const FftConvParams p) {
    using S = FftShape<N>;
    constexpr int NT = S::NT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* buf = reinterpret_cast<double2*>(smem_raw);
    const int tid = threadIdx.x;
    const int lw = p.lw;
    const int L = N - 2 * lw;
    const int seg = blockIdx.x;
    const int line_a = 2 * blockIdx.y, line_b = line_a + 1;
    const bool has_b = line_b < p.n_lines;
    const int g0 = p.out0 + seg * L - lw;  // global position of sample n = 0
    const float* ra = p.in + (int64_t)line_a * p.ld_in;
    const float* rb = p.in + (int64_t)(has_b ? line_b : line_a) * p.ld_in;
    const bool interior = g0 >= 0 && g0 + N <= p.n_glob && g0 >= p.in0 && g0 + N <= p.in0 + p.in_len;
    const double2* __restrict__ tw = p.tw;

    fft_forward_outer<N>(buf, tw, tid, [&](int n) -> double2 {
        if constexpr (SETUP) {
            const int m = n < N - n ? n : N - n;
            return make_double2(m <= lw ? p.w[m] : 0.0, 0.0);
        } else {
            int idx;
            if (interior) {
                idx = g0 - p.in0 + n;
            } else {
                idx = reflect_index(g0 + n, p.n_glob) - p.in0;
                idx = idx < 0 ? 0 : (idx >= p.in_len ? p.in_len - 1 : idx);  // only feeds outputs that are not stored
            }
            return make_double2((double)__ldg(ra + idx), has_b ? (double)__ldg(rb + idx) : 0.0);
        }
    });

    // ---- innermost stage (stride 1) forward + multiplier + innermost inverse stage, in registers
    for (int u = tid; u < N / 8; u += NT) {
        double2 v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = buf[pad(8 * u + q)];
        dft8<false>(v);
        if constexpr (SETUP) {
#pragma unroll
            for (int k = 0; k < 8; ++k) p.H[8 * u + k] = v[k].x * (1.0 / N);
        } else {
            const double2* hp = reinterpret_cast<const double2*>(p.H + 8 * u);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const double2 h = __ldg(hp + k);
                v[2 * k].x *= h.x, v[2 * k].y *= h.x;
                v[2 * k + 1].x *= h.y, v[2 * k + 1].y *= h.y;
            }
            dft8<true>(v);
#pragma unroll
            for (int q = 0; q < 8; ++q) buf[pad(8 * u + q)] = v[q];
        }
    }
    if constexpr (SETUP) return;
    __syncthreads();

    const int o_base = seg * L - lw;  // output column of sample n: o_base + n
    float* oa = p.out + (int64_t)line_a * p.ld_out;
    float* ob = p.out + (int64_t)line_b * p.ld_out;
    fft_inverse_outer<N>(buf, tw, tid, [&](int n, double2 y) {
        const int o = o_base + n;
        if (n >= lw && n < lw + L && o < p.out_len) {
            oa[o] = (float)y.x;
            if (has_b) ob[o] = (float)y.y;
        }
    });
}

template <int N>
static int launch_fft_conv(const FftConvParams& p, cudaStream_t s) {
    using S = FftShape<N>;
    static bool attr_set[64] = {false};
    int dev = 0;
    TOPO_CUDA(cudaGetDevice(&dev));
    if (dev >= 64 || !attr_set[dev]) {
        TOPO_CUDA(cudaFuncSetAttribute(fft_conv_kernel<N, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::SMEM));
        TOPO_CUDA(cudaFuncSetAttribute(fft_conv_kernel<N, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::SMEM));
        if (dev < 64) attr_set[dev] = true;
    }
    TOPO_LAUNCH("gauss_fft_twiddles", s, fft_twiddle_kernel<<<ceil_div(N, 256), 256, 0, s>>>(p.tw, N));
    TOPO_LAUNCH("gauss_fft_setup", s, (fft_conv_kernel<N, true><<<1, S::NT, S::SMEM, s>>>(p)));
    const int L = N - 2 * p.lw;
    dim3 grid(ceil_div(p.out_len, L), ceil_div(p.n_lines, 2));
    TOPO_CHECK(grid.y <= 65535, "too many lines for one launch");
    TOPO_LAUNCH("gauss_fft", s, (fft_conv_kernel<N, false><<<grid, S::NT, S::SMEM, s>>>(p)));
    return 0;
}

// Transform length for a radius: the shortest of 2048 / 4096 / 8192 that keeps at least ~60 % of a segment useful.
int fft_conv_length(int lw) {
    if (lw <= 400) return 2048;
    if (lw <= 1000) return 4096;
    if (lw <= 3072) return 8192;
    return 0;
}

size_t fft_conv_table_bytes(int lw) {
    const int n = fft_conv_length(lw);
    return n ? (size_t)n * (sizeof(double2) + sizeof(double)) : 0;
}

// One overlap-save pass along the contiguous axis.  tables: fft_conv_table_bytes(lw) bytes of device memory.
int fft_conv_rows(const float* in, int64_t ld_in, float* out, int64_t ld_out, int n_lines, int n_glob, int in0, int in_len,
                  int out0, int out_len, const double* w, int lw, void* tables, cudaStream_t s) {
    const int n = fft_conv_length(lw);
    TOPO_CHECK(n > 0, "radius %d too wide for the FFT path", lw);
    FftConvParams p;
    p.in = in, p.out = out, p.ld_in = ld_in, p.ld_out = ld_out, p.n_lines = n_lines, p.n_glob = n_glob;
    p.in0 = in0, p.in_len = in_len, p.out0 = out0, p.out_len = out_len, p.w = w, p.lw = lw;
    p.tw = reinterpret_cast<double2*>(tables);
    p.H = reinterpret_cast<double*>(p.tw + n);
    switch (n) {
        case 2048: return launch_fft_conv<2048>(p, s);
        case 4096: return launch_fft_conv<4096>(p, s);
        default: return launch_fft_conv<8192>(p, s);
    }
}

}  // namespace topo
