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

#include "common.cuh"

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

__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(fma(a.x, b.x, -(a.y * b.y)), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ double2 cmul_conj(double2 a, double2 b) {  // a * conj(b)
    return make_double2(fma(a.x, b.x, a.y * b.y), fma(a.y, b.x, -(a.x * b.y)));
}

// 8-point DFT in registers, natural order in and out.  INV: conjugate kernel (no 1/8).
template <bool INV>
__device__ __forceinline__ void dft8(double2 (&v)[8]) {
    constexpr double c = 0.70710678118654752440;
    auto rot = [](double2 a) {  // * (-i) forward, * (+i) inverse
        return INV ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x);
    };
    auto w8 = [&](double2 a) {  // * exp(-+ i pi / 4)
        return INV ? make_double2((a.x - a.y) * c, (a.x + a.y) * c) : make_double2((a.x + a.y) * c, (a.y - a.x) * c);
    };
    double2 a0 = cadd(v[0], v[4]), a4 = csub(v[0], v[4]);
    double2 a1 = cadd(v[1], v[5]), a5 = w8(csub(v[1], v[5]));
    double2 a2 = cadd(v[2], v[6]), a6 = rot(csub(v[2], v[6]));
    double2 a3 = cadd(v[3], v[7]), a7 = rot(w8(csub(v[3], v[7])));
    double2 b0 = cadd(a0, a2), b2 = csub(a0, a2);
    double2 b1 = cadd(a1, a3), b3 = rot(csub(a1, a3));
    double2 b4 = cadd(a4, a6), b6 = csub(a4, a6);
    double2 b5 = cadd(a5, a7), b7 = rot(csub(a5, a7));
    v[0] = cadd(b0, b1), v[4] = csub(b0, b1);
    v[2] = cadd(b2, b3), v[6] = csub(b2, b3);
    v[1] = cadd(b4, b5), v[5] = csub(b4, b5);
    v[3] = cadd(b6, b7), v[7] = csub(b6, b7);
}

template <bool INV>
__device__ __forceinline__ void dft4(double2 (&v)[4]) {
    auto rot = [](double2 a) { return INV ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x); };
    double2 a0 = cadd(v[0], v[2]), a2 = csub(v[0], v[2]);
    double2 a1 = cadd(v[1], v[3]), a3 = rot(csub(v[1], v[3]));
    v[0] = cadd(a0, a1), v[2] = csub(a0, a1);
    v[1] = cadd(a2, a3), v[3] = csub(a2, a3);
}

__device__ __forceinline__ int pad(int i) { return i + (i >> 3); }

template <int N>
struct FftShape {
    static constexpr int LEAD = (N == 8192 || N == 1024) ? 2 : (N == 2048 ? 4 : 1);  // N = LEAD * 8^k
    static constexpr int NT = N >= 8192 ? 512 : 256;                                  // threads per CTA
    static constexpr int M8 = N / LEAD;                                               // length the radix-8 stages start from
    static constexpr size_t SMEM = (size_t)(N + N / 8) * sizeof(double2);
};

// SETUP: transform the wrapped kernel h[n] = w[min(n, N - n)] (0 beyond lw) and store Re / N as the multiplier.
template <int N, bool SETUP>
__global__ void __launch_bounds__(FftShape<N>::NT, N >= 8192 ? 1 : 2) fft_conv_kernel(const FftConvParams p) {
    using S = FftShape<N>;
    constexpr int NT = S::NT, LEAD = S::LEAD, M8 = S::M8;
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

    auto load_in = [&](int n) -> double2 {
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
    };

    // ---- forward, leading radix-2 / radix-4 stage straight from global memory
    if constexpr (LEAD == 2) {
        constexpr int St = N / 2;
        for (int j = tid; j < St; j += NT) {
            const double2 x0 = load_in(j), x1 = load_in(j + St);
            buf[pad(j)] = cadd(x0, x1);
            buf[pad(j + St)] = cmul(csub(x0, x1), tw[j]);
        }
        __syncthreads();
    } else if constexpr (LEAD == 4) {
        constexpr int St = N / 4;
        for (int j = tid; j < St; j += NT) {
            double2 v[4] = {load_in(j), load_in(j + St), load_in(j + 2 * St), load_in(j + 3 * St)};
            dft4<false>(v);
            buf[pad(j)] = v[0];
#pragma unroll
            for (int k = 1; k < 4; ++k) buf[pad(j + k * St)] = cmul(v[k], tw[j * k]);
        }
        __syncthreads();
    }

    // ---- forward radix-8 stages, sub-transform length M = M8, M8/8, ... 64 (stride St = M / 8 >= 8)
    {
        int M = M8;
        if constexpr (LEAD == 1) {  // first stage reads global memory
            constexpr int St = N / 8;
            for (int j = tid; j < St; j += NT) {
                double2 v[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) v[q] = load_in(j + q * St);
                dft8<false>(v);
                buf[pad(j)] = v[0];
#pragma unroll
                for (int k = 1; k < 8; ++k) buf[pad(j + k * St)] = cmul(v[k], tw[j * k]);
            }
            __syncthreads();
            M = N / 8;
        }
        for (; M > 8; M >>= 3) {
            const int St = M >> 3, stride = N / M;
            for (int u = tid; u < N / 8; u += NT) {
                const int j = u & (St - 1), base = (u - j) * 8 + j;  // block * M + j with block = u / St
                double2 v[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) v[q] = buf[pad(base + q * St)];
                dft8<false>(v);
                buf[pad(base)] = v[0];
                const int t1 = j * stride;
#pragma unroll
                for (int k = 1; k < 8; ++k) buf[pad(base + k * St)] = cmul(v[k], tw[t1 * k]);
            }
            __syncthreads();
        }
    }

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

    // ---- inverse radix-8 stages M = 64, 512, ... (mirror of the forward ones); the last one (M = N when there is no
    // leading stage) writes global memory
    const int o_base = seg * L - lw;  // output column of sample n: o_base + n
    float* oa = p.out + (int64_t)line_a * p.ld_out;
    float* ob = p.out + (int64_t)line_b * p.ld_out;
    auto store_out = [&](int n, double2 y) {
        const int o = o_base + n;
        if (n >= lw && n < lw + L && o < p.out_len) {
            oa[o] = (float)y.x;
            if (has_b) ob[o] = (float)y.y;
        }
    };
    for (int M = 64; M <= M8; M <<= 3) {
        const int St = M >> 3, stride = N / M;
        const bool to_global = (LEAD == 1) && (M == N);
        for (int u = tid; u < N / 8; u += NT) {
            const int j = u & (St - 1), base = (u - j) * 8 + j;
            double2 v[8];
            v[0] = buf[pad(base)];
            const int t1 = j * stride;
#pragma unroll
            for (int k = 1; k < 8; ++k) v[k] = cmul_conj(buf[pad(base + k * St)], tw[t1 * k]);
            dft8<true>(v);
            if (to_global) {
#pragma unroll
                for (int q = 0; q < 8; ++q) store_out(base + q * St, v[q]);
            } else {
#pragma unroll
                for (int q = 0; q < 8; ++q) buf[pad(base + q * St)] = v[q];
            }
        }
        if (!to_global) __syncthreads();
    }

    // ---- inverse leading stage to global memory
    if constexpr (LEAD == 2) {
        constexpr int St = N / 2;
        for (int j = tid; j < St; j += NT) {
            const double2 y0 = buf[pad(j)], y1 = cmul_conj(buf[pad(j + St)], tw[j]);
            store_out(j, cadd(y0, y1));
            store_out(j + St, csub(y0, y1));
        }
    } else if constexpr (LEAD == 4) {
        constexpr int St = N / 4;
        for (int j = tid; j < St; j += NT) {
            double2 v[4];
            v[0] = buf[pad(j)];
#pragma unroll
            for (int k = 1; k < 4; ++k) v[k] = cmul_conj(buf[pad(j + k * St)], tw[j * k]);
            dft4<true>(v);
#pragma unroll
            for (int q = 0; q < 4; ++q) store_out(j + q * St, v[q]);
        }
    }
}

__global__ void fft_twiddle_kernel(double2* __restrict__ tw, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s, c;
    sincospi(-2.0 * (double)i / (double)n, &s, &c);
    tw[i] = make_double2(c, s);
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
