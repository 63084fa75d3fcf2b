// Shared-memory float64 FFT building blocks (sm_100a): in-place radix-8 stages (a leading radix 2 / 4 stage when N is
// not a power of 8).  Forward = decimation in frequency, natural order in, DIGIT-REVERSED order out; inverse =
// decimation in time, digit-reversed in, natural out; a point-wise product of two spectra taken in digit-reversed
// order therefore needs no permutation pass (profiles/proto/fft_conv.py pins the index conventions).  Element i of a
// line lives at i + (i >> 3) in shared memory: every stage moves 16-byte elements without bank conflicts, including
// the stride-1 stage.  Used by gauss_fft.cu (overlap-save Gaussian passes) and valley_fft.cu (2-D correlation bank).
#pragma once

#include <math.h>

#include "common.cuh"

namespace topo {

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

// 6-point DFT (two 3-point transforms + a radix-2 combination), natural order in and out.  INV: conjugate kernel.
template <bool INV>
__device__ __forceinline__ void dft3(double2 a, double2 b, double2 c, double2& x0, double2& x1, double2& x2) {
    constexpr double h = 0.86602540378443864676;  // sqrt(3) / 2
    const double2 s = cadd(b, c), d = csub(b, c);
    x0 = cadd(a, s);
    const double2 m = make_double2(a.x - 0.5 * s.x, a.y - 0.5 * s.y);
    const double2 t = make_double2(h * d.x, h * d.y);
    // forward: x1 = m - i t, x2 = m + i t; inverse: the other way round
    const double2 p = make_double2(m.x + t.y, m.y - t.x), q = make_double2(m.x - t.y, m.y + t.x);
    x1 = INV ? q : p;
    x2 = INV ? p : q;
}

template <bool INV>
__device__ __forceinline__ void dft6(double2 (&v)[6]) {
    constexpr double h = 0.86602540378443864676;
    double2 e[3], o[3];
    dft3<INV>(v[0], v[2], v[4], e[0], e[1], e[2]);
    dft3<INV>(v[1], v[3], v[5], o[0], o[1], o[2]);
    // w6^k o[k], w6 = exp(-+ i pi / 3) = (1/2, -+ sqrt(3)/2), w6^2 = (-1/2, -+ sqrt(3)/2)
    const double sg = INV ? 1.0 : -1.0;
    const double2 o1 = make_double2(0.5 * o[1].x - sg * h * o[1].y, 0.5 * o[1].y + sg * h * o[1].x);
    const double2 o2 = make_double2(-0.5 * o[2].x - sg * h * o[2].y, -0.5 * o[2].y + sg * h * o[2].x);
    v[0] = cadd(e[0], o[0]), v[3] = csub(e[0], o[0]);
    v[1] = cadd(e[1], o1), v[4] = csub(e[1], o1);
    v[2] = cadd(e[2], o2), v[5] = csub(e[2], o2);
}

__device__ __forceinline__ int pad(int i) { return i + (i >> 3); }

// w^1 .. w^7 from ONE table entry: the stages are bound by shared-memory / L1 traffic (ncu: l1tex 92 %, FP64 pipe
// 18 %), so six complex multiplications are cheaper than six more 16-byte table loads per butterfly; the powers carry
// ~3 ulp instead of 0.5 (the transform's error stays ~1e-15 relative)
__device__ __forceinline__ void twiddle_powers(double2 w1, double2 (&w)[8]) {
    w[1] = w1;
    w[2] = cmul(w1, w1);
    w[3] = cmul(w[2], w1);
    w[4] = cmul(w[2], w[2]);
    w[5] = cmul(w[4], w1);
    w[6] = cmul(w[3], w[3]);
    w[7] = cmul(w[4], w[3]);
}

template <int N>
struct FftShape {
    static constexpr int LEAD = (N == 8192 || N == 1024) ? 2 : (N == 2048 ? 4 : (N == 3072 ? 6 : 1));  // N = LEAD * 8^k
    static constexpr int NT = N >= 8192 ? 512 : 256;                                  // threads per CTA
    static constexpr int M8 = N / LEAD;                                               // length the radix-8 stages start from
    static constexpr size_t SMEM = (size_t)(N + N / 8) * sizeof(double2);
};


// Forward stages from natural-order input down to sub-transforms of length 8 (the stride-1 stage is left to the
// caller, who usually fuses it with a point-wise operation).  load_in(n) -> sample n.  Ends with __syncthreads().
template <int N, class Load>
__device__ __forceinline__ void fft_forward_outer(double2* buf, const double2* __restrict__ tw, int tid, Load load_in) {
    using S = FftShape<N>;
    constexpr int NT = S::NT, LEAD = S::LEAD, M8 = S::M8;
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
            const double2 w1 = tw[j], w2 = cmul(w1, w1), w3 = cmul(w2, w1);
            buf[pad(j + St)] = cmul(v[1], w1);
            buf[pad(j + 2 * St)] = cmul(v[2], w2);
            buf[pad(j + 3 * St)] = cmul(v[3], w3);
        }
        __syncthreads();
    }
    else if constexpr (LEAD == 6) {
        constexpr int St = N / 6;
        for (int j = tid; j < St; j += NT) {
            double2 v[6];
#pragma unroll
            for (int q = 0; q < 6; ++q) v[q] = load_in(j + q * St);
            dft6<false>(v);
            buf[pad(j)] = v[0];
            const double2 w1 = tw[j], w2 = cmul(w1, w1), w3 = cmul(w2, w1), w4 = cmul(w2, w2), w5 = cmul(w4, w1);
            buf[pad(j + St)] = cmul(v[1], w1);
            buf[pad(j + 2 * St)] = cmul(v[2], w2);
            buf[pad(j + 3 * St)] = cmul(v[3], w3);
            buf[pad(j + 4 * St)] = cmul(v[4], w4);
            buf[pad(j + 5 * St)] = cmul(v[5], w5);
        }
        __syncthreads();
    }
    int M = M8;
    if constexpr (LEAD == 1) {  // first radix-8 stage reads the input directly
        constexpr int St = N / 8;
        for (int j = tid; j < St; j += NT) {
            double2 v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = load_in(j + q * St);
            dft8<false>(v);
            buf[pad(j)] = v[0];
            double2 w[8];
            twiddle_powers(tw[j], w);
#pragma unroll
            for (int k = 1; k < 8; ++k) buf[pad(j + k * St)] = cmul(v[k], w[k]);
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
            double2 w[8];
            twiddle_powers(tw[j * stride], w);
#pragma unroll
            for (int k = 1; k < 8; ++k) buf[pad(base + k * St)] = cmul(v[k], w[k]);
        }
        __syncthreads();
    }
}

// Inverse stages from sub-transforms of length 8 (already done by the caller, buffer synchronised) up to N; the
// last stage hands every output to store_out(n, value) instead of shared memory.  No 1/N.
template <int N, class Store>
__device__ __forceinline__ void fft_inverse_outer(double2* buf, const double2* __restrict__ tw, int tid, Store store_out) {
    using S = FftShape<N>;
    constexpr int NT = S::NT, LEAD = S::LEAD, M8 = S::M8;
    for (int M = 64; M <= M8; M <<= 3) {
        const int St = M >> 3, stride = N / M;
        const bool to_out = (LEAD == 1) && (M == N);
        for (int u = tid; u < N / 8; u += NT) {
            const int j = u & (St - 1), base = (u - j) * 8 + j;
            double2 v[8];
            v[0] = buf[pad(base)];
            double2 w[8];
            twiddle_powers(tw[j * stride], w);
#pragma unroll
            for (int k = 1; k < 8; ++k) v[k] = cmul_conj(buf[pad(base + k * St)], w[k]);
            dft8<true>(v);
            if (to_out) {
#pragma unroll
                for (int q = 0; q < 8; ++q) store_out(base + q * St, v[q]);
            } else {
#pragma unroll
                for (int q = 0; q < 8; ++q) buf[pad(base + q * St)] = v[q];
            }
        }
        if (!to_out) __syncthreads();
    }
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
            const double2 w1 = tw[j], w2 = cmul(w1, w1), w3 = cmul(w2, w1);
            v[1] = cmul_conj(buf[pad(j + St)], w1);
            v[2] = cmul_conj(buf[pad(j + 2 * St)], w2);
            v[3] = cmul_conj(buf[pad(j + 3 * St)], w3);
            dft4<true>(v);
#pragma unroll
            for (int q = 0; q < 4; ++q) store_out(j + q * St, v[q]);
        }
    } else if constexpr (LEAD == 6) {
        constexpr int St = N / 6;
        for (int j = tid; j < St; j += NT) {
            double2 v[6];
            v[0] = buf[pad(j)];
            const double2 w1 = tw[j], w2 = cmul(w1, w1), w3 = cmul(w2, w1), w4 = cmul(w2, w2), w5 = cmul(w4, w1);
            v[1] = cmul_conj(buf[pad(j + St)], w1);
            v[2] = cmul_conj(buf[pad(j + 2 * St)], w2);
            v[3] = cmul_conj(buf[pad(j + 3 * St)], w3);
            v[4] = cmul_conj(buf[pad(j + 4 * St)], w4);
            v[5] = cmul_conj(buf[pad(j + 5 * St)], w5);
            dft6<true>(v);
#pragma unroll
            for (int q = 0; q < 6; ++q) store_out(j + q * St, v[q]);
        }
    }
}

// tw[i] = exp(-2 pi i / n), i = 0 .. n-1
static __global__ void fft_twiddle_kernel(double2* __restrict__ tw, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s, c;
    sincospi(-2.0 * (double)i / (double)n, &s, &c);
    tw[i] = make_double2(c, s);
}

}  // namespace topo
