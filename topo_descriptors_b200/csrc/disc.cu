// TPI and STD: zero-padded disc sums (reference: topo.py:144-181 tpi, 272-307 std, 191-213
// circular_kernel; scipy.signal.convolve(mode="same") centring).
//
// Method (north star: "per-row prefix-sum span-sum kernel"):
//   every DEM value becomes one or more unsigned 32-bit integers (fixed point / exact integer parts),
//   each row gets an exclusive prefix sum in WRAP-AROUND uint32 arithmetic, and a disc sum is
//   sum over kernel rows of  P[row][hi+1] - P[row][lo]  (exact modulo 2^32, and the true span sum
//   is < 2^32 by construction of the scale).  The row differences are accumulated in 32-bit integers
//   when the whole disc sum provably fits (one IADD3 per pixel and kernel row), else in 64-bit integers.
//   The epilogue is float64 with an exact integer numerator.  Integer arithmetic makes the result
//   independent of tiling and of the row-band partition (bit-identical on 1 or 8 GPUs).
//
// Modes (what the uint32 planes hold), t = trunc(z) (the reference's astype("int32"), topo.py:300),
// frac = z - t:
//   TPI_I : t - tmin                                  1 plane, exact (integer-valued DEM)
//   TPI_Q : q = rn(z * 2^S) - c0*2^S                  1 plane, |error of the mean| <= 2^-(S+1)
//   TPI_X : t - tmin,  (frac + 1) * 2^Sf              2 planes, exact (used when S would be < 10)
//   STD_I : t - tmin,  (t - cmid)^2                   2 planes, exact (integer-valued DEM)
//   STD_F : t - tmin,  (t - cmid)^2, (frac+1)*2^Sf    3 planes, exact
//
// Execution shapes (DESIGN.md section 4 has the measurements):
//   tiny     : odd sizes 5..13, one-plane modes: register sliding sums, no prefix at all.
//   fused    : CTA = 128 x 32 output tile; tile + halo is loaded with 128-bit loads, converted, scanned with warp
//              shuffles into shared memory; a thread owns 2 pixels x RB rows; conflict-free LDS.
//   two-pass : prefix planes go to HBM in two copies, P and P shifted by one element, so that the pair
//              (P[j], P[j+1]) is always one aligned 64-bit load; the walk gathers them through L1/L2 with LDG.64
//              (measured 29.9 words/clk/SM vs 19.5 for LDG.32).  CTAs are rasterised in 16-tile-wide super-columns
//              so that the rows in flight stay L2-resident.  Odd discs use the hybrid decomposition: an O(1) core
//              (inscribed square, or -- in multi-size sweeps -- an octagon from sheared summed-area tables) + row
//              caps + column caps (+ corner diagonals from diagonal prefix tables), see hybrid_walk.
//   plane cache (topo_disc_cache): every plane and table above is independent of the disc size, so a multi-size
//              sweep builds them once, for the halo of its largest size.
#include <math.h>

#include <algorithm>
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "common.cuh"
#include "fft2d.cuh"

namespace topo {

// PL_*: single planes, used by the two-pass path which processes one plane per launch
// PL_QL / PL_QH: the centred square split into its low / high 16 bits, for DEMs whose size x range^2 overflows the
// 32-bit span sums of PL_Q (an Alpine 0..4800 m range from size ~800, 0..8848 m from size ~400)
enum DiscMode { TPI_Q = 0, TPI_X = 1, STD_I = 2, STD_F = 3, TPI_I = 4, PL_T = 5, PL_Q = 6, PL_F = 7, PL_QL = 8, PL_QH = 9 };

template <int MODE>
struct ModeTraits;
template <>
struct ModeTraits<TPI_Q> { static constexpr int NARR = 1; static constexpr int RB = 8; };
template <>
struct ModeTraits<TPI_I> { static constexpr int NARR = 1; static constexpr int RB = 8; };
template <>
struct ModeTraits<TPI_X> { static constexpr int NARR = 2; static constexpr int RB = 4; };
template <>
struct ModeTraits<STD_I> { static constexpr int NARR = 2; static constexpr int RB = 4; };
template <>
struct ModeTraits<STD_F> { static constexpr int NARR = 3; static constexpr int RB = 4; };
template <>
struct ModeTraits<PL_T> { static constexpr int NARR = 1; static constexpr int RB = 8; };
template <>
struct ModeTraits<PL_Q> { static constexpr int NARR = 1; static constexpr int RB = 8; };
template <>
struct ModeTraits<PL_F> { static constexpr int NARR = 1; static constexpr int RB = 8; };
template <>
struct ModeTraits<PL_QL> { static constexpr int NARR = 1; static constexpr int RB = 8; };
template <>
struct ModeTraits<PL_QH> { static constexpr int NARR = 1; static constexpr int RB = 8; };

constexpr int kTW = 128;        // output tile width: 64 threads x 2 pixels
constexpr int kTH = 32;         // output tile height: 4 row groups x 8 rows
constexpr int kThreads = 256;   // 8 warps
constexpr int kMaxSize = 8191;  // span table lives in shared memory (4 B per kernel row)
constexpr int kSuperCols = 16;  // two-pass raster: tiles per super-column

struct DiscParams {
    const float* dem;
    float* out;
    uint32_t* planes;  // two-pass: global row-prefix planes (2 copies per plane); fused: unused
    uint32_t* cplanes;  // hybrid: column-prefix planes (2 copies per plane), prefix_rows + 1 rows each
    unsigned long long* sat;  // hybrid: summed-area planes (64-bit), prefix_rows + 1 rows each
    int64_t cplane_stride, sat_stride;  // elements between consecutive copies / planes
    int asq;     // hybrid: half-side of the square inscribed in the disc, floor(mid / sqrt(2)); octagon: u
    // octagon decomposition (plane cache only): O = {|i| <= u, |j| <= u, |i| + |j| <= u + v}, v = isqrt(mid^2 - u^2)
    int oct;         // 1 = the diagonal tables below exist and the walk uses the octagon
    int oct_v;       // v
    int oct_ndiag;   // corner diagonals |i| + |j| = u + v + 1 .. u + v + ndiag
    uint32_t* eplanes;           // diagonal prefix of the plane values: E1 A, E1 B, E2 A, E2 B (plane_stride apart)
    unsigned long long* dplanes; // diagonal sums of the 64-bit row prefix: D1, D2 (sat_stride apart)
    unsigned long long* partial;  // two-pass, multi-plane modes: raw disc sums, [plane][out_rows][nx]
    int64_t partial_stride;
    unsigned long long* tsum;  // optional: raw sums of the T plane (trunc(z) - tmin), shared between tpi and std
    unsigned long long* fsum;  // optional, float DEMs: raw sums of the fraction plane, shared likewise
    unsigned long long* qsum;  // optional, integer DEMs on the FFT route: raw sums of the (low half of the) square plane
    int fplane;                // index of the fraction plane among this mode's planes (-1: none)
    int nrows;   // two-pass: rows of the prefix planes
    int64_t ld_in, ld_out;
    int64_t plane_stride;  // elements between consecutive plane copies
    int nx, gny, in_gy0, in_rows, out_gy0, out_rows;
    int k;       // kernel size
    int c;       // (k-1)/2 : scipy 'same' centring
    int mid;     // k/2     : circular_kernel's middle
    int square;  // size < 5
    int halo;    // k/2, rows/cols of halo on each side
    int haloL;   // left halo rounded up to a multiple of 8 (keeps 128-bit loads / 64-bit pairs aligned)
    int pitch;   // prefix row pitch in elements (multiple of 8)
    int prow0;   // two-pass: global row of prefix row 0
    int tiles_x, tiles_y;
    // conversion constants
    float scale;   // 2^S
    int c0i;       // c0 * 2^S            (TPI_Q)
    int tmin;      // offset of the integer part
    int cmid;      // offset inside the square
    float fscale;  // 2^Sf
    // epilogue constants
    double inv_scale, inv_fscale, n, inv_nm1;
    long long n_ll;
    double nc0_scaled, n_tmin;  // N*c0 (TPI_Q: times 2^-S folded in) and N*tmin as doubles, for the 32-bit epilogue
    double inv_n_nm1;  // 1 / (N * (N - 1))
    int fpack;         // tiny float std: bits of the fraction field of the packed word
    int qsplit;        // STD: the square plane is split in PL_QL + PL_QH (raw sums of PL_QH: partial plane NARR)
    int exact64;       // N*B - a^2 fits in 64-bit integers
    int excl;          // TPI: offset (excl, excl) of the excluded "mid point" (0 odd size, -1 even size)
};

// ---- span table: kernel row i -> (dxlo, dxhi) of its run of ones, as DEM column offsets ------------
// out[y,x] = sum_{i,j} K[i,j] * d[y + c - i, x + c - j]
__device__ __forceinline__ void disc_row_columns(int k, int mid, int square, int i, int& jlo, int& jhi) {
    if (square) {
        jlo = 0;
        jhi = k - 1;
    } else {
        int di = i - mid;
        int rem = mid * mid - di * di;  // >= 0 for every row of the kernel
        int w = (int)floorf(sqrtf((float)rem));
        while (w * w > rem) --w;
        while ((w + 1) * (w + 1) <= rem) ++w;
        jlo = mid - w;
        jhi = mid + w;
        if (jhi > k - 1) jhi = k - 1;  // even sizes: the disc is clipped on the right/bottom
    }
}

__device__ __forceinline__ void kernel_row_span(const DiscParams& p, int i, int& dxlo, int& dxhi) {
    int jlo, jhi;
    disc_row_columns(p.k, p.mid, p.square, i, jlo, jhi);
    dxlo = p.c - jhi;
    dxhi = p.c - jlo;
}

__device__ __forceinline__ void build_span_table(const DiscParams& p, int* tab) {
    for (int i = threadIdx.x; i < p.k; i += blockDim.x) {
        int lo, hi;
        kernel_row_span(p, i, lo, hi);
        tab[i] = (lo & 0xffff) | (hi << 16);
    }
}

// ---- value -> uint32 planes ----------------------------------------------------------------------
template <int MODE>
__device__ __forceinline__ void convert(const DiscParams& p, float z, uint32_t (&v)[ModeTraits<MODE>::NARR]) {
    if constexpr (MODE == TPI_Q) {
        v[0] = (uint32_t)(__float2int_rn(z * p.scale) - p.c0i);
    } else if constexpr (MODE == PL_Q) {
        const int d = __float2int_rz(z) - p.cmid;
        v[0] = (uint32_t)(d * d);
    } else if constexpr (MODE == PL_QL) {
        const int d = __float2int_rz(z) - p.cmid;
        v[0] = (uint32_t)(d * d) & 0xffffu;
    } else if constexpr (MODE == PL_QH) {
        const int d = __float2int_rz(z) - p.cmid;
        v[0] = (uint32_t)(d * d) >> 16;
    } else if constexpr (MODE == PL_F) {
        const float f1 = (z - (float)__float2int_rz(z)) + 1.0f;
        v[0] = (uint32_t)__float2int_rn(f1 * p.fscale);
    } else {
        const int t = __float2int_rz(z);
        v[0] = (uint32_t)(t - p.tmin);
        if constexpr (MODE == STD_I || MODE == STD_F) {
            const int d = t - p.cmid;
            v[1] = (uint32_t)(d * d);
        }
        if constexpr (MODE == TPI_X || MODE == STD_F) {
            const float f1 = (z - (float)t) + 1.0f;  // exact: frac in (-1, 1)
            v[ModeTraits<MODE>::NARR - 1] = (uint32_t)__float2int_rn(f1 * p.fscale);
        }
    }
}

// Load 4 consecutive values of global row `gy`, columns x..x+3 (x % 4 == 0), zero outside the image.
__device__ __forceinline__ float4 load4_zero(const DiscParams& p, int gy, int x, bool aligned) {
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    // outside the image: zero padding.  Outside the band: only reached by tile rows whose outputs are
    // never stored (the host checked that every needed row is present).
    if (gy < 0 || gy >= p.gny || gy < p.in_gy0 || gy >= p.in_gy0 + p.in_rows) return r;
    const float* row = p.dem + (int64_t)(gy - p.in_gy0) * p.ld_in;
    if (aligned && x >= 0 && x + 3 < p.nx) return ldg4(row + x);
    if (x >= 0 && x < p.nx) r.x = __ldg(row + x);
    if (x + 1 >= 0 && x + 1 < p.nx) r.y = __ldg(row + x + 1);
    if (x + 2 >= 0 && x + 2 < p.nx) r.z = __ldg(row + x + 2);
    if (x + 3 >= 0 && x + 3 < p.nx) r.w = __ldg(row + x + 3);
    return r;
}

// One warp scans one row: W elements starting at global column xs (multiple of 4) into the exclusive
// prefix dst[a][0..W] of each plane a (rows have `pitch` >= W + 8 elements, 32-byte aligned).  With
// COPIES == 2 the shifted copy dstB[a][k] = P[k+1] is written too.  8 elements per lane and chunk.
// r64 != nullptr: the un-wrapped 64-bit row prefix is stored as well (input of the summed-area table).
template <int MODE, int COPIES>
__device__ __forceinline__ void scan_row(const DiscParams& p, int gy, int xs, int W, uint32_t* dst,
                                         int64_t plane_stride, bool aligned, int lane,
                                         unsigned long long* r64 = nullptr, int64_t r64_stride = 0) {
    constexpr int NARR = ModeTraits<MODE>::NARR;
    unsigned long long carry[NARR];
#pragma unroll
    for (int a = 0; a < NARR; ++a) carry[a] = 0ull;
    const int nchunks = (W + 255) >> 8;
    for (int ch = 0; ch < nchunks; ++ch) {
        const int lc = (ch << 8) + (lane << 3);
        float4 z0 = make_float4(0.f, 0.f, 0.f, 0.f), z1 = z0;
        if (lc < W) z0 = load4_zero(p, gy, xs + lc, aligned);
        if (lc + 4 < W) z1 = load4_zero(p, gy, xs + lc + 4, aligned);
        const float zz[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
        uint32_t v[8][NARR];
#pragma unroll
        for (int j = 0; j < 8; ++j) convert<MODE>(p, zz[j], v[j]);
#pragma unroll
        for (int a = 0; a < NARR; ++a) {
            // elements past W must not contribute (they convert to a non-zero "zero" otherwise)
            uint32_t s[8];
            uint32_t run = 0u;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                run += (lc + j < W) ? v[j][a] : 0u;
                s[j] = run;  // inclusive within the lane (8 values < 2^32)
            }
            unsigned long long base64;
            uint32_t base;
            if (r64 != nullptr) {  // warp-uniform
                unsigned long long incl = run;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += t;
                }
                base64 = carry[a] + incl - run;
                base = (uint32_t)base64;
                carry[a] += __shfl_sync(0xffffffffu, incl, 31);
            } else {
                const uint32_t incl = warp_inclusive_scan_u32(run, lane);
                base = (uint32_t)carry[a] + incl - run;  // exclusive prefix at element lc
                base64 = base;
                carry[a] = (uint32_t)carry[a] + __shfl_sync(0xffffffffu, incl, 31);
            }
            if (lc < W) {
                uint32_t* d = dst + (int64_t)(COPIES * a) * plane_stride + lc;
                *reinterpret_cast<uint4*>(d) = make_uint4(base, base + s[0], base + s[1], base + s[2]);
                *reinterpret_cast<uint4*>(d + 4) = make_uint4(base + s[3], base + s[4], base + s[5], base + s[6]);
                if constexpr (COPIES == 2) {
                    uint32_t* e = d + plane_stride;
                    *reinterpret_cast<uint4*>(e) = make_uint4(base + s[0], base + s[1], base + s[2], base + s[3]);
                    *reinterpret_cast<uint4*>(e + 4) = make_uint4(base + s[4], base + s[5], base + s[6], base + s[7]);
                }
                if (r64 != nullptr) {
                    unsigned long long* q = r64 + a * r64_stride + lc;
                    *reinterpret_cast<ulonglong2*>(q) = make_ulonglong2(base64, base64 + s[0]);
                    *reinterpret_cast<ulonglong2*>(q + 2) = make_ulonglong2(base64 + s[1], base64 + s[2]);
                    *reinterpret_cast<ulonglong2*>(q + 4) = make_ulonglong2(base64 + s[3], base64 + s[4]);
                    *reinterpret_cast<ulonglong2*>(q + 6) = make_ulonglong2(base64 + s[5], base64 + s[6]);
                }
            }
        }
    }
    // closing element P[W8] (W8 = W rounded up to 8): total of the row (P[W] itself when W % 8 == 0)
    if (lane == 0) {
        const int W8 = (W + 7) & ~7;
#pragma unroll
        for (int a = 0; a < NARR; ++a) {
            dst[(int64_t)(COPIES * a) * plane_stride + W8] = (uint32_t)carry[a];
            if (r64 != nullptr) r64[a * r64_stride + W8] = carry[a];
        }
    }
}

// ---- epilogue -------------------------------------------------------------------------------------
template <int MODE>
__device__ __forceinline__ float finish(const DiscParams& p, const unsigned long long (&acc)[ModeTraits<MODE>::NARR],
                                        int gy, int x) {
    const double n = p.n;
    if constexpr (MODE == TPI_Q || MODE == TPI_X || MODE == TPI_I) {
        double sum_z;
        if constexpr (MODE == TPI_Q) {
            const long long tot = (long long)acc[0] + p.n_ll * (long long)p.c0i;
            sum_z = (double)tot * p.inv_scale;
        } else {
            const long long st = (long long)acc[0] + p.n_ll * (long long)p.tmin;
            sum_z = (double)st;
            if constexpr (MODE == TPI_X) sum_z += (double)acc[1] * p.inv_fscale - n;
        }
        const float* row = p.dem + (int64_t)(gy - p.in_gy0) * p.ld_in;
        const double z = (double)__ldg(row + x);
        double ze = z;
        if (p.excl != 0) {  // even size: the excluded "mid point" is the pixel at (-1, -1)
            const int ey = gy + p.excl, ex = x + p.excl;
            ze = (ey >= 0 && ex >= 0) ? (double)__ldg(p.dem + (int64_t)(ey - p.in_gy0) * p.ld_in + ex) : 0.0;
        }
        // z - conv/(N-1) as (z*(N-1) - conv)/(N-1): the numerator is exact, so a constant or planar
        // neighbourhood gives exactly 0
        return (float)(fma(z, n - 1.0, -(sum_z - ze)) * p.inv_nm1);
    } else {
        // N*(N-1)*var = N*sum(t^2) - (sum x)^2, evaluated around cmid so that the integer part
        //   N*B - a^2,  a = sum(t - cmid), B = sum((t - cmid)^2)
        // is exact in 64-bit integers (=> exactly 0 on flat terrain); the fractional parts enter as
        //   - F*(2a + F + 2*N*cmid),  F = sum frac(x)
        const long long a = (long long)acc[0] - p.n_ll * (long long)(p.cmid - p.tmin);
        double num;
        if (p.exact64)
            num = (double)(p.n_ll * (long long)acc[1] - a * a);
        else
            num = n * (double)acc[1] - (double)a * (double)a;
        if constexpr (MODE == STD_F) {
            const double F = (double)acc[2] * p.inv_fscale - n;
            num -= F * (2.0 * (double)a + F + 2.0 * n * (double)p.cmid);
        }
        const double var = num * p.inv_n_nm1;
        return (float)sqrt(fmax(var, 0.0));
    }
}

// ---- the span walk: RB output rows x 2 adjacent pixels ------------------------------------------------
// prow: prefix row of the FIRST of the RB output rows at kernel offset dy = 0; lcx: prefix column of the
// first pixel (two-pass: even, second pixel adjacent; fused: second pixel 64 columns to the right).  ACC: bit a set => plane a's whole disc sum fits 32 bits (one IADD3 per row).
template <int MODE, int ACC, bool GLOBAL>
__device__ __forceinline__ void span_walk(const DiscParams& p, const uint32_t* __restrict__ P, int64_t plane_stride,
                                          int pitch, const int* __restrict__ tab, int prow, int lcx,
                                          unsigned long long (&out)[ModeTraits<MODE>::RB][2][ModeTraits<MODE>::NARR]) {
    constexpr int NARR = ModeTraits<MODE>::NARR;
    constexpr int RB = ModeTraits<MODE>::RB;
    uint32_t s32[RB][2][NARR];
    unsigned long long s64[RB][2][NARR];
#pragma unroll
    for (int b = 0; b < RB; ++b)
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int a = 0; a < NARR; ++a) s32[b][q][a] = 0u, s64[b][q][a] = 0ull;

    const int k = p.k;
    // kernel row i touches DEM row offset dy = c - i
    if constexpr (GLOBAL) {
        // copy A of plane a at P + 2a*stride (P[j]), copy B at P + (2a+1)*stride (P[j+1]): the pair
        // (P[j], P[j+1]) is an aligned 64-bit word of A when j is even, of B (at j-1) when j is odd.
        const uint32_t* rowp[RB];
#pragma unroll
        for (int b = 0; b < RB; ++b) rowp[b] = P + (int64_t)(prow + p.c + b) * pitch + lcx;
        const int64_t down = pitch;
#pragma unroll 1
        for (int i = 0; i < k; ++i) {
            const int e = tab[i];
            const int lo = (int)(short)(e & 0xffff);
            const int hi1 = (e >> 16) + 1;
            // lcx is even: parity of the element index = parity of the offset
            const int64_t offR = (hi1 & 1) ? (plane_stride + hi1 - 1) : (int64_t)hi1;
            const int64_t offL = (lo & 1) ? (plane_stride + lo - 1) : (int64_t)lo;
#pragma unroll
            for (int b = 0; b < RB; ++b) {
#pragma unroll
                for (int a = 0; a < NARR; ++a) {
                    const uint32_t* q = rowp[b] + 2 * a * plane_stride;
                    const uint2 hv = __ldg(reinterpret_cast<const uint2*>(q + offR));
                    const uint2 lv = __ldg(reinterpret_cast<const uint2*>(q + offL));
                    if ((ACC >> a) & 1) {
                        s32[b][0][a] += hv.x - lv.x;
                        s32[b][1][a] += hv.y - lv.y;
                    } else {
                        s64[b][0][a] += (unsigned long long)(uint32_t)(hv.x - lv.x);
                        s64[b][1][a] += (unsigned long long)(uint32_t)(hv.y - lv.y);
                    }
                }
                rowp[b] -= down;
            }
        }
    } else {
        int o = (prow + p.c) * pitch + lcx;
#pragma unroll 2
        for (int i = 0; i < k; ++i, o -= pitch) {
            const int e = tab[i];
            const int lo = (int)(short)(e & 0xffff);
            const int hi1 = (e >> 16) + 1;
#pragma unroll
            for (int b = 0; b < RB; ++b) {
                const int ob = o + b * pitch;
#pragma unroll
                for (int a = 0; a < NARR; ++a) {
                    // lanes = consecutive columns (conflict-free); the thread's second pixel is 64 columns on
                    const uint32_t* q = P + a * plane_stride + ob;
                    const uint32_t h0 = q[hi1], h1 = q[hi1 + kTW / 2], l0 = q[lo], l1 = q[lo + kTW / 2];
                    if ((ACC >> a) & 1) {
                        s32[b][0][a] += h0 - l0;
                        s32[b][1][a] += h1 - l1;
                    } else {
                        s64[b][0][a] += (unsigned long long)(uint32_t)(h0 - l0);
                        s64[b][1][a] += (unsigned long long)(uint32_t)(h1 - l1);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int b = 0; b < RB; ++b)
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int a = 0; a < NARR; ++a)
                out[b][q][a] = ((ACC >> a) & 1) ? (unsigned long long)s32[b][q][a] : s64[b][q][a];
}

// Store the RB x 2 results of one batch.  gyb: global row of batch row 0; rows < gy_first are duplicates
// produced by the bottom clamp and are skipped.  XS: distance between the thread's two pixels.
template <int MODE, int XS>
__device__ __forceinline__ void store_batch(const DiscParams& p,
                                            const unsigned long long (&acc)[ModeTraits<MODE>::RB][2][ModeTraits<MODE>::NARR],
                                            int gyb, int gy_first, int y_end, int x) {
    constexpr int RB = ModeTraits<MODE>::RB;
#pragma unroll
    for (int b = 0; b < RB; ++b) {
        const int gy = gyb + b;
        if (gy < gy_first || gy >= y_end) continue;
        float* o = p.out + (int64_t)(gy - p.out_gy0) * p.ld_out + x;
        if (XS == 1) {
            if (x + 1 < p.nx) {
                const float r0 = finish<MODE>(p, acc[b][0], gy, x), r1 = finish<MODE>(p, acc[b][1], gy, x + 1);
                if ((reinterpret_cast<uintptr_t>(o) & 7) == 0)
                    *reinterpret_cast<float2*>(o) = make_float2(r0, r1);
                else
                    o[0] = r0, o[1] = r1;
            } else if (x < p.nx) {
                o[0] = finish<MODE>(p, acc[b][0], gy, x);
            }
        } else {
            if (x < p.nx) o[0] = finish<MODE>(p, acc[b][0], gy, x);
            if (x + XS < p.nx) o[XS] = finish<MODE>(p, acc[b][1], gy, x + XS);
        }
    }
}

// ---- fused kernel --------------------------------------------------------------------------------
template <int MODE, int ACC>
__global__ void __launch_bounds__(kThreads) disc_fused_kernel(const DiscParams p) {
    constexpr int NARR = ModeTraits<MODE>::NARR;
    constexpr int RB = ModeTraits<MODE>::RB;
    extern __shared__ __align__(32) unsigned char smem_raw[];
    int* tab = reinterpret_cast<int*>(smem_raw);
    const int tab_elems = (p.k + 7) & ~7;
    uint32_t* P = reinterpret_cast<uint32_t*>(smem_raw) + tab_elems;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x0 = blockIdx.x * kTW;
    const int y0 = p.out_gy0 + blockIdx.y * kTH;  // global row of the tile's first output row
    const int R = kTH + 2 * p.halo;
    const int W = p.haloL + kTW + p.halo;
    const int64_t plane_stride = (int64_t)R * p.pitch;
    const bool aligned = ((p.ld_in & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.dem) & 15) == 0);

    build_span_table(p, tab);
    for (int r = warp; r < R; r += kThreads / 32)
        scan_row<MODE, 1>(p, y0 - p.halo + r, x0 - p.haloL, W, P + (int64_t)r * p.pitch, plane_stride, aligned, lane);
    __syncthreads();

    const int tp = threadIdx.x & 63;  // the thread owns tile columns tp and tp + 64
    const int yg = threadIdx.x >> 6;  // row group 0..3
    const int x = x0 + tp;
    const int y_end = p.out_gy0 + p.out_rows;
#pragma unroll 1
    for (int bt = 0; bt < kTH / 4; bt += RB) {
        const int ty = yg * (kTH / 4) + bt;  // first tile row of this batch
        if (y0 + ty >= y_end) break;
        unsigned long long acc[RB][2][NARR];
        span_walk<MODE, ACC, false>(p, P, plane_stride, p.pitch, tab, ty + p.halo, tp + p.haloL, acc);
        store_batch<MODE, kTW / 2>(p, acc, y0 + ty, y0 + ty, y_end, x);
    }
}

// ---- tiny discs (odd sizes 5..13): direct sliding sums in registers ---------------------------------------
// No prefix sums.  A thread owns one column of a 128 x 128 tile strip and walks down the rows: for every input
// row it builds the symmetric row sums R_w = sum_{|j| <= w} q[Y][x+j] for w = 0..M by widening (2M loads,
// 2M adds) and adds R_{w_r} to the 2M+1 output rows y = Y - r that are open, held in rotating registers
// (all indices are compile-time after unrolling by 2M+1).  ~4M+2 integer ops and 2M+1 conflict-free LDS per
// pixel and plane; squares of STD are derived on the fly (one IMAD) instead of being loaded.
constexpr int kTinyTile = 128;  // tile edge; 8 warps = 4 column groups x 2 row strips of 64 rows
constexpr int kTinyStrip = 64;

template <int M>
struct TinyWidths {
    // w[r] = half-width of kernel row at vertical offset |r|, circular_kernel(2M+1): floor(sqrt(M^2 - r^2))
    __host__ __device__ static constexpr int w(int r) {
        int rem = M * M - r * r, v = 0;
        while ((v + 1) * (v + 1) <= rem) ++v;
        return v;
    }
};

template <int MODE>
struct TinyTraits;
// LP: planes kept in shared memory; NS: sums kept per output pixel; SQ: the square plane is derived from plane 0.
// PACK (float std): ONE word per cell, (t - tmin) << fpack | fraction -- a second plane would leave a single CTA per SM.
// The sums kept are T, Q and W = sum of the packed words (mod 2^32): the fraction sum is W - (T << fpack).
template <>
struct TinyTraits<TPI_I> { static constexpr int LP = 1, NS = 1; static constexpr bool SQ = false, PACK = false; };
template <>
struct TinyTraits<TPI_Q> { static constexpr int LP = 1, NS = 1; static constexpr bool SQ = false, PACK = false; };
template <>
struct TinyTraits<TPI_X> { static constexpr int LP = 2, NS = 2; static constexpr bool SQ = false, PACK = false; };
template <>
struct TinyTraits<STD_I> { static constexpr int LP = 1, NS = 2; static constexpr bool SQ = true, PACK = false; };
template <>
struct TinyTraits<STD_F> { static constexpr int LP = 1, NS = 3; static constexpr bool SQ = true, PACK = true; };

// the planes a tiny kernel LOADS (the square plane of STD is derived from plane 0)
template <int MODE>
__device__ __forceinline__ void convert_tiny(const DiscParams& p, float z, uint32_t (&v)[TinyTraits<MODE>::LP]) {
    if constexpr (MODE == TPI_Q) {
        v[0] = (uint32_t)(__float2int_rn(z * p.scale) - p.c0i);
    } else {
        const int t = __float2int_rz(z);
        v[0] = (uint32_t)(t - p.tmin);
        if constexpr (TinyTraits<MODE>::PACK)
            v[0] = (v[0] << p.fpack) | (uint32_t)__float2int_rn(((z - (float)t) + 1.0f) * p.fscale);
        else if constexpr (TinyTraits<MODE>::LP == 2)
            v[1] = (uint32_t)__float2int_rn(((z - (float)t) + 1.0f) * p.fscale);
    }
}

template <int MODE, int M>
__global__ void __launch_bounds__(256) disc_tiny_kernel(const DiscParams p) {
    constexpr int D = 2 * M + 1;
    constexpr int LP = TinyTraits<MODE>::LP;
    constexpr bool SQ = TinyTraits<MODE>::SQ, PACK = TinyTraits<MODE>::PACK;
    constexpr int NS = TinyTraits<MODE>::NS;       // sums kept per output pixel
    constexpr int NARR = ModeTraits<MODE>::NARR;   // == NS, in the order (T, Q, F | W) / (T, F) / (q)
    static_assert(NS == NARR, "plane bookkeeping");
    constexpr int TC = kTinyTile + 2 * M;          // tile columns
    constexpr int PITCH = TC + 1;
    constexpr int STEPS = ((kTinyStrip + 2 * M + D - 1) / D) * D;  // input rows walked per strip (multiple of D)
    constexpr int TR = kTinyStrip + STEPS;         // tile rows: strip 1 starts 64 rows below strip 0
    extern __shared__ __align__(32) unsigned char smem_raw[];
    uint32_t* tile = reinterpret_cast<uint32_t*>(smem_raw);  // [LP][TR][PITCH]

    const int x0 = blockIdx.x * kTinyTile;
    const int y0 = p.out_gy0 + blockIdx.y * kTinyTile;  // global row of the tile's first output row
    const int in_end = p.in_gy0 + p.in_rows;

    // ---- stage the converted tile: rows y0-M .. y0-M+TR-1, columns x0-M .. x0-M+TC-1, zero padding outside.
    // Two rows x ceil(TC/32) loads are issued before the first use (the loop bounds are compile-time).
    {
        constexpr int CI = (TC + 31) / 32;
        const int lane_ = threadIdx.x & 31;
        for (int r = (threadIdx.x >> 5) * 2; r < TR; r += 16) {
            float z[2][CI];
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const int gy = y0 - M + r + rr;
                const bool row_ok = r + rr < TR && gy >= 0 && gy < p.gny && gy >= p.in_gy0 && gy < in_end;
                const float* src = p.dem + (int64_t)(gy - p.in_gy0) * p.ld_in;
#pragma unroll
                for (int ci = 0; ci < CI; ++ci) {
                    const int c = ci * 32 + lane_, gx = x0 - M + c;
                    z[rr][ci] = (row_ok && c < TC && gx >= 0 && gx < p.nx) ? __ldg(src + gx) : 0.f;
                }
            }
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                if (r + rr >= TR) break;
#pragma unroll
                for (int ci = 0; ci < CI; ++ci) {
                    const int c = ci * 32 + lane_;
                    if (c < TC) {
                        uint32_t v[LP];
                        convert_tiny<MODE>(p, z[rr][ci], v);
#pragma unroll
                        for (int a = 0; a < LP; ++a) tile[(a * TR + r + rr) * PITCH + c] = v[a];
                    }
                }
            }
        }
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cg = warp & 3, strip = warp >> 2;
    const int tx = cg * 32 + lane;                 // tile column of this thread's pixel column
    const int x = x0 + tx;
    const int ys = y0 + strip * kTinyStrip;        // global row of the strip's first output row
    const int y_end = p.out_gy0 + p.out_rows;
    if (ys >= y_end) return;
    const uint32_t* col = tile + (strip * kTinyStrip) * PITCH + tx + M;  // input row Y=-M of the strip, centre column
    const int dq = p.cmid - p.tmin;                // (t - cmid) = (t - tmin) - dq

    uint32_t acc[D][NS];
    float zc[D];  // TPI: the pixel's own elevation, fetched when its slot opens (2M steps before it is needed)
#pragma unroll
    for (int s = 0; s < D; ++s) {
        zc[s] = 0.f;
#pragma unroll
        for (int a = 0; a < NS; ++a) acc[s][a] = 0u;
    }
    constexpr bool IS_TPI = (MODE == TPI_I || MODE == TPI_Q || MODE == TPI_X);
    const int xz = x < p.nx ? x : p.nx - 1;

#pragma unroll 1
    for (int j = 0; j < STEPS / D; ++j) {
#pragma unroll
        for (int s = 0; s < D; ++s) {
            // input row Y = D*j + s - M (relative to the strip); tile row D*j + s
            const uint32_t* row = col + (D * j + s) * PITCH;
            uint32_t R[M + 1][NS];
            // widening: R[w] = R[w-1] + q[-w] + q[+w]
#pragma unroll
            for (int w = 0; w <= M; ++w) {
                if constexpr (PACK) {
                    const int fb = p.fpack;
                    if (w == 0) {
                        const uint32_t c0 = row[0], t0 = c0 >> fb;
                        const int d0 = (int)t0 - dq;
                        R[0][0] = t0, R[0][1] = (uint32_t)(d0 * d0), R[0][2] = c0;
                    } else {
                        const uint32_t l = row[-w], r = row[w], tl = l >> fb, tr = r >> fb;
                        const int dl = (int)tl - dq, dr = (int)tr - dq;
                        R[w][0] = R[w - 1][0] + tl + tr;
                        R[w][1] = R[w - 1][1] + (uint32_t)(dl * dl) + (uint32_t)(dr * dr);
                        R[w][2] = R[w - 1][2] + l + r;
                    }
                    continue;
                }
#pragma unroll
                for (int a = 0; a < LP; ++a) {
                    const uint32_t* pl = row + a * TR * PITCH;
                    if (w == 0) {
                        const uint32_t c0 = pl[0];
                        R[0][a == 0 ? 0 : NS - 1] = c0;
                        if (SQ && a == 0) {
                            const int d0 = (int)c0 - dq;
                            R[0][1] = (uint32_t)(d0 * d0);
                        }
                    } else {
                        const uint32_t l = pl[-w], r = pl[w];
                        R[w][a == 0 ? 0 : NS - 1] = R[w - 1][a == 0 ? 0 : NS - 1] + l + r;
                        if (SQ && a == 0) {
                            const int dl = (int)l - dq, dr = (int)r - dq;
                            R[w][1] = R[w - 1][1] + (uint32_t)(dl * dl) + (uint32_t)(dr * dr);
                        }
                    }
                }
            }
            // output rows y = Y - r, r = -M..M, live in slot (s - M - r) mod D; r = -M opens the slot, r = +M closes it
#pragma unroll
            for (int r = -M; r <= M; ++r) {
                constexpr int BIG = 4 * D;
                const int slot = (s - M - r + BIG) % D;
                const int w = TinyWidths<M>::w(r < 0 ? -r : r);
#pragma unroll
                for (int a = 0; a < NS; ++a) {
                    if (r == -M)
                        acc[slot][a] = R[w][a];
                    else
                        acc[slot][a] += R[w][a];
                }
                if (IS_TPI && r == -M) {
                    const int gyo = ys + D * j + s;  // output row whose slot opens now: Y + M = D*j + s (strip-relative)
                    zc[slot] = (gyo < y_end && D * j + s < kTinyStrip) ? __ldg(p.dem + (int64_t)(gyo - p.in_gy0) * p.ld_in + xz) : 0.f;
                }
                if (r == M) {
                    const int yo = D * j + s - 2 * M;  // output row (strip-relative) that just received its last row
                    const int gy = ys + yo;
                    if (yo >= 0 && yo < kTinyStrip && gy < y_end && x < p.nx) {
                        float res;
                        if constexpr (MODE == TPI_I || MODE == TPI_Q || MODE == TPI_X) {
                            // odd size: the excluded mid point is the pixel itself
                            const double z = (double)zc[slot];
                            double sz = (double)acc[slot][0];
                            if constexpr (MODE == TPI_Q)
                                sz = fma(sz, p.inv_scale, p.nc0_scaled);
                            else
                                sz += p.n_tmin;
                            if constexpr (MODE == TPI_X) sz += (double)acc[slot][1] * p.inv_fscale - p.n;
                            res = (float)(fma(z, p.n - 1.0, z - sz) * p.inv_nm1);
                        } else {
                            unsigned long long a64[NARR];
#pragma unroll
                            for (int a = 0; a < NS; ++a) a64[a] = acc[slot][a];
                            if constexpr (PACK) a64[2] = (uint32_t)(acc[slot][2] - (acc[slot][0] << p.fpack));  // the fraction sum
                            res = finish<MODE>(p, a64, gy, x);
                        }
                        p.out[(int64_t)(gy - p.out_gy0) * p.ld_out + x] = res;
                    }
                }
            }
        }
    }
}

// ---- two-pass kernels ------------------------------------------------------------------------------
// pass 1: prefix planes (2 copies) for global rows [prow0, prow0 + nrows) and columns [-haloL, nx + halo)
template <int MODE>
__global__ void __launch_bounds__(kThreads) disc_prefix_kernel(const DiscParams p, int nrows) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    if (row >= nrows) return;
    const bool aligned = ((p.ld_in & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.dem) & 15) == 0);
    const int W = p.haloL + p.nx + p.halo;
    // hybrid: the 64-bit row prefix of plane row `row` goes to row + 1 of the summed-area buffer (row 0 = 0)
    unsigned long long* r64 = p.sat ? p.sat + (int64_t)(row + 1) * p.pitch : nullptr;
    scan_row<MODE, 2>(p, p.prow0 + row, -p.haloL, W, p.planes + (int64_t)row * p.pitch, p.plane_stride, aligned, lane,
                      r64, p.sat_stride);
}

// ---- hybrid decomposition (odd discs that run two-pass) ------------------------------------------------
// disc = inscribed square [-a, a]^2 (a = floor(mid / sqrt 2))  +  top/bottom caps (rows |r| > a: row-prefix
// spans)  +  left/right caps (columns |c| > a: column-prefix spans).  Lookups per pixel: 4 + 8 (mid - a)
// instead of 4 mid + 2  (1.7x fewer).  Needs, besides the row-prefix planes:
//   CP[Y][X] = sum_{i < Y} q[i][X]            uint32 wrap-around, 2 shifted copies, rows 0..nrows
//   S [Y][X] = sum_{i < Y} sum_{j < X} q[i][j] uint64 (exact),                     rows 0..nrows
// both built by a chunked column scan: per-chunk column totals -> scan of the totals -> apply.
constexpr int kColChunk = 64;

// element (row, 4 columns starting at col) of plane a as produced by `convert` (zero past the row width W)
template <int MODE>
__device__ __forceinline__ void load_q4(const DiscParams& p, int row, int col, int W, bool aligned,
                                        uint32_t (&q)[4][ModeTraits<MODE>::NARR]) {
    const float4 z = load4_zero(p, p.prow0 + row, col - p.haloL, aligned);
    const float zz[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        convert<MODE>(p, zz[j], q[j]);
        if (col + j >= W) {
#pragma unroll
            for (int a = 0; a < ModeTraits<MODE>::NARR; ++a) q[j][a] = 0u;
        }
    }
}

// phase 1: column totals of each chunk of kColChunk rows.  grid (pitch / 4 / 256, nchunks).
//   tot_q [a][chunk][col] (uint32)  : totals of q          -> column prefix
//   tot_r [a][chunk][col] (uint64)  : totals of the 64-bit row prefix (rows 1.. of p.sat) -> summed-area table
template <int MODE>
__global__ void __launch_bounds__(256) disc_colsum_kernel(const DiscParams p, uint32_t* __restrict__ tot_q,
                                                          unsigned long long* __restrict__ tot_r, int nchunks) {
    constexpr int NARR = ModeTraits<MODE>::NARR;
    const int col = (blockIdx.x * 256 + threadIdx.x) * 4;
    if (col >= p.pitch) return;
    const int chunk = blockIdx.y;
    const int r0 = chunk * kColChunk, r1 = min(r0 + kColChunk, p.nrows);
    const bool aligned = ((p.ld_in & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.dem) & 15) == 0);
    const int W = p.haloL + p.nx + p.halo;
    uint32_t sq[4][NARR];
    unsigned long long sr[4][NARR];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int a = 0; a < NARR; ++a) sq[j][a] = 0u, sr[j][a] = 0ull;
    for (int r = r0; r < r1; ++r) {
        uint32_t q[4][NARR];
        load_q4<MODE>(p, r, col, W, aligned, q);
#pragma unroll
        for (int a = 0; a < NARR; ++a) {
            const unsigned long long* s = p.sat + a * p.sat_stride + (int64_t)(r + 1) * p.pitch + col;
            const ulonglong2 u0 = *reinterpret_cast<const ulonglong2*>(s), u1 = *reinterpret_cast<const ulonglong2*>(s + 2);
            sr[0][a] += u0.x, sr[1][a] += u0.y, sr[2][a] += u1.x, sr[3][a] += u1.y;
#pragma unroll
            for (int j = 0; j < 4; ++j) sq[j][a] += q[j][a];
        }
    }
#pragma unroll
    for (int a = 0; a < NARR; ++a) {
        const int64_t o = ((int64_t)a * nchunks + chunk) * p.pitch + col;
        *reinterpret_cast<uint4*>(tot_q + o) = make_uint4(sq[0][a], sq[1][a], sq[2][a], sq[3][a]);
        *reinterpret_cast<ulonglong2*>(tot_r + o) = make_ulonglong2(sr[0][a], sr[1][a]);
        *reinterpret_cast<ulonglong2*>(tot_r + o + 2) = make_ulonglong2(sr[2][a], sr[3][a]);
    }
}

// phase 2: exclusive scan of the chunk totals along the chunk axis (one thread per column and plane)
__global__ void __launch_bounds__(256) disc_chunkscan_kernel(uint32_t* __restrict__ tot_q, unsigned long long* __restrict__ tot_r,
                                                             int nchunks, int pitch, int narr) {
    const int col = blockIdx.x * 256 + threadIdx.x;
    if (col >= pitch) return;
    for (int a = 0; a < narr; ++a) {
        uint32_t rq = 0u;
        unsigned long long rr = 0ull;
        for (int c = 0; c < nchunks; ++c) {
            const int64_t o = ((int64_t)a * nchunks + c) * pitch + col;
            const uint32_t vq = tot_q[o];
            const unsigned long long vr = tot_r[o];
            tot_q[o] = rq, tot_r[o] = rr;
            rq += vq, rr += vr;
        }
    }
}

// phase 3: inclusive column scan inside each chunk, offset by the scanned totals.
//   CP row Y+1 (both copies) = sum of q rows 0..Y ;  S row Y+1 = sum of 64-bit row prefixes 0..Y (in place).
template <int MODE>
__global__ void __launch_bounds__(256) disc_colapply_kernel(const DiscParams p, const uint32_t* __restrict__ tot_q,
                                                            const unsigned long long* __restrict__ tot_r, int nchunks) {
    constexpr int NARR = ModeTraits<MODE>::NARR;
    const int col = (blockIdx.x * 256 + threadIdx.x) * 4;
    if (col >= p.pitch) return;
    const int chunk = blockIdx.y;
    const int r0 = chunk * kColChunk, r1 = min(r0 + kColChunk, p.nrows);
    const bool aligned = ((p.ld_in & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.dem) & 15) == 0);
    const int W = p.haloL + p.nx + p.halo;
    uint32_t sq[4][NARR];
    unsigned long long sr[4][NARR];
#pragma unroll
    for (int a = 0; a < NARR; ++a) {
        const int64_t o = ((int64_t)a * nchunks + chunk) * p.pitch + col;
        const uint4 t = *reinterpret_cast<const uint4*>(tot_q + o);
        sq[0][a] = t.x, sq[1][a] = t.y, sq[2][a] = t.z, sq[3][a] = t.w;
        const ulonglong2 u0 = *reinterpret_cast<const ulonglong2*>(tot_r + o), u1 = *reinterpret_cast<const ulonglong2*>(tot_r + o + 2);
        sr[0][a] = u0.x, sr[1][a] = u0.y, sr[2][a] = u1.x, sr[3][a] = u1.y;
        if (chunk == 0) {  // row 0 of CP and S is the empty sum
            uint32_t* cA = p.cplanes + (int64_t)(2 * a) * p.cplane_stride + col;
            *reinterpret_cast<uint4*>(cA) = make_uint4(0u, 0u, 0u, 0u);
            *reinterpret_cast<uint4*>(cA + p.cplane_stride) = make_uint4(0u, 0u, 0u, 0u);
            unsigned long long* s = p.sat + a * p.sat_stride + col;
            *reinterpret_cast<ulonglong2*>(s) = make_ulonglong2(0ull, 0ull);
            *reinterpret_cast<ulonglong2*>(s + 2) = make_ulonglong2(0ull, 0ull);
        }
    }
    for (int r = r0; r < r1; ++r) {
        uint32_t q[4][NARR];
        load_q4<MODE>(p, r, col, W, aligned, q);
#pragma unroll
        for (int a = 0; a < NARR; ++a) {
            unsigned long long* s = p.sat + a * p.sat_stride + (int64_t)(r + 1) * p.pitch + col;
            const ulonglong2 u0 = *reinterpret_cast<const ulonglong2*>(s), u1 = *reinterpret_cast<const ulonglong2*>(s + 2);
            sr[0][a] += u0.x, sr[1][a] += u0.y, sr[2][a] += u1.x, sr[3][a] += u1.y;
            *reinterpret_cast<ulonglong2*>(s) = make_ulonglong2(sr[0][a], sr[1][a]);
            *reinterpret_cast<ulonglong2*>(s + 2) = make_ulonglong2(sr[2][a], sr[3][a]);
#pragma unroll
            for (int j = 0; j < 4; ++j) sq[j][a] += q[j][a];
            uint32_t* cA = p.cplanes + (int64_t)(2 * a) * p.cplane_stride + (int64_t)(r + 1) * p.pitch + col;
            *reinterpret_cast<uint4*>(cA) = make_uint4(sq[0][a], sq[1][a], sq[2][a], sq[3][a]);
            // shifted copy: B[X] = A[X + 1]
            uint32_t* cB = cA + p.cplane_stride;
            if (col > 0) cB[-1] = sq[0][a];
            cB[0] = sq[1][a], cB[1] = sq[2][a], cB[2] = sq[3][a];
        }
    }
}

// ---- octagon tables: running sums along the diagonals --------------------------------------------------------
//   DIR = +1 (blockIdx.z = 0): T[r][c] = v[r][c] + T[r-1][c-1]      (c - r constant)
//   DIR = -1 (blockIdx.z = 1): T[r][c] = v[r][c] + T[r-1][c+1]      (c + r constant)
// One thread per (diagonal, chunk of kDiagChunk rows); lanes = neighbouring diagonals => at every row the warp
// touches consecutive columns (coalesced).  Three phases like the column scan: chunk totals -> exclusive scan of the
// totals along each diagonal -> running sums offset by the scanned totals.  Sums that start outside the plane start
// at 0: every lookup is a difference of two entries of the same diagonal, so that cancels.
// WIDE = false: v = plane value (uint32 wrap-around, two shifted copies)  -> E1 / E2  (corner diagonals of the disc)
// WIDE = true : v = 64-bit row prefix (rows 1.. of p.sat, BEFORE the column scan turns it into the summed-area table)
//               -> D1 / D2 (sheared summed-area tables: a 45-degree trapezoid costs 4 lookups)
constexpr int kDiagChunk = 256;

template <int MODE>
__device__ __forceinline__ uint32_t plane_value_at(const DiscParams& p, int r, int c) {
    const int gy = p.prow0 + r, gx = c - p.haloL;
    float z = 0.f;  // zero padding outside the image (rows outside the band are never used, see load4_zero)
    if (gy >= 0 && gy < p.gny && gy >= p.in_gy0 && gy < p.in_gy0 + p.in_rows && gx >= 0 && gx < p.nx)
        z = __ldg(p.dem + (int64_t)(gy - p.in_gy0) * p.ld_in + gx);
    uint32_t q[ModeTraits<MODE>::NARR];
    convert<MODE>(p, z, q);
    return q[0];
}

// PHASE 0: chunk totals -> tot[dir][chunk][t];  PHASE 1: running sums, starting from the scanned totals
template <int MODE, bool WIDE, int PHASE>
__global__ void __launch_bounds__(256) disc_diag_kernel(const DiscParams p, unsigned long long* __restrict__ tot, int ndiag) {
    const int W = p.haloL + p.nx + p.halo;
    const int wc = WIDE ? W + 1 : W;  // columns of the source
    const int dir = blockIdx.z ? -1 : 1;
    const int t = blockIdx.x * 256 + threadIdx.x;
    if (t >= p.nrows + wc - 1) return;
    // column at row 0: c0 = t - (nrows - 1) for DIR +1 (then c = c0 + r), c0 = t for DIR -1 (then c = c0 - r)
    const int c0 = dir > 0 ? t - (p.nrows - 1) : t;
    const int r_begin = max(dir > 0 ? max(0, -c0) : max(0, c0 - (wc - 1)), (int)blockIdx.y * kDiagChunk);
    const int r_end = min(dir > 0 ? min(p.nrows, wc - c0) : min(p.nrows, c0 + 1), ((int)blockIdx.y + 1) * kDiagChunk);
    unsigned long long* slot = tot + ((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * ndiag + t;
    unsigned long long acc = PHASE == 1 ? *slot : 0ull;
    auto source = [&](int r) -> unsigned long long {
        const int c = c0 + dir * r;
        if constexpr (WIDE) return __ldg(p.sat + (int64_t)(r + 1) * p.pitch + c);
        else return plane_value_at<MODE>(p, r, c);
    };
    auto store = [&](int r, unsigned long long v) {
        const int c = c0 + dir * r;
        if constexpr (WIDE) {
            (p.dplanes + (blockIdx.z ? p.sat_stride : 0))[(int64_t)r * p.pitch + c] = v;
        } else {
            uint32_t* dstA = p.eplanes + (blockIdx.z ? 2 * p.plane_stride : 0) + (int64_t)r * p.pitch;
            dstA[c] = (uint32_t)v;
            if (c > 0) dstA[p.plane_stride + c - 1] = (uint32_t)v;  // shifted copy: B[c] = A[c + 1]
        }
    };
    int r = r_begin;
    for (; r + 8 <= r_end; r += 8) {
        unsigned long long v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = source(r + k);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            acc += v[k];
            if (PHASE == 1) store(r + k, acc);
        }
    }
    for (; r < r_end; ++r) {
        acc += source(r);
        if (PHASE == 1) store(r, acc);
    }
    if (PHASE == 0) *slot = acc;
}

// exclusive scan of the chunk totals along every diagonal (both directions)
__global__ void __launch_bounds__(256) disc_diag_chunkscan_kernel(unsigned long long* __restrict__ tot, int nchunks, int ndiag) {
    const int t = blockIdx.x * 256 + threadIdx.x;
    if (t >= ndiag) return;
    unsigned long long* col = tot + (int64_t)blockIdx.y * nchunks * ndiag + t;
    unsigned long long run = 0ull;
    for (int c = 0; c < nchunks; ++c) {
        const unsigned long long v = col[(int64_t)c * ndiag];
        col[(int64_t)c * ndiag] = run;
        run += v;
    }
}

// corner diagonals of the octagon: for d = u + v + 1 + n the lattice points (|i|, |j|) = (q, d - q) inside the disc
// and inside |i|, |j| <= u form the run q = lo .. hi; dtab[n] = lo | hi << 16 (hi < lo: empty)
__device__ __forceinline__ void build_diag_table(const DiscParams& p, int* dtab) {
    const int u = p.asq, s = p.asq + p.oct_v, m2 = p.mid * p.mid;
    for (int n = threadIdx.x; n < p.oct_ndiag; n += blockDim.x) {
        const int d = s + 1 + n;
        int lo = max(0, d - u), hi = min(u, d);
        while (lo <= hi && lo * lo + (d - lo) * (d - lo) > m2) ++lo;
        while (hi >= lo && hi * hi + (d - hi) * (d - hi) > m2) --hi;
        dtab[n] = (hi < lo) ? (1 | (0 << 16)) : (lo | (hi << 16));
    }
}

// ---- two-pass span kernels: ONE plane per launch, 8 rows x 2 adjacent pixels per thread -----------------------
// All lookups are aligned 64-bit pairs (P[j], P[j+1]) from the A/B copies.  Accumulators are 32-bit when the
// whole disc sum of the plane fits (ACC32), else 64-bit.
constexpr int kRB = 8;

struct PairAcc {
    uint32_t s32[kRB][2];
    unsigned long long s64[kRB][2];
};

// Stops the compiler from folding a row base back into every per-lookup address computation.
__device__ __forceinline__ const uint32_t* opaque_ptr(const uint32_t* p) {
    asm volatile("" : "+l"(p));
    return p;
}

template <bool ACC32>
__device__ __forceinline__ void pair_add(PairAcc& A, int b, const uint2 hv, const uint2 lv) {
    if (ACC32) {
        A.s32[b][0] += hv.x - lv.x;
        A.s32[b][1] += hv.y - lv.y;
    } else {
        A.s64[b][0] += (unsigned long long)(uint32_t)(hv.x - lv.x);
        A.s64[b][1] += (unsigned long long)(uint32_t)(hv.y - lv.y);
    }
}

// Row-prefix spans of kernel rows [i_begin, i_end), GROUPED BY PREFIX ROW: output row b and kernel row i
// read prefix row R0 + b - i, so for t = i - b fixed the 8 (b, i = t + b) lookups hit the SAME prefix row at
// neighbouring columns: the 16 loads of a step share 2-4 cache lines (immediate L1 reuse, independent of
// what the other warps stream through L1).  The offsets of i = t .. t+7 live in a rotating register window.
// lead / lag: the row groups of a CTA (8 rows apart) start `lead` steps early and stop `lag` steps late so that
// at every step ALL warps of the CTA read the same prefix row: its lines are fetched from L2 once per CTA
// instead of once per row group.  Steps without any valid (b, i) slot skip their loads.
template <bool ACC32>
__device__ __forceinline__ void row_walk_grouped(const DiscParams& p, const int* __restrict__ tab, int R0, int lcx,
                                                 int i_begin, int i_end, int lead, int lag, PairAcc& A) {
    const int pitch = p.pitch;
    const int pstride = (int)p.plane_stride;  // host guarantees it fits 31 bits
    // Element offsets INCLUDING the thread's column, kept unsigned: the address of a lookup is then
    // row base (64-bit, shared by the 16 loads of a step) + 4 * offset = ONE IMAD.WIDE.U32 on the FMA pipe.
    // (Signed offsets added to a per-thread pointer cost 4 ALU-pipe instructions per load and made this walk
    // ALU-bound: ncu 79% ALU pipe, profiles/r01_ncu_full_summary.csv.)
    const uint32_t ulcx = (uint32_t)lcx;
    uint32_t offR[kRB], offL[kRB];
    auto fetch = [&](int i, uint32_t& oR, uint32_t& oL) {
        if (i >= i_begin && i < i_end) {
            const int e = tab[i];
            const int lo = (int)(short)(e & 0xffff);
            const int hi1 = (e >> 16) + 1;
            // lcx is even: parity of the element index = parity of the offset; odd -> shifted copy B
            oR = ulcx + (uint32_t)((hi1 & 1) ? (pstride + hi1 - 1) : hi1);
            oL = ulcx + (uint32_t)((lo & 1) ? (pstride + lo - 1) : lo);
        } else {
            oR = ulcx, oL = ulcx;  // both loads hit the same word: contributes 0
        }
    };
    const int t0 = i_begin - (kRB - 1) - lead;
#pragma unroll
    for (int b = 0; b < kRB; ++b) fetch(t0 + b, offR[b], offL[b]);
    const int last_row = p.nrows - 1;
#pragma unroll 1
    for (int t = t0; t < i_end + lag; t += kRB) {
#pragma unroll
        for (int s = 0; s < kRB; ++s) {
            if (t + s + kRB > i_begin && t + s < i_end) {  // some slot i = t+s+b is valid (warp-uniform)
                int r = R0 - (t + s);  // slots outside the range carry zero contributions: keep them in bounds
                r = r < 0 ? 0 : (r > last_row ? last_row : r);
                const uint32_t* rowp = opaque_ptr(p.planes + (int64_t)r * pitch);
#pragma unroll
                for (int b = 0; b < kRB; ++b) {
                    const uint2 hv = __ldg(reinterpret_cast<const uint2*>(rowp + offR[b]));
                    const uint2 lv = __ldg(reinterpret_cast<const uint2*>(rowp + offL[b]));
                    pair_add<ACC32>(A, b, hv, lv);
                }
            }
#pragma unroll
            for (int b = 0; b < kRB - 1; ++b) offR[b] = offR[b + 1], offL[b] = offL[b + 1];
            fetch(t + s + kRB, offR[kRB - 1], offL[kRB - 1]);
        }
    }
}

// Hybrid: inscribed square from the 64-bit summed-area table + row caps (grouped walk) + column caps.
// Octagon variant (p.oct, plane cache only): the O(1) part is the octagon |i| <= u, |j| <= u, |i| + |j| <= u + v
// (rectangle from the summed-area table + two 45-degree trapezoids from the sheared tables D1 / D2), the caps
// shrink to |i| > u and |j| > u, and the four corner slivers are walked along their diagonals with the diagonal
// prefix tables E1 / E2: 4 (mid - u) + 4 ndiag lines instead of 4 (mid - a) -- 284 vs 472 at size 801.
// (profiles/proto/octagon.py is the CPU prototype that pins the index conventions.)
template <bool ACC32>
__device__ __forceinline__ void hybrid_walk(const DiscParams& p, const int* __restrict__ tab, const int* __restrict__ dtab,
                                            int prow, int lcx, int lead, int lag, PairAcc& A) {
    const int a_sq = p.asq, mid = p.mid, pitch = p.pitch;
    const int rv = p.oct ? p.oct_v : a_sq;  // half-height of the summed-area rectangle (its half-width is a_sq = u)
    // ---- rectangle: rows [py-rv, py+rv], columns [lcx-a, lcx+a] (+1 for the second pixel)
#pragma unroll
    for (int b = 0; b < kRB; ++b) {
        const unsigned long long* top = p.sat + (int64_t)(prow + b - rv) * pitch + lcx;      // S row py-rv
        const unsigned long long* bot = p.sat + (int64_t)(prow + b + rv + 1) * pitch + lcx;  // S row py+rv+1
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const unsigned long long v = __ldg(bot + q + a_sq + 1) - __ldg(top + q + a_sq + 1) - __ldg(bot + q - a_sq) +
                                         __ldg(top + q - a_sq);
            A.s32[b][q] = (uint32_t)v;
            A.s64[b][q] = v;
        }
    }
    if (p.oct) {
        const int u = a_sq, v = p.oct_v, s = u + v;
        const unsigned long long* D1 = p.dplanes;
        const unsigned long long* D2 = p.dplanes + p.sat_stride;
        if (u > v) {
            // top trapezoid: rows i = -u .. -v-1, columns x - (s+i) .. x + (s+i); bottom: rows v+1 .. u, half-width s - i
#pragma unroll
            for (int b = 0; b < kRB; ++b) {
                const int64_t y = prow + b;
                const unsigned long long* t1 = D1 + (y - v - 1) * pitch + lcx;  // row y + i1      (i1 = -v-1)
                const unsigned long long* t0 = D1 + (y - u - 1) * pitch + lcx;  // row y + i0 - 1  (i0 = -u)
                const unsigned long long* s1 = D2 + (y - v - 1) * pitch + lcx;
                const unsigned long long* s0 = D2 + (y - u - 1) * pitch + lcx;
                const unsigned long long* b1 = D2 + (y + u) * pitch + lcx;      // row y + j1      (j1 = u)
                const unsigned long long* b0 = D2 + (y + v) * pitch + lcx;      // row y + j0 - 1  (j0 = v+1)
                const unsigned long long* c1 = D1 + (y + u) * pitch + lcx;
                const unsigned long long* c0 = D1 + (y + v) * pitch + lcx;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const unsigned long long top = (__ldg(t1 + q + s - v) - __ldg(t0 + q + s - u)) -
                                                   (__ldg(s1 + q - s + v + 1) - __ldg(s0 + q - s + u + 1));
                    const unsigned long long bot = (__ldg(b1 + q + s - u + 1) - __ldg(b0 + q + s - v + 1)) -
                                                   (__ldg(c1 + q - s + u) - __ldg(c0 + q - s + v));
                    A.s32[b][q] += (uint32_t)(top + bot);
                    A.s64[b][q] += top + bot;
                }
            }
        }
        // corner slivers: diagonal |i| + |j| = d, run q = lo .. hi of |i| in every quadrant.  Quadrant by quadrant:
        // neighbouring diagonals end on neighbouring rows / columns, so their lookups reuse the same L1 lines.
        const int estride = (int)p.plane_stride;
#pragma unroll 1
        for (int k = 0; k < 4; ++k) {
            const uint32_t* tbl = p.eplanes + (k < 2 ? 2 * p.plane_stride : 0);  // quadrants 0, 1: E2; 2, 3: E1
#pragma unroll 1
            for (int n = 0; n < p.oct_ndiag; ++n) {
                const int e = dtab[n];
                const int lo = e & 0xffff, hi = e >> 16;
                if (hi < lo) continue;
                const int d = s + 1 + n;
                // (row, column) offsets of the far / near end of the run, see octagon.py
                int rh, rl, ch, cl;
                if (k == 0) rh = hi, rl = lo - 1, ch = d - hi, cl = d - lo + 1;
                else if (k == 1) rh = -lo, rl = -hi - 1, ch = -d + lo, cl = -d + hi + 1;
                else if (k == 2) rh = hi, rl = lo - 1, ch = -d + hi, cl = -d + lo - 1;
                else rh = -lo, rl = -hi - 1, ch = d - lo, cl = d - hi - 1;
                const uint32_t oh = (uint32_t)(lcx + ((ch & 1) ? (estride + ch - 1) : ch));
                const uint32_t ol = (uint32_t)(lcx + ((cl & 1) ? (estride + cl - 1) : cl));
                const uint32_t* ph = opaque_ptr(tbl + (int64_t)(prow + rh) * pitch);
                const uint32_t* pl = opaque_ptr(tbl + (int64_t)(prow + rl) * pitch);
#pragma unroll
                for (int b = 0; b < kRB; ++b) {
                    const uint2 hv = __ldg(reinterpret_cast<const uint2*>(ph + oh));
                    const uint2 lv = __ldg(reinterpret_cast<const uint2*>(pl + ol));
                    pair_add<ACC32>(A, b, hv, lv);
                    ph += pitch, pl += pitch;
                }
            }
        }
    }
    // ---- top / bottom caps: kernel rows |i - mid| > a
    const int ncap = mid - a_sq;
    row_walk_grouped<ACC32>(p, tab, prow + p.c, lcx, 0, ncap, lead, lag, A);
    row_walk_grouped<ACC32>(p, tab, prow + p.c, lcx, p.k - ncap, p.k, lead, lag, A);
    // ---- left / right caps: columns |cc| > a, column-prefix spans of half-height h = half-width of row mid+cc
    {
        const uint32_t* colp = p.cplanes + (int64_t)prow * pitch;  // row base; the column goes into the offset
        const int cstride = (int)p.cplane_stride;
#pragma unroll 1
        for (int cc = a_sq + 1; cc <= mid; ++cc) {
            const int h = tab[mid + cc] >> 16;  // dxhi of kernel row mid+cc = its half-width (odd size)
            const uint32_t* up = colp - (int64_t)h * pitch;
            const uint32_t* dn = colp + (int64_t)(h + 1) * pitch;
            const uint32_t* upb[kRB];
            const uint32_t* dnb[kRB];
#pragma unroll
            for (int b = 0; b < kRB; ++b) upb[b] = opaque_ptr(up + (int64_t)b * pitch), dnb[b] = opaque_ptr(dn + (int64_t)b * pitch);
#pragma unroll
            for (int sgn = 0; sgn < 2; ++sgn) {
                const int c = sgn ? cc : -cc;
                // lcx even: parity of the column = parity of c; unsigned element offset -> IMAD.WIDE.U32 addressing
                const uint32_t oc = (uint32_t)(lcx + ((c & 1) ? (cstride + c - 1) : c));
#pragma unroll
                for (int b = 0; b < kRB; ++b) {
                    const uint2 hv = __ldg(reinterpret_cast<const uint2*>(dnb[b] + oc));
                    const uint2 lv = __ldg(reinterpret_cast<const uint2*>(upb[b] + oc));
                    pair_add<ACC32>(A, b, hv, lv);
                }
            }
        }
    }
}

// FIN: TPI_Q / TPI_I = finish in place (single-plane descriptors); -1 = store the raw sums of plane `plane`.
template <bool ACC32, bool HYBRID, int FIN>
__global__ void __launch_bounds__(kThreads, HYBRID ? 3 : 1) disc_span_kernel(const DiscParams p, int plane) {
    extern __shared__ __align__(32) unsigned char smem_raw[];
    int* tab = reinterpret_cast<int*>(smem_raw);
    int* dtab = tab + ((p.k + 7) & ~7);  // corner diagonals of the octagon (hybrid walk with a plane cache)
    build_span_table(p, tab);
    if (HYBRID && p.oct) build_diag_table(p, dtab);
    __syncthreads();

    // super-column raster: tiles of kSuperCols columns are walked top to bottom before moving right
    int tile_x, tile_y;
    {
        const int id = blockIdx.x;
        const int full = p.tiles_x / kSuperCols;
        const int per_sc = kSuperCols * p.tiles_y;
        if (id < full * per_sc) {
            const int sc = id / per_sc, r = id - sc * per_sc;
            tile_y = r / kSuperCols;
            tile_x = sc * kSuperCols + (r - tile_y * kSuperCols);
        } else {
            const int r = id - full * per_sc, width = p.tiles_x - full * kSuperCols;
            tile_y = r / width;
            tile_x = full * kSuperCols + (r - tile_y * width);
        }
    }

    const int tp = threadIdx.x & 63;
    const int yg = threadIdx.x >> 6;
    const int x = tile_x * kTW + 2 * tp;
    int xc = x;  // clamp: keeps every address inside the planes (even, so pairs stay aligned)
    if (xc > p.nx - 1) xc = (p.nx - 1) & ~1;
    const int y0 = p.out_gy0 + tile_y * kTH;
    const int y_end = p.out_gy0 + p.out_rows;
    const int gy0 = y0 + yg * kRB;
    if (gy0 >= y_end) return;
    // clamp the batch so that all rows stay inside the planes (duplicates are not stored)
    int gyb = gy0;
    if (gyb + kRB > y_end) gyb = y_end - kRB;
    if (gyb < p.out_gy0) gyb = p.out_gy0;  // out_rows < 8: planes are padded (see host)

    PairAcc A;
#pragma unroll
    for (int b = 0; b < kRB; ++b) A.s32[b][0] = A.s32[b][1] = 0u, A.s64[b][0] = A.s64[b][1] = 0ull;
    const int prow = gyb - p.prow0, lcx = xc + p.haloL;
    // lead / lag would put the CTA's row groups in lockstep on the same prefix row (fewer L2 fetches); measured
    // slower on B200 (the walk is bound by L1 wavefronts, not by L2 traffic, and lockstep adds 24 steps): off.
    const int lead = 0, lag = 0;
    if constexpr (HYBRID)
        hybrid_walk<ACC32>(p, tab, dtab, prow, lcx, lead, lag, A);
    else
        row_walk_grouped<ACC32>(p, tab, prow + p.c, lcx, 0, p.k, lead, lag, A);

    if (x != xc) return;
#pragma unroll
    for (int b = 0; b < kRB; ++b) {
        const int gy = gyb + b;
        if (gy < gy0 || gy >= y_end) continue;
        const unsigned long long v0 = ACC32 ? (unsigned long long)A.s32[b][0] : A.s64[b][0];
        const unsigned long long v1 = ACC32 ? (unsigned long long)A.s32[b][1] : A.s64[b][1];
        if constexpr (FIN >= 0) {
            float* o = p.out + (int64_t)(gy - p.out_gy0) * p.ld_out + x;
            const unsigned long long a0[1] = {v0}, a1[1] = {v1};
            if (x + 1 < p.nx) {
                const float r0 = finish<FIN>(p, a0, gy, x), r1 = finish<FIN>(p, a1, gy, x + 1);
                if ((reinterpret_cast<uintptr_t>(o) & 7) == 0)
                    *reinterpret_cast<float2*>(o) = make_float2(r0, r1);
                else
                    o[0] = r0, o[1] = r1;
            } else {
                o[0] = finish<FIN>(p, a0, gy, x);
            }
            if (p.tsum) {  // keep the raw T-plane sums for a following std() of the same size
                unsigned long long* t = p.tsum + (int64_t)(gy - p.out_gy0) * p.nx + x;
                t[0] = v0;
                if (x + 1 < p.nx) t[1] = v1;
            }
        } else {
            const int64_t idx = (int64_t)(gy - p.out_gy0) * p.nx + x;
            unsigned long long* o = (plane == 0 && p.tsum)          ? p.tsum + idx
                                    : (plane == p.fplane && p.fsum) ? p.fsum + idx
                                                                    : p.partial + plane * p.partial_stride + idx;
            o[0] = v0;
            if (x + 1 < p.nx) o[1] = v1;
        }
    }
}

// Epilogue of the multi-plane modes on the two-pass path: reads the raw plane sums.
template <int MODE>
__global__ void __launch_bounds__(256) disc_finish_kernel(const DiscParams p) {
    constexpr int NARR = ModeTraits<MODE>::NARR;
    const int64_t total = (int64_t)p.out_rows * p.nx;
    for (int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * 256) {
        const int r = (int)(idx / p.nx), x = (int)(idx - (int64_t)r * p.nx);
        unsigned long long acc[NARR];
#pragma unroll
        for (int a = 0; a < NARR; ++a)
            acc[a] = (a == 0 && p.tsum)            ? p.tsum[idx]
                     : (a == p.fplane && p.fsum)   ? p.fsum[idx]
                     : (a == 1 && p.qsum)          ? p.qsum[idx]
                                                   : p.partial[a * p.partial_stride + idx];
        if constexpr (MODE == STD_I || MODE == STD_F) {
            if (p.qsplit) acc[1] += p.partial[NARR * p.partial_stride + idx] << 16;
        }
        p.out[(int64_t)r * p.ld_out + x] = finish<MODE>(p, acc, p.out_gy0 + r, x);
    }
}

// ---- FFT route: exact disc sums of the integer planes by float64 2-D overlap-save convolution -------------------
// The plane values are non-negative integers, so every disc sum is an integer; a float64 FFT convolution of a T x T
// window reproduces it with an absolute error far below 1/2 (bound checked by the planner: ~1e-14 * T * max value *
// sqrt(N)), and rounding to the nearest integer gives EXACTLY the sums the prefix-plane walk accumulates -- the same
// float64 epilogue then yields bit-identical TPI / STD.  Cost: one inverse 2-D transform per size and plane PAIR (two
// planes ride in the real and imaginary parts), independent of the size: 8.3 ms per pair at 16384^2 against 29-34 ms
// per plane for the size-801 walk.  The forward transforms of the planes are shared by all sizes of a sweep (cache).
struct DfftGeom {
    int T, H, V, tiles_y, tiles_x;  // transform length along x, window halo, outputs per tile along x
    int Ty, Vy;                     // transform length / outputs per tile along y (a thin band takes a shorter transform)
};

__device__ __forceinline__ double plane_value_rt(const DiscParams& p, int mode, float z) {
    switch (mode) {
        case PL_T: return (double)(uint32_t)(__float2int_rz(z) - p.tmin);
        case PL_Q: {
            const int d = __float2int_rz(z) - p.cmid;
            return (double)(uint32_t)(d * d);
        }
        case PL_QL: {
            const int d = __float2int_rz(z) - p.cmid;
            return (double)((uint32_t)(d * d) & 0xffffu);
        }
        case PL_QH: {
            const int d = __float2int_rz(z) - p.cmid;
            return (double)((uint32_t)(d * d) >> 16);
        }
        case PL_F: {
            const float f1 = (z - (float)__float2int_rz(z)) + 1.0f;
            return (double)(uint32_t)__float2int_rn(f1 * p.fscale);
        }
        default: return 0.0;
    }
}

// first forward pass of a plane pair: window row `line` of tile `plane`; zero elevation outside the image (the
// reference's zero padding: it converts to the planes' own "zero").
// twin (a plane that travels alone, mode_b < 0): the imaginary part carries the SAME plane of ANOTHER tile -- complex
// plane c holds tile 2c (real) and tile 2c + 1 (imaginary).  The disc mask is real, so the inverse transform of the
// product returns the two tiles' disc sums in its real and imaginary parts: a lone plane costs half the transforms
// (forward and inverse) and half the spectrum.
template <int N>
__global__ void __launch_bounds__(FftShape<N>::NT, N >= 8192 ? 1 : 3)
    dfft_fwd_planes_kernel(const DiscParams p, const DfftGeom g, int mode_a, int mode_b, int twin, double2* __restrict__ dst,
                           const double2* __restrict__ tw) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* buf = reinterpret_cast<double2*>(smem_raw);
    const int line = blockIdx.x, plane = blockIdx.y;
    const int ntiles = g.tiles_y * g.tiles_x;
    const int tile_a = twin ? 2 * plane : plane, tile_b = 2 * plane + 1;
    const float *row_a, *row_b = nullptr;
    bool ok_a, ok_b = false;
    int c0_a, c0_b = 0;
    auto locate = [&](int tile, const float*& row, bool& ok, int& c0) {
        const int ty = tile / g.tiles_x, tx = tile - ty * g.tiles_x;
        const int gy = p.out_gy0 + ty * g.Vy - g.H + line;
        c0 = tx * g.V - g.H;
        ok = gy >= 0 && gy < p.gny && gy >= p.in_gy0 && gy < p.in_gy0 + p.in_rows;
        row = p.dem + (int64_t)(ok ? gy - p.in_gy0 : 0) * p.ld_in;
    };
    locate(tile_a, row_a, ok_a, c0_a);
    const bool have_b = twin && tile_b < ntiles;
    if (have_b) locate(tile_b, row_b, ok_b, c0_b);
    fft2d_forward_line<N>(buf, tw, threadIdx.x, [&](int n) -> double2 {
        const int gx = c0_a + n;
        const float z = (ok_a && gx >= 0 && gx < p.nx) ? __ldg(row_a + gx) : 0.f;
        double im = 0.0;
        if (have_b) {
            const int gxb = c0_b + n;
            im = plane_value_rt(p, mode_a, (ok_b && gxb >= 0 && gxb < p.nx) ? __ldg(row_b + gxb) : 0.f);
        } else if (mode_b >= 0) {
            im = plane_value_rt(p, mode_b, z);
        }
        return make_double2(plane_value_rt(p, mode_a, z), im);
    }, dst + ((int64_t)plane * gridDim.x + line) * N);  // (grid: Ty lines x planes)
}

// first forward pass of the disc mask (circular_kernel / the square of sizes < 5), placed so that scipy's "same" crop
// of the true convolution lands on window offset (H, H)
template <int N>
__global__ void __launch_bounds__(FftShape<N>::NT, N >= 8192 ? 1 : 3)
    dfft_fwd_disc_kernel(const DiscParams p, const DfftGeom g, double2* __restrict__ dst, const double2* __restrict__ tw) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* buf = reinterpret_cast<double2*>(smem_raw);
    const int line = blockIdx.x;
    const int i = line - (g.H - p.c);  // kernel row
    double2* out = dst + (int64_t)line * N;
    if (i < 0 || i >= p.k) {
        for (int n = threadIdx.x; n < N; n += FftShape<N>::NT) out[n] = make_double2(0.0, 0.0);
        return;
    }
    int jlo, jhi;
    disc_row_columns(p.k, p.mid, p.square, i, jlo, jhi);
    fft2d_forward_line<N>(buf, tw, threadIdx.x, [&](int n) -> double2 {
        const int j = n - (g.H - p.c);
        return make_double2((j >= jlo && j <= jhi) ? 1.0 : 0.0, 0.0);
    }, out);
}

// second inverse pass + rounding to the exact integer sums: window pixel (line, n) -> output pixel.
// FIN < 0: store the raw sums of the pair.  FIN = descriptor mode: this is the LAST pair of the descriptor -- gather the
// other planes' sums (left by an earlier pair of this call or by the tpi of a tpi + std pair) and finish in place, which
// saves the 16 B/px round trip of the raw sums and the separate finish pass.
// TWIN: the real and imaginary parts are the sums of plane mode_a on tiles 2c and 2c + 1 (see dfft_fwd_planes_kernel).
template <int N, int FIN, bool TWIN>
__global__ void __launch_bounds__(FftShape<N>::NT, N >= 8192 ? 1 : 3)
    dfft_store_kernel(const DiscParams p, const DfftGeom g, const double2* __restrict__ src, const double2* __restrict__ tw,
                      unsigned long long* __restrict__ dest_a, unsigned long long* __restrict__ dest_b, int mode_a, int mode_b,
                      double scale) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* buf = reinterpret_cast<double2*>(smem_raw);
    const int line = blockIdx.x, plane = blockIdx.y;
    const int x_lo = 2 * g.H;
    // output row / first column / column bound of a tile for this window line (row < 0: nothing to produce)
    auto locate = [&](int tile, int& gy, int& ox0, int& x_hi) {
        const int ty = tile / g.tiles_x, tx = tile - ty * g.tiles_x;
        const int oy0 = p.out_gy0 + ty * g.Vy;
        const int oy1 = min(oy0 + g.Vy, p.out_gy0 + p.out_rows);
        gy = oy0 + line - 2 * g.H;
        if (tile >= g.tiles_y * g.tiles_x || gy < oy0 || gy >= oy1) gy = -1;
        ox0 = tx * g.V;
        x_hi = x_lo + min(g.V, p.nx - ox0);
    };
    int gy_a, ox_a, xh_a, gy_b = -1, ox_b = 0, xh_b = 0;
    locate(TWIN ? 2 * plane : plane, gy_a, ox_a, xh_a);
    if constexpr (TWIN) locate(2 * plane + 1, gy_b, ox_b, xh_b);
    if (gy_a < 0 && gy_b < 0) return;  // CTA-uniform
    const double2* __restrict__ in = src + ((int64_t)plane * gridDim.x + line) * N;  // (grid: Ty lines x planes)
    if constexpr (FIN >= 0) {
        // The epilogue reads, per pixel, the sums another pass left in HBM (and the DEM for TPI) right after the last
        // butterfly stage, one dependent load per output: fetch those rows into L2 now, while the transform runs.
        auto prefetch = [&](const void* base, int64_t first, int count, int elem) {
            const char* b = reinterpret_cast<const char*>(base) + first * elem;
            for (int64_t o = (int64_t)threadIdx.x * 128; o < (int64_t)count * elem; o += (int64_t)FftShape<N>::NT * 128)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(b + o));
        };
        auto prefetch_row = [&](int gy, int ox0, int x_hi) {
            if (gy < 0) return;
            const int count = x_hi - x_lo;
            const int64_t first = (int64_t)(gy - p.out_gy0) * p.nx + ox0;
            constexpr int NARR = ModeTraits<FIN>::NARR;
#pragma unroll
            for (int a = 0; a < NARR; ++a) {
                const int pm = a == 0 ? PL_T : (a == p.fplane ? PL_F : (p.qsplit ? PL_QL : PL_Q));
                if (pm == mode_a || pm == mode_b) continue;
                const unsigned long long* srcp = (a == 0 && p.tsum)          ? p.tsum
                                                 : (a == p.fplane && p.fsum) ? p.fsum
                                                 : (a == 1 && p.qsum)        ? p.qsum
                                                                             : p.partial + a * p.partial_stride;
                prefetch(srcp, first, count, 8);
            }
            if constexpr (FIN == TPI_I || FIN == TPI_X || FIN == TPI_Q)
                prefetch(p.dem, (int64_t)(gy - p.in_gy0) * p.ld_in + ox0, count, 4);
        };
        prefetch_row(gy_a, ox_a, xh_a);
        if constexpr (TWIN) prefetch_row(gy_b, ox_b, xh_b);
    }
    // one output pixel: va / vb = this transform's sums of planes ma / mb at the pixel (mb < 0: none)
    auto emit = [&](int gy, int x, unsigned long long va, unsigned long long vb, int ma, int mb) {
        const int64_t idx = (int64_t)(gy - p.out_gy0) * p.nx + x;
        if (dest_a) dest_a[idx] = va;
        if (mb >= 0 && dest_b) dest_b[idx] = vb;
        if constexpr (FIN >= 0) {
            constexpr int NARR = ModeTraits<FIN>::NARR;
            unsigned long long acc[NARR];
#pragma unroll
            for (int a = 0; a < NARR; ++a) {
                const int pm = a == 0 ? PL_T : (a == p.fplane ? PL_F : (p.qsplit ? PL_QL : PL_Q));  // the plane of slot a
                if (pm == ma)
                    acc[a] = va;
                else if (pm == mb)
                    acc[a] = vb;
                else
                    acc[a] = (a == 0 && p.tsum)          ? p.tsum[idx]
                             : (a == p.fplane && p.fsum) ? p.fsum[idx]
                             : (a == 1 && p.qsum)        ? p.qsum[idx]
                                                         : p.partial[a * p.partial_stride + idx];
            }
            if constexpr (FIN == STD_I || FIN == STD_F) {
                if (p.qsplit)  // the high half of a split square plane: this pair's, or left by an earlier one
                    acc[1] += (ma == PL_QH ? va : mb == PL_QH ? vb : p.partial[NARR * p.partial_stride + idx]) << 16;
            }
            p.out[(int64_t)(gy - p.out_gy0) * p.ld_out + x] = finish<FIN>(p, acc, gy, x);
        }
    };
    fft2d_inverse_line_from<N>(buf, tw, threadIdx.x, in, [&](int n, double2 y) {
        if (n < x_lo) return;
        const unsigned long long va = (unsigned long long)llrint(y.x * scale), vb = (unsigned long long)llrint(y.y * scale);
        if constexpr (TWIN) {
            if (gy_a >= 0 && n < xh_a) emit(gy_a, ox_a - x_lo + n, va, 0ull, mode_a, -1);
            if (gy_b >= 0 && n < xh_b) emit(gy_b, ox_b - x_lo + n, vb, 0ull, mode_a, -1);
        } else {
            if (n < xh_a) emit(gy_a, ox_a - x_lo + n, va, vb, mode_a, mode_b);
        }
    });
}

// ---- host side ---------------------------------------------------------------------------------------
struct DiscPlan {
    DiscParams p;
    int mode, acc;
    bool acc_qh;  // split squares: the PL_QH disc sums fit 32 bits
    int qh_kind;  // plane-cache region of PL_QH
    bool fused, hybrid, tiny, cached;
    size_t smem;
    int prefix_rows;  // two-pass
    int nchunks;      // hybrid: column-scan chunks
    // workspace layout (byte offsets)
    size_t off_cp, off_sat, off_totq, off_totr, off_e, off_d, off_dtot, off_partial;
    int ndchunks, ndiag_threads;  // octagon tables: row chunks and diagonals of the diagonal scans
    size_t ws_bytes;
    // FFT route
    bool fft;
    DfftGeom fg;
    int fft_pairs;                  // plane pairs of this DEM class (1: integer-valued, 2: float or split squares)
    int fft_mb0;                    // the plane that rides with T in pair 0 (the same for every call on this DEM: the
                                    // cached spectrum of pair 0 is shared by tpi and std): F, Q, QL or none
    size_t fft_plane_bytes;         // one spectrum: tiles x T x T x 16
    size_t off_tw, off_dhat, off_x, off_y, off_k1, off_k2;
};

// Sizes from here take the FFT route (one inverse 2-D transform per plane pair, ~8.3 ms per pair at 16384^2 whatever the
// size; a single call also pays the forward transforms of its planes).  Measured crossovers on B200: the cached octagon
// walk costs 4.4 / 5.6 / 8.5 ms per PLANE at sizes 41 / 81 / 161 plus ~9 ms of table builds per plane kind and sweep.
constexpr int kDiscFftMin = 128;        // single calls
constexpr int kDiscFftMinCached = 33;   // inside a sweep whose plane spectra are cached

static int dfft_length(int halo) {
    if (halo <= 128) return 2048;
    if (halo <= 640) return 4096;
    if (halo <= 2048) return 8192;
    return 0;
}

// |error| of a float64 FFT convolution output, with a safety factor of 4: must stay well below 1/2 for the rounding
// to recover the exact integer sum
static bool dfft_exact(int T, double vmax, double n) { return 4.0 * 24.0 * 1.11e-16 * (double)T * vmax * sqrt(n) < 0.25; }

static long long disc_count(int k) {
    // N = number of ones of circular_kernel(k)
    if (k < 5) return (long long)k * k;
    const int mid = k / 2;
    long long n = 0;
    for (int i = 0; i < k; ++i) {
        const long long di = i - mid, rem = (long long)mid * mid - di * di;
        long long w = (long long)floor(sqrt((double)rem));
        while (w * w > rem) --w;
        while ((w + 1) * (w + 1) <= rem) ++w;
        long long jlo = mid - w, jhi = mid + w;
        if (jhi > k - 1) jhi = k - 1;
        n += jhi - jlo + 1;
    }
    return n;
}

constexpr int kCachedTwoPassMin = 41;  // measured on B200: tpi+std 13.4 -> 10.5 ms at size 41, 22.0 -> 12.8 ms at size 81
constexpr size_t kFusedSmemBudget = 108 * 1024;  // keep 2 CTAs per SM (2 x (108 + 1) KB <= 228 KB); larger discs go two-pass.
                                                 // (108, not 101: the three planes of a float std at size 21 need 104.9 KB --
                                                 // 22.7 ms of prefix planes + walks otherwise)

static int max_rb(int mode) { return (mode == TPI_Q || mode == TPI_I) ? 8 : 4; }
static int narr_of(int mode) { return (mode == TPI_Q || mode == TPI_I) ? 1 : (mode == STD_F ? 3 : 2); }

// Geometry that does not depend on the data range (used by the workspace query too).
static int plan_geometry(const topo_view* v, int size, int narr, int rb, DiscPlan& pl, int plane_halo = 0,
                         bool force_two_pass = false) {
    DiscParams& p = pl.p;
    p.nx = v->nx, p.gny = v->gny, p.in_gy0 = v->in_gy0, p.in_rows = v->in_rows;
    p.out_gy0 = v->out_gy0, p.out_rows = v->out_rows;
    p.k = size, p.c = (size - 1) / 2, p.mid = size / 2, p.square = size < 5;
    p.halo = size / 2;
    p.haloL = (p.halo + 7) & ~7;
    p.excl = p.c - p.mid;
    p.tiles_x = ceil_div(p.nx, kTW);
    p.tiles_y = ceil_div(p.out_rows, kTH);
    const int tab_bytes = ((size + 7) & ~7) * 4;
    {
        const int R = kTH + 2 * p.halo;
        const int W = p.haloL + kTW + p.halo;
        const int pitch = ((W + 7) & ~7) + 8;
        const size_t bytes = (size_t)tab_bytes + (size_t)R * pitch * 4 * narr;
        // with a plane cache the prefix planes are already there: odd discs from kCachedTwoPassMin on walk them
        // (octagon: ~1.2 * size lookups) instead of re-scanning tile + halo in shared memory (2 * size lookups)
        const bool prefer_planes = plane_halo > 0 && (size & 1) && size >= kCachedTwoPassMin;
        if (bytes <= kFusedSmemBudget && !prefer_planes && !force_two_pass) {
            pl.fused = true;
            pl.hybrid = false;
            p.pitch = pitch;
            pl.smem = bytes;
            pl.ws_bytes = 0;
            pl.prefix_rows = 0;
            return 0;
        }
    }
    // two-pass: one plane per launch, the workspace is reused by the planes.  With a plane cache the planes are
    // laid out for the halo of the largest disc that will use them (the kernels only need halo >= size / 2).
    pl.fused = false;
    if (plane_halo > p.halo) {
        p.halo = plane_halo;
        p.haloL = (p.halo + 7) & ~7;
    }
    const int W = p.haloL + p.nx + p.halo;
    p.pitch = ((W + 7) & ~7) + 8;
    p.prow0 = p.out_gy0 - p.halo;
    int rows = p.out_rows + 2 * p.halo;
    if (p.out_rows < kRB) rows += kRB - p.out_rows;  // batch clamp may read up to 8 rows from out_gy0
    pl.prefix_rows = rows;
    p.nrows = rows;
    p.plane_stride = (int64_t)rows * p.pitch;
    auto align = [](size_t b) { return (b + 255) & ~(size_t)255; };
    size_t bytes = align((size_t)p.plane_stride * 4 * 2);
    pl.smem = tab_bytes;
    // hybrid decomposition for odd discs: square + row caps + column caps
    pl.hybrid = (size & 1) && !p.square;
    if (pl.hybrid) {
        p.asq = (int)floor((double)p.mid / sqrt(2.0));
        while (2ll * p.asq * p.asq > (long long)p.mid * p.mid) --p.asq;
        while (2ll * (p.asq + 1) * (p.asq + 1) <= (long long)p.mid * p.mid) ++p.asq;
    }
    // a cached plane region always has the full (hybrid) layout, whatever the size that happens to use it
    if (pl.hybrid || plane_halo > 0) {
        pl.nchunks = ceil_div(rows, kColChunk);
        p.cplane_stride = (int64_t)(rows + 1) * p.pitch;
        p.sat_stride = (int64_t)(rows + 1) * p.pitch;
        pl.off_cp = bytes;
        bytes = align(pl.off_cp + (size_t)p.cplane_stride * 4 * 2);
        pl.off_sat = bytes;
        bytes = align(pl.off_sat + (size_t)p.sat_stride * 8);
        pl.off_totq = bytes;
        bytes = align(pl.off_totq + (size_t)pl.nchunks * p.pitch * 4);
        pl.off_totr = bytes;
        bytes = align(pl.off_totr + (size_t)pl.nchunks * p.pitch * 8);
    }
    // octagon tables (plane cache only: they are worth building when several sizes share them)
    p.oct = 0;
    if (plane_halo > 0) {
        pl.off_e = bytes;
        bytes = align(pl.off_e + (size_t)p.plane_stride * 4 * 4);  // E1 A/B, E2 A/B
        pl.off_d = bytes;
        bytes = align(pl.off_d + (size_t)p.sat_stride * 8 * 2);    // D1, D2
        pl.ndchunks = ceil_div(rows, kDiagChunk);
        pl.ndiag_threads = rows + p.pitch;
        pl.off_dtot = bytes;
        bytes = align(pl.off_dtot + (size_t)2 * pl.ndchunks * pl.ndiag_threads * 8);  // chunk totals of the diagonal scans
        if (pl.hybrid && option_enabled(kOptOctagon)) {
            // u minimising the walked lines 4 (mid - u) + 4 (dmax - u - v), v = isqrt(mid^2 - u^2) <= u
            const long long m = p.mid, m2 = m * m;
            auto isqrt = [](long long n) {
                long long r = (long long)floor(sqrt((double)n));
                while (r * r > n) --r;
                while ((r + 1) * (r + 1) <= n) ++r;
                return r;
            };
            long long dmax = 0;
            for (long long q = 0; q <= m; ++q) dmax = std::max(dmax, q + isqrt(m2 - q * q));
            long long best = -1, bu = 0, bv = 0;
            for (long long u = (long long)floor((double)m / sqrt(2.0)); u <= m; ++u) {
                const long long v = isqrt(m2 - u * u);
                if (v > u) continue;
                const long long lines = 4 * (m - u) + 4 * (dmax - (u + v));
                if (best < 0 || lines < best) best = lines, bu = u, bv = v;
            }
            p.oct = 1;
            p.asq = (int)bu;
            p.oct_v = (int)bv;
            p.oct_ndiag = (int)(dmax - (bu + bv));
            pl.smem += (size_t)((p.oct_ndiag + 7) & ~7) * 4;
        }
    }
    pl.off_partial = bytes;
    if (narr > 1) {
        p.partial_stride = (int64_t)p.out_rows * p.nx;
        bytes += (size_t)narr * p.partial_stride * 8;
    }
    pl.ws_bytes = bytes;
    (void)rb;
    return 0;
}

static int check_band(const topo_view* v, int halo) {
    const int need_lo = v->out_gy0 - halo > 0 ? v->out_gy0 - halo : 0;
    const int hi = v->out_gy0 + v->out_rows + halo;
    const int need_hi = hi < v->gny ? hi : v->gny;
    TOPO_CHECK(v->in_gy0 <= need_lo && v->in_gy0 + v->in_rows >= need_hi,
               "input band [%d,%d) does not cover the halo rows [%d,%d)", v->in_gy0, v->in_gy0 + v->in_rows,
               need_lo, need_hi);
    return 0;
}

static int ilog2_floor(double x) {
    int e;
    frexp(x, &e);  // x = m * 2^e, m in [0.5, 1)
    return e - 1;
}

constexpr double kU32 = 4294967295.0;

// FFT route for this size, with windows laid out for `halo` rows / columns around a tile (the size's own half, or the
// half of the largest size of a sweep that shares the plane spectra)
static bool fft_route(int size, int halo, bool cached = false) {
    return option_enabled(kOptDiscFft) && size >= (cached ? kDiscFftMinCached : kDiscFftMin) && dfft_length(halo) > 0;
}

static void plan_fft_geometry(const topo_view* v, int size, int narr, int plane_halo, int pairs, DiscPlan& pl) {
    DiscParams& p = pl.p;
    p.nx = v->nx, p.gny = v->gny, p.in_gy0 = v->in_gy0, p.in_rows = v->in_rows;
    p.out_gy0 = v->out_gy0, p.out_rows = v->out_rows;
    p.k = size, p.c = (size - 1) / 2, p.mid = size / 2, p.square = size < 5;
    p.halo = size / 2;
    p.excl = p.c - p.mid;
    pl.fft = true, pl.fused = false, pl.hybrid = false, pl.tiny = false, pl.cached = plane_halo > 0;
    pl.smem = 0, pl.prefix_rows = 0;
    DfftGeom& g = pl.fg;
    g.H = plane_halo > 0 ? plane_halo : size / 2;
    g.T = dfft_length(g.H);
    g.V = g.T - 2 * g.H;
    // along y a 4096-row window may be more than a thin row band needs (2048 rows + 2 x 400 of halo on 8 GPUs): a
    // 3072-point transform (radix-6 leading stage) is taken when it covers the rows with fewer window rows in total
    g.Ty = g.T;
    if (g.T == 4096 && 3072 - 2 * g.H > 0 &&
        (long long)ceil_div(v->out_rows, 3072 - 2 * g.H) * 3072 < (long long)ceil_div(v->out_rows, g.V) * 4096)
        g.Ty = 3072;
    g.Vy = g.Ty - 2 * g.H;
    g.tiles_y = ceil_div(v->out_rows, g.Vy), g.tiles_x = ceil_div(v->nx, g.V);
    pl.fft_pairs = pairs;
    pl.fft_plane_bytes = (size_t)g.tiles_y * g.tiles_x * g.T * g.Ty * sizeof(double2);
    auto align = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t one = (size_t)g.T * g.Ty * sizeof(double2);
    size_t off = 0;
    pl.off_tw = off, off = align(off + (size_t)(g.T + g.Ty) * sizeof(double2));  // twiddles of both lengths
    pl.off_dhat = off;
    if (!pl.cached) off = align(off + pairs * pl.fft_plane_bytes);  // with a cache the plane spectra live there
    pl.off_x = off, off = align(off + pl.fft_plane_bytes);
    pl.off_y = off, off = align(off + pl.fft_plane_bytes);
    pl.off_k1 = off, off = align(off + one);
    pl.off_k2 = off, off = align(off + one);
    pl.off_partial = off;
    p.partial_stride = (int64_t)p.out_rows * p.nx;
    off += (size_t)narr * p.partial_stride * 8;
    pl.ws_bytes = off;
}

// Fill the data-dependent constants.  what: 0 = TPI, 1 = STD.
// span_size: the disc size the fixed-point scales are laid out for -- `size` itself, or the largest size of a sweep
// that shares its planes through a topo_disc_cache (then every size of the sweep sees the same planes).
static int plan_disc_impl(const topo_view* v, int size, int what, int all_integer, double zmin, double zmax, DiscPlan& pl,
                          int cache_size, int span_size, bool pair) {
    TOPO_CHECK(size >= 2 && size <= kMaxSize, "kernel size %d outside [2, %d]", size, kMaxSize);
    TOPO_CHECK(isfinite(zmin) && isfinite(zmax) && zmin <= zmax, "DEM range is not finite");
    const double n = (double)disc_count(size);
    const double span = (double)span_size;  // longest run of a kernel row
    // integer-part range, including the zero padding value
    const double tlo = fmin(0.0, trunc(zmin)), thi = fmax(0.0, trunc(zmax));
    const double trange = thi - tlo;
    int mode, acc = 0;
    DiscParams& p = pl.p;
    memset(&p, 0, sizeof(p));
    // shared planes / spectra: laid out for the halo of the largest size of the sweep
    const int plane_halo = cache_size >= size ? cache_size / 2 : 0;
    const bool fft = fft_route(size, plane_halo > 0 ? plane_halo : size / 2, plane_halo > 0);
    pl.fft = false;
    double vmax[3] = {0, 0, 0};  // largest value a plane element can take
    double vmax_qh = 0;
    int qsplit = 0;
    if (what == 0 && !all_integer) {
        const double c0 = fmin(0.0, floor(zmin));
        const double range = fmax(0.0, zmax) - c0 + 1.0;
        int S = ilog2_floor(kU32 / (span * range));
        if (S > 20) S = 20;
        // a slightly coarser scale lets the whole disc sum live in 32 bits (one IADD3 per row)
        int S32 = ilog2_floor(kU32 / (n * range));
        if (S32 > 20) S32 = 20;
        if (S32 >= 13 && span_size == size) S = S32;  // (depends on the size: not for shared planes)
        // pair: tpi shares the T and fraction sums with a std of the same size (exact TPI_X); the FFT route carries two
        // planes per transform anyway
        if (S >= 10 && !pair && !fft) {
            mode = TPI_Q;
            p.scale = (float)ldexp(1.0, S);
            p.c0i = (int)ldexp(c0, S);
            p.inv_scale = ldexp(1.0, -S);
            vmax[0] = ldexp(range, S);
        } else {
            mode = TPI_X;
        }
    } else if (what == 0) {
        mode = TPI_I;
    } else {
        mode = all_integer ? STD_I : STD_F;
    }
    if (mode != TPI_Q) {
        TOPO_CHECK(span * (trange + 1.0) < kU32, "size %d x DEM range %.0f overflows the 32-bit span sums", size, trange);
        p.tmin = (int)tlo;
        p.cmid = (int)(tlo + floor(trange / 2.0));
        vmax[0] = trange + 1.0;
        if (mode == STD_I || mode == STD_F) {
            const double half = floor(trange / 2.0) + 1.0;
            if (span * half * half < kU32) {
                vmax[1] = half * half;
            } else {
                // the square itself must fit 32 bits; its 16-bit halves then always fit the span sums (span <= 8191)
                TOPO_CHECK(half <= 65535.0, "DEM range %.0f too wide for the integer squares of std", trange);
                qsplit = 1;
                vmax[1] = 65535.0;
                vmax_qh = floor(half * half / 65536.0);
            }
        }
        if (mode == TPI_X || mode == STD_F) {
            int Sf = 30 - (ilog2_floor(span) + 1);
            if (Sf > 23) Sf = 23;
            p.fscale = (float)ldexp(1.0, Sf);
            p.inv_fscale = ldexp(1.0, -Sf);
            vmax[narr_of(mode) - 1] = ldexp(2.0, Sf);
        }
    }
    if (fft) {
        // the rounded FFT output must be the exact integer sum: planes whose magnitude would endanger that are split
        // (squares) or coarsened (fraction; 2^-12 m per pixel is still far inside the tolerance)
        const int T = dfft_length(plane_halo > 0 ? plane_halo : size / 2);
        const double nb = (double)disc_count(span_size);
        if ((mode == STD_I || mode == STD_F) && !qsplit && !dfft_exact(T, vmax[1], nb)) {
            const double half = floor(trange / 2.0) + 1.0;
            TOPO_CHECK(half <= 65535.0, "DEM range %.0f too wide for the integer squares of std", trange);
            qsplit = 1;
            vmax[1] = 65535.0;
            vmax_qh = floor(half * half / 65536.0);
        }
        const double half_q = floor(trange / 2.0) + 1.0;
        if (mode == TPI_I) {
            // the square plane rides along with T (and is kept for a following std): split it exactly when std would
            if (!(span * half_q * half_q < kU32) || !dfft_exact(T, half_q * half_q, nb)) qsplit = half_q <= 65535.0 ? 1 : 0;
        }
        pl.fft_mb0 = !all_integer ? PL_F : (half_q > 65535.0 ? -1 : (qsplit ? PL_QL : PL_Q));
        if (mode == TPI_X || mode == STD_F) {
            int Sf = ilog2_floor((double)p.fscale);
            while (Sf > 12 && !dfft_exact(T, ldexp(2.0, Sf), nb)) --Sf;
            p.fscale = (float)ldexp(1.0, Sf);
            p.inv_fscale = ldexp(1.0, -Sf);
        }
    }
    for (int a = 0; a < narr_of(mode); ++a)
        if (n * vmax[a] < kU32) acc |= 1 << a;
    pl.mode = mode;
    pl.cached = plane_halo > 0;
    if (fft) {
        // plane pairs: (T, Q | QL | F) and, for float DEMs or split squares, (Q | QL | QH, QH | -)
        const int pairs = (all_integer ? 1 : 2) + ((all_integer && qsplit) ? 1 : 0);
        plan_fft_geometry(v, size, narr_of(mode) + ((mode == STD_I || mode == STD_F) ? qsplit : 0), plane_halo, pairs, pl);
    } else {
        if (plan_geometry(v, size, narr_of(mode) + qsplit, max_rb(mode), pl, plane_halo, qsplit != 0)) return -1;
    }
    if (pl.fused) pl.cached = false;
    p.qsplit = qsplit;
    p.fplane = mode == TPI_X ? 1 : mode == STD_F ? 2 : -1;
    pl.acc_qh = qsplit && n * vmax_qh < kU32;
    pl.qh_kind = all_integer ? 2 : 4;
    pl.tiny = false;
    if (pl.fused) {
        // instantiated accumulator layouts of the fused kernels: none, plane 0 only, all planes
        const int full = (1 << narr_of(mode)) - 1;
        // tiny odd discs whose sums fit 32 bits: direct register sliding sums instead of prefix sums
        // (one loaded plane only: with two the 128 x 128 tile leaves a single CTA per SM and the prefix kernel wins)
        pl.tiny = (size & 1) && size >= 5 && size <= 13 && acc == full && (mode == TPI_I || mode == TPI_Q || mode == STD_I) &&
                  option_enabled(kOptTiny);
        if (mode == STD_F && (size & 1) && size >= 5 && size <= 13 && !qsplit && option_enabled(kOptTiny)) {
            // float std: integer part and fraction packed in one word, (t - tmin) << fpack | (frac + 1) * 2^Sf with
            // fpack = Sf + 2.  Exact -- and then bit-identical to the fused kernel's 2^23 fraction plane -- when every
            // float32 fraction has at most Sf bits, i.e. |z| >= 2^(23 - Sf) (or z = 0, the padding); the three sums of
            // a disc of N <= 137 cells must fit 32 bits.
            const int tb = ilog2_floor(trange + 1.0) + 1;
            int Sf = 30 - tb;
            if (Sf > 23) Sf = 23;
            const double zlow = ldexp(1.0, 23 - Sf);
            const double half = floor(trange / 2.0) + 1.0;
            if (Sf >= 12 && (zmin >= zlow || zmax <= -zlow) && n * ldexp(2.0, Sf) < kU32 && n * half * half < kU32 &&
                n * (trange + 1.0) < kU32) {
                pl.tiny = true;
                p.fpack = Sf + 2;
                p.fscale = (float)ldexp(1.0, Sf);
                p.inv_fscale = ldexp(1.0, -Sf);
                acc = full;
            }
        }
        if (acc != full) acc &= 1;
    }
    pl.acc = acc;
    p.n = n;
    p.n_ll = (long long)n;
    p.inv_n_nm1 = 1.0 / (n * (n - 1.0));
    p.exact64 = (n * (floor(trange / 2.0) + 1.0)) < 3.0e9;
    p.inv_nm1 = 1.0 / (n - 1.0);
    p.nc0_scaled = n * (double)p.c0i * p.inv_scale;
    p.n_tmin = n * (double)p.tmin;
    return 0;
}

// Fill the data-dependent constants.  what: 0 = TPI, 1 = STD.  With a plane cache the plan is first made for shared
// planes (scales of the largest size); a size that runs fused anyway, or whose shared layout does not fit the 32-bit
// span sums, gets its own plan.
static int plan_disc(const topo_view* v, int size, int what, int all_integer, double zmin, double zmax, DiscPlan& pl,
                     int cache_size = 0, bool pair = false) {
    if (cache_size >= size && size >= 2) {
        // a sweep whose largest size takes the FFT route keeps plane SPECTRA in its cache: only FFT sizes can use it
        const bool cache_is_fft = fft_route(cache_size, cache_size / 2);
        if (!cache_is_fft || size >= kDiscFftMinCached) {
            if (plan_disc_impl(v, size, what, all_integer, zmin, zmax, pl, cache_size, cache_size, pair) == 0 && !pl.fused) return 0;
        }
    }
    return plan_disc_impl(v, size, what, all_integer, zmin, zmax, pl, 0, size, pair);
}

static const char* mode_name(int mode) {
    switch (mode) {
        case TPI_Q: return "TPI_Q";
        case TPI_X: return "TPI_X";
        case TPI_I: return "TPI_I";
        case STD_I: return "STD_I";
        case STD_F: return "STD_F";
        case PL_T: return "T";
        case PL_Q: return "Q";
        case PL_QL: return "QL";
        case PL_QH: return "QH";
        default: return "F";
    }
}

static const char* kernel_label(const char* kind, int mode, int acc) {
    // stable storage for the profiler's kernel names
    static char names[64][48];
    static int used = 0;
    char buf[48];
    snprintf(buf, sizeof(buf), "%s<%s,acc%d>", kind, mode_name(mode), acc);
    for (int i = 0; i < used; ++i)
        if (strcmp(names[i], buf) == 0) return names[i];
    if (used == 64) return "disc";
    strcpy(names[used], buf);
    return names[used++];
}

template <int MODE, int ACC>
static int launch_fused(const DiscPlan& pl, cudaStream_t s) {
    const DiscParams& p = pl.p;
    static bool attr_set[64] = {false};
    int dev = 0;
    TOPO_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && !attr_set[dev]) {
        TOPO_CUDA(cudaFuncSetAttribute(disc_fused_kernel<MODE, ACC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(227 * 1024)));
        attr_set[dev] = true;
    }
    dim3 grid(p.tiles_x, p.tiles_y);
    TOPO_LAUNCH(kernel_label("disc_fused", MODE, ACC), s, disc_fused_kernel<MODE, ACC><<<grid, kThreads, pl.smem, s>>>(p));
    return 0;
}

template <int MODE>
static int launch_fused_acc(const DiscPlan& pl, cudaStream_t s) {
    constexpr int FULL = (1 << ModeTraits<MODE>::NARR) - 1;
    if (pl.acc == FULL) return launch_fused<MODE, FULL>(pl, s);
    if constexpr (FULL != 1) {
        if (pl.acc == 1) return launch_fused<MODE, 1>(pl, s);
    }
    return launch_fused<MODE, 0>(pl, s);
}

template <int MODE, int M>
static int launch_tiny(const DiscPlan& pl, cudaStream_t s) {
    const DiscParams& p = pl.p;
    constexpr int D = 2 * M + 1;
    constexpr int STEPS = ((kTinyStrip + 2 * M + D - 1) / D) * D;
    constexpr size_t smem = (size_t)TinyTraits<MODE>::LP * (kTinyStrip + STEPS) * (kTinyTile + 2 * M + 1) * 4;
    static bool attr_set[64] = {false};
    int dev = 0;
    TOPO_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && !attr_set[dev]) {
        TOPO_CUDA(cudaFuncSetAttribute(disc_tiny_kernel<MODE, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
        attr_set[dev] = true;
    }
    dim3 grid(ceil_div(p.nx, kTinyTile), ceil_div(p.out_rows, kTinyTile));
    TOPO_LAUNCH(kernel_label("disc_tiny", MODE, 2 * M + 1), s, disc_tiny_kernel<MODE, M><<<grid, 256, smem, s>>>(p));
    return 0;
}

template <int MODE>
static int launch_tiny_m(const DiscPlan& pl, cudaStream_t s) {
    switch (pl.p.mid) {
        case 2: return launch_tiny<MODE, 2>(pl, s);
        case 3: return launch_tiny<MODE, 3>(pl, s);
        case 4: return launch_tiny<MODE, 4>(pl, s);
        case 5: return launch_tiny<MODE, 5>(pl, s);
        default: return launch_tiny<MODE, 6>(pl, s);
    }
}

// One plane of the two-pass path: prefix planes (+ column prefix and summed-area table for the hybrid walk),
// then the gather kernel.  PM: plane mode; FIN: descriptor mode finished in place, or -1 for raw sums.
template <int PM, int FIN>
static int launch_plane(const DiscPlan& pl, int plane, bool acc32, cudaStream_t s, topo_disc_cache* cache = nullptr) {
    DiscParams p = pl.p;
    bool need_rows = true, need_cols = pl.hybrid;
    int publish = 0;  // cache->valid bits, set once every launch of this plane was accepted
    if (cache) {
        // plane kind 0 = trunc(z) - tmin, 1 = its centred square, 2 = fraction, 3 = quantised elevation (float DEMs):
        // one region each, same internal layout as the workspace
        // (+ the high half of a split square, after the kinds the DEM uses anyway: region 2 for integer-valued DEMs,
        // region 4 for float DEMs)
        const int kind = (PM == PL_Q || PM == PL_QL) ? 1 : PM == PL_F ? 2 : PM == TPI_Q ? 3 : PM == PL_QH ? pl.qh_kind : 0;
        TOPO_CHECK(cache->bytes >= (size_t)(kind + 1) * pl.off_partial, "plane cache too small: need %zu bytes, got %zu",
                   (size_t)(kind + 1) * pl.off_partial, cache->bytes);
        unsigned char* region = reinterpret_cast<unsigned char*>(cache->mem) + (size_t)kind * pl.off_partial;
        p.planes = reinterpret_cast<uint32_t*>(region);
        p.cplanes = reinterpret_cast<uint32_t*>(region + pl.off_cp);
        p.sat = reinterpret_cast<unsigned long long*>(region + pl.off_sat);
        p.eplanes = reinterpret_cast<uint32_t*>(region + pl.off_e);
        p.dplanes = reinterpret_cast<unsigned long long*>(region + pl.off_d);
        const int row_bit = 1 << (2 * kind), col_bit = 2 << (2 * kind);
        need_rows = !(cache->valid & row_bit);
        need_cols = pl.hybrid && !(cache->valid & col_bit);
        publish = row_bit | (pl.hybrid ? col_bit : 0);
    }
    const int warps = kThreads / 32;
    if (need_rows)
        TOPO_LAUNCH(kernel_label("disc_prefix", PM, 0), s,
                    disc_prefix_kernel<PM><<<ceil_div(pl.prefix_rows, warps), kThreads, 0, s>>>(p, pl.prefix_rows));
    const int grid = p.tiles_x * p.tiles_y;
    const size_t smem = pl.smem;
    if (pl.hybrid) {
        if (need_cols) {
            unsigned char* ws = reinterpret_cast<unsigned char*>(p.planes);
            uint32_t* tot_q = reinterpret_cast<uint32_t*>(ws + pl.off_totq);
            unsigned long long* tot_r = reinterpret_cast<unsigned long long*>(ws + pl.off_totr);
            dim3 cgrid(ceil_div(p.pitch / 4, 256), pl.nchunks);
            if (cache) {
                // octagon tables: D1 / D2 from the 64-bit row prefix (still in the summed-area buffer), E1 / E2 from the DEM
                unsigned long long* dtot = reinterpret_cast<unsigned long long*>(ws + pl.off_dtot);
                const int nd = pl.ndiag_threads;
                dim3 dgrid(ceil_div(nd, 256), pl.ndchunks, 2), sgrid(ceil_div(nd, 256), 2);
                TOPO_LAUNCH("disc_diag64_totals", s, (disc_diag_kernel<PM, true, 0><<<dgrid, 256, 0, s>>>(p, dtot, nd)));
                TOPO_LAUNCH("disc_diag_chunkscan", s, disc_diag_chunkscan_kernel<<<sgrid, 256, 0, s>>>(dtot, pl.ndchunks, nd));
                TOPO_LAUNCH("disc_diag64_apply", s, (disc_diag_kernel<PM, true, 1><<<dgrid, 256, 0, s>>>(p, dtot, nd)));
                TOPO_LAUNCH("disc_diag32_totals", s, (disc_diag_kernel<PM, false, 0><<<dgrid, 256, 0, s>>>(p, dtot, nd)));
                TOPO_LAUNCH("disc_diag_chunkscan", s, disc_diag_chunkscan_kernel<<<sgrid, 256, 0, s>>>(dtot, pl.ndchunks, nd));
                TOPO_LAUNCH("disc_diag32_apply", s, (disc_diag_kernel<PM, false, 1><<<dgrid, 256, 0, s>>>(p, dtot, nd)));
            }
            TOPO_LAUNCH("disc_colsum", s, disc_colsum_kernel<PM><<<cgrid, 256, 0, s>>>(p, tot_q, tot_r, pl.nchunks));
            TOPO_LAUNCH("disc_chunkscan", s,
                        disc_chunkscan_kernel<<<ceil_div(p.pitch, 256), 256, 0, s>>>(tot_q, tot_r, pl.nchunks, p.pitch, 1));
            TOPO_LAUNCH("disc_colapply", s, disc_colapply_kernel<PM><<<cgrid, 256, 0, s>>>(p, tot_q, tot_r, pl.nchunks));
        }
        if (acc32)
            TOPO_LAUNCH(kernel_label("disc_hybrid", PM, 1), s, disc_span_kernel<true, true, FIN><<<grid, kThreads, smem, s>>>(p, plane));
        else
            TOPO_LAUNCH(kernel_label("disc_hybrid", PM, 0), s, disc_span_kernel<false, true, FIN><<<grid, kThreads, smem, s>>>(p, plane));
    } else {
        if (acc32)
            TOPO_LAUNCH(kernel_label("disc_span", PM, 1), s, disc_span_kernel<true, false, FIN><<<grid, kThreads, smem, s>>>(p, plane));
        else
            TOPO_LAUNCH(kernel_label("disc_span", PM, 0), s, disc_span_kernel<false, false, FIN><<<grid, kThreads, smem, s>>>(p, plane));
    }
    if (cache) cache->valid |= publish;
    return 0;
}

// tsum_op: 0 = no sharing, 1 = compute the T plane and keep its raw sums in p.tsum, 2 = reuse p.tsum
static int launch_two_pass(const DiscPlan& pl, int tsum_op, cudaStream_t s, topo_disc_cache* cache) {
    const bool a0 = pl.acc & 1, a1 = (pl.acc >> 1) & 1, a2 = (pl.acc >> 2) & 1;
    const bool reuse = tsum_op == 2;
    int rc = 0;
    switch (pl.mode) {
        case TPI_Q: return launch_plane<TPI_Q, TPI_Q>(pl, 0, a0, s, cache);
        case TPI_I:
            if (!reuse) return launch_plane<PL_T, TPI_I>(pl, 0, a0, s, cache);
            TOPO_LAUNCH("disc_finish<TPI_I>", s, disc_finish_kernel<TPI_I><<<kNumSMs * 8, 256, 0, s>>>(pl.p));
            return 0;
        case TPI_X:
            if (!reuse && (rc = launch_plane<PL_T, -1>(pl, 0, a0, s, cache))) return rc;
            if (!(reuse && pl.p.fsum) && (rc = launch_plane<PL_F, -1>(pl, 1, a1, s, cache))) return rc;
            TOPO_LAUNCH("disc_finish<TPI_X>", s, disc_finish_kernel<TPI_X><<<kNumSMs * 8, 256, 0, s>>>(pl.p));
            return 0;
        case STD_I:
            if (!reuse && (rc = launch_plane<PL_T, -1>(pl, 0, a0, s, cache))) return rc;
            if (pl.p.qsplit) {
                if ((rc = launch_plane<PL_QL, -1>(pl, 1, a1, s, cache))) return rc;
                if ((rc = launch_plane<PL_QH, -1>(pl, 2, pl.acc_qh, s, cache))) return rc;
            } else if ((rc = launch_plane<PL_Q, -1>(pl, 1, a1, s, cache))) {
                return rc;
            }
            TOPO_LAUNCH("disc_finish<STD_I>", s, disc_finish_kernel<STD_I><<<kNumSMs * 8, 256, 0, s>>>(pl.p));
            return 0;
        default:
            if (!reuse && (rc = launch_plane<PL_T, -1>(pl, 0, a0, s, cache))) return rc;
            if (pl.p.qsplit) {
                if ((rc = launch_plane<PL_QL, -1>(pl, 1, a1, s, cache))) return rc;
                if ((rc = launch_plane<PL_QH, -1>(pl, 3, pl.acc_qh, s, cache))) return rc;
            } else if ((rc = launch_plane<PL_Q, -1>(pl, 1, a1, s, cache))) {
                return rc;
            }
            if (!(reuse && pl.p.fsum) && (rc = launch_plane<PL_F, -1>(pl, 2, a2, s, cache))) return rc;
            TOPO_LAUNCH("disc_finish<STD_F>", s, disc_finish_kernel<STD_F><<<kNumSMs * 8, 256, 0, s>>>(pl.p));
            return 0;
    }
}

// ---- FFT route: launches ------------------------------------------------------------------------------------------
// NX / NY: transform lengths along x (lines = window rows) and along y (lines = window columns)
template <int NX, int NY>
static int launch_fft_route_n(const DiscPlan& pl, int tsum_op, cudaStream_t s, topo_disc_cache* cache, unsigned char* ws) {
    using SX = FftShape<NX>;
    using SY = FftShape<NY>;
    const DiscParams& p = pl.p;
    const DfftGeom& g = pl.fg;
    static bool attr_set[64] = {false};
    int dev = 0;
    TOPO_CUDA(cudaGetDevice(&dev));
    if (dev >= 64 || !attr_set[dev]) {
        TOPO_CUDA(cudaFuncSetAttribute(dfft_fwd_planes_kernel<NX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SX::SMEM));
        TOPO_CUDA(cudaFuncSetAttribute(dfft_fwd_disc_kernel<NX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SX::SMEM));
        TOPO_CUDA(cudaFuncSetAttribute((dfft_store_kernel<NX, -1, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SX::SMEM));
        TOPO_CUDA(cudaFuncSetAttribute((dfft_store_kernel<NX, TPI_I, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SX::SMEM));
        TOPO_CUDA(cudaFuncSetAttribute((dfft_store_kernel<NX, TPI_X, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SX::SMEM));
        TOPO_CUDA(cudaFuncSetAttribute((dfft_store_kernel<NX, STD_I, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SX::SMEM));
        TOPO_CUDA(cudaFuncSetAttribute((dfft_store_kernel<NX, STD_F, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SX::SMEM));
        TOPO_CUDA(cudaFuncSetAttribute((dfft_store_kernel<NX, STD_I, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SX::SMEM));
        TOPO_CUDA(cudaFuncSetAttribute((dfft_store_kernel<NX, STD_F, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SX::SMEM));
        if (fft2d_set_smem_attributes<NY>()) return -2;
        if (dev < 64) attr_set[dev] = true;
    }
    double2* twx = reinterpret_cast<double2*>(ws + pl.off_tw);
    double2* twy = NX == NY ? twx : twx + NX;
    double2* X = reinterpret_cast<double2*>(ws + pl.off_x);
    double2* Y = reinterpret_cast<double2*>(ws + pl.off_y);
    double2* K1 = reinterpret_cast<double2*>(ws + pl.off_k1);
    double2* K2 = reinterpret_cast<double2*>(ws + pl.off_k2);
    const size_t spectra_bytes = (size_t)pl.fft_pairs * pl.fft_plane_bytes;
    const size_t mask_bytes = (size_t)NX * NY * sizeof(double2);
    bool have_mask = false;
    if (cache) {
        // the mask spectrum lives behind the plane spectra: tpi(size) and std(size) of a pair build it once
        TOPO_CHECK(cache->bytes >= spectra_bytes + mask_bytes, "plane cache too small: need %zu bytes, got %zu",
                   spectra_bytes + mask_bytes, cache->bytes);
        K1 = reinterpret_cast<double2*>(reinterpret_cast<unsigned char*>(cache->mem) + spectra_bytes);
        have_mask = cache->mask_size == p.k;
    }
    const int tiles = g.tiles_y * g.tiles_x;
    TOPO_CHECK(tiles <= 65535, "too many tiles for one launch: split the DEM in row bands");
    const bool reuse = tsum_op == 2;  // a previous call of the pair left the sums of plane pair 0 in tsum / qsum / fsum
    const int q_lo = p.qsplit ? PL_QL : PL_Q;
    unsigned long long* part = p.partial;
    const int64_t ps = p.partial_stride;

    // the plane pairs this descriptor needs: {pair index, mode a, mode b, destination a, destination b}.
    // A plane that travels alone in pair 1 (the squares of a float DEM, the high half of split squares) is laid out as
    // twin tiles: two tiles of the same plane per complex plane (see dfft_fwd_planes_kernel).
    struct Job { int pair, ma, mb; unsigned long long *da, *db; };
    Job jobs[2];
    int njobs = 0;
    switch (pl.mode) {
        // (the LAST job finishes the descriptor in its store pass and only keeps raw sums a later call will reuse;
        // earlier jobs leave their sums in tsum / qsum / fsum or the workspace)
        case TPI_I:
            if (!reuse) jobs[njobs++] = {0, PL_T, pl.fft_mb0, p.tsum, pl.fft_mb0 >= 0 ? p.qsum : nullptr};
            break;
        case TPI_X:
            if (!reuse) jobs[njobs++] = {0, PL_T, PL_F, p.tsum, p.fsum};
            break;
        case STD_I:
            if (!reuse) {
                if (p.qsplit)
                    jobs[njobs++] = {0, PL_T, q_lo, p.tsum ? p.tsum : part, p.qsum ? p.qsum : part + ps};
                else
                    jobs[njobs++] = {0, PL_T, q_lo, p.tsum, p.qsum};
            }
            if (p.qsplit) jobs[njobs++] = {1, PL_QH, -1, nullptr, nullptr};
            break;
        default:  // STD_F
            if (!reuse) jobs[njobs++] = {0, PL_T, PL_F, p.tsum ? p.tsum : part, p.fsum ? p.fsum : part + 2 * ps};
            jobs[njobs++] = {1, q_lo, p.qsplit ? PL_QH : -1, nullptr, nullptr};
            break;
    }
    // window planes are [NY rows][NX columns]; their spectra and the first inverse pass work on [NX lines][NY]
    if (njobs > 0) {
        TOPO_LAUNCH("disc_fft_twiddles", s, fft_twiddle_kernel<<<ceil_div(NX, 256), 256, 0, s>>>(twx, NX));
        if (NX != NY) TOPO_LAUNCH("disc_fft_twiddles", s, fft_twiddle_kernel<<<ceil_div(NY, 256), 256, 0, s>>>(twy, NY));
        if (!have_mask) {
            // spectrum of the disc mask (its lines stored by eights: the order the product pass reads them in)
            if (cache) cache->mask_size = 0;
            TOPO_LAUNCH("disc_fft_mask", s, (dfft_fwd_disc_kernel<NX><<<dim3(NY, 1), SX::NT, SX::SMEM, s>>>(p, g, K1, twx)));
            TOPO_LAUNCH("disc_fft_transpose", s, fft2d_transpose_kernel<<<dim3(NX / 32, NY / 32, 1), dim3(32, 8), 0, s>>>(K1, K2, NY, NX));
            TOPO_LAUNCH("disc_fft_fwd", s, (fft2d_fwd_cplx_kernel<NY, true><<<dim3(NX, 1), SY::NT, SY::SMEM, s>>>(K2, K1, twy)));
            if (cache) cache->mask_size = p.k;
        }
    }
    const double scale = 1.0 / ((double)NX * (double)NY);
    for (int j = 0; j < njobs; ++j) {
        const Job& job = jobs[j];
        const bool twin = job.pair == 1 && job.mb < 0;
        const int planes = twin ? (tiles + 1) / 2 : tiles;  // complex planes of this job
        double2* dhat;
        bool have = false;
        if (cache) {
            dhat = reinterpret_cast<double2*>(reinterpret_cast<unsigned char*>(cache->mem) + (size_t)job.pair * pl.fft_plane_bytes);
            have = (cache->valid >> (16 + job.pair)) & 1;
        } else {
            dhat = reinterpret_cast<double2*>(ws + pl.off_dhat + (size_t)job.pair * pl.fft_plane_bytes);
        }
        if (!have) {
            TOPO_LAUNCH("disc_fft_planes", s, (dfft_fwd_planes_kernel<NX><<<dim3(NY, planes), SX::NT, SX::SMEM, s>>>(p, g, job.ma, job.mb, twin ? 1 : 0, X, twx)));
            TOPO_LAUNCH("disc_fft_transpose", s, fft2d_transpose_kernel<<<dim3(NX / 32, NY / 32, planes), dim3(32, 8), 0, s>>>(X, Y, NY, NX));
            TOPO_LAUNCH("disc_fft_fwd", s, (fft2d_fwd_cplx_kernel<NY><<<dim3(NX, planes), SY::NT, SY::SMEM, s>>>(Y, dhat, twy)));
            if (cache) cache->valid |= 1 << (16 + job.pair);
        }
        // only the window rows the store pass turns into pixels travel through the first pass's stores and the transpose
        const int ct0 = (2 * g.H) / 32, ct1 = ceil_div(2 * g.H + g.Vy, 32);
        TOPO_LAUNCH("disc_fft_inv", s, (fft2d_inv_product_kernel<NY><<<dim3(planes, NX), SY::NT, SY::SMEM, s>>>(dhat, K1, X, twy, ct0 * 32, ct1 * 32)));
        TOPO_LAUNCH("disc_fft_transpose", s, fft2d_transpose_kernel<<<dim3(ct1 - ct0, NX / 32, planes), dim3(32, 8), 0, s>>>(X, Y, NX, NY, ct0));
        const dim3 sgrid(NY, planes);
#define TOPO_DFFT_STORE(LABEL, FIN, TWIN) \
    TOPO_LAUNCH(LABEL, s, (dfft_store_kernel<NX, FIN, TWIN><<<sgrid, SX::NT, SX::SMEM, s>>>(p, g, Y, twx, job.da, job.db, job.ma, job.mb, scale)))
        if (j + 1 < njobs) {
            TOPO_CHECK(!twin, "internal: a twin-tile job must be the last one of its descriptor");
            TOPO_DFFT_STORE("disc_fft_store", -1, false);
        } else if (twin) {
            if (pl.mode == STD_I) {
                TOPO_DFFT_STORE("disc_fft_finish<STD_I,twin>", STD_I, true);
            } else {
                TOPO_CHECK(pl.mode == STD_F, "internal: twin tiles outside std");
                TOPO_DFFT_STORE("disc_fft_finish<STD_F,twin>", STD_F, true);
            }
        } else {
            switch (pl.mode) {
                case TPI_I: TOPO_DFFT_STORE("disc_fft_finish<TPI_I>", TPI_I, false); break;
                case TPI_X: TOPO_DFFT_STORE("disc_fft_finish<TPI_X>", TPI_X, false); break;
                case STD_I: TOPO_DFFT_STORE("disc_fft_finish<STD_I>", STD_I, false); break;
                default: TOPO_DFFT_STORE("disc_fft_finish<STD_F>", STD_F, false); break;
            }
        }
#undef TOPO_DFFT_STORE
    }
    if (njobs == 0) {  // every plane sum was left by the other descriptor of the pair: finish only
        switch (pl.mode) {
            case TPI_I: TOPO_LAUNCH("disc_finish<TPI_I>", s, disc_finish_kernel<TPI_I><<<kNumSMs * 8, 256, 0, s>>>(p)); break;
            case TPI_X: TOPO_LAUNCH("disc_finish<TPI_X>", s, disc_finish_kernel<TPI_X><<<kNumSMs * 8, 256, 0, s>>>(p)); break;
            case STD_I: TOPO_LAUNCH("disc_finish<STD_I>", s, disc_finish_kernel<STD_I><<<kNumSMs * 8, 256, 0, s>>>(p)); break;
            default: TOPO_LAUNCH("disc_finish<STD_F>", s, disc_finish_kernel<STD_F><<<kNumSMs * 8, 256, 0, s>>>(p)); break;
        }
    }
    return 0;
}

static int launch_fft_route(const DiscPlan& pl, int tsum_op, cudaStream_t s, topo_disc_cache* cache, unsigned char* ws) {
    switch (pl.fg.T) {
        case 2048: return launch_fft_route_n<2048, 2048>(pl, tsum_op, s, cache, ws);
        case 4096:
            return pl.fg.Ty == 3072 ? launch_fft_route_n<4096, 3072>(pl, tsum_op, s, cache, ws)
                                    : launch_fft_route_n<4096, 4096>(pl, tsum_op, s, cache, ws);
        default: return launch_fft_route_n<8192, 8192>(pl, tsum_op, s, cache, ws);
    }
}

static int run_disc(const float* dem, int64_t ld_in, float* out, int64_t ld_out, const topo_view* v, int size,
                    int what, int all_integer, double zmin, double zmax, unsigned long long* tsum, int tsum_op,
                    topo_disc_cache* cache, void* ws, size_t ws_bytes, void* stream) {
    TOPO_CHECK(dem && out, "null pointer");
    if (validate_view(v)) return -1;
    TOPO_CHECK(ld_in >= v->nx && ld_out >= v->nx, "row pitch smaller than nx");
    if (v->out_rows == 0) return 0;
    if (size == 1) {
        // kernel sum - 1 == 0: the reference divides by zero and returns NaN everywhere
        return topo_fill_f32(out, v->out_rows, v->nx, ld_out, NAN, stream);
    }
    DiscPlan pl;
    if (cache && !cache->mem) cache = nullptr;
    if (plan_disc(v, size, what, all_integer, zmin, zmax, pl, cache ? cache->max_size : 0, tsum_op != 0)) return -1;
    if (!pl.cached) cache = nullptr;
    if (check_band(v, pl.p.halo)) return -1;
    pl.p.dem = dem, pl.p.out = out, pl.p.ld_in = ld_in, pl.p.ld_out = ld_out;
    cudaStream_t s = (cudaStream_t)stream;
    if (pl.fft) {
        TOPO_CHECK(ws != nullptr && ws_bytes >= pl.ws_bytes, "workspace too small: need %zu bytes, got %zu", pl.ws_bytes, ws_bytes);
        TOPO_CHECK((reinterpret_cast<uintptr_t>(ws) & 255) == 0, "workspace must be 256-byte aligned");
        if (cache) TOPO_CHECK((reinterpret_cast<uintptr_t>(cache->mem) & 255) == 0, "plane cache must be 256-byte aligned");
        pl.p.partial = reinterpret_cast<unsigned long long*>((unsigned char*)ws + pl.off_partial);
        if (tsum_op != 0) {
            TOPO_CHECK(tsum != nullptr, "tsum_op %d needs a plane-sum buffer", tsum_op);
            pl.p.tsum = tsum;
            // the second shared plane sits right behind the first (topo_disc_shares_tsum == 2): the fraction plane of
            // float DEMs, the (low half of the) square plane of integer-valued ones
            if (!all_integer)
                pl.p.fsum = tsum + (int64_t)v->out_rows * v->nx;
            else
                pl.p.qsum = tsum + (int64_t)v->out_rows * v->nx;
        }
        return launch_fft_route(pl, tsum_op, s, cache, reinterpret_cast<unsigned char*>(ws));
    }
    if (!pl.fused) {
        // with a plane cache the workspace only holds the raw sums of the multi-plane modes
        const size_t ws_need = cache ? pl.ws_bytes - pl.off_partial : pl.ws_bytes;
        TOPO_CHECK(ws_need == 0 || (ws != nullptr && ws_bytes >= ws_need), "workspace too small: need %zu bytes, got %zu",
                   ws_need, ws_bytes);
        if (cache) {
            TOPO_CHECK((reinterpret_cast<uintptr_t>(cache->mem) & 255) == 0, "plane cache must be 256-byte aligned");
        }
        TOPO_CHECK((reinterpret_cast<uintptr_t>(ws) & 31) == 0, "workspace must be 32-byte aligned");
        TOPO_CHECK((long long)pl.p.tiles_x * pl.p.tiles_y < 2147483647ll, "too many tiles");
        TOPO_CHECK((long long)pl.p.plane_stride + pl.p.pitch < 2147483647ll && (long long)pl.p.cplane_stride + pl.p.pitch < 2147483647ll,
                   "band too large for 32-bit plane offsets: split the DEM in row bands");
        pl.p.planes = (uint32_t*)ws;
        if (pl.hybrid) {
            pl.p.cplanes = reinterpret_cast<uint32_t*>((unsigned char*)ws + pl.off_cp);
            pl.p.sat = reinterpret_cast<unsigned long long*>((unsigned char*)ws + pl.off_sat);
        }
        pl.p.partial = reinterpret_cast<unsigned long long*>((unsigned char*)ws + (cache ? 0 : pl.off_partial));
        if (tsum_op != 0) {
            TOPO_CHECK(tsum != nullptr, "tsum_op %d needs a T-plane sum buffer", tsum_op);
            TOPO_CHECK(pl.mode != TPI_Q, "this size/DEM does not use the T plane (see topo_disc_shares_tsum)");
            pl.p.tsum = tsum;
            // float DEMs keep the fraction-plane sums right behind the T-plane sums (topo_disc_shares_tsum == 2)
            if (!all_integer) pl.p.fsum = tsum + (int64_t)v->out_rows * v->nx;
        }
        return launch_two_pass(pl, tsum_op, s, cache);
    }
    TOPO_CHECK(tsum_op == 0, "T-plane sums are only shared on the two-pass path (see topo_disc_shares_tsum)");
    if (pl.tiny) {
        switch (pl.mode) {
            case TPI_Q: return launch_tiny_m<TPI_Q>(pl, s);
            case TPI_I: return launch_tiny_m<TPI_I>(pl, s);
            case TPI_X: return launch_tiny_m<TPI_X>(pl, s);
            case STD_I: return launch_tiny_m<STD_I>(pl, s);
            default: return launch_tiny_m<STD_F>(pl, s);
        }
    }
    switch (pl.mode) {
        case TPI_Q: return launch_fused_acc<TPI_Q>(pl, s);
        case TPI_I: return launch_fused_acc<TPI_I>(pl, s);
        case TPI_X: return launch_fused_acc<TPI_X>(pl, s);
        case STD_I: return launch_fused_acc<STD_I>(pl, s);
        default: return launch_fused_acc<STD_F>(pl, s);
    }
}

}  // namespace topo

using namespace topo;

extern "C" {

// The three queries below make the SAME plan as run_disc (same range, integrality and cache size), so what they
// report is what the call will do -- including its fall-back to an un-cached plan when the shared layout overflows.
static bool quiet_plan(const topo_view* v, int size, int what, int all_integer, double zmin, double zmax, int cache_max_size,
                       DiscPlan& pl, bool pair = false) {
    if (!v || size < 2 || size > kMaxSize || validate_view(v)) return false;
    memset(&pl, 0, sizeof(pl));
    return plan_disc(v, size, what, all_integer, zmin, zmax, pl, cache_max_size, pair) == 0;
}

size_t topo_disc_workspace_bytes(const topo_view* v, int size, int what, int all_integer, double zmin, double zmax,
                                 int cache_max_size, int tsum_op) {
    DiscPlan pl;
    if (!quiet_plan(v, size, what, all_integer, zmin, zmax, cache_max_size, pl, tsum_op != 0)) return 0;
    if (pl.fft) return pl.ws_bytes;  // (already without the plane spectra when they live in the cache)
    if (pl.fused) return 0;
    // with a plane cache the planes live there and the workspace only holds the raw plane sums
    return pl.cached ? pl.ws_bytes - pl.off_partial : pl.ws_bytes;
}

int topo_disc_shares_tsum(const topo_view* v, int size, int all_integer, double zmin, double zmax, int cache_max_size) {
    // tpi and std of the same size both run two-pass on the same planes => std can reuse tpi's raw plane sums:
    // the T plane of integer-valued DEMs (returns 1), the T and fraction planes of float DEMs (returns 2: tpi then
    // runs as the exact two-plane TPI_X instead of the one-plane quantised TPI_Q -- 3 gather passes per pair, not 4)
    DiscPlan a, b;
    if (!quiet_plan(v, size, 0, all_integer, zmin, zmax, cache_max_size, a, true)) return 0;
    if (!quiet_plan(v, size, 1, all_integer, zmin, zmax, cache_max_size, b, true)) return 0;
    if (a.fused || b.fused || a.p.tmin != b.p.tmin || a.fft != b.fft) return 0;
    if (a.mode == TPI_I && b.mode == STD_I) return a.fft ? 2 : 1;  // FFT route: the square plane rides along with T
    if (a.mode == TPI_X && b.mode == STD_F && a.p.fscale == b.p.fscale) return 2;
    return 0;
}

size_t topo_disc_cache_bytes(const topo_view* v, int max_size, int all_integer, double zmin, double zmax) {
    DiscPlan pl;
    if (!quiet_plan(v, max_size, 1, all_integer, zmin, zmax, max_size, pl)) return 0;
    // plane spectra, one per pair, + the mask spectrum of the size in flight
    if (pl.fft) return pl.cached ? (size_t)pl.fft_pairs * pl.fft_plane_bytes + (size_t)pl.fg.T * pl.fg.Ty * sizeof(double2) : 0;
    if (pl.fused || !pl.cached) return 0;
    // integer-valued DEMs use two plane kinds (trunc(z) - tmin, its square); float DEMs add the fraction and the
    // quantised-elevation planes; a split square adds its high half
    return (size_t)((all_integer ? 2 : 4) + (pl.p.qsplit ? 1 : 0)) * pl.off_partial;
}

int topo_disc_plan_info(const topo_view* v, int size, int what, int all_integer, double zmin, double zmax,
                        int cache_max_size, int tsum_op, long long* info) {
    // host-only introspection of the plan run_disc would execute (no launch, no GPU needed): used by the CPU tests
    TOPO_CHECK(info != nullptr, "null pointer");
    if (validate_view(v)) return -1;
    DiscPlan pl;
    memset(&pl, 0, sizeof(pl));
    if (plan_disc(v, size, what, all_integer, zmin, zmax, pl, cache_max_size, tsum_op != 0)) return -1;
    info[0] = pl.mode, info[1] = pl.fused, info[2] = pl.hybrid, info[3] = pl.tiny, info[4] = pl.cached;
    info[5] = pl.p.oct, info[6] = pl.p.asq, info[7] = pl.p.oct_v, info[8] = pl.p.oct_ndiag, info[9] = pl.acc;
    info[10] = (long long)pl.smem, info[11] = pl.p.halo, info[12] = pl.p.pitch, info[13] = pl.prefix_rows;
    info[14] = (long long)pl.ws_bytes, info[15] = (long long)pl.off_partial, info[16] = pl.p.qsplit;
    info[17] = pl.fft, info[18] = pl.fft ? pl.fg.T : 0, info[19] = pl.fft ? pl.fg.tiles_y * pl.fg.tiles_x : 0;
    return 0;
}

int topo_tpi_f32(const float* dem, int64_t ld_in, float* out, int64_t ld_out, const topo_view* v, int size,
                 int all_integer, double zmin, double zmax, unsigned long long* tsum, int tsum_op,
                 topo_disc_cache* cache, void* ws, size_t ws_bytes, void* stream) {
    return run_disc(dem, ld_in, out, ld_out, v, size, 0, all_integer, zmin, zmax, tsum, tsum_op, cache, ws, ws_bytes,
                    stream);
}

int topo_std_f32(const float* dem, int64_t ld_in, float* out, int64_t ld_out, const topo_view* v, int size,
                 int all_integer, double zmin, double zmax, unsigned long long* tsum, int tsum_op,
                 topo_disc_cache* cache, void* ws, size_t ws_bytes, void* stream) {
    return run_disc(dem, ld_in, out, ld_out, v, size, 1, all_integer, zmin, zmax, tsum, tsum_op, cache, ws, ws_bytes,
                    stream);
}

}  // extern "C"
