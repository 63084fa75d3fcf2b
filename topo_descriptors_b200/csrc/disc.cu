// TPI and STD: zero-padded disc sums (reference: topo.py:144-181 tpi, 272-307 std, 191-213
// circular_kernel; scipy.signal.convolve(mode="same") centring).
//
// Method (north star: "per-row prefix-sum span-sum kernel"):
//   every DEM value becomes one or more unsigned 32-bit integers (fixed point / exact integer parts),
//   each tile row gets an exclusive prefix sum in WRAP-AROUND uint32 arithmetic, and a disc sum is
//   sum over kernel rows of  P[row][hi+1] - P[row][lo]  (exact modulo 2^32, and the true span sum
//   is < 2^32 by construction of the scale), accumulated in 64-bit integers.  The epilogue is
//   float64.  Integer arithmetic makes the result independent of tiling and of the row-band
//   partition (bit-identical on 1 or 8 GPUs).
//
// Modes (what the uint32 planes hold):
//   TPI_Q : q = rn(z * 2^S) - c0*2^S                 1 plane, |error of the mean| <= 2^-(S+1)
//   TPI_X : t - tmin,  (frac + 1) * 2^Sf             2 planes, exact (used when S would be < 10)
//   STD_I : t - tmin,  (t - cmid)^2                  2 planes, exact (integer-valued DEM)
//   STD_F : t - tmin,  (t - cmid)^2, (frac+1)*2^Sf   3 planes, exact
//   with t = trunc(z) (the reference's astype("int32"), topo.py:300) and frac = z - t.
//
// Two execution shapes:
//   fused    : CTA = 128-column x TH-row output tile; tile + halo is loaded with 128-bit loads,
//              converted, scanned with warp shuffles into shared memory, then each thread walks the
//              kernel rows for RB output rows of one column (conflict-free LDS, lanes = columns).
//   two-pass : for discs whose halo does not fit shared memory the prefix planes go to global
//              memory (workspace) and the span walk reads them through L1/L2.
#include <math.h>
#include <string.h>

#include <type_traits>

#include "common.cuh"

namespace topo {

enum DiscMode { TPI_Q = 0, TPI_X = 1, STD_I = 2, STD_F = 3 };

template <int MODE>
struct ModeTraits;
template <>
struct ModeTraits<TPI_Q> { static constexpr int NARR = 1; static constexpr int RB = 8; };
template <>
struct ModeTraits<TPI_X> { static constexpr int NARR = 2; static constexpr int RB = 4; };
template <>
struct ModeTraits<STD_I> { static constexpr int NARR = 2; static constexpr int RB = 4; };
template <>
struct ModeTraits<STD_F> { static constexpr int NARR = 3; static constexpr int RB = 4; };

constexpr int kTW = 128;        // output tile width  (4 warps of columns)
constexpr int kThreads = 256;   // 8 warps: 4 across x, 2 row groups
constexpr int kMaxSize = 8191;  // span table lives in shared memory (4 B per kernel row)

struct DiscParams {
    const float* dem;
    float* out;
    uint32_t* planes;  // two-pass: global prefix planes; fused: unused
    int64_t ld_in, ld_out;
    int64_t plane_stride;  // elements between planes
    int nx, gny, in_gy0, in_rows, out_gy0, out_rows;
    int k;       // kernel size
    int c;       // (k-1)/2 : scipy 'same' centring
    int mid;     // k/2     : circular_kernel's middle
    int square;  // size < 5
    int halo;    // k/2, rows/cols of halo on each side
    int haloL;   // left halo rounded up to a multiple of 4 (keeps 128-bit loads aligned)
    int TH;      // tile rows (fused) / rows per CTA (two-pass)
    int pitch;   // prefix row pitch in elements
    int prow0;   // two-pass: global row of prefix row 0
    // conversion constants
    float scale;   // 2^S
    int c0i;       // c0 * 2^S            (TPI_Q)
    int tmin;      // offset of the integer part
    int cmid;      // offset inside the square
    float fscale;  // 2^Sf
    // epilogue constants
    double inv_scale, inv_fscale, n, inv_nm1;
    long long n_ll;
    double inv_n_nm1;  // 1 / (N * (N - 1))
    int exact64;       // N*B - a^2 fits in 64-bit integers
    int excl;  // TPI: offset (excl, excl) of the excluded "mid point" (0 odd size, -1 even size)
};

// ---- span table: kernel row i -> (dxlo, dxhi) of its run of ones, as DEM column offsets ------------
// out[y,x] = sum_{i,j} K[i,j] * d[y + c - i, x + c - j]
__device__ __forceinline__ void kernel_row_span(const DiscParams& p, int i, int& dxlo, int& dxhi) {
    int jlo, jhi;
    if (p.square) {
        jlo = 0;
        jhi = p.k - 1;
    } else {
        int di = i - p.mid;
        int rem = p.mid * p.mid - di * di;  // >= 0 for every row of the kernel
        int w = (int)floorf(sqrtf((float)rem));
        while (w * w > rem) --w;
        while ((w + 1) * (w + 1) <= rem) ++w;
        jlo = p.mid - w;
        jhi = p.mid + w;
        if (jhi > p.k - 1) jhi = p.k - 1;  // even sizes: the disc is clipped on the right/bottom
    }
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
// prefix dst[a][0..W] of each plane a (dst rows have `pitch` elements, 16-byte aligned).
template <int MODE>
__device__ __forceinline__ void scan_row(const DiscParams& p, int gy, int xs, int W, uint32_t* dst,
                                         int64_t plane_stride, bool aligned, int lane) {
    constexpr int NARR = ModeTraits<MODE>::NARR;
    uint32_t carry[NARR];
#pragma unroll
    for (int a = 0; a < NARR; ++a) carry[a] = 0u;
    const int nchunks = (W + 127) >> 7;
    for (int ch = 0; ch < nchunks; ++ch) {
        const int lc = (ch << 7) + (lane << 2);
        float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        if (lc < W) z = load4_zero(p, gy, xs + lc, aligned);
        uint32_t v0[NARR], v1[NARR], v2[NARR], v3[NARR];
        convert<MODE>(p, z.x, v0);
        convert<MODE>(p, z.y, v1);
        convert<MODE>(p, z.z, v2);
        convert<MODE>(p, z.w, v3);
#pragma unroll
        for (int a = 0; a < NARR; ++a) {
            // elements past W must not contribute (they convert to a non-zero "zero" otherwise)
            uint32_t e0 = (lc + 0 < W) ? v0[a] : 0u;
            uint32_t e1 = (lc + 1 < W) ? v1[a] : 0u;
            uint32_t e2 = (lc + 2 < W) ? v2[a] : 0u;
            uint32_t e3 = (lc + 3 < W) ? v3[a] : 0u;
            const uint32_t s1 = e0 + e1, s2 = s1 + e2, tot = s2 + e3;
            const uint32_t incl = warp_inclusive_scan_u32(tot, lane);
            const uint32_t base = carry[a] + incl - tot;  // exclusive prefix at element lc
            if (lc < W) {
                uint4 o = make_uint4(base, base + e0, base + s1, base + s2);
                *reinterpret_cast<uint4*>(dst + a * plane_stride + lc) = o;
            }
            carry[a] += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
    // closing element P[W4] (W4 = W rounded up to 4): total of the row
    if (lane == 0) {
        const int W4 = (W + 3) & ~3;
#pragma unroll
        for (int a = 0; a < NARR; ++a) dst[a * plane_stride + W4] = carry[a];
    }
}

// ---- epilogue -------------------------------------------------------------------------------------
template <int MODE>
__device__ __forceinline__ float finish(const DiscParams& p, const unsigned long long (&acc)[ModeTraits<MODE>::NARR],
                                        int gy, int x) {
    const double n = p.n;
    if constexpr (MODE == TPI_Q || MODE == TPI_X) {
        double sum_z;
        if constexpr (MODE == TPI_Q) {
            const long long tot = (long long)acc[0] + p.n_ll * (long long)p.c0i;
            sum_z = (double)tot * p.inv_scale;
        } else {
            const long long st = (long long)acc[0] + p.n_ll * (long long)p.tmin;
            sum_z = (double)st + ((double)acc[1] * p.inv_fscale - n);
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

// ---- the span walk: RB output rows of one column -----------------------------------------------------
// P: prefix planes; row index `prow` of the prefix row that corresponds to the FIRST of the RB output
// rows at kernel offset dy = 0; lcx: prefix column of the pixel itself.
template <int MODE, bool GLOBAL>
__device__ __forceinline__ void span_walk(const DiscParams& p, const uint32_t* __restrict__ P, int64_t plane_stride,
                                          int pitch, const int* __restrict__ tab, int prow, int lcx,
                                          unsigned long long (&acc)[ModeTraits<MODE>::RB][ModeTraits<MODE>::NARR]) {
    constexpr int NARR = ModeTraits<MODE>::NARR;
    constexpr int RB = ModeTraits<MODE>::RB;
    using idx_t = typename std::conditional<GLOBAL, int64_t, int>::type;  // shared memory: 32-bit indices
#pragma unroll
    for (int b = 0; b < RB; ++b)
#pragma unroll
        for (int a = 0; a < NARR; ++a) acc[b][a] = 0ull;

    const uint32_t* Pa[NARR];
#pragma unroll
    for (int a = 0; a < NARR; ++a) Pa[a] = P + a * plane_stride;

    const int k = p.k;
    // kernel row i touches DEM row offset dy = c - i
    idx_t o = (idx_t)(prow + p.c) * pitch + lcx;
#pragma unroll 2
    for (int i = 0; i < k; ++i, o -= pitch) {
        const int e = tab[i];
        const int lo = (int)(short)(e & 0xffff);
        const int hi1 = (e >> 16) + 1;
#pragma unroll
        for (int b = 0; b < RB; ++b) {
            const idx_t ob = o + (idx_t)b * pitch;
#pragma unroll
            for (int a = 0; a < NARR; ++a) {
                uint32_t hi_v, lo_v;
                if constexpr (GLOBAL) {
                    hi_v = __ldg(Pa[a] + ob + hi1);
                    lo_v = __ldg(Pa[a] + ob + lo);
                } else {
                    hi_v = Pa[a][ob + hi1];
                    lo_v = Pa[a][ob + lo];
                }
                acc[b][a] += (unsigned long long)(uint32_t)(hi_v - lo_v);
            }
        }
    }
}

// ---- fused kernel --------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(kThreads) disc_fused_kernel(const DiscParams p) {
    constexpr int NARR = ModeTraits<MODE>::NARR;
    constexpr int RB = ModeTraits<MODE>::RB;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int* tab = reinterpret_cast<int*>(smem_raw);
    const int tab_elems = (p.k + 3) & ~3;
    uint32_t* P = reinterpret_cast<uint32_t*>(smem_raw) + tab_elems;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x0 = blockIdx.x * kTW;
    const int y0 = p.out_gy0 + blockIdx.y * p.TH;  // global row of the tile's first output row
    const int R = p.TH + 2 * p.halo;
    const int W = p.haloL + kTW + p.halo;
    const int64_t plane_stride = (int64_t)R * p.pitch;
    const bool aligned = ((p.ld_in & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.dem) & 15) == 0);

    build_span_table(p, tab);
    for (int r = warp; r < R; r += kThreads / 32)
        scan_row<MODE>(p, y0 - p.halo + r, x0 - p.haloL, W, P + (int64_t)r * p.pitch, plane_stride, aligned, lane);
    __syncthreads();

    const int tx = threadIdx.x & (kTW - 1);
    const int yg = threadIdx.x >> 7;  // 0..1
    const int x = x0 + tx;
    const int rows_per_thread = p.TH >> 1;
    const int y_end = p.out_gy0 + p.out_rows;
    for (int bt = 0; bt < rows_per_thread; bt += RB) {
        const int ty = yg * rows_per_thread + bt;  // first tile row of this batch
        if (y0 + ty >= y_end) break;
        unsigned long long acc[RB][NARR];
        span_walk<MODE, false>(p, P, plane_stride, p.pitch, tab, ty + p.halo, tx + p.haloL, acc);
        if (x < p.nx) {
#pragma unroll
            for (int b = 0; b < RB; ++b) {
                const int gy = y0 + ty + b;
                if (gy < y_end) p.out[(int64_t)(gy - p.out_gy0) * p.ld_out + x] = finish<MODE>(p, acc[b], gy, x);
            }
        }
    }
}

// ---- two-pass kernels ------------------------------------------------------------------------------
// pass 1: prefix planes for global rows [prow0, prow0 + nrows) and columns [-haloL, nx + halo)
template <int MODE>
__global__ void __launch_bounds__(kThreads) disc_prefix_kernel(const DiscParams p, int nrows) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    if (row >= nrows) return;
    const bool aligned = ((p.ld_in & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.dem) & 15) == 0);
    const int W = p.haloL + p.nx + p.halo;
    scan_row<MODE>(p, p.prow0 + row, -p.haloL, W, p.planes + (int64_t)row * p.pitch, p.plane_stride, aligned, lane);
}

// pass 2: span walk over the global planes
template <int MODE>
__global__ void __launch_bounds__(kThreads) disc_span_kernel(const DiscParams p) {
    constexpr int NARR = ModeTraits<MODE>::NARR;
    constexpr int RB = ModeTraits<MODE>::RB;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int* tab = reinterpret_cast<int*>(smem_raw);
    build_span_table(p, tab);
    __syncthreads();

    const int tx = threadIdx.x & (kTW - 1);
    const int yg = threadIdx.x >> 7;
    const int x = blockIdx.x * kTW + tx;
    const int xc = x < p.nx ? x : p.nx - 1;  // clamp: keeps every address inside the planes
    const int rows_per_thread = p.TH >> 1;
    const int y0 = p.out_gy0 + blockIdx.y * p.TH;
    const int y_end = p.out_gy0 + p.out_rows;
    for (int bt = 0; bt < rows_per_thread; bt += RB) {
        const int ty = yg * rows_per_thread + bt;
        const int gy0 = y0 + ty;
        if (gy0 >= y_end) break;
        // clamp the batch so that all RB rows stay inside the planes (duplicates are not stored)
        int gyb = gy0;
        if (gyb + RB > y_end) gyb = y_end - RB;
        if (gyb < p.out_gy0) gyb = p.out_gy0;  // out_rows < RB: planes are padded (see host)
        unsigned long long acc[RB][NARR];
        span_walk<MODE, true>(p, p.planes, p.plane_stride, p.pitch, tab, gyb - p.prow0, xc + p.haloL, acc);
        if (x < p.nx) {
#pragma unroll
            for (int b = 0; b < RB; ++b) {
                const int gy = gyb + b;
                if (gy >= gy0 && gy < y_end) p.out[(int64_t)(gy - p.out_gy0) * p.ld_out + x] = finish<MODE>(p, acc[b], gy, x);
            }
        }
    }
}

// ---- host side ---------------------------------------------------------------------------------------
struct DiscPlan {
    DiscParams p;
    int mode;
    bool fused;
    size_t smem;
    int prefix_rows;  // two-pass
    size_t ws_bytes;
};

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

constexpr size_t kSmemBudget = 200 * 1024;  // leave room for a second small CTA / static smem

static int max_rb(int mode) { return mode == TPI_Q ? 8 : 4; }
static int narr_of(int mode) { return mode == TPI_Q ? 1 : (mode == STD_F ? 3 : 2); }

// Geometry that does not depend on the data range (used by the workspace query too).
static int plan_geometry(const topo_view* v, int size, int narr, int rb, DiscPlan& pl) {
    DiscParams& p = pl.p;
    p.nx = v->nx, p.gny = v->gny, p.in_gy0 = v->in_gy0, p.in_rows = v->in_rows;
    p.out_gy0 = v->out_gy0, p.out_rows = v->out_rows;
    p.k = size, p.c = (size - 1) / 2, p.mid = size / 2, p.square = size < 5;
    p.halo = size / 2;
    p.haloL = (p.halo + 3) & ~3;
    p.excl = p.c - p.mid;
    const int tab_bytes = ((size + 3) & ~3) * 4;
    // fused: pick the tallest tile that fits
    pl.fused = false;
    for (int th : {32, 16}) {
        if (th / 2 < rb) continue;
        const int R = th + 2 * p.halo;
        const int W = p.haloL + kTW + p.halo;
        const int pitch = ((W + 3) & ~3) + 4;
        const size_t bytes = (size_t)tab_bytes + (size_t)R * pitch * 4 * narr;
        if (bytes <= kSmemBudget) {
            pl.fused = true;
            p.TH = th, p.pitch = pitch;
            pl.smem = bytes;
            pl.ws_bytes = 0;
            pl.prefix_rows = 0;
            return 0;
        }
    }
    // two-pass
    p.TH = 32;
    const int W = p.haloL + p.nx + p.halo;
    p.pitch = ((W + 3) & ~3) + 4;
    p.prow0 = p.out_gy0 - p.halo;
    int rows = p.out_rows + 2 * p.halo;
    if (p.out_rows < rb) rows += rb - p.out_rows;  // batch clamp may read up to RB rows from out_gy0
    pl.prefix_rows = rows;
    p.plane_stride = (int64_t)rows * p.pitch;
    pl.ws_bytes = (size_t)p.plane_stride * 4 * narr;
    pl.smem = tab_bytes;
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

// Fill the data-dependent constants.  what: 0 = TPI, 1 = STD.
static int plan_disc(const topo_view* v, int size, int what, int all_integer, double zmin, double zmax, DiscPlan& pl) {
    TOPO_CHECK(size >= 2 && size <= kMaxSize, "kernel size %d outside [2, %d]", size, kMaxSize);
    TOPO_CHECK(isfinite(zmin) && isfinite(zmax) && zmin <= zmax, "DEM range is not finite");
    const double n = (double)disc_count(size);
    const double span = (double)size;  // longest run of a kernel row
    // integer-part range, including the zero padding value
    const double tlo = fmin(0.0, trunc(zmin)), thi = fmax(0.0, trunc(zmax));
    const double trange = thi - tlo;
    int mode;
    DiscParams& p = pl.p;
    memset(&p, 0, sizeof(p));
    if (what == 0) {
        const double c0 = fmin(0.0, floor(zmin));
        const double range = fmax(0.0, zmax) - c0 + 1.0;
        int S = ilog2_floor(4294967295.0 / (span * range));
        if (S > 20) S = 20;
        if (S >= 10) {
            mode = TPI_Q;
            p.scale = (float)ldexp(1.0, S);
            p.c0i = (int)ldexp(c0, S);
            p.inv_scale = ldexp(1.0, -S);
        } else {
            mode = TPI_X;
        }
    } else {
        mode = all_integer ? STD_I : STD_F;
    }
    if (mode != TPI_Q) {
        TOPO_CHECK(span * (trange + 1.0) < 4294967295.0, "size %d x DEM range %.0f overflows the 32-bit span sums",
                   size, trange);
        p.tmin = (int)tlo;
        p.cmid = (int)(tlo + floor(trange / 2.0));
        if (mode == STD_I || mode == STD_F) {
            const double half = floor(trange / 2.0) + 1.0;
            TOPO_CHECK(span * half * half < 4294967295.0,
                       "size %d x (DEM range %.0f)^2 overflows the 32-bit span sums of squares", size, trange);
        }
        int Sf = 30 - (ilog2_floor(span) + 1);
        if (Sf > 23) Sf = 23;
        p.fscale = (float)ldexp(1.0, Sf);
        p.inv_fscale = ldexp(1.0, -Sf);
    }
    pl.mode = mode;
    if (plan_geometry(v, size, narr_of(mode), max_rb(mode), pl)) return -1;
    p.n = n;
    p.n_ll = (long long)n;
    p.inv_n_nm1 = 1.0 / (n * (n - 1.0));
    p.exact64 = (n * (floor(trange / 2.0) + 1.0)) < 3.0e9;
    p.inv_nm1 = 1.0 / (n - 1.0);
    return 0;
}

static const char* const kFusedName[4] = {"disc_fused<TPI_Q>", "disc_fused<TPI_X>", "disc_fused<STD_I>", "disc_fused<STD_F>"};
static const char* const kPrefixName[4] = {"disc_prefix<TPI_Q>", "disc_prefix<TPI_X>", "disc_prefix<STD_I>", "disc_prefix<STD_F>"};
static const char* const kSpanName[4] = {"disc_span<TPI_Q>", "disc_span<TPI_X>", "disc_span<STD_I>", "disc_span<STD_F>"};

template <int MODE>
static int launch_disc(const DiscPlan& pl, cudaStream_t s) {
    const DiscParams& p = pl.p;
    dim3 grid(ceil_div(p.nx, kTW), ceil_div(p.out_rows, p.TH));
    if (pl.fused) {
        static bool attr_set[64] = {false};
        int dev = 0;
        TOPO_CUDA(cudaGetDevice(&dev));
        if (dev < 64 && !attr_set[dev]) {
            TOPO_CUDA(cudaFuncSetAttribute(disc_fused_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)(227 * 1024)));
            attr_set[dev] = true;
        }
        TOPO_LAUNCH(kFusedName[MODE], s, disc_fused_kernel<MODE><<<grid, kThreads, pl.smem, s>>>(p));
    } else {
        const int warps = kThreads / 32;
        TOPO_LAUNCH(kPrefixName[MODE], s,
                    disc_prefix_kernel<MODE><<<ceil_div(pl.prefix_rows, warps), kThreads, 0, s>>>(p, pl.prefix_rows));
        TOPO_LAUNCH(kSpanName[MODE], s, disc_span_kernel<MODE><<<grid, kThreads, pl.smem, s>>>(p));
    }
    return 0;
}

static int run_disc(const float* dem, int64_t ld_in, float* out, int64_t ld_out, const topo_view* v, int size,
                    int what, int all_integer, double zmin, double zmax, void* ws, size_t ws_bytes, void* stream) {
    TOPO_CHECK(dem && out, "null pointer");
    if (validate_view(v)) return -1;
    TOPO_CHECK(ld_in >= v->nx && ld_out >= v->nx, "row pitch smaller than nx");
    if (v->out_rows == 0) return 0;
    if (size == 1) {
        // kernel sum - 1 == 0: the reference divides by zero and returns NaN everywhere
        return topo_fill_f32(out, v->out_rows, v->nx, ld_out, NAN, stream);
    }
    DiscPlan pl;
    if (plan_disc(v, size, what, all_integer, zmin, zmax, pl)) return -1;
    if (check_band(v, pl.p.halo)) return -1;
    pl.p.dem = dem, pl.p.out = out, pl.p.ld_in = ld_in, pl.p.ld_out = ld_out;
    if (!pl.fused) {
        TOPO_CHECK(ws != nullptr && ws_bytes >= pl.ws_bytes, "workspace too small: need %zu bytes, got %zu",
                   pl.ws_bytes, ws_bytes);
        TOPO_CHECK((reinterpret_cast<uintptr_t>(ws) & 15) == 0, "workspace must be 16-byte aligned");
        pl.p.planes = (uint32_t*)ws;
    }
    cudaStream_t s = (cudaStream_t)stream;
    switch (pl.mode) {
        case TPI_Q: return launch_disc<TPI_Q>(pl, s);
        case TPI_X: return launch_disc<TPI_X>(pl, s);
        case STD_I: return launch_disc<STD_I>(pl, s);
        default: return launch_disc<STD_F>(pl, s);
    }
}

}  // namespace topo

using namespace topo;

extern "C" {

size_t topo_disc_workspace_bytes(const topo_view* v, int size, int what) {
    if (!v || size < 2 || size > kMaxSize) return 0;
    // worst case over the modes `what` can select (the mode depends on the data range)
    size_t worst = 0;
    const int modes_tpi[2] = {TPI_Q, TPI_X}, modes_std[2] = {STD_I, STD_F};
    for (int m = 0; m < 2; ++m) {
        const int mode = what == 0 ? modes_tpi[m] : modes_std[m];
        DiscPlan pl;
        memset(&pl, 0, sizeof(pl));
        plan_geometry(v, size, narr_of(mode), max_rb(mode), pl);
        if (pl.ws_bytes > worst) worst = pl.ws_bytes;
    }
    return worst;
}

int topo_tpi_f32(const float* dem, int64_t ld_in, float* out, int64_t ld_out, const topo_view* v, int size,
                 double zmin, double zmax, void* ws, size_t ws_bytes, void* stream) {
    return run_disc(dem, ld_in, out, ld_out, v, size, 0, 0, zmin, zmax, ws, ws_bytes, stream);
}

int topo_std_f32(const float* dem, int64_t ld_in, float* out, int64_t ld_out, const topo_view* v, int size,
                 int all_integer, double zmin, double zmax, void* ws, size_t ws_bytes, void* stream) {
    return run_disc(dem, ld_in, out, ld_out, v, size, 1, all_integer, zmin, zmax, ws, ws_bytes, stream);
}

}  // extern "C"
