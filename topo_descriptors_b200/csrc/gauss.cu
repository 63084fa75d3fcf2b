// Gaussian smoothing, gradient / slope / aspect, Sobel.
// Reference: topo.py:62-80 (dem), 597-644 (gradient), 658-685 (sobel), 688-712 (_normalize_dxy);
// third-party semantics restated from scipy.ndimage (gaussian_filter -> correlate1d per axis, axis 0
// first, float64 line buffers and accumulators, output rounded to the input dtype after each axis,
// mode="reflect") and numpy.gradient (central differences, first-order one-sided at the edges).
//
// Separable passes, register-blocked: every thread owns K consecutive outputs ALONG the filter axis and
// slides over K + 2*lw inputs, keeping a rotating window of K weights in registers, so each loaded
// sample feeds K float64 FMAs.  Axis 0: lanes = columns (coalesced global loads, neighbouring row
// groups share lines through L1).  Axis 1: the row segment + halo is staged in shared memory (reflect
// applied while staging) and lanes = rows with an odd pitch (conflict-free), outputs are transposed
// back through shared memory for coalesced stores.
#include <math.h>
#include <string.h>

#include <type_traits>

#include "common.cuh"
#include "tma.cuh"

namespace topo {

// gauss_fft.cu: overlap-save float64 FFT pass along the contiguous axis (wide radii)
int fft_conv_length(int lw);
size_t fft_conv_table_bytes(int lw);
int fft_conv_rows(const float* in, int64_t ld_in, float* out, int64_t ld_out, int n_lines, int n_glob, int in0, int in_len,
                  int out0, int out_len, const double* w, int lw, void* tables, cudaStream_t s);

constexpr int kFftMinRadius = 40;  // from here the FFT pass (~1.9 ms per axis at 16384^2, any radius) beats 2*lw+1 float64 taps per pixel
constexpr int kK = 16;  // outputs per thread along the filter axis (micro-benchmark: 41 DFMA/clk/SM vs 28 at K = 8)
constexpr int kAxis1SmemMaxRadius = 64;  // wider axis-1 filters go through a transpose

struct GaussParams {
    const float* in;
    float* out;
    int64_t ld_in, ld_out;
    int nx, gny;
    int in_gy0, in_rows;    // rows present in `in`
    int out_gy0, out_rows;  // rows to produce
    const double* w;        // device: w[0] = centre ... w[lw]
    int lw;
};

// Steps a thread walks: K + 2*lw inputs, rounded up to a multiple of K.
__host__ __device__ __forceinline__ int gauss_steps(int lw) { return ((kK + 2 * lw + kK - 1) / kK) * kK; }

// Full kernel laid out by step index n (input offset t = n - lw): wfull[n] = w[|n - lw|], zero past 2*lw.
// The inner loops then read it linearly (immediate offsets, no index arithmetic).
__device__ __forceinline__ void stage_weights(const GaussParams& p, double* wfull, int tid, int nthreads) {
    const int nsteps = gauss_steps(p.lw);
    for (int n = tid; n < nsteps; n += nthreads) {
        int a = n - p.lw;
        a = a < 0 ? -a : a;
        wfull[n] = (a <= p.lw) ? p.w[a] : 0.0;
    }
}

// ---- axis 0 (along y) -----------------------------------------------------------------------------
// block (32, 8): 64 columns (each thread owns columns x and x + 32, which share the weight window) x 8 row
// groups of K rows: one LDS.64 + two LDG + two F2F feed 2*K DFMA.  Interior row groups (the whole input
// window lies inside the image and the band) walk a pointer down the column; only groups that touch the
// global top or bottom edge pay for the reflect index arithmetic.
constexpr int kA0Cols = 64;

// NANSAFE: the DEM holds non-finite values; taps outside an output's own window (zero weight) must not touch
// it (0 * NaN = NaN would spread NaN further than scipy's +-lw).
template <bool NANSAFE>
__global__ void __launch_bounds__(256, 2) gauss_axis0_kernel(const GaussParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* wfull = reinterpret_cast<double*>(smem_raw);
    __shared__ long long rowoff[8][kK];  // per warp: element offset of the input row of each step of a group
    stage_weights(p, wfull, threadIdx.x + threadIdx.y * 32, 256);
    __syncthreads();

    const int x = blockIdx.x * kA0Cols + threadIdx.x;
    const int gy0 = p.out_gy0 + (blockIdx.y * 8 + threadIdx.y) * kK;  // first output row of this thread
    if (gy0 >= p.out_gy0 + p.out_rows) return;                        // warp-uniform
    const int xa = x < p.nx ? x : p.nx - 1;
    const int xb = x + 32 < p.nx ? x + 32 : p.nx - 1;
    const int lw = p.lw;

    double acc0[kK], acc1[kK], wr[kK];
#pragma unroll
    for (int k = 0; k < kK; ++k) acc0[k] = 0.0, acc1[k] = 0.0, wr[k] = 0.0;

    const int in_end = p.in_gy0 + p.in_rows;
    const int nsteps = gauss_steps(lw);
    const int first = gy0 - lw;
    const double* wp = wfull;
    long long* ro = rowoff[threadIdx.y];
    for (int n0 = 0; n0 < nsteps; n0 += kK, wp += kK) {
        // lanes 0..K-1 resolve the input rows of this group (reflect at the global edges; rows outside the
        // band can only carry zero weight or feed outputs that are not stored: clamp them into the band)
        __syncwarp();
        if (threadIdx.x < kK) {
            int g = reflect_index(first + n0 + (int)threadIdx.x, p.gny);
            g = g < p.in_gy0 ? p.in_gy0 : (g >= in_end ? in_end - 1 : g);
            ro[threadIdx.x] = (long long)(g - p.in_gy0) * p.ld_in;
        }
        __syncwarp();
#pragma unroll
        for (int s = 0; s < kK; ++s) {
#pragma unroll
            for (int k = kK - 1; k > 0; --k) wr[k] = wr[k - 1];
            wr[0] = wp[s];
            const float* row = p.in + ro[s];
            const double da = (double)__ldg(row + xa), db = (double)__ldg(row + xb);
#pragma unroll
            for (int k = 0; k < kK; ++k) {
                if (!NANSAFE || (unsigned)(n0 + s - k) <= (unsigned)(2 * lw)) {
                    acc0[k] = fma(wr[k], da, acc0[k]);
                    acc1[k] = fma(wr[k], db, acc1[k]);
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < kK; ++k) {
        const int gy = gy0 + k;
        if (gy < p.out_gy0 + p.out_rows) {
            float* o = p.out + (int64_t)(gy - p.out_gy0) * p.ld_out;
            if (x < p.nx) o[x] = (float)acc0[k];
            if (x + 32 < p.nx) o[x + 32] = (float)acc1[k];
        }
    }
}

// 64 x 64 tiled transpose (rows x cols -> cols x rows); lets the wide-radius passes along y run on contiguous lines.
// 128-bit loads along the rows, 128-bit stores along the transposed rows (each thread gathers a column quad from the
// padded tile: conflict-free), edge tiles and unaligned rasters element by element.
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ in, int64_t ld_in,
                                                        float* __restrict__ out, int64_t ld_out, int rows, int cols) {
    __shared__ float t[64][65];
    const int c0 = blockIdx.x * 64, r0 = blockIdx.y * 64;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const bool full = c0 + 64 <= cols && r0 + 64 <= rows && (ld_in & 3) == 0 && (ld_out & 3) == 0 &&
                      (reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    if (full) {
        const int q = tid & 15, rr = tid >> 4;  // 16 quads per tile row, 16 rows per sweep
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = rr + 16 * i;
            const float4 v = ldg4(in + (int64_t)(r0 + r) * ld_in + c0 + 4 * q);
            t[r][4 * q] = v.x, t[r][4 * q + 1] = v.y, t[r][4 * q + 2] = v.z, t[r][4 * q + 3] = v.w;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = rr + 16 * i;  // output row = input column
            const float4 v = make_float4(t[4 * q][c], t[4 * q + 1][c], t[4 * q + 2][c], t[4 * q + 3][c]);
            *reinterpret_cast<float4*>(out + (int64_t)(c0 + c) * ld_out + r0 + 4 * q) = v;
        }
        return;
    }
    for (int i = tid; i < 64 * 64; i += 256) {
        const int r = i >> 6, c = i & 63;
        if (r0 + r < rows && c0 + c < cols) t[r][c] = __ldg(in + (int64_t)(r0 + r) * ld_in + c0 + c);
    }
    __syncthreads();
    for (int i = tid; i < 64 * 64; i += 256) {
        const int c = i >> 6, r = i & 63;
        if (r0 + r < rows && c0 + c < cols) out[(int64_t)(c0 + c) * ld_out + r0 + r] = t[r][c];
    }
}

// ---- axis 1 (along x), radius <= kAxis1SmemMaxRadius ---------------------------------------------------
// block 256 = 8 warps.  Tile: 32 rows (lanes) x 8*K output columns; warp w owns columns [w*K, w*K+K).
// K = 8 for small radii (the walk visits K + 2*lw taps rounded up to K: less padding), 16 otherwise.
template <int K>
__host__ __device__ __forceinline__ int axis1_steps(int lw) { return ((K + 2 * lw + K - 1) / K) * K; }

template <int K, bool NANSAFE>
__global__ void __launch_bounds__(256) gauss_axis1_kernel(const GaussParams p, int pitch) {
    constexpr int COLS = 8 * K;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* wfull = reinterpret_cast<double*>(smem_raw);
    const int nsteps = axis1_steps<K>(p.lw);
    float* tile = reinterpret_cast<float*>(wfull + nsteps);  // [32][pitch]
    float* otile = tile + (size_t)32 * pitch;                // [32][COLS + 1]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lw = p.lw;
    for (int n = threadIdx.x; n < nsteps; n += 256) {
        int a = n - lw;
        a = a < 0 ? -a : a;
        wfull[n] = (a <= lw) ? p.w[a] : 0.0;
    }

    const int x0 = blockIdx.x * COLS;
    const int r0 = blockIdx.y * 32;  // first row (relative to the output band) of this tile
    const int span = COLS + nsteps - K;  // columns a thread group can touch: c0 + n, n < nsteps
    // stage: warp per row, lanes along x, reflect at the global left/right edges
    for (int r = warp; r < 32; r += 8) {
        const int row = r0 + r;
        float* dst = tile + (size_t)r * pitch;
        if (row < p.out_rows) {
            const float* src = p.in + (int64_t)(p.out_gy0 + row - p.in_gy0) * p.ld_in;
            for (int c = lane; c < span; c += 32) dst[c] = __ldg(src + reflect_index(x0 - lw + c, p.nx));
        } else {
            for (int c = lane; c < span; c += 32) dst[c] = 0.f;
        }
    }
    __syncthreads();

    {
        const int c0 = warp * K;  // first output column (tile-relative) of this thread
        const float* tp = tile + (size_t)lane * pitch + c0;  // input x0 + c0 + (n - lw) sits at tile column c0 + n
        const double* wp = wfull;
        double acc[K], wr[K];
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = 0.0, wr[k] = 0.0;
        for (int n0 = 0; n0 < nsteps; n0 += K, wp += K, tp += K) {
#pragma unroll
            for (int s = 0; s < K; ++s) {
#pragma unroll
                for (int k = K - 1; k > 0; --k) wr[k] = wr[k - 1];
                wr[0] = wp[s];
                const double dv = (double)tp[s];
#pragma unroll
                for (int k = 0; k < K; ++k)
                    if (!NANSAFE || (unsigned)(n0 + s - k) <= (unsigned)(2 * lw)) acc[k] = fma(wr[k], dv, acc[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < K; ++k) otile[lane * (COLS + 1) + c0 + k] = (float)acc[k];
    }
    __syncthreads();
    for (int r = warp; r < 32; r += 8) {
        const int row = r0 + r;
        if (row >= p.out_rows) break;
        float* dst = p.out + (int64_t)row * p.ld_out;
        for (int c = lane; c < COLS; c += 32)
            if (x0 + c < p.nx) dst[x0 + c] = otile[r * (COLS + 1) + c];
    }
}

template <int K>
static int launch_axis1(const GaussParams& p, int nan_safe, cudaStream_t s) {
    constexpr int COLS = 8 * K;
    const int nsteps = axis1_steps<K>(p.lw);
    int pitch = COLS + nsteps - K;
    pitch |= 1;  // odd pitch: lanes (rows) hit distinct banks
    const size_t smem = (size_t)nsteps * sizeof(double) + ((size_t)32 * pitch + (size_t)32 * (COLS + 1)) * sizeof(float);
    static bool attr_set[64] = {false};
    int dev = 0;
    TOPO_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && !attr_set[dev]) {
        TOPO_CUDA(cudaFuncSetAttribute(gauss_axis1_kernel<K, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        TOPO_CUDA(cudaFuncSetAttribute(gauss_axis1_kernel<K, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set[dev] = true;
    }
    dim3 grid(ceil_div(p.nx, COLS), ceil_div(p.out_rows, 32));
    if (nan_safe)
        TOPO_LAUNCH("gauss_axis1<nansafe>", s, gauss_axis1_kernel<K, true><<<grid, 256, smem, s>>>(p, pitch));
    else
        TOPO_LAUNCH("gauss_axis1", s, gauss_axis1_kernel<K, false><<<grid, 256, smem, s>>>(p, pitch));
    return 0;
}

// ---- derivative / slope / aspect epilogue ------------------------------------------------------------
struct GradParams {
    const float* gx;  // differentiated along x
    const float* gy;  // differentiated along y
    float *dx, *dy, *slope, *aspect;
    int64_t ld_in, ld_out;
    int nx, gny, in_gy0, in_rows, out_gy0, out_rows;
    const double* res_x;
    const double* res_y;
    const float* res_xf;  // optional float32 copies (exactly equal to the float64 arrays)
    const float* res_yf;
    int res_x_2d, res_y_2d, normalize;
};

// float32 value / float64 resolution, rounded to float32.  When the resolution is exactly representable in
// float32 (25.0, 30.0, float32-derived UTM spacings ...) the IEEE float32 division gives the same correctly
// rounded quotient without the float64 divide.
__device__ __forceinline__ float div_by_res(float v, double r) {
    const float rf = (float)r;
    if ((double)rf == r) return __fdiv_rn(v, rf);
    return (float)((double)v / r);
}

// slope / aspect of one pixel in the reference's float32 operation order (topo.py:639-642)
__device__ __forceinline__ void slope_aspect(float dx, float dy, float& slope, float& aspect) {
    // np.arctan(np.sqrt(dx**2 + dy**2)) * (180 / np.pi)
    const float h2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
    slope = __fmul_rn(atanf(sqrtf(h2)), 57.29577951308232f);
    // (180 + np.degrees(np.arctan2(dx, dy))) % 360; np.degrees on float32 multiplies by 180.0f / NPY_PIf =
    // 57.2957763671875f (one ulp below float32(180/pi) used for the slope)
    const float deg = __fmul_rn(atan2f(dx, dy), 57.2957763671875f);
    float a = __fadd_rn(180.0f, deg);
    if (a >= 360.0f) a -= 360.0f;  // a is in [0, 360]: Python's % 360 only maps 360 -> 0; NaN stays NaN
    aspect = a;
}

__device__ __forceinline__ void finish_gradient(const GradParams& p, float dx, float dy, int gy, int x) {
    if (p.normalize) {
        // `dx /= res` with a float64 resolution array: float64 division, rounded back to float32
        const double rx = p.res_x_2d ? p.res_x[(int64_t)gy * p.nx + x] : p.res_x[x];
        const double ry = p.res_y_2d ? p.res_y[(int64_t)gy * p.nx + x] : p.res_y[gy];
        dx = div_by_res(dx, rx);
        dy = div_by_res(dy, ry);
    }
    const int64_t o = (int64_t)(gy - p.out_gy0) * p.ld_out + x;
    p.dx[o] = dx;
    p.dy[o] = dy;
    if (p.slope) {
        float sl, as;
        slope_aspect(dx, dy, sl, as);
        p.slope[o] = sl;
        p.aspect[o] = as;
    }
}

// generic per-pixel path (image borders, unaligned rasters, float64-only resolutions)
template <bool SOBEL>
__device__ __forceinline__ void gradient_pixel(const GradParams& p, int gy, int x) {
    float dx, dy;
    if (SOBEL) {
        const int xm = reflect_index(x - 1, p.nx), xp = reflect_index(x + 1, p.nx);
        const float* r0 = p.gx + (int64_t)(reflect_index(gy - 1, p.gny) - p.in_gy0) * p.ld_in;
        const float* r1 = p.gx + (int64_t)(gy - p.in_gy0) * p.ld_in;
        const float* r2 = p.gx + (int64_t)(reflect_index(gy + 1, p.gny) - p.in_gy0) * p.ld_in;
        const double a = __ldg(r0 + xm), b = __ldg(r0 + x), c = __ldg(r0 + xp);
        const double d = __ldg(r1 + xm), f = __ldg(r1 + xp);
        const double g = __ldg(r2 + xm), h = __ldg(r2 + x), i = __ldg(r2 + xp);
        // exact in float64 (<= 6 terms of 24-bit values), one rounding to float32 like ndimage
        dx = (float)(((c + 2.0 * f + i) - (a + 2.0 * d + g)) * 0.125);
        dy = (float)(((g + 2.0 * h + i) - (a + 2.0 * b + c)) * 0.125);
    } else {
        const float* rx = p.gx + (int64_t)(gy - p.in_gy0) * p.ld_in;
        if (x == 0)
            dx = __fsub_rn(__ldg(rx + 1), __ldg(rx));
        else if (x == p.nx - 1)
            dx = __fsub_rn(__ldg(rx + x), __ldg(rx + x - 1));
        else
            dx = __fmul_rn(__fsub_rn(__ldg(rx + x + 1), __ldg(rx + x - 1)), 0.5f);
        const float* cy = p.gy + (int64_t)(gy - p.in_gy0) * p.ld_in + x;
        if (gy == 0)
            dy = __fsub_rn(__ldg(cy + p.ld_in), __ldg(cy));
        else if (gy == p.gny - 1)
            dy = __fsub_rn(__ldg(cy), __ldg(cy - p.ld_in));
        else
            dy = __fmul_rn(__fsub_rn(__ldg(cy + p.ld_in), __ldg(cy - p.ld_in)), 0.5f);
    }
    finish_gradient(p, dx, dy, gy, x);
}

// ---- fused small-radius gradient: Gaussian axis 0 -> float32 -> axis 1 -> float32 -> np.gradient -> slope / aspect ----
// A persistent CTA walks 30 x 128 output tiles and produces the four outputs from ONE read of the DEM tile + halo
// (20 B/px of HBM traffic instead of the 36 B/px of the three-kernel route).  The raw tile arrives by TMA
// (cp.async.bulk.tensor.2d + mbarrier; with two buffers the next tile is in flight while this one is computed; tiles
// that touch an image edge are staged by hand with the reflect rule).  The raw tile (float32), the axis-0 result
// (rounded to float32 like scipy's intermediate array, kept as float64 so the second pass needs no conversion) and the
// smoothed tile stay in shared memory.  Every tap is a float64 FMA in the same order as gauss_axis0 / gauss_axis1
// (input index ascending), so the results are bit identical to the three-kernel route.  A thread owns 8 consecutive
// outputs along the filter axis; the half kernel sits in (uniform) registers (template radius LWT >= lw, weights
// beyond lw are zero: fma(0, x, acc) = acc) and all tap indices are compile-time, so exactly 8 (2 LWT + 1) DFMA are
// issued per 8 + 2 LWT samples.
constexpr int kFuTH = 30, kFuTW = 128;       // output tile
constexpr int kFuGR = kFuTH + 2;             // smoothed rows held per tile (one more on each side for the differences)
constexpr int kFuGC = 136;                   // smoothed columns computed (17 groups of 8): column c <-> x0 - 4 + c
constexpr int kFuGP = 140;                   // pitch of the smoothed tile (multiple of 4: aligned 128-bit reads)
constexpr int kFusedMaxRadius = 21;

template <int LWT>
struct FusedShape {
    static constexpr int NBUF = LWT <= 5 ? 2 : 1;              // raw-tile buffers (2: the next tile's copy overlaps the whole tile;
                                                                // 1: it is issued after the axis-0 pass) -- sized for 2 CTAs per SM
    using AT = typename std::conditional<(LWT <= 13), double, float>::type;  // axis-0 result: float64 saves the second pass its
                                                                             // conversions, float32 halves the buffer
    static constexpr int HL = (4 + LWT + 3) & ~3;               // raw column C <-> global column x0 - HL + C (16-byte aligned box)
    static constexpr int OFF = HL - 4 - LWT;                    // smoothed column c, tap n reads raw / axis-0 column c + n + OFF
    static constexpr int RAW_ROWS = kFuGR + 2 * LWT;            // raw row R <-> global row y0 - 1 - LWT + R
    static constexpr int NEED_COLS = kFuTW + 2 + 2 * LWT + OFF + 3;  // last raw column a needed output reads, + 1
    static constexpr int RAW_COLS = (kFuGC + 2 * LWT + OFF + 3) & ~3;
    static constexpr int A_PITCH = RAW_COLS | 1;                // elements; lanes = rows in the axis-1 pass: odd pitch
    static constexpr size_t RAW_BYTES = (size_t)RAW_ROWS * RAW_COLS * sizeof(float);
    static constexpr size_t RAW_STRIDE = (RAW_BYTES + 127) & ~(size_t)127;  // TMA destinations are 128-byte aligned
    static constexpr size_t OFF_A = NBUF * RAW_STRIDE;
    static constexpr size_t OFF_G = (OFF_A + (size_t)kFuGR * A_PITCH * sizeof(AT) + 15) & ~(size_t)15;
    static constexpr size_t OFF_BAR = (OFF_G + (size_t)kFuGR * kFuGP * sizeof(float) + 15) & ~(size_t)15;
    static constexpr size_t BYTES = OFF_BAR + 2 * sizeof(uint64_t);
};

// acc[k] += w[|s - k - LWT|] * x for the outputs k = 0..7 that sample s (0 .. 7 + 2 LWT) reaches
template <int LWT, int S>
__device__ __forceinline__ void fused_taps(double (&acc)[8], const double (&w)[LWT + 1], double x) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int n = S - k;
        if (n >= 0 && n <= 2 * LWT) acc[k] = fma(w[n < LWT ? LWT - n : n - LWT], x, acc[k]);
    }
}

template <int LWT, int S = 0>
struct FusedWalk {
    template <class Load>
    __device__ __forceinline__ static void run(double (&acc)[8], const double (&w)[LWT + 1], Load load) {
        fused_taps<LWT, S>(acc, w, load(S));
        if constexpr (S + 1 < 8 + 2 * LWT) FusedWalk<LWT, S + 1>::run(acc, w, load);
    }
};

struct FusedParams {
    GradParams g;
    const double* w;
    int lw, vec_ok, use_tma, tiles_x, tiles_y;
};

template <int LWT>
__global__ void __launch_bounds__(256) gauss_grad_fused_kernel(const __grid_constant__ CUtensorMap tmap, const FusedParams q) {
    using SH = FusedShape<LWT>;
    const GradParams& p = q.g;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    using AT = typename SH::AT;
    AT* A = reinterpret_cast<AT*>(smem_raw + SH::OFF_A);
    float* G = reinterpret_cast<float*>(smem_raw + SH::OFF_G);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + SH::OFF_BAR);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int in_end = p.in_gy0 + p.in_rows;
    const int y_end = p.out_gy0 + p.out_rows;
    const int ntiles = q.tiles_x * q.tiles_y;

    double w[LWT + 1];
#pragma unroll
    for (int i = 0; i <= LWT; ++i) w[i] = i <= q.lw ? __ldg(q.w + i) : 0.0;

    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();

    // a tile whose box lies inside the image and the band is copied by TMA; the others are staged by hand (reflect)
    auto tile_origin = [&](int t, int& x0, int& y0) {
        const int by = t / q.tiles_x;
        x0 = (t - by * q.tiles_x) * kFuTW;
        y0 = p.out_gy0 + by * kFuTH;
    };
    auto by_tma = [&](int t) {
        int x0, y0;
        tile_origin(t, x0, y0);
        const int r0 = y0 - 1 - LWT;
        return q.use_tma && x0 - SH::HL >= 0 && x0 - SH::HL + SH::NEED_COLS <= p.nx && r0 >= 0 && r0 >= p.in_gy0 &&
               r0 + SH::RAW_ROWS <= p.gny && r0 + SH::RAW_ROWS <= in_end;
    };
    auto issue = [&](int t, int buf) {  // thread 0 only
        int x0, y0;
        tile_origin(t, x0, y0);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&bar[buf], (uint32_t)SH::RAW_BYTES);
        tma_load_2d(smem_raw + buf * SH::RAW_STRIDE, &tmap, x0 - SH::HL, y0 - 1 - LWT - p.in_gy0, &bar[buf]);
    };

    uint32_t phases = 0u;  // bit b: parity the next completion of bar[b] will have
    const int t0 = blockIdx.x, stride = gridDim.x;
    if (t0 < ntiles && by_tma(t0) && tid == 0) issue(t0, 0);
    for (int it = 0, t = t0; t < ntiles; ++it, t += stride) {
        const int buf = SH::NBUF == 2 ? (it & 1) : 0;
        float* raw = reinterpret_cast<float*>(smem_raw + buf * SH::RAW_STRIDE);
        int x0, y0;
        tile_origin(t, x0, y0);
        if (SH::NBUF == 2) {  // next tile in flight while this one is computed (its buffer was last read two syncs ago)
            const int nt = t + stride;
            if (nt < ntiles && by_tma(nt) && tid == 0) issue(nt, buf ^ 1);
        }
        if (by_tma(t)) {
            mbar_wait(&bar[buf], (phases >> buf) & 1u);
            phases ^= 1u << buf;
        } else {
            for (int R = warp; R < SH::RAW_ROWS; R += 8) {
                int g = reflect_index(y0 - 1 - LWT + R, p.gny);
                g = g < p.in_gy0 ? p.in_gy0 : (g >= in_end ? in_end - 1 : g);  // outside the band: rows no stored output reads
                const float* src = p.gx + (int64_t)(g - p.in_gy0) * p.ld_in;
                float* dst = raw + R * SH::RAW_COLS;
                for (int C = lane; C < SH::RAW_COLS; C += 32) dst[C] = __ldg(src + reflect_index(x0 - SH::HL + C, p.nx));
            }
            __syncthreads();
        }

        // ---- axis 0: A[r][C] = float32(sum_n w[n] raw[r + n][C]); item = (row group of 8, column), lanes = columns
        for (int item = tid; item < (kFuGR / 8) * SH::RAW_COLS; item += 256) {
            const int g = item / SH::RAW_COLS, C = item - g * SH::RAW_COLS;
            const float* col = raw + (g * 8) * SH::RAW_COLS + C;
            double acc[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = 0.0;
            FusedWalk<LWT>::run(acc, w, [&](int s) { return (double)col[s * SH::RAW_COLS]; });
#pragma unroll
            for (int k = 0; k < 8; ++k) A[(g * 8 + k) * SH::A_PITCH + C] = (AT)(float)acc[k];
        }
        __syncthreads();
        if (SH::NBUF == 1) {  // single buffer: the raw tile is free now, start the next copy under the remaining phases
            const int nt = t + stride;
            if (nt < ntiles && by_tma(nt) && tid == 0) issue(nt, 0);
        }

        // ---- axis 1: G[r][c] = float32(sum_n w[n] A[r][c + n + OFF]); item = (column group of 8, row), lanes = rows
        for (int item = tid; item < (kFuGC / 8) * kFuGR; item += 256) {
            const int gx = item / kFuGR, r = item - gx * kFuGR;
            const AT* row = A + r * SH::A_PITCH + gx * 8 + SH::OFF;
            double acc[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = 0.0;
            FusedWalk<LWT>::run(acc, w, [&](int s) { return (double)row[s]; });
#pragma unroll
            for (int k = 0; k < 8; ++k) G[r * kFuGP + gx * 8 + k] = (float)acc[k];
        }
        __syncthreads();

        // ---- np.gradient + resolution + slope / aspect; G[ty + 1][tx + 4] is the smoothed value of output (ty, tx).
        // 4 pixels per thread; interior quads of aligned rasters with float32 resolutions take 128-bit loads / stores.
        for (int i = tid; i < kFuTH * (kFuTW / 4); i += 256) {
            const int ty = i / (kFuTW / 4), tx = (i - ty * (kFuTW / 4)) * 4;
            const int gy = y0 + ty, x = x0 + tx;
            if (gy >= y_end || x >= p.nx) continue;
            const float* c = G + (ty + 1) * kFuGP + tx + 4;
            const bool quad = q.vec_ok && x >= 1 && x + 4 < p.nx && gy >= 1 && gy + 1 < p.gny;
            if (quad) {
                const float4 mid = *reinterpret_cast<const float4*>(c);
                const float4 up = *reinterpret_cast<const float4*>(c - kFuGP), dn = *reinterpret_cast<const float4*>(c + kFuGP);
                const float v[6] = {c[-1], mid.x, mid.y, mid.z, mid.w, c[4]};
                const float u[4] = {up.x, up.y, up.z, up.w}, d[4] = {dn.x, dn.y, dn.z, dn.w};
                float dxv[4], dyv[4], rxv[4], ryv[4], sl[4], as[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    dxv[j] = __fmul_rn(__fsub_rn(v[j + 2], v[j]), 0.5f);
                    dyv[j] = __fmul_rn(__fsub_rn(d[j], u[j]), 0.5f);
                }
                const float4 tx4 = ldg4(p.res_xf + (p.res_x_2d ? (int64_t)gy * p.nx : 0) + x);
                rxv[0] = tx4.x, rxv[1] = tx4.y, rxv[2] = tx4.z, rxv[3] = tx4.w;
                if (p.res_y_2d) {
                    const float4 t4 = ldg4(p.res_yf + (int64_t)gy * p.nx + x);
                    ryv[0] = t4.x, ryv[1] = t4.y, ryv[2] = t4.z, ryv[3] = t4.w;
                } else {
                    ryv[0] = ryv[1] = ryv[2] = ryv[3] = __ldg(p.res_yf + gy);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    dxv[j] = __fdiv_rn(dxv[j], rxv[j]);
                    dyv[j] = __fdiv_rn(dyv[j], ryv[j]);
                }
                const int64_t o = (int64_t)(gy - p.out_gy0) * p.ld_out + x;
                st4_streaming(p.dx + o, make_float4(dxv[0], dxv[1], dxv[2], dxv[3]));
                st4_streaming(p.dy + o, make_float4(dyv[0], dyv[1], dyv[2], dyv[3]));
                if (p.slope) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) slope_aspect(dxv[j], dyv[j], sl[j], as[j]);
                    st4_streaming(p.slope + o, make_float4(sl[0], sl[1], sl[2], sl[3]));
                    st4_streaming(p.aspect + o, make_float4(as[0], as[1], as[2], as[3]));
                }
                continue;
            }
#pragma unroll 1
            for (int j = 0; j < 4; ++j) {
                const int xx = x + j;
                if (xx >= p.nx) break;
                float dx, dy;
                if (xx == 0)
                    dx = __fsub_rn(c[j + 1], c[j]);
                else if (xx == p.nx - 1)
                    dx = __fsub_rn(c[j], c[j - 1]);
                else
                    dx = __fmul_rn(__fsub_rn(c[j + 1], c[j - 1]), 0.5f);
                if (gy == 0)
                    dy = __fsub_rn(c[j + kFuGP], c[j]);
                else if (gy == p.gny - 1)
                    dy = __fsub_rn(c[j], c[j - kFuGP]);
                else
                    dy = __fmul_rn(__fsub_rn(c[j + kFuGP], c[j - kFuGP]), 0.5f);
                finish_gradient(p, dx, dy, gy, xx);
            }
        }
        // (no barrier: the next iteration writes the other raw buffer / A only after its own barriers, and G is not
        // written before the barrier that follows the next axis-0 pass)
        if (SH::NBUF == 1) __syncthreads();  // hand-staged tiles write the single raw buffer right away
    }
}

// block 256 = 64 pixel quads x 4 rows: tile 256 x 4.  Interior tiles of aligned rasters take the vector path
// (128-bit loads and streaming stores, 4 pixels per thread); everything else goes pixel by pixel.
template <bool SOBEL>
__global__ void __launch_bounds__(256) gradient_kernel(const GradParams p, int vec_ok) {
    const int tq = threadIdx.x & 63;
    const int x0 = blockIdx.x * 256;
    const int gy = p.out_gy0 + blockIdx.y * 4 + (threadIdx.x >> 6);
    const int ty0 = p.out_gy0 + blockIdx.y * 4;
    const bool interior = vec_ok && x0 >= 4 && x0 + 256 + 4 <= p.nx && ty0 >= 1 && ty0 + 4 + 1 <= p.gny &&
                          ty0 + 4 <= p.out_gy0 + p.out_rows && ty0 - 1 >= p.in_gy0 && ty0 + 5 <= p.in_gy0 + p.in_rows;
    if (!interior) {
        if (gy >= p.out_gy0 + p.out_rows) return;
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
            const int x = x0 + 4 * tq + j;
            if (x < p.nx) gradient_pixel<SOBEL>(p, gy, x);
        }
        return;
    }
    const int x = x0 + 4 * tq;
    float dxv[4], dyv[4];
    if (SOBEL) {
        const float* r0 = p.gx + (int64_t)(gy - 1 - p.in_gy0) * p.ld_in + x;
        const float* r1 = r0 + p.ld_in;
        const float* r2 = r1 + p.ld_in;
        const float4 m0 = ldg4(r0), m1 = ldg4(r1), m2 = ldg4(r2);
        const float v0[6] = {__ldg(r0 - 1), m0.x, m0.y, m0.z, m0.w, __ldg(r0 + 4)};
        const float v1[6] = {__ldg(r1 - 1), m1.x, m1.y, m1.z, m1.w, __ldg(r1 + 4)};
        const float v2[6] = {__ldg(r2 - 1), m2.x, m2.y, m2.z, m2.w, __ldg(r2 + 4)};
        double cs[6], rd[6];  // column sums (1,2,1) and row differences, exact in float64
#pragma unroll
        for (int t = 0; t < 6; ++t) {
            const double a = (double)v0[t], b = (double)v1[t], c = (double)v2[t];
            cs[t] = (a + c) + 2.0 * b;
            rd[t] = c - a;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            dxv[j] = (float)((cs[j + 2] - cs[j]) * 0.125);
            dyv[j] = (float)(((rd[j] + rd[j + 2]) + 2.0 * rd[j + 1]) * 0.125);
        }
    } else {
        const float* rx = p.gx + (int64_t)(gy - p.in_gy0) * p.ld_in + x;
        const float4 m = ldg4(rx);
        const float v[6] = {__ldg(rx - 1), m.x, m.y, m.z, m.w, __ldg(rx + 4)};
        const float* cy = p.gy + (int64_t)(gy - p.in_gy0) * p.ld_in + x;
        const float4 up = ldg4(cy - p.ld_in), dn = ldg4(cy + p.ld_in);
        const float u[4] = {up.x, up.y, up.z, up.w}, d[4] = {dn.x, dn.y, dn.z, dn.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            dxv[j] = __fmul_rn(__fsub_rn(v[j + 2], v[j]), 0.5f);
            dyv[j] = __fmul_rn(__fsub_rn(d[j], u[j]), 0.5f);
        }
    }
    float sl[4], as[4];
    if (p.normalize) {
        float rxv[4], ryv[4];
        if (p.res_x_2d) {
            const float4 t = ldg4(p.res_xf + (int64_t)gy * p.nx + x);
            rxv[0] = t.x, rxv[1] = t.y, rxv[2] = t.z, rxv[3] = t.w;
        } else {
            const float4 t = ldg4(p.res_xf + x);
            rxv[0] = t.x, rxv[1] = t.y, rxv[2] = t.z, rxv[3] = t.w;
        }
        if (p.res_y_2d) {
            const float4 t = ldg4(p.res_yf + (int64_t)gy * p.nx + x);
            ryv[0] = t.x, ryv[1] = t.y, ryv[2] = t.z, ryv[3] = t.w;
        } else {
            ryv[0] = ryv[1] = ryv[2] = ryv[3] = __ldg(p.res_yf + gy);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            dxv[j] = __fdiv_rn(dxv[j], rxv[j]);
            dyv[j] = __fdiv_rn(dyv[j], ryv[j]);
        }
    }
    const int64_t o = (int64_t)(gy - p.out_gy0) * p.ld_out + x;
    st4_streaming(p.dx + o, make_float4(dxv[0], dxv[1], dxv[2], dxv[3]));
    st4_streaming(p.dy + o, make_float4(dyv[0], dyv[1], dyv[2], dyv[3]));
    if (p.slope) {
#pragma unroll
        for (int j = 0; j < 4; ++j) slope_aspect(dxv[j], dyv[j], sl[j], as[j]);
        st4_streaming(p.slope + o, make_float4(sl[0], sl[1], sl[2], sl[3]));
        st4_streaming(p.aspect + o, make_float4(as[0], as[1], as[2], as[3]));
    }
}

// The weight table of the column kernel lives in shared memory: 8 B per tap.  Up to 48 KB (radius 3064) without
// ceremony, beyond that (the reference's own example goes to 100 km = radius 4001 on a 25 m grid) the kernel opts in
// to 224 KB (radius ~14 300; one CTA per SM from 113 KB on).
static int axis0_smem_opt_in(size_t smem, int lw) {
    constexpr int kMax = 224 * 1024;  // 227 KB per CTA minus the kernel's static shared memory
    TOPO_CHECK(smem <= (size_t)kMax, "gaussian radius %d too large (the weight table must fit 224 KB of shared memory)", lw);
    if (smem <= 48 * 1024) return 0;
    static bool attr_set[64] = {false};
    int dev = 0;
    TOPO_CUDA(cudaGetDevice(&dev));
    if (dev >= 64 || !attr_set[dev]) {
        TOPO_CUDA(cudaFuncSetAttribute(gauss_axis0_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMax));
        TOPO_CUDA(cudaFuncSetAttribute(gauss_axis0_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMax));
        if (dev < 64) attr_set[dev] = true;
    }
    return 0;
}

static int check_rows_reflect(const topo_view* v, int lo_off, int hi_off, const char* what) {
    // rows out_gy0+lo_off .. out_gy0+out_rows-1+hi_off, reflected into the image, must be in the band
    const int a = v->out_gy0 + lo_off, b = v->out_gy0 + v->out_rows - 1 + hi_off;
    int need_lo = a < 0 ? 0 : a, need_hi = b >= v->gny ? v->gny - 1 : b;
    if (-a > v->gny || b - v->gny + 1 > v->gny) need_lo = 0, need_hi = v->gny - 1;  // multiple reflections
    if (a < 0) {  // reflection of [a, -1] is [0, -a-1]
        int r = -a - 1;
        if (r >= v->gny) r = v->gny - 1;
        if (r > need_hi) need_hi = r;
    }
    if (b >= v->gny) {  // reflection of [gny, b] is [2*gny-1-b, gny-1]
        int r = 2 * v->gny - 1 - b;
        if (r < 0) r = 0;
        if (r < need_lo) need_lo = r;
    }
    TOPO_CHECK(v->in_gy0 <= need_lo && v->in_gy0 + v->in_rows > need_hi,
               "%s: input band [%d,%d) does not cover rows [%d,%d]", what, v->in_gy0, v->in_gy0 + v->in_rows,
               need_lo, need_hi);
    return 0;
}

// Which passes take the FFT route.  NaN-exact smoothing stays on the direct kernels: a transform would spread a
// non-finite sample over its whole segment instead of scipy's +-lw.
static bool fft_radius(int lw) { return lw >= kFftMinRadius && fft_conv_length(lw) > 0 && option_enabled(kOptGaussFft); }

struct GaussWs {
    size_t tmp;       // axis-0 result, out_rows x pitch (when both axes run)
    size_t t1, t2;    // transposed planes
    size_t tables;    // FFT twiddles + multipliers (one set, reused by the two axes)
    size_t total;
};

static GaussWs gauss_ws_layout(const topo_view* v, int lw_y, int lw_x, bool do_y, bool do_x, bool allow_fft) {
    auto align = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t rows = (size_t)(v->out_rows > 0 ? v->out_rows : 1), in_rows = (size_t)v->in_rows;
    const size_t pitch = ((size_t)v->nx + 3) & ~(size_t)3;
    const bool fy = allow_fft && do_y && fft_radius(lw_y), fx = allow_fft && do_x && fft_radius(lw_x);
    GaussWs w{};
    size_t off = 0;
    w.tmp = off;
    if (do_y && do_x) off = align(off + pitch * rows * sizeof(float));
    // transposed planes: FFT axis 0 needs (nx x in_rows) and (nx x out_rows); the direct wide axis-1 route needs
    // (nx x out_rows) twice
    const size_t tp_in = ((in_rows + 3) & ~(size_t)3), tp_out = ((rows + 3) & ~(size_t)3);
    size_t n1 = 0, n2 = 0;
    if (fy) n1 = tp_in * v->nx, n2 = tp_out * v->nx;
    if (do_x && !fx && lw_x > kAxis1SmemMaxRadius) n1 = n1 > tp_out * v->nx ? n1 : tp_out * v->nx, n2 = n2 > tp_out * v->nx ? n2 : tp_out * v->nx;
    w.t1 = off;
    off = align(off + n1 * sizeof(float));
    w.t2 = off;
    off = align(off + n2 * sizeof(float));
    w.tables = off;
    size_t tb = 0;
    if (fy) tb = fft_conv_table_bytes(lw_y);
    if (fx && fft_conv_table_bytes(lw_x) > tb) tb = fft_conv_table_bytes(lw_x);
    off = align(off + tb);
    w.total = off;
    return w;
}

}  // namespace topo

using namespace topo;

extern "C" {

size_t topo_gauss_workspace_bytes(const topo_view* v, int lw_y, int lw_x) {
    if (!v) return 0;
    // the route depends on nan_safe, which the query does not know: room for either
    const size_t a = gauss_ws_layout(v, lw_y, lw_x, lw_y >= 0, lw_x >= 0, true).total;
    const size_t b = gauss_ws_layout(v, lw_y, lw_x, lw_y >= 0, lw_x >= 0, false).total;
    return a > b ? a : b;
}

int topo_gauss_f32(const float* in, int64_t ld_in, float* out, int64_t ld_out, const topo_view* v,
                   const double* w_y, int lw_y, const double* w_x, int lw_x, int nan_safe, void* ws, size_t ws_bytes,
                   void* stream) {
    TOPO_CHECK(in && out, "null pointer");
    if (validate_view(v)) return -1;
    TOPO_CHECK(ld_in >= v->nx && ld_out >= v->nx, "row pitch smaller than nx");
    TOPO_CHECK(lw_y >= 0 && lw_x >= 0, "negative radius");
    if (v->out_rows == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const bool do_y = w_y != nullptr, do_x = w_x != nullptr;
    TOPO_CHECK(in != out, "in-place smoothing is not supported");
    const GaussWs L = gauss_ws_layout(v, lw_y, lw_x, do_y, do_x, !nan_safe);
    TOPO_CHECK(L.total == 0 || (ws && ws_bytes >= L.total), "workspace too small: need %zu bytes, got %zu", L.total, ws_bytes);
    TOPO_CHECK(L.total == 0 || (reinterpret_cast<uintptr_t>(ws) & 15) == 0, "workspace must be 16-byte aligned");
    unsigned char* wsb = reinterpret_cast<unsigned char*>(ws);
    const bool fy = !nan_safe && do_y && fft_radius(lw_y), fx = !nan_safe && do_x && fft_radius(lw_x);
    const int64_t pitch = ((int64_t)v->nx + 3) & ~(int64_t)3;
    const int64_t tp_in = ((int64_t)v->in_rows + 3) & ~(int64_t)3, tp_out = ((int64_t)v->out_rows + 3) & ~(int64_t)3;

    const float* cur = in;
    int64_t cur_ld = ld_in;
    int cur_gy0 = v->in_gy0, cur_rows = v->in_rows;
    if (do_y) {
        if (check_rows_reflect(v, -lw_y, lw_y, "gaussian axis 0")) return -1;
        float* dst = do_x ? reinterpret_cast<float*>(wsb + L.tmp) : out;
        const int64_t dst_ld = do_x ? pitch : ld_out;
        if (fy) {
            // transpose the band, filter its lines (= the image columns) with the FFT pass, transpose back
            float* t1 = reinterpret_cast<float*>(wsb + L.t1);
            float* t2 = reinterpret_cast<float*>(wsb + L.t2);
            dim3 tg(ceil_div(v->nx, 64), ceil_div(v->in_rows, 64));
            TOPO_LAUNCH("transpose", s, transpose_kernel<<<tg, dim3(32, 8), 0, s>>>(in, ld_in, t1, tp_in, v->in_rows, v->nx));
            if (fft_conv_rows(t1, tp_in, t2, tp_out, v->nx, v->gny, v->in_gy0, v->in_rows, v->out_gy0, v->out_rows, w_y, lw_y,
                              wsb + L.tables, s))
                return -1;
            dim3 tg2(ceil_div(v->out_rows, 64), ceil_div(v->nx, 64));
            TOPO_LAUNCH("transpose", s, transpose_kernel<<<tg2, dim3(32, 8), 0, s>>>(t2, tp_out, dst, dst_ld, v->nx, v->out_rows));
        } else {
            GaussParams p{cur, dst, cur_ld, dst_ld, v->nx, v->gny, cur_gy0, cur_rows, v->out_gy0, v->out_rows, w_y, lw_y};
            dim3 grid(ceil_div(v->nx, kA0Cols), ceil_div(v->out_rows, 8 * kK));
            const size_t smem = (size_t)gauss_steps(lw_y) * sizeof(double);
            if (axis0_smem_opt_in(smem, lw_y)) return -1;
            if (nan_safe)
                TOPO_LAUNCH("gauss_axis0<nansafe>", s, gauss_axis0_kernel<true><<<grid, dim3(32, 8), smem, s>>>(p));
            else
                TOPO_LAUNCH("gauss_axis0", s, gauss_axis0_kernel<false><<<grid, dim3(32, 8), smem, s>>>(p));
        }
        cur = dst, cur_ld = dst_ld, cur_gy0 = v->out_gy0, cur_rows = v->out_rows;
    } else {
        TOPO_CHECK(v->in_gy0 <= v->out_gy0 && v->in_gy0 + v->in_rows >= v->out_gy0 + v->out_rows,
                   "input band does not cover the output rows");
    }
    const float* src = cur + (int64_t)(v->out_gy0 - cur_gy0) * cur_ld;  // first output row in `cur`
    if (fx) {
        if (fft_conv_rows(src, cur_ld, out, ld_out, v->out_rows, v->nx, 0, v->nx, 0, v->nx, w_x, lw_x, wsb + L.tables, s)) return -1;
    } else if (do_x && lw_x > kAxis1SmemMaxRadius) {
        // wide radius, NaN-exact: transpose -> column kernel -> transpose back
        float* t1 = reinterpret_cast<float*>(wsb + L.t1);
        float* t2 = reinterpret_cast<float*>(wsb + L.t2);
        dim3 tg(ceil_div(v->nx, 64), ceil_div(v->out_rows, 64));
        TOPO_LAUNCH("transpose", s, transpose_kernel<<<tg, dim3(32, 8), 0, s>>>(src, cur_ld, t1, tp_out, v->out_rows, v->nx));
        GaussParams p{t1, t2, tp_out, tp_out, v->out_rows, v->nx, 0, v->nx, 0, v->nx, w_x, lw_x};
        dim3 grid(ceil_div(v->out_rows, kA0Cols), ceil_div(v->nx, 8 * kK));
        const size_t smem = (size_t)gauss_steps(lw_x) * sizeof(double);
        if (axis0_smem_opt_in(smem, lw_x)) return -1;
        if (nan_safe)
            TOPO_LAUNCH("gauss_axis0<nansafe>", s, gauss_axis0_kernel<true><<<grid, dim3(32, 8), smem, s>>>(p));
        else
            TOPO_LAUNCH("gauss_axis0", s, gauss_axis0_kernel<false><<<grid, dim3(32, 8), smem, s>>>(p));
        dim3 tg2(ceil_div(v->out_rows, 64), ceil_div(v->nx, 64));
        TOPO_LAUNCH("transpose", s, transpose_kernel<<<tg2, dim3(32, 8), 0, s>>>(t2, tp_out, out, ld_out, v->nx, v->out_rows));
    } else if (do_x) {
        GaussParams p{cur, out, cur_ld, ld_out, v->nx, v->gny, cur_gy0, cur_rows, v->out_gy0, v->out_rows, w_x, lw_x};
        const int rc = lw_x <= 12 ? launch_axis1<8>(p, nan_safe, s) : launch_axis1<16>(p, nan_safe, s);
        if (rc) return rc;
    } else if (!do_y) {
        TOPO_CUDA(cudaMemcpy2DAsync(out, ld_out * sizeof(float),
                                    in + (int64_t)(v->out_gy0 - v->in_gy0) * ld_in, ld_in * sizeof(float),
                                    (size_t)v->nx * sizeof(float), v->out_rows, cudaMemcpyDeviceToDevice, s));
    }
    return 0;
}

static int run_grad(bool sobel, const float* gx, const float* gy, int64_t ld_in, float* dx, float* dy, float* slope,
                    float* aspect, int64_t ld_out, const topo_view* v, const double* res_x, int res_x_2d,
                    const double* res_y, int res_y_2d, const float* res_xf, const float* res_yf, int normalize,
                    void* stream) {
    TOPO_CHECK(gx && gy && dx && dy, "null pointer");
    if (validate_view(v)) return -1;
    TOPO_CHECK(v->nx >= 2 && v->gny >= 2, "gradient needs at least 2 x 2 pixels");
    TOPO_CHECK(!normalize || (res_x && res_y), "missing resolution arrays");
    TOPO_CHECK((slope == nullptr) == (aspect == nullptr), "slope and aspect go together");
    if (v->out_rows == 0) return 0;
    if (check_rows_reflect(v, -1, 1, sobel ? "sobel" : "gradient")) return -1;
    GradParams p{gx, gy, dx, dy, slope, aspect, ld_in, ld_out, v->nx, v->gny, v->in_gy0, v->in_rows,
                 v->out_gy0, v->out_rows, res_x, res_y, res_xf, res_yf, res_x_2d, res_y_2d, normalize};
    // vector path: 16-byte aligned rows everywhere and float32 resolutions at hand
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    int vec_ok = (ld_in % 4 == 0) && (ld_out % 4 == 0) && (v->nx % 4 == 0) && al16(gx) && al16(gy) && al16(dx) && al16(dy) &&
                 (!slope || (al16(slope) && al16(aspect)));
    if (normalize) vec_ok = vec_ok && res_xf && res_yf && al16(res_xf) && al16(res_yf);
    dim3 grid(ceil_div(v->nx, 256), ceil_div(v->out_rows, 4));
    cudaStream_t s = (cudaStream_t)stream;
    if (sobel)
        TOPO_LAUNCH("sobel_gradient", s, gradient_kernel<true><<<grid, 256, 0, s>>>(p, vec_ok));
    else
        TOPO_LAUNCH("grad_from_smooth", s, gradient_kernel<false><<<grid, 256, 0, s>>>(p, vec_ok));
    return 0;
}

// Rows of smoothed DEM the gradient of rows [out_gy0, out_gy0 + out_rows) reads (one more on each side, clamped)
static void grad_smooth_rows(const topo_view* v, int& g0, int& g1) {
    g0 = v->out_gy0 - 1 < 0 ? 0 : v->out_gy0 - 1;
    g1 = v->out_gy0 + v->out_rows + 1 > v->gny ? v->gny : v->out_gy0 + v->out_rows + 1;
}

static bool gradient_is_fused(int lw, int nan_safe) {
    return !nan_safe && lw >= 1 && lw <= kFusedMaxRadius && option_enabled(kOptGradFused);
}

size_t topo_gradient_workspace_bytes(const topo_view* v, int lw) {
    if (!v) return 0;
    // the three-kernel route (wide radii, NaN-exact smoothing, or the fused shape switched off): smoothed rows + the
    // Gaussian's own workspace.  The fused shape needs none; the query does not know nan_safe: room for either.
    int g0, g1;
    grad_smooth_rows(v, g0, g1);
    topo_view sv = *v;
    sv.out_gy0 = g0, sv.out_rows = g1 - g0;
    const size_t pitch = ((size_t)v->nx + 3) & ~(size_t)3;
    const size_t smooth = (pitch * (size_t)(g1 - g0) * sizeof(float) + 255) & ~(size_t)255;
    return smooth + topo_gauss_workspace_bytes(&sv, lw, lw);
}

int topo_gradient_f32(const float* dem, int64_t ld_in, float* dx, float* dy, float* slope, float* aspect, int64_t ld_out,
                      const topo_view* v, const double* w, int lw, int nan_safe, const double* res_x, int res_x_2d,
                      const double* res_y, int res_y_2d, const float* res_xf, const float* res_yf, void* ws,
                      size_t ws_bytes, void* stream) {
    TOPO_CHECK(dem && dx && dy && w && res_x && res_y, "null pointer");
    if (validate_view(v)) return -1;
    TOPO_CHECK(v->nx >= 2 && v->gny >= 2, "gradient needs at least 2 x 2 pixels");
    TOPO_CHECK(lw >= 0, "negative radius");
    TOPO_CHECK((slope == nullptr) == (aspect == nullptr), "slope and aspect go together");
    TOPO_CHECK(ld_in >= v->nx && ld_out >= v->nx, "row pitch smaller than nx");
    if (v->out_rows == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    int g0, g1;
    grad_smooth_rows(v, g0, g1);
    if (gradient_is_fused(lw, nan_safe)) {
        if (check_rows_reflect(v, -lw - 1, lw + 1, "fused gradient")) return -1;
        GradParams p{dem, dem, dx, dy, slope, aspect, ld_in, ld_out, v->nx, v->gny, v->in_gy0, v->in_rows,
                     v->out_gy0, v->out_rows, res_x, res_y, res_xf, res_yf, res_x_2d, res_y_2d, 1};
        auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
        FusedParams fp;
        fp.g = p, fp.w = w, fp.lw = lw;
        fp.vec_ok = (ld_out % 4 == 0) && (v->nx % 4 == 0) && al16(dx) && al16(dy) && (!slope || (al16(slope) && al16(aspect))) &&
                    res_xf && res_yf && al16(res_xf) && al16(res_yf);
        fp.tiles_x = ceil_div(v->nx, kFuTW), fp.tiles_y = ceil_div(v->out_rows, kFuTH);
        TOPO_CHECK((long long)fp.tiles_x * fp.tiles_y < 2147483647ll, "too many tiles");
        const int ntiles = fp.tiles_x * fp.tiles_y;
#define TOPO_FUSED_LAUNCH(LWT)                                                                                           \
    do {                                                                                                                 \
        using SH = FusedShape<LWT>;                                                                                      \
        static bool attr_set[64] = {false};                                                                              \
        int dev = 0;                                                                                                     \
        TOPO_CUDA(cudaGetDevice(&dev));                                                                                  \
        if (dev >= 64 || !attr_set[dev]) {                                                                               \
            TOPO_CUDA(cudaFuncSetAttribute(gauss_grad_fused_kernel<LWT>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                           (int)SH::BYTES));                                                             \
            if (dev < 64) attr_set[dev] = true;                                                                          \
        }                                                                                                                \
        CUtensorMap tmap;                                                                                                \
        memset(&tmap, 0, sizeof(tmap));                                                                                  \
        fp.use_tma = make_tmap_2d_f32(&tmap, dem, (uint64_t)v->nx, (uint64_t)v->in_rows, (uint64_t)ld_in, SH::RAW_COLS,  \
                                      SH::RAW_ROWS) ? 1 : 0;                                                             \
        const int per_sm = (int)(227 * 1024 / (SH::BYTES + 1024));                                                       \
        int ctas = kNumSMs * (per_sm < 1 ? 1 : per_sm);                                                                  \
        if (ctas > ntiles) ctas = ntiles;                                                                                \
        TOPO_LAUNCH("gauss_grad_fused", s, gauss_grad_fused_kernel<LWT><<<ctas, 256, SH::BYTES, s>>>(tmap, fp));         \
    } while (0)
        if (lw <= 5) TOPO_FUSED_LAUNCH(5);
        else if (lw <= 9) TOPO_FUSED_LAUNCH(9);
        else if (lw <= 13) TOPO_FUSED_LAUNCH(13);
        else TOPO_FUSED_LAUNCH(21);
#undef TOPO_FUSED_LAUNCH
        return 0;
    }
    // three kernels: smooth rows [g0, g1), then differentiate
    topo_view sv = *v;
    sv.out_gy0 = g0, sv.out_rows = g1 - g0;
    const int64_t pitch = ((int64_t)v->nx + 3) & ~(int64_t)3;
    const size_t smooth = ((size_t)pitch * (size_t)(g1 - g0) * sizeof(float) + 255) & ~(size_t)255;
    TOPO_CHECK(ws && ws_bytes >= smooth + topo_gauss_workspace_bytes(&sv, lw, lw), "workspace too small: need %zu bytes, got %zu",
               smooth + topo_gauss_workspace_bytes(&sv, lw, lw), ws_bytes);
    TOPO_CHECK((reinterpret_cast<uintptr_t>(ws) & 255) == 0, "workspace must be 256-byte aligned");
    float* sm = reinterpret_cast<float*>(ws);
    int rc = topo_gauss_f32(dem, ld_in, sm, pitch, &sv, w, lw, w, lw, nan_safe, reinterpret_cast<unsigned char*>(ws) + smooth,
                            ws_bytes - smooth, stream);
    if (rc) return rc;
    topo_view gv = *v;
    gv.in_gy0 = g0, gv.in_rows = g1 - g0;
    return run_grad(false, sm, sm, pitch, dx, dy, slope, aspect, ld_out, &gv, res_x, res_x_2d, res_y, res_y_2d, res_xf, res_yf, 1,
                    stream);
}

int topo_grad_from_smooth_f32(const float* gx, const float* gy, int64_t ld_in, float* dx, float* dy, float* slope,
                              float* aspect, int64_t ld_out, const topo_view* v, const double* res_x, int res_x_2d,
                              const double* res_y, int res_y_2d, const float* res_xf, const float* res_yf, void* stream) {
    return run_grad(false, gx, gy, ld_in, dx, dy, slope, aspect, ld_out, v, res_x, res_x_2d, res_y, res_y_2d, res_xf,
                    res_yf, 1, stream);
}

int topo_sobel_gradient_f32(const float* dem, int64_t ld_in, float* dx, float* dy, float* slope, float* aspect,
                            int64_t ld_out, const topo_view* v, const double* res_x, int res_x_2d,
                            const double* res_y, int res_y_2d, const float* res_xf, const float* res_yf, int normalize,
                            void* stream) {
    return run_grad(true, dem, dem, ld_in, dx, dy, slope, aspect, ld_out, v, res_x, res_x_2d, res_y, res_y_2d, res_xf,
                    res_yf, normalize, stream);
}

}  // extern "C"
