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

#include "common.cuh"

namespace topo {

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

// 32 x 32 tiled transpose (rows x cols -> cols x rows); lets the wide-radius axis-1 pass reuse the
// column kernel above at full lane occupancy.
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ in, int64_t ld_in,
                                                        float* __restrict__ out, int64_t ld_out, int rows, int cols) {
    __shared__ float t[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        if (r < rows && c < cols) t[i][threadIdx.x] = __ldg(in + (int64_t)r * ld_in + c);
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (c < cols && r < rows) out[(int64_t)c * ld_out + r] = t[threadIdx.x][i];
    }
}

// ---- axis 1 (along x), radius <= kAxis1SmemMaxRadius ---------------------------------------------------
// block 256 = 8 warps.  Tile: 32 rows (lanes) x 128 output columns; warp w owns columns [w*K, w*K+K).
constexpr int kA1Cols = 8 * kK;

template <bool NANSAFE>
__global__ void __launch_bounds__(256) gauss_axis1_kernel(const GaussParams p, int pitch) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* wfull = reinterpret_cast<double*>(smem_raw);
    const int nsteps = gauss_steps(p.lw);
    float* tile = reinterpret_cast<float*>(wfull + nsteps);  // [32][pitch]
    float* otile = tile + (size_t)32 * pitch;                // [32][kA1Cols + 1]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lw = p.lw;
    stage_weights(p, wfull, threadIdx.x, 256);

    const int x0 = blockIdx.x * kA1Cols;
    const int r0 = blockIdx.y * 32;  // first row (relative to the output band) of this tile
    const int span = kA1Cols + nsteps - kK;  // columns a thread group can touch: c0 + n, n < nsteps
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
        const int c0 = warp * kK;  // first output column (tile-relative) of this thread
        const float* tp = tile + (size_t)lane * pitch + c0;  // input x0 + c0 + (n - lw) sits at tile column c0 + n
        const double* wp = wfull;
        double acc[kK], wr[kK];
#pragma unroll
        for (int k = 0; k < kK; ++k) acc[k] = 0.0, wr[k] = 0.0;
        for (int n0 = 0; n0 < nsteps; n0 += kK, wp += kK, tp += kK) {
#pragma unroll
            for (int s = 0; s < kK; ++s) {
#pragma unroll
                for (int k = kK - 1; k > 0; --k) wr[k] = wr[k - 1];
                wr[0] = wp[s];
                const double dv = (double)tp[s];
#pragma unroll
                for (int k = 0; k < kK; ++k)
                    if (!NANSAFE || (unsigned)(n0 + s - k) <= (unsigned)(2 * lw)) acc[k] = fma(wr[k], dv, acc[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < kK; ++k) otile[lane * (kA1Cols + 1) + c0 + k] = (float)acc[k];
    }
    __syncthreads();
    for (int r = warp; r < 32; r += 8) {
        const int row = r0 + r;
        if (row >= p.out_rows) break;
        float* dst = p.out + (int64_t)row * p.ld_out;
        for (int c = lane; c < kA1Cols; c += 32)
            if (x0 + c < p.nx) dst[x0 + c] = otile[r * (kA1Cols + 1) + c];
    }
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
        // np.arctan(np.sqrt(dx**2 + dy**2)) * (180 / np.pi), every step in float32
        const float h2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        p.slope[o] = __fmul_rn(atanf(sqrtf(h2)), 57.29577951308232f);
    }
    if (p.aspect) {
        // (180 + np.degrees(np.arctan2(dx, dy))) % 360 in float32, Python modulo
        // np.degrees on float32 multiplies by 180.0f / NPY_PIf = 57.2957763671875f (one ulp below
        // float32(180/pi) used for the slope above)
        const float deg = __fmul_rn(atan2f(dx, dy), 57.2957763671875f);
        // a is in [0, 360] (|deg| <= 180), so Python's % 360 only maps 360 -> 0; NaN stays NaN
        float a = __fadd_rn(180.0f, deg);
        if (a >= 360.0f) a -= 360.0f;
        p.aspect[o] = a;
    }
}

__global__ void __launch_bounds__(256) grad_from_smooth_kernel(const GradParams p) {
    const int x = blockIdx.x * 64 + (threadIdx.x & 63);
    const int gy = p.out_gy0 + blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x >= p.nx || gy >= p.out_gy0 + p.out_rows) return;
    const float* rx = p.gx + (int64_t)(gy - p.in_gy0) * p.ld_in;
    float dx, dy;
    if (x == 0)
        dx = __fsub_rn(__ldg(rx + 1), __ldg(rx));
    else if (x == p.nx - 1)
        dx = __fsub_rn(__ldg(rx + x), __ldg(rx + x - 1));
    else
        dx = __fdiv_rn(__fsub_rn(__ldg(rx + x + 1), __ldg(rx + x - 1)), 2.0f);
    const float* cy = p.gy + (int64_t)(gy - p.in_gy0) * p.ld_in + x;
    if (gy == 0)
        dy = __fsub_rn(__ldg(cy + p.ld_in), __ldg(cy));
    else if (gy == p.gny - 1)
        dy = __fsub_rn(__ldg(cy), __ldg(cy - p.ld_in));
    else
        dy = __fdiv_rn(__fsub_rn(__ldg(cy + p.ld_in), __ldg(cy - p.ld_in)), 2.0f);
    finish_gradient(p, dx, dy, gy, x);
}

__global__ void __launch_bounds__(256) sobel_gradient_kernel(const GradParams p) {
    const int x = blockIdx.x * 64 + (threadIdx.x & 63);
    const int gy = p.out_gy0 + blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x >= p.nx || gy >= p.out_gy0 + p.out_rows) return;
    const int xm = reflect_index(x - 1, p.nx), xp = reflect_index(x + 1, p.nx);
    const float* r0 = p.gx + (int64_t)(reflect_index(gy - 1, p.gny) - p.in_gy0) * p.ld_in;
    const float* r1 = p.gx + (int64_t)(gy - p.in_gy0) * p.ld_in;
    const float* r2 = p.gx + (int64_t)(reflect_index(gy + 1, p.gny) - p.in_gy0) * p.ld_in;
    const double a = __ldg(r0 + xm), b = __ldg(r0 + x), c = __ldg(r0 + xp);
    const double d = __ldg(r1 + xm), f = __ldg(r1 + xp);
    const double g = __ldg(r2 + xm), h = __ldg(r2 + x), i = __ldg(r2 + xp);
    // exact in float64 (<= 6 terms of 24-bit values), one rounding to float32 like ndimage
    const float dx = (float)(((c + 2.0 * f + i) - (a + 2.0 * d + g)) * 0.125);
    const float dy = (float)(((g + 2.0 * h + i) - (a + 2.0 * b + c)) * 0.125);
    finish_gradient(p, dx, dy, gy, x);
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

}  // namespace topo

using namespace topo;

extern "C" {

size_t topo_gauss_workspace_bytes(const topo_view* v, int lw_y, int lw_x) {
    if (!v) return 0;
    (void)lw_y;
    const size_t rows = (size_t)(v->out_rows > 0 ? v->out_rows : 1);
    // axis-0 result for the output rows (pitch = nx rounded to 4) ...
    const size_t pitch = ((size_t)v->nx + 3) & ~(size_t)3;
    size_t bytes = pitch * rows * sizeof(float);
    // ... plus two transposed planes when the axis-1 radius takes the transpose route
    if (lw_x > kAxis1SmemMaxRadius) bytes += 2 * ((rows + 3) & ~(size_t)3) * (size_t)v->nx * sizeof(float);
    return bytes;
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

    const float* cur = in;
    int64_t cur_ld = ld_in;
    int cur_gy0 = v->in_gy0, cur_rows = v->in_rows;
    if (do_y) {
        if (check_rows_reflect(v, -lw_y, lw_y, "gaussian axis 0")) return -1;
        float* dst = out;
        int64_t dst_ld = ld_out;
        if (do_x) {
            const int64_t pitch = ((int64_t)v->nx + 3) & ~(int64_t)3;
            TOPO_CHECK(ws && ws_bytes >= (size_t)pitch * v->out_rows * sizeof(float), "workspace too small");
            dst = (float*)ws;
            dst_ld = pitch;
        }
        GaussParams p{cur, dst, cur_ld, dst_ld, v->nx, v->gny, cur_gy0, cur_rows, v->out_gy0, v->out_rows, w_y, lw_y};
        dim3 grid(ceil_div(v->nx, kA0Cols), ceil_div(v->out_rows, 8 * kK));
        const size_t smem = (size_t)gauss_steps(lw_y) * sizeof(double);
        TOPO_CHECK(smem <= 48 * 1024, "gaussian radius %d too large", lw_y);
        if (nan_safe)
            TOPO_LAUNCH("gauss_axis0<nansafe>", s, gauss_axis0_kernel<true><<<grid, dim3(32, 8), smem, s>>>(p));
        else
            TOPO_LAUNCH("gauss_axis0", s, gauss_axis0_kernel<false><<<grid, dim3(32, 8), smem, s>>>(p));
        cur = dst, cur_ld = dst_ld, cur_gy0 = v->out_gy0, cur_rows = v->out_rows;
    } else {
        TOPO_CHECK(v->in_gy0 <= v->out_gy0 && v->in_gy0 + v->in_rows >= v->out_gy0 + v->out_rows,
                   "input band does not cover the output rows");
    }
    if (do_x && lw_x > kAxis1SmemMaxRadius) {
        // wide radius: transpose -> column kernel -> transpose back (two extra 8 B/px passes are noise next
        // to 2*lw+1 float64 FMAs per pixel)
        const int64_t pitch = ((int64_t)v->nx + 3) & ~(int64_t)3;
        const int64_t tpitch = ((int64_t)v->out_rows + 3) & ~(int64_t)3;
        const size_t need = (size_t)pitch * v->out_rows * sizeof(float) + 2 * (size_t)tpitch * v->nx * sizeof(float);
        TOPO_CHECK(ws && ws_bytes >= need, "workspace too small: need %zu bytes, got %zu", need, ws_bytes);
        float* t1 = (float*)ws + (size_t)pitch * v->out_rows;
        float* t2 = t1 + (size_t)tpitch * v->nx;
        const float* src = cur + (int64_t)(v->out_gy0 - cur_gy0) * cur_ld;
        dim3 tg(ceil_div(v->nx, 32), ceil_div(v->out_rows, 32));
        TOPO_LAUNCH("transpose", s, transpose_kernel<<<tg, dim3(32, 8), 0, s>>>(src, cur_ld, t1, tpitch, v->out_rows, v->nx));
        GaussParams p{t1, t2, tpitch, tpitch, v->out_rows, v->nx, 0, v->nx, 0, v->nx, w_x, lw_x};
        dim3 grid(ceil_div(v->out_rows, kA0Cols), ceil_div(v->nx, 8 * kK));
        const size_t smem = (size_t)gauss_steps(lw_x) * sizeof(double);
        TOPO_CHECK(smem <= 48 * 1024, "gaussian radius %d too large", lw_x);
        if (nan_safe)
            TOPO_LAUNCH("gauss_axis0<nansafe>", s, gauss_axis0_kernel<true><<<grid, dim3(32, 8), smem, s>>>(p));
        else
            TOPO_LAUNCH("gauss_axis0", s, gauss_axis0_kernel<false><<<grid, dim3(32, 8), smem, s>>>(p));
        dim3 tg2(ceil_div(v->out_rows, 32), ceil_div(v->nx, 32));
        TOPO_LAUNCH("transpose", s, transpose_kernel<<<tg2, dim3(32, 8), 0, s>>>(t2, tpitch, out, ld_out, v->nx, v->out_rows));
    } else if (do_x) {
        const int nsteps = gauss_steps(lw_x);
        int pitch = kA1Cols + nsteps - kK;
        pitch |= 1;  // odd pitch: lanes (rows) hit distinct banks
        const size_t smem = (size_t)nsteps * sizeof(double) + ((size_t)32 * pitch + (size_t)32 * (kA1Cols + 1)) * sizeof(float);
        static bool attr_set[64] = {false};
        int dev = 0;
        TOPO_CUDA(cudaGetDevice(&dev));
        if (dev < 64 && !attr_set[dev]) {
            TOPO_CUDA(cudaFuncSetAttribute(gauss_axis1_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            TOPO_CUDA(cudaFuncSetAttribute(gauss_axis1_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            attr_set[dev] = true;
        }
        GaussParams p{cur, out, cur_ld, ld_out, v->nx, v->gny, cur_gy0, cur_rows, v->out_gy0, v->out_rows, w_x, lw_x};
        dim3 grid(ceil_div(v->nx, kA1Cols), ceil_div(v->out_rows, 32));
        if (nan_safe)
            TOPO_LAUNCH("gauss_axis1<nansafe>", s, gauss_axis1_kernel<true><<<grid, 256, smem, s>>>(p, pitch));
        else
            TOPO_LAUNCH("gauss_axis1", s, gauss_axis1_kernel<false><<<grid, 256, smem, s>>>(p, pitch));
    } else if (!do_y) {
        TOPO_CUDA(cudaMemcpy2DAsync(out, ld_out * sizeof(float),
                                    in + (int64_t)(v->out_gy0 - v->in_gy0) * ld_in, ld_in * sizeof(float),
                                    (size_t)v->nx * sizeof(float), v->out_rows, cudaMemcpyDeviceToDevice, s));
    }
    return 0;
}

static int run_grad(bool sobel, const float* gx, const float* gy, int64_t ld_in, float* dx, float* dy, float* slope,
                    float* aspect, int64_t ld_out, const topo_view* v, const double* res_x, int res_x_2d,
                    const double* res_y, int res_y_2d, int normalize, void* stream) {
    TOPO_CHECK(gx && gy && dx && dy, "null pointer");
    if (validate_view(v)) return -1;
    TOPO_CHECK(v->nx >= 2 && v->gny >= 2, "gradient needs at least 2 x 2 pixels");
    TOPO_CHECK(!normalize || (res_x && res_y), "missing resolution arrays");
    if (v->out_rows == 0) return 0;
    if (check_rows_reflect(v, -1, 1, sobel ? "sobel" : "gradient")) return -1;
    GradParams p{gx, gy, dx, dy, slope, aspect, ld_in, ld_out, v->nx, v->gny, v->in_gy0, v->in_rows,
                 v->out_gy0, v->out_rows, res_x, res_y, res_x_2d, res_y_2d, normalize};
    dim3 grid(ceil_div(v->nx, 64), ceil_div(v->out_rows, 4));
    cudaStream_t s = (cudaStream_t)stream;
    if (sobel)
        TOPO_LAUNCH("sobel_gradient", s, sobel_gradient_kernel<<<grid, 256, 0, s>>>(p));
    else
        TOPO_LAUNCH("grad_from_smooth", s, grad_from_smooth_kernel<<<grid, 256, 0, s>>>(p));
    return 0;
}

int topo_grad_from_smooth_f32(const float* gx, const float* gy, int64_t ld_in, float* dx, float* dy, float* slope,
                              float* aspect, int64_t ld_out, const topo_view* v, const double* res_x, int res_x_2d,
                              const double* res_y, int res_y_2d, void* stream) {
    return run_grad(false, gx, gy, ld_in, dx, dy, slope, aspect, ld_out, v, res_x, res_x_2d, res_y, res_y_2d, 1, stream);
}

int topo_sobel_gradient_f32(const float* dem, int64_t ld_in, float* dx, float* dy, float* slope, float* aspect,
                            int64_t ld_out, const topo_view* v, const double* res_x, int res_x_2d,
                            const double* res_y, int res_y_2d, int normalize, void* stream) {
    return run_grad(true, dem, dem, ld_in, dx, dy, slope, aspect, ld_out, v, res_x, res_x_2d, res_y, res_y_2d,
                    normalize, stream);
}

}  // extern "C"
