// Valley / ridge index for LARGE kernels: the rotated-kernel bank applied by 2-D overlap-save FFT convolution.
// Reference: topo.py:441-452 -- 180 x `signal.convolve(dem3d, kernels_rot, "same")` (scipy picks its FFT method, so
// the reference's cost does not depend on the kernel size) with a running strict-'>' maximum over the angles.
// The direct bank of valley.cu costs O(size^2) per pixel and angle (684 ms per 4096^2 at size 41, 24 s at size 161);
// here every (angle, channel) kernel costs one inverse 2-D transform of each image tile, whatever its size.
//
//   tiles   : T x T windows of the zero-padded, z-scored DEM (T = 2048 / 4096 / 8192 by kernel extent), overlapping
//             by the kernel extent; their spectra D^ are computed once per call and kept.
//   kernels : two real kernels ride in one complex transform (real + i * imaginary: the DEM is real, so the real and
//             imaginary parts of the inverse transform are the two convolutions).  Every kernel is placed so that
//             scipy's "same" crop lands on ONE window offset (HB, WB) for all angles: a pixel's two values are then
//             read by the same thread, and the running (max, argmax) needs no atomics.
//   2-D FFT : row transform -> complex transpose -> row transform, all float64 in shared memory (fft_smem.cuh);
//             spectra stay in digit-reversed order in both dimensions, so no permutation pass exists.
//   per kernel pair: K^ (2 row passes + transpose), then for all tiles at once: D^ * K^ fused into the first inverse
//             pass, transpose, second inverse pass fused with the fold into the running maximum.
#include <math.h>

#include "fft2d.cuh"

namespace topo {

struct VfftGeom {
    int T;                 // transform length (window edge)
    int HT, HB, WT, WB;    // window rows above / below (columns left / right of) an output pixel that kernels reach
    int V_y, V_x;          // output rows / columns per tile
    int tiles_y, tiles_x;
    int nx, gny, in_gy0, in_rows, out_gy0, out_rows;
};

struct VfftKernelPair {
    const float* ka;  // h_a x w_a, row-major (true-convolution orientation, as scipy receives it)
    const float* kb;  // may be null
    int ha, wa, hb, wb;
    float angle_a, angle_b;
};

enum { VSRC_DEM = 0, VSRC_KERN = 1 };

struct VfftFwdParams {
    VfftGeom g;
    const float* dem;
    int64_t ld_in;
    VfftKernelPair kp;
    double2* dst;        // [planes][T][T], digit-reversed along the line
    const double2* tw;
};

// One line (row of a T x T plane) per CTA: forward transform, output in digit-reversed order.
template <int N, int SRC>
__global__ void __launch_bounds__(FftShape<N>::NT, N >= 8192 ? 1 : 3) vfft_fwd_kernel(const VfftFwdParams p) {
    using S = FftShape<N>;
    constexpr int NT = S::NT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* buf = reinterpret_cast<double2*>(smem_raw);
    const int tid = threadIdx.x;
    const int line = blockIdx.x, plane = blockIdx.y;
    const VfftGeom& g = p.g;
    double2* out = p.dst + ((int64_t)plane * N + line) * N;

    bool all_zero = false;  // CTA-uniform: lines that are entirely zero transform to zero
    int gy = 0, c0 = 0;
    const float* row = nullptr;
    if constexpr (SRC == VSRC_DEM) {
        const int ty = plane / g.tiles_x, tx = plane - ty * g.tiles_x;
        gy = g.out_gy0 + ty * g.V_y - g.HT + line;  // global row of this window row
        c0 = tx * g.V_x - g.WT;                     // global column of sample 0
        all_zero = gy < 0 || gy >= g.gny || gy < g.in_gy0 || gy >= g.in_gy0 + g.in_rows;
        if (!all_zero) row = p.dem + (int64_t)(gy - g.in_gy0) * p.ld_in;
    } else if constexpr (SRC == VSRC_KERN) {
        const int ia = line - (g.HB - (p.kp.ha - 1) / 2), ib = p.kp.kb ? line - (g.HB - (p.kp.hb - 1) / 2) : -1;
        all_zero = !((ia >= 0 && ia < p.kp.ha) || (ib >= 0 && ib < p.kp.hb));
    }
    if (all_zero) {
        for (int i = tid; i < N; i += NT) out[i] = make_double2(0.0, 0.0);
        return;
    }
    const double2* __restrict__ tw = p.tw;
    fft2d_forward_line<N>(buf, tw, tid, [&](int n) -> double2 {
        if constexpr (SRC == VSRC_DEM) {
            const int gx = c0 + n;
            return make_double2((gx >= 0 && gx < g.nx) ? (double)__ldg(row + gx) : 0.0, 0.0);
        } else {
            // kernel element (i, j) sits at window (i + HB - cy, j + WB - cx): one crop offset for every kernel
            double re = 0.0, im = 0.0;
            const int ia = line - (g.HB - (p.kp.ha - 1) / 2), ja = n - (g.WB - (p.kp.wa - 1) / 2);
            if (ia >= 0 && ia < p.kp.ha && ja >= 0 && ja < p.kp.wa) re = (double)__ldg(p.kp.ka + (int64_t)ia * p.kp.wa + ja);
            if (p.kp.kb) {
                const int ib = line - (g.HB - (p.kp.hb - 1) / 2), jb = n - (g.WB - (p.kp.wb - 1) / 2);
                if (ib >= 0 && ib < p.kp.hb && jb >= 0 && jb < p.kp.wb) im = (double)__ldg(p.kp.kb + (int64_t)ib * p.kp.wb + jb);
            }
            return make_double2(re, im);
        }
    }, out);
}

struct VfftInvParams {
    VfftGeom g;
    const double2* a;   // PRODUCT: D^ [planes][T][T];  FOLD: transposed first-pass result [planes][T][T]
    const double2* k;   // PRODUCT: K^ [T][T]
    double2* dst;       // PRODUCT: [planes][T][T] natural order along the line
    const double2* tw;
    float* norm;        // FOLD: running maximum (raw) and its angle
    float* dir;
    int64_t ld_out;
    float angle_a, angle_b;
    int has_b;
    double scale;       // 1 / T^2
};

// Second inverse pass, one line per CTA: sample n of line Y is the convolution value of window pixel (Y, n): real part =
// kernel a, imaginary part = kernel b; folded into the running strict-'>' (max, argmax) of the output pixel it belongs to.
template <int N>
__global__ void __launch_bounds__(FftShape<N>::NT, N >= 8192 ? 1 : 3) vfft_fold_kernel(const VfftInvParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* buf = reinterpret_cast<double2*>(smem_raw);
    const int tid = threadIdx.x;
    const int line = blockIdx.x, plane = blockIdx.y;
    const VfftGeom& g = p.g;
    const double2* __restrict__ src = p.a + ((int64_t)plane * N + line) * N;
    // window row `line` <-> output row gy = tile's first output row + line - (HT + HB)
    const int ty = plane / g.tiles_x, tx = plane - ty * g.tiles_x;
    const int oy0 = g.out_gy0 + ty * g.V_y;
    const int gy = oy0 + line - (g.HT + g.HB);
    const int oy1 = min(oy0 + g.V_y, g.out_gy0 + g.out_rows);
    if (gy < oy0 || gy >= oy1) return;  // halo rows of the window: nothing to fold (CTA-uniform)
    const int ox0 = tx * g.V_x;
    const int x_lo = g.WT + g.WB;  // window column of output column ox0
    const int x_hi = x_lo + min(g.V_x, g.nx - ox0);
    float* nrow = p.norm + (int64_t)(gy - g.out_gy0) * p.ld_out + ox0 - x_lo;
    float* drow = p.dir + (int64_t)(gy - g.out_gy0) * p.ld_out + ox0 - x_lo;
    fft2d_inverse_line_from<N>(buf, p.tw, tid, src, [&](int n, double2 y) {
        if (n < x_lo || n >= x_hi) return;
        float best = nrow[n], bdir = drow[n];
        const float va = (float)(y.x * p.scale);
        if (va > best) best = va, bdir = p.angle_a;
        if (p.has_b) {
            const float vb = (float)(y.y * p.scale);
            if (vb > best) best = vb, bdir = p.angle_b;
        }
        nrow[n] = best, drow[n] = bdir;
    });
}

__global__ void vfft_init_kernel(float* __restrict__ norm, float* __restrict__ dir, int64_t ld, int rows, int nx, int finish) {
    const int64_t total = (int64_t)rows * nx;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int y = (int)(i / nx), x = (int)(i - (int64_t)y * nx);
        const int64_t o = (int64_t)y * ld + x;
        if (finish) {
            norm[o] = fmaxf(norm[o], 0.f);  // np.ndarray.clip(index_norm, min=0), topo.py:452
        } else {
            norm[o] = -INFINITY, dir[o] = 0.f;
        }
    }
}

static int vfft_length(int hmax, int wmax) {
    const int k = hmax > wmax ? hmax : wmax;
    if (k <= 1024) return 2048;
    if (k <= 2048) return 4096;
    if (k <= 4096) return 8192;
    return 0;
}

static int vfft_geometry(const topo_view* v, int hmax, int wmax, VfftGeom& g) {
    g.T = vfft_length(hmax, wmax);
    TOPO_CHECK(g.T > 0, "kernel extent %d x %d too large for the FFT path", hmax, wmax);
    // a kernel of h rows reaches cy = (h-1)/2 rows below and h-1-cy rows above an output pixel ("same" crop of a true
    // convolution); bounds over the bank
    g.HB = (hmax - 1) / 2, g.HT = hmax - 1 - (hmax - 1) / 2;
    g.WB = (wmax - 1) / 2, g.WT = wmax - 1 - (wmax - 1) / 2;
    if (g.HT < hmax / 2) g.HT = hmax / 2;  // smaller kernels of the bank: h-1-cy <= hmax/2
    if (g.WT < wmax / 2) g.WT = wmax / 2;
    g.V_y = g.T - g.HT - g.HB, g.V_x = g.T - g.WT - g.WB;
    g.nx = v->nx, g.gny = v->gny, g.in_gy0 = v->in_gy0, g.in_rows = v->in_rows, g.out_gy0 = v->out_gy0, g.out_rows = v->out_rows;
    g.tiles_y = ceil_div(v->out_rows, g.V_y), g.tiles_x = ceil_div(v->nx, g.V_x);
    return 0;
}

struct VfftWs {
    size_t tw, dhat, x, y, k1, k2, total;
};

static VfftWs vfft_ws_layout(const VfftGeom& g) {
    auto align = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t plane = (size_t)g.T * g.T * sizeof(double2), planes = (size_t)g.tiles_y * g.tiles_x;
    VfftWs w;
    size_t off = 0;
    w.tw = off, off = align(off + (size_t)g.T * sizeof(double2));
    w.dhat = off, off = align(off + planes * plane);
    w.x = off, off = align(off + planes * plane);
    w.y = off, off = align(off + planes * plane);
    w.k1 = off, off = align(off + plane);
    w.k2 = off, off = align(off + plane);
    w.total = off;
    return w;
}

template <int N>
static int vfft_run(const VfftGeom& g, const float* dem, int64_t ld_in, float* norm, float* dir, int64_t ld_out,
                    const float* kernels, const long long* kern_off, const int* kern_hw, int n_kernels, unsigned char* ws,
                    cudaStream_t s) {
    using S = FftShape<N>;
    static bool attr_set[64] = {false};
    int dev = 0;
    TOPO_CUDA(cudaGetDevice(&dev));
    if (dev >= 64 || !attr_set[dev]) {
        TOPO_CUDA(cudaFuncSetAttribute(vfft_fwd_kernel<N, VSRC_DEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::SMEM));
        TOPO_CUDA(cudaFuncSetAttribute(vfft_fwd_kernel<N, VSRC_KERN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::SMEM));
        TOPO_CUDA(cudaFuncSetAttribute(vfft_fold_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::SMEM));
        if (fft2d_set_smem_attributes<N>()) return -2;
        if (dev < 64) attr_set[dev] = true;
    }
    const VfftWs L = vfft_ws_layout(g);
    double2* tw = reinterpret_cast<double2*>(ws + L.tw);
    double2* dhat = reinterpret_cast<double2*>(ws + L.dhat);
    double2* X = reinterpret_cast<double2*>(ws + L.x);
    double2* Y = reinterpret_cast<double2*>(ws + L.y);
    double2* K1 = reinterpret_cast<double2*>(ws + L.k1);
    double2* K2 = reinterpret_cast<double2*>(ws + L.k2);
    const int planes = g.tiles_y * g.tiles_x;
    TOPO_CHECK(planes <= 65535, "too many tiles for one launch");
    TOPO_LAUNCH("valley_fft_twiddles", s, fft_twiddle_kernel<<<ceil_div(N, 256), 256, 0, s>>>(tw, N));
    TOPO_LAUNCH("valley_fft_init", s, vfft_init_kernel<<<kNumSMs * 8, 256, 0, s>>>(norm, dir, ld_out, g.out_rows, g.nx, 0));

    // ---- spectra of the image tiles: rows -> transpose -> rows
    VfftFwdParams f{};
    f.g = g, f.dem = dem, f.ld_in = ld_in, f.tw = tw;
    const dim3 tgrid(N / 32, N / 32, planes), kgrid(N / 32, N / 32, 1);
    f.dst = X;
    TOPO_LAUNCH("valley_fft_fwd", s, (vfft_fwd_kernel<N, VSRC_DEM><<<dim3(N, planes), S::NT, S::SMEM, s>>>(f)));
    TOPO_LAUNCH("valley_fft_transpose", s, fft2d_transpose_kernel<<<tgrid, dim3(32, 8), 0, s>>>(X, Y, N, N));
    TOPO_LAUNCH("valley_fft_fwd", s, (fft2d_fwd_cplx_kernel<N><<<dim3(N, planes), S::NT, S::SMEM, s>>>(Y, dhat, tw)));

    VfftInvParams q{};
    q.g = g, q.tw = tw, q.norm = norm, q.dir = dir, q.ld_out = ld_out, q.scale = 1.0 / ((double)N * (double)N);
    for (int i = 0; i < n_kernels; i += 2) {
        VfftKernelPair kp{};
        kp.ka = kernels + kern_off[i], kp.ha = kern_hw[3 * i], kp.wa = kern_hw[3 * i + 1], kp.angle_a = (float)kern_hw[3 * i + 2];
        if (i + 1 < n_kernels) {
            kp.kb = kernels + kern_off[i + 1], kp.hb = kern_hw[3 * i + 3], kp.wb = kern_hw[3 * i + 4];
            kp.angle_b = (float)kern_hw[3 * i + 5];
        }
        // K^ of the pair
        f.kp = kp, f.dst = K1;
        TOPO_LAUNCH("valley_fft_fwd", s, (vfft_fwd_kernel<N, VSRC_KERN><<<dim3(N, 1), S::NT, S::SMEM, s>>>(f)));
        TOPO_LAUNCH("valley_fft_transpose", s, fft2d_transpose_kernel<<<kgrid, dim3(32, 8), 0, s>>>(K1, K2, N, N));
        TOPO_LAUNCH("valley_fft_fwd", s, (fft2d_fwd_cplx_kernel<N, true><<<dim3(N, 1), S::NT, S::SMEM, s>>>(K2, K1, tw)));
        // all tiles: D^ * K^ -> inverse along the rows' axis -> transpose -> inverse + fold
        // (only the window rows the fold pass uses travel through the first pass's stores and the transpose)
        const int ct0 = (g.HT + g.HB) / 32, ct1 = ceil_div(g.HT + g.HB + g.V_y, 32);
        TOPO_LAUNCH("valley_fft_inv", s, (fft2d_inv_product_kernel<N><<<dim3(planes, N), S::NT, S::SMEM, s>>>(dhat, K1, X, tw, ct0 * 32, ct1 * 32)));
        TOPO_LAUNCH("valley_fft_transpose", s, fft2d_transpose_kernel<<<dim3(ct1 - ct0, N / 32, planes), dim3(32, 8), 0, s>>>(X, Y, N, N, ct0));
        q.a = Y, q.angle_a = kp.angle_a, q.angle_b = kp.angle_b, q.has_b = kp.kb != nullptr;
        TOPO_LAUNCH("valley_fft_fold", s, (vfft_fold_kernel<N><<<dim3(N, planes), S::NT, S::SMEM, s>>>(q)));
    }
    TOPO_LAUNCH("valley_fft_init", s, vfft_init_kernel<<<kNumSMs * 8, 256, 0, s>>>(norm, dir, ld_out, g.out_rows, g.nx, 1));
    return 0;
}

}  // namespace topo

using namespace topo;

extern "C" {

size_t topo_valley_ridge_fft_workspace_bytes(const topo_view* v, int hmax, int wmax) {
    if (!v || hmax < 1 || wmax < 1 || validate_view(v)) return 0;
    VfftGeom g;
    if (vfft_geometry(v, hmax, wmax, g)) return 0;
    return vfft_ws_layout(g).total;
}

int topo_valley_ridge_fft_f32(const float* dem_norm, int64_t ld_in, float* norm, float* dir, int64_t ld_out, const topo_view* v,
                              const float* kernels, const long long* kern_off, const int* kern_hw, int n_kernels, int hmax,
                              int wmax, void* ws, size_t ws_bytes, void* stream) {
    TOPO_CHECK(dem_norm && norm && dir && kernels && kern_off && kern_hw, "null pointer");
    if (validate_view(v)) return -1;
    TOPO_CHECK(n_kernels >= 1 && hmax >= 1 && wmax >= 1, "bad kernel bank");
    TOPO_CHECK(ld_in >= v->nx && ld_out >= v->nx, "row pitch smaller than nx");
    if (v->out_rows == 0) return 0;
    VfftGeom g;
    if (vfft_geometry(v, hmax, wmax, g)) return -1;
    for (int i = 0; i < n_kernels; ++i) {
        const int h = kern_hw[3 * i], w = kern_hw[3 * i + 1];
        TOPO_CHECK(h >= 1 && h <= hmax && w >= 1 && w <= wmax, "kernel %d: extent %d x %d outside the bank maximum", i, h, w);
    }
    {
        const int lo = v->out_gy0 - g.HT > 0 ? v->out_gy0 - g.HT : 0;
        const int hi = v->out_gy0 + v->out_rows + g.HB < v->gny ? v->out_gy0 + v->out_rows + g.HB : v->gny;
        TOPO_CHECK(v->in_gy0 <= lo && v->in_gy0 + v->in_rows >= hi, "valley_ridge: band [%d,%d) does not cover rows [%d,%d)",
                   v->in_gy0, v->in_gy0 + v->in_rows, lo, hi);
    }
    const size_t need = vfft_ws_layout(g).total;
    TOPO_CHECK(ws && ws_bytes >= need, "workspace too small: need %zu bytes, got %zu", need, ws_bytes);
    TOPO_CHECK((reinterpret_cast<uintptr_t>(ws) & 255) == 0, "workspace must be 256-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    unsigned char* w = reinterpret_cast<unsigned char*>(ws);
    switch (g.T) {
        case 2048: return vfft_run<2048>(g, dem_norm, ld_in, norm, dir, ld_out, kernels, kern_off, kern_hw, n_kernels, w, s);
        case 4096: return vfft_run<4096>(g, dem_norm, ld_in, norm, dir, ld_out, kernels, kern_off, kern_hw, n_kernels, w, s);
        default: return vfft_run<8192>(g, dem_norm, ld_in, norm, dir, ld_out, kernels, kern_off, kern_hw, n_kernels, w, s);
    }
}

}  // extern "C"
