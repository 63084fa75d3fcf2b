/*
 * topo_b200.h  --  C ABI of libtopo_b200.so, the sm_100a implementation of the raster-filter hot
 * path of MeteoSwiss/topo-descriptors.
 *
 * The reference has no FFI: its boundary is the Python function API of topo_descriptors/topo.py.
 * Each entry point below replaces the *array kernel* behind one of those functions; the host shim
 * (topo_descriptors_b200/topo.py) keeps the reference signatures and calls these through ctypes.
 * Citations are reference file:line (topo.py = topo_descriptors/topo.py).
 *
 * Conventions
 *  - All image pointers are DEVICE pointers to row-major float32 rasters; `ld*` = row pitch in
 *    elements.  Nothing here allocates: scratch is passed in (`ws`, size from the *_workspace_bytes
 *    twin).  Launches are asynchronous on `stream` (a cudaStream_t passed as void*).
 *  - Return value: 0 = ok, < 0 = error; topo_last_error() returns a thread-local message.  Nothing
 *    throws across the ABI.
 *  - Row bands (multi-GPU sharding): every stencil takes a *view*.  The input
 *    buffer holds global rows [in_gy0, in_gy0+in_rows) of an image that is `gny` rows tall; the
 *    call writes global rows [out_gy0, out_gy0+out_rows) to `out` (row 0 of `out` = out_gy0).
 *    Border rules (zero padding / reflect / one-sided differences / the Sx frame) are always
 *    applied in GLOBAL coordinates, so a band + halo gives bit-identical pixels to the full image.
 *    The caller guarantees the input band covers every in-image row the stencil touches.
 *    Whole image on one GPU: in_gy0 = out_gy0 = 0, in_rows = out_rows = gny.
 */
#ifndef TOPO_B200_H
#define TOPO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct topo_view {
    int nx;       /* columns */
    int gny;      /* rows of the GLOBAL image */
    int in_gy0;   /* global row of input row 0 */
    int in_rows;  /* rows present in the input buffer */
    int out_gy0;  /* global row of output row 0 */
    int out_rows; /* rows to compute */
} topo_view;

/* Planes shared by the tpi / std calls of ONE DEM band at several sizes (a multi-scale sweep): the integer planes
 * (trunc(z) - tmin, its centred square, and for float DEMs the fraction / the quantised elevation with the
 * fixed-point scale of max_size) do not depend on the disc size, so they are built by the first call that needs them
 * and reused by the later ones.  When max_size takes the FFT route (the default from size 128) what is kept is the
 * float64 SPECTRUM of the tiled planes (one slot per plane pair) plus the mask spectrum of the size in flight;
 * otherwise the prefix planes with the column-prefix, summed-area and diagonal tables of the octagon walk.
 * The caller owns `mem` (DEVICE, 256-byte aligned,
 * >= topo_disc_cache_bytes(v, max_size, all_integer, zmin, zmax)), starts with valid = 0 and passes the same struct, DEM,
 * view, all_integer flag and range to every call; the input band must cover the halo of max_size.
 * NULL (or mem = NULL) = no sharing. */
typedef struct topo_disc_cache {
    void* mem;
    size_t bytes;
    int max_size; /* the planes are laid out for discs up to this size */
    int valid;    /* in/out: bits 2k / 2k+1 = row prefix / column-side tables of plane kind k
                     (0 trunc(z) - tmin, 1 its square (or the low 16 bits of a split square), 2 fraction,
                     3 quantised elevation; the high half of a split square takes the next free kind);
                     bits 16 / 17: the spectrum of plane pair 0 / 1 (FFT route) */
    int mask_size; /* in/out, FFT route: the disc size whose mask spectrum sits behind the plane spectra (tpi(size) and
                      std(size) of a pair build it once); start with 0 */
} topo_disc_cache;

/* ---- library ------------------------------------------------------------------------------ */
int topo_version(void);
const char* topo_last_error(void);
/* Number of kernel launches issued through this library by the calling process (bench.py's
 * `gpu_launches` claim). */
long long topo_launch_count(void);
/* Per-kernel timing for bench.py: while enabled, every launch is bracketed by CUDA events on its
 * stream.  topo_profile_dump synchronises on them, writes one "<kernel> <ms>" line per launch (in
 * launch order) into buf and clears the records. */
int topo_profile_enable(int on);
int topo_profile_dump(char* buf, size_t cap);
/* Execution-shape switches, all on by default: "octagon" (octagon core of the shared-plane disc walk), "tiny"
 * (register sliding sums for sizes 5..13), "sx_tma" (TMA-staged Sx tile), "gauss_fft" (float64 FFT overlap-save for
 * wide Gaussian radii), "grad_fused" (single-kernel small-radius gradient), "disc_fft" (exact disc sums of sizes >= 128
 * by float64 FFT convolution of the integer planes instead of the prefix-plane walk).  Every setting gives the same results
 * through another kernel shape (the tests flip them to compare shapes bit for bit); nothing is read from the
 * process environment.  Returns 0, or -1 for an unknown name. */
int topo_set_option(const char* name, int value);
/* Measurement aid for bench.py: launches 148 x 8 CTAs of 8 independent DFMA chains x iters per thread on `stream`
 * (scratch: DEVICE, 1 double, never written) and returns the float64 flops of the launch in *flops (HOST); timed by
 * the caller with CUDA events = the FP64 pipe rate the wide Gaussian is compared with. */
int topo_probe_dfma(int iters, double* scratch, double* flops, void* stream);

/* ---- DEM statistics (one pass, cached by the host per uploaded DEM) ------------------------ */
/* out_stats (DEVICE, 8 doubles): [0] min [1] max [2] #non-finite [3] #non-integer-valued
 * [4] sum [5] sum of squares [6] n [7] reserved.  Deterministic two-stage reduction.
 * Feeds: the fixed-point scale of the disc sums, the all-NaN rule of the FFT-based reference
 * functions (topo.py:175,301-302,443), and valley_ridge's global z-score (topo.py:429). */
size_t topo_dem_stats_workspace_bytes(int rows, int nx);
int topo_dem_stats_f32(const float* dem, int rows, int nx, int64_t ld, double* out_stats, void* ws,
                       size_t ws_bytes, void* stream);

/* out[i] = value (NaN re-stamp of whole outputs, the reference's FFT behaviour on NaN input). */
int topo_fill_f32(float* out, int rows, int nx, int64_t ld, float value, void* stream);
/* out[rows[i], cols[i]] = value : `array[ind_nans] = np.nan` of the compute_* drivers
 * (topo.py:57,139,267,385,591), on device. */
int topo_stamp_f32(float* out, int64_t ld, const int* rows, const int* cols, int64_t n, float value,
                   void* stream);

/* ---- pre-stage (helpers.py:17-31, 137-154): mask + NaN census + nearest-neighbour fill along x -------
 * A cell is MISSING when it is NaN, or -- with use_mask -- when it is <= mask_below
 * (`dem.where(dem > CFG.min_elevation)`, helpers.py:31).
 * topo_fill_na_f32: out = dem with every missing cell replaced by the valid cell of its row nearest in x
 *   (`interpolate_na(dim="x", method="nearest", fill_value="extrapolate")`, helpers.py:152-154: scipy
 *   interp1d "nearest" -- exact half-way points take the lower-x neighbour, cells beyond the first / last
 *   valid one take that one, rows without a valid cell stay NaN).  x (DEVICE, optional): the nx x-coordinates;
 *   NULL = uniform ascending grid.  row_missing (DEVICE, optional, `rows` ints): missing cells per row.
 * topo_nan_indices_f32: `ind_nans = np.where(np.isnan(dem))` (helpers.py:150) in row-major order.
 *   Call 1 (out_rows = out_cols = NULL): row_offsets[0..rows] = exclusive prefix sum of row_missing
 *   (row_offsets[rows] = total: the caller reads it and sizes the outputs).  Call 2: fills out_rows / out_cols. */
int topo_fill_na_f32(const float* dem, int64_t ld_in, float* out, int64_t ld_out, int rows, int nx,
                     const double* x, int use_mask, float mask_below, int* row_missing, void* stream);
int topo_nan_indices_f32(const float* dem, int64_t ld_in, int rows, int nx, int use_mask, float mask_below,
                         const int* row_missing, int64_t* row_offsets, int* out_rows, int* out_cols,
                         void* stream);

/* ---- TPI (topo.py:144-181) and STD (topo.py:272-307): disc sums ------------------------------
 * Zero-padded "same" convolution with circular_kernel(size) (topo.py:191-213; a square for
 * size < 5), scipy centring for even sizes.  Per-row prefix sums in 32-bit fixed point
 * (wrap-around arithmetic, so every span difference is exact) + one span difference per kernel
 * row, accumulated in 64-bit integers; float64 epilogue.
 *   TPI: out = dem - (disc_sum - excluded_centre) / (N - 1)
 *   STD: out = sqrt(max((sum trunc(x)^2 - (sum x)^2 / N) / (N - 1), 0))      (float32 result; the
 *        host up-casts to float64 like the reference).  trunc(x) reproduces
 *        `dem.astype("int32") ** 2` (topo.py:300): the squares use the truncated elevation.
 * zmin/zmax: GLOBAL finite range of the DEM (from topo_dem_stats_f32); all_integer: 1 if every
 * value is integral (SRTM-like DEMs: exact one-plane TPI, two-plane STD).
 * Sizes 5..13 (odd) use register sliding sums, small sizes run fused (tile + halo prefix in shared
 * memory), larger sizes run two passes through `ws` (prefix planes in HBM, gathered with 64-bit loads;
 * odd discs as an inscribed square / octagon from summed-area tables + caps).  cache: see
 * topo_disc_cache above -- with it the planes live in the cache and `ws` only holds raw plane sums;
 * pass the cache's max_size to the *_workspace_bytes / shares_tsum queries (0 without a cache). */
/* The queries take the same all_integer / zmin / zmax / cache size as the call they describe and make the same plan,
 * so the answer is exactly what the call needs (0 for the fused shapes).  STD on a DEM whose size x range^2 would
 * overflow the 32-bit span sums of the squares (e.g. a 0..4800 m range from size ~800) splits the square plane in
 * 16-bit halves: one more gather pass, same exact result. */
size_t topo_disc_workspace_bytes(const topo_view* v, int size, int what /*0 tpi, 1 std*/, int all_integer,
                                 double zmin, double zmax,
                                 int cache_max_size /* max_size of the topo_disc_cache that will be passed, else 0 */,
                                 int tsum_op /* the tsum_op the call will pass */);
/* tpi(size) and std(size) of one DEM share raw plane sums: trunc(z) - tmin for integer-valued DEMs (returns 1), and
 * for float DEMs also the fraction plane (returns 2; tpi then runs as the exact two-plane TPI_X rather than the
 * one-plane quantised TPI_Q, so a tpi + std pair walks 3 planes instead of 4).  When the return value n is > 0 the
 * first call can keep the sums (tsum_op = 1, tsum = n * out_rows * nx uint64 on the DEVICE) and the second reuse them
 * (tsum_op = 2).  0: no sharing for this size / DEM; tsum_op = 0: off. */
int topo_disc_shares_tsum(const topo_view* v, int size, int all_integer, double zmin, double zmax,
                          int cache_max_size /* 0: no plane cache */);
/* 0: this DEM / size cannot share planes (the calls then run un-cached; pass cache = NULL). */
size_t topo_disc_cache_bytes(const topo_view* v, int max_size, int all_integer, double zmin, double zmax);
/* Host-only introspection (no launch): the plan topo_tpi_f32 (what = 0) / topo_std_f32 (what = 1) would run.
 * info[20]: 0 mode (0 TPI_Q, 1 TPI_X, 2 STD_I, 3 STD_F, 4 TPI_I), 1 fused, 2 hybrid, 3 tiny, 4 uses the plane
 * cache, 5 octagon walk, 6 u (octagon) or a (inscribed square), 7 v, 8 corner diagonals, 9 32-bit accumulator
 * mask, 10 dynamic shared memory of the walk, 11 plane halo, 12 plane pitch, 13 plane rows, 14 workspace
 * bytes, 15 bytes of one cached plane region, 16 split square planes, 17 FFT route, 18 its transform length, 19 its
 * tiles. */
int topo_disc_plan_info(const topo_view* v, int size, int what, int all_integer, double zmin, double zmax,
                        int cache_max_size, int tsum_op, long long* info);
int topo_tpi_f32(const float* dem, int64_t ld_in, float* out, int64_t ld_out, const topo_view* v,
                 int size, int all_integer, double zmin, double zmax, unsigned long long* tsum,
                 int tsum_op, topo_disc_cache* cache, void* ws, size_t ws_bytes, void* stream);
int topo_std_f32(const float* dem, int64_t ld_in, float* out, int64_t ld_out, const topo_view* v,
                 int size, int all_integer, double zmin, double zmax, unsigned long long* tsum,
                 int tsum_op, topo_disc_cache* cache, void* ws, size_t ws_bytes, void* stream);

/* ---- Gaussian smoothing (topo.py:62-80; scipy.ndimage.gaussian_filter semantics) -------------
 * Separable, radius int(4*sigma+0.5), float64 weights and accumulation, axis 0 then axis 1 with a
 * float32 round in between, reflect borders (d c b a | a b c d | d c b a) in global coordinates.
 * w_y / w_x: DEVICE arrays of lw+1 float64 half-kernels (w[0] = centre tap) computed by the host
 * exactly as scipy's _gaussian_kernel1d does; a null pointer skips that axis (sigma <= 1e-15).
 * nan_safe != 0: the input may hold non-finite values; the kernels then predicate every tap so that NaN
 * spreads exactly +-lw like scipy (needed because the register-blocked walk also visits zero-weight taps).
 * `ws` holds the axis-0 result (+ two transposed planes when the axis-1 radius exceeds 64). */
size_t topo_gauss_workspace_bytes(const topo_view* v, int lw_y, int lw_x);
int topo_gauss_f32(const float* in, int64_t ld_in, float* out, int64_t ld_out, const topo_view* v,
                   const double* w_y, int lw_y, const double* w_x, int lw_x, int nan_safe, void* ws,
                   size_t ws_bytes, void* stream);

/* ---- gradient / slope / aspect (topo.py:597-644, 688-712) -------------------------------------
 * From smoothed surfaces gx (differentiated along x) and gy (along y; same pointer when
 * sig_ratio == 1): numpy.gradient central differences (one-sided at the global edges), division
 * by the signed grid resolution, slope = atan(hypot) in degrees, aspect = (180 + deg(atan2(dx,
 * dy))) mod 360 -- all in float32 in the reference's operation order.  res_x: (nx) or, if
 * res_x_2d, (gny, nx) float64 addressed by GLOBAL row; res_y: (gny) or (gny, nx). */
int topo_grad_from_smooth_f32(const float* gx, const float* gy, int64_t ld_in, float* dx, float* dy,
                              float* slope, float* aspect, int64_t ld_out, const topo_view* v,
                              const double* res_x, int res_x_2d, const double* res_y, int res_y_2d,
                              const float* res_xf, const float* res_yf, void* stream);
/* The whole isotropic gradient (sig_ratio == 1, topo.py:630-631 + 637-642) from the raw DEM: Gaussian axis 0 ->
 * float32 -> axis 1 -> float32 -> numpy.gradient -> resolution -> slope / aspect.  Radii up to 21 px run as ONE
 * kernel that keeps the tile, the axis-0 result and the smoothed tile in shared memory (20 B/px of HBM traffic
 * instead of 36; same taps in the same order: bit-identical to the separate kernels); wider radii, NaN-exact
 * smoothing (nan_safe) and "grad_fused" switched off run topo_gauss_f32 + topo_grad_from_smooth_f32 inside `ws`
 * (>= topo_gradient_workspace_bytes, 256-byte aligned).  w: DEVICE half kernel of lw + 1 float64 taps.  The band
 * must cover rows out_gy0 - lw - 1 .. out_gy0 + out_rows + lw (reflected at the global edges). */
size_t topo_gradient_workspace_bytes(const topo_view* v, int lw);
int topo_gradient_f32(const float* dem, int64_t ld_in, float* dx, float* dy, float* slope, float* aspect,
                      int64_t ld_out, const topo_view* v, const double* w, int lw, int nan_safe,
                      const double* res_x, int res_x_2d, const double* res_y, int res_y_2d, const float* res_xf,
                      const float* res_yf, void* ws, size_t ws_bytes, void* stream);
/* Sobel branch, sigma <= 1 (topo.py:628-629, 658-685): ndimage.convolve with K/8 and K.T/8,
 * reflect borders, float64 accumulation; same normalisation / slope / aspect epilogue fused.
 * res_xf / res_yf (both entry points, optional): float32 copies of the resolution arrays, to be passed
 * only when every value is exactly representable in float32 -- the float32 division then rounds
 * identically to numpy's float64 one and the kernels can take their 128-bit vector path. */
int topo_sobel_gradient_f32(const float* dem, int64_t ld_in, float* dx, float* dy, float* slope,
                            float* aspect, int64_t ld_out, const topo_view* v, const double* res_x,
                            int res_x_2d, const double* res_y, int res_y_2d, const float* res_xf,
                            const float* res_yf, int normalize, void* stream);

/* ---- Sx (topo.py:775-858, 928-953) --------------------------------------------------------------
 * n_az azimuth sectors in one launch.  offsets: (dy, dx) int pairs of the de-duplicated ray
 * samples of all sectors, inv_dist: 1/distance per sample (float32), az_begin[n_az+1]: CSR ranges
 * (all DEVICE arrays).  dy_min..dx_max: extreme offsets over all samples (band check, box size).
 * When a 128 x 16 tile plus the sample extents fits a 256 x 256 TMA box (and 150 KB), the DEM tile is
 * staged in shared memory by one cp.async.bulk.tensor per CTA and all azimuths of a CTA scan it from
 * shared memory; otherwise the samples are gathered through L1.  The angle is atanf in float32.
 * out: [n_az] planes `az_stride` elements apart.  Pixels within `window` of any GLOBAL edge are 0;
 * NaN samples are skipped, a NaN centre or an empty sample list gives NaN (np.nanmax semantics). */
int topo_sx_f32(const float* dem, int64_t ld_in, float* out, int64_t ld_out, int64_t az_stride,
                const topo_view* v, const int* offsets, const float* inv_dist, const int* az_begin,
                int n_az, int window, float height, int dy_min, int dy_max, int dx_min, int dx_max,
                void* stream);

/* ---- valley / ridge (topo.py:389-453) ------------------------------------------------------------
 * z-score with the global mean/std (topo.py:429), then ALL angles in one launch with a running
 * strict-'>' max and argmax over the angles, norm clipped at 0 (topo.py:441-452).
 * bank (DEVICE): for angle a, at element offset bank_off[a] (multiple of 4), a [w][hp][4] float32
 * array: kernel column j, row i, 4 channel slots (n_ch used, rest 0); rows h..hp-1 are zero, hp is
 * a multiple of 4 and >= h + 3.  The kernels are already channel-mixed (the reference's 3-D
 * convolution sums neighbouring flat-list kernels, topo.py:431,443) and flipped, so the device
 * correlates:  out[y,x] = sum Kf[i,j] * d[y + i - h/2, x + j - w/2], zero outside the image.
 * bank_hw (DEVICE): 4 ints per angle: h, w, hp, c0.  hmax/wmax: maxima of h and w over the angles.
 * bank_cols (DEVICE): for column j of angle a, at int index 2 * (c0 + j): (lo, n) -- the rows lo .. lo+n-1
 * of that kernel column are walked, lo and n multiples of 4, lo + n <= hp, every weight outside
 * [lo, lo + n - 4] exactly zero (the corners of a rotated kernel: ~37% of the bounding boxes).
 * dir receives the angle index (degrees, angles are 0..n_angles-1).
 * Flat lists longer than 4 (the reference accepts any length, topo.py:389-396): the host packs the channels in
 * banks of <= 4 and calls once per bank with group_flags bit 0 = "norm / dir hold the raw running (max, argmax) of
 * the previous banks" and bit 1 = "leave them raw for the next bank" (0 for a single bank); ties between banks
 * resolve to the lower angle, which is the reference's strict '>' over the angles in order. */
/* Device-side rotation of the kernel bank (topo.py:521-531 + the channel mixing of topo.py:431,443) for the FFT route:
 * coef (DEVICE): the n_kernels source kernels (h x w each) after scipy's quadratic-spline prefilter, float64;
 * angles (DEVICE): n_angles records of 64 bytes {double m00, m01, m10, m11, off0, off1; int64 out_off; int32 oh, ow}
 * = rotation matrix, offset and output box of scipy.ndimage.rotate(reshape=True) for each angle (computed by the host
 * with scipy's own cosdg / sindg) and the element offset of the angle's [n_kernels][oh][ow] block.  scratch and out
 * (DEVICE): as many floats as the blocks take.  out receives the rotated (order-2 spline, exactly scipy's arithmetic),
 * masked (cval), z-scored and channel-mixed kernels in scipy's orientation -- the `kernels` of
 * topo_valley_ridge_fft_f32. */
int topo_rotate_bank_f32(const double* coef, int n_kernels, int h, int w, const void* angles, int n_angles, float cval,
                         float* scratch, float* out, void* stream);
/* Large kernels: the same bank applied by 2-D overlap-save FFT convolution (float64 transforms in shared memory,
 * two real kernels per complex transform, spectra of the image tiles computed once): the cost per (angle, channel)
 * kernel is one inverse 2-D transform per tile, whatever the kernel size -- like the reference's own
 * signal.convolve, which takes its FFT method (topo.py:443).  kernels (DEVICE): the channel-mixed kernels in scipy's
 * orientation (true convolution, NOT flipped), row-major, kernel i at element offset kern_off[i]; kern_off / kern_hw
 * (HOST): per kernel its offset and (h, w, angle index), in angle order; hmax / wmax: maxima over the bank (up to 4096).
 * ws: DEVICE, 256-byte aligned, >= the query (3 x tiles x T^2 x 16 B, T = 2048 / 4096 / 8192). */
size_t topo_valley_ridge_fft_workspace_bytes(const topo_view* v, int hmax, int wmax);
int topo_valley_ridge_fft_f32(const float* dem_norm, int64_t ld_in, float* norm, float* dir, int64_t ld_out,
                              const topo_view* v, const float* kernels, const long long* kern_off, const int* kern_hw,
                              int n_kernels, int hmax, int wmax, void* ws, size_t ws_bytes, void* stream);
int topo_zscore_f32(const float* dem, int64_t ld_in, float* out, int64_t ld_out, int rows, int nx,
                    float mean, float std, void* stream);
int topo_valley_ridge_f32(const float* dem_norm, int64_t ld_in, float* norm, float* dir, int64_t ld_out,
                          const topo_view* v, const float* bank, const int* bank_hw,
                          const int64_t* bank_off, const int* bank_cols, int n_angles, int n_ch,
                          int hmax, int wmax, int group_flags, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TOPO_B200_H */
