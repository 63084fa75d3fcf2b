"""Parity of the CUDA path (through the public API -> C ABI) against the CPU oracle and the golden
vectors produced by the reference itself.  Tolerances are the north star's:

    TPI / STD          max abs error <= 1e-3 m
    dx, dy, slope      <= 1e-4
    aspect, Sx         <= 1e-3 degrees (aspect: away from flat cells)
    NaN / border masks exact

Run on the GPU box:  python -m pytest tests -m gpu
"""

import numpy as np
import pytest

from oracle import oracle as O
from topo_descriptors_b200 import _xr, device as dev, helpers as hlp, topo
from topo_descriptors_b200.device import DeviceDEM
from topo_descriptors_b200.synth import dem_dataset, fractal_dem

pytestmark = pytest.mark.gpu

TOL_M = 1e-3
TOL_D = 1e-4
TOL_DEG = 1e-3


@pytest.fixture(scope="module")
def c1():
    """BASELINE config 1/2 DEM: 900 x 1440, 30 m, seed 0 (float and integer-valued variants)."""
    z = fractal_dem(900, 1440, seed=0)
    zi = fractal_dem(900, 1440, seed=0, integer=True)
    return z, zi, dem_dataset(z, res=30.0)


def maxdiff(a, b):
    return float(np.nanmax(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64))))


# ---------------------------------------------------------------------------------------------
# TPI
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("size", [3, 5, 6, 7, 17, 33])
def test_tpi_golden_tight(golden, size):
    got = topo.tpi(golden["in__zi"], size)
    assert got.dtype == np.float32 and got.shape == golden["in__zi"].shape
    assert maxdiff(got, golden[f"tpi_tight__{size}"]) <= 1e-4  # float32 rounding of the result only


@pytest.mark.parametrize("size", [3, 5, 6, 7, 17, 33])
def test_tpi_golden_literal(golden, size):
    got = topo.tpi(golden["in__z"], size)
    assert maxdiff(got, O.tpi_exact(golden["in__z"], size)) <= TOL_M
    assert maxdiff(got, golden[f"tpi_lit__{size}"]) <= 3e-3  # the reference's own float32-FFT noise


@pytest.mark.parametrize("size", [2, 3, 4, 5, 7, 17, 35, 67])
def test_tpi_c1(c1, size):
    z = c1[0]
    got = topo.tpi(z, size)
    assert maxdiff(got, O.tpi_exact(z, size)) <= TOL_M


def test_tpi_two_pass_large_disc():
    z = fractal_dem(600, 700, seed=5)
    for size in (201, 301):
        got = topo.tpi(z, size)
        assert maxdiff(got, O.tpi_exact(z, size)) <= TOL_M


def test_tpi_two_pass_even_size_and_ragged_width():
    z = fractal_dem(333, 517, seed=6)  # nx not a multiple of 4: unaligned rows
    for size in (17, 200, 6):
        assert maxdiff(topo.tpi(z, size), O.tpi_exact(z, size)) <= TOL_M


def test_tpi_negative_and_wide_range():
    z = fractal_dem(200, 300, seed=8, zmin=-420.0, zmax=8800.0)
    for size in (5, 17, 67):
        assert maxdiff(topo.tpi(z, size), O.tpi_exact(z, size)) <= TOL_M
    # range so wide that the quantised mode would be too coarse: exact two-plane mode (TPI_X)
    zw = fractal_dem(200, 300, seed=8, zmin=-2.0e5, zmax=3.0e6)
    want = O.tpi_exact(zw, 17)
    assert maxdiff(topo.tpi(zw, 17), want) <= 0.26  # float32 resolution of a 3e6 elevation is 0.25


def test_tpi_constant_dem_border_values():
    z = np.full((64, 80), 100.0, dtype=np.float32)
    got = topo.tpi(z, 17)
    assert np.all(got[8:-8, 8:-8] == 0.0)
    want = O.tpi_exact(z, 17)
    assert maxdiff(got, want) <= 1e-4
    assert abs(float(got[0, 0]) - float(want[0, 0])) <= 1e-4 and want[0, 0] > 60  # 64.3: zero padding


def test_tpi_plane_interior_is_zero():
    yy, xx = np.mgrid[:300, :400]
    z = (1000.0 + 0.75 * xx - 0.5 * yy).astype(np.float32)
    for size in (5, 17, 67, 201):
        m = size // 2
        got = topo.tpi(z, size)
        assert np.abs(got[m:-m, m:-m]).max() <= 2e-4


def test_tpi_nan_input_gives_all_nan_and_size_one():
    z = fractal_dem(50, 60, seed=2)
    zn = z.copy()
    zn[10, 10] = np.nan
    assert np.isnan(topo.tpi(zn, 5)).all()
    assert np.isnan(topo.tpi(z, 1)).all()
    assert np.isnan(topo.std(zn, 5)).all()


def test_tpi_sigma_and_types(golden):
    z = golden["in__z"]
    got = topo.tpi(z, 7, sigma=1.75)
    assert maxdiff(got, O.tpi_exact(z, 7, sigma=1.75)) <= TOL_M
    assert maxdiff(got, golden["tpi_lit_sigma__7_1.75"]) <= 3e-3
    out64 = topo.tpi(z.astype(np.float64), 7)
    assert out64.dtype == np.float64
    da = _xr.DataArray(z, ("y", "x"))
    out_da = topo.tpi(da, 7)
    assert isinstance(out_da, _xr.DataArray) and out_da.dims == ("y", "x")
    assert np.array_equal(out_da.values, topo.tpi(z, 7))


# ---------------------------------------------------------------------------------------------
# STD
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("size", [3, 5, 7, 17])
def test_std_golden(golden, size):
    got = topo.std(golden["in__zc"], size)
    assert got.dtype == np.float64
    assert maxdiff(got, O.std_exact(golden["in__zc"], size)) <= 1e-5
    assert maxdiff(got, golden[f"std_f64c__{size}"]) <= 1e-4
    got = topo.std(golden["in__zi"], size)
    assert maxdiff(got, O.std_exact(golden["in__zi"], size)) <= TOL_M
    assert maxdiff(got, golden[f"std_f64__{size}"]) <= 1e-2  # reference fed float64: its FFT noise
    # float DEM: the int32-truncation quirk is active and reproduced
    got = topo.std(golden["in__z"], size)
    assert maxdiff(got, O.std_exact(golden["in__z"], size)) <= TOL_M


@pytest.mark.parametrize("size", [2, 3, 4, 6, 7, 17, 35, 67])
def test_std_c1(c1, size):
    z, zi, _ = c1
    assert maxdiff(topo.std(zi, size), O.std_exact(zi, size)) <= TOL_M
    assert maxdiff(topo.std(z, size), O.std_exact(z, size)) <= TOL_M


def test_std_two_pass_and_flat():
    zi = fractal_dem(500, 600, seed=4, integer=True)
    assert maxdiff(topo.std(zi, 201), O.std_exact(zi, 201)) <= TOL_M
    z = fractal_dem(300, 380, seed=4)
    assert maxdiff(topo.std(z, 151), O.std_exact(z, 151)) <= TOL_M
    flat = np.full((80, 90), 1234.0, dtype=np.float32)
    got = topo.std(flat, 17)
    assert np.all(got[8:-8, 8:-8] == 0.0)
    assert maxdiff(got, O.std_exact(flat, 17)) <= TOL_M
    zneg = fractal_dem(120, 160, seed=3, zmin=-300.5, zmax=900.25)
    assert maxdiff(topo.std(zneg, 9), O.std_exact(zneg, 9)) <= TOL_M


def test_tpi_std_share_disc_sums():
    """tpi(size) + std(size) on an integer DEM share the T-plane disc sums: identical to the unshared calls."""
    zi = fractal_dem(420, 500, seed=14, integer=True)
    d = DeviceDEM(dev.to_device(zi))
    for size in (151, 201):
        t_ref = dev.tpi(d, size, share=False)
        s_ref = dev.std(d, size, share=False)
        t1 = dev.tpi(d, size)            # computes and keeps the sums
        assert d._tsum is not None
        s1 = dev.std(d, size)            # reuses them
        assert d._tsum is None
        s2 = dev.std(d, size)            # nothing cached: computes (and keeps) again
        t2 = dev.tpi(d, size)            # reuses
        for a, b in ((t1, t_ref), (t2, t_ref), (s1, s_ref), (s2, s_ref)):
            assert bool((a == b).all())
        assert maxdiff(s1.cpu().numpy(), O.std_exact(zi, size)) <= TOL_M


def test_disc_plane_cache_is_transparent():
    """A multi-scale sweep shares size-independent data -- the plane spectra of the FFT route, or the prefix planes and
    octagon tables of the walk (disc_fft off): identical to the unshared calls, in any order of sizes, for odd, even
    and fused sizes, whole images and row bands."""
    from topo_descriptors_b200 import _lib

    zi = fractal_dem(450, 520, seed=15, integer=True)
    plain = DeviceDEM(dev.to_device(zi))
    sizes = [151, 7, 201, 120, 301, 33]
    _lib.set_option("disc_fft", False)
    try:
        want = {s: (dev.tpi(plain, s, share=False), dev.std(plain, s, share=False)) for s in sizes}
        shared = DeviceDEM(dev.to_device(zi)).share_disc_planes(max(sizes))
        for s in sizes:
            assert bool((dev.tpi(shared, s) == want[s][0]).all()) and bool((dev.std(shared, s) == want[s][1]).all()), s
        assert shared._plane_cache is not None and shared._plane_cache[1].valid == 15
        shared.release_disc_planes()
        # the cached walk uses the octagon decomposition (diagonal tables); the square + caps walk must agree bit for bit
        _lib.set_option("octagon", False)
        try:
            square = DeviceDEM(dev.to_device(zi)).share_disc_planes(max(sizes))
            for s in (301, 151, 201):
                assert bool((dev.tpi(square, s) == want[s][0]).all()) and bool((dev.std(square, s) == want[s][1]).all()), s
        finally:
            _lib.set_option("octagon", True)
    finally:
        _lib.set_option("disc_fft", True)
    # default: the sweep caches the spectrum of the (T, Q) plane pair; every size from 33 up reuses it
    shared = DeviceDEM(dev.to_device(zi)).share_disc_planes(max(sizes))
    for s in sizes:
        assert bool((dev.tpi(shared, s) == want[s][0]).all()) and bool((dev.std(shared, s) == want[s][1]).all()), s
    assert shared._plane_cache is not None and shared._plane_cache[1].valid == 1 << 16
    shared.release_disc_planes()
    # a row band with enough halo for the largest size
    lo, hi, halo = 120, 300, max(sizes) // 2
    band = _band(plain.tensor, lo, hi, halo, plain.stats).share_disc_planes(max(sizes))
    for s in (201, 151, 301):
        assert bool((dev.std(band, s, lo, hi - lo) == want[s][1][lo:hi]).all()), s
        assert bool((dev.tpi(band, s, lo, hi - lo) == want[s][0][lo:hi]).all()), s
    # float DEMs share planes too, with the fixed-point scales of the largest size: within tolerance at every size
    z = fractal_dem(420, 460, seed=16)
    f = DeviceDEM(dev.to_device(z)).share_disc_planes(301)
    for s in (301, 41, 151, 88, 9):
        assert maxdiff(dev.tpi(f, s).cpu().numpy(), O.tpi_exact(z, s)) <= TOL_M, s
        assert maxdiff(dev.std(f, s).cpu().numpy(), O.std_exact(z, s)) <= TOL_M, s
    assert f._plane_cache is not None and f._plane_cache[1].valid != 0
    # a range so wide that the quantised plane is too coarse: exact two-plane TPI (T + fraction planes), shared as well
    zw = fractal_dem(300, 340, seed=17, zmin=-2.0e5, zmax=3.0e6)
    g = DeviceDEM(dev.to_device(zw)).share_disc_planes(151)
    for s in (151, 61):
        assert maxdiff(dev.tpi(g, s).cpu().numpy(), O.tpi_exact(zw, s)) <= 0.26, s  # float32 resolution of 3e6 is 0.25


def test_std_sigma(golden):
    got = topo.std(golden["in__zc"], 7, sigma=1.75)
    assert maxdiff(got, O.std_exact(golden["in__zc"], 7, sigma=1.75)) <= TOL_M


# ---------------------------------------------------------------------------------------------
# Gaussian / gradient / sobel
# ---------------------------------------------------------------------------------------------
def _ulp_close(got, want, max_frac=1e-3):
    """Same float32 value everywhere except rare one-ulp rounding flips (float64 summation order)."""
    got, want = np.asarray(got), np.asarray(want)
    bad = got != want
    assert bad.mean() <= max_frac, f"{bad.mean():.2e} of the pixels differ"
    if bad.any():
        assert np.all(np.abs(got[bad] - want[bad]) <= np.spacing(np.abs(want[bad])) * 1.01)


def test_gaussian_matches_scipy_semantics(golden):
    z = golden["in__z"]
    _ulp_close(topo.dem(z, 3.3), golden["dem__3.3"])
    _ulp_close(topo.dem(z, 20.0), golden["dem__20"])  # radius 80 > 64 rows: multiple reflections
    _ulp_close(topo.dem(z, (2.0, 7.5)), O.gaussian_filter_restated(z, (2.0, 7.5)))


@pytest.mark.parametrize("sigma", [1.75, 16.75, 50.0])
def test_gaussian_c1(c1, sigma):
    z = c1[0]
    _ulp_close(topo.dem(z, sigma), O.gaussian_filter_restated(z, sigma))


def test_gaussian_nan_spreads_exactly_like_scipy():
    z = fractal_dem(150, 170, seed=31)
    z[40, 50] = np.nan
    z[149, 0] = np.nan
    z[70:72, 100] = np.inf
    for sigma in (0.6, 1.75, 4.25, (2.0, 0.0), 20.0):
        want = O.gaussian_filter_restated(z, sigma)
        got = topo.dem(z, sigma)
        assert np.array_equal(np.isnan(got), np.isnan(want)), sigma
        assert np.array_equal(np.isinf(got), np.isinf(want)), sigma
        ok = np.isfinite(want)
        if ok.any():  # sigma 20: the radius-80 window already covers every pixel
            _ulp_close(got[ok], want[ok])


def test_sobel_bit_exact(golden):
    dx, dy = topo.sobel(golden["in__z"])
    assert np.array_equal(dx, golden["sobel__dx"]) and np.array_equal(dy, golden["sobel__dy"])


def _check_gradient(got, want, flat_thr=0.01):
    dx, dy, slope, aspect = got
    for a in got:
        assert a.dtype == np.float32
    assert maxdiff(dx, want[0]) <= TOL_D and maxdiff(dy, want[1]) <= TOL_D
    assert maxdiff(slope, want[2]) <= TOL_D
    steep = np.hypot(want[0], want[1]) > flat_thr
    d = np.abs(aspect.astype(np.float64) - want[3].astype(np.float64))
    d = np.minimum(d, 360.0 - d)
    assert d[steep].max() <= TOL_DEG
    return float(d.max())


@pytest.mark.parametrize("sigma,ratio", [(0.75, 1), (1.75, 1), (4.25, 1), (4.25, 1.5), (16.75, 1)])
def test_gradient_golden(golden, sigma, ratio):
    res = {"x": golden["scale_to_pixel__res_x"], "y": golden["scale_to_pixel__res_y"]}
    got = topo.gradient(golden["in__z"], sigma, res, sig_ratio=ratio)
    want = [golden[f"gradient__{sigma}_{ratio}_{nm}"] for nm in ("dx", "dy", "slope", "aspect")]
    _check_gradient(got, want)


def test_gradient_res2d_and_flat(golden):
    res = {"x": golden["gradient_res2d__x"], "y": golden["gradient_res2d__y"]}
    got = topo.gradient(golden["in__z"], 1.75, res)
    want = [golden[f"gradient_res2d__{nm}"] for nm in ("dx", "dy", "slope", "aspect")]
    _check_gradient(got, want)
    flat = np.full((16, 24), 512.25, dtype=np.float32)
    out = topo.gradient(flat, 1.75, {"x": np.full(24, 30.0), "y": np.full(16, -30.0)})
    assert np.array_equal(out[2], golden["gradient_flat__slope"])
    assert np.array_equal(out[3], golden["gradient_flat__aspect"])  # sign-of-zero convention: 0 deg


@pytest.mark.parametrize("sigma", [0.75, 1.75, 16.75])
def test_gradient_c2(c1, sigma):
    z, _, ds = c1
    _, res = hlp.scale_to_pixel([200], ds)
    _check_gradient(topo.gradient(z, sigma, res), O.gradient_exact(z, sigma, res))


def test_gradient_plane_is_exact():
    yy, xx = np.mgrid[:200, :260]
    z = (500.0 + 3.0 * xx + 1.5 * yy).astype(np.float32)
    res = {"x": np.full(260, 30.0), "y": np.full(200, -30.0)}
    for sigma in (0.5, 2.0):
        dx, dy, slope, aspect = topo.gradient(z, sigma, res)
        inner = (slice(20, -20), slice(20, -20))
        assert np.abs(dx[inner] - 0.1).max() <= 1e-5 and np.abs(dy[inner] + 0.05).max() <= 1e-5
        assert np.abs(slope[inner] - np.degrees(np.arctan(np.hypot(0.1, 0.05)))).max() <= 1e-4


# ---------------------------------------------------------------------------------------------
# Sx
# ---------------------------------------------------------------------------------------------
def test_sx_golden(golden):
    z, x, y = golden["in__z"], golden["in__x"], golden["in__y"]
    ds = _xr.Dataset({"alti": (("y", "x"), z)}, coords={"x": x, "y": y}, attrs={"crs": "epsg:2056"})
    for i, (az, rad, rmin, h, arc, steps) in enumerate(golden["sx__cases"]):
        got = topo.sx(ds, az, rad, height=h, azimuth_arc=arc, azimuth_steps=int(steps), radius_min=rmin)
        want = golden[f"sx__case{i}"]
        assert got.dtype == np.float32 and got.shape == want.shape
        assert maxdiff(got, want) <= TOL_DEG, i
        assert np.array_equal(got == 0, want == 0)  # the zero frame
    dsn = _xr.Dataset({"alti": (("y", "x"), golden["sx_nan__in"])}, coords={"x": x, "y": y}, attrs={"crs": "epsg:2056"})
    got, want = topo.sx(dsn, 45, 150), golden["sx_nan__out"]
    assert np.array_equal(np.isnan(got), np.isnan(want)) and maxdiff(got, want) <= TOL_DEG
    got, want = topo.sx(ds, 0, 150, radius_min=1000.0), golden["sx_allmasked__out"]
    assert np.array_equal(np.isnan(got), np.isnan(want)) and np.array_equal(got == 0, want == 0)


def test_sx_c3_crop_many_azimuths():
    z = fractal_dem(512, 640, seed=1)
    ds = dem_dataset(z, res=30.0)
    azs = list(range(0, 360, 45))
    got = topo.sx(ds, azs, 500.0)
    assert got.shape == (len(azs), 512, 640)
    x, y = ds["x"].values, ds["y"].values
    for k, az in enumerate(azs):
        want = O.sx_exact(z, x, y, az, 500.0)
        assert maxdiff(got[k], want) <= TOL_DEG, az
    assert np.array_equal(topo.sx(ds, 135, 500.0), got[3])


def test_sx_plane_analytic():
    yy, xx = np.mgrid[:128, :160]
    z = (100.0 + 6.0 * xx).astype(np.float32)  # rises 6 m per 30 m pixel towards the east
    ds = dem_dataset(z, res=30.0)
    got = topo.sx(ds, 90.0, 300.0, height=10.0, azimuth_arc=0.0)  # single ray due east, 10 px
    # samples at 1..9 px (the source pixel itself is excluded, topo.py:909); (6k-10)/(30k) grows with k
    want = np.degrees(np.arctan((6.0 * 9 - 10.0) / 270.0))
    assert np.abs(got[10:-10, 10:-10] - want).max() <= TOL_DEG
    assert not got[:10].any() and not got[:, -10:].any()


# ---------------------------------------------------------------------------------------------
# valley / ridge
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case,size,mode,flats,sigma", [
    ("valley7", 7, "valley", (0, 0.15, 0.3), None),
    ("ridge9", 9, "ridge", (0, 0.2, 0.4), 1.125),
    ("valley5f2", 5, "valley", (0, 0.3), None),
])
def test_valley_ridge_golden(golden, case, size, mode, flats, sigma):
    z = golden["in__z"]
    norm, direction = topo.valley_ridge(z, size, mode, list(flats), sigma)
    assert norm.dtype == np.float32 and direction.dtype == np.float32
    o_norm, o_dir, gap = O.valley_ridge_exact(z, size, mode, flats, sigma, return_gap=True)
    assert maxdiff(norm, o_norm) <= TOL_M
    assert maxdiff(norm, golden[f"valley_ridge__{case}_norm"]) <= TOL_M
    decidable = gap > 1e-2
    assert np.array_equal(direction[decidable], o_dir[decidable])
    assert np.array_equal(direction[decidable], golden[f"valley_ridge__{case}_dir"][decidable])
    assert (direction != o_dir).mean() < 0.01


def test_valley_ridge_larger_kernel_and_nan():
    z = fractal_dem(96, 128, seed=11)
    norm, direction = topo.valley_ridge(z, 21, "ridge")
    o_norm, o_dir, gap = O.valley_ridge_exact(z, 21, "ridge", return_gap=True)
    assert maxdiff(norm, o_norm) <= TOL_M
    assert np.array_equal(direction[gap > 1e-2], o_dir[gap > 1e-2])
    zn = z.copy()
    zn[3, 3] = np.nan
    norm, _ = topo.valley_ridge(zn, 7, "valley")
    assert not norm.any()  # reference: FFT spreads the NaN, nothing compares greater, clip(-inf) = 0


# ---------------------------------------------------------------------------------------------
# row bands: a band + halo must reproduce the whole-image pixels bit for bit
# ---------------------------------------------------------------------------------------------
def _band(full, lo, hi, halo, stats):
    a, b = max(0, lo - halo), min(full.shape[0], hi + halo)
    return DeviceDEM(full[a:b].contiguous(), gny=full.shape[0], gy0=a, stats=stats)


def test_bands_bit_identical():
    z = fractal_dem(301, 420, seed=12)
    whole = DeviceDEM(dev.to_device(z))
    stats = whole.stats
    cuts = [0, 97, 200, 301]
    for size in (7, 33, 151):
        ref_t = dev.tpi(whole, size)
        ref_s = dev.std(whole, size)
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            band = _band(whole.tensor, lo, hi, size // 2, stats)
            assert (dev.tpi(band, size, lo, hi - lo) == ref_t[lo:hi]).all()
            assert (dev.std(band, size, lo, hi - lo) == ref_s[lo:hi]).all()
    sigma = 4.25
    lw = dev.gauss_radius(sigma)
    ref_g = dev.gauss(whole, sigma, sigma)
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        band = _band(whole.tensor, lo, hi, lw, stats)
        assert (dev.gauss(band, sigma, sigma, lo, hi - lo) == ref_g[lo:hi]).all()


def test_cached_sweep_on_thin_bands_like_8_gpus():
    """The partition tests/mgpu_check.py uses at 8 ranks (1500 rows -> bands of ~187 rows, halo 150 of size 301),
    emulated on one GPU: every band runs the cached multi-size sweep and must equal the whole-image result."""
    from topo_descriptors_b200 import bands

    ny, nx = 1500, 1111
    zi = fractal_dem(ny, nx, seed=5, integer=True)
    whole = DeviceDEM(dev.to_device(zi))
    sizes = [5, 21, 67, 201, 301]
    want = {s: (dev.tpi(whole, s, share=False), dev.std(whole, s, share=False)) for s in sizes}
    halo = bands.sweep_halo(sizes, [])
    for rank in range(8):
        ctx = bands.BandContext(ny, nx, rank, 8)
        band = _band(whole.tensor, ctx.r0, ctx.r1, halo, whole.stats).share_disc_planes(max(sizes))
        for s in sizes:
            assert bool((dev.tpi(band, s, ctx.r0, ctx.rows) == want[s][0][ctx.r0 : ctx.r1]).all()), (rank, s)
            assert bool((dev.std(band, s, ctx.r0, ctx.rows) == want[s][1][ctx.r0 : ctx.r1]).all()), (rank, s)
        band.release_disc_planes()


def test_bands_valley_sx_bit_identical():
    """Row bands of valley_ridge and Sx equal the same rows of the whole-image result."""
    from topo_descriptors_b200 import bands

    z = fractal_dem(301, 420, seed=13)
    ds = dem_dataset(z, res=30.0)
    whole = DeviceDEM(dev.to_device(z))
    cuts = [0, 97, 200, 301]
    # valley / ridge through the band entry point (world 1 = whole image) and through band views
    ctx = bands.BandContext(301, 420)
    for sigma in (None, 1.5):
        want = topo.valley_ridge(z, 9, "valley", [0, 0.15, 0.3], sigma)
        got = bands.valley_ridge_band(whole.tensor, ctx, 9, "valley", [0, 0.15, 0.3], sigma)
        assert np.array_equal(got[0].cpu().numpy(), want[0]) and np.array_equal(got[1].cpu().numpy(), want[1])
    st = whole.stats
    mean = st["sum"] / st["n"]
    sd = np.sqrt(st["sumsq"] / st["n"] - mean * mean)
    normed = dev.zscore(whole, np.float32(mean), np.float32(sd))
    bank = topo._device_bank(9, "valley", [0, 0.15, 0.3], whole.tensor.device)
    ref_n, ref_d = dev.valley_ridge(normed, bank)
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        band = _band(normed.tensor, lo, hi, bank["hmax"] // 2, st)
        n, d = dev.valley_ridge(band, bank, lo, hi - lo)
        assert (n == ref_n[lo:hi]).all() and (d == ref_d[lo:hi]).all()
    # Sx
    plan = topo._sx_plan(ds, [0.0, 135.0, 270.0], 300.0, 10.0, 15, 0.0)
    ref = topo._sx_device(whole, plan, 10.0)
    assert (bands.sx_band(whole.tensor, ctx, plan, 10.0) == ref).all()
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        band = _band(whole.tensor, lo, hi, plan[3], None)
        assert (topo._sx_device(band, plan, 10.0, lo, hi - lo) == ref[:, lo:hi]).all()
    assert bands.azimuth_share(range(10), bands.BandContext(10, 4, 1, 4)) == [1, 5, 9]


# ---------------------------------------------------------------------------------------------
# device pre-stage: mask + NaN census + nearest fill (helpers.py:17-31, 137-154)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(40, 97), (33, 1000), (7, 31), (3, 4099)])
def test_fill_na_resident(shape):
    from topo_descriptors_b200 import prestage

    rng = np.random.default_rng(shape[1])
    z = rng.uniform(-50, 1000, shape).astype(np.float32)
    z[rng.uniform(size=shape) < 0.3] = np.nan
    z[1, :] = np.nan                      # a row without any valid cell stays NaN
    z[2, :] = np.nan
    z[2, shape[1] // 2] = 3.0             # a single valid cell fills its row
    z[0, : shape[1] * 2 // 3] = np.nan    # long gaps at both ends
    z[min(4, shape[0] - 1), shape[1] // 3 :] = np.nan
    base = dem_dataset(z, res=30.0)
    xs = base["x"].values
    warped = xs + 7.0 * np.sin(np.arange(xs.size))  # non-uniform but still ascending coordinates
    for x, thr in ((xs, None), (xs[::-1].copy(), None), (warped, None), (xs, 100.0)):
        ds = _xr.Dataset({"alti": (("y", "x"), z)}, coords={"x": x, "y": base["y"].values}, attrs=base.attrs)
        ind, filled = prestage.fill_na_resident(ds, mask_below=thr)
        want_ind, want = O.fill_na_exact(z, x, thr)
        assert np.array_equal(ind[0].cpu().numpy(), want_ind[0]) and np.array_equal(ind[1].cpu().numpy(), want_ind[1])
        got = hlp.get_da(filled).values
        assert isinstance(got, DeviceDEM)
        assert np.array_equal(got.numpy(), want, equal_nan=True)
    # nothing missing: identity and empty index tensors
    clean = np.nan_to_num(z, nan=5.0)
    ind, filled = prestage.fill_na_resident(dem_dataset(clean, res=30.0))
    assert ind[0].numel() == 0 and np.array_equal(hlp.get_da(filled).values.numpy(), clean)


def test_compute_driver_on_resident_dataset(tmp_path):
    """The script's pipeline with the DEM resident from the pre-stage on equals the host-marshalled one."""
    from topo_descriptors_b200 import prestage

    z = fractal_dem(120, 160, seed=22)
    z[5, 7] = np.nan
    z[60, 61:90] = np.nan
    ds = dem_dataset(z, res=30.0)
    ind_h, ds_h = hlp.fill_na(ds)
    ind_d, ds_d = prestage.fill_na_resident(ds)
    (tmp_path / "h").mkdir()
    (tmp_path / "d").mkdir()
    for ind, d, out in ((ind_h, ds_h, tmp_path / "h"), (ind_d, ds_d, tmp_path / "d")):
        topo.compute_tpi(d, [200, 500], ind_nans=ind, outdir=out)
        topo.compute_gradient(d, [200], ind_nans=ind, outdir=out)
        topo.compute_sx(d, 0, 150, outdir=out)
    for f in sorted((tmp_path / "h").iterdir()):
        with np.load(f) as a, np.load(tmp_path / "d" / f.name) as b:
            for k in a.files:
                assert np.array_equal(a[k], b[k], equal_nan=a[k].dtype.kind == "f"), (f.name, k)


# ---------------------------------------------------------------------------------------------
# compute_* drivers
# ---------------------------------------------------------------------------------------------
def test_compute_drivers(tmp_path, c1):
    z = fractal_dem(120, 160, seed=21)
    zn = z.copy()
    zn[5, 7] = np.nan
    zn[60, 61:64] = np.nan
    ind_nans, ds = hlp.fill_na(dem_dataset(zn, res=30.0))
    zf = hlp.get_da(ds).values
    x, y = ds["x"].values, ds["y"].values
    crop = {"x": slice(x[10], x[149]), "y": slice(y[8], y[111])}
    scales = [200, 500]
    px, res = hlp.scale_to_pixel(scales, ds)

    topo.compute_tpi(ds, scales, ind_nans=ind_nans, crop=crop, outdir=tmp_path)
    topo.compute_std(ds, scales, smth_factors=[None, 0.5], ind_nans=ind_nans, outdir=tmp_path)
    topo.compute_gradient(ds, scales, ind_nans=ind_nans, outdir=tmp_path)
    topo.compute_dem(ds, 200, ind_nans=ind_nans, outdir=tmp_path)
    topo.compute_valley_ridge(ds, 200, "valley", ind_nans=ind_nans, outdir=tmp_path)
    topo.compute_sx(ds, 0, 150, outdir=tmp_path)

    with np.load(tmp_path / "topo_TPI_200M.npz") as f:
        want = O.tpi_exact(zf, px[0])
        want[ind_nans] = np.nan
        want = want[8:112, 10:150]
        got = f["TPI_200M"]
        assert got.shape == want.shape and got.dtype == np.float32
        assert np.array_equal(np.isnan(got), np.isnan(want)) and maxdiff(got, want) <= TOL_M
        assert str(f["units"]) == "m"
    with np.load(tmp_path / "topo_STD_500M_SMTHFACT0.5.npz") as f:
        got = f["STD_500M_SMTHFACT0.5"]
        want = O.std_exact(zf, px[1], sigma=0.5 * px[1] / 4)
        want[ind_nans] = np.nan
        assert got.dtype == np.float64 and np.isnan(got[5, 7]) and maxdiff(got, want) <= TOL_M
    names = sorted(p.name for p in tmp_path.iterdir())
    for expect in ("topo_WE_DERIVATIVE_200M_SIGRATIO1.npz", "topo_SN_DERIVATIVE_500M_SIGRATIO1.npz",
                   "topo_SLOPE_200M_SIGRATIO1.npz", "topo_ASPECT_500M_SIGRATIO1.npz", "topo_DEM_200M.npz",
                   "topo_VALLEY_NORM_200M.npz", "topo_VALLEY_DIR_200M.npz", "topo_SX_RADIUS150_AZIMUTH0.npz",
                   "topo_STD_200M.npz", "topo_TPI_500M.npz"):
        assert expect in names, (expect, names)
    with np.load(tmp_path / "topo_SLOPE_200M_SIGRATIO1.npz") as f:
        want = O.gradient_exact(zf, px[0] / 4, res)[2]
        want[ind_nans] = np.nan
        assert maxdiff(f["SLOPE_200M_SIGRATIO1"], want) <= TOL_D and str(f["units"]) == "degree"
    with np.load(tmp_path / "topo_SX_RADIUS150_AZIMUTH0.npz") as f:
        assert maxdiff(f["SX_RADIUS150_AZIMUTH0"], O.sx_exact(zf, x, y, 0, 150)) <= TOL_DEG


# ---------------------------------------------------------------------------------------------
# BASELINE full sizes: size-independent properties
# ---------------------------------------------------------------------------------------------
def test_full_size_properties_16384():
    """Config 4 shape (16384^2): plane => TPI 0 in the interior (fused and two-pass discs), exact
    gradient; band == whole for a strip; constant => STD 0."""
    import torch

    n = 16384
    xx = torch.arange(n, device="cuda", dtype=torch.float32)
    plane = (500.0 + 0.5 * xx[None, :] + 0.25 * xx[:, None]).contiguous()  # exactly representable
    d = DeviceDEM(plane, stats={"min": 500.0, "max": 500.0 + 0.75 * (n - 1), "nonfinite": 0, "nonint": 1,
                                "sum": 0.0, "sumsq": 0.0, "n": n * n})
    for size in (21, 801):
        m = size // 2
        out = dev.tpi(d, size)
        assert float(out[m:-m, m:-m].abs().max()) <= 1e-3
        strip = dev.tpi(_band(plane, 5000, 5064, m, d.stats), size, 5000, 64)
        assert bool((strip == out[5000:5064]).all())
        del out, strip
    res = torch.full((n,), 25.0, dtype=torch.float64, device="cuda")
    dx, dy, slope, aspect = dev.gradient_from_smooth(d, d, res, 0, -res, 0)
    assert float((dx[1:-1, 1:-1] - 0.02).abs().max()) <= 1e-6
    assert float((dy[1:-1, 1:-1] + 0.01).abs().max()) <= 1e-6
    del dx, dy, slope, aspect
    const = torch.full((n, n), 777.0, device="cuda")
    dc = DeviceDEM(const, stats={"min": 777.0, "max": 777.0, "nonfinite": 0, "nonint": 0, "sum": 0.0, "sumsq": 0.0,
                                 "n": n * n})
    out = dev.std(dc, 401)
    assert float(out[200:-200, 200:-200].abs().max()) == 0.0
