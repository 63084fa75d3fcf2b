"""Host logic of the drop-in boundary (no GPU): helpers, Sx geometry, valley/ridge kernel bank,
Dataset duck-types, output names, and the C-ABI library's symbol table."""

import os
import re
import sys

import numpy as np
import pytest

from oracle import oracle as O
from topo_descriptors_b200 import CFG, _geometry as geo, _lib, _xr, helpers as hlp, topo
from topo_descriptors_b200.synth import dem_dataset, fractal_dem

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- the reference's own known-answer tests against the new host code (test/test_topo.py, test_helpers.py)
def test_sx_distance():
    out = topo._sx_distance(150.0, 50.0, 40.0)
    expected = np.array([256.1249695, 219.31712199, 188.67962264, 167.63054614, 160.0,
                         167.63054614, 188.67962264, 219.31712199, 256.1249695])
    assert np.all(np.isclose(out[0, :], expected))
    assert out.dtype == np.float64


def test_sx_bresenhamlines():
    out = topo._sx_bresenhamlines(np.array([[8, 9], [17, 22]]), np.array([15, 15]))
    expected = np.array([[9, 10], [10, 11], [11, 12], [12, 12], [13, 13], [14, 14],
                         [17, 21], [16, 20], [16, 19], [16, 18], [16, 17], [15, 16]])
    assert np.all(out == expected)
    assert out.dtype == np.int64


def test_sx_source_idx_delta():
    out = topo._sx_source_idx_delta(np.array([3.0, 4.0, 5.0, 6.0]), 500, 20, 30)
    assert np.all(out == np.array([[17, 1], [17, 2], [17, 2], [17, 3]]))
    assert out.dtype == np.int64


def test_round_up_to_odd():
    out = hlp.round_up_to_odd(np.arange(0.1, 10, 0.7))
    assert out.dtype == np.int64
    assert all(a == b for a, b in zip(out, [1, 1, 1, 3, 3, 3, 5, 5, 5, 7, 7, 7, 9, 9, 9]))


# ---- golden vectors (reference outputs) --------------------------------------------------------
def test_helpers_against_golden(golden):
    z = golden["in__z"]
    ds = _xr.Dataset({"alti": (("y", "x"), z)}, coords={"x": golden["in__x"], "y": golden["in__y"]},
                     attrs={"crs": "epsg:2056"})
    px, res = hlp.scale_to_pixel(list(golden["scale_to_pixel__scales"]), ds)
    assert px.dtype == np.int64 and np.array_equal(px, golden["scale_to_pixel__px"])
    assert np.array_equal(res["x"], golden["scale_to_pixel__res_x"])
    assert np.array_equal(res["y"], golden["scale_to_pixel__res_y"])
    sig = hlp.get_sigmas([None, 0.5, 0, 1, 2.5], px)
    want = golden["get_sigmas__out"]
    for s, w in zip(sig, want):
        assert (s is None and np.isnan(w)) or s == w
    assert np.array_equal(hlp.round_up_to_odd(golden["round_up_to_odd__in"]), golden["round_up_to_odd__out"])


@pytest.mark.parametrize("size", [3, 4, 5, 6, 7, 17, 33])
def test_circular_kernel(golden, size):
    k = topo.circular_kernel(size)
    assert k.dtype == np.float32 and np.array_equal(k, golden[f"circular_kernel__{size}"])


def test_sx_geometry_against_golden(golden):
    assert np.array_equal(geo.sx_distance(150.0, 30.0, -30.0), golden["sx_distance__150_30_-30"])
    assert np.array_equal(geo.sx_source_idx_delta(np.linspace(-5, 5, 15), 150.0, 30.0, -30.0),
                          golden["sx_source_idx_delta__b"])
    assert np.array_equal(geo.sx_bresenhamlines(np.array([[8, 9], [17, 22]]), np.array([15, 15])),
                          golden["sx_bresenhamlines__a"])


def test_sx_lines_match_oracle_on_random_sectors():
    rng = np.random.default_rng(3)
    for _ in range(40):
        radius = float(rng.integers(60, 900))
        dx = float(rng.choice([20.0, 25.0, 30.0, 50.0]))
        dy = -float(rng.choice([20.0, 25.0, 30.0, 40.0]))
        az = float(rng.uniform(0, 360))
        arc = float(rng.choice([0.0, 10.0, 30.0]))
        steps = int(rng.integers(1, 20))
        azs = np.linspace(az - arc / 2, az + arc / 2, 1 if arc == 0 else steps)
        dist = geo.sx_distance(radius, dx, dy)
        assert np.array_equal(dist, O.sx_distance(radius, dx, dy))
        centre = np.floor(np.array(dist.shape) / 2)
        src = (centre + geo.sx_source_idx_delta(azs, radius, dx, dy)).astype(int)
        assert np.array_equal(geo.sx_bresenhamlines(src, centre), O.sx_bresenhamlines(src, centre))


def test_sx_samples_dedupe_and_mask():
    dist = geo.sx_distance(150.0, 30.0, -30.0)
    dist[dist < 60.0] = np.nan
    centre = np.floor(np.array(dist.shape) / 2)
    src = (centre + geo.sx_source_idx_delta(np.linspace(40, 50, 15), 150.0, 30.0, -30.0)).astype(int)
    lines = geo.sx_bresenhamlines(src, centre)
    off, inv, window = geo.sx_samples(dist, lines)
    assert window == 5 and off.dtype == np.int32 and inv.dtype == np.float32
    assert len(np.unique(off, axis=0)) == len(off) <= len(lines)
    d = np.hypot(off[:, 0] * 30.0, off[:, 1] * 30.0)
    assert np.all(d >= 60.0) and np.allclose(inv, 1.0 / d, rtol=1e-6)


def test_valley_bank(golden):
    k = geo.valley_kernels(7, [0, 0.15, 0.3])
    assert k.dtype == np.float32 and np.allclose(k, golden["valley_kernels__7"], atol=1e-6)
    assert np.allclose(topo._ridge_kernels(7, [0, 0.15, 0.3]), -golden["valley_kernels__7"], atol=1e-6)
    for ang in (0, 30, 45, 90, 137):
        got = topo._rotate_kernels(k, np.float32(ang))
        assert got.dtype == np.float32 and np.allclose(got, golden[f"rotate_kernels__7_{ang}"], atol=2e-6)
    bank = geo.build_valley_bank(7, "valley", [0, 0.15, 0.3])
    assert bank["n_angles"] == 180 and bank["n_ch"] == 3
    for a in (0, 33, 90, 179):
        h, w, hp, c0 = bank["hw"][a]
        assert hp % 4 == 0 and hp >= h + 3 and bank["off"][a] % 4 == 0
        blk = bank["data"][bank["off"][a] : bank["off"][a] + w * hp * 4].reshape(w, hp, 4)
        # the walked row range of every kernel column holds all its non-zero weights + 3 trailing zero rows
        for j in range(w):
            lo, n = bank["cols"][c0 + j]
            assert lo % 4 == 0 and n % 4 == 0 and lo + n <= hp
            assert not blk[j, :lo].any() and not blk[j, max(lo + n - 3, 0):].any()
        want = O.valley_ridge_channel_kernels(O.rotate_kernels(O.valley_kernels(7, [0, 0.15, 0.3]), np.float32(a)))
        got = np.transpose(blk[:, :h, :3], (2, 1, 0))[:, ::-1, ::-1]  # undo flip + layout
        assert np.allclose(got, want, atol=1e-5)
        assert not blk[:, h:, :].any() and not blk[:, :, 3].any()
    with pytest.raises(ValueError):
        geo.build_valley_bank(7, "gully", [0])


def test_output_names(golden):
    names = [
        topo._dem_name(200), topo._tpi_name(200, None), topo._tpi_name(2000, 0.5), topo._std_name(200, 1),
        *topo._valley_ridge_names(1000, "valley", 0.5), *topo._gradient_names(200, 1), *topo._gradient_names(2000, 1.5),
        topo._sx_name(500.0, 45.0),
    ]
    assert names == list(golden["names__all"])


# ---- data model ------------------------------------------------------------------------------------
def test_check_dem_errors():
    z = np.zeros((4, 5), np.float32)
    good = dem_dataset(z)
    hlp.check_dem(good)
    with pytest.raises(ValueError):
        hlp.check_dem(z)
    with pytest.raises(ValueError):
        hlp.check_dem(_xr.Dataset({"a": (("x", "y"), z)}, attrs={"crs": "epsg:2056"}))
    with pytest.raises(KeyError):
        hlp.check_dem(_xr.Dataset({"a": (("y", "x"), z)}, attrs={}))
    with pytest.raises(ValueError):
        hlp.check_dem(_xr.Dataset({"a": (("y", "x"), z)}, attrs={"crs": "wgs84"}))
    with pytest.raises(TypeError):
        topo.sx(z, 0, 100)
    with pytest.raises(ValueError):
        topo.valley_ridge(z, 5, "gully")


def test_cfg_defaults():
    assert CFG.scale_std == 4 and CFG.min_elevation == -100


def _krueger_utm(lat, lon, zone):
    """Karney / Krueger series to n^4 (sub-millimetre): the cross-check for helpers._wgs84_to_utm."""
    a, f, k0 = 6378137.0, 1.0 / 298.257223563, 0.9996
    n = f / (2.0 - f)
    A = a / (1.0 + n) * (1.0 + n**2 / 4.0 + n**4 / 64.0)
    alpha = (n / 2.0 - 2.0 * n**2 / 3.0 + 5.0 * n**3 / 16.0 + 41.0 * n**4 / 180.0,
             13.0 * n**2 / 48.0 - 3.0 * n**3 / 5.0 + 557.0 * n**4 / 1440.0,
             61.0 * n**3 / 240.0 - 103.0 * n**4 / 140.0, 49561.0 * n**4 / 161280.0)
    lam = np.deg2rad(lon) - np.deg2rad((zone - 1) * 6.0 - 180.0 + 3.0)
    phi = np.deg2rad(lat)
    e = np.sqrt(f * (2.0 - f))
    t = np.sinh(np.arctanh(np.sin(phi)) - e * np.arctanh(e * np.sin(phi)))
    xi, eta = np.arctan2(t, np.cos(lam)), np.arctanh(np.sin(lam) / np.sqrt(1.0 + t * t))
    E, N = eta.copy(), xi.copy()
    for j, aj in enumerate(alpha, start=1):
        E = E + aj * np.cos(2 * j * xi) * np.sinh(2 * j * eta)
        N = N + aj * np.sin(2 * j * xi) * np.cosh(2 * j * eta)
    return 500000.0 + k0 * A * E, k0 * A * N


def test_wgs84_resolution():
    # utm.from_latlon(51.2, 7.5) == (395201.3103811303, 5673135.241182375, 32, 'U')  (utm package docs)
    e, n = hlp._wgs84_to_utm(51.2, 7.5)
    assert abs(e - 395201.3103811303) < 1e-8 and abs(n - 5673135.241182375) < 1e-8
    # independent check of the series: Krueger's transverse Mercator (n^4) agrees to a millimetre inside the zone
    ke, kn = _krueger_utm(np.array([51.2, 46.5, 10.0]), np.array([7.5, 8.9, 11.9]), 32)
    se, sn = hlp._wgs84_to_utm(np.array([51.2, 46.5, 10.0]), np.array([7.5, 8.9, 11.9]))
    assert np.abs(ke - se).max() < 2e-3 and np.abs(kn - sn).max() < 2e-3
    # zone rule of utm.latlon_to_zone_number: first element of the grid, Norway / Svalbard exceptions
    assert hlp._utm_zone_number(np.array([[46.0, 46.0]]), np.array([[5.9, 6.1]])) == 31  # Swiss DEM starting below 6 E
    assert hlp._utm_zone_number(46.0, 6.1) == 32 and hlp._utm_zone_number(60.0, 4.0) == 32  # 32V
    assert hlp._utm_zone_number(75.0, 10.0) == 33 and hlp._utm_zone_number(75.0, 8.0) == 31
    assert hlp._utm_zone_number(0.0, 180.0) == 1 and hlp._utm_zone_number(-30.0, -70.5) == 19
    # a grid that straddles 6 E keeps the zone (central meridian 3 E) of its first column
    e2, _ = hlp._wgs84_to_utm(np.array([[46.0, 46.0]]), np.array([[5.9, 6.1]]))
    assert e2[0, 1] - e2[0, 0] > 15000 and e2[0, 0] > 700000
    with pytest.raises(ValueError):
        hlp._wgs84_to_utm(np.array([-1.0, 1.0]), np.array([7.0, 7.0]))
    lon = 8.0 + np.arange(40) / 3600.0
    lat = 46.5 - np.arange(30) / 3600.0
    ds = _xr.Dataset({"a": (("y", "x"), np.zeros((30, 40), np.float32))}, coords={"x": lon, "y": lat},
                     attrs={"crs": "EPSG:4326"})
    px, res = hlp.scale_to_pixel([500], ds)
    assert res["x"].shape == (30, 40) and res["y"].shape == (30, 40)
    assert 20 < res["x"].mean() < 23 and -32 < res["y"].mean() < -30  # 1 arc-second at 46.5 N
    assert px[0] % 2 == 1


def test_fill_na_nearest():
    z = np.arange(24, dtype=np.float32).reshape(3, 8)
    z[0, 0:2] = np.nan
    z[1, 3] = np.nan
    z[2, 6:] = np.nan
    ind, filled = hlp.fill_na(dem_dataset(z))
    assert len(ind[0]) == 5
    f = hlp.get_da(filled).values
    assert not np.isnan(f).any()
    assert f[0, 0] == 2 and f[0, 1] == 2 and f[1, 3] == 10 and f[2, 6] == 21 and f[2, 7] == 21


def test_fill_na_matches_scipy_nearest():
    """helpers.fill_na against the oracle's scipy interp1d('nearest', extrapolate) restatement of xarray's
    interpolate_na: long gaps, all-NaN rows, a single valid cell, both coordinate orders."""
    rng = np.random.default_rng(0)
    z = rng.uniform(0, 1000, (40, 97)).astype(np.float32)
    z[rng.uniform(size=z.shape) < 0.3] = np.nan
    z[5, :] = np.nan
    z[6, :] = np.nan
    z[6, 50] = 3.0
    z[7, :40] = np.nan
    z[8, 60:] = np.nan
    base = dem_dataset(z, res=30.0)
    for x in (base["x"].values, base["x"].values[::-1].copy()):
        ds = _xr.Dataset({"alti": (("y", "x"), z)}, coords={"x": x, "y": base["y"].values}, attrs=base.attrs)
        ind, filled = hlp.fill_na(ds)
        want_ind, want = O.fill_na_exact(z, x)
        assert np.array_equal(ind[0], want_ind[0]) and np.array_equal(ind[1], want_ind[1])
        assert np.array_equal(hlp.get_da(filled).values, want, equal_nan=True)
    # non-uniform and unsorted x coordinates (the row-by-row path sorts per row like interp1d)
    xs = base["x"].values
    warped = xs + 7.0 * np.sin(np.arange(xs.size))
    perm = rng.permutation(xs.size)
    for x, zz in ((warped, z), (xs[perm], z[:, perm])):
        ds = _xr.Dataset({"alti": (("y", "x"), zz)}, coords={"x": x, "y": base["y"].values}, attrs=base.attrs)
        ind, filled = hlp.fill_na(ds)
        want_ind, want = O.fill_na_exact(zz, x)
        assert np.array_equal(ind[1], want_ind[1])
        assert np.array_equal(hlp.get_da(filled).values, want, equal_nan=True)


def test_to_netcdf_npz_sink(tmp_path):
    z = fractal_dem(12, 16, seed=1)
    ds = dem_dataset(z, res=30.0)
    x, y = ds["x"].values, ds["y"].values
    crop = {"x": slice(x[2], x[9]), "y": slice(y[3], y[8])}  # y descends: slice(hi, lo)
    p = hlp.to_netcdf(z * 2, ds, "tpi_200m", crop, tmp_path, "m")
    assert p.name == "topo_TPI_200M.npz"
    with np.load(p) as f:
        assert f["TPI_200M"].shape == (6, 8) and np.array_equal(f["TPI_200M"], (z * 2)[3:9, 2:10])
        assert str(f["units"]) == "m" and np.array_equal(f["x"], x[2:10])
    # the same crop applied before the copy back (device-side crop of the compute_* drivers)
    w = ds.sel_window(crop)
    assert w == {"x": (2, 10), "y": (3, 9)}
    p2 = hlp.to_netcdf((z * 2)[3:9, 2:10], ds, "tpi_300m", crop, tmp_path, "m", window=w)
    with np.load(p) as f, np.load(p2) as g:
        assert np.array_equal(f["TPI_200M"], g["TPI_300M"]) and np.array_equal(f["y"], g["y"]) and np.array_equal(f["x"], g["x"])
    assert ds.sel_window({"x": slice(x[-1] + 1, None)}) == {"x": (0, 0)}


# ---- the C-ABI library --------------------------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "topo_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(topo_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 15
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/topo_b200.h but not exported"
    assert declared == set(_lib.PROTOTYPES), "ctypes prototypes out of sync with the header"
    assert lib.topo_version() >= 100
    v = _lib.View(1440, 900, 0, 900, 0, 900)
    import ctypes

    assert lib.topo_disc_workspace_bytes(ctypes.byref(v), 17, 0, 1, 200.0, 3400.0, 0, 0) == 0  # fused: no workspace
    assert lib.topo_disc_workspace_bytes(ctypes.byref(v), 801, 0, 1, 200.0, 3400.0, 0, 0) > 4 * 900 * 1440
    # with a plane cache the planes live there: the workspace shrinks and mid sizes walk the cached planes
    rng = (200.0, 3400.0)
    # FFT route (default for sizes >= 128): the cache holds one plane spectrum per plane pair (one tile: 4096 columns x
    # 3072 rows -- 900 rows + 2 x 400 of halo fit the shorter transform along y) + the mask spectrum of the size in flight
    # (plane spectra: one pair for integer-valued DEMs, two for float ones)
    assert lib.topo_disc_cache_bytes(ctypes.byref(v), 801, 1, *rng) == 2 * 4096 * 3072 * 16
    assert lib.topo_disc_cache_bytes(ctypes.byref(v), 801, 0, *rng) == 3 * 4096 * 3072 * 16
    tall = _lib.View(1440, 16384, 0, 16384, 0, 16384)  # a whole 16384-row image: five 4096-row windows beat eight of 3072
    assert lib.topo_disc_cache_bytes(ctypes.byref(tall), 801, 1, *rng) == (5 + 1) * 4096 * 4096 * 16
    band = _lib.View(1440, 16384, 1648, 2848, 2048, 2048)  # a 2048-row band of it (8 GPUs): one 3072-row window
    assert lib.topo_disc_cache_bytes(ctypes.byref(band), 801, 1, *rng) == 2 * 4096 * 3072 * 16
    assert lib.topo_disc_shares_tsum(ctypes.byref(v), 401, 1, *rng, 801) == 2  # T and the square plane ride together
    _lib.set_option("disc_fft", False)  # the prefix-plane walk and its plane cache
    assert lib.topo_disc_cache_bytes(ctypes.byref(v), 801, 1, *rng) > 10 * 4 * 900 * 1440
    assert lib.topo_disc_cache_bytes(ctypes.byref(v), 801, 0, *rng) == 2 * lib.topo_disc_cache_bytes(ctypes.byref(v), 801, 1, *rng)
    assert lib.topo_disc_workspace_bytes(ctypes.byref(v), 41, 1, 1, *rng, 801, 0) >= 2 * 8 * 900 * 1440  # raw sums of two planes
    assert lib.topo_disc_shares_tsum(ctypes.byref(v), 41, 1, *rng, 0) == 0
    assert lib.topo_disc_shares_tsum(ctypes.byref(v), 41, 1, *rng, 801) == 1
    assert lib.topo_disc_shares_tsum(ctypes.byref(v), 41, 0, *rng, 801) == 2  # float DEM: T and fraction planes
    _lib.set_option("disc_fft", True)


def test_c_abi_argument_errors_are_reported_before_any_launch():
    """Every entry point validates its arguments first: bad calls return an error code and a message through
    topo_last_error() without touching the GPU (so this runs on a CPU-only box)."""
    import ctypes

    fake = ctypes.c_void_p(0x1000)  # never dereferenced: the checks fail first
    null = ctypes.c_void_p(0)
    v = _lib.View(64, 48, 0, 48, 0, 48)
    vp = ctypes.byref(v)

    def fails(name, *args, match):
        with pytest.raises(_lib.TopoError, match=match or None):
            _lib.call(name, *args)

    fails("topo_tpi_f32", null, 64, fake, 64, vp, 5, 0, 0.0, 1.0, null, 0, None, null, 0, null, match="null pointer")
    fails("topo_tpi_f32", fake, 64, fake, 64, vp, 9000, 0, 0.0, 1.0, null, 0, None, null, 0, null, match="kernel size")
    fails("topo_std_f32", fake, 64, fake, 64, vp, 5, 0, float("nan"), 1.0, null, 0, None, null, 0, null, match="not finite")
    fails("topo_std_f32", fake, 32, fake, 64, vp, 5, 0, 0.0, 1.0, null, 0, None, null, 0, null, match="pitch")
    bad = _lib.View(64, 48, 10, 20, 0, 48)  # the band does not contain the requested rows
    fails("topo_tpi_f32", fake, 64, fake, 64, ctypes.byref(bad), 5, 0, 0.0, 1.0, null, 0, None, null, 0, null, match="")
    fails("topo_tpi_f32", fake, 64, fake, 64, vp, 301, 1, 0.0, 100.0, null, 0, None, null, 0, null, match="workspace too small")
    fails("topo_sx_f32", fake, 64, fake, 64, 64 * 48, vp, fake, fake, fake, 0, 3, ctypes.c_float(10.0), -3, 3, -3, 3, null,
          match="n_az")
    fails("topo_sx_f32", fake, 64, fake, 64, 64 * 48, vp, fake, fake, fake, 1, 3, ctypes.c_float(10.0), -4, 3, -3, 3, null,
          match="exceed the window")
    fails("topo_valley_ridge_f32", fake, 64, fake, fake, 64, vp, fake, fake, fake, fake, 180, 5, 9, 9, 0, null,
          match="5 channels in one bank")
    fails("topo_fill_na_f32", fake, 64, fake, 64, 48, 200000, null, 0, ctypes.c_float(0.0), null, null, match="")
    # size 1: kernel sum - 1 == 0 in the reference; handled by the fill entry point, which needs a real buffer -> not called here
    assert _lib.load().topo_last_error()


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        topo.tpi(np.zeros((8, 8), np.float32), 3)
    with pytest.raises(RuntimeError):
        topo.gradient(np.zeros((8, 8), np.float32), 2.0, {"x": np.ones(8), "y": np.ones(8)})


def test_disc_queries_agree_with_the_plan_over_a_size_x_range_grid():
    """The workspace / cache / tsum queries make the same plan as the call itself (host-only introspection): a
    shared-plane sweep whose layout overflows for one size falls back consistently, and STD on wide ranges splits
    the square plane instead of failing (an Alpine 0..4800 m DEM from size ~800, 0..8848 m from size ~400)."""
    import ctypes

    lib = _lib.load()
    v = _lib.View(4096, 4096, 0, 4096, 0, 4096)
    vp = ctypes.byref(v)
    info = (ctypes.c_longlong * 32)()
    for fft_on in (True, False):
      _lib.set_option("disc_fft", fft_on)
      for integer in (1, 0):
          for zmin, zmax in ((200.0, 3400.0), (200.0, 4700.0), (0.0, 4800.0), (0.0, 8848.0), (-11000.0, 8848.0)):
              for hint in (0, 801, 2001):
                  for size in (5, 21, 41, 201, 401, 801, 2001):
                      if hint and size > hint:
                          continue
                      for what in (0, 1):
                          assert lib.topo_disc_plan_info(vp, size, what, integer, zmin, zmax, hint, 0, info) == 0, (
                              size, what, integer, zmin, zmax, hint, lib.topo_last_error())
                          fused, cached, ws, off_partial = info[1], info[4], info[14], info[15]
                          need = ws if info[17] else (0 if fused else (ws - off_partial if cached else ws))
                          got = lib.topo_disc_workspace_bytes(vp, size, what, integer, zmin, zmax, hint, 0)
                          assert got == need, (size, what, integer, zmin, zmax, hint, got, need)
                      if lib.topo_disc_shares_tsum(vp, size, integer, zmin, zmax, hint):
                          a = (ctypes.c_longlong * 32)()
                          b = (ctypes.c_longlong * 32)()
                          lib.topo_disc_plan_info(vp, size, 0, integer, zmin, zmax, hint, 1, a)
                          lib.topo_disc_plan_info(vp, size, 1, integer, zmin, zmax, hint, 1, b)
                          # TPI_I + STD_I (integer DEM) or TPI_X + STD_F (float DEM), both two-pass
                          assert (a[0], b[0], a[1], b[1]) == ((4, 2, 0, 0) if integer else (1, 3, 0, 0)) and a[17] == b[17]
    _lib.set_option("disc_fft", False)
    # the advisor's cases: std(801) on 200..4800 and std(2001) now plan (split squares) instead of raising
    for size in (801, 2001):
        assert lib.topo_disc_plan_info(vp, size, 1, 1, 200.0, 4800.0, 0, 0, info) == 0 and info[16] == 1
    assert lib.topo_disc_plan_info(vp, 801, 1, 1, 200.0, 3400.0, 0, 0, info) == 0 and info[16] == 0
    # a cache laid out for a split square holds one more plane region
    base = lib.topo_disc_cache_bytes(vp, 801, 1, 200.0, 3400.0)
    assert base > 0 and lib.topo_disc_cache_bytes(vp, 801, 1, 0.0, 4800.0) * 2 == base * 3
    _lib.set_option("disc_fft", True)
    # FFT route: exactness of the rounded sums decides the split (an 0..8848 m range at size 801), not the 32-bit spans
    assert lib.topo_disc_plan_info(vp, 801, 1, 1, 200.0, 3400.0, 0, 0, info) == 0 and (info[17], info[16]) == (1, 0)
    assert lib.topo_disc_plan_info(vp, 801, 1, 1, 0.0, 8848.0, 0, 0, info) == 0 and (info[17], info[16]) == (1, 1)
    # float std of sizes 5 .. 13: the packed-word register kernel when every fraction is exactly representable in the
    # fraction field (|z| >= 2^(23 - Sf)), else the fused kernel; sizes above 13 stay fused
    for (zmin, zmax, size, tiny) in ((200.0, 3400.0, 9, 1), (-3400.0, -200.0, 13, 1), (0.0, 500.0, 9, 0), (-50.0, 900.0, 5, 0),
                                     (200.0, 3400.0, 15, 0), (200.0, 3400.0, 8, 0)):
        assert lib.topo_disc_plan_info(vp, size, 1, 0, zmin, zmax, 0, 0, info) == 0 and (info[0], info[1], info[3]) == (3, 1, tiny), (
            zmin, zmax, size, list(info)[:4])


def test_tiler_band_plan_and_stats_merge():
    from topo_descriptors_b200 import tiler

    plan = tiler.band_plan(1000, 40, 300)
    assert plan == [(0, 300, 0, 340), (300, 600, 260, 640), (600, 900, 560, 940), (900, 1000, 860, 1000)]
    assert tiler.band_plan(10, 400, 4) == [(0, 4, 0, 10), (4, 8, 0, 10), (8, 10, 0, 10)]  # halo wider than the image
    with pytest.raises(ValueError):
        tiler.band_plan(10, 1, 0)
    a = {"min": 1.0, "max": 5.0, "nonfinite": 0, "nonint": 2, "sum": 10.0, "sumsq": 30.0, "n": 4}
    b = {"min": -2.0, "max": 3.0, "nonfinite": 1, "nonint": 0, "sum": 1.0, "sumsq": 5.0, "n": 3}
    m = tiler.merge_stats([a, b])
    assert m == {"min": -2.0, "max": 5.0, "nonfinite": 1, "nonint": 2, "sum": 11.0, "sumsq": 35.0, "n": 7}
    assert not topo._out_of_core(np.zeros((4, 4), np.float32))


def test_rotation_plan_equals_scipy_rotate_setup():
    """_geometry.plan_rotations restates the set-up of scipy.ndimage.rotate(reshape=True): same output boxes as scipy for
    every angle; profiles/proto/rotate_restated.py (run on CPU) shows the full restatement -- coordinates, quadratic
    B-spline weights, mirrored taps -- bit-identical to scipy, which is what csrc/rotate.cu executes."""
    from scipy import ndimage

    from topo_descriptors_b200 import _geometry as geo

    k = geo.valley_kernels(21, [0, 0.3])
    angles = np.arange(0, 180, dtype=np.float32)
    recs, total = geo.plan_rotations(k.shape[1:], k.shape[0], angles)
    assert recs.dtype.itemsize == 64
    pos = 0
    for r, angle in zip(recs, angles):
        want = ndimage.rotate(k, angle, axes=(1, 2), reshape=True, order=2, mode="constant", cval=-9999)
        assert (int(r["oh"]), int(r["ow"])) == want.shape[1:] and int(r["out_off"]) == pos
        pos += want.size
    assert pos == total
    sys.path.insert(0, os.path.join(ROOT, "profiles", "proto"))
    import rotate_restated

    for angle in (0.0, 17.0, 45.0, 90.0, 133.0):
        want = ndimage.rotate(k, angle, axes=(1, 2), reshape=True, order=2, mode="constant", cval=-9999)
        assert np.array_equal(rotate_restated.rotate_restated(k, angle), want)
