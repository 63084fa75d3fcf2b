"""Seeded random sweeps over small, awkward rasters (1-row / 1-column images, images smaller than the
kernel, widths that are not multiples of 4 / 32 / 128, even and odd sizes, integer and float DEMs): the CUDA
path against the CPU oracle, same tolerances as test_gpu_parity.py.  These shapes exercise every clamp, halo
and tail path of the kernels; the reference's own tests only use 3 x 3 ... 9 x 9 arrays.
"""

import numpy as np
import pytest

from oracle import oracle as O
from topo_descriptors_b200 import _xr, helpers as hlp, topo
from topo_descriptors_b200.synth import dem_dataset

pytestmark = pytest.mark.gpu

TOL_M, TOL_D, TOL_DEG = 1e-3, 1e-4, 1e-3


def maxdiff(a, b):
    d = np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64))
    return float(np.nanmax(d)) if d.size else 0.0


def _dem(rng, ny, nx, integer):
    z = rng.uniform(-200.0, 3000.0, (ny, nx))
    # some spatial structure so that sums do not just average out
    z += 500.0 * np.sin(np.arange(nx) / 7.0)[None, :] + 300.0 * np.cos(np.arange(ny) / 5.0)[:, None]
    z = np.rint(z) if integer else z
    return z.astype(np.float32)


SHAPES = [(1, 1), (1, 37), (41, 1), (2, 3), (3, 130), (7, 129), (17, 257), (33, 31), (64, 64), (70, 301), (9, 1025)]


@pytest.mark.parametrize("seed", range(4))
def test_tpi_std_random_small(seed):
    rng = np.random.default_rng(100 + seed)
    for ny, nx in SHAPES:
        for size in sorted({2, 3, 4, 5, 6, 9, 13, int(rng.integers(14, 60)), int(rng.integers(60, 140))}):
            integer = bool(rng.integers(0, 2))
            z = _dem(rng, ny, nx, integer)
            got_t, got_s = topo.tpi(z, size), topo.std(z, size)
            assert got_t.shape == z.shape and got_s.shape == z.shape
            assert maxdiff(got_t, O.tpi_exact(z, size)) <= TOL_M, (ny, nx, size, integer)
            assert maxdiff(got_s, O.std_exact(z, size)) <= TOL_M, (ny, nx, size, integer)


@pytest.mark.parametrize("seed", range(3))
def test_gaussian_gradient_random_small(seed):
    rng = np.random.default_rng(200 + seed)
    for ny, nx in SHAPES:
        if ny < 2 or nx < 2:
            continue  # np.gradient needs two samples per axis (the reference raises as well)
        z = _dem(rng, ny, nx, bool(rng.integers(0, 2)))
        res = {"x": np.full(nx, 30.0), "y": np.full(ny, -30.0)}
        for sigma in (0.75, float(rng.uniform(1.1, 3.0)), float(rng.uniform(3.0, 9.0)), 20.5):
            want = O.gradient_exact(z, sigma, res)
            got = topo.gradient(z, sigma, res)
            assert maxdiff(got[0], want[0]) <= TOL_D and maxdiff(got[1], want[1]) <= TOL_D, (ny, nx, sigma)
            assert maxdiff(got[2], want[2]) <= TOL_D, (ny, nx, sigma)
            steep = np.hypot(want[0], want[1]) > 0.2
            d = np.abs(got[3].astype(np.float64) - want[3].astype(np.float64))
            d = np.minimum(d, 360.0 - d)
            assert (d[steep] <= TOL_DEG).all(), (ny, nx, sigma)
            sm = topo.dem(z, sigma)
            assert maxdiff(sm, O.gaussian_filter_restated(z, sigma)) <= 1e-3, (ny, nx, sigma)


def test_gaussian_one_row_and_one_column():
    rng = np.random.default_rng(7)
    for shape in ((1, 50), (50, 1), (1, 1), (2, 2)):
        z = _dem(rng, *shape, False)
        for sigma in (1.0, 4.0, 30.0):
            assert maxdiff(topo.dem(z, sigma), O.gaussian_filter_restated(z, sigma)) <= 1e-3, (shape, sigma)


@pytest.mark.parametrize("seed", range(2))
def test_sx_random_small(seed):
    rng = np.random.default_rng(300 + seed)
    for ny, nx in ((40, 50), (33, 129), (64, 257), (21, 21)):
        z = _dem(rng, ny, nx, False)
        z[rng.uniform(size=z.shape) < 0.02] = np.nan
        ds = dem_dataset(z, res=30.0)
        x, y = ds["x"].values, ds["y"].values
        for az, radius in ((float(rng.uniform(0, 360)), 150.0), (float(rng.uniform(0, 360)), 330.0), (90.0, 90.0)):
            got = topo.sx(ds, az, radius)
            want = O.sx_exact(z, x, y, az, radius)
            assert np.array_equal(np.isnan(got), np.isnan(want)), (ny, nx, az, radius)
            assert maxdiff(got, want) <= TOL_DEG, (ny, nx, az, radius)
            assert np.array_equal(got == 0, want == 0)
    # a window larger than the image: everything is frame
    small = dem_dataset(_dem(rng, 8, 9, False), res=30.0)
    assert not topo.sx(small, 10.0, 500.0).any()


@pytest.mark.parametrize("seed", range(2))
def test_valley_ridge_random_small(seed):
    rng = np.random.default_rng(400 + seed)
    for (ny, nx), size in (((20, 30), 5), ((33, 65), 7), ((9, 140), 9), ((64, 64), 11), ((5, 5), 7)):
        z = _dem(rng, ny, nx, False)
        mode = "valley" if rng.integers(0, 2) else "ridge"
        norm, direction = topo.valley_ridge(z, size, mode)
        o_norm, o_dir, gap = O.valley_ridge_exact(z, size, mode, (0, 0.15, 0.3), None, return_gap=True)
        assert maxdiff(norm, o_norm) <= TOL_M, (ny, nx, size, mode)
        decidable = gap > 1e-2
        assert np.array_equal(direction[decidable], o_dir[decidable]), (ny, nx, size, mode)


def test_fill_na_then_descriptors_on_ragged_nan_pattern():
    """The script's pipeline (fill_na -> compute) on a DEM with NaN stripes, blocks and an empty row."""
    from topo_descriptors_b200 import prestage

    rng = np.random.default_rng(5)
    z = _dem(rng, 90, 131, False)
    z[:, 17] = np.nan
    z[30:40, 50:90] = np.nan
    z[77, :] = np.nan
    ds = dem_dataset(z, res=30.0)
    ind_h, ds_h = hlp.fill_na(ds)
    ind_d, ds_d = prestage.fill_na_resident(ds)
    assert np.array_equal(ind_d[0].cpu().numpy(), ind_h[0]) and np.array_equal(ind_d[1].cpu().numpy(), ind_h[1])
    filled = hlp.get_da(ds_d).values.numpy()
    assert np.array_equal(filled, hlp.get_da(ds_h).values, equal_nan=True)
    assert np.isnan(filled[77]).all() and not np.isnan(np.delete(filled, 77, axis=0)).any()
    # the all-NaN row makes the FFT-based descriptors NaN everywhere, like the reference
    assert np.isnan(topo.tpi(filled, 5)).all()
    assert isinstance(ds_d, _xr.Dataset)
