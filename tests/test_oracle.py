"""Pin the CPU oracle (oracle/oracle.py) to the reference: golden vectors produced by the
reference itself (tests/golden/ref_small.npz, oracle/make_golden.py) and the reference's four
known-answer tests (test/test_topo.py:6-67, test/test_helpers.py:6-11).  CPU only."""

import numpy as np
import pytest

from oracle import oracle as O


# ---- the reference's own known-answer tests, verbatim vectors ------------------------------
def test_sx_distance_known_answer():
    out = O.sx_distance(150.0, 50.0, 40.0)
    expected = np.array([256.1249695, 219.31712199, 188.67962264, 167.63054614, 160.0,
                         167.63054614, 188.67962264, 219.31712199, 256.1249695])
    assert np.all(np.isclose(out[0, :], expected))
    assert out.dtype == np.float64


def test_sx_bresenhamlines_known_answer():
    out = O.sx_bresenhamlines(np.array([[8, 9], [17, 22]]), np.array([15, 15]))
    expected = np.array([[9, 10], [10, 11], [11, 12], [12, 12], [13, 13], [14, 14],
                         [17, 21], [16, 20], [16, 19], [16, 18], [16, 17], [15, 16]])
    assert np.all(out == expected)
    assert out.dtype == np.int64


def test_sx_source_idx_delta_known_answer():
    out = O.sx_source_idx_delta(np.array([3.0, 4.0, 5.0, 6.0]), 500, 20, 30)
    assert np.all(out == np.array([[17, 1], [17, 2], [17, 2], [17, 3]]))
    assert out.dtype == np.int64


def test_round_up_to_odd_known_answer():
    out = O.round_up_to_odd(np.arange(0.1, 10, 0.7))
    assert out.dtype == np.int64
    assert list(out) == [1, 1, 1, 3, 3, 3, 5, 5, 5, 7, 7, 7, 9, 9, 9]


# ---- golden vectors from the reference -----------------------------------------------------
@pytest.mark.parametrize("size", [3, 4, 5, 6, 7, 17, 33])
def test_circular_kernel(golden, size):
    assert np.array_equal(O.circular_kernel(size), golden[f"circular_kernel__{size}"])


@pytest.mark.parametrize("size", [3, 5, 6, 7, 17, 33])
def test_tpi_tight_pin(golden, size):
    """Exact pin (reference forced onto scipy's exact direct path), odd AND even sizes."""
    got = O.tpi_exact(golden["in__zi"], size)
    assert np.abs(got - golden[f"tpi_tight__{size}"]).max() < 1e-9


@pytest.mark.parametrize("size", [3, 5, 6, 7, 17, 33])
def test_tpi_literal_within_reference_noise(golden, size):
    """The literal reference (float32 FFT) carries ~1e-3 m of its own noise."""
    got = O.tpi_exact(golden["in__z"], size)
    assert np.abs(got - golden[f"tpi_lit__{size}"]).max() < 3e-3
    lit = O.tpi_literal(golden["in__z"], size)
    assert np.abs(lit - golden[f"tpi_lit__{size}"]).max() < 3e-3


def test_tpi_with_sigma(golden):
    got = O.tpi_exact(golden["in__z"], 7, sigma=1.75)
    assert np.abs(got - golden["tpi_lit_sigma__7_1.75"]).max() < 3e-3


def test_conv_direct_vs_fft64(golden):
    z = golden["in__z"]
    for size in (5, 6, 17):
        k = O.circular_kernel(size)
        a, b = O.conv2_same_direct(z, k), O.conv2_same_fft64(z, k)
        assert np.abs(a - b).max() < 1e-7


@pytest.mark.parametrize("size", [3, 5, 7, 17])
def test_std_pins(golden, size):
    # small-range integer DEM: the reference's float64 run is accurate to ~3e-5 m
    got = O.std_exact(golden["in__zc"], size)
    assert got.dtype == np.float64
    assert np.abs(got - golden[f"std_f64c__{size}"]).max() < 1e-4
    # SRTM-like integer DEM, reference fed float64: ~4e-3 m of kernel-FFT noise
    got = O.std_exact(golden["in__zi"], size)
    assert np.abs(got - golden[f"std_f64__{size}"]).max() < 1e-2
    # float DEM (int32-truncation quirk active), reference fed float64
    got = O.std_exact(golden["in__z"], size)
    assert np.abs(got - golden[f"std_f64flt__{size}"]).max() < 0.1
    # literal float32 reference: cancellation noise up to ~0.1 m at these sizes
    assert np.abs(O.std_exact(golden["in__zi"], size) - golden[f"std_lit__{size}"]).max() < 0.5


def test_std_with_sigma(golden):
    got = O.std_exact(golden["in__zc"], 7, sigma=1.75)
    assert np.abs(got - golden["std_f64c_sigma__7_1.75"]).max() < 5e-3


def test_gaussian_restated_bit_exact(golden):
    z = golden["in__z"]
    assert np.array_equal(O.dem_smooth(z, 3.3), golden["dem__3.3"])
    assert np.array_equal(O.dem_smooth(z, 20.0), golden["dem__20"])


def test_gaussian_restated_vs_installed_scipy():
    from scipy import ndimage

    rng = np.random.default_rng(0)
    a = (rng.random((37, 53)) * 3000).astype(np.float32)
    for sigma in (0.6, 1.75, 5.0, (2.0, 7.5), 30.0):
        assert np.array_equal(O.gaussian_filter_restated(a, sigma), ndimage.gaussian_filter(a, sigma))


def test_np_gradient_restated():
    rng = np.random.default_rng(1)
    a = (rng.random((9, 11)) * 100).astype(np.float32)
    gy, gx = np.gradient(a)
    assert np.array_equal(O.np_gradient_restated(a, 0), gy)
    assert np.array_equal(O.np_gradient_restated(a, 1), gx)


def test_sobel(golden):
    dx, dy = O.sobel_exact(golden["in__z"])
    assert np.array_equal(dx, golden["sobel__dx"])
    assert np.array_equal(dy, golden["sobel__dy"])
    yy, xx = np.mgrid[:20, :30]
    px, py = O.sobel_exact((3 * xx + 2 * yy).astype(np.float32))
    assert np.allclose(px[1:-1, 1:-1], 3) and np.allclose(py[1:-1, 1:-1], 2)


@pytest.mark.parametrize("sigma,ratio", [(0.75, 1), (1.75, 1), (4.25, 1), (4.25, 1.5), (16.75, 1)])
def test_gradient_bit_exact(golden, sigma, ratio):
    res = {"x": golden["scale_to_pixel__res_x"], "y": golden["scale_to_pixel__res_y"]}
    out = O.gradient_exact(golden["in__z"], sigma, res, ratio)
    for nm, arr in zip(("dx", "dy", "slope", "aspect"), out):
        assert arr.dtype == np.float32
        assert np.array_equal(arr, golden[f"gradient__{sigma}_{ratio}_{nm}"]), nm


def test_gradient_res2d_and_flat(golden):
    res = {"x": golden["gradient_res2d__x"], "y": golden["gradient_res2d__y"]}
    out = O.gradient_exact(golden["in__z"], 1.75, res)
    for nm, arr in zip(("dx", "dy", "slope", "aspect"), out):
        assert np.array_equal(arr, golden[f"gradient_res2d__{nm}"]), nm
    flat = np.full((16, 24), 512.25, dtype=np.float32)
    out = O.gradient_exact(flat, 1.75, {"x": np.full(24, 30.0), "y": np.full(16, -30.0)})
    assert np.array_equal(out[3], golden["gradient_flat__aspect"])
    assert np.array_equal(out[2], golden["gradient_flat__slope"])


def test_valley_kernels_and_rotation(golden):
    assert np.allclose(O.valley_kernels(7, [0, 0.15, 0.3]), golden["valley_kernels__7"], atol=1e-6)
    assert np.allclose(O.valley_kernels(11, [0, 0.2, 0.4]), golden["valley_kernels__11"], atol=1e-6)
    k = O.valley_kernels(7, [0, 0.15, 0.3])
    for ang in (0, 30, 45, 90, 137):
        ref = golden[f"rotate_kernels__7_{ang}"]
        got = O.rotate_kernels(k, np.float32(ang))
        assert got.shape == ref.shape and got.dtype == np.float32
        assert np.allclose(got, ref, atol=2e-6)


@pytest.mark.parametrize("case,size,mode,flats,sigma", [
    ("valley7", 7, "valley", (0, 0.15, 0.3), None),
    ("ridge9", 9, "ridge", (0, 0.2, 0.4), 1.125),
    ("valley5f2", 5, "valley", (0, 0.3), None),
])
def test_valley_ridge(golden, case, size, mode, flats, sigma):
    """Pins the 3-D-convolution channel mixing (incl. F=2) and the strict '>' argmax."""
    norm, direction, gap = O.valley_ridge_exact(golden["in__z"], size, mode, flats, sigma, return_gap=True)
    rn, rd = golden[f"valley_ridge__{case}_norm"], golden[f"valley_ridge__{case}_dir"]
    assert np.abs(norm - rn).max() < 2e-4
    decidable = gap > 1e-3
    assert np.array_equal(direction[decidable], rd[decidable])
    assert (direction != rd).mean() < 0.01


def test_valley_ridge_unknown_mode():
    with pytest.raises(ValueError):
        O.valley_ridge_exact(np.zeros((8, 8), np.float32), 5, "gully")


def test_sx_helpers(golden):
    assert np.array_equal(O.sx_distance(150.0, 50.0, 40.0), golden["sx_distance__150_50_40"])
    assert np.array_equal(O.sx_distance(150.0, 30.0, -30.0), golden["sx_distance__150_30_-30"])
    assert np.array_equal(O.sx_source_idx_delta(np.linspace(-5, 5, 15), 150.0, 30.0, -30.0),
                          golden["sx_source_idx_delta__b"])
    assert np.array_equal(O.sx_bresenhamlines(np.array([[8, 9], [17, 22]]), np.array([15, 15])),
                          golden["sx_bresenhamlines__a"])


def test_sx(golden):
    z, x, y = golden["in__z"], golden["in__x"], golden["in__y"]
    for i, (az, rad, rmin, h, arc, steps) in enumerate(golden["sx__cases"]):
        got = O.sx_exact(z, x, y, az, rad, height=h, azimuth_arc=arc, azimuth_steps=int(steps), radius_min=rmin)
        assert got.dtype == np.float32
        assert np.array_equal(got, golden[f"sx__case{i}"]), i
    got = O.sx_exact(golden["sx_nan__in"], x, y, 45, 150)
    assert np.array_equal(got, golden["sx_nan__out"], equal_nan=True)
    got = O.sx_exact(z, x, y, 0, 150, radius_min=1000.0)
    assert np.array_equal(got, golden["sx_allmasked__out"], equal_nan=True)


def test_reference_still_agrees_when_present(golden):
    """In the authoring container, re-run the reference and check the fixture is current."""
    from oracle import ref_runner

    if not ref_runner.available():
        pytest.skip("reference tree not present (GPU box)")
    import warnings

    warnings.simplefilter("ignore")
    topo, _ = ref_runner.load()
    assert np.array_equal(topo.tpi(golden["in__z"], 7), golden["tpi_lit__7"])
    assert np.array_equal(topo.dem(golden["in__z"], 3.3), golden["dem__3.3"])
