"""The octagon decomposition of a disc sum (csrc/disc.cu: hybrid_walk with p.oct) restated on the CPU with numpy
tables (profiles/proto/octagon.py): rectangle + two 45-degree trapezoids from sheared summed-area tables + row /
column caps + corner diagonals must tile the reference's circular_kernel exactly, for every radius."""

import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("octagon_proto", os.path.join(ROOT, "profiles", "proto", "octagon.py"))
octagon = importlib.util.module_from_spec(spec)
spec.loader.exec_module(octagon)


@pytest.mark.parametrize("m", [2, 3, 7, 20, 33, 40, 61])
def test_octagon_tiles_the_disc_exactly(m):
    rng = np.random.default_rng(m)
    pl = octagon.plan(m)
    u, v, s, diag, lines = pl
    assert v <= u <= m and u * u + v * v <= m * m
    n = 2 * m + 9
    z = rng.integers(0, 1000, (n + 2, n + 2))
    tables = octagon.tables(z)
    for y, x in ((m + 2, m + 3), (m + 4, m + 1), (n - m - 1, n - m - 2)):
        assert octagon.disc_sum_octagon(z, m, y, x, tables, pl) == octagon.disc_sum_direct(z, m, y, x)


def test_octagon_walks_fewer_lines_than_the_square():
    for m in (80, 200, 400):
        _, _, _, _, lines = octagon.plan(m)
        square_lines = 4 * (m - int(m / np.sqrt(2)))
        assert lines < 0.65 * square_lines


def _cover_count(m):
    """How many pieces of the octagon decomposition cover each lattice point of [-m, m]^2 (mirrors the device
    logic: plan_geometry's choice of u, v; build_diag_table's trimmed runs; the cap ranges of hybrid_walk)."""
    u, v, s, diag, _ = octagon.plan(m)
    n = 2 * m + 1
    ii, jj = np.mgrid[-m : m + 1, -m : m + 1]
    ai, aj = np.abs(ii), np.abs(jj)
    cover = np.zeros((n, n), dtype=np.int32)
    cover += (ai <= v) & (aj <= u)                                   # summed-area rectangle
    cover += (ai > v) & (ai <= u) & (aj <= s - ai)                   # the two trapezoids
    h = np.floor(np.sqrt(np.maximum(m * m - np.arange(-m, m + 1) ** 2, 0).astype(np.float64))).astype(np.int64)
    h = np.where((h + 1) ** 2 <= m * m - np.arange(-m, m + 1) ** 2, h + 1, h)
    h = np.where(h ** 2 > m * m - np.arange(-m, m + 1) ** 2, h - 1, h)
    cover += (ai > u) & (aj <= h[:, None])                           # row caps: full disc rows
    cover += (aj > u) & (ai <= h[None, :])                           # column caps: full disc columns
    dmax = s + len(diag)
    for d in range(s + 1, dmax + 1):                                 # corner diagonals, build_diag_table's trimming
        lo, hi = max(0, d - u), min(u, d)
        while lo <= hi and lo * lo + (d - lo) * (d - lo) > m * m:
            lo += 1
        while hi >= lo and hi * hi + (d - hi) * (d - hi) > m * m:
            hi -= 1
        if hi >= lo:
            cover += (ai + aj == d) & (ai >= lo) & (ai <= hi)
    return cover, (ii * ii + jj * jj <= m * m)


@pytest.mark.parametrize("m", list(range(2, 80)) + [100, 120, 150, 200, 241, 400, 777, 1000])
def test_octagon_pieces_cover_every_disc_pixel_exactly_once(m):
    cover, disc = _cover_count(m)
    assert np.array_equal(cover, disc.astype(np.int32)), m


def test_cxx_planner_matches_the_prototype_for_every_odd_size():
    """The host planner of csrc/disc.cu (through the host-only topo_disc_plan_info) picks the same octagon as the
    CPU prototype, for every odd size the cached two-pass walk accepts, and stays inside the hardware limits."""
    import ctypes
    import math

    from topo_descriptors_b200 import _lib

    lib = _lib.load()
    info = (ctypes.c_longlong * 32)()
    _lib.set_option("disc_fft", False)  # this test is about the prefix-plane walk (sizes >= 128 take the FFT route by default)
    try:
        _planner_checks(lib, info, octagon, math, ctypes, _lib)
    finally:
        _lib.set_option("disc_fft", True)
    # default route for large sizes: FFT convolution of the integer planes, T = 4096 windows for a 20 km sweep
    v = _lib.View(2048, 4096, 0, 4096, 0, 4096)
    assert lib.topo_disc_plan_info(ctypes.byref(v), 801, 1, 1, 0.0, 3400.0, 801, 0, info) == 0
    assert (info[17], info[18], info[4], info[16]) == (1, 4096, 1, 0) and info[19] == 2 * 1
    assert lib.topo_disc_plan_info(ctypes.byref(v), 41, 1, 1, 0.0, 3400.0, 801, 0, info) == 0 and (info[17], info[4]) == (1, 1)
    assert lib.topo_disc_plan_info(ctypes.byref(v), 21, 1, 1, 0.0, 3400.0, 801, 0, info) == 0 and (info[17], info[4]) == (0, 0)
    assert lib.topo_disc_plan_info(ctypes.byref(v), 41, 1, 1, 0.0, 3400.0, 0, 0, info) == 0 and info[17] == 0
    assert lib.topo_disc_plan_info(ctypes.byref(v), 161, 0, 0, 0.0, 3400.5, 0, 0, info) == 0 and (info[17], info[18], info[0]) == (1, 2048, 1)


def _planner_checks(lib, info, octagon, math, ctypes, _lib):
    sizes = list(range(41, 1203, 2)) + [2001, 3001, 4095, 8191]
    for size in sizes:
        v = _lib.View(2048, 4096, 0, 4096, 0, 4096)
        assert lib.topo_disc_plan_info(ctypes.byref(v), size, 0, 1, 0.0, 500.0, size, 0, info) == 0, size
        mode, fused, hybrid, tiny, cached, oct_, u, vv, ndiag, acc, smem = list(info)[:11]
        assert (mode, fused, hybrid, tiny, cached, oct_) == (4, 0, 1, 0, 1, 1), size
        m = size // 2
        pu, pv, ps, _, _ = octagon.plan(m)
        dmax = max(q + math.isqrt(m * m - q * q) for q in range(m + 1))
        assert (u, vv, ndiag) == (pu, pv, dmax - ps), size
        assert vv <= u <= m and u * u + vv * vv <= m * m and (u + 1) ** 2 * 2 > m * m
        assert smem <= 200 * 1024 and info[11] == m
    # without a cache: the inscribed square; small sizes: fused / tiny; even sizes: plain row spans
    v = _lib.View(2048, 4096, 0, 4096, 0, 4096)
    assert lib.topo_disc_plan_info(ctypes.byref(v), 801, 0, 1, 0.0, 500.0, 0, 0, info) == 0
    assert (info[2], info[4], info[5], info[6]) == (1, 0, 0, int(400 / math.sqrt(2)))
    assert lib.topo_disc_plan_info(ctypes.byref(v), 9, 0, 1, 0.0, 500.0, 801, 0, info) == 0 and (info[1], info[3]) == (1, 1)
    assert lib.topo_disc_plan_info(ctypes.byref(v), 21, 1, 1, 0.0, 500.0, 801, 0, info) == 0 and (info[1], info[3]) == (1, 0)
    assert lib.topo_disc_plan_info(ctypes.byref(v), 400, 0, 1, 0.0, 500.0, 801, 0, info) == 0 and (info[1], info[2], info[4]) == (0, 0, 1)
    # float DEM: quantised plane with the scale of the sweep's largest size, or the exact two-plane mode for wide ranges
    assert lib.topo_disc_plan_info(ctypes.byref(v), 201, 0, 0, 0.0, 3400.5, 801, 0, info) == 0 and info[0] == 0 and info[4] == 1
    assert lib.topo_disc_plan_info(ctypes.byref(v), 201, 0, 0, -2e5, 3e6, 801, 0, info) == 0 and info[0] == 1
    assert lib.topo_disc_plan_info(ctypes.byref(v), 201, 1, 0, 0.0, 3400.5, 801, 0, info) == 0 and info[0] == 3
