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
