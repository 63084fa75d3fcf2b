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
