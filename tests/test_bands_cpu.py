"""Row-band host logic on CPU: partitioning, neighbour halo exchange and the statistics all-reduce,
world_size 2 and 3 over gloo (the N > 1 path of bench.py / bands.sweep without the kernels)."""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import numpy as np
import torch.multiprocessing as mp

from topo_descriptors_b200 import bands


def test_partition_rows():
    assert bands.partition_rows(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert bands.partition_rows(16384, 8)[3] == (6144, 8192)
    for gny, w in [(1, 1), (7, 2), (900, 8), (16384, 8), (5, 5)]:
        parts = bands.partition_rows(gny, w)
        assert parts[0][0] == 0 and parts[-1][1] == gny
        assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
        p = np.array(parts)
        sizes = p[:, 1] - p[:, 0]
        assert sizes.max() - sizes.min() <= 1
    # weighted bands (e.g. by measured D2H rate): proportional, contiguous, at least one row each
    assert bands.partition_rows(100, 4, [2, 2, 3, 3]) == [(0, 20), (20, 40), (40, 70), (70, 100)]
    assert bands.partition_rows(5, 4, [0, 0, 1, 0]) == [(0, 1), (1, 2), (2, 4), (4, 5)]
    w = [12.1] * 4 + [18.1] * 4
    parts = bands.partition_rows(16384, 8, w)
    assert parts[0][0] == 0 and parts[-1][1] == 16384 and all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
    assert abs((parts[7][1] - parts[7][0]) / (parts[0][1] - parts[0][0]) - 18.1 / 12.1) < 0.01
    ctx = bands.BandContext(100, 8, rank=1, world=4)
    assert (ctx.r0, ctx.r1, ctx.rows) == (25, 50, 25)
    assert ctx.halo_extent(10) == (15, 60) and bands.BandContext(100, 8, 0, 4).halo_extent(10) == (0, 35)


def test_sweep_halo():
    assert bands.sweep_halo([5, 801], []) == 400
    assert bands.sweep_halo([], [200.25]) == 802  # Gaussian radius 801 + 1 row for the central difference
    assert bands.sweep_halo([17], [0.75]) == 8


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, gny, nx, halo, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(gny * nx, dtype=torch.float32).reshape(gny, nx)
        ctx = bands.BandContext(gny, nx, rank, world)
        core = full[ctx.r0 : ctx.r1].clone()
        band, gy0 = bands.exchange_halo(core, ctx, halo)
        a, b = ctx.halo_extent(halo)
        ok = gy0 == a and band.shape == (b - a, nx) and torch.equal(band, full[a:b])
        local = {"min": float(core.min()), "max": float(core.max()), "nonfinite": 0, "nonint": rank,
                 "sum": float(core.double().sum()), "sumsq": float((core.double() ** 2).sum()), "n": core.numel()}
        g = bands.global_stats(local, ctx)
        ok = ok and g["min"] == 0.0 and g["max"] == float(gny * nx - 1) and g["n"] == gny * nx
        ok = ok and g["nonint"] == sum(range(world)) and abs(g["sum"] - float(full.double().sum())) < 1e-6
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,gny,halo", [(2, 37, 5), (3, 40, 4), (3, 30, 17), (2, 8, 0)])
def test_halo_exchange_gloo(world, gny, halo):
    """halo 17 over 10-row bands: the halo spans more than the adjacent band (multi-hop case)."""
    port = _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, gny, 6, halo, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(out.get(r) for r in range(world)), dict(out)


def test_single_rank_is_identity():
    ctx = bands.BandContext(12, 5)
    core = torch.zeros(12, 5)
    band, gy0 = bands.exchange_halo(core, ctx, 3)
    assert band is core and gy0 == 0
    st = {"min": 1.0, "max": 2.0, "nonfinite": 0, "nonint": 0, "sum": 3.0, "sumsq": 5.0, "n": 60}
    assert bands.global_stats(st, ctx) == st


def test_round_robin_share_and_numa_binding_are_safe_without_gpus():
    ctx = bands.BandContext(100, 10, 2, 4)
    assert bands.azimuth_share(range(0, 360, 45), ctx) == [90, 270]
    assert bands.azimuth_share([], ctx) == []
    # no NVML device here: best effort, must not raise
    assert bands.bind_to_gpu_numa(0) in (False, True)
