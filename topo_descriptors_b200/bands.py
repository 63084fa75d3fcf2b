"""Row-band sharding of one large DEM over the GPUs of a box.

One process per GPU (``torch.distributed``, NCCL on GPUs, gloo in the CPU tests).  Rank r owns the
contiguous rows ``[r0, r1)`` of the global raster; before a descriptor runs, each rank receives the
``halo`` rows above and below from its neighbours (NVLink send/recv through
``torch.distributed.batch_isend_irecv``) and then calls the same kernels as the single-GPU path with a
``topo_view`` that carries the global geometry, so borders (zero padding, reflect, one-sided
differences, the Sx frame) are evaluated in global coordinates and every pixel is bit-identical to the
single-GPU result.  The only collective besides the neighbour exchange is the tiny all-reduce that
makes the DEM statistics (range -> fixed-point scale, mean/std -> z-score) global.

The reference's own precedent is ``dask.array.map_overlap(depth=2*size)`` in ``tpi`` (topo.py:177-178).
"""

import numpy as np


def partition_rows(gny, world, weights=None):
    """Contiguous row ranges [(r0, r1)] * world.  Balanced by default (the first ``gny % world`` ranks get one more
    row); with ``weights`` the bands are proportional to them (e.g. to each rank's measured device -> host rate when the
    job is bound by getting the results off the GPUs), every rank keeping at least one row."""
    if weights is not None:
        w = np.maximum(np.asarray(weights, dtype=np.float64), 0.0)
        if len(w) != int(world) or not np.isfinite(w).all() or w.sum() <= 0:
            raise ValueError("weights must be one non-negative finite number per rank")
        edges = np.rint(np.cumsum(w) / w.sum() * int(gny)).astype(np.int64)
        edges[-1] = int(gny)
        out, r = [], 0
        for k in range(int(world)):
            e = int(min(max(edges[k], r + 1), int(gny) - (int(world) - 1 - k)))
            out.append((r, e))
            r = e
        return out
    base, extra = divmod(int(gny), int(world))
    out, r = [], 0
    for k in range(world):
        n = base + (1 if k < extra else 0)
        out.append((r, r + n))
        r += n
    return out


class BandContext:
    """Where this rank sits in the row-band decomposition."""

    def __init__(self, gny, nx, rank=0, world=1, weights=None):
        self.gny, self.nx, self.rank, self.world = int(gny), int(nx), int(rank), int(world)
        self.parts = partition_rows(gny, world, weights)
        self.r0, self.r1 = self.parts[rank]

    @property
    def rows(self):
        return self.r1 - self.r0

    def halo_extent(self, halo):
        """Global rows [a, b) this rank needs for a stencil reaching ``halo`` rows up and down."""
        return max(0, self.r0 - halo), min(self.gny, self.r1 + halo)


def exchange_halo(core, ctx, halo, out=None):
    """core: (rows, nx) tensor with this rank's own rows.  Returns (band, gy0): the rows
    [max(0, r0-halo), min(gny, r1+halo)) assembled from the neighbours' edge rows (into ``out`` when given: the
    static input buffer of a captured sweep).

    ``halo`` may span several neighbouring bands (a 20 km scale over thin bands): every rank sends
    to each rank whose extended range overlaps its rows.  Single rank: returns ``core`` itself.
    """
    import torch
    import torch.distributed as dist

    if ctx.world == 1 or halo <= 0:
        if out is not None:
            out.copy_(core)
            return out, ctx.r0
        return core, ctx.r0
    a, b = ctx.halo_extent(halo)
    band = out if out is not None else torch.empty((b - a, ctx.nx), dtype=core.dtype, device=core.device)
    band[ctx.r0 - a : ctx.r1 - a] = core
    ops = []
    keep = []
    for other in range(ctx.world):
        if other == ctx.rank:
            continue
        o0, o1 = ctx.parts[other]
        # rows of `other` that I need
        lo, hi = max(a, o0), min(b, o1)
        if lo < hi:
            ops.append(dist.P2POp(dist.irecv, band[lo - a : hi - a], other))
        # rows of mine that `other` needs
        oa, ob = max(0, o0 - halo), min(ctx.gny, o1 + halo)
        lo, hi = max(oa, ctx.r0), min(ob, ctx.r1)
        if lo < hi:
            chunk = core[lo - ctx.r0 : hi - ctx.r0]
            keep.append(chunk)
            ops.append(dist.P2POp(dist.isend, chunk, other))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return band, a


def global_stats(local, ctx, device=None):
    """All-reduce per-band DEM statistics (device.dem_stats dicts) into global ones."""
    if ctx.world == 1:
        return dict(local)
    import torch
    import torch.distributed as dist

    dev = device if device is not None else "cpu"
    # two collectives and ONE device -> host read: (-min, max) under MAX, the counts and sums under SUM
    ext = torch.tensor([-local["min"], local["max"]], dtype=torch.float64, device=dev)
    sm = torch.tensor([local["nonfinite"], local["nonint"], local["sum"], local["sumsq"], local["n"]],
                      dtype=torch.float64, device=dev)
    dist.all_reduce(ext, op=dist.ReduceOp.MAX)
    dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    out = torch.cat([ext, sm]).cpu().numpy()
    return {"min": float(-out[0]), "max": float(out[1]), "nonfinite": int(out[2]), "nonint": int(out[3]),
            "sum": float(out[4]), "sumsq": float(out[5]), "n": int(out[6])}


# ---------------------------------------------------------------------------------------------
# the multi-scale sweep on one band (what bench.py times; also the engine of sharded compute_*)
# ---------------------------------------------------------------------------------------------
def sweep_halo(sizes, sigmas):
    """Rows of raw DEM a rank needs beyond its own for a TPI/STD sweep over ``sizes`` and a gradient
    sweep over ``sigmas``: disc radius, resp. Gaussian radius + 1 for the central difference."""
    from . import device as dev

    h = 0
    for s in sizes:
        h = max(h, int(s) // 2)
    for sg in sigmas:
        h = max(h, (dev.gauss_radius(sg) + 1) if sg > 1 else 1)
    return h


def sweep(core, ctx, sizes, sigmas, res_x, res_y, what=("tpi", "std", "gradient"), sink=None, stats=None):
    """Run the TPI / STD / gradient multi-scale sweep for this rank's rows.

    core : (rows, nx) float32 CUDA tensor, the rank's own rows of the DEM
    sink : callable(name, scale_index, tensor) receiving every output band (rows r0..r1); default drops
    Returns the number of descriptor calls issued.
    """
    from . import device as dev
    from .device import DeviceDEM

    halo = sweep_halo(sizes if ("tpi" in what or "std" in what) else [], sigmas if "gradient" in what else [])
    band, gy0 = exchange_halo(core, ctx, halo)
    if stats is None:
        stats = global_stats(dev.dem_stats(core), ctx, device=core.device)
    ddem = DeviceDEM(band, gny=ctx.gny, gy0=gy0, stats=stats)
    return _sweep_body(ddem, ctx, sizes, sigmas, res_x, res_y, what, sink)


def _sweep_body(ddem, ctx, sizes, sigmas, res_x, res_y, what, sink):
    """The kernels of one sweep on a band that already holds its halo rows and the global statistics (no host
    synchronisation, no collective: this is what ``SweepGraph`` captures)."""
    from . import device as dev

    if len(sizes) > 1 and ("tpi" in what or "std" in what):
        ddem.share_disc_planes(max(int(s) for s in sizes))  # prefix planes built once for all sizes
    calls = 0
    rx, rx2d = res_x
    ry, ry2d = res_y
    # largest scales first: their kernels run for milliseconds, so the host gets ahead of the GPU right after the
    # statistics read-back instead of feeding it one short launch at a time (matters on thin bands: 8 GPUs)
    for i, size in sorted(enumerate(sizes), key=lambda e: -int(e[1])):
        if "tpi" in what:
            out = dev.tpi(ddem, size, ctx.r0, ctx.rows, pair_std="std" in what)
            calls += 1
            if sink:
                sink("tpi", i, out)
        if "std" in what:
            out = dev.std(ddem, size, ctx.r0, ctx.rows)
            calls += 1
            if sink:
                sink("std", i, out)
    ddem.release_disc_planes()
    if "gradient" in what:
        for i, sigma in sorted(enumerate(sigmas), key=lambda e: -float(e[1])):
            if sigma <= 1:
                outs = dev.sobel_gradient(ddem, rx, rx2d, ry, ry2d, True, ctx.r0, ctx.rows)
            else:
                outs = dev.gradient(ddem, sigma, rx, rx2d, ry, ry2d, ctx.r0, ctx.rows)
            calls += 1
            if sink:
                for nm, o in zip(("dx", "dy", "slope", "aspect"), outs):
                    sink(nm, i, o)
    return calls


class SweepGraph:
    """The kernel sequence of :func:`sweep` captured once in a CUDA graph and replayed per step.

    On thin bands (8 GPUs: 2048 rows) the 30 descriptor calls are ~360 launches of 0.05 - 5 ms each, and Python +
    launch latency become visible next to them; the sequence is static for a given band geometry and DEM statistics
    (range and integrality fix every kernel parameter), so it is captured after one eager pass and replayed.  The halo
    exchange (NCCL) and the statistics all-reduce stay outside the graph, every step.  A DEM whose statistics differ
    re-captures.  ``keep=True`` holds every output band in the graph's memory pool (``outputs[(name, scale index)]``,
    overwritten by each replay); otherwise outputs are dropped as they are produced, like ``sweep`` without a sink.
    """

    def __init__(self, keep=False):
        self.key = None
        self.graph = None
        self.band = None
        self.launches = 0
        self.keep = bool(keep)
        self.outputs = {}

    def run(self, core, ctx, sizes, sigmas, res_x, res_y, what=("tpi", "std", "gradient")):
        import torch

        from . import _lib, device as dev
        from .device import DeviceDEM

        halo = sweep_halo(sizes if ("tpi" in what or "std" in what) else [], sigmas if "gradient" in what else [])
        a, b = ctx.halo_extent(halo) if ctx.world > 1 else (ctx.r0, ctx.r1)
        if self.band is None or tuple(self.band.shape) != (b - a, ctx.nx):
            self.band = torch.empty((b - a, ctx.nx), dtype=core.dtype, device=core.device)
            self.key = None
        band, gy0 = exchange_halo(core, ctx, halo, out=self.band)
        stats = global_stats(dev.dem_stats(core), ctx, device=core.device)
        key = (gy0, tuple(int(s) for s in sizes), tuple(float(s) for s in sigmas), tuple(what), stats["min"], stats["max"],
               stats["nonfinite"] > 0, stats["nonint"] > 0)
        if key != self.key:
            ddem = DeviceDEM(band, gny=ctx.gny, gy0=gy0, stats=stats)
            _sweep_body(ddem, ctx, sizes, sigmas, res_x, res_y, what, None)  # eager: fills the weight / attribute caches
            torch.cuda.synchronize()
            ddem = DeviceDEM(band, gny=ctx.gny, gy0=gy0, stats=stats)
            ddem._assume_fits = True  # no memory queries while capturing
            self.graph = torch.cuda.CUDAGraph()
            self.outputs = {}
            sink = (lambda name, i, t: self.outputs.__setitem__((name, i), t)) if self.keep else None
            n0 = _lib.launch_count()
            with torch.cuda.graph(self.graph):
                _sweep_body(ddem, ctx, sizes, sigmas, res_x, res_y, what, sink)
            self.launches = _lib.launch_count() - n0
            self.key = key
        self.graph.replay()
        return 3 * len(sizes) if len(what) == 3 else None


# ---------------------------------------------------------------------------------------------
# valley / ridge and Sx on one band
# ---------------------------------------------------------------------------------------------
def valley_ridge_band(core, ctx, size, mode, flat_list=(0, 0.15, 0.3), sigma=None):
    """Valley / ridge norm and direction for this rank's rows (reference topo.py:389-453).

    Halo = half the tallest rotated kernel (+ the Gaussian radius when ``sigma`` pre-smooths); the z-score
    uses the GLOBAL mean / std of the (smoothed) DEM: every rank reduces its own rows, one all-reduce.
    Returns (norm, direction) tensors of shape (rows, nx).
    """
    from . import device as dev, topo
    from .device import DeviceDEM

    bank = topo._device_bank(size, mode, list(flat_list), core.device)
    khalo = int(bank["hmax"]) // 2
    lw = dev.gauss_radius(sigma) if sigma else 0
    band, gy0 = exchange_halo(core, ctx, khalo + lw)
    raw_stats = global_stats(dev.dem_stats(core), ctx, device=core.device)
    ddem = DeviceDEM(band, gny=ctx.gny, gy0=gy0, stats=raw_stats)
    s0, s1 = ctx.halo_extent(khalo)  # rows of the (smoothed) DEM the kernels read
    if sigma:
        smooth = dev.gauss(ddem, sigma, sigma, s0, s1 - s0)
        own = smooth[ctx.r0 - s0 : ctx.r1 - s0]
        stats = global_stats(dev.dem_stats(own), ctx, device=core.device)
        src = DeviceDEM(smooth, gny=ctx.gny, gy0=s0, stats=stats)
    else:
        stats = raw_stats
        src = DeviceDEM(band[s0 - gy0 : s1 - gy0], gny=ctx.gny, gy0=s0, stats=stats)
    if stats["nonfinite"] > 0:  # the reference's FFT spreads NaN everywhere: norm = clip(-inf) = 0 (topo.py:441-452)
        zero = dev._new(ctx.rows, ctx.nx, core)
        dev.fill(zero, 0.0)
        return zero, zero.clone()
    mean64 = stats["sum"] / stats["n"]
    var64 = max(stats["sumsq"] / stats["n"] - mean64 * mean64, 0.0)
    normed = dev.zscore(src, np.float32(mean64), np.float32(np.sqrt(var64)))
    return dev.valley_ridge(normed, bank, ctx.r0, ctx.rows)


def sx_band(core, ctx, plan, height=10.0):
    """Sx for this rank's rows and every azimuth of ``plan`` (topo._sx_plan): halo = the sample window.
    Returns a tensor (n_azimuths, rows, nx).  (For many azimuths on a DEM that fits one GPU, dealing the
    azimuths round-robin -- ``azimuth_share`` -- needs no communication at all.)"""
    from . import topo
    from .device import DeviceDEM

    window = int(plan[3])
    band, gy0 = exchange_halo(core, ctx, window)
    ddem = DeviceDEM(band, gny=ctx.gny, gy0=gy0)
    return topo._sx_device(ddem, plan, height, ctx.r0, ctx.rows)


def azimuth_share(azimuths, ctx):
    """Round-robin deal of independent work items (azimuths, scales) over the ranks."""
    return list(azimuths)[ctx.rank :: ctx.world]


def bind_to_gpu_numa(index):
    """Pin this process to the CPU cores NVML reports as local to GPU ``index`` (one process per GPU): pinned
    host buffers are then first-touched on the GPU's own NUMA node, so that the 8 ranks of a box do not push
    their results through one socket.  Returns True when the affinity was set (best effort, never raises)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        handle = None
        try:  # NVML ignores CUDA_VISIBLE_DEVICES: go through the UUID of the CUDA device when torch exposes it
            import torch

            handle = pynvml.nvmlDeviceGetHandleByUUID("GPU-" + str(torch.cuda.get_device_properties(int(index)).uuid))
        except Exception:
            handle = pynvml.nvmlDeviceGetHandleByIndex(int(index))
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        return True
    except Exception:  # no NVML, no permission, or a single-socket host: nothing to do
        return False
