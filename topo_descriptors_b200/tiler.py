"""Out-of-core row-band tiler: descriptors of a DEM that does not fit (or should not sit) in HBM.

The reference's only mechanism for rasters beyond memory is ``dask.array.map_overlap(conv_fn, dem.data,
depth=size*2, boundary="none")`` in ``tpi`` (topo.py:177-178): overlapping chunks, each convolved on its own.
This module is that idea for every descriptor of the path, on one GPU: the DEM stays on the host (a numpy array,
a ``numpy.memmap``, or anything 2-D that slices to something ``numpy.asarray`` accepts -- a dask array works the
same way, ``asarray`` triggers its ``compute``), contiguous row bands are streamed through HBM with the halo the
stencil needs, and every band runs the same kernels as a resident DEM with a ``topo_view`` that carries the global
geometry -- so zero padding, reflection, one-sided differences and the Sx frame are evaluated in global coordinates
and the assembled result is bit-identical to the single-pass one (unlike ``map_overlap(boundary="none")``, whose
chunk edges see a truncated neighbourhood unless ``depth`` over-covers the kernel).

Pass 1 streams the DEM once for the global statistics (range and integrality fix the fixed-point planes of TPI /
STD, the all-NaN rule and the valley/ridge z-score); pass 2 computes band by band.  Uploads run on a second stream
from pinned staging buffers, one band ahead of the kernels; results land in a caller-supplied array (e.g. a
``numpy.memmap``) or a fresh host array.
"""

import numpy as np

from . import device as dev
from .device import DeviceDEM


def band_plan(gny, halo, band_rows):
    """[(r0, r1, a, b)]: output rows [r0, r1) of each band and the input rows [a, b) it reads (halo clipped to
    the image).  ``band_rows`` output rows per band (the last band may be shorter)."""
    gny, halo, band_rows = int(gny), max(0, int(halo)), int(band_rows)
    if band_rows <= 0:
        raise ValueError("band_rows must be positive")
    out = []
    r0 = 0
    while r0 < gny:
        r1 = min(gny, r0 + band_rows)
        out.append((r0, r1, max(0, r0 - halo), min(gny, r1 + halo)))
        r0 = r1
    return out


def merge_stats(parts):
    """Combine per-band DEM statistics (device.dem_stats dicts) into the statistics of the whole raster."""
    return {
        "min": min(p["min"] for p in parts), "max": max(p["max"] for p in parts),
        "nonfinite": sum(p["nonfinite"] for p in parts), "nonint": sum(p["nonint"] for p in parts),
        "sum": float(np.sum([p["sum"] for p in parts])), "sumsq": float(np.sum([p["sumsq"] for p in parts])),
        "n": sum(p["n"] for p in parts),
    }


def default_band_rows(nx, halo, bytes_per_px=96, budget_bytes=None):
    """Output rows per band so that a band with its halo and the descriptor's scratch (``bytes_per_px``: DEM, outputs
    and the widest workspace of the path, ~96 B/px for a cached disc sweep) stays inside ``budget_bytes`` (default:
    half of the free HBM)."""
    if budget_bytes is None:
        torch = dev.require_cuda()
        free, _total = torch.cuda.mem_get_info()
        budget_bytes = free // 2
    rows = int(budget_bytes // (int(bytes_per_px) * int(nx))) - 2 * int(halo)
    return max(rows, max(64, int(halo) // 4))


class HostDEM:
    """A 2-D host raster to be processed out of core: remembers the global statistics between descriptor calls."""

    def __init__(self, array_like, band_rows=None):
        shape = tuple(array_like.shape)
        if len(shape) != 2:
            raise ValueError("dem must be 2-D")
        self.src = array_like
        self.gny, self.nx = int(shape[0]), int(shape[1])
        self.band_rows = band_rows
        self._stats = None

    def rows(self, a, b):
        """Rows [a, b) as a contiguous float32 host array (float64 / lazy inputs are converted here, band by band)."""
        rows = np.ascontiguousarray(np.asarray(self.src[a:b]), dtype=np.float32)
        return rows if rows.flags.writeable else rows.copy()  # (a read-only memmap view: torch wants a writable buffer)

    @property
    def stats(self):
        if self._stats is None:
            step = self.band_rows or default_band_rows(self.nx, 0, bytes_per_px=8)
            parts = [dev.dem_stats(dev.to_device(self.rows(a, min(self.gny, a + step)))) for a in range(0, self.gny, step)]
            self._stats = merge_stats(parts)
        return self._stats


def run(dem, halo, fn, n_out=1, out=None, band_rows=None, out_dtype=np.float32, need_stats=True, lead=None):
    """Stream ``dem`` (HostDEM or array-like) through ``fn(ddem, r0, rows) -> tensor | [tensors]`` band by band.

    ``halo``: rows of raw DEM the stencil reaches above / below an output row.  ``out``: host array(s) to fill
    (created when None).  ``lead``: leading shape of each output (e.g. ``(n_az,)`` for Sx: tensors are then
    ``lead + (rows, nx)``).  Returns the list of host arrays (or the single array when ``n_out == 1``).
    """
    torch = dev.require_cuda()
    host = dem if isinstance(dem, HostDEM) else HostDEM(dem, band_rows)
    rows_per_band = band_rows or host.band_rows or default_band_rows(host.nx, halo)
    plan = band_plan(host.gny, halo, rows_per_band)
    stats = host.stats if need_stats else None
    lead = tuple(lead or ())
    outs = out
    if outs is None:
        outs = [np.empty(lead + (host.gny, host.nx), dtype=out_dtype) for _ in range(n_out)]
    elif n_out == 1 and not isinstance(outs, (list, tuple)):
        outs = [outs]
    copy_stream = torch.cuda.Stream()
    main = torch.cuda.current_stream()

    def upload(i):
        _r0, _r1, a, b = plan[i]
        pinned = torch.from_numpy(host.rows(a, b)).pin_memory()
        with torch.cuda.stream(copy_stream):
            t = pinned.to("cuda", non_blocking=True)
            done = copy_stream.record_event()
        return t, done, pinned

    nxt = upload(0)
    for i, (r0, r1, a, _b) in enumerate(plan):
        band, ready, keep = nxt
        if i + 1 < len(plan):
            nxt = upload(i + 1)  # the next band's H2D runs under this band's kernels
        main.wait_event(ready)
        band.record_stream(main)
        ddem = DeviceDEM(band, gny=host.gny, gy0=a, stats=stats)
        res = fn(ddem, r0, r1 - r0)
        res = list(res) if isinstance(res, (list, tuple)) else [res]
        for o, t in zip(outs, res):
            o[..., r0:r1, :] = t.cpu().numpy().astype(out_dtype, copy=False)
        del ddem, band, keep
    return outs[0] if n_out == 1 else list(outs)


# ---------------------------------------------------------------------------------------------
# the descriptors (same arguments as topo.*; host array-like in, host arrays out)
# ---------------------------------------------------------------------------------------------
def tpi(dem, size, band_rows=None, out=None):
    """topo.tpi out of core (topo.py:144-181; no pre-smoothing: smooth with :func:`gauss` first)."""
    return run(dem, int(size) // 2, lambda d, r0, n: dev.tpi(d, int(size), r0, n, share=False), out=out, band_rows=band_rows)


def std(dem, size, band_rows=None, out=None):
    """topo.std out of core (topo.py:272-307); float64 like the reference."""
    return run(dem, int(size) // 2, lambda d, r0, n: dev.std(d, int(size), r0, n, share=False), out=out, band_rows=band_rows,
               out_dtype=np.float64)


def gauss(dem, sigma, band_rows=None, out=None):
    """topo.dem out of core (topo.py:62-80)."""
    sig = (sigma, sigma) if np.isscalar(sigma) else tuple(sigma)
    return run(dem, dev.gauss_radius(sig[0]), lambda d, r0, n: dev.gauss(d, sig[0], sig[1], r0, n), out=out, band_rows=band_rows)


def gradient(dem, sigma, res_meters, band_rows=None, out=None):
    """topo.gradient (sig_ratio = 1) out of core (topo.py:597-644): [dx, dy, slope, aspect]."""
    torch = dev.require_cuda()
    device = torch.device("cuda", torch.cuda.current_device())
    rx, ry = dev._Res(res_meters["x"], device), dev._Res(res_meters["y"], device)
    if sigma <= 1:
        return run(dem, 1, lambda d, r0, n: dev.sobel_gradient(d, rx, rx.is_2d, ry, ry.is_2d, True, r0, n), n_out=4, out=out,
                   band_rows=band_rows)
    return run(dem, dev.gauss_radius(sigma) + 1, lambda d, r0, n: dev.gradient(d, sigma, rx, rx.is_2d, ry, ry.is_2d, r0, n),
               n_out=4, out=out, band_rows=band_rows)


def sweep(dem, sizes, band_rows=None):
    """TPI and STD at several sizes with ONE pass over the DEM per band: the size-independent prefix planes are built
    once per band (``share_disc_planes``) and every size walks them.  Returns {size: (tpi, std)}."""
    sizes = [int(s) for s in sizes]
    host = dem if isinstance(dem, HostDEM) else HostDEM(dem, band_rows)
    res = {s: (np.empty((host.gny, host.nx), np.float32), np.empty((host.gny, host.nx), np.float64)) for s in sizes}

    def fn(d, r0, n):
        if len(sizes) > 1:
            d.share_disc_planes(max(sizes))
        for s in sizes:
            res[s][0][r0 : r0 + n] = dev.tpi(d, s, r0, n, pair_std=True).cpu().numpy()
            res[s][1][r0 : r0 + n] = dev.std(d, s, r0, n).cpu().numpy()
        d.release_disc_planes()
        return []

    run(host, max(sizes) // 2, fn, n_out=0, out=[], band_rows=band_rows)
    return res


def valley_ridge(dem, size, mode, flat_list=(0, 0.15, 0.3), band_rows=None):
    """topo.valley_ridge out of core (topo.py:389-453, no pre-smoothing): the z-score uses the global statistics of
    pass 1."""
    from . import topo

    host = dem if isinstance(dem, HostDEM) else HostDEM(dem, band_rows)
    st = host.stats
    if st["nonfinite"] > 0:
        zero = np.zeros((host.gny, host.nx), np.float32)
        return [zero, zero.copy()]
    mean64 = st["sum"] / st["n"]
    sd = np.sqrt(max(st["sumsq"] / st["n"] - mean64 * mean64, 0.0))
    torch = dev.require_cuda()
    bank = topo._device_bank(size, mode, list(flat_list), torch.device("cuda", torch.cuda.current_device()))

    def fn(d, r0, n):
        normed = dev.zscore(d, np.float32(mean64), np.float32(sd))
        return dev.valley_ridge(normed, bank, r0, n)

    return run(host, int(bank["hmax"]) // 2, fn, n_out=2, band_rows=band_rows)


def sx(dem_ds, azimuth, radius, height=10.0, azimuth_arc=10.0, azimuth_steps=15, radius_min=0.0, band_rows=None):
    """topo.sx out of core (topo.py:775-858): ``dem_ds`` is a Dataset whose DEM variable may be any host array-like."""
    from . import _xr, helpers as hlp, topo

    if not _xr.is_dataset(dem_ds):
        raise TypeError("Argument 'dem_ds' must be a xr.Dataset.")
    many = hasattr(azimuth, "__iter__")
    centres = list(azimuth) if many else [azimuth]
    plan = topo._sx_plan(dem_ds, centres, radius, azimuth_arc, azimuth_steps, radius_min)
    values = hlp.get_da(dem_ds).values
    out = run(values, int(plan[3]), lambda d, r0, n: topo._sx_device(d, plan, height, r0, n), band_rows=band_rows,
              need_stats=False, lead=(len(centres),))
    return out if many else out[0]
