"""Device-side plumbing: a DEM (or a row band of one) resident in HBM, and thin wrappers that launch
the kernels of libtopo_b200.so on torch-owned buffers.

PyTorch is used only for device memory, streams and (in ``bands.py``) ``torch.distributed``; every
computation below is a call through the C ABI (``include/topo_b200.h``).  No CPU fallback exists: a
missing library or GPU raises.
"""

import ctypes

import numpy as np

from . import _lib
from ._lib import View


def _torch():
    import torch

    return torch


def require_cuda():
    torch = _torch()
    if not torch.cuda.is_available():
        raise RuntimeError("topo_descriptors_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    _lib.load()
    return torch


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream():
    return ctypes.c_void_p(_torch().cuda.current_stream().cuda_stream)


def to_device(array, pin=False):
    """Host ndarray -> contiguous float32 CUDA tensor (H2D on the current stream)."""
    torch = require_cuda()
    a = np.ascontiguousarray(array, dtype=np.float32)
    t = torch.from_numpy(a)
    if pin:
        t = t.pin_memory()
    return t.to("cuda", non_blocking=pin)


class DeviceDEM:
    """A float32 raster band in HBM.

    ``tensor``: (rows, nx) CUDA tensor holding global rows [gy0, gy0 + rows) of an image that is
    ``gny`` rows tall (``gy0 = 0, gny = rows`` for a whole DEM).  ``stats`` are the GLOBAL DEM
    statistics (lazily computed for a whole DEM; must be supplied for a band so that every band uses
    the same fixed-point scale and z-score).
    """

    def __init__(self, tensor, gny=None, gy0=0, stats=None):
        torch = require_cuda()
        if isinstance(tensor, np.ndarray):
            tensor = to_device(tensor)
        if tensor.dtype != torch.float32 or tensor.dim() != 2 or not tensor.is_cuda:
            raise TypeError("DeviceDEM needs a 2-D float32 CUDA tensor")
        if tensor.stride(1) != 1:
            tensor = tensor.contiguous()
        self.tensor = tensor
        self.rows, self.nx = int(tensor.shape[0]), int(tensor.shape[1])
        self.ld = int(tensor.stride(0))
        self.gy0 = int(gy0)
        self.gny = int(self.rows if gny is None else gny)
        self._stats = stats

    is_device_dem = True  # lets the Dataset container keep it as the values of the DEM variable

    def share_disc_planes(self, max_size):
        """Announce tpi / std calls at several sizes up to ``max_size`` on this band: the size-independent prefix
        planes are then built once and kept (7 B/px x 2 planes of HBM at most) until ``release_disc_planes``."""
        self._plane_hint = int(max_size)
        return self

    def release_disc_planes(self):
        self._plane_hint = 0
        self._plane_cache = None
        self._tsum = None

    @property
    def shape(self):
        return (self.rows, self.nx)

    @property
    def dtype(self):
        return np.dtype(np.float32)

    def numpy(self):
        """One D2H copy of the band."""
        return self.tensor.cpu().numpy()

    @property
    def is_whole(self):
        return self.gy0 == 0 and self.rows == self.gny

    @property
    def stats(self):
        if self._stats is None:
            if not self.is_whole:
                raise RuntimeError("a row band needs the global DEM statistics (see bands.global_stats)")
            self._stats = dem_stats(self.tensor)
        return self._stats

    def view(self, out_gy0=None, out_rows=None):
        """topo_view computing global rows [out_gy0, out_gy0+out_rows) from this band."""
        if out_gy0 is None:
            out_gy0, out_rows = self.gy0, self.rows
        return View(self.nx, self.gny, self.gy0, self.rows, int(out_gy0), int(out_rows))


def dem_stats(tensor):
    """min / max / non-finite / non-integer / sum / sumsq / n of a device raster (one D2H of 64 B)."""
    torch = require_cuda()
    rows, nx = int(tensor.shape[0]), int(tensor.shape[1])
    L = _lib.load()
    ws_bytes = L.topo_dem_stats_workspace_bytes(rows, nx)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=tensor.device)
    out = torch.empty(8, dtype=torch.float64, device=tensor.device)
    _lib.call("topo_dem_stats_f32", _ptr(tensor), rows, nx, int(tensor.stride(0)), _ptr(out), _ptr(ws), ws_bytes,
              _stream())
    s = out.cpu().numpy()
    return {
        "min": float(s[0]), "max": float(s[1]), "nonfinite": int(s[2]), "nonint": int(s[3]),
        "sum": float(s[4]), "sumsq": float(s[5]), "n": int(s[6]),
    }


def _new(rows, nx, like):
    return _torch().empty((rows, nx), dtype=_torch().float32, device=like.device)


def fill(t, value):
    _lib.call("topo_fill_f32", _ptr(t), int(t.shape[0]), int(t.shape[1]), int(t.stride(0)), float(value), _stream())


def fill_na(tensor, x_dev=None, mask_below=None, want_indices=True):
    """Device pre-stage of helpers.py:17-31 + 137-154: cells that are NaN (or <= ``mask_below``) are replaced by
    the nearest valid cell of their row (ties: lower x).  Returns (filled tensor, (rows, cols) int32 device
    tensors of the missing cells in np.where order or None, number of missing cells)."""
    torch = require_cuda()
    rows, nx = int(tensor.shape[0]), int(tensor.shape[1])
    out = torch.empty((rows, nx), dtype=torch.float32, device=tensor.device)
    counts = torch.empty((max(rows, 1),), dtype=torch.int32, device=tensor.device)
    use_mask = 0 if mask_below is None else 1
    thr = 0.0 if mask_below is None else float(mask_below)
    _lib.call("topo_fill_na_f32", _ptr(tensor), int(tensor.stride(0)), _ptr(out), int(out.stride(0)), rows, nx,
              _ptr(x_dev), use_mask, thr, _ptr(counts), _stream())
    offsets = torch.empty((rows + 1,), dtype=torch.int64, device=tensor.device)
    _lib.call("topo_nan_indices_f32", _ptr(tensor), int(tensor.stride(0)), rows, nx, use_mask, thr, _ptr(counts),
              _ptr(offsets), None, None, _stream())
    total = int(offsets[-1].item())
    if not want_indices or total == 0:
        return out, None, total
    r = torch.empty((total,), dtype=torch.int32, device=tensor.device)
    c = torch.empty((total,), dtype=torch.int32, device=tensor.device)
    _lib.call("topo_nan_indices_f32", _ptr(tensor), int(tensor.stride(0)), rows, nx, use_mask, thr, _ptr(counts),
              _ptr(offsets), _ptr(r), _ptr(c), _stream())
    return out, (r, c), total


def stamp(t, rows_dev, cols_dev, value=float("nan")):
    """t[rows, cols] = value on device (``array[ind_nans] = np.nan`` of the compute_* drivers)."""
    n = int(rows_dev.numel())
    if n:
        _lib.call("topo_stamp_f32", _ptr(t), int(t.stride(0)), _ptr(rows_dev), _ptr(cols_dev), n, float(value),
                  _stream())


# ---- Gaussian weights, exactly scipy's _gaussian_kernel1d (scipy/ndimage/_filters.py) -------------
_WEIGHT_CACHE = {}


def gaussian_half_kernel(sigma, truncate=4.0):
    """(w, lw): w[0] centre .. w[lw]; float64, normalised by the float64 sum of the full kernel."""
    sigma = float(sigma)
    lw = int(truncate * sigma + 0.5)
    x = np.arange(-lw, lw + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x**2)
    phi = phi / phi.sum()
    return np.ascontiguousarray(phi[lw:]), lw


def _device_weights(sigma, device):
    if sigma is None or float(sigma) <= 1e-15:
        return None, 0
    key = (float(sigma), str(device))
    hit = _WEIGHT_CACHE.get(key)
    if hit is None:
        w, lw = gaussian_half_kernel(sigma)
        hit = (_torch().from_numpy(w).to(device), lw)
        if len(_WEIGHT_CACHE) > 256:
            _WEIGHT_CACHE.clear()
        _WEIGHT_CACHE[key] = hit
    return hit


def gauss_radius(sigma):
    return 0 if sigma is None or float(sigma) <= 1e-15 else int(4.0 * float(sigma) + 0.5)


def gauss(dem, sigma_y, sigma_x, out_gy0=None, out_rows=None, out=None):
    """ndimage.gaussian_filter(dem, (sigma_y, sigma_x)) for global rows [out_gy0, out_gy0+out_rows)."""
    torch = require_cuda()
    v = dem.view(out_gy0, out_rows)
    wy, lwy = _device_weights(sigma_y, dem.tensor.device)
    wx, lwx = _device_weights(sigma_x, dem.tensor.device)
    if out is None:
        out = _new(v.out_rows, dem.nx, dem.tensor)
    L = _lib.load()
    ws_bytes = L.topo_gauss_workspace_bytes(ctypes.byref(v), lwy, lwx)
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dem.tensor.device)
    # NaN-exact taps only when the DEM is known (or not yet known) to hold non-finite values
    st = dem._stats if dem._stats is not None else (dem.stats if dem.is_whole else None)
    nan_safe = 1 if (st is None or st["nonfinite"] > 0) else 0
    _lib.call("topo_gauss_f32", _ptr(dem.tensor), dem.ld, _ptr(out), int(out.stride(0)), ctypes.byref(v), _ptr(wy), lwy,
              _ptr(wx), lwx, nan_safe, _ptr(ws), ws_bytes, _stream())
    return out


def _nan_result(v, like, n=1):
    outs = [_new(v.out_rows, v.nx, like) for _ in range(n)]
    for o in outs:
        fill(o, float("nan"))
    return outs


def _tsum_plan(dem, v, size, st, share, cache_size=0, pair=False):
    """Plane-sum sharing between tpi(size) and std(size) of the same DEM band (T plane of integer-valued DEMs; T and
    fraction planes of float DEMs, only when the caller announced the pair -- a lone float tpi is cheaper unpaired):
    returns (tensor or None, op, keep) with op 0 = off, 1 = compute + keep, 2 = reuse; ``keep`` is what the caller
    publishes as ``dem._tsum`` AFTER the launch succeeded (a failed call must not leave an unwritten buffer behind)."""
    torch = _torch()
    integer = st["nonint"] == 0
    if not share or not (integer or pair):
        return None, 0, None
    L = _lib.load()
    planes = L.topo_disc_shares_tsum(ctypes.byref(v), int(size), 1 if integer else 0, st["min"], st["max"], int(cache_size))
    if not planes:
        return None, 0, None
    key = (int(size), v.out_gy0, v.out_rows)
    cached = getattr(dem, "_tsum", None)
    dem._tsum = None  # consumed (a tpi+std pair is the use case), or replaced once the new sums exist
    if cached is not None and cached[0] == key:
        return cached[1], 2, None
    t = torch.empty((planes, v.out_rows, dem.nx), dtype=torch.int64, device=dem.tensor.device)
    return t, 1, (key, t)


def _disc(name, what, dem, size, out_gy0, out_rows, out, share, pair=False):
    torch = require_cuda()
    v = dem.view(out_gy0, out_rows)
    st = dem.stats
    if out is None:
        out = _new(v.out_rows, dem.nx, dem.tensor)
    if st["nonfinite"] > 0:
        fill(out, float("nan"))
        return out
    L = _lib.load()
    cache = _plane_cache(dem, v, size, st)
    cache_size = cache.max_size if cache is not None else 0
    integer = 1 if st["nonint"] == 0 else 0
    # std of a float DEM reuses what a paired tpi left behind; it never starts a pair itself
    tsum, op, keep = _tsum_plan(dem, v, size, st, share, cache_size, pair or (what == 1 and getattr(dem, "_tsum", None) is not None))
    if what == 1 and not integer and op == 1:
        tsum, op, keep = None, 0, None
    ws_bytes = L.topo_disc_workspace_bytes(ctypes.byref(v), int(size), what, integer, st["min"], st["max"], cache_size, op)
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dem.tensor.device)
    _lib.call(name, _ptr(dem.tensor), dem.ld, _ptr(out), int(out.stride(0)), ctypes.byref(v), int(size),
              integer, st["min"], st["max"], _ptr(tsum), op,
              ctypes.byref(cache) if cache is not None else None, _ptr(ws), ws_bytes, _stream())
    if keep is not None:
        dem._tsum = keep
    return out


def _fits_in_hbm(nbytes, device):
    """True if ``nbytes`` can still be allocated on ``device`` (free HBM + what torch's allocator holds unused).
    Fails open: any problem with the query means "yes"."""
    try:
        torch = _torch()
        free, _total = torch.cuda.mem_get_info(device)
        pooled = torch.cuda.memory_reserved(device) - torch.cuda.memory_allocated(device)
        return int(nbytes) <= int(free) + max(int(pooled), 0)
    except Exception:
        return True


def _plane_cache(dem, v, size, st):
    """The ``topo_disc_cache`` of this band, if the caller announced a multi-scale sweep with
    ``dem.share_disc_planes(max_size)``: the size-independent prefix planes are then built once per sweep."""
    hint = getattr(dem, "_plane_hint", 0)
    if not hint or size > hint:
        return None
    key = (v.in_gy0, v.in_rows, v.out_gy0, v.out_rows, hint)
    held = getattr(dem, "_plane_cache", None)
    if held is not None and held[0] == key:
        return held[1]
    halo = hint // 2  # the band must cover the halo of the largest disc
    if v.in_gy0 > max(0, v.out_gy0 - halo) or v.in_gy0 + v.in_rows < min(v.gny, v.out_gy0 + v.out_rows + halo):
        return None
    nbytes = _lib.load().topo_disc_cache_bytes(ctypes.byref(v), int(hint), 1 if st["nonint"] == 0 else 0, st["min"],
                                               st["max"])
    if nbytes == 0 or not (getattr(dem, "_assume_fits", False) or _fits_in_hbm(2 * nbytes, dem.tensor.device)):
        return None  # (a DEM too large for shared planes runs every size on its own workspace, as without the hint)
    mem = _torch().empty(nbytes + 256, dtype=_torch().uint8, device=dem.tensor.device)
    base = (mem.data_ptr() + 255) & ~255
    cache = _lib.DiscCache(ctypes.c_void_p(base), nbytes, int(hint), 0)
    dem._plane_cache = (key, cache, mem)  # `mem` keeps the allocation alive
    return cache


def tpi(dem, size, out_gy0=None, out_rows=None, out=None, share=True, pair_std=False):
    """Device TPI of global rows [out_gy0, out_gy0+out_rows); all-NaN if the DEM has non-finite values
    (the reference's FFT convolution spreads them over the whole output).  ``share``: on integer-valued
    DEMs keep / reuse the disc sums that tpi(size) and std(size) have in common (one gather pass less per pair).
    ``pair_std``: a std(size) call follows on this DEM -- float DEMs then run the exact two-plane TPI whose plane
    sums the std reuses (3 gather passes per pair instead of 4)."""
    return _disc("topo_tpi_f32", 0, dem, size, out_gy0, out_rows, out, share, pair_std)


def std(dem, size, out_gy0=None, out_rows=None, out=None, share=True):
    """Device STD (float32; the host shim up-casts to float64 like the reference)."""
    return _disc("topo_std_f32", 1, dem, size, out_gy0, out_rows, out, share)


class _Res:
    """Grid resolution array on the device: float64 (reference semantics) plus, when every value is exactly
    representable in float32, a float32 copy that lets the gradient kernels use their vector path."""

    def __init__(self, res, device):
        r = np.ascontiguousarray(np.asarray(res, dtype=np.float64))
        self.f64 = _torch().from_numpy(r).to(device)
        self.is_2d = int(r.ndim == 2)
        r32 = r.astype(np.float32)
        self.f32 = _torch().from_numpy(r32).to(device) if np.array_equal(r32.astype(np.float64), r) else None


def _res_to_device(res, device):
    """Back-compatible helper: (float64 device tensor, is_2d)."""
    r = _Res(res, device)
    return r.f64, r.is_2d


def _as_res(x, is_2d=None):
    if isinstance(x, _Res):
        return x
    r = _Res.__new__(_Res)
    r.f64, r.is_2d = x, int(x.dim() == 2 if is_2d is None else is_2d)
    f32 = x.to(_torch().float32)
    r.f32 = f32 if bool((f32.to(_torch().float64) == x).all()) else None
    return r


def gradient_from_smooth(gx, gy, res_x_dev, res_x_2d, res_y_dev, res_y_2d, out_gy0=None, out_rows=None):
    """[dx, dy, slope, aspect] from smoothed bands gx (d/dx) and gy (d/dy) (same band geometry).
    res_*_dev: `_Res` objects (preferred) or float64 device tensors."""
    require_cuda()
    v = gx.view(out_gy0, out_rows)
    rx, ry = _as_res(res_x_dev, res_x_2d), _as_res(res_y_dev, res_y_2d)
    f32 = rx.f32 is not None and ry.f32 is not None
    outs = [_new(v.out_rows, gx.nx, gx.tensor) for _ in range(4)]
    _lib.call("topo_grad_from_smooth_f32", _ptr(gx.tensor), _ptr(gy.tensor), gx.ld, _ptr(outs[0]), _ptr(outs[1]),
              _ptr(outs[2]), _ptr(outs[3]), int(outs[0].stride(0)), ctypes.byref(v), _ptr(rx.f64), rx.is_2d,
              _ptr(ry.f64), ry.is_2d, _ptr(rx.f32 if f32 else None), _ptr(ry.f32 if f32 else None), _stream())
    return outs


def gradient(dem, sigma, res_x_dev, res_x_2d, res_y_dev, res_y_2d, out_gy0=None, out_rows=None):
    """[dx, dy, slope, aspect] of the Gaussian-smoothed DEM (isotropic sigma > 1) for global rows
    [out_gy0, out_gy0+out_rows): one fused kernel for radii up to 21 px, else smoothing + differences inside the
    library's workspace.  The band must reach ``gauss_radius(sigma) + 1`` rows beyond the output rows."""
    torch = require_cuda()
    v = dem.view(out_gy0, out_rows)
    w, lw = _device_weights(sigma, dem.tensor.device)
    rx, ry = _as_res(res_x_dev, res_x_2d), _as_res(res_y_dev, res_y_2d)
    f32 = rx.f32 is not None and ry.f32 is not None
    st = dem._stats if dem._stats is not None else (dem.stats if dem.is_whole else None)
    nan_safe = 1 if (st is None or st["nonfinite"] > 0) else 0
    outs = [_new(v.out_rows, dem.nx, dem.tensor) for _ in range(4)]
    ws_bytes = _lib.load().topo_gradient_workspace_bytes(ctypes.byref(v), lw)
    ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device=dem.tensor.device)
    base = (ws.data_ptr() + 255) & ~255
    _lib.call("topo_gradient_f32", _ptr(dem.tensor), dem.ld, _ptr(outs[0]), _ptr(outs[1]), _ptr(outs[2]), _ptr(outs[3]),
              int(outs[0].stride(0)), ctypes.byref(v), _ptr(w), lw, nan_safe, _ptr(rx.f64), rx.is_2d, _ptr(ry.f64), ry.is_2d,
              _ptr(rx.f32 if f32 else None), _ptr(ry.f32 if f32 else None), ctypes.c_void_p(base), ws_bytes, _stream())
    return outs


def sobel_gradient(dem, res_x_dev=None, res_x_2d=0, res_y_dev=None, res_y_2d=0, normalize=True, out_gy0=None,
                   out_rows=None):
    """Sobel derivatives; with ``normalize`` also the resolution division + slope + aspect."""
    require_cuda()
    v = dem.view(out_gy0, out_rows)
    n_out = 4 if normalize else 2
    outs = [_new(v.out_rows, dem.nx, dem.tensor) for _ in range(n_out)]
    null = ctypes.c_void_p(0)
    if normalize:
        rx, ry = _as_res(res_x_dev, res_x_2d), _as_res(res_y_dev, res_y_2d)
        f32 = rx.f32 is not None and ry.f32 is not None
        res_args = (_ptr(rx.f64), rx.is_2d, _ptr(ry.f64), ry.is_2d, _ptr(rx.f32 if f32 else None),
                    _ptr(ry.f32 if f32 else None))
    else:
        res_args = (null, 0, null, 0, null, null)
    _lib.call("topo_sobel_gradient_f32", _ptr(dem.tensor), dem.ld, _ptr(outs[0]), _ptr(outs[1]),
              _ptr(outs[2]) if normalize else null, _ptr(outs[3]) if normalize else null,
              int(outs[0].stride(0)), ctypes.byref(v), *res_args, 1 if normalize else 0, _stream())
    return outs


def sx(dem, offsets_dev, inv_dist_dev, az_begin_dev, n_az, window, height, extents, out_gy0=None, out_rows=None):
    """Sx for n_az azimuth sectors in one launch -> tensor (n_az, out_rows, nx).
    extents = (dy_min, dy_max, dx_min, dx_max) over all samples."""
    torch = require_cuda()
    v = dem.view(out_gy0, out_rows)
    out = torch.empty((n_az, v.out_rows, dem.nx), dtype=torch.float32, device=dem.tensor.device)
    dy_min, dy_max, dx_min, dx_max = (int(e) for e in extents)
    _lib.call("topo_sx_f32", _ptr(dem.tensor), dem.ld, _ptr(out), int(out.stride(1)), int(out.stride(0)),
              ctypes.byref(v), _ptr(offsets_dev), _ptr(inv_dist_dev), _ptr(az_begin_dev), int(n_az),
              int(window), float(height), dy_min, dy_max, dx_min, dx_max, _stream())
    return out


def zscore(dem, mean, sd):
    """(dem - mean) / std in float32 (topo.py:429) -> DeviceDEM with the same band geometry."""
    require_cuda()
    out = _new(dem.rows, dem.nx, dem.tensor)
    _lib.call("topo_zscore_f32", _ptr(dem.tensor), dem.ld, _ptr(out), int(out.stride(0)), dem.rows, dem.nx,
              float(mean), float(sd), _stream())
    return DeviceDEM(out, gny=dem.gny, gy0=dem.gy0, stats=dem._stats)


VALLEY_FFT_MIN_EXTENT = 52  # rotated kernels at least this tall / wide (size >= ~37) go through the FFT route: measured on
# B200 per 2048^2, 180 angles x 3 flats: FFT ~180 ms at any size; direct bank 65 ms (size 21), 221 (41), 465 (61), 1144 (81)
VALLEY_DEVICE_ROTATION = True  # ... and their bank is rotated on the GPU (host scipy otherwise)


def valley_ridge_fft(dem_norm, bank, out_gy0=None, out_rows=None):
    """Running (max, argmax) over the bank by 2-D overlap-save FFT convolution: cost independent of the kernel size."""
    torch = require_cuda()
    v = dem_norm.view(out_gy0, out_rows)
    norm = _new(v.out_rows, dem_norm.nx, dem_norm.tensor)
    direction = _new(v.out_rows, dem_norm.nx, dem_norm.tensor)
    plain = bank["plain"]
    hmax, wmax = int(bank["hmax"]), int(bank["wmax"])
    ws_bytes = _lib.load().topo_valley_ridge_fft_workspace_bytes(ctypes.byref(v), hmax, wmax)
    ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device=dem_norm.tensor.device)
    base = (ws.data_ptr() + 255) & ~255
    off = np.ascontiguousarray(plain["off"], dtype=np.int64)
    hw = np.ascontiguousarray(plain["hw"], dtype=np.int32)
    _lib.call("topo_valley_ridge_fft_f32", _ptr(dem_norm.tensor), dem_norm.ld, _ptr(norm), _ptr(direction), int(norm.stride(0)),
              ctypes.byref(v), _ptr(plain["data"]), off.ctypes.data_as(ctypes.c_void_p), hw.ctypes.data_as(ctypes.c_void_p),
              int(len(off)), hmax, wmax, ctypes.c_void_p(base), ws_bytes, _stream())
    return norm, direction


def valley_ridge(dem_norm, bank, out_gy0=None, out_rows=None):
    """Running (max, argmax) over the rotated-kernel bank -> (norm, dir) tensors.  ``bank``: one packed bank, or a dict
    whose ``"groups"`` lists several (flat lists longer than 4 run in groups of 4 channels sharing the running maximum)."""
    require_cuda()
    packed = "data" in bank or "groups" in bank  # a bank rotated on the device only has the FFT route's layout
    if "plain" in bank and (not packed or max(int(bank["hmax"]), int(bank["wmax"])) >= VALLEY_FFT_MIN_EXTENT):
        return valley_ridge_fft(dem_norm, bank, out_gy0, out_rows)
    v = dem_norm.view(out_gy0, out_rows)
    norm = _new(v.out_rows, dem_norm.nx, dem_norm.tensor)
    direction = _new(v.out_rows, dem_norm.nx, dem_norm.tensor)
    groups = bank["groups"] if "groups" in bank else [bank]
    for g, b in enumerate(groups):
        flags = (1 if g > 0 else 0) | (2 if g + 1 < len(groups) else 0)
        _lib.call("topo_valley_ridge_f32", _ptr(dem_norm.tensor), dem_norm.ld, _ptr(norm), _ptr(direction),
                  int(norm.stride(0)), ctypes.byref(v), _ptr(b["data"]), _ptr(b["hw"]), _ptr(b["off"]),
                  _ptr(b["cols"]), int(b["n_angles"]), int(b["n_ch"]), int(b["hmax"]), int(b["wmax"]), flags, _stream())
    return norm, direction
