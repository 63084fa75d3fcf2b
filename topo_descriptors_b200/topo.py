"""Drop-in for ``topo_descriptors.topo`` (reference: topo_descriptors/topo.py) with every array
kernel running on a B200 through libtopo_b200.so.

Same function names, argument meaning, defaults, return types and exceptions as the reference
(SURVEY.md section 8b).  Inputs are host arrays (numpy, an xarray DataArray, or the built-in
``_xr.DataArray``); each call moves the DEM to HBM, launches the CUDA kernels and brings the result
back, so callers see reference types.  To keep a DEM resident across calls pass a
:class:`~topo_descriptors_b200.device.DeviceDEM` instead of an array: results then stay on the
device as torch tensors (this is what the ``compute_*`` drivers do across scales).

There is no CPU implementation in this package: without the built library or a CUDA device every
function raises.
"""

import logging

import numpy as np

from . import CFG
from . import _geometry as geo
from . import _xr
from . import device as dev
from . import helpers as hlp
from .device import DeviceDEM

logger = logging.getLogger(__name__)


# ---------------------------------------------------------------------------------------------
# array <-> device marshalling
# ---------------------------------------------------------------------------------------------
class _Marshal:
    """Remembers how the caller passed the DEM so the result can go back in the same form."""

    def __init__(self, dem):
        self.on_device = isinstance(dem, DeviceDEM)
        self.container = None
        if self.on_device:
            self.ddem = dem
            self.dtype = np.dtype(np.float32)
            return
        if _xr.is_dataarray(dem):
            self.container = dem
            values = dem.values
        else:
            values = np.asarray(dem)
        if values.ndim != 2:
            raise ValueError("dem must be a 2-D array")
        self.dtype = values.dtype if values.dtype.kind == "f" else np.dtype(np.float64)
        self.ddem = DeviceDEM(dev.to_device(values))

    def back(self, tensor, dtype=None):
        """Device tensor -> what the reference would have returned for this input type."""
        if self.on_device:
            return tensor
        arr = tensor.cpu().numpy()
        dtype = self.dtype if dtype is None else dtype
        if arr.dtype != dtype:
            arr = arr.astype(dtype)
        if self.container is not None:
            return self.container.copy(data=arr)
        return arr


def _out_of_core(dem):
    """True for inputs that should be streamed through HBM in row bands instead of being uploaded whole: a
    ``numpy.memmap``, a dask array (the reference's ``map_overlap`` case, topo.py:177-178) or a DataArray backed by
    one.  Such inputs go through :mod:`tiler` (same kernels, global-coordinate views, bit-identical results)."""
    data = dem.data if _xr.is_dataarray(dem) and not isinstance(dem, _xr.DataArray) else dem
    return isinstance(data, np.memmap) or (hasattr(data, "chunks") and hasattr(data, "compute"))


def _lazy_source(dem):
    return dem.data if _xr.is_dataarray(dem) and not isinstance(dem, _xr.DataArray) else dem


def _smoothed(ddem, sigma):
    """Optional Gaussian pre-smoothing (topo.py:172-173, 297-298, 426-427).  The smoothed surface is
    a new DEM: its statistics (range, integrality) are recomputed."""
    if not sigma:
        return ddem
    if not ddem.is_whole:
        raise ValueError("pre-smoothing of a row band goes through bands.py (needs a wider halo)")
    return DeviceDEM(dev.gauss(ddem, sigma, sigma))


# ---------------------------------------------------------------------------------------------
# smoothed DEM
# ---------------------------------------------------------------------------------------------
def compute_dem(dem_ds, scales, ind_nans=[], crop=None, outdir="."):
    """Smoothed DEM for every scale, one file each (topo.py:16-59)."""
    hlp.check_dem(dem_ds)
    logger.info(f"***Starting dem computation for scales {scales} meters***")
    if not hasattr(scales, "__iter__"):
        scales = [scales]

    scales_pxl, _ = hlp.scale_to_pixel(scales, dem_ds)
    sigmas = scales_pxl / CFG.scale_std
    ddem, nans = _resident(dem_ds, ind_nans)

    for idx, sigma in enumerate(sigmas):
        logger.info(f"Computing scale {scales[idx]} meters")
        out = dev.gauss(ddem, sigma, sigma)
        _finish_output(out, nans, dem_ds, _dem_name(scales[idx]), crop, outdir, "m")


def dem(dem, sigma):
    """Gaussian-smoothed DEM, ``scipy.ndimage.gaussian_filter(dem, sigma)`` semantics (topo.py:62-80)."""
    if _out_of_core(dem):
        from . import tiler

        return tiler.gauss(_lazy_source(dem), sigma)
    m = _Marshal(dem)
    sig = (sigma, sigma) if np.isscalar(sigma) else tuple(sigma)
    return m.back(dev.gauss(m.ddem, sig[0], sig[1]))


def _dem_name(scale):
    return f"DEM_{scale}M"


# ---------------------------------------------------------------------------------------------
# TPI
# ---------------------------------------------------------------------------------------------
def compute_tpi(dem_ds, scales, smth_factors=None, ind_nans=[], crop=None, outdir="."):
    """TPI for every scale, one file each (topo.py:88-141)."""
    hlp.check_dem(dem_ds)
    logger.info(f"***Starting TPI computation for scales {scales} meters***")
    if not hasattr(scales, "__iter__"):
        scales = [scales]
    if not hasattr(smth_factors, "__iter__"):
        smth_factors = [smth_factors] * len(scales)

    scales_pxl, _ = hlp.scale_to_pixel(scales, dem_ds)
    sigmas = hlp.get_sigmas(smth_factors, scales_pxl)
    ddem, nans = _resident(dem_ds, ind_nans)
    _share_planes(ddem, scales_pxl, sigmas)

    for idx, scale_pxl in enumerate(scales_pxl):
        logger.info(f"Computing scale {scales[idx]} meters with smoothing factor {smth_factors[idx]} ...")
        out = tpi(ddem, scale_pxl, sigma=sigmas[idx])
        _finish_output(out, nans, dem_ds, _tpi_name(scales[idx], smth_factors[idx]), crop, outdir, "m")


@hlp.timer
def tpi(dem, size, sigma=None):
    """Topographic position index: elevation minus the mean of the neighbours inside a disc of
    diameter ``size`` pixels, centre excluded (topo.py:144-181).

    Zero-padded borders normalised by the full neighbour count, like the reference's
    ``signal.convolve(mode="same")``; any non-finite input value makes the whole output NaN (FFT
    semantics).  Result dtype follows the input (float32 in -> float32 out).

    Deviations from the reference, by design: (i) a float64 DEM is rounded to float32 on upload and the result is
    cast back to float64 (the reference would carry float64 through its FFT; the difference is the float32
    representation error of the elevations, <= 2.5e-4 m below 4096 m); (ii) the reference's ``convolve`` may pick
    its direct method for very small kernels, in which case a NaN only poisons its own footprint there -- here every
    size follows the FFT rule (all NaN), which is what the compute_* drivers assume when they re-stamp ``ind_nans``
    on a filled DEM.
    """
    if _out_of_core(dem):
        from . import tiler

        src = _lazy_source(dem)
        return tiler.tpi(tiler.gauss(src, sigma) if sigma else src, int(size))
    m = _Marshal(dem)
    return m.back(dev.tpi(_smoothed(m.ddem, sigma), int(size)))


def _tpi_name(scale, smth_factor):
    add = f"_SMTHFACT{smth_factor:.3g}" if smth_factor else ""
    return f"TPI_{scale}M{add}"


def circular_kernel(size):
    """Disc mask of diameter ``size`` (a square for size < 5), float32 (topo.py:191-213).  Host-side
    convenience only: the device derives the per-row half-widths itself."""
    middle = int(size / 2)
    if size < 5:
        return np.ones((size, size), dtype=np.float32)
    ii, jj = np.ogrid[:size, :size]
    return ((ii - middle) ** 2 + (jj - middle) ** 2 <= middle**2).astype(np.float32)


# ---------------------------------------------------------------------------------------------
# STD
# ---------------------------------------------------------------------------------------------
def compute_std(dem_ds, scales, smth_factors=None, ind_nans=[], crop=None, outdir="."):
    """Local standard deviation for every scale, one file each (topo.py:216-269)."""
    hlp.check_dem(dem_ds)
    logger.info(f"***Starting STD computation for scales {scales} meters***")
    if not hasattr(scales, "__iter__"):
        scales = [scales]
    if not hasattr(smth_factors, "__iter__"):
        smth_factors = [smth_factors] * len(scales)

    scales_pxl, _ = hlp.scale_to_pixel(scales, dem_ds)
    sigmas = hlp.get_sigmas(smth_factors, scales_pxl)
    ddem, nans = _resident(dem_ds, ind_nans)
    _share_planes(ddem, scales_pxl, sigmas)

    for idx, scale_pxl in enumerate(scales_pxl):
        logger.info(f"Computing scale {scales[idx]} meters with smoothing factor {smth_factors[idx]} ...")
        out = std(ddem, scale_pxl, sigma=sigmas[idx])
        _finish_output(out, nans, dem_ds, _std_name(scales[idx], smth_factors[idx]), crop, outdir, "m",
                       dtype=np.float64)


@hlp.timer
def std(dem, size, sigma=None):
    """Standard deviation inside a rolling disc of diameter ``size`` pixels (topo.py:272-307):
    sqrt(clip((sum trunc(x)^2 - (sum x)^2/N) / (N-1), 0)), including the reference's
    ``dem.astype("int32")`` truncation inside the squares.  Returns float64 like the reference (the
    kernel computes the variance in exact integer / float64 arithmetic and stores float32).
    float64 DEMs are rounded to float32 on upload (see ``tpi``); non-finite input gives an all-NaN result.
    """
    if _out_of_core(dem):
        from . import tiler

        src = _lazy_source(dem)
        return tiler.std(tiler.gauss(src, sigma) if sigma else src, int(size))
    m = _Marshal(dem)
    return m.back(dev.std(_smoothed(m.ddem, sigma), int(size)), dtype=np.float64)


def _std_name(scale, smth_factor):
    add = f"_SMTHFACT{smth_factor:.3g}" if smth_factor else ""
    return f"STD_{scale}M{add}"


# ---------------------------------------------------------------------------------------------
# valley / ridge
# ---------------------------------------------------------------------------------------------
def compute_valley_ridge(dem_ds, scales, mode, flat_list=[0, 0.15, 0.3], smth_factors=None, ind_nans=[], crop=None,
                         outdir="."):
    """Valley or ridge index (norm + direction) for every scale (topo.py:317-386)."""
    hlp.check_dem(dem_ds)
    logger.info(f"***Starting {mode} index computation for scales {scales} meters***")
    if not hasattr(scales, "__iter__"):
        scales = [scales]
    if not hasattr(smth_factors, "__iter__"):
        smth_factors = [smth_factors] * len(scales)

    scales_pxl, _ = hlp.scale_to_pixel(scales, dem_ds)
    sigmas = hlp.get_sigmas(smth_factors, scales_pxl)
    ddem, nans = _resident(dem_ds, ind_nans)

    for idx, scale_pxl in enumerate(scales_pxl):
        logger.info(f"Computing scale {scales[idx]} meters with smoothing factor {smth_factors[idx]} ...")
        names = _valley_ridge_names(scales[idx], mode, smth_factors[idx])
        outs = valley_ridge(ddem, scale_pxl, mode, flat_list, sigmas[idx])
        for out, name in zip(outs, names):
            _finish_output(out, nans, dem_ds, name, crop, outdir, "1")


_BANK_CACHE = {}


def _device_bank_rotated_on_device(size, mode, flat_list, device):
    """The bank of the FFT route for large kernels, rotated on the GPU: the host only runs scipy's
    spline prefilter on the F source kernels and the 180 rotation set-ups."""
    import ctypes

    import torch

    from . import _lib

    base = geo.valley_kernels(int(size), list(flat_list))
    if mode == "ridge":
        base = base * np.float32(-1)
    F, H, W = base.shape
    angles = np.arange(0, 180, dtype=np.float32)
    recs, total = geo.plan_rotations((H, W), F, angles)
    coef = torch.from_numpy(geo.spline_coefficients(base)).to(device)
    recs_dev = torch.from_numpy(recs.view(np.uint8).reshape(-1).copy()).to(device)
    scratch = torch.empty(total, dtype=torch.float32, device=device)
    out = torch.empty(total, dtype=torch.float32, device=device)
    _lib.call("topo_rotate_bank_f32", dev._ptr(coef), F, H, W, dev._ptr(recs_dev), len(recs), ctypes.c_float(-9999.0),
              dev._ptr(scratch), dev._ptr(out), dev._stream())
    off, hw = [], []
    for a, r in enumerate(recs):
        n = int(r["oh"]) * int(r["ow"])
        for m in range(F):
            off.append(int(r["out_off"]) + m * n)
            hw.append((int(r["oh"]), int(r["ow"]), a))
    return {"plain": {"data": out, "off": np.array(off, dtype=np.int64), "hw": np.array(hw, dtype=np.int32)},
            "n_angles": len(recs), "n_ch": F, "hmax": int(recs["oh"].max()), "wmax": int(recs["ow"].max())}


def _device_bank(size, mode, flat_list, device):
    key = (int(size), mode, tuple(float(f) for f in flat_list), str(device))
    bank = _BANK_CACHE.get(key)
    if bank is None and int(size) * 1.42 >= dev.VALLEY_FFT_MIN_EXTENT and dev.VALLEY_DEVICE_ROTATION:
        bank = _device_bank_rotated_on_device(size, mode, flat_list, device)
        if len(_BANK_CACHE) > 16:
            _BANK_CACHE.clear()
        _BANK_CACHE[key] = bank
    if bank is None:
        import torch

        host = geo.build_valley_bank(int(size), mode, flat_list)

        def upload(h):
            b = dict(h)
            for k in ("data", "hw", "off", "cols"):
                b[k] = torch.from_numpy(h[k]).to(device)
            return b

        bank = dict(host, groups=[upload(g) for g in host["groups"]]) if "groups" in host else upload(host)
        plain = host["plain"]  # FFT route: kernels on the device, offsets / extents stay on the host
        bank["plain"] = {"data": torch.from_numpy(plain["data"]).to(device), "off": plain["off"], "hw": plain["hw"]}
        if len(_BANK_CACHE) > 16:
            _BANK_CACHE.clear()
        _BANK_CACHE[key] = bank
    return bank


@hlp.timer
def valley_ridge(dem, size, mode, flat_list=[0, 0.15, 0.3], sigma=None):
    """Valley or ridge index (topo.py:389-453): the z-scored DEM is matched against V / U shaped
    kernels of width ``size`` rotated through 0..179 degrees; returns ``[norm, direction]`` (float32)
    where direction is the best angle (0 = W-E, 90 = S-N, clockwise).

    The whole 180-angle bank runs in one kernel launch.  Reproduces the reference's 3-D convolution
    (neighbouring flat-list kernels are summed per channel) and strict '>' argmax.
    """
    if mode not in ("valley", "ridge"):
        raise ValueError(f"Unknown mode {mode!r}")
    m = _Marshal(dem)
    ddem = _smoothed(m.ddem, sigma)
    st = ddem.stats
    if st["nonfinite"] > 0:
        # the FFT spreads NaN everywhere: no comparison ever succeeds, norm = clip(-inf) = 0
        # (direction is uninitialised memory in the reference; 0 here)
        zero = dev._new(ddem.rows, ddem.nx, ddem.tensor)
        dev.fill(zero, 0.0)
        return [m.back(zero, np.float32), m.back(zero.clone(), np.float32)]
    mean64 = st["sum"] / st["n"]
    var64 = max(st["sumsq"] / st["n"] - mean64 * mean64, 0.0)
    normed = dev.zscore(ddem, np.float32(mean64), np.float32(np.sqrt(var64)))
    bank = _device_bank(size, mode, flat_list, ddem.tensor.device)
    norm, direction = dev.valley_ridge(normed, bank)
    return [m.back(norm, np.float32), m.back(direction, np.float32)]


def _valley_ridge_names(scale, mode, smth_factor):
    add = f"_SMTHFACT{smth_factor:.3g}" if smth_factor else ""
    return [f"{mode}_NORM_{scale}M{add}", f"{mode}_DIR_{scale}M{add}"]


def _valley_kernels(size, flat_list):
    """Z-scored V / U kernels (topo.py:466-499)."""
    return geo.valley_kernels(size, flat_list)


def _ridge_kernels(size, flat_list):
    """Flipped-sign valley kernels (topo.py:502-518)."""
    return geo.valley_kernels(size, flat_list) * -1


def _rotate_kernels(kernel, angle):
    """Rotated, re-normalised kernel stack (topo.py:521-531)."""
    return geo.rotate_kernels(kernel, angle)


# ---------------------------------------------------------------------------------------------
# gradient / slope / aspect
# ---------------------------------------------------------------------------------------------
def compute_gradient(dem_ds, scales, sig_ratios=1, ind_nans=[], crop=None, outdir="."):
    """W-E / S-N derivatives, slope and aspect for every scale (topo.py:534-594)."""
    hlp.check_dem(dem_ds)
    logger.info(f"***Starting gradients computation for scales {scales} meters***")
    if not hasattr(scales, "__iter__"):
        scales = [scales]
    if not hasattr(sig_ratios, "__iter__"):
        sig_ratios = [sig_ratios] * len(scales)

    scales_pxl, res_meters = hlp.scale_to_pixel(scales, dem_ds)
    sigmas = scales_pxl / CFG.scale_std
    ddem, nans = _resident(dem_ds, ind_nans)
    all_units = ["1", "1", "degree", "degree"]

    for idx, sigma in enumerate(sigmas):
        logger.info(f"Computing scale {scales[idx]} meters with sigma ratio {sig_ratios[idx]} ...")
        names = _gradient_names(scales[idx], sig_ratios[idx])
        outs = gradient(ddem, sigma, res_meters, sig_ratio=sig_ratios[idx])
        for out, name, units in zip(outs, names, all_units):
            _finish_output(out, nans, dem_ds, name, crop, outdir, units)


@hlp.timer
def gradient(dem, sigma, res_meters, sig_ratio=1):
    """Directional derivatives, slope and aspect (topo.py:597-644): ``[dx, dy, slope, aspect]``.

    sigma <= 1: Sobel.  Otherwise central differences of the Gaussian-smoothed DEM (anisotropic
    smoothing when ``sig_ratio != 1``), divided by the signed grid resolution ``res_meters`` (the
    second return value of ``helpers.scale_to_pixel``); slope in degrees, aspect clockwise from
    north.  float32, evaluated in the reference's own operation order.
    """
    if _out_of_core(dem) and sig_ratio == 1:
        from . import tiler

        return tiler.gradient(_lazy_source(dem), sigma, res_meters)
    m = _Marshal(dem)
    d = m.ddem
    device = d.tensor.device
    rx, ry = dev._Res(res_meters["x"], device), dev._Res(res_meters["y"], device)
    rx2d, ry2d = rx.is_2d, ry.is_2d
    if sigma <= 1:
        outs = dev.sobel_gradient(d, rx, rx2d, ry, ry2d, normalize=True)
    else:
        if not d.is_whole:
            raise ValueError("gradient of a row band goes through bands.py (needs a halo)")
        if sig_ratio == 1:
            outs = dev.gradient(d, sigma, rx, rx2d, ry, ry2d)
        else:
            sigma_perp = sigma * sig_ratio
            gx = DeviceDEM(dev.gauss(d, sigma_perp, sigma))
            gy = DeviceDEM(dev.gauss(d, sigma, sigma_perp))
            outs = dev.gradient_from_smooth(gx, gy, rx, rx2d, ry, ry2d)
    return [m.back(o, np.float32) for o in outs]


def _gradient_names(scale, sig_ratio):
    tag = f"{scale}M_SIGRATIO{sig_ratio:.3g}"
    return [f"WE_DERIVATIVE_{tag}", f"SN_DERIVATIVE_{tag}", f"SLOPE_{tag}", f"ASPECT_{tag}"]


def sobel(dem):
    """Sobel derivatives ``(dx, dy)`` in metres per pixel, reflect borders (topo.py:658-685)."""
    m = _Marshal(dem)
    dx, dy = dev.sobel_gradient(m.ddem, normalize=False)
    return m.back(dx, np.float32), m.back(dy, np.float32)


# ---------------------------------------------------------------------------------------------
# Sx
# ---------------------------------------------------------------------------------------------
def compute_sx(dem_ds, azimuth, radius, height=10.0, azimuth_arc=10.0, azimuth_steps=15, radius_min=0.0, crop=None,
               outdir="."):
    """Sx for one azimuth, saved to file (topo.py:715-772)."""
    hlp.check_dem(dem_ds)
    logger.info(f"***Starting Sx computation for azimuth {azimuth} meters and radius {radius}***")
    array = sx(dem_ds, azimuth, radius, height=height, azimuth_arc=azimuth_arc, azimuth_steps=azimuth_steps,
               radius_min=radius_min)
    hlp.to_netcdf(array, dem_ds, _sx_name(radius, azimuth), crop, outdir, "degree")


def _sx_plan(dem_ds, azimuths_centre, radius, azimuth_arc, azimuth_steps, radius_min):
    """Host geometry of topo.py:828-853 for one or several sector centres -> CSR sample lists."""
    if azimuth_arc == 0:
        azimuth_steps = 1
    _, res_meters = hlp.scale_to_pixel(radius, dem_ds)
    dx = res_meters["x"].mean()
    dy = res_meters["y"].mean()
    window_distance = geo.sx_distance(radius, dx, dy)
    window_distance[window_distance < radius_min] = np.nan
    centre = np.floor(np.array(window_distance.shape) / 2)
    offs, invs, begin = [], [], [0]
    window = int(window_distance.shape[0] / 2)
    for az in azimuths_centre:
        azimuths = np.linspace(az - azimuth_arc / 2, az + azimuth_arc / 2, azimuth_steps)
        source = (centre + geo.sx_source_idx_delta(azimuths, radius, dx, dy)).astype(int)
        lines = geo.sx_bresenhamlines(source, centre)
        o, inv, window = geo.sx_samples(window_distance, lines)
        offs.append(o)
        invs.append(inv)
        begin.append(begin[-1] + len(inv))
    offsets = np.concatenate(offs) if offs else np.zeros((0, 2), np.int32)
    inv = np.concatenate(invs) if invs else np.zeros(0, np.float32)
    return offsets.reshape(-1, 2), inv, np.array(begin, dtype=np.int32), window


def _sx_device(ddem, plan, height, out_gy0=None, out_rows=None):
    import torch

    offsets, inv, begin, window = plan
    device = ddem.tensor.device
    off_d = torch.from_numpy(np.ascontiguousarray(offsets, dtype=np.int32)).to(device)
    inv_d = torch.from_numpy(np.ascontiguousarray(inv, dtype=np.float32)).to(device)
    beg_d = torch.from_numpy(begin).to(device)
    n_az = len(begin) - 1
    if len(offsets):
        extents = (offsets[:, 0].min(), offsets[:, 0].max(), offsets[:, 1].min(), offsets[:, 1].max())
    else:
        extents = (0, 0, 0, 0)
    return dev.sx(ddem, off_d, inv_d, beg_d, n_az, window, height, extents, out_gy0, out_rows)


@hlp.timer
def sx(dem_ds, azimuth, radius, height=10.0, azimuth_arc=10.0, azimuth_steps=15, radius_min=0.0):
    """Sx (Winstral): the maximum slope angle, in degrees, towards the terrain lying within
    ``radius`` metres in the sector ``azimuth`` +- ``azimuth_arc``/2 (topo.py:775-858).

    ``dem_ds`` must be a Dataset (``TypeError`` otherwise, like the reference).  A frame of
    radius-in-pixels cells along every edge is 0; NaN terrain samples are ignored; a NaN centre
    gives NaN.  Returns a float32 array.  ``azimuth`` may also be a sequence: all sectors then run
    in one launch and the result has a leading azimuth axis.
    """
    if not _xr.is_dataset(dem_ds):
        raise TypeError("Argument 'dem_ds' must be a xr.Dataset.")
    many = hasattr(azimuth, "__iter__")
    centres = list(azimuth) if many else [azimuth]
    plan = _sx_plan(dem_ds, centres, radius, azimuth_arc, azimuth_steps, radius_min)
    values = hlp.get_da(dem_ds).values
    ddem = values if isinstance(values, DeviceDEM) else DeviceDEM(dev.to_device(values))
    out = _sx_device(ddem, plan, height).cpu().numpy()
    return out if many else out[0]


def _sx_distance(radius, dx, dy):
    return geo.sx_distance(radius, dx, dy)


def _sx_source_idx_delta(azimuths, radius, dx, dy):
    return geo.sx_source_idx_delta(azimuths, radius, dx, dy)


def _sx_bresenhamlines(start, end):
    return geo.sx_bresenhamlines(start, end)


def _sx_name(radius, azimuth):
    return f"SX_RADIUS{int(radius)}_AZIMUTH{int(azimuth)}"


# ---------------------------------------------------------------------------------------------
# shared by the compute_* drivers: DEM residency and the output stage
# ---------------------------------------------------------------------------------------------
def _share_planes(ddem, scales_pxl, sigmas):
    """Several unsmoothed scales of one DEM: the disc prefix planes do not depend on the scale, build them once."""
    if len(scales_pxl) > 1 and all(sg is None for sg in sigmas):
        ddem.share_disc_planes(int(max(scales_pxl)))


def _resident(dem_ds, ind_nans):
    """Upload the DEM once for all scales; upload the NaN re-stamp indices once (topo.py:129,139)."""
    import torch

    values = hlp.get_da(dem_ds).values
    ddem = values if isinstance(values, DeviceDEM) else DeviceDEM(dev.to_device(values))
    nans = None
    if ind_nans is not None and len(ind_nans) == 2 and len(ind_nans[0]):
        if torch.is_tensor(ind_nans[0]):  # already on the device (prestage.fill_na_resident)
            nans = (ind_nans[0].to(torch.int32), ind_nans[1].to(torch.int32))
        else:
            rows = torch.from_numpy(np.ascontiguousarray(ind_nans[0], dtype=np.int32)).to(ddem.tensor.device)
            cols = torch.from_numpy(np.ascontiguousarray(ind_nans[1], dtype=np.int32)).to(ddem.tensor.device)
            nans = (rows, cols)
    return ddem, nans


def _finish_output(tensor, nans, dem_ds, name, crop, outdir, units, dtype=np.float32):
    """``array[ind_nans] = np.nan`` and the crop on the device, one D2H of what is kept, then the reference's
    writer (topo.py:139-140 + helpers.py:57-59)."""
    if nans is not None:
        dev.stamp(tensor, nans[0], nans[1])
    window = dem_ds.sel_window(crop) if crop and isinstance(dem_ds, _xr.Dataset) else None
    if window is not None:
        dims = hlp.get_da(dem_ds).dims
        (r0, r1), (c0, c1) = (window.get(d, (0, n)) for d, n in zip(dims, tensor.shape))
        tensor = tensor[r0:r1, c0:c1]
    array = tensor.cpu().numpy()
    if array.dtype != dtype:
        array = array.astype(dtype)
    hlp.to_netcdf(array, dem_ds, name, crop, outdir, units, window=window)
