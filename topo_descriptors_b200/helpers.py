"""Host-side helpers of the drop-in boundary (reference: topo_descriptors/helpers.py).

Same names, argument meaning and return types as the reference helpers the hot path uses
(SURVEY.md section 8 row a17).  All of this is tiny scalar / coordinate work and stays on the host.
The NetCDF reader/writer needs xarray (absent from the build image): ``to_netcdf`` falls back to a
``.npz`` sink with the same naming so the ``compute_*`` drivers remain usable.
"""

import datetime as dt
import functools
import logging
import time
from pathlib import Path

import numpy as np

from . import CFG
from . import _xr

logger = logging.getLogger(__name__)


def get_da(dem_ds):
    """First data variable of the Dataset, whatever its name (helpers.py:191-196)."""
    return dem_ds[list(dem_ds)[0]]


def check_dem(dem):
    """Validate the DEM data model (helpers.py:171-188): a Dataset, one 2-D ('y','x') field,
    a ``crs`` attribute carrying an ``epsg:`` code.  Same exception types as the reference."""
    if not _xr.is_dataset(dem):
        raise ValueError("dem must be a xr.Dataset")
    if tuple(get_da(dem).dims) != ("y", "x"):
        raise ValueError("dem dimensions must be ('y', 'x')")
    if "crs" not in dem.attrs:
        raise KeyError("missing 'crs' (case sensitive) attribute in dem")
    if "epsg:" not in dem.attrs["crs"].lower():
        raise ValueError("missing 'epsg:' (case insensitive) key in the 'crs' attribute")


def round_up_to_odd(f):
    """Nearest odd integer, numpy round-half-even on (f-1)/2 (helpers.py:108-111;
    pinned by test/test_helpers.py:6-11)."""
    half = np.round((np.asarray(f, dtype=np.float64) - 1.0) / 2.0)
    return np.asarray(half * 2.0 + 1.0, dtype=np.int64)


def _wgs84_to_utm(lat, lon):
    """Forward transverse-Mercator (WGS84 -> UTM easting/northing), zone of the grid centre.

    Stands in for ``utm.from_latlon`` (helpers.py:96; the package is not installed here).
    Krueger series to n^4: sub-millimetre inside a zone, far below what the pixel-resolution
    estimate needs.  Like ``utm`` the zone is chosen once (from the mean longitude / latitude)
    so that the whole grid shares one projection.
    """
    lat = np.asarray(lat, dtype=np.float64)
    lon = np.asarray(lon, dtype=np.float64)
    a = 6378137.0
    f = 1.0 / 298.257223563
    k0 = 0.9996
    n = f / (2.0 - f)
    A = a / (1.0 + n) * (1.0 + n**2 / 4.0 + n**4 / 64.0)
    alpha = (
        n / 2.0 - 2.0 * n**2 / 3.0 + 5.0 * n**3 / 16.0 + 41.0 * n**4 / 180.0,
        13.0 * n**2 / 48.0 - 3.0 * n**3 / 5.0 + 557.0 * n**4 / 1440.0,
        61.0 * n**3 / 240.0 - 103.0 * n**4 / 140.0,
        49561.0 * n**4 / 161280.0,
    )
    zone = int((float(np.mean(lon)) + 180.0) / 6.0) % 60 + 1
    lon0 = np.deg2rad((zone - 1) * 6.0 - 180.0 + 3.0)
    phi = np.deg2rad(lat)
    lam = np.deg2rad(lon) - lon0
    e = np.sqrt(f * (2.0 - f))
    t = np.sinh(np.arctanh(np.sin(phi)) - e * np.arctanh(e * np.sin(phi)))
    xi = np.arctan2(t, np.cos(lam))
    eta = np.arctanh(np.sin(lam) / np.sqrt(1.0 + t * t))
    E = eta.copy()
    N = xi.copy()
    for j, aj in enumerate(alpha, start=1):
        E = E + aj * np.cos(2 * j * xi) * np.sinh(2 * j * eta)
        N = N + aj * np.sin(2 * j * xi) * np.cosh(2 * j * eta)
    easting = 500000.0 + k0 * A * E
    northing = k0 * A * N
    if float(np.mean(lat)) < 0:
        northing = northing + 10000000.0
    return easting, northing


def scale_to_pixel(scales, dem_ds):
    """Metres -> closest odd number of pixels, plus the per-point grid resolution in metres
    (helpers.py:68-105).  Projected CRS: 1-D signed resolutions from ``np.gradient`` of the
    coordinates.  EPSG:4326: coordinates are projected to UTM on the full meshgrid first and the
    resolutions are 2-D float32-derived arrays, as in the reference.
    """
    check_dem(dem_ds)
    x_coords = np.asarray(dem_ds["x"].values)
    y_coords = np.asarray(dem_ds["y"].values)
    if "epsg:4326" in dem_ds.attrs["crs"].lower():
        logger.debug("Reprojecting coordinates from WGS84 to UTM to obtain units of meters")
        lon, lat = np.meshgrid(x_coords, y_coords)
        x_coords, y_coords = _wgs84_to_utm(lat, lon)
        x_coords, y_coords = x_coords.astype(np.float32), y_coords.astype(np.float32)

    x_res = np.gradient(x_coords, axis=x_coords.ndim - 1)
    y_res = np.gradient(y_coords, axis=0)
    mean_res = np.mean(np.abs([x_res.mean(), y_res.mean()]))
    logger.debug(f"Estimated resolution: {mean_res:.0f} meters.")
    return round_up_to_odd(np.array(scales) / mean_res), {"x": x_res, "y": y_res}


def get_sigmas(smth_factors, scales_pxl):
    """Smoothing factors -> Gaussian sigmas in pixels, ``None`` where the factor is falsy
    (helpers.py:114-134): sigma = factor * scale_pxl / CFG.scale_std."""
    factors = np.array([fact if fact else np.nan for fact in smth_factors], dtype=np.float64)
    sigmas = factors * np.asarray(scales_pxl) / CFG.scale_std
    return [None if np.isnan(sigma) else sigma for sigma in sigmas]


def timer(func):
    """Log the wall time of a descriptor call at INFO, same message as helpers.py:157-168."""

    @functools.wraps(func)
    def wrapper_timer(*args, **kwargs):
        t_start = time.monotonic()
        value = func(*args, **kwargs)
        elapsed = str(dt.timedelta(seconds=time.monotonic() - t_start)).split(".", 2)[0]
        logger.info(f"Computed in {elapsed} (HH:mm:ss)")
        return value

    return wrapper_timer


def fill_na(dem_ds):
    """Indices of the NaNs + the DEM with NaNs filled by nearest neighbour along x, extrapolating
    at the row ends (helpers.py:137-154: ``interpolate_na(dim="x", method="nearest",
    fill_value="extrapolate")``).  Ties (equidistant neighbours) take the left value, as scipy's
    ``interp1d(kind="nearest")`` does (it rounds half down).
    """
    da = get_da(dem_ds)
    values = np.asarray(da.values)
    nan_mask = np.isnan(values)
    ind_nans = np.where(nan_mask)
    if _xr.have_xarray() and not isinstance(dem_ds, _xr.Dataset):  # pragma: no cover
        return ind_nans, dem_ds.interpolate_na(dim="x", method="nearest", fill_value="extrapolate")

    x = np.asarray(dem_ds["x"].values, dtype=np.float64)
    dx = np.diff(x)
    if len(ind_nans[0]) == 0:
        filled = values.copy()
    elif x.size > 1 and (np.all(dx > 0) or np.all(dx < 0)):
        filled = _fill_nearest_monotonic(values, nan_mask, x)
    else:
        filled = _fill_nearest_rows(values, nan_mask, x, ind_nans)
    name = list(dem_ds)[0]
    out = _xr.Dataset({name: (da.dims, filled)}, coords=dem_ds.coords, attrs=dem_ds.attrs)
    return ind_nans, out


def _fill_nearest_monotonic(values, nan_mask, x):
    """Nearest valid cell along x for every missing cell, all rows at once (x strictly monotonic): running
    last-valid / next-valid indices, the nearer one in x wins, exact half-way points take the lower-x neighbour,
    rows without a valid cell stay NaN."""
    ny, nx = values.shape
    idx = np.arange(nx)
    left = np.maximum.accumulate(np.where(nan_mask, -1, idx[None, :]), axis=1)
    right = np.minimum.accumulate(np.where(nan_mask, nx, idx[None, :])[:, ::-1], axis=1)[:, ::-1]
    rows, cols = np.nonzero(nan_mask)
    li, ri = left[rows, cols], right[rows, cols]
    has_l, has_r = li >= 0, ri < nx
    xl = x[np.clip(li, 0, nx - 1)]
    xr = x[np.clip(ri, 0, nx - 1)]
    dl, dr = np.abs(x[cols] - xl), np.abs(xr - x[cols])
    take_left = has_l & (~has_r | (dl < dr) | ((dl == dr) & (xl < xr)))
    pick = np.where(take_left, li, ri)
    ok = has_l | has_r
    filled = values.copy()
    filled[rows[ok], cols[ok]] = values[rows[ok], np.clip(pick[ok], 0, nx - 1)]
    return filled


def _fill_nearest_rows(values, nan_mask, x, ind_nans):
    """Row-by-row variant for unsorted x coordinates (sorted per row like scipy's interp1d does)."""
    filled = values.copy()
    for row in np.unique(ind_nans[0]):
        good = ~nan_mask[row]
        if not good.any():
            continue
        xg = x[good]
        vg = values[row, good]
        order = np.argsort(xg)
        xg, vg = xg[order], vg[order]
        q = x[nan_mask[row]]
        # nearest with half-way points going to the lower-x neighbour
        mid = (xg[:-1] + xg[1:]) / 2.0
        idx = np.searchsorted(mid, q, side="left")
        filled[row, nan_mask[row]] = vg[idx]
    return filled


def get_dem_netcdf(path_dem):
    """Load a DEM NetCDF as float32 and mask values <= CFG.min_elevation (helpers.py:17-31).
    Needs xarray + a NetCDF backend; raises ImportError otherwise."""
    xr = _xr.xarray_module()
    if xr is None:
        raise ImportError("get_dem_netcdf needs xarray, which is not installed")
    dem_ds = xr.open_dataset(path_dem).astype(np.float32).squeeze(drop=True)  # pragma: no cover
    return dem_ds.where(dem_ds > CFG.min_elevation)  # pragma: no cover


def to_netcdf(array, dem_ds, name, crop=None, outdir=".", units=None, window=None):
    """Save one descriptor array with the DEM's coordinates and attributes (helpers.py:34-65).

    ``window`` = ``{dim: (start, stop)}`` (from ``Dataset.sel_window(crop)``) says that ``array`` was already
    cropped on the device: only the coordinates are cut here.

    File name ``topo_<NAME>.nc`` with NAME upper-cased.  With real xarray inputs the file is a
    NetCDF written by xarray; with the built-in Dataset container the same content goes to
    ``topo_<NAME>.npz`` (arrays: the variable, its coordinates; plus ``units`` and ``crs``).
    Returns the path written.
    """
    name = str.upper(name)
    outdir = Path(outdir)
    dims = get_da(dem_ds).dims
    xr = _xr.xarray_module()
    if xr is not None and isinstance(dem_ds, xr.Dataset):  # pragma: no cover
        ds = xr.Dataset({name: (dims, array)}, coords=dem_ds.coords, attrs=dem_ds.attrs).sel(crop)
        if units is not None:
            ds[name].attrs.update(units=units)
        path = outdir / f"topo_{name}.nc"
        ds.to_netcdf(path)
    else:
        if window is not None:
            coords = {}
            for cname, c in dem_ds.coords.items():
                lo, hi = window.get(cname, (0, c.values.shape[0]))
                coords[cname] = _xr.DataArray(c.values[lo:hi], c.dims, c.attrs)
            ds = _xr.Dataset({name: (dims, array)}, coords=coords, attrs=dem_ds.attrs)
        else:
            ds = _xr.Dataset({name: (dims, array)}, coords=dem_ds.coords, attrs=dem_ds.attrs).sel(crop)
        if units is not None:
            ds[name].attrs.update(units=units)
        path = outdir / f"topo_{name}.npz"
        payload = {name: ds[name].values}
        for cname, c in ds.coords.items():
            payload[cname] = c.values
        payload["units"] = np.array("" if units is None else units)
        payload["crs"] = np.array(ds.attrs.get("crs", ""))
        np.savez(path, **payload)
    logger.info(f"saved: {path}")
    return path
