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


def _utm_zone_number(latitude, longitude):
    """Zone rule of ``utm.latlon_to_zone_number``: array input -> the FIRST element decides (the whole grid then
    shares one projection), with the Norway (32V) and Svalbard (31X-37X) exceptions."""
    latitude = float(np.asarray(latitude).flat[0])
    longitude = float(np.asarray(longitude).flat[0])
    if longitude == 180:
        longitude = -180.0
    if 56 <= latitude < 64 and 3 <= longitude < 12:
        return 32
    if 72 <= latitude <= 84 and longitude >= 0:
        if longitude < 9:
            return 31
        if longitude < 21:
            return 33
        if longitude < 33:
            return 35
        if longitude < 42:
            return 37
    return int((longitude + 180) / 6) % 60 + 1


def _wgs84_to_utm(lat, lon):
    """WGS84 -> UTM easting / northing, a restatement of ``utm.from_latlon`` (helpers.py:96).

    The ``utm`` package (un-pinned in the reference's requirements.txt, absent from this image) publishes a
    truncated Snyder series with E = 0.00669438, R = 6378137, K0 = 0.9996; the same series is evaluated here in the
    same operation order, and it reproduces the package's documented known answer
    ``from_latlon(51.2, 7.5) == (395201.3103811303, 5673135.241182375, 32, 'U')`` to 1e-9 m (tests/test_host.py).
    Zone: ``_utm_zone_number`` (first element of array input, Norway / Svalbard exceptions); latitudes of mixed
    sign raise like the package.
    """
    lat = np.asarray(lat, dtype=np.float64)
    lon = np.asarray(lon, dtype=np.float64)
    if lat.size and lat.min() < 0 <= lat.max():
        raise ValueError("latitudes must all have the same sign")
    K0 = 0.9996
    E = 0.00669438
    E2 = E * E
    E3 = E2 * E
    E_P2 = E / (1 - E)
    M1 = 1 - E / 4 - 3 * E2 / 64 - 5 * E3 / 256
    M2 = 3 * E / 8 + 3 * E2 / 32 + 45 * E3 / 1024
    M3 = 15 * E2 / 256 + 45 * E3 / 1024
    M4 = 35 * E3 / 3072
    R = 6378137

    lat_rad = np.radians(lat)
    lat_sin = np.sin(lat_rad)
    lat_cos = np.cos(lat_rad)
    lat_tan = lat_sin / lat_cos
    lat_tan2 = lat_tan * lat_tan
    lat_tan4 = lat_tan2 * lat_tan2
    zone = _utm_zone_number(lat, lon)
    central_lon_rad = np.radians((zone - 1) * 6 - 180 + 3)
    n = R / np.sqrt(1 - E * lat_sin**2)
    c = E_P2 * lat_cos**2
    a = lat_cos * ((np.radians(lon) - central_lon_rad + np.pi) % (2 * np.pi) - np.pi)
    a2 = a * a
    a3 = a2 * a
    a4 = a3 * a
    a5 = a4 * a
    a6 = a5 * a
    m = R * (M1 * lat_rad - M2 * np.sin(2 * lat_rad) + M3 * np.sin(4 * lat_rad) - M4 * np.sin(6 * lat_rad))
    easting = K0 * n * (a + a3 / 6 * (1 - lat_tan2 + c) + a5 / 120 * (5 - 18 * lat_tan2 + lat_tan4 + 72 * c - 58 * E_P2)) + 500000
    northing = K0 * (m + n * lat_tan * (a2 / 2 + a4 / 24 * (5 - lat_tan2 + 9 * c + 4 * c**2)
                                        + a6 / 720 * (61 - 58 * lat_tan2 + lat_tan4 + 600 * c - 330 * E_P2)))
    if lat.size and lat.max() < 0:
        northing = northing + 10000000
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
