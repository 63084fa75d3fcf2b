"""Device pre-stage of the reference script (scripts/compute_topo_descriptors.py:17-19):

    dem_ds = hlp.get_dem_netcdf(path)               # ... .where(dem_ds > CFG.min_elevation)   helpers.py:31
    ind_nans, dem_ds = hlp.fill_na(dem_ds)          # np.where(isnan) + interpolate_na along x  helpers.py:137-154
    tp.compute_*(dem_ds, ..., ind_nans=ind_nans)

``fill_na_resident`` does the mask, the NaN census and the nearest-neighbour fill in one pass over the DEM in
HBM and returns a Dataset whose DEM variable is a ``DeviceDEM`` plus device-side index tensors: every
``compute_*`` driver and every descriptor accepts both, so the DEM is uploaded once and only results travel
back.  ``helpers.fill_na`` stays the host-side mirror for callers that want numpy back.
"""

import numpy as np

from . import _xr, device as dev, helpers as hlp
from .device import DeviceDEM


def fill_na_resident(dem_ds, mask_below=None):
    """Device version of ``hlp.fill_na`` (helpers.py:137-154), optionally preceded by the
    ``dem > mask_below`` mask of ``get_dem_netcdf`` (helpers.py:31).

    Returns ``(ind_nans, dem_ds_filled)``: ``ind_nans`` = (rows, cols) int32 CUDA tensors in
    ``np.where`` order (empty tuple-of-arrays semantics: two empty tensors when nothing is missing);
    ``dem_ds_filled`` has the same coordinates / attributes and a ``DeviceDEM`` as its DEM variable.
    """
    import torch

    hlp.check_dem(dem_ds)
    da = hlp.get_da(dem_ds)
    values = da.values
    tensor = values.tensor if isinstance(values, DeviceDEM) else dev.to_device(values)
    x = np.asarray(dem_ds["x"].values, dtype=np.float64)
    dx = np.diff(x)
    uniform = x.size < 2 or (dx[0] > 0 and np.all(dx == dx[0]))
    x_dev = None if uniform else torch.from_numpy(np.ascontiguousarray(x)).to(tensor.device)
    filled, idx, _ = dev.fill_na(tensor, x_dev, mask_below)
    if idx is None:
        empty = torch.empty((0,), dtype=torch.int32, device=tensor.device)
        idx = (empty, empty.clone())
    name = list(dem_ds)[0]
    out = _xr.Dataset({name: (da.dims, DeviceDEM(filled))}, coords=dem_ds.coords, attrs=dem_ds.attrs)
    return idx, out
