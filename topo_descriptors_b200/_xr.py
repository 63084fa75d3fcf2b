"""Dataset / DataArray duck-typing for the drop-in boundary.

The reference takes ``xarray.Dataset`` DEMs (topo.py:825, helpers.py:171-188).  xarray is an
optional dependency here: when it is importable the real classes are accepted; in any case the
two small containers below implement exactly the surface the hot path touches
(SURVEY.md section 8b): ``list(ds)[0]``, ``ds[name].dims/.values/.data``, ``ds["x"].values``,
``ds["y"].values``, ``ds.attrs``, ``ds.coords``.
"""

import numpy as np

try:  # pragma: no cover - xarray is absent from the build image
    import xarray as _xarray
except Exception:  # noqa: BLE001
    _xarray = None


class DataArray:
    """Minimal labelled array: ``values``/``data``, ``dims``, ``attrs``."""

    def __init__(self, values, dims=(), attrs=None):
        # a device-resident DEM (device.DeviceDEM) is kept as it is: the pre-stage hands such Datasets to compute_*
        self.values = values if getattr(values, "is_device_dem", False) else np.asarray(values)
        self.dims = tuple(dims)
        self.attrs = dict(attrs or {})

    @property
    def data(self):
        return self.values

    @property
    def shape(self):
        return self.values.shape

    @property
    def dtype(self):
        return self.values.dtype

    def copy(self, data=None):
        return DataArray(self.values.copy() if data is None else data, self.dims, self.attrs)

    def mean(self):
        return self.values.mean()

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.values, dtype=dtype)

    def __sub__(self, other):
        return DataArray(self.values - np.asarray(other), self.dims, self.attrs)

    def __repr__(self):
        return f"DataArray(dims={self.dims}, shape={self.values.shape}, dtype={self.values.dtype})"


class Dataset:
    """Minimal Dataset: ordered data variables + 1-D coordinates + attrs.

    ``Dataset({"alti": (("y", "x"), array)}, coords={"x": x, "y": y}, attrs={"crs": "epsg:2056"})``
    """

    def __init__(self, data_vars=None, coords=None, attrs=None):
        self._vars = {}
        for name, spec in (data_vars or {}).items():
            if isinstance(spec, DataArray):
                self._vars[name] = spec
            else:
                dims, values = spec[0], spec[1]
                self._vars[name] = DataArray(values, dims)
        self.coords = {}
        for name, values in (coords or {}).items():
            if isinstance(values, DataArray):
                self.coords[name] = values
            else:
                self.coords[name] = DataArray(np.asarray(values), (name,))
        self.attrs = dict(attrs or {})

    def __iter__(self):
        return iter(self._vars)

    def __len__(self):
        return len(self._vars)

    def __contains__(self, name):
        return name in self._vars or name in self.coords

    def __getitem__(self, name):
        if name in self._vars:
            return self._vars[name]
        return self.coords[name]

    def __setitem__(self, name, spec):
        if isinstance(spec, DataArray):
            self._vars[name] = spec
        else:
            self._vars[name] = DataArray(spec[1], spec[0])

    def sel_window(self, indexers=None):
        """The index ranges ``{dim: (start, stop)}`` that ``sel`` keeps, or None when a selection is not one
        contiguous run (non-monotonic coordinate).  Lets the drivers crop on the device before the D2H copy."""
        out = {}
        for name, sl in (indexers or {}).items():
            c = self.coords[name].values
            if not isinstance(sl, slice):
                raise TypeError("only slice indexers are supported")
            asc = c.size < 2 or c[-1] >= c[0]
            keep = np.ones(c.shape, dtype=bool)
            if sl.start is not None:
                keep &= (c >= sl.start) if asc else (c <= sl.start)
            if sl.stop is not None:
                keep &= (c <= sl.stop) if asc else (c >= sl.stop)
            idx = np.nonzero(keep)[0]
            if idx.size == 0:
                out[name] = (0, 0)
            elif idx[-1] - idx[0] + 1 != idx.size:
                return None
            else:
                out[name] = (int(idx[0]), int(idx[-1]) + 1)
        return out

    def sel(self, indexers=None):
        """Label-based crop with ``{coord: slice(lo, hi)}`` like ``xr.Dataset.sel`` (helpers.py:57-59).

        Slices are inclusive on both ends and follow the coordinate's own order
        (a descending ``y`` needs ``slice(hi, lo)``, exactly as in xarray).
        """
        if not indexers:
            return self
        index = {}
        for name, sl in indexers.items():
            c = self.coords[name].values
            if not isinstance(sl, slice):
                raise TypeError("only slice indexers are supported")
            lo, hi = sl.start, sl.stop
            asc = c.size < 2 or c[-1] >= c[0]
            keep = np.ones(c.shape, dtype=bool)
            if lo is not None:
                keep &= (c >= lo) if asc else (c <= lo)
            if hi is not None:
                keep &= (c <= hi) if asc else (c >= hi)
            index[name] = np.nonzero(keep)[0]
        out_vars = {}
        for name, da in self._vars.items():
            v = da.values
            for axis, dim in enumerate(da.dims):
                if dim in index:
                    v = np.take(v, index[dim], axis=axis)
            out_vars[name] = DataArray(v, da.dims, da.attrs)
        out_coords = {}
        for name, da in self.coords.items():
            v = da.values
            if name in index:
                v = v[index[name]]
            out_coords[name] = DataArray(v, da.dims, da.attrs)
        return Dataset(out_vars, out_coords, self.attrs)

    def __repr__(self):
        return f"Dataset(vars={list(self._vars)}, coords={list(self.coords)}, attrs={self.attrs})"


def is_dataset(obj):
    """``isinstance(obj, xr.Dataset)`` for real xarray objects and for the container above."""
    if isinstance(obj, Dataset):
        return True
    return _xarray is not None and isinstance(obj, _xarray.Dataset)


def is_dataarray(obj):
    if isinstance(obj, DataArray):
        return True
    return _xarray is not None and isinstance(obj, _xarray.DataArray)


def have_xarray():
    return _xarray is not None


def xarray_module():
    return _xarray
