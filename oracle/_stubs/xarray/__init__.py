"""Minimal stand-in for xarray: TEST INFRASTRUCTURE ONLY.

xarray is not installed in this image.  These two classes are just enough for the
reference package to import and for its ``isinstance(..., xr.Dataset)`` checks
(/root/reference/topo_descriptors/topo.py:825, helpers.py:179) to work with the
duck-typed Dataset in ``oracle/fake_xr.py``.  Never on the product import path.
"""


class DataArray:  # pragma: no cover - marker type
    pass


class Dataset:  # pragma: no cover - marker type
    pass


def open_dataset(*args, **kwargs):  # pragma: no cover
    raise NotImplementedError("xarray stub: no I/O")
