"""Import the UNMODIFIED reference  --  TEST / BENCH INFRASTRUCTURE ONLY.

From /root/reference in the authoring container; on the GPU box (no /root/reference) from the byte-for-byte staged
archive ``oracle/_ref/topo_descriptors_ref.zip`` made by ``oracle/build_ref.py`` (git-ignored, shipped by gpurun).  Four stub packages
under ``oracle/_stubs`` stand in for the reference's missing imports (xarray, dask, utm,
yaconfigobject).  Used by ``oracle/make_golden.py`` to generate the golden
vectors in ``tests/golden`` and by ``tests/test_oracle.py`` (skipped when the reference is absent).
"""

import os
import sys
import warnings

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_STUBS = os.path.join(_HERE, "_stubs")


def _root():
    """sys.path entry that provides ``topo_descriptors``: the reference tree, or the staged archive (zipimport)."""
    if os.path.isdir("/root/reference/topo_descriptors"):
        return "/root/reference"
    staged = os.path.join(_HERE, "_ref", "topo_descriptors_ref.zip")
    return staged if os.path.isfile(staged) else None


REFERENCE_ROOT = _root()


def available():
    return REFERENCE_ROOT is not None


def load():
    """Return (topo, helpers) modules of the reference."""
    if not available():
        raise RuntimeError("reference tree not present")
    for p in (REFERENCE_ROOT, _STUBS):
        if p not in sys.path:
            sys.path.insert(0, p)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import topo_descriptors.helpers as hlp
        import topo_descriptors.topo as topo
    return topo, hlp


def fake_dataset(z, x, y, crs="epsg:2056", name="alti"):
    """Duck-typed Dataset subclassing the stub ``xarray.Dataset`` so the reference's isinstance
    checks pass (topo.py:825, helpers.py:179)."""
    load()
    import xarray as xr  # the stub

    class _DA:
        def __init__(self, values, dims):
            self.values = values
            self.dims = dims
            self.data = values

    class _DS(xr.Dataset):
        def __init__(self):
            self.attrs = {"crs": crs}
            self._v = {name: _DA(z, ("y", "x")), "x": _DA(np.asarray(x), ("x",)), "y": _DA(np.asarray(y), ("y",))}
            self.coords = {"x": self._v["x"], "y": self._v["y"]}

        def __iter__(self):
            return iter([name])

        def __getitem__(self, key):
            return self._v[key]

    return _DS()
