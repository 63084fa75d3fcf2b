"""dask.array stub: only the names topo.py:6,177-178 touches."""


class Array:  # pragma: no cover - marker type
    pass


def map_overlap(*args, **kwargs):  # pragma: no cover
    raise NotImplementedError("dask stub")
