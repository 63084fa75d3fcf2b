"""utm stub (test infrastructure only): helpers.py:9,96 imports it; projected CRS never calls it."""


def from_latlon(*args, **kwargs):  # pragma: no cover
    raise NotImplementedError("utm stub: use a projected CRS (e.g. epsg:2056)")
