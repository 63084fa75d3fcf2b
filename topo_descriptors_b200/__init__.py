"""topo_descriptors_b200 - B200-native raster-filter hot path of topo-descriptors.

Drop-in for the array kernels and ``compute_*`` drivers of ``topo_descriptors.topo``
(reference: topo_descriptors/topo.py, topo_descriptors/helpers.py).  Every descriptor
runs as a hand-written sm_100a CUDA kernel behind the C-ABI in ``include/topo_b200.h``;
there is no CPU fallback: importing :mod:`topo_descriptors_b200.topo` works anywhere,
calling a descriptor without the built library or without a GPU raises.

``CFG`` mirrors the reference's two configuration constants
(topo_descriptors/config/topo_descriptors.conf:1-5, loaded at __init__.py:15).
"""

__version__ = "0.1.0"


class _Config:
    """The two constants of the reference configuration file.

    ``min_elevation`` (helpers.py:31) - values <= this are filtered when loading a DEM.
    ``scale_std``     (topo.py:49,573; helpers.py:131) - standard deviations per unit scale.
    Attributes can be overwritten in place (the override hook).
    """

    def __init__(self):
        self.min_elevation = -100
        self.scale_std = 4

    def __repr__(self):
        return f"CFG(min_elevation={self.min_elevation}, scale_std={self.scale_std})"


CFG = _Config()
