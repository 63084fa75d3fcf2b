"""yaconfigobject stub (test infrastructure only).

Carries the two constants of /root/reference/topo_descriptors/config/topo_descriptors.conf:1-5.
"""


class Config:
    def __init__(self, name=None, **kwargs):
        self.min_elevation = -100
        self.scale_std = 4
