"""ctypes binding of libtopo_b200.so (C ABI: include/topo_b200.h).

There is NO CPU fallback: if the library has not been built, or no CUDA device is visible, every
descriptor call raises.  ``load()`` itself works without a GPU (symbols only), which is what the
CPU-side test-suite checks.
"""

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int64, c_longlong, c_size_t, c_void_p

from . import _build


class TopoError(RuntimeError):
    """A libtopo_b200 entry point returned an error code."""


class View(ctypes.Structure):
    """``topo_view`` (include/topo_b200.h): a row band of a global image."""

    _fields_ = [
        ("nx", c_int),
        ("gny", c_int),
        ("in_gy0", c_int),
        ("in_rows", c_int),
        ("out_gy0", c_int),
        ("out_rows", c_int),
    ]


_VP = POINTER(View)


class DiscCache(ctypes.Structure):
    """``topo_disc_cache`` (include/topo_b200.h): prefix planes shared by tpi / std calls at several sizes."""

    _fields_ = [("mem", c_void_p), ("bytes", c_size_t), ("max_size", c_int), ("valid", c_int), ("mask_size", c_int)]


_CP = POINTER(DiscCache)

# name -> (restype, argtypes); mirrors include/topo_b200.h declaration by declaration
PROTOTYPES = {
    "topo_version": (c_int, []),
    "topo_last_error": (c_char_p, []),
    "topo_launch_count": (c_longlong, []),
    "topo_set_option": (c_int, [c_char_p, c_int]),
    "topo_probe_dfma": (c_int, [c_int, c_void_p, c_void_p, c_void_p]),
    "topo_profile_enable": (c_int, [c_int]),
    "topo_profile_dump": (c_int, [c_char_p, c_size_t]),
    "topo_dem_stats_workspace_bytes": (c_size_t, [c_int, c_int]),
    "topo_dem_stats_f32": (c_int, [c_void_p, c_int, c_int, c_int64, c_void_p, c_void_p, c_size_t, c_void_p]),
    "topo_fill_f32": (c_int, [c_void_p, c_int, c_int, c_int64, c_float, c_void_p]),
    "topo_stamp_f32": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_float, c_void_p]),
    "topo_disc_workspace_bytes": (c_size_t, [_VP, c_int, c_int, c_int, c_double, c_double, c_int, c_int]),
    "topo_disc_shares_tsum": (c_int, [_VP, c_int, c_int, c_double, c_double, c_int]),
    "topo_disc_cache_bytes": (c_size_t, [_VP, c_int, c_int, c_double, c_double]),
    "topo_disc_plan_info": (c_int, [_VP, c_int, c_int, c_int, c_double, c_double, c_int, c_int, c_void_p]),
    "topo_tpi_f32": (c_int, [c_void_p, c_int64, c_void_p, c_int64, _VP, c_int, c_int, c_double, c_double,
                             c_void_p, c_int, _CP, c_void_p, c_size_t, c_void_p]),
    "topo_std_f32": (c_int, [c_void_p, c_int64, c_void_p, c_int64, _VP, c_int, c_int, c_double, c_double,
                             c_void_p, c_int, _CP, c_void_p, c_size_t, c_void_p]),
    "topo_gauss_workspace_bytes": (c_size_t, [_VP, c_int, c_int]),
    "topo_gauss_f32": (c_int, [c_void_p, c_int64, c_void_p, c_int64, _VP, c_void_p, c_int, c_void_p, c_int, c_int,
                               c_void_p, c_size_t, c_void_p]),
    "topo_grad_from_smooth_f32": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_int64, _VP, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "topo_gradient_workspace_bytes": (c_size_t, [_VP, c_int]),
    "topo_gradient_f32": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, _VP, c_void_p, c_int,
                                  c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "topo_sobel_gradient_f32": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, _VP,
                                        c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "topo_sx_f32": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, _VP, c_void_p, c_void_p, c_void_p,
                            c_int, c_int, c_float, c_int, c_int, c_int, c_int, c_void_p]),
    "topo_fill_na_f32": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_void_p, c_int, c_float, c_void_p,
                                 c_void_p]),
    "topo_nan_indices_f32": (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p]),
    "topo_zscore_f32": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_float, c_float, c_void_p]),
    "topo_rotate_bank_f32": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_float, c_void_p, c_void_p, c_void_p]),
    "topo_valley_ridge_fft_workspace_bytes": (c_size_t, [_VP, c_int, c_int]),
    "topo_valley_ridge_fft_f32": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, _VP, c_void_p, c_void_p, c_void_p,
                                          c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "topo_valley_ridge_f32": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, _VP, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
}

_LIB = None


def lib_path():
    return _build.LIB


def load():
    """Load the shared library and declare every prototype.  Raises if it is not built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing: build it with `python -m topo_descriptors_b200._build` "
            "(there is no CPU fallback)"
        )
    cdll = ctypes.CDLL(path)
    for name, (restype, argtypes) in PROTOTYPES.items():
        fn = getattr(cdll, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _LIB = cdll
    return cdll


def call(name, *args):
    """Call an int-returning entry point; raise TopoError with the library's message on failure."""
    cdll = load()
    rc = getattr(cdll, name)(*args)
    if rc != 0:
        msg = cdll.topo_last_error()
        raise TopoError(f"{name} failed ({rc}): {msg.decode() if msg else 'unknown error'}")


def set_option(name, value):
    """Execution-shape switch of the library (include/topo_b200.h: topo_set_option); results do not change."""
    call("topo_set_option", name.encode(), int(bool(value)))


def launch_count():
    return int(load().topo_launch_count())


def profile_enable(on=True):
    load().topo_profile_enable(1 if on else 0)


def profile_dump(aggregate=True):
    """Kernel timings since the last dump.  aggregate=True: {name: {"launches", "ms", "max_ms"}};
    aggregate=False: [(name, ms)] in launch order."""
    buf = ctypes.create_string_buffer(1 << 20)
    call("topo_profile_dump", buf, len(buf))
    records = []
    for line in buf.value.decode().splitlines():
        name, ms = line.rsplit(" ", 1)
        records.append((name, float(ms)))
    if not aggregate:
        return records
    out = {}
    for name, ms in records:
        a = out.setdefault(name, {"launches": 0, "ms": 0.0, "max_ms": 0.0})
        a["launches"] += 1
        a["ms"] += ms
        a["max_ms"] = max(a["max_ms"], ms)
    return out
