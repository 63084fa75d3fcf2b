"""Seeded synthetic DEMs: spectral-synthesis fractal terrain.

Used by the tests, ``bench.py`` and the oracle's golden-vector generator so that the GPU path and
the CPU reference see the *same* input.  numpy only.
"""

import numpy as np

from ._xr import Dataset


def fractal_dem(ny, nx, seed=0, zmin=200.0, zmax=3400.0, beta=1.7, integer=False, dtype=np.float32):
    """Fractal DEM: uniform random phases on the rfft2 half-plane, amplitude ~ k^-beta, DC = 0,
    inverse FFT, affine rescale to [zmin, zmax].  ``integer=True`` rounds to whole metres
    (SRTM-like; the only input class on which the reference ``std`` is meaningful).
    """
    rng = np.random.default_rng(seed)
    ky = np.fft.fftfreq(ny)[:, None]
    kx = np.fft.rfftfreq(nx)[None, :]
    k = np.hypot(ky, kx)
    k[0, 0] = 1.0
    amp = k ** (-beta)
    amp[0, 0] = 0.0
    phase = rng.uniform(0.0, 2.0 * np.pi, size=amp.shape)
    spec = amp * np.exp(1j * phase)
    z = np.fft.irfft2(spec, s=(ny, nx))
    lo, hi = z.min(), z.max()
    z = zmin + (z - lo) * ((zmax - zmin) / (hi - lo))
    if integer:
        z = np.rint(z)
    return np.ascontiguousarray(z.astype(dtype))


def tiled_fractal_dem(ny, nx, seed=0, tile=2048, rows=None, **kw):
    """Large DEMs without a large FFT: a ``tile`` x ``tile`` fractal patch mirrored/tiled to
    (ny, nx) plus a smooth large-scale trend.  Deterministic, cheap, keeps realistic local relief;
    used only where the full spectral synthesis would take minutes of host time (16384^2 and up).
    ``rows=(r0, r1)`` returns only those rows of the (ny, nx) DEM (what one rank of a row-band job needs).
    """
    r0, r1 = (0, ny) if rows is None else (int(rows[0]), int(rows[1]))
    base = fractal_dem(tile, tile, seed=seed, **kw).astype(np.float32)
    sym = np.concatenate([base, base[:, ::-1]], axis=1)
    sym = np.concatenate([sym, sym[::-1, :]], axis=0)  # 2*tile periodic, continuous
    rx = -(-nx // sym.shape[1])
    z = np.tile(sym[np.arange(r0, r1) % sym.shape[0]], (1, rx))[:, :nx]
    yy = np.linspace(0.0, 1.0, ny, dtype=np.float32)[r0:r1, None]
    xx = np.linspace(0.0, 1.0, nx, dtype=np.float32)[None, :]
    z = z + np.float32(150.0) * np.sin(np.float32(2 * np.pi) * yy) * np.cos(np.float32(2 * np.pi) * xx)
    if kw.get("integer"):
        z = np.rint(z)
    return np.ascontiguousarray(z.astype(np.float32))


def dem_dataset(z, res=30.0, x0=2600000.0, y0=1200000.0, crs="epsg:2056", name="alti"):
    """Wrap an array as a north-up projected-CRS Dataset: x ascending, y descending (y_res < 0)."""
    ny, nx = z.shape
    x = x0 + res * np.arange(nx, dtype=np.float64)
    y = y0 - res * np.arange(ny, dtype=np.float64)
    return Dataset({name: (("y", "x"), z)}, coords={"x": x, "y": y}, attrs={"crs": crs})
