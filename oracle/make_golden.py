"""Generate tests/golden/ref_small.npz by RUNNING THE REFERENCE ITSELF  --  test infrastructure.

Run in the authoring container only (needs /root/reference):

    python oracle/make_golden.py

Every array stored is the output of an unmodified ``topo_descriptors`` function (imported through
``oracle/ref_runner.py``) on a small seeded DEM that is stored alongside.  Key naming:
``<function>__<case>``.  The reference's tests pin nothing on tpi/std/gradient/valley_ridge/sx
(SURVEY.md section 4), so these vectors are the parity pin for the oracle and, through it, for
the CUDA path.

Tight pins where the reference's float32 FFT noise would hide errors:

* ``tpi_tight``: the reference fed ``int64(dem) << 40``: scipy then picks exact direct
  convolution in float64 (choose_conv_method's integer-overflow rule), result / 2^40 is the exact
  TPI of the integer DEM (error < 1e-12).
* ``std_f64*``: the reference fed a float64 DEM (float64 FFT of the data; the kernel's float32
  FFT leaves ~3e-3 m on a 200..3400 m DEM and ~3e-5 on the 0..40 m integer DEM ``zc``).
"""

import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_runner  # noqa: E402
from topo_descriptors_b200.synth import fractal_dem  # noqa: E402

NY, NX, RES = 64, 96, 30.0


def main():
    warnings.simplefilter("ignore")
    topo, hlp = ref_runner.load()
    g = {}

    z = fractal_dem(NY, NX, seed=7)
    zi = fractal_dem(NY, NX, seed=7, integer=True)
    zc = np.rint(fractal_dem(NY, NX, seed=9, zmin=0.0, zmax=40.0)).astype(np.float32)
    x = 2600000.0 + RES * np.arange(NX, dtype=np.float64)
    y = 1200000.0 - RES * np.arange(NY, dtype=np.float64)
    g["in__z"], g["in__zi"], g["in__zc"], g["in__x"], g["in__y"] = z, zi, zc, x, y
    ds = ref_runner.fake_dataset(z, x, y)

    # ---- helpers (a17) -------------------------------------------------------------------
    scales = [100, 200, 500, 2000, 3100]
    px, res = hlp.scale_to_pixel(scales, ds)
    g["scale_to_pixel__scales"] = np.array(scales)
    g["scale_to_pixel__px"] = px
    g["scale_to_pixel__res_x"] = res["x"]
    g["scale_to_pixel__res_y"] = res["y"]
    sig = hlp.get_sigmas([None, 0.5, 0, 1, 2.5], px)
    g["get_sigmas__out"] = np.array([np.nan if s is None else s for s in sig])
    g["round_up_to_odd__in"] = np.arange(0.1, 10, 0.7)
    g["round_up_to_odd__out"] = hlp.round_up_to_odd(np.arange(0.1, 10, 0.7))
    for size in (3, 4, 5, 6, 7, 17, 33):
        g[f"circular_kernel__{size}"] = topo.circular_kernel(size)

    # ---- tpi (a2) ------------------------------------------------------------------------
    for size in (3, 5, 6, 7, 17, 33):
        g[f"tpi_lit__{size}"] = topo.tpi(z, size)
        g[f"tpi_tight__{size}"] = topo.tpi(zi.astype(np.int64) << 40, size) / 2.0**40
    g["tpi_lit_sigma__7_1.75"] = topo.tpi(z, 7, sigma=1.75)

    # ---- std (a3) ------------------------------------------------------------------------
    for size in (3, 5, 7, 17):
        g[f"std_lit__{size}"] = topo.std(zi, size)
        g[f"std_f64__{size}"] = topo.std(zi.astype(np.float64), size)
        g[f"std_f64c__{size}"] = topo.std(zc.astype(np.float64), size)
        g[f"std_f64flt__{size}"] = topo.std(z.astype(np.float64), size)
    g["std_f64c_sigma__7_1.75"] = topo.std(zc.astype(np.float64), 7, sigma=1.75)

    # ---- dem / gradient / sobel (a4-a7) ----------------------------------------------------
    g["dem__3.3"] = topo.dem(z, 3.3)
    g["dem__20"] = topo.dem(z, 20.0)  # radius 80 > ny: multiple reflections
    sdx, sdy = topo.sobel(z)
    g["sobel__dx"], g["sobel__dy"] = sdx, sdy
    for sigma, ratio in [(0.75, 1), (1.75, 1), (4.25, 1), (4.25, 1.5), (16.75, 1)]:
        out = topo.gradient(z, sigma, res, sig_ratio=ratio)
        for nm, arr in zip(("dx", "dy", "slope", "aspect"), out):
            g[f"gradient__{sigma}_{ratio}_{nm}"] = arr
    flat = np.full((16, 24), 512.25, dtype=np.float32)
    out = topo.gradient(flat, 1.75, {"x": np.full(24, 30.0), "y": np.full(16, -30.0)})
    g["gradient_flat__aspect"] = out[3]
    g["gradient_flat__slope"] = out[2]
    res2d = {"x": np.full((NY, NX), 28.5) + np.linspace(0, 3, NX)[None, :],
             "y": np.full((NY, NX), -31.0) - np.linspace(0, 1, NY)[:, None]}
    out = topo.gradient(z, 1.75, res2d)
    g["gradient_res2d__x"], g["gradient_res2d__y"] = res2d["x"], res2d["y"]
    for nm, arr in zip(("dx", "dy", "slope", "aspect"), out):
        g[f"gradient_res2d__{nm}"] = arr

    # ---- valley / ridge (a8-a10) -----------------------------------------------------------
    g["valley_kernels__7"] = topo._valley_kernels(7, [0, 0.15, 0.3])
    g["valley_kernels__11"] = topo._valley_kernels(11, [0, 0.2, 0.4])
    for ang in (0, 30, 45, 90, 137):
        g[f"rotate_kernels__7_{ang}"] = topo._rotate_kernels(topo._valley_kernels(7, [0, 0.15, 0.3]), np.float32(ang))
    vn, vd = topo.valley_ridge(z, 7, "valley")
    g["valley_ridge__valley7_norm"], g["valley_ridge__valley7_dir"] = vn, vd
    vn, vd = topo.valley_ridge(z, 9, "ridge", flat_list=[0, 0.2, 0.4], sigma=1.125)
    g["valley_ridge__ridge9_norm"], g["valley_ridge__ridge9_dir"] = vn, vd
    vn, vd = topo.valley_ridge(z, 5, "valley", flat_list=[0, 0.3])
    g["valley_ridge__valley5f2_norm"], g["valley_ridge__valley5f2_dir"] = vn, vd

    # ---- sx (a11-a15) ----------------------------------------------------------------------
    g["sx_distance__150_50_40"] = topo._sx_distance(150.0, 50.0, 40.0)
    g["sx_distance__150_30_-30"] = topo._sx_distance(150.0, 30.0, -30.0)
    g["sx_source_idx_delta__a"] = topo._sx_source_idx_delta(np.array([3.0, 4.0, 5.0, 6.0]), 500, 20, 30)
    g["sx_source_idx_delta__b"] = topo._sx_source_idx_delta(np.linspace(-5, 5, 15), 150.0, 30.0, -30.0)
    g["sx_bresenhamlines__a"] = topo._sx_bresenhamlines(np.array([[8, 9], [17, 22]]), np.array([15, 15]))
    sx_cases = [(0, 150, 0.0, 10.0, 10.0, 15), (45, 150, 0.0, 10.0, 10.0, 15), (200, 300, 60.0, 10.0, 10.0, 15),
                (270, 150, 0.0, 2.0, 0.0, 15), (135, 240, 0.0, 10.0, 30.0, 7)]
    g["sx__cases"] = np.array(sx_cases, dtype=np.float64)
    for i, (az, rad, rmin, h, arc, steps) in enumerate(sx_cases):
        g[f"sx__case{i}"] = topo.sx(ds, az, rad, height=h, azimuth_arc=arc, azimuth_steps=steps, radius_min=rmin)
    zn = z.copy()
    zn[20, 30] = np.nan
    zn[40, 70] = np.nan
    g["sx_nan__in"] = zn
    g["sx_nan__out"] = topo.sx(ref_runner.fake_dataset(zn, x, y), 45, 150)
    g["sx_allmasked__out"] = topo.sx(ds, 0, 150, radius_min=1000.0)

    # ---- output names of the compute_* drivers (a16) ---------------------------------------
    names = [
        topo._dem_name(200), topo._tpi_name(200, None), topo._tpi_name(2000, 0.5), topo._std_name(200, 1),
        *topo._valley_ridge_names(1000, "valley", 0.5), *topo._gradient_names(200, 1), *topo._gradient_names(2000, 1.5),
        topo._sx_name(500.0, 45.0),
    ]
    g["names__all"] = np.array(names)

    out = os.path.join(ROOT, "tests", "golden", "ref_small.npz")
    np.savez_compressed(out, **g)
    print(f"wrote {out}: {len(g)} arrays, {os.path.getsize(out) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
