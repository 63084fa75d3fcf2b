"""Example: the reference's scripts/compute_topo_descriptors.py on the B200 path.

Same sequence of calls as the reference script (only the two import lines differ); the DEM comes from
``DEM.nc`` when xarray + a NetCDF backend are installed, else from the synthetic generator so that the example
runs anywhere a B200 is present.  With ``--resident`` the NaN census / fill runs on the GPU and the DEM stays
in HBM for every descriptor (prestage.fill_na_resident).

    python scripts/compute_topo_descriptors.py --size 4096 --outdir out
"""

import argparse
import logging
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import topo_descriptors_b200.helpers as hlp  # noqa: E402   (reference: import topo_descriptors.helpers as hlp)
import topo_descriptors_b200.topo as tp  # noqa: E402      (reference: import topo_descriptors.topo as tp)

logger = logging.getLogger(__name__)

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--dem", default="DEM.nc")
    ap.add_argument("--size", type=int, default=4096, help="edge of the synthetic DEM when --dem cannot be read")
    ap.add_argument("--outdir", default=".")
    ap.add_argument("--resident", action="store_true", help="NaN census + fill on the GPU, DEM stays in HBM")
    args = ap.parse_args()
    logging.basicConfig(level=logging.INFO)
    logging.captureWarnings(True)
    os.makedirs(args.outdir, exist_ok=True)

    # get the DEM
    try:
        dem_ds = hlp.get_dem_netcdf(args.dem)
    except (ImportError, OSError) as exc:
        from topo_descriptors_b200.synth import dem_dataset, tiled_fractal_dem

        logger.info(f"cannot read {args.dem} ({exc}); using a synthetic {args.size} x {args.size} DEM at 25 m")
        dem_ds = dem_dataset(tiled_fractal_dem(args.size, args.size, seed=2), res=25.0)
    if args.resident:
        from topo_descriptors_b200 import prestage

        ind_nans, dem_ds = prestage.fill_na_resident(dem_ds)
    else:
        ind_nans, dem_ds = hlp.fill_na(dem_ds)

    # define the target domain (here: the inner 80 % of the DEM)
    x, y = dem_ds["x"].values, dem_ds["y"].values
    nx, ny = len(x), len(y)
    domain = {"x": slice(x[nx // 10], x[-nx // 10]), "y": slice(y[ny // 10], y[-ny // 10])}

    # define the convolution scales in meters (the reference's own list, 100 m ... 100 km)
    scales_meters = [100, 300, 500, 1000, 2000, 4000, 6000, 10000, 20000, 30000, 60000, 100000]
    # valley / ridge: the rotated kernel of a scale must fit the largest FFT window (4096 px)
    res = abs(float(x[1] - x[0]))
    vr_scales = [sc for sc in scales_meters[3:] if sc / res * 1.4143 <= 4096]
    if len(vr_scales) < len(scales_meters[3:]):
        logger.info(f"valley/ridge: scales {scales_meters[3 + len(vr_scales):]} m exceed the 4096 px kernel extent of the FFT route, skipped")

    # smoothed DEM
    tp.compute_dem(dem_ds, scales_meters, ind_nans=ind_nans, crop=domain, outdir=args.outdir)
    # raw TPI, TPI with prior smoothing
    tp.compute_tpi(dem_ds, scales_meters, smth_factors=None, ind_nans=ind_nans, crop=domain, outdir=args.outdir)
    tp.compute_tpi(dem_ds, scales_meters, smth_factors=1, ind_nans=ind_nans, crop=domain, outdir=args.outdir)
    # gradients with symmetric kernels
    tp.compute_gradient(dem_ds, scales_meters, sig_ratios=1, ind_nans=ind_nans, crop=domain, outdir=args.outdir)
    # standard deviation of surface
    tp.compute_std(dem_ds, scales_meters, ind_nans=ind_nans, crop=domain, outdir=args.outdir)
    # valley / ridge index with prior smoothing
    tp.compute_valley_ridge(dem_ds, vr_scales, mode="valley", flat_list=[0, 0.2, 0.4], smth_factors=0.5,
                            ind_nans=ind_nans, crop=domain, outdir=args.outdir)
    tp.compute_valley_ridge(dem_ds, vr_scales, mode="ridge", flat_list=[0, 0.15, 0.3], smth_factors=0.5,
                            ind_nans=ind_nans, crop=domain, outdir=args.outdir)
    # Sx for one azimuth
    tp.compute_sx(dem_ds, 0, 1000, crop=domain, outdir=args.outdir)
