"""What would it cost in fidelity to accumulate the Gaussian taps beyond 2.5 sigma in float32?  

Model of the proposed kernel on the CPU: taps |k| <= 2.5 sigma in float64 (as today), the outer taps in float32
blocks of 16 around the centre value x_c (block sums added to the float64 accumulator), float32 rounding between the
two axes like scipy.  Reports the fraction of pixels whose float32 result differs from scipy's sequence and the
largest differences of the smoothed DEM, of dx / dy and of the aspect (where |grad| > 0.2).
"""
import sys, os
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle as O  # noqa: E402
from topo_descriptors_b200.synth import fractal_dem  # noqa: E402


def weights(sigma):
    lw = int(4.0 * sigma + 0.5)
    x = np.arange(-lw, lw + 1, dtype=np.float64)
    w = np.exp(-0.5 * (x / sigma) ** 2)
    return w / w.sum(), lw


def axis_pass(a, sigma, split):
    """Correlate along axis 0 with reflect borders; a is float32 (ny, nx); returns float32."""
    w, lw = weights(sigma)
    ny = a.shape[0]
    idx = np.arange(-lw, ny + lw)
    idx = np.where(idx < 0, -idx - 1, idx)
    idx = np.where(idx >= ny, 2 * ny - 1 - idx, idx)
    while (idx < 0).any() or (idx >= ny).any():
        idx = np.where(idx < 0, -idx - 1, idx)
        idx = np.where(idx >= ny, 2 * ny - 1 - idx, idx)
    pad = a[idx]
    inner = int(2.5 * sigma)
    acc = np.zeros(a.shape, np.float64)
    xc = a.astype(np.float32)
    tail_w = 0.0
    blk = np.zeros(a.shape, np.float32)
    nblk = 0
    for k in range(-lw, lw + 1):
        src = pad[k + lw : k + lw + ny]
        if not split or abs(k) <= inner:
            acc += w[k + lw] * src.astype(np.float64)
        else:
            blk += np.float32(w[k + lw]) * (src - xc)      # float32 FMA model: product and sum rounded to float32
            tail_w += float(np.float32(w[k + lw]))
            nblk += 1
            if nblk == 16:
                acc += blk.astype(np.float64)
                blk[:] = 0
                nblk = 0
    if split:
        acc += blk.astype(np.float64) + tail_w * xc.astype(np.float64)
    return acc.astype(np.float32)


def smooth(z, sigma, split):
    t = axis_pass(z, sigma, split)
    return axis_pass(t.T.copy(), sigma, split).T.copy()


if __name__ == "__main__":
    z = fractal_dem(600, 800, seed=0)
    for sigma in (16.75, 50.25):
        ref = smooth(z, sigma, False)
        new = smooth(z, sigma, True)
        assert np.array_equal(ref, O.gaussian_filter_restated(z, sigma)) or np.abs(ref - O.gaussian_filter_restated(z, sigma)).max() < 1e-3
        diff = np.abs(ref.astype(np.float64) - new.astype(np.float64))
        gy, gx = np.gradient(ref.astype(np.float32))
        hy, hx = np.gradient(new.astype(np.float32))
        dx, dy = gx / 25.0, gy / -25.0
        ex, ey = hx / 25.0, hy / -25.0
        asp = lambda a, b: (180 + np.degrees(np.arctan2(a, b))) % 360  # noqa: E731
        steep = np.hypot(dx, dy) > 0.2
        da = np.abs(asp(dx, dy) - asp(ex, ey))
        da = np.minimum(da, 360 - da)
        print(f"sigma {sigma}: pixels that differ {100 * (diff > 0).mean():.2f} %, max |dem diff| {diff.max():.2e} m, "
              f"max |d(dx,dy)| {max(np.abs(dx - ex).max(), np.abs(dy - ey).max()):.2e}, "
              f"max aspect diff where |grad| > 0.2: {da[steep].max() if steep.any() else 0:.2e} deg ({steep.mean() * 100:.1f} % of px)")
