"""CPU prototype of the device kernel-bank rotation (csrc/rotate.cu): scipy.ndimage.rotate(kernels, angle, axes=(1, 2),
reshape=True, order=2, mode="constant", cval=-9999) restated -- the reference's call in _rotate_kernels (topo.py:524-526).

Host part (cheap, kept on scipy so that it is bit-identical by construction): cosdg / sindg, the output shape, the
offset, and the quadratic-spline prefilter of the F source kernels (angle independent).  Device part restated here in
numpy: per output pixel the source coordinate (same operation order as NI_GeometricTransform's affine branch, no fused
multiply-add), the outside test of mode="constant", the three quadratic B-spline weights per axis and the 9-tap sum with
mirrored coefficient indices at the edges.  Run: python profiles/proto/rotate_restated.py  (compares with scipy for
every angle 0..179 and several sizes: identical validity masks, values to ~1e-15).
"""
import numpy as np
from scipy import ndimage, special


def rotation_setup(shape_hw, angle):
    """What scipy.ndimage.rotate computes before calling affine_transform (scipy/ndimage/_interpolation.py rotate)."""
    c, s = special.cosdg(angle), special.sindg(angle)
    rot = np.array([[c, s], [-s, c]])
    iy, ix = shape_hw
    out_bounds = rot @ [[0, 0, iy, iy], [0, ix, 0, ix]]
    out_shape = (np.ptp(out_bounds, axis=1) + 0.5).astype(int)
    out_center = rot @ ((out_shape - 1) / 2)
    in_center = (np.array(shape_hw) - 1) / 2
    return rot, in_center - out_center, out_shape


def mirror(i, n):
    """Index extension of the spline coefficients for taps beyond the array (mirror: d c b | a b c d | c b a)."""
    if n == 1:
        return np.zeros_like(i)
    p = 2 * n - 2
    i = np.abs(i) % p
    return np.where(i >= n, p - i, i)


def rotate_restated(kernels, angle, cval=-9999.0):
    kernels = np.asarray(kernels)
    F, H, W = kernels.shape
    rot, off, (oh, ow) = rotation_setup((H, W), angle)
    coef = np.stack([ndimage.spline_filter(k.astype(np.float64), order=2, output=np.float64, mode="constant") for k in kernels])
    oy, ox = np.mgrid[0:oh, 0:ow].astype(np.float64)
    # icoor[h] = 0 + o_y * m[h][0]; += o_x * m[h][1]; += shift[h]   (separate roundings)
    cy = (oy * rot[0, 0] + ox * rot[0, 1]) + off[0]
    cx = (oy * rot[1, 0] + ox * rot[1, 1]) + off[1]
    outside = (cy < 0) | (cy > H - 1) | (cx < 0) | (cx > W - 1)

    def taps(c):
        start = np.floor(c + 0.5).astype(np.int64) - 1
        x = c - np.floor(c + 0.5)
        w1 = 0.75 - x * x
        t = 0.5 - x
        w0 = 0.5 * t * t
        w2 = 1.0 - w0 - w1
        return start, (w0, w1, w2)

    sy, wy = taps(cy)
    sx, wx = taps(cx)
    out = np.empty((F, oh, ow), dtype=np.float64)
    for f in range(F):
        acc = np.zeros((oh, ow))
        for i in range(3):
            yi = mirror(sy + i, H)
            for j in range(3):
                xj = mirror(sx + j, W)
                acc = acc + (coef[f][yi, xj] * wy[i]) * wx[j]
        out[f] = np.where(outside, cval, acc)
    return out.astype(kernels.dtype)


if __name__ == "__main__":
    import sys

    sys.path.insert(0, ".")
    from topo_descriptors_b200 import _geometry as geo

    worst = 0.0
    for size in (5, 9, 21, 41, 67):
        k = geo.valley_kernels(size, [0, 0.15, 0.3])
        for angle in np.arange(0, 180, dtype=np.float32):
            want = ndimage.rotate(k, angle, axes=(1, 2), reshape=True, order=2, mode="constant", cval=-9999)
            got = rotate_restated(k, angle)
            assert got.shape == want.shape, (size, angle, got.shape, want.shape)
            assert np.array_equal(got == -9999, want == -9999), (size, angle, "mask")
            d = np.abs(got.astype(np.float64) - want)[want != -9999]
            worst = max(worst, float(d.max()) if d.size else 0.0)
            bits = np.mean(got != want)
            assert d.max() <= 1e-6, (size, angle, d.max())
        print(f"size {size}: masks identical for 180 angles, worst value diff so far {worst:.2e}, last bit-diff fraction {bits:.2e}")
