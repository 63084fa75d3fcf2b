"""Host-side geometry that feeds the device kernels: the Sx ray samples and the valley/ridge
rotated-kernel bank.  Tiny, setup-time work (a few thousand integers / floats); the reference does
the same on the host (topo.py:828-853 for Sx, 466-531 for the kernel bank).
"""

import numpy as np


# ---------------------------------------------------------------------------------------------
# Sx geometry (reference: topo.py:861-925; pinned by test/test_topo.py)
# ---------------------------------------------------------------------------------------------
def sx_distance(radius, dx, dy):
    """Distance in metres of every cell of the (2*r_px+1)^2 search window from its centre
    (topo.py:861-878).  float64, shape (W, W), W = len(arange(2*r_px+1))."""
    r_px = max(radius / abs(dy), radius / abs(dx))
    extent = 2 * r_px + 1
    centre = np.floor(extent / 2)
    idx = np.arange(extent) - centre
    rows = (idx * dy)[:, None]
    cols = (idx * dx)[None, :]
    return np.sqrt(rows**2 + cols**2)


def sx_source_idx_delta(azimuths, radius, dx, dy):
    """(row, col) index offsets of the pixels at distance ``radius`` in the directions ``azimuths``
    (topo.py:881-892).  dy, dx are the SIGNED resolutions: on a north-up raster dy < 0, so azimuth 0
    points towards decreasing row.  int64, shape (n, 2)."""
    az = np.deg2rad(np.asarray(azimuths, dtype=np.float64))
    d_row = np.rint(radius / dy * np.cos(az))
    d_col = np.rint(radius / dx * np.sin(az))
    return np.stack([d_row, d_col], axis=1).astype(np.int64)


def sx_bresenhamlines(start, end):
    """Pixels on the straight lines from each ``start`` towards ``end`` (topo.py:895-925).

    Unit steps along each line's dominant axis (the other coordinate rounded half-to-even), starting
    one step away from the source; every line takes the same number of steps (the longest line's), a
    step is kept while the L1 distance to ``end`` does not grow w.r.t. the previous step, and the
    end pixel itself is dropped.  Returns an (M, 2) array with ``start``'s dtype, lines concatenated
    in input order.
    """
    start = np.asarray(start)
    end = np.asarray(end)
    delta = (end - start).astype(np.float64)
    reach = np.abs(delta).max(axis=1)
    n_steps = int(reach.max()) if reach.size else 0
    unit = np.divide(delta, reach[:, None], out=np.zeros_like(delta), where=reach[:, None] > 0)
    t = np.arange(1, n_steps + 1, dtype=np.float64)
    pts = np.rint(start[:, None, :] + unit[:, None, :] * t[None, :, None]).astype(start.dtype)  # (n, steps, 2)
    l1 = np.abs(pts - end).sum(axis=2)
    keep = np.ones(l1.shape, dtype=bool)
    keep[:, 1:] = l1[:, 1:] <= l1[:, :-1]
    keep &= ~np.all(pts == end.astype(start.dtype), axis=2)
    return pts[keep]


def sx_samples(window_distance, lines):
    """De-duplicated ray samples for the device scan: offsets (dy, dx) relative to the centre pixel
    and 1/distance.  Samples masked by ``radius_min`` (NaN distance) are dropped: np.nanmax skips
    them (topo.py:845, 951).  Duplicates (lines overlap near the centre) do not change a maximum.
    Sorted by (dy, dx) so that consecutive samples touch neighbouring DEM rows."""
    window = int(window_distance.shape[0] / 2)
    lines = np.asarray(lines, dtype=np.int64).reshape(-1, 2)
    if lines.size:
        lines = np.unique(lines, axis=0)
        dist = window_distance[lines[:, 0], lines[:, 1]]
        ok = ~np.isnan(dist)
        lines, dist = lines[ok], dist[ok]
    else:
        dist = np.zeros(0)
    offsets = (lines - window).astype(np.int32)
    with np.errstate(divide="ignore"):
        inv = (1.0 / dist).astype(np.float32)
    return np.ascontiguousarray(offsets), np.ascontiguousarray(inv), window


# ---------------------------------------------------------------------------------------------
# valley / ridge kernel bank (reference: topo.py:466-531)
# ---------------------------------------------------------------------------------------------
def valley_kernels(size, flat_list):
    """Z-scored V / U profiles (topo.py:466-499): |row - m| constant along the columns (0 deg = an
    east-west valley), with a flat bottom of half-width floor(floor(size*flat/2)+0.5) rows.
    float32, shape (F, size, size).  Like the reference, odd sizes only."""
    size = int(size)
    if size % 2 == 0:
        raise ValueError("valley/ridge kernels need an odd size (compute_valley_ridge always passes one)")
    m = size // 2
    out = np.empty((len(flat_list), size, size), dtype=np.float32)
    rows = np.abs(np.arange(size) - m).astype(np.float64)
    for f, flat in enumerate(flat_list):
        hw = int(np.floor(np.floor(size * flat / 2) + 0.5))
        prof = rows.copy()
        prof[m - hw : m + hw + 1] = prof[m - hw]
        k = np.repeat(prof[:, None], size, axis=1)
        out[f] = ((k - k.mean()) / k.std()).astype(np.float32)
    return out


def rotate_kernels(kernels, angle):
    """Rotate the kernel stack by ``angle`` degrees in its (row, col) plane with quadratic-spline
    interpolation, mark everything that falls outside the source square, z-score each kernel over its
    valid support and zero the rest (topo.py:521-531).  float32, shape (F, h', w')."""
    from scipy import ndimage  # the reference's own dependency; setup-time only

    sentinel = -9999
    rot = ndimage.rotate(kernels, angle, axes=(1, 2), reshape=True, order=2, mode="constant", cval=sentinel)
    valid = rot != sentinel
    out = np.zeros(rot.shape, dtype=np.float32)
    for f in range(rot.shape[0]):
        vals = rot[f][valid[f]].astype(np.float64)
        out[f][valid[f]] = ((vals - vals.mean()) / vals.std()).astype(np.float32)
    return out


ROT_ANGLE_DTYPE = np.dtype([("m", np.float64, (4,)), ("off", np.float64, (2,)), ("out_off", np.int64), ("oh", np.int32),
                            ("ow", np.int32)])  # 64 bytes: `RotAngle` of csrc/rotate.cu


def plan_rotations(shape_hw, n_kernels, angles):
    """What ``scipy.ndimage.rotate(reshape=True)`` computes before it interpolates (scipy/ndimage/_interpolation.py):
    rotation matrix from cosdg / sindg, output box, offset -- for every angle, plus the element offset of the angle's
    ``[n_kernels][oh][ow]`` block.  Returns (records, total elements)."""
    from scipy import special

    recs = np.zeros(len(angles), dtype=ROT_ANGLE_DTYPE)
    iy, ix = (int(v) for v in shape_hw)
    pos = 0
    for r, angle in zip(recs, angles):
        c, s = special.cosdg(angle), special.sindg(angle)
        rot = np.array([[c, s], [-s, c]])
        out_bounds = rot @ [[0, 0, iy, iy], [0, ix, 0, ix]]
        out_shape = (np.ptp(out_bounds, axis=1) + 0.5).astype(int)
        out_center = rot @ ((out_shape - 1) / 2)
        offset = (np.array([iy, ix]) - 1) / 2 - out_center
        r["m"] = rot.ravel()
        r["off"] = offset
        r["out_off"] = pos
        r["oh"], r["ow"] = out_shape
        pos += int(n_kernels) * int(out_shape[0]) * int(out_shape[1])
    return recs, pos


def spline_coefficients(kernels):
    """Quadratic-spline prefilter of the source kernels, exactly as ``ndimage.rotate(order=2, mode="constant")`` applies
    it (angle independent; float64)."""
    from scipy import ndimage

    return np.stack([ndimage.spline_filter(np.asarray(k, dtype=np.float64), order=2, output=np.float64, mode="constant")
                     for k in kernels])


def mix_channels(kernels_rot):
    """What the reference's 3-D ``signal.convolve(dem3d, kernels_rot, "same")`` (topo.py:431,443)
    applies per output channel: the DEM is broadcast along the flat-list axis, so the convolution
    along that axis sums kernels: channel m = sum of k_j over 0 <= m + s - j <= F-1, s = (F-1)//2
    (F = 3: k0+k1, k0+k1+k2, k1+k2)."""
    F = kernels_rot.shape[0]
    s = (F - 1) // 2
    out = np.zeros(kernels_rot.shape, dtype=np.float64)
    for m in range(F):
        lo, hi = max(0, m + s - (F - 1)), min(F - 1, m + s)
        out[m] = kernels_rot[lo : hi + 1].astype(np.float64).sum(axis=0)
    return out.astype(np.float32)


def build_valley_bank(size, mode, flat_list, angles=None, rotate=None):
    """Pack the 180-angle bank for ``topo_valley_ridge_f32`` (layout in include/topo_b200.h):
    per angle a [w][hp][4] float32 block of the channel-mixed kernels, flipped in both axes so the
    device correlates, rows h..hp-1 zero; plus, per kernel column, the 4-aligned row range (lo, n) outside
    which every weight is exactly zero (the corners of the rotated bounding box), so the device skips them.
    Flat lists longer than 4: ``"groups"`` holds one such bank per 4 mixed channels (the device keeps one running
    maximum across them).  ``rotate``: callable(kernels, angle) replacing :func:`rotate_kernels` (the device rotation)."""
    if mode not in ("valley", "ridge"):
        raise ValueError(f"Unknown mode {mode!r}")
    flat_list = list(flat_list)
    base = valley_kernels(size, flat_list)
    if mode == "ridge":
        base = base * np.float32(-1)
    if angles is None:
        angles = np.arange(0, 180, dtype=np.float32)
    rotate = rotate_kernels if rotate is None else rotate
    mixed_all = [mix_channels(rotate(base, ang)) for ang in angles]
    if base.shape[0] > 4:
        groups = [_pack_bank([m[c : c + 4] for m in mixed_all]) for c in range(0, base.shape[0], 4)]
        bank = {"groups": groups, "n_angles": len(mixed_all), "n_ch": base.shape[0],
                "hmax": max(g["hmax"] for g in groups), "wmax": max(g["wmax"] for g in groups)}
    else:
        bank = _pack_bank(mixed_all)
    bank["plain"] = _plain_bank(mixed_all)
    return bank


def _plain_bank(mixed_all):
    """The same kernels for the FFT route (``topo_valley_ridge_fft_f32``): scipy's orientation (not flipped), row-major,
    concatenated in angle order; per kernel its element offset and (h, w, angle index)."""
    flat, off, hw = [], [], []
    pos = 0
    for a, mixed in enumerate(mixed_all):
        for m in range(mixed.shape[0]):
            k = np.ascontiguousarray(mixed[m], dtype=np.float32)
            flat.append(k.ravel())
            off.append(pos)
            hw.append((k.shape[0], k.shape[1], a))
            pos += k.size
    return {"data": np.concatenate(flat), "off": np.array(off, dtype=np.int64), "hw": np.array(hw, dtype=np.int32)}


def _pack_bank(mixed_all):
    n_ch = mixed_all[0].shape[0]
    blocks, hw, off, cols = [], [], [], []
    pos = 0
    for mixed in mixed_all:
        flipped = mixed[:, ::-1, ::-1]
        F, h, w = flipped.shape
        hp = 4 * ((h + 3 + 3) // 4)
        blk = np.zeros((w, hp, 4), dtype=np.float32)
        blk[:, :h, :F] = np.transpose(flipped, (2, 1, 0))
        blocks.append(blk.ravel())
        hw.append((h, w, hp, len(cols)))
        nonzero = (blk != 0).any(axis=2)
        for j in range(w):
            rows = np.flatnonzero(nonzero[j])
            if len(rows) == 0:
                cols.append((0, 0))
                continue
            # the sliding window of 4 pixels needs weight rows first .. last + 3 (the tail rows are zero)
            lo = int(rows[0]) & ~3
            n = 4 * ((int(rows[-1]) + 4 - lo + 3) // 4)
            assert lo + n <= hp
            cols.append((lo, n))
        off.append(pos)
        pos += blk.size
    hw = np.array(hw, dtype=np.int32)
    return {
        "data": np.concatenate(blocks), "hw": hw, "off": np.array(off, dtype=np.int64),
        "cols": np.array(cols, dtype=np.int32).reshape(-1, 2),
        "n_angles": len(off), "n_ch": n_ch, "hmax": int(hw[:, 0].max()), "wmax": int(hw[:, 1].max()),
    }
