"""CPU models of the arithmetic behind the round-2 FFT disc route and the packed tiny kernel: the identities the CUDA
code relies on, restated in numpy (the kernels themselves are compared with the oracle and with the prefix-plane walk
in the GPU suite)."""

import numpy as np


def _dft3(a, b, c, inv):
    """csrc/fft_smem.cuh: dft3 -- s = b + c, d = b - c, m = a - s/2, t = (sqrt 3 / 2) d; x1 = m -+ i t, x2 = m +- i t."""
    h = np.sqrt(3.0) / 2.0
    s, d = b + c, b - c
    m, t = a - 0.5 * s, h * d
    p, q = m - 1j * t, m + 1j * t
    return a + s, (q if inv else p), (p if inv else q)


def _dft6(v, inv):
    """csrc/fft_smem.cuh: dft6 -- two 3-point transforms of the even / odd samples + a radix-2 combination."""
    e = _dft3(v[0], v[2], v[4], inv)
    o = _dft3(v[1], v[3], v[5], inv)
    w = np.exp((1j if inv else -1j) * np.pi / 3.0)
    out = [None] * 6
    for k in range(3):
        out[k] = e[k] + w**k * o[k]
        out[k + 3] = e[k] - w**k * o[k]
    return np.array(out)


def test_radix6_butterfly_is_the_six_point_dft():
    rng = np.random.default_rng(0)
    for _ in range(20):
        v = rng.normal(size=6) + 1j * rng.normal(size=6)
        assert np.allclose(_dft6(v, False), np.fft.fft(v), atol=1e-13)
        assert np.allclose(_dft6(v, True), np.fft.ifft(v) * 6, atol=1e-13)


def test_leading_radix6_stage_followed_by_length_512_transforms_is_the_3072_point_dft():
    """Decimation in frequency with a leading radix-6 stage (FftShape<3072>): sub-block k of length 512 holds
    dft6(x[j + q 512])[k] * w^(j k); its 512-point transform gives the frequencies 6 f + k."""
    rng = np.random.default_rng(1)
    n, st = 3072, 512
    x = rng.normal(size=n) + 1j * rng.normal(size=n)
    want = np.fft.fft(x)
    j = np.arange(st)
    blocks = np.stack([x[q * st : (q + 1) * st] for q in range(6)])  # blocks[q][j] = x[j + q * 512]
    y = np.stack([_dft6(blocks[:, jj], False) for jj in range(st)], axis=1)  # y[k][j]
    for k in range(6):
        sub = np.fft.fft(y[k] * np.exp(-2j * np.pi * j * k / n))
        assert np.allclose(sub, want[k::6], atol=1e-9)


def test_twin_tiles_two_tiles_of_one_real_plane_in_one_complex_transform():
    """csrc/disc.cu, dfft_fwd_planes_kernel: the mask is real, so the cyclic convolution of a + i b with it is
    (a * m) + i (b * m) -- the rounded real / imaginary parts are the exact integer sums of the two tiles."""
    rng = np.random.default_rng(2)
    n = 64
    a = rng.integers(0, 3400, size=(n, n)).astype(np.float64)
    b = rng.integers(0, 2**22, size=(n, n)).astype(np.float64)
    m = (rng.random((n, n)) < 0.2).astype(np.float64)
    z = np.fft.ifft2(np.fft.fft2(a + 1j * b) * np.fft.fft2(m))
    ca = np.fft.ifft2(np.fft.fft2(a) * np.fft.fft2(m)).real
    cb = np.fft.ifft2(np.fft.fft2(b) * np.fft.fft2(m)).real
    assert np.array_equal(np.rint(z.real), np.rint(ca)) and np.array_equal(np.rint(z.imag), np.rint(cb))
    # and the rounded values are the integer sums themselves
    direct = sum(np.roll(np.roll(a, dy, 0), dx, 1) for dy in range(n) for dx in range(n) if m[dy, dx])
    assert np.array_equal(np.rint(z.real), direct)


def test_packed_word_of_the_tiny_float_std():
    """csrc/disc.cu, disc_tiny_kernel<STD_F>: cell word = (t - tmin) << fpack | F, F = (frac + 1) 2^Sf <= 2^(Sf+1),
    fpack = Sf + 2; with W = sum of words and T = sum of (t - tmin), both modulo 2^32, the fraction sum is
    (W - (T << fpack)) mod 2^32 as long as it is below 2^32 itself."""
    rng = np.random.default_rng(3)
    sf, tmin = 18, 0
    fpack = sf + 2
    for n_cells in (21, 69, 137):
        z = rng.uniform(200.0, 3400.0, size=n_cells).astype(np.float32)
        t = np.trunc(z).astype(np.int64)
        f = np.rint(((z - t.astype(np.float32)) + np.float32(1.0)).astype(np.float64) * 2.0**sf).astype(np.int64)
        assert int(f.max()) <= 2 ** (sf + 1) and int((t - tmin).max()) < 2 ** (32 - fpack)
        words = ((t - tmin) << fpack) | f
        assert np.array_equal(words, ((t - tmin) << fpack) + f)  # the fields do not overlap
        W = int(words.sum()) % 2**32
        T = int((t - tmin).sum()) % 2**32
        assert (W - ((T << fpack) % 2**32)) % 2**32 == int(f.sum()) < 2**32
        # every float32 fraction of |z| >= 2^(23 - Sf) is a multiple of 2^-Sf: the field is exact
        assert np.array_equal(f.astype(np.float64) * 2.0**-sf - 1.0, z.astype(np.float64) - t)


def test_window_rows_3072_versus_4096():
    """csrc/disc.cu, plan_fft_geometry: window rows (transform length along y x tile rows) for a band of `rows` output
    rows and halo 400 -- what the planner's choice buys on 8, 4, 2 and 1 GPUs of a 16384-row DEM."""

    def window_rows(rows, ty, h=400):
        return -(-rows // (ty - 2 * h)) * ty

    assert (window_rows(2048, 3072), window_rows(2048, 4096)) == (3072, 4096)      # 8 GPUs: 3072
    assert (window_rows(4096, 3072), window_rows(4096, 4096)) == (6144, 8192)      # 4 GPUs: 3072
    assert (window_rows(8192, 3072), window_rows(8192, 4096)) == (12288, 12288)    # 2 GPUs: tie -> 4096
    assert (window_rows(16384, 3072), window_rows(16384, 4096)) == (24576, 20480)  # 1 GPU: 4096
