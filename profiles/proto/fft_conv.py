"""CPU prototype of the in-place shared-memory FFT the wide Gaussian uses (csrc/gauss_fft.cu): forward
decimation-in-frequency stages leave the spectrum digit-reversed, the point-wise product is taken in that order
(the multiplier goes through the same forward transform), inverse decimation-in-time stages mirror the forward ones
and return natural order -- no permutation pass at all.  Two real lines ride in one complex transform (the filter is
real).  Pins the index conventions; run: python profiles/proto/fft_conv.py
"""
import numpy as np


def plan(N):
    """Radix list: a leading radix 2 / 4 stage when N is not a power of 8, then radix 8."""
    r, n = [], N
    while n % 8 == 0 and n > 1:
        n //= 8
    lead = n  # 1, 2 or 4
    assert lead in (1, 2, 4)
    k = int(round(np.log(N // lead) / np.log(8)))
    return ([lead] if lead > 1 else []) + [8] * k


def forward(x, radices, tw):
    N = len(x)
    x = x.copy()
    M = N
    for R in radices:
        S = M // R
        for u in range(N // R):
            block, j = divmod(u, S)
            base = block * M + j
            q = x[base + S * np.arange(R)]
            y = np.array([np.sum(q * np.exp(-2j * np.pi * np.arange(R) * k / R)) for k in range(R)])
            y = y * tw[(j * np.arange(R) * (N // M)) % N]
            x[base + S * np.arange(R)] = y
        M = S
    return x


def inverse(x, radices, tw):
    N = len(x)
    x = x.copy()
    Ms = []
    M = N
    for R in radices:
        Ms.append(M)
        M //= R
    for R, M in zip(reversed(radices), reversed(Ms)):
        S = M // R
        for u in range(N // R):
            block, j = divmod(u, S)
            base = block * M + j
            y = x[base + S * np.arange(R)] * np.conj(tw[(j * np.arange(R) * (N // M)) % N])
            q = np.array([np.sum(y * np.exp(2j * np.pi * np.arange(R) * k / R)) for k in range(R)])
            x[base + S * np.arange(R)] = q
    return x


def check(N, lw, seed=0):
    rng = np.random.default_rng(seed)
    tw = np.exp(-2j * np.pi * np.arange(N) / N)
    rad = plan(N)
    sigma = lw / 4.0
    xs = np.arange(-lw, lw + 1)
    w = np.exp(-0.5 * xs**2 / sigma**2)
    w /= w.sum()
    h = np.zeros(N)
    h[0] = w[lw]
    for m in range(1, lw + 1):
        h[m] = h[N - m] = w[lw + m]
    H = forward(h.astype(complex), rad, tw).real / N  # digit-reversed order, 1/N folded in
    a, b = rng.uniform(200, 3400, N), rng.uniform(200, 3400, N)
    y = inverse(forward(a + 1j * b, rad, tw) * H, rad, tw)
    ref_a = np.array([np.dot(w, a[n - lw : n + lw + 1]) for n in range(lw, N - lw)])
    ref_b = np.array([np.dot(w, b[n - lw : n + lw + 1]) for n in range(lw, N - lw)])
    ea = np.abs(y.real[lw : N - lw] - ref_a).max()
    eb = np.abs(y.imag[lw : N - lw] - ref_b).max()
    print(f"N={N} radices={rad} lw={lw}: max err {ea:.2e} {eb:.2e}")
    assert ea < 1e-9 and eb < 1e-9


if __name__ == "__main__":
    check(64, 5)
    check(128, 9)
    check(256, 20)
    check(512, 33)
