"""CPU prototype of the octagon decomposition of a disc sum (index conventions for disc.cu).

disc = {(i, j): i^2 + j^2 <= m^2};  octagon O = {|i| <= u, |j| <= u, |i| + |j| <= u + v}, v = isqrt(m^2 - u^2)
  S(y, x) = SAT rect (|i| <= v, |j| <= u)
          + top / bottom trapezoids (rows v < |i| <= u, half-width u + v - |i|) from the sheared tables D1, D2
          + row caps (|i| > u) from the row prefix P, + column caps (|j| > u) from the column prefix C
          + corner diagonals (|i| + |j| = d > u + v, |i| <= u, |j| <= u) from the diagonal prefixes E1, E2
"""
import math
import numpy as np


def isqrt(n):
    return math.isqrt(n)


def plan(m):
    best = None
    dmax = max(p + isqrt(m * m - p * p) for p in range(m + 1))
    for u in range(int(m / math.sqrt(2)), m + 1):
        v = isqrt(m * m - u * u)
        if v > u:
            continue
        lines = 4 * (m - u) + 4 * (dmax - (u + v))
        if best is None or lines < best[0]:
            best = (lines, u, v)
    lines, u, v = best
    s = u + v
    diag = []  # per d: |i| range [p_lo, p_hi] of the corner points on |i| + |j| = d
    for d in range(s + 1, dmax + 1):
        ps = [p for p in range(max(0, d - u), min(u, d) + 1) if p * p + (d - p) * (d - p) <= m * m]
        # points with p = 0 or d - p = 0 lie on an axis: they belong to caps (|i| > u or |j| > u) -> excluded by p <= u, d-p <= u
        if ps:
            assert ps == list(range(ps[0], ps[-1] + 1))
            diag.append((d, ps[0], ps[-1]))
    return u, v, s, diag, lines


def tables(z):
    z = z.astype(np.int64)
    R, W = z.shape
    P = np.zeros((R, W + 1), np.int64)          # P[r][c] = sum z[r][:c]
    P[:, 1:] = np.cumsum(z, axis=1)
    C = np.zeros((R + 1, W), np.int64)          # C[r][c] = sum z[:r][c]
    C[1:] = np.cumsum(z, axis=0)
    S = np.zeros((R + 1, W + 1), np.int64)      # S[r][c] = sum z[:r][:c]
    S[1:, 1:] = np.cumsum(np.cumsum(z, axis=0), axis=1)
    # E1[r][c] = z[r][c] + E1[r-1][c-1]   (main diagonal, running down-right)
    # E2[r][c] = z[r][c] + E2[r-1][c+1]   (anti-diagonal, running down-left)
    E1 = z.copy()
    E2 = z.copy()
    for r in range(1, R):
        E1[r, 1:] += E1[r - 1, :-1]
        E2[r, :-1] += E2[r - 1, 1:]
    # D1[r][c] = P[r][c] + D1[r-1][c-1],  D2[r][c] = P[r][c] + D2[r-1][c+1]   (c in 0..W)
    D1 = P.copy()
    D2 = P.copy()
    for r in range(1, R):
        D1[r, 1:] += D1[r - 1, :-1]
        D2[r, :-1] += D2[r - 1, 1:]
    return P, C, S, E1, E2, D1, D2


def disc_sum_octagon(z, m, y, x, T, pl):
    P, C, S, E1, E2, D1, D2 = T
    u, v, s, diag, _ = pl
    tot = 0
    # rect rows [y-v, y+v], cols [x-u, x+u]
    tot += S[y + v + 1][x + u + 1] - S[y - v][x + u + 1] - S[y + v + 1][x - u] + S[y - v][x - u]
    # top trapezoid: rows i in [-u, -v-1], cols [x - (s+i), x + (s+i)]
    #   sum_i P[y+i][x+s+i+1]  : (row, col) both +1 per step -> D1 ;  sum_i P[y+i][x-s-i] : col -1 per step -> D2
    i0, i1 = -u, -v - 1
    if i1 >= i0:
        a = D1[y + i1][x + s + i1 + 1] - (D1[y + i0 - 1][x + s + i0] if True else 0)
        b = D2[y + i1][x - s - i1] - D2[y + i0 - 1][x - s - i0 + 1]
        tot += a - b
        # bottom trapezoid: rows i in [v+1, u], half-width s - i: cols [x-(s-i), x+(s-i)]
        #   sum_i P[y+i][x+s-i+1] : col -1 per step -> D2 ; sum_i P[y+i][x-s+i] : col +1 per step -> D1
        j0, j1 = v + 1, u
        a = D2[y + j1][x + s - j1 + 1] - D2[y + j0 - 1][x + s - j0 + 2]
        b = D1[y + j1][x - s + j1] - D1[y + j0 - 1][x - s + j0 - 1]
        tot += a - b
    # row caps |i| > u
    for i in range(u + 1, m + 1):
        h = isqrt(m * m - i * i)
        for r in (y - i, y + i):
            tot += P[r][x + h + 1] - P[r][x - h]
    # column caps |j| > u (their rows are all |i| <= u because u > m / sqrt 2)
    for j in range(u + 1, m + 1):
        h = isqrt(m * m - j * j)
        for c in (x - j, x + j):
            tot += C[y + h + 1][c] - C[y - h][c]
    # corner diagonals
    for d, p_lo, p_hi in diag:
        # quadrant (i>0, j>0): points (p, d-p), p = p_lo..p_hi: row +1, col -1 per step -> E2
        tot += E2[y + p_hi][x + d - p_hi] - E2[y + p_lo - 1][x + d - p_lo + 1]
        # quadrant (i<0, j<0): points (-p, -(d-p)): as p grows row -1, col +1 -> E2 with ends swapped
        tot += E2[y - p_lo][x - d + p_lo] - E2[y - p_hi - 1][x - d + p_hi + 1]
        # quadrant (i>0, j<0): points (p, -(d-p)): row +1, col +1 -> E1
        tot += E1[y + p_hi][x - d + p_hi] - E1[y + p_lo - 1][x - d + p_lo - 1]
        # quadrant (i<0, j>0): points (-p, d-p): row -1, col -1 as p grows -> E1 with ends swapped
        tot += E1[y - p_lo][x + d - p_lo] - E1[y - p_hi - 1][x + d - p_hi - 1]
    return tot


def disc_sum_direct(z, m, y, x):
    tot = 0
    for i in range(-m, m + 1):
        h = isqrt(m * m - i * i)
        tot += int(z[y + i, x - h : x + h + 1].sum())
    return tot


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for m in (20, 40, 61, 80, 120, 200, 400):
        pl = plan(m)
        u, v, s, diag, lines = pl
        print(f"m={m}: u={u} v={v} s={s} diagonals={len(diag)} lines={lines} (square hybrid: {8 * (m - int(m / math.sqrt(2)))})")
        if m > 130:
            continue
        n = 2 * m + 12
        z = rng.integers(0, 1000, (n + 2, n + 2))
        T = tables(z)
        for (y, x) in ((m + 2, m + 3), (m + 5, m + 1), (n - m - 1, n - m - 2)):
            a, b = disc_sum_octagon(z, m, y, x, T, pl), disc_sum_direct(z, m, y, x)
            assert a == b, (m, y, x, a, b)
    print("octagon decomposition == direct disc sum")
