"""Design study for the fast device Voigt: piecewise polynomial tables for G(x) = 1 - 2 x F(x)
(F = Dawson) and accuracy of the small-y expansion against mpmath.  Offline tool."""
import sys
import numpy as np
import mpmath as mp
from numpy.polynomial import chebyshev as C

mp.mp.dps = 50

def G_exact(x):
    x = mp.mpf(x)
    if x == 0:
        return mp.mpf(1)
    F = mp.sqrt(mp.pi) / 2 * mp.exp(-x * x) * mp.erfi(x)
    return 1 - 2 * x * F

def fit_interval(lo, hi, deg):
    # Chebyshev interpolation at deg+1 nodes in high precision, converted to monomials in t = x - mid
    mid, half = (lo + hi) / 2, (hi - lo) / 2
    k = np.arange(deg + 1)
    nodes = np.cos(np.pi * (k + 0.5) / (deg + 1))
    vals = [G_exact(mid + half * float(n)) for n in nodes]
    # solve in mp for monomial coefficients in u = t/half
    A = mp.matrix(deg + 1, deg + 1)
    for i, n in enumerate(nodes):
        for j in range(deg + 1):
            A[i, j] = mp.mpf(float(n)) ** j
    c = mp.lu_solve(A, mp.matrix(vals))
    return [float(c[j] / mp.mpf(half) ** j) for j in range(deg + 1)], mid

def test(delta, deg, xmax):
    worst = 0
    for k in range(int(xmax / delta)):
        lo, hi = k * delta, (k + 1) * delta
        coef, mid = fit_interval(lo, hi, deg)
        xs = np.linspace(lo, hi, 41)
        t = xs - mid
        p = np.zeros_like(t)
        for cj in coef[::-1]:
            p = p * t + cj
        ex = np.array([float(G_exact(x)) for x in xs])
        # error measure: relative where |G| is not near its zero, else absolute/0.05
        err = np.abs(p - ex) / np.maximum(np.abs(ex), 0.02)
        worst = max(worst, err.max())
    return worst

if __name__ == "__main__":
    for delta, deg in [(0.25, 8), (0.25, 9), (0.25, 10), (0.125, 6), (0.125, 7), (0.125, 8), (0.5, 11), (0.5, 12)]:
        print(delta, deg, "%.2e" % test(delta, deg, 8.0), "%.2e" % test(delta, deg, 16.0) if delta >= 0.25 else "")
