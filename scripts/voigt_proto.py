"""Numpy prototype of the fast small-y Voigt expansion, checked against mpmath.
H(x,y) = U(x) Pe(s) - (2y/sqrt(pi)) [ G(x) Po(s) - Qo(s) ],  s = x^2, U = exp(-s), G = 1 - 2xF(x)."""
import numpy as np, mpmath as mp, sys
mp.mp.dps = 40
SPI = float(mp.sqrt(mp.pi))

def w_re(x, y):
    z = mp.mpc(x, y)
    return float(mp.re(mp.exp(-z*z) * mp.erfc(-1j*z)))

def G_mp(x):
    x = mp.mpf(x)
    return float(1 - 2*x*(mp.sqrt(mp.pi)/2*mp.exp(-x*x)*mp.erfi(x))) if x != 0 else 1.0

def hermite_polys(nmax):
    # p_n, q_n as numpy poly1d in x: p0=1,q0=0; p_{n+1}=p_n' - 2x p_n ; q_{n+1} = q_n' + (2/sqrt(pi)) p_n
    x = np.poly1d([1, 0])
    p = [np.poly1d([1.0])]; q = [np.poly1d([0.0])]
    for n in range(nmax):
        p.append(p[n].deriv() - 2*x*p[n])
        q.append(q[n].deriv() + (2/SPI)*p[n])
    return p, q

def fast(x, y, nmax, Gfun):
    """Series in y through order nmax, using exact U and G (tests truncation only)."""
    p, q = hermite_polys(nmax)
    s = x*x
    U = np.exp(-s)
    G = Gfun(x)
    F = np.where(x != 0, (1 - G)/(2*np.where(x == 0, 1, x)), 0.0)
    V = 2/SPI*F
    tot = np.zeros_like(x)
    fact = 1.0
    for n in range(nmax+1):
        if n > 0: fact *= n
        if n % 2 == 0:
            tot += (-1)**(n//2) * y**n / fact * p[n](x) * U
        else:
            tot -= (-1)**((n-1)//2) * y**n / fact * (p[n](x)*V + q[n](x))
    return tot

if __name__ == "__main__":
    rng = np.random.default_rng(0)
    xs = np.concatenate([np.linspace(0, 8, 161), rng.uniform(0, 16, 200)])
    Gv = np.array([G_mp(x) for x in xs])
    for y in [1e-4, 1e-3, 3e-3, 1e-2, 2e-2, 3e-2, 5e-2]:
        ex = np.array([w_re(x, y) for x in xs])
        row = []
        for nmax in (5, 6, 7, 9):
            got = fast(xs, y, nmax, lambda x: Gv)
            row.append("%.1e" % np.max(np.abs(got-ex)/ex))
        print("y=%g" % y, row)
