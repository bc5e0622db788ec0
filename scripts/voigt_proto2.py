"""Numpy model of the device fast Voigt (same formulas, same table) vs the reference profile."""
import re, sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
src = open(os.path.join(ROOT, "fake_spectra_b200/csrc/fsb_voigt_tables.h")).read()
body = src[src.index("{")+1:src.index("};")]
body = re.sub(r"//.*", "", body)
tab = np.array([float(v) for v in body.replace("\n", " ").split(",") if v.strip()]).reshape(11, 64)
SPI = np.sqrt(np.pi)

def coeffs(y):
    y2 = y*y
    pe = [(6 + 6*y2 + 3*y2**2 + y2**3)/6, -y2*(2 + 2*y2 + y2**2), 2*y2**2*(1 + y2)/3, -4*y2**3/45]
    a = [-y*(6 + 6*y2 + 3*y2**2 + y2**3)/3, 2*y**3*(2 + 2*y2 + y2**2)/3, -4*y**5*(1 + y2)/15, 8*y**7/315]
    b = [y**3*(70 + 49*y2 + 19*y2**2)/105, -2*y**5*(7 + 6*y2)/105, 4*y**7/315]
    return pe, [c/SPI for c in a], [c/SPI for c in b]

def fast(x, y):
    x = np.abs(np.asarray(x, float)); y = np.asarray(y, float)
    s = x*x
    pe, a, b = coeffs(y)
    out = np.zeros_like(x)
    xU2 = 37.0 - np.log(np.maximum(y, 1e-300))
    near = x < 16.0
    k = np.minimum((x*4).astype(int), 63)
    t = x - (k + 0.5)*0.25
    G = np.zeros_like(x)
    for j in range(10, -1, -1):
        G = G*t + tab[j, k]
    Pe = ((pe[3]*s + pe[2])*s + pe[1])*s + pe[0]
    A = ((a[3]*s + a[2])*s + a[1])*s + a[0]
    B = (b[2]*s + b[1])*s + b[0]
    U = np.where(s < xU2, np.exp(-s), 0.0)
    Hn = U*Pe + (G*A + B)
    # far wings: asymptotic series in 1/z^2
    r2 = s + y*y
    inv = 1/r2
    zr, zi = x*inv, -y*inv
    ur, ui = zr*zr - zi*zi, 2*zr*zi
    cs = [1.0, 0.5, 0.75, 1.875, 6.5625, 29.53125, 162.421875, 1055.7421875, 7918.06640625]
    Sr = np.full_like(x, cs[-1]); Si = np.zeros_like(x)
    for c in cs[-2::-1]:
        Sr, Si = Sr*ur - Si*ui + c, Sr*ui + Si*ur
    Hf = -(zr*Si + zi*Sr)/SPI
    return np.where(near, Hn, Hf)

if __name__ == "__main__":
    from oracle import Reference
    ref = Reference()
    rng = np.random.default_rng(3)
    n = 2000000
    x = np.concatenate([rng.uniform(0, 20, n), rng.uniform(0, 2000, n//4), rng.uniform(0,1e-3,1000), np.arange(0,16.5,0.25), np.arange(0,16.5,0.25)-1e-13])
    x = np.abs(x)
    for ymax_exp in [(-7,-4), (-4,-3), (-3,-2), (np.log10(0.01), np.log10(0.02)), (np.log10(0.02), np.log10(0.03)), (np.log10(0.03), np.log10(0.05))]:
        y = 10**rng.uniform(ymax_exp[0], ymax_exp[1], x.size)
        h = ref.profile(x, y)
        got = fast(x, y)
        rel = np.abs(got - h)/h
        i = rel.argmax()
        print("y in 10^[%.2f,%.2f]: max rel %.2e at x=%.4f y=%.3g ; 99.99pct %.2e" % (ymax_exp[0], ymax_exp[1], rel.max(), x[i], y[i], np.quantile(rel, 0.9999)))
    y = np.zeros_like(x); h = ref.profile(x, y); got = fast(x, y)
    m = h > 0
    print("y=0: max rel", np.max(np.abs(got[m]-h[m])/h[m]), "zero pattern", np.array_equal(got == 0, h == 0))
