"""Shared, seeded parity cases (inputs at the native boundary) used by the oracle tests, the
golden-fixture generator and the GPU parity tests."""
import numpy as np

from fake_spectra_b200 import synthetic as syn

# (lambda_cm, gamma, fosc, amumass): SURVEY App. F / reference atom.dat
LINES = {
    "HI1215": (1215.6701e-8, 6.265e8, 0.4164, 1.00794),
    "HI1025": (1025.7223e-8, 1.897e8, 0.07912, 1.00794),
    "CIV1548": (1548.2049e-8, 2.642e8, 0.1899, 12.011),
    "MgII2796": (2796.3542699e-8, 2.68e8, 0.6155, 24.305),
}
TAUTAIL = 1e-7  # reference spectra.py:135


def params(case, line="HI1215", kernel=1, res=1.0, gamma_zero=False, nbins=None, tautail=TAUTAIL):
    """Scalar arguments in _Particle_Interpolate order (reference py_module.cpp:115)."""
    lam, gam, fosc, amu = LINES[line]
    cos = syn.Cosmology()
    velfac = float(cos.velfac)
    if nbins is None:
        nbins = int(case["box"] * velfac / res)
    return dict(nbins=nbins, kernel=kernel, box=case["box"], velfac=velfac, atime=cos.atime, lambda_cm=lam,
                gamma=0.0 if gamma_zero else gam, fosc=fosc, amumass=amu, tautail=tautail)


def random_case(nside=16, nlos=40, axis="cycle", seed=42, los_seed=23, metal_scale=1.0):
    d = syn.boundary_arrays(nside, seed=seed, metal_scale=metal_scale)
    cofm, ax = syn.random_sightlines(d["box"], nlos, seed=los_seed, axis=axis)
    d["cofm"], d["axis"] = cofm, ax
    return d


def grid_case(nside=12, nspec=6, seed=7):
    """All three axes on a regular grid: many sightlines sit exactly on coordinate 0.0
    (reference griddedspectra.py:59-88)."""
    d = syn.boundary_arrays(nside, seed=seed)
    cofm, ax = syn.grid_sightlines(d["box"], nspec, axis=-1)
    d["cofm"], d["axis"] = cofm, ax
    return d


def edge_case():
    """Hand-built geometry: duplicates, box faces, exact-boundary predicates."""
    box = 10000.0
    cofm = np.array([[4000, 4000, 4000], [4000, 4000, 4000], [4000, 4020, 4010], [4000, 4000, 4010],
                     [0, 0.4, 0.1], [0, 10000 - 0.4, 0.3], [0, 2000, 1000], [1000, 2000, 500],
                     [1000, 2000, 500], [1000, 2000, 500], [3000, 5000, 550], [8000, 5500, 9000],
                     [6000, 5500, 9000], [0, 0, 0], [0, 10000, 10000], [5000, 0, 10000],
                     [9999.5, 3.0, 9998.0], [2.0, 9999.0, 1.0]], dtype=np.float64)
    axis = np.array([1, 1, 1, 1, 1, 1, 1, 3, 2, 1, 3, 3, 1, 1, 2, 3, 2, 3], dtype=np.int32)
    pos = np.array([[500, 2000, 1000], [5000, 2000, 1000], [5010, 2010, 990], [4000, 4000, 4000],
                    [1000, 9999.9, 9999.9], [1000.5, 2000, 501], [7500, 7500, 7500], [4008, 4008, 4008],
                    [2000, 9999, 9999.8], [4000, 4016, 4010], [4000, 3984, 4010], [4000, 4000, 4026],
                    [1.0, 1.0, 1.0], [9999.0, 9999.0, 9999.0], [5000, 0.5, 9999.5], [0.25, 9999.75, 0.5],
                    [9998.5, 2.0, 9999.0], [4000, 4012, 4016]], dtype=np.float32)
    h = np.array([1, 1, 20, 25, 0.6, 1.5, 7, 10, 0.8, 16, 16, 16, 3, 3, 2, 4, 5, 20], dtype=np.float32)
    rng = np.random.default_rng(5)
    n = pos.shape[0]
    return {"box": box, "cofm": cofm, "axis": axis, "pos": pos, "h": h,
            "vel": (50 * rng.standard_normal((n, 3))).astype(np.float32),
            "dens": (1e11 * (1 + rng.random(n))).astype(np.float32),
            "temp": (1e4 * (0.5 + rng.random(n))).astype(np.float32)}


def rel_err(a, b):
    """max |a-b|/|b| over b != 0, plus whether the zero patterns agree."""
    a, b = np.asarray(a), np.asarray(b)
    m = b != 0
    rel = float(np.max(np.abs(a - b)[m] / np.abs(b[m]))) if m.any() else 0.0
    return rel, bool(np.array_equal(a == 0, b == 0))
