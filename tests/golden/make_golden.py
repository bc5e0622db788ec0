"""Generates the committed golden fixtures in tests/golden/ by running the UNMODIFIED reference
(oracle/_ref/libfsref.so, built by oracle/build.py from /root/reference) on seeded inputs, and by
extracting the known-answer table held by the reference's own Faddeeva self-test.

Run in the build container (where /root/reference is mounted):

    python tests/golden/make_golden.py

Outputs (small .npz files, committed):
  faddeeva_w_kat.npz   the 57-point w(z) table of reference Faddeeva.cpp:1919-2108 (z and w(z))
  voigt_sweep.npz      Re w(x+iy) of the reference on a seeded (x, y) sweep covering all regimes
  case_random16.npz    16^3 particles, 40 sightlines cycling axes 1,2,3: candidate lists, near_lines,
                       tau + colden for kernels 0/1/3 and several lines (inputs stored too)
  case_grid12.npz      12^3 particles, 3*6^2 gridded sightlines (coordinates on 0.0)
  case_edge.npz        hand-built geometry (duplicates, box faces, exact-boundary predicates)
  case_voronoi8.npz    8^3 cells, kernel 2: assign_cells arrays, tau and colden
"""
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
from oracle import Reference  # noqa: E402

REF_SRC = "/root/reference/fake_spectra"


def _c_float(tok):
    tok = tok.strip()
    if tok in ("NaN", "-NaN"):
        return float("nan")
    if tok == "Inf":
        return float("inf")
    if tok == "-Inf":
        return float("-inf")
    return float(tok)


def faddeeva_table():
    """Parse the first z[NTST]/w[NTST] initialisers of the TEST_FADDEEVA main()."""
    src = open(os.path.join(REF_SRC, "Faddeeva.cpp")).read()
    start = src.index("w(z) tests")
    zblock = src[src.index("cmplx z[NTST] = {", start):]
    zblock = zblock[:zblock.index("};")]
    wblock = src[src.index("cmplx w[NTST] = {", start):]
    wblock = wblock[:wblock.index("};")]
    wblock = re.sub(r"/\*.*?\*/", "", wblock, flags=re.S)
    pair = re.compile(r"C\(\s*([^,()]+?)\s*,\s*([^,()]+?)\s*\)", re.S)
    z = [(_c_float(a), _c_float(b)) for a, b in pair.findall(zblock)]
    w = [(_c_float(a), _c_float(b)) for a, b in pair.findall(wblock)]
    assert len(z) == 57 and len(w) == 57, (len(z), len(w))
    z, w = np.array(z), np.array(w)
    np.savez(os.path.join(HERE, "faddeeva_w_kat.npz"), z_re=z[:, 0], z_im=z[:, 1], w_re=w[:, 0], w_im=w[:, 1])


def voigt_sweep(ref):
    rng = np.random.default_rng(2024)
    x = np.concatenate([rng.uniform(-12, 12, 6000), rng.uniform(-300, 300, 1500), rng.normal(0, 1e-3, 300),
                        np.array([0.0, 0.0, 5e-4, 4.9e-4, 6.0, 6.0000001, 8.0, 8.0000001, 10.0, 9.9999999, 28.0, 28.1])])
    y = 10 ** rng.uniform(-6, 1.2, x.size)
    y[:200] = 0.0                       # gamma = 0 (turn_off_selfshield): exp(-x^2) branch
    y[200:400] = 10 ** rng.uniform(-12, -9, 200)
    np.savez(os.path.join(HERE, "voigt_sweep.npz"), x=x, y=y, h=ref.profile(x, y))


def run_case(ref, d, name, configs, with_lists=True):
    out = {k: d[k] for k in ("pos", "vel", "dens", "temp", "h", "cofm", "axis")}
    out["box"] = np.float64(d["box"])
    if with_lists:
        offsets, part, dr2 = ref.near_particles(d["cofm"], d["axis"], d["box"], d["pos"], d["h"])
        out.update(offsets=offsets, part=part, dr2=dr2,
                   near_lines=ref.near_lines(d["box"], d["pos"], d["h"], d["axis"], d["cofm"]))
    for tag, kw in configs.items():
        p = cases.params(d, **kw)
        out["tau_" + tag] = ref.compute_tau(p["nbins"], p["kernel"], p["box"], p["velfac"], p["atime"], p["lambda_cm"],
                                            p["gamma"], p["fosc"], p["amumass"], p["tautail"], d["pos"], d["vel"],
                                            d["dens"], d["temp"], d["h"], d["axis"], d["cofm"])
        out["colden_" + tag] = ref.compute_colden(p["nbins"], p["kernel"], p["box"], p["velfac"], p["atime"],
                                                  p["lambda_cm"], p["gamma"], p["fosc"], p["amumass"], p["tautail"],
                                                  d["pos"], d["dens"], d["h"], d["axis"], d["cofm"])
    np.savez_compressed(os.path.join(HERE, name), **out)


# tag -> params() keyword arguments; shared with tests/test_golden.py
RANDOM16_CONFIGS = {
    "cubic_HI1215": dict(line="HI1215", kernel=1),
    "tophat_HI1215": dict(line="HI1215", kernel=0),
    "quintic_HI1025": dict(line="HI1025", kernel=3),
    "cubic_MgII2796_res10": dict(line="MgII2796", kernel=1, res=10.0),   # sub-sampled pixels
    "cubic_CIV1548": dict(line="CIV1548", kernel=1),
    "cubic_HI1215_gamma0": dict(line="HI1215", kernel=1, gamma_zero=True),
    "cubic_HI1215_odd": dict(line="HI1215", kernel=1, nbins=277),
    "cubic_HI1215_tiny": dict(line="HI1215", kernel=1, nbins=24),
}
GRID12_CONFIGS = {"cubic_HI1215": dict(line="HI1215", kernel=1), "quintic_HI1215": dict(line="HI1215", kernel=3)}
EDGE_CONFIGS = {"cubic_HI1215": dict(line="HI1215", kernel=1, nbins=2000, tautail=1e-5),
                "tophat_HI1215": dict(line="HI1215", kernel=0, nbins=2000, tautail=1e-5)}
VORONOI_CONFIGS = {"voronoi_HI1215": dict(line="HI1215", kernel=2)}


def voronoi_case(ref):
    d = cases.random_case(nside=8, nlos=6, axis="cycle", seed=11, los_seed=3)
    run_case(ref, d, "case_voronoi8.npz", VORONOI_CONFIGS)
    z = dict(np.load(os.path.join(HERE, "case_voronoi8.npz")))
    for line in range(d["cofm"].shape[0]):
        _, arr = ref.assign_cells(d["cofm"], d["axis"], d["box"], line, d["pos"], d["h"])
        z["cells_%d" % line] = arr
    np.savez_compressed(os.path.join(HERE, "case_voronoi8.npz"), **z)


def lattice_case():
    """Cells on a regular lattice, sightlines through points equidistant from 4, 2 and 1 cell columns, on all three
    axes: at every march point of assign_cells several candidates are at EXACTLY the same distance, and the reference
    gives the point to the first of them (strict <, index_table.cpp:181-190).  All coordinates are exactly
    representable, so the distances tie bit for bit."""
    box, n, a = 64.0, 4, 16.0
    g = (np.arange(n) + 0.5) * a
    pos = np.array([[x, y, z] for x in g for y in g for z in g], dtype=np.float32)
    npart = pos.shape[0]
    rng = np.random.default_rng(21)
    perp = [(16.0, 16.0), (16.0, 24.0), (24.0, 24.0), (32.0, 48.0), (8.0, 40.0), (17.5, 30.25)]
    cofm, axis = [], []
    for ax in (1, 2, 3):
        for (u, v) in perp:
            c = [0.0, 0.0, 0.0]
            others = [i for i in range(3) if i != ax - 1]
            c[others[0]], c[others[1]] = u, v
            cofm.append(c)
            axis.append(ax)
    return {"box": box, "cofm": np.array(cofm, dtype=np.float64), "axis": np.array(axis, dtype=np.int32), "pos": pos,
            "h": np.full(npart, 14.0, dtype=np.float32),
            "vel": (30 * rng.standard_normal((npart, 3))).astype(np.float32),
            "dens": (1e11 * (1 + rng.random(npart))).astype(np.float32),
            "temp": (1e4 * (0.5 + rng.random(npart))).astype(np.float32)}


def voronoi_lattice_case(ref):
    d = lattice_case()
    run_case(ref, d, "case_voronoi_lattice.npz", VORONOI_CONFIGS)
    z = dict(np.load(os.path.join(HERE, "case_voronoi_lattice.npz")))
    for line in range(d["cofm"].shape[0]):
        _, arr = ref.assign_cells(d["cofm"], d["axis"], d["box"], line, d["pos"], d["h"])
        z["cells_%d" % line] = arr
    np.savez_compressed(os.path.join(HERE, "case_voronoi_lattice.npz"), **z)


def main():
    ref = Reference()
    faddeeva_table()
    voigt_sweep(ref)
    run_case(ref, cases.random_case(), "case_random16.npz", RANDOM16_CONFIGS)
    run_case(ref, cases.grid_case(), "case_grid12.npz", GRID12_CONFIGS)
    run_case(ref, cases.edge_case(), "case_edge.npz", EDGE_CONFIGS)
    voronoi_case(ref)
    voronoi_lattice_case(ref)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print("%-24s %8d bytes" % (f, os.path.getsize(os.path.join(HERE, f))))


if __name__ == "__main__":
    main()
