"""Oracle (C restatement) against the committed golden fixtures, which were produced by the
UNMODIFIED reference (tests/golden/make_golden.py).  CPU only.

Tolerances: candidate lists, near_lines and dr^2 bit-exact; tau / colden 1e-12 relative with an
identical zero pattern (the reference's own -ffast-math noise is 3.5e-14)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(__file__))
import cases  # noqa: E402
from golden import make_golden as mg  # noqa: E402

FIXTURES = [("case_random16.npz", mg.RANDOM16_CONFIGS), ("case_grid12.npz", mg.GRID12_CONFIGS),
            ("case_edge.npz", mg.EDGE_CONFIGS), ("case_voronoi8.npz", mg.VORONOI_CONFIGS),
            ("case_voronoi_lattice.npz", mg.VORONOI_CONFIGS)]


def load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name))
    d = {k: z[k] for k in z.files}
    d["box"] = float(d["box"])
    return d


@pytest.mark.parametrize("name,configs", FIXTURES)
def test_candidate_lists(oracle, golden_dir, name, configs):
    d = load(golden_dir, name)
    off, part, dr2 = oracle.near_particles(d["cofm"], d["axis"], d["box"], d["pos"], d["h"])
    assert np.array_equal(off, d["offsets"])
    assert np.array_equal(part, d["part"])
    assert np.array_equal(dr2, d["dr2"])          # bit-exact doubles
    assert np.array_equal(oracle.near_lines(d["box"], d["pos"], d["h"], d["axis"], d["cofm"]), d["near_lines"])


@pytest.mark.parametrize("name,configs", FIXTURES)
def test_tau_colden(oracle, golden_dir, name, configs):
    d = load(golden_dir, name)
    for tag, kw in configs.items():
        p = cases.params(d, **kw)
        tau = oracle.compute_tau(**p, pos=d["pos"], vel=d["vel"], dens=d["dens"], temp=d["temp"], h=d["h"],
                                 axis=d["axis"], cofm=d["cofm"])
        rel, same_zero = cases.rel_err(tau, d["tau_" + tag])
        assert same_zero and rel < 1e-12, (name, tag, rel)
        col = oracle.compute_colden(**p, pos=d["pos"], dens=d["dens"], h=d["h"], axis=d["axis"], cofm=d["cofm"])
        rel, same_zero = cases.rel_err(col, d["colden_" + tag])
        assert same_zero and rel < 1e-12, (name, tag, rel)


@pytest.mark.parametrize("fixture", ["case_voronoi8.npz", "case_voronoi_lattice.npz"])
def test_voronoi_cells(oracle, golden_dir, fixture):
    d = load(golden_dir, fixture)
    for line in range(d["cofm"].shape[0]):
        err, arr = oracle.assign_cells(d["cofm"], d["axis"], d["box"], line, d["pos"], d["h"])
        assert err == 0
        assert np.array_equal(arr, d["cells_%d" % line])   # float32, bit-exact


def test_voigt_sweep(oracle, golden_dir):
    z = np.load(os.path.join(golden_dir, "voigt_sweep.npz"))
    got = oracle.profile(z["x"], z["y"])
    rel, same_zero = cases.rel_err(got, z["h"])
    assert same_zero and rel < 1e-13, rel
