"""The pin of every GPU parity test, run again on the GPU box: the C restatement the GPU tests compare with
(oracle/_build/libfsoracle.so) against the UNMODIFIED reference (oracle/_ref/libfsref.so, which travels with the
snapshot).  No GPU work here; the marker only makes the driver's `-m gpu` tier execute it too, and a direct
GPU-vs-reference check of tau and column density follows."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(__file__))
import cases  # noqa: E402
import test_oracle_vs_ref as pin  # noqa: E402

pytestmark = pytest.mark.gpu


def test_oracle_is_pinned_to_the_reference(oracle, reference):
    pin.test_candidates_larger(oracle, reference)
    pin.test_negative_weights_and_cold_gas(oracle, reference)
    pin.test_voigt_dense_sweep(oracle, reference)
    pin.test_random_snapshot(oracle, reference, 1, "HI1215", 1.0)
    pin.test_random_snapshot(oracle, reference, 3, "MgII2796", 10.0)


@pytest.mark.parametrize("line,res", [("HI1215", 1.0), ("CIV1548", 1.0), ("MgII2796", 1.0), ("MgII2796", 5.0)])
def test_gpu_against_the_reference_itself(reference, line, res):
    """The CUDA path against libfsref.so directly (not through the restatement), metal lines included: their kernels
    are much wider than the thermal width, so their marches take the node-by-node mixed route."""
    from fake_spectra_b200 import _spectra_priv as priv
    d = cases.random_case(nside=20, nlos=60, axis="cycle", seed=11, los_seed=5, metal_scale=1.0 if line == "HI1215" else 1e-4)
    p = cases.params(d, line=line, res=res)
    args = (p["nbins"], p["kernel"], p["box"], p["velfac"], p["atime"], p["lambda_cm"], p["gamma"], p["fosc"], p["amumass"],
            p["tautail"], d["pos"], d["vel"], d["dens"], d["temp"], d["h"], d["axis"], d["cofm"])
    got = priv._Particle_Interpolate(1, *args)
    want = reference.compute_tau(**p, pos=d["pos"], vel=d["vel"], dens=d["dens"], temp=d["temp"], h=d["h"], axis=d["axis"],
                                 cofm=d["cofm"])
    rel, same_zero = cases.rel_err(got, want)
    assert same_zero and rel < 1e-10, rel
    got = priv._Particle_Interpolate(0, *args)
    want = reference.compute_colden(**p, pos=d["pos"], dens=d["dens"], h=d["h"], axis=d["axis"], cofm=d["cofm"])
    rel, same_zero = cases.rel_err(got, want)
    assert same_zero and rel < 1e-10, rel
