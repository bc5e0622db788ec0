"""Host classes on the GPU: the resident route (particles + candidate index kept in HBM) and the
drop-in route (host arrays through _Particle_Interpolate) against the same classes driven by the
CPU oracle.  Tolerance 1e-10 relative with identical zero pattern (weighted means: 1e-6, they are
float32 quotients)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(__file__))
import cases  # noqa: E402
import hostcases  # noqa: E402
from fake_spectra_b200 import griddedspectra, randspectra  # noqa: E402

pytestmark = pytest.mark.gpu


def rand(nseg=1, **kw):
    return randspectra.RandSpectra(0, hostcases.snapshot(12, nseg), numlos=20, thresh=0., res=1.5, quiet=True, **kw)


@pytest.fixture(scope="module")
def want(oracle):
    rs = rand(backend=hostcases.OracleBackend(oracle))
    return {"tau": rs.get_tau("H", 1, 1215), "taub": rs.get_tau("H", 1, 1025), "col": rs.get_col_density("H", 1),
            "temp": rs.get_temp("H", 1), "vel": rs.get_velocity("H", 1), "dwd": rs.get_dens_weighted_density("H", 1)}


@pytest.mark.parametrize("resident", [True, False])
@pytest.mark.parametrize("nseg", [1, 3])
def test_spectra_routes_match_oracle(want, resident, nseg):
    rs = rand(nseg, resident=resident)
    for got, ref in ((rs.get_tau("H", 1, 1215), want["tau"]), (rs.get_col_density("H", 1), want["col"])):
        rel, same_zero = cases.rel_err(got, ref)
        assert same_zero and rel < 1e-10, rel
    both = rs.get_tau_lines("H", 1, [1215, 1025])
    rel, same_zero = cases.rel_err(both[1025], want["taub"])
    assert same_zero and rel < 1e-10, rel
    assert np.allclose(rs.get_temp("H", 1), want["temp"], rtol=1e-6)
    assert np.allclose(rs.get_velocity("H", 1), want["vel"], rtol=1e-4, atol=1e-3)
    assert np.allclose(rs.get_dens_weighted_density("H", 1), want["dwd"], rtol=1e-6)
    if resident:
        assert len(rs._engines) == nseg  # particles and index stayed in HBM across the calls


def test_engine_cache_reuses_the_index(monkeypatch):
    """Plain Spectra use (no replace_not_DLA): one upload and one candidate index per (segment, ion) serve every
    quantity; changing the sightlines drops them; the LRU bound releases the oldest engine."""
    from fake_spectra_b200 import native
    built = []
    orig = native.CandidateIndex.__init__

    def counting(self, *a, **k):
        built.append(1)
        return orig(self, *a, **k)
    monkeypatch.setattr(native.CandidateIndex, "__init__", counting)
    rs = rand(2, resident=True)
    rs.get_tau("H", 1, 1215)
    n0 = len(built)
    assert n0 == 2  # one per segment
    rs.get_tau("H", 1, 1025), rs.get_col_density("H", 1), rs.get_temp("H", 1), rs.get_velocity("H", 1)
    assert len(built) == n0
    rs.set_sightlines(rs.cofm[:7], rs.axis[:7])
    rs.get_tau("H", 1, 1215)
    assert len(built) == 2 * n0
    rs.max_engines = 1
    rs.set_sightlines(rs.cofm, rs.axis)
    rs.get_tau("H", 1, 1215)
    assert len(rs._engines) == 1


def test_gridded_all_axes_and_kernels(oracle):
    for kernel in ("cubic", "quintic", "tophat"):
        snap = hostcases.snapshot(10)
        ref = griddedspectra.GriddedSpectra(0, snap, nspec=3, res=2.0, axis=-1, quiet=True, kernel=kernel,
                                            backend=hostcases.OracleBackend(oracle)).get_tau("H", 1, 1215)
        got = griddedspectra.GriddedSpectra(0, snap, nspec=3, res=2.0, axis=-1, quiet=True, kernel=kernel).get_tau("H", 1, 1215)
        rel, same_zero = cases.rel_err(got, ref)
        assert same_zero and rel < 1e-10, (kernel, rel)


def test_voronoi_host(oracle):
    snap = hostcases.snapshot(8, nsegments=2, arepo=True)
    kw = dict(numlos=6, thresh=0., res=4.0, quiet=True, kernel="voronoi")
    ref = randspectra.RandSpectra(0, snap, backend=hostcases.OracleBackend(oracle), **kw)
    got = randspectra.RandSpectra(0, snap, **kw)
    for a, b in ((got.get_tau("H", 1, 1215), ref.get_tau("H", 1, 1215)), (got.get_col_density("H", 1), ref.get_col_density("H", 1))):
        rel, same_zero = cases.rel_err(a, b)
        assert same_zero and rel < 1e-10, rel


def test_product_never_imports_the_oracle():
    import subprocess
    code = ("import sys; import fake_spectra_b200.spectra, fake_spectra_b200.randspectra, fake_spectra_b200.griddedspectra, "
            "fake_spectra_b200.native; assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported'")
    subprocess.run([sys.executable, "-c", code], check=True, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
