"""Snapshot fields -> interpolation inputs on the device (csrc/fsb_prep.cu) against the host numpy route
that follows the reference (Spectra._read_particle_data = spectra.py:550-617 with abstractsnapshot.py:114-154,253-282,
gas_properties.py:104-146, convert_cloudy.py:167-200).

Tolerances: selection, positions, Gadget smoothing lengths and velocities are bit-equal (same float32 / float64
operations); Arepo smoothing lengths Volume^(1/3) within one float32 ulp (numpy's float32 power is a SIMD routine that is
not correctly rounded -- it differs from glibc's powf for 20 % of arguments -- while the kernel returns the correctly
rounded power); temperatures bit-equal; species densities within 2.5e-7 relative (two float32 ulps: the Rahmati neutral fraction is
evaluated in double by numpy and by the kernel and rounded to float32 once, the libraries' pow / exp differ in the last
bits of the double); ion fractions from a smooth Cloudy
table within 2e-5 relative (numpy's float32 log10 is an ulp off the correctly rounded value the kernel uses for half of
its arguments; the difference scales with the table's slope), and within 1e-6 of the restatement with correctly rounded
logarithms on a table with Cloudy's -30 plateau (27 dex per cell at its edge).  The optical depths
computed from the device-prepared arrays are then checked against the CPU oracle run on the SAME arrays at 1e-10, so
the end-to-end chain keeps the kernel's parity bar."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(__file__))
import cases  # noqa: E402
import hostcases  # noqa: E402
from test_cloudy import make_table  # noqa: E402
from fake_spectra_b200 import cloudy, randspectra  # noqa: E402

pytestmark = pytest.mark.gpu


def snapshot(nside=14, nseg=1, arepo=False, dense=True):
    snap = hostcases.snapshot(nside, nseg, arepo=arepo)
    if dense:
        # push every seventh particle above the star-formation threshold (Rahmati branch of get_reproc_HI) and give a
        # few particles no carbon (the _filter_particles branch) and a non-positive internal energy (temperature floor)
        snap.fields["Density"][::7] *= np.float32(3e4)
        snap.fields["GFM_Metals"][::5, 2] = 0
        snap.fields["InternalEnergy"][::11] = 0
    return snap


def spectra(snap, **kw):
    table, reds = make_table()
    rs = randspectra.RandSpectra(0, snap, numlos=24, thresh=0., res=2., quiet=True, **kw)
    rs.cloudy_table = cloudy.CloudyTable(rs.red, table=table, reds=reds)
    return rs


@pytest.mark.parametrize("arepo", [False, True])
@pytest.mark.parametrize("elem,ion", [("H", 1), ("He", -1), ("C", 4), ("Mg", 2), ("C", -1)])
def test_device_prep_matches_host_route(elem, ion, arepo):
    rs = spectra(snapshot(arepo=arepo, nseg=2))
    assert rs._device_prep_ok(elem, ion)
    for fn in range(2):
        want = rs._read_particle_data(fn, elem, ion, True)
        got = rs._device_particle_data(fn, elem, ion)
        assert got[5] == want[5]
        pos, vel, den, temp, hh = [g.cpu().numpy() for g in got[:5]]
        assert np.array_equal(pos, want[0])
        assert np.array_equal(hh, want[4]) if not arepo else np.max(np.abs(hh - want[4]) / want[4]) < 1.2e-7
        assert np.array_equal(vel, want[1])
        assert np.array_equal(temp, want[3])
        assert temp.min() >= 1.0 and (temp == 1.0).any()
        tol = 2e-5 if (ion > 0 and elem != "H") else 2.5e-7
        assert den.dtype == np.float32 and (np.all(den > 0) if ion > 0 else (elem != "C" or (den == 0).any()))
        ok = want[2] > 0
        assert np.array_equal(den == 0, want[2] == 0) and np.max(np.abs(den[ok] - want[2][ok]) / want[2][ok]) < tol
    if elem == "H":
        thr = rs.gasprop.PhysDensThresh / 0.76 / rs.gasprop._density_conversion()
        assert (rs.snapshot_set.fields["Density"] > thr).any()  # the self-shielding branch was exercised


def test_ion_lookup_on_a_plateau_table():
    """N V of the synthetic table carries a -30 plateau with a sharp edge: the kernel against the numpy restatement of what
    it evaluates (test_cloudy.bspline_eval, itself pinned to scipy's map_coordinates) with correctly rounded float32
    logarithms, for every particle of the snapshot."""
    import torch
    from test_cloudy import bspline_eval
    from fake_spectra_b200 import _lib, native
    rs = spectra(snapshot())
    tb = rs.cloudy_table
    snap, gp = rs.snapshot_set, rs.gasprop
    dev = torch.device("cuda", 0)
    up = lambda name: torch.from_numpy(snap.get_data(0, name, segment=0)).to(dev)  # noqa: E731
    cfg = _lib.Prep()
    cfg.velocity_factor, cfg.dens_conv, cfg.rscale = np.sqrt(rs.atime), gp._density_conversion(), rs.rscale
    cfg.unit_ienergy = rs.units.UnitInternalEnergy_in_cgs
    cfg.temp_factor = (rs.units.gamma - 1) * rs.units.protonmass / rs.units.boltzmann
    cfg.hy_mass, cfg.amumass = 0.76, np.float32(14.0067)
    metals = up("GFM_Metals")
    ion_table, _owner = tb.device_table("N", 5, dev)
    hh = native.smoothing_lengths(up("SmoothingLength"), mode=0)
    got = native.prepare_particles(cfg, None, up("Position"), up("Velocities"), up("Density"), up("InternalEnergy"),
                                   up("ElectronAbundance"), None, hh, metals[:, 3], ion_table=ion_table)
    den = gp.get_code_rhoH(0, segment=0).astype(np.float32)
    temp = gp.get_temp(0, segment=0).astype(np.float32)
    temp[temp <= 0] = 1
    assert np.array_equal(got[3].cpu().numpy(), temp)
    ed = (den * np.float32(rs.rscale)) * snap.get_data(0, "GFM_Metals", segment=0)[:, 3]
    nh = np.clip(den, np.float32(tb.get_dens_bounds()[0]), np.float32(tb.get_dens_bounds()[1])) * np.float32(0.774132)
    tt = np.clip(temp, np.float32(tb.get_temp_bounds()[0]), np.float32(tb.get_temp_bounds()[1]))
    c0 = (np.log10(nh.astype(np.float64)).astype(np.float32) - tb.dens[0]) * (tb.dens.size - 1) / (tb.dens[-1] - tb.dens[0])
    c1 = (np.log10(tt.astype(np.float64)).astype(np.float32) - tb.temp[0]) * (tb.temp.size - 1) / (tb.temp[-1] - tb.temp[0])
    ions = bspline_eval(tb, tb.spline_coefficients("N", 5), c0, c1)
    want = (ed * np.float32(10 ** ions)) / np.float32(14.0067)
    assert (ions < -20).any() and (ions > -8).any()  # both sides of the plateau's edge were sampled
    assert np.max(np.abs(got[2].cpu().numpy() - want) / want) < 1e-6


@pytest.mark.parametrize("peculiar", [False, True])
def test_device_prep_from_a_bigfile_snapshot(tmp_path, peculiar):
    """A snapshot on disc (MP-Gadget layout): header values are numpy float64, so numpy forms temperatures in double and
    velocities as v / a in double; the kernel's second arithmetic mode reproduces both bit for bit."""
    from fake_spectra_b200 import abstractsnapshot as absn, synthetic
    path = synthetic.write_bigfile(snapshot(arepo=False), str(tmp_path), num=1, nfile=3, peculiar=peculiar)
    table, reds = make_table()
    rs = randspectra.RandSpectra(1, path, numlos=24, thresh=0., res=2., quiet=True)
    rs.cloudy_table = cloudy.CloudyTable(rs.red, table=table, reds=reds)
    assert isinstance(rs.snapshot_set, absn.BigFileSnapshot) and isinstance(rs.units.UnitInternalEnergy_in_cgs, np.floating)
    for elem, ion in (("H", 1), ("C", 4)):
        want = rs._read_particle_data(0, elem, ion, True)
        got = [g.cpu().numpy() for g in rs._device_particle_data(0, elem, ion)[:5]]
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[4], want[4])
        assert np.array_equal(got[1], want[1])
        assert np.array_equal(got[3], want[3].astype(np.float32))
        assert np.max(np.abs(got[2] - want[2]) / want[2]) < (2.5e-7 if elem == "H" else 2e-5)
    tau = rs.get_tau("H", 1, 1215)
    assert rs._engines and np.isfinite(tau).all() and tau.max() > 0


def test_smoothing_length_from_masses():
    """Third case of get_smooth_length (abstractsnapshot.py:276-281): no Volume, no SmoothingLength."""
    import torch
    from fake_spectra_b200 import native
    rng = np.random.default_rng(2)
    mass = rng.uniform(0.5, 2, 5000).astype(np.float32)
    dens = rng.uniform(1e-3, 7, 5000).astype(np.float32)
    got = native.smoothing_lengths(torch.from_numpy(mass).cuda(), torch.from_numpy(dens).cuda(), mode=2).cpu().numpy()
    want = np.power(mass / dens, 1. / 3)
    assert np.max(np.abs(got - want) / want) < 1.2e-7
    assert np.array_equal(got, np.power((mass / dens).astype(np.float64), np.float64(np.float32(1. / 3))).astype(np.float32))


@pytest.mark.parametrize("elem,ion,line", [("H", 1, 1215), ("C", 4, 1548)])
def test_tau_from_device_prepared_particles(oracle, elem, ion, line):
    """Spectra.get_tau on the device-prepared route == the CPU oracle on the very same prepared arrays (1e-10), and
    within the input-rounding tolerance of the host-prepared route."""
    rs = spectra(snapshot())
    got = rs.get_tau(elem, ion, line)
    assert rs._engines, "the resident route with device preparation did not run"
    pos, vel, den, temp, hh = [g.cpu().numpy() for g in rs._device_particle_data(0, elem, ion)[:5]]
    ln = rs.lines[(elem, ion)][line]
    want = oracle.compute_tau(rs.nbins, rs.kernel_int, rs.box, rs.velfac, rs.atime, ln.lambda_X * 1e-8, ln.gamma_X, ln.fosc_X,
                              rs.lines.get_mass(elem), rs.tautail, pos, vel, den, temp, hh, rs.axis, rs.cofm)
    rel, same_zero = cases.rel_err(got, want)
    assert same_zero and rel < 1e-10, rel
    host = spectra(snapshot(), device_prep=False).get_tau(elem, ion, line)
    big = host > 1e-6 * host.max()
    assert np.max(np.abs(got[big] - host[big]) / host[big]) < 1e-4
