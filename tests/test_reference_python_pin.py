"""The host-side restatements of SURVEY 8f row f3 (gas_properties, cloudy, unitsystem, spec_utils, line_data) against
outputs of the reference's own PYTHON modules, imported unmodified from /root/reference in the build container by
tests/golden/make_golden_prep.py and committed as tests/golden/prep_reference.npz (inputs stored next to the outputs).

These pin the host-prepared route of Spectra._read_particle_data; the device route (fsb_prepare_particles) is checked
against that host route on the GPU tier (tests/test_gpu_prep.py), which closes the chain reference Python -> host route
-> CUDA kernel.  Tolerances: everything numpy computes the same way is bit-equal; res_corr is within 1e-13 (a
different summation order of the same sampled, truncated, renormalised Gaussian)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(__file__))
import hostcases  # noqa: E402
from test_cloudy import make_table  # noqa: E402
from fake_spectra_b200 import cloudy, gas_properties, line_data, spec_utils, unitsystem  # noqa: E402


@pytest.fixture(scope="module")
def gold(golden_dir):
    z = np.load(os.path.join(golden_dir, "prep_reference.npz"))
    return {k: z[k] for k in z.files}


def snapshot_of(gold):
    snap = hostcases.snapshot(10, 1)
    for name in ("Density", "InternalEnergy", "ElectronAbundance", "NeutralHydrogenAbundance"):
        snap.fields[name] = np.array(gold[name])
    return snap


@pytest.mark.parametrize("sf", [True, False])
def test_gas_properties_match_the_reference(gold, sf):
    snap = snapshot_of(gold)
    gp = gas_properties.GasProperties(float(gold["redshift"]), snap, float(gold["hubble"]), units=unitsystem.UnitSystem(), sf_neutral=sf)
    tag = "sf" if sf else "nosf"
    assert gp.PhysDensThresh == float(gold["PhysDensThresh"])
    assert float(gp.gray_opac) == float(gold["gray_opac"]) and float(gp.gamma_UVB) == float(gold["gamma_UVB"])
    for got, name in ((gp.get_code_rhoH(0, segment=0), "rhoH_"), (gp.get_temp(0, segment=0), "temp_"),
                      (gp.get_reproc_HI(0, segment=0), "reprocHI_")):
        want = gold[name + tag]
        assert got.dtype == want.dtype and np.array_equal(got, want), name
    if sf:  # the self-shielding branch was exercised and changes the snapshot's values
        assert (gold["reprocHI_sf"] != gold["reprocHI_nosf"]).sum() > 50


def test_no_uvb_coverage_makes_star_forming_gas_neutral(gold, capsys):
    gp = gas_properties.GasProperties(9.0, snapshot_of(gold), float(gold["hubble"]), units=unitsystem.UnitSystem())
    assert not gp.redshift_coverage
    assert np.array_equal(gp.get_reproc_HI(0, segment=0), gold["reprocHI_z9"])


def test_unit_system_matches_the_reference(gold):
    us = unitsystem.UnitSystem()
    got = np.array([us.UnitDensity_in_cgs, us.UnitInternalEnergy_in_cgs, us.hubble(2.5, 0.3), us.absorption_distance(20000., 3.),
                    us.redshift_distance(20000., 3., 0.3), us.rho_crit(0.7), us.light, us.protonmass, us.boltzmann, us.gravcgs,
                    us.h100, us.gamma])
    assert np.array_equal(got, gold["units"])
    assert np.array_equal(us.hubble(np.array([0., 1., 2.5]), 0.3), gold["units_hubble_array"])


@pytest.mark.parametrize("z", [2.4, 0.0, 3.0])
def test_cloudy_lookup_matches_the_reference(gold, z):
    table, _ = make_table(nred=4)
    assert np.array_equal(np.array([table.sum(), np.abs(table).max()]), gold["cloudy_table_checksum"])
    ct = cloudy.CloudyTable(z, table=table)  # redshifts 0, 1, 2, 3 like the reference without zz* directories
    assert np.array_equal(np.array(ct.get_dens_bounds() + ct.get_temp_bounds()), gold["cloudy_bounds"])
    for elem, ion in (("C", 4), ("Mg", 2), ("N", 5), ("H", 1)):
        key = "cloudy_z%g_%s%d" % (z, elem, ion)
        assert np.array_equal(ct._slice(elem, ion), gold[key + "_table"]), key
        got = ct.ion(elem, ion, np.array(gold[key + "_rho"]), gold[key + "_temp"])
        assert np.array_equal(got, gold[key]), key


def test_res_corr_matches_the_reference(gold):
    for dv, fwhm in ((1.0, 8.0), (2.5, 8.0), (10.0, 8.0), (1.0, 0.9)):
        want = gold["res_corr_%g_%g" % (dv, fwhm)]
        got = spec_utils.res_corr(gold["res_flux"], dv, fwhm)
        assert np.max(np.abs(got - want)) < 1e-13, (dv, fwhm)


def test_line_table_matches_the_reference(gold):
    """The shipped table (data/lines_9species.dat) is the reference's: same 171 lines of 28 ions, same values, so that
    get_observer_tau chooses among the same transitions; the built-in fallback carries the same values for its subset."""
    ld = line_data.LineData()
    ref = {(str(e), int(v[0]), int(v[1])): v[2:] for e, v in zip(gold["lines_species"], gold["lines_values"])}
    checked = 0
    for (elem, ion), lines in ld.lines.items():
        for lam, line in lines.items():
            want = ref[(elem, ion, lam)]
            assert (line.lambda_X, line.fosc_X, line.gamma_X) == tuple(want), (elem, ion, lam)
            checked += 1
    assert checked == len(ref) == 171  # the shipped table holds every line the reference holds, no more
    assert np.array_equal(np.array([ld.get_mass(e) for e in ("H", "He", "C", "N", "O", "Ne", "Mg", "Si", "Fe")]), gold["masses"])
    for (elem, ion), entries in line_data._BUILTIN.items():
        for lam, fosc, gam in entries:
            assert tuple(ref[(elem, ion, int(lam))]) == (lam, fosc, gam), (elem, ion, lam)


# ---- the known answers of the reference's own Python tests (fake_spectra/tests/test_spectra.py:92-131) ----------------
def test_reference_kat_rho_crit_and_absorption_distance():
    units = unitsystem.UnitSystem()
    assert units.rho_crit(0.7) == 9.204285430050004e-30          # testRhoCrit
    assert units.rho_crit(1.0) == 1.8784255979693885e-29
    assert units.absorption_distance(25000, 3) == 0.13377926628219666   # testAbsDist
    assert units.absorption_distance(25000, 2) == 0.07525083728373562
    assert units.absorption_distance(25000, 3) / units.absorption_distance(12500, 3) == 2.


def test_reference_kat_rolled_spectra():
    tau = np.zeros((2, 50))                                       # testRolledSpectra
    tau[0, 0] = 1
    tau[1, 0] = 1
    tau[1, -1] = 2
    roll, tau_new = spec_utils.get_rolled_spectra(tau)
    assert np.all(roll == np.array([25, -24]))
    assert tau_new[0, 25] == 1 and tau_new[1, 25] == 2 and tau_new[1, 26] == 1
    assert np.sum(np.abs(tau_new)) == 4


def test_reference_kat_res_corr():
    tau = np.zeros((2, 50))                                       # testrescorr
    tau[0, 25] = 2
    tau[1, 23] = 3
    tau2 = spec_utils.res_corr(tau, 2, 8)
    assert abs(np.sum(tau2[0, :]) / np.sum(tau[0, :]) - 1) < 1e-6 and abs(np.sum(tau2[1, :]) / np.sum(tau[1, :]) - 1) < 1e-6
    for i in (0, 1):
        assert np.size(np.where(tau2[i, :] > 0)) == 15
