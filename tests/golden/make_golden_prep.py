"""Generates tests/golden/prep_reference.npz by importing the reference's own PYTHON modules where they lie
(/root/reference/fake_spectra: gas_properties, unitsystem, cloudy_tables.convert_cloudy, spec_utils, line_data) and
running them on seeded synthetic snapshot fields: the quantities of SURVEY 8f row f3 that precede the interpolation
(hydrogen number density, temperature, reprocessed neutral fraction, Cloudy ion fractions) plus res_corr and the
line table.  The reference PACKAGE cannot be imported (its __init__ pulls in h5py, absent here), so a stub package
object with the reference directory as its __path__ is registered and the submodules are imported individually,
unmodified.  Run in the build container:

    python tests/golden/make_golden_prep.py

The fixture stores the inputs next to the outputs, so the tests need neither the reference nor the generator.
"""
import importlib
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = "/root/reference/fake_spectra"


def reference_modules():
    pkg = types.ModuleType("fake_spectra")
    pkg.__path__ = [REF]
    sys.modules["fake_spectra"] = pkg
    sub = types.ModuleType("fake_spectra.cloudy_tables")
    sub.__path__ = [os.path.join(REF, "cloudy_tables")]
    sys.modules["fake_spectra.cloudy_tables"] = sub
    names = ("unitsystem", "gas_properties", "cloudy_tables.convert_cloudy", "spec_utils", "line_data")
    return {n.split(".")[-1]: importlib.import_module("fake_spectra." + n) for n in names}


def main():
    import hostcases
    from test_cloudy import make_table
    ref = reference_modules()
    snap = hostcases.snapshot(10, 1)
    snap.fields["Density"][::7] *= np.float32(3e4)      # above the star-formation threshold: Rahmati branch
    snap.fields["InternalEnergy"][::11] = 0
    out = {k: snap.fields[k] for k in ("Density", "InternalEnergy", "ElectronAbundance", "NeutralHydrogenAbundance")}
    redshift = 1. / snap.get_header_attr("Time") - 1.
    hubble = snap.get_header_attr("HubbleParam")
    out["redshift"], out["hubble"] = redshift, hubble
    # --- gas_properties.GasProperties on the snapshot (duck-typed absnap), default unit system
    for sf in (True, False):
        gp = ref["gas_properties"].GasProperties(redshift, snap, hubble, units=ref["unitsystem"].UnitSystem(), sf_neutral=sf)
        tag = "sf" if sf else "nosf"
        out["rhoH_" + tag] = gp.get_code_rhoH(0, segment=0)
        out["reprocHI_" + tag] = gp.get_reproc_HI(0, segment=0)
        out["temp_" + tag] = gp.get_temp(0, segment=0)
        out["PhysDensThresh"] = gp.PhysDensThresh
        out["gray_opac"], out["gamma_UVB"] = float(gp.gray_opac), float(gp.gamma_UVB)
    # the same at a redshift the UVB table does not cover (star-forming gas fully neutral)
    gp9 = ref["gas_properties"].GasProperties(9.0, snap, hubble, units=ref["unitsystem"].UnitSystem())
    out["reprocHI_z9"] = gp9.get_reproc_HI(0, segment=0)
    # --- unit system
    us = ref["unitsystem"].UnitSystem()
    out["units"] = np.array([us.UnitDensity_in_cgs, us.UnitInternalEnergy_in_cgs, us.hubble(2.5, 0.3), us.absorption_distance(20000., 3.),
                             us.redshift_distance(20000., 3., 0.3), us.rho_crit(0.7), us.light, us.protonmass, us.boltzmann,
                             us.gravcgs, us.h100, us.gamma])
    out["units_hubble_array"] = us.hubble(np.array([0., 1., 2.5]), 0.3)
    # --- convert_cloudy.CloudyTable on a synthetic table in the reference's cache format
    table, _ = make_table(nred=4)  # redshifts 0, 1, 2, 3 (the reference numbers them when there are no zz* directories)
    with tempfile.TemporaryDirectory() as tmp:
        np.savez(os.path.join(tmp, "cloudy_table.npz"), table=table)
        for z in (2.4, 0.0, 3.0):
            ct = ref["convert_cloudy"].CloudyTable(z, tmp + "/")
            rng = np.random.default_rng(int(z * 10))
            rho = (10 ** rng.uniform(-6.9, 3.7, 3000)).astype(np.float32)
            temp = (10 ** rng.uniform(3.0, 8.5, 3000)).astype(np.float32)
            for (elem, ion) in (("C", 4), ("Mg", 2), ("N", 5), ("H", 1)):
                key = "cloudy_z%g_%s%d" % (z, elem, ion)
                out[key + "_rho"], out[key + "_temp"] = rho, temp
                out[key] = ct.ion(elem, ion, np.array(rho), temp)
                out[key + "_table"] = ct.red_table[:, :, ct.species.index(elem), ion - 1]  # the slice the lookup interpolates
            out["cloudy_bounds"] = np.array(ct.get_dens_bounds() + ct.get_temp_bounds())
    out["cloudy_table_checksum"] = np.array([table.sum(), np.abs(table).max()])  # tests rebuild it: test_cloudy.make_table(nred=4)
    # --- spec_utils.res_corr
    rng = np.random.default_rng(2)
    flux = np.exp(-np.exp(rng.normal(-1, 1.2, (6, 300))))
    out["res_flux"] = flux
    for dv, fwhm in ((1.0, 8.0), (2.5, 8.0), (10.0, 8.0), (1.0, 0.9)):
        out["res_corr_%g_%g" % (dv, fwhm)] = ref["spec_utils"].res_corr(flux, dv, fwhm)
    # --- line_data.LineData: every line of the species the host classes use
    ld = ref["line_data"].LineData()
    rows = []
    for (elem, ion), lines in sorted(ld.lines.items()):
        for lam, line in sorted(lines.items()):
            rows.append((elem, ion, lam, line.lambda_X, line.fosc_X, line.gamma_X))
    out["lines_species"] = np.array([r[0] for r in rows])
    out["lines_values"] = np.array([r[1:] for r in rows], dtype=np.float64)
    out["masses"] = np.array([ld.get_mass(e) for e in ("H", "He", "C", "N", "O", "Ne", "Mg", "Si", "Fe")])
    np.savez_compressed(os.path.join(HERE, "prep_reference.npz"), **out)
    print("wrote prep_reference.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
