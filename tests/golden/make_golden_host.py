"""Generates tests/golden/host_reference.npz: outputs of the reference's own PYTHON host classes (spectra.Spectra,
randspectra.RandSpectra, griddedspectra.GriddedSpectra, fluxstatistics) running on top of the reference's own C++.

The reference package cannot be imported as it stands in this container (h5py is absent and its CPython extension
_spectra_priv is not built), so the generator assembles it from its unmodified parts:
  * a stub package object whose __path__ is /root/reference/fake_spectra, so that the submodules are imported from
    where they lie without the package __init__;
  * an empty stand-in for h5py (only the file I/O of the classes touches it, which is not exercised);
  * a stand-in for fake_spectra._spectra_priv whose _Particle_Interpolate / _near_lines call the UNMODIFIED reference
    C++ (oracle/_ref/libfsref.so = absorption.cpp, index_table.cpp, part_int.cpp, Faddeeva.cpp compiled where they lie)
    through this repository's ctypes shim instead of py_module.cpp's argument parsing; _rescale_mean_flux (py_module.cpp:
    235-282, not part of that library) comes from the C restatement, which tests/test_oracle_stats.py pins to the
    reference's own known answers;
  * abstractsnapshot.AbstractSnapshotFactory replaced by the identity, so that the in-memory synthetic snapshot (which
    has the AbstractSnapshot duck-type) is used instead of a file.
Everything else -- sightline selection, unit handling, particle preparation, segment accumulation, weighted fields,
observer tau, damped-absorber masking, flux statistics -- is the reference's code, run unmodified.

    python tests/golden/make_golden_host.py
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


CONVERTED = []


def reference_package():
    from oracle import Oracle, Reference
    ref, orc = Reference(), Oracle()
    pkg = types.ModuleType("fake_spectra")
    pkg.__path__ = [REF]
    sys.modules["fake_spectra"] = pkg
    sub = types.ModuleType("fake_spectra.cloudy_tables")
    sub.__path__ = [os.path.join(REF, "cloudy_tables")]
    sys.modules["fake_spectra.cloudy_tables"] = sub
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    priv = types.ModuleType("fake_spectra._spectra_priv")

    def particle_interpolate(compute_tau, nbins, kernel, box, velfac, atime, lam, gamma, fosc, amumass, tautail, pos, vel, dens,
                             temp, h, axis, cofm):
        # py_module.cpp:122-125 insists on float32.  Under NumPy >= 2 the reference's weighted fields arrive as float64
        # (spectra.py:954: a float32 array times the numpy float64 scalar sqrt(atime) is promoted; NumPy 1 kept float32
        # there), which the real extension would refuse: they are rounded to float32 here, once, and counted.
        if dens.dtype != np.float32:
            CONVERTED.append(str(dens.dtype))
            dens = dens.astype(np.float32)
        for a in (pos, vel, dens, temp, h):
            assert a.dtype == np.float32
        if compute_tau:
            return ref.compute_tau(nbins, kernel, box, velfac, atime, lam, gamma, fosc, amumass, tautail, pos, vel, dens, temp, h,
                                   axis, cofm)
        return ref.compute_colden(nbins, kernel, box, velfac, atime, lam, gamma, fosc, amumass, tautail, pos, dens, h, axis, cofm)

    priv._Particle_Interpolate = particle_interpolate
    priv._near_lines = lambda box, pos, hh, axis, cofm: ref.near_lines(box, pos, hh, axis, cofm)
    priv._rescale_mean_flux = lambda tau, mf, n, tol, thresh: orc.mean_flux_scale(np.ravel(tau)[:int(n)], mf, tol, thresh)
    sys.modules["fake_spectra._spectra_priv"] = priv
    absn = importlib.import_module("fake_spectra.abstractsnapshot")
    absn.AbstractSnapshotFactory = lambda num, base, comm=None: base
    return {n: importlib.import_module("fake_spectra." + n) for n in ("spectra", "randspectra", "griddedspectra", "fluxstatistics")}


def main():
    import hostcases
    from test_cloudy import make_table
    mods = reference_package()
    out = {}
    tmp = tempfile.mkdtemp()
    table, _ = make_table(nred=4)
    np.savez(os.path.join(tmp, "cloudy_table.npz"), table=table)
    common = dict(quiet=True, savefile="unused.hdf5", savedir=tmp, cdir=tmp + "/")

    # ---- RandSpectra on two segments: H I, the weighted fields, the statistics, a metal ion through the Cloudy table
    rs = mods["randspectra"].RandSpectra(0, hostcases.snapshot(12, 2), numlos=20, thresh=0., res=1.5, **common)
    out["rand_cofm"], out["rand_axis"], out["rand_nbins"] = rs.cofm, rs.axis, rs.nbins
    out["rand_header"] = np.array([rs.box, rs.red, rs.hubble, rs.velfac, rs.vmax, rs.dvbin, rs.rscale, rs.Hz, rs.OmegaM, rs.omegab])
    out["rand_tau_1215"] = np.array(rs.get_tau("H", 1, 1215))
    out["rand_tau_1025"] = np.array(rs.get_tau("H", 1, 1025))
    out["rand_colden_H1"] = np.array(rs.get_col_density("H", 1))
    out["rand_colden_Hall"] = np.array(rs.get_col_density("H", -1))
    out["rand_temp_H1"] = np.array(rs.get_temp("H", 1))
    out["rand_vel_H1"] = np.array(rs.get_velocity("H", 1))
    out["rand_dwd_H1"] = np.array(rs.get_dens_weighted_density("H", 1))
    out["rand_tau_C4_1548"] = np.array(rs.get_tau("C", 4, 1548))
    out["rand_colden_C4"] = np.array(rs.get_col_density("C", 4))
    out["rand_colden_Z"] = np.array(rs.get_col_density("Z", -1))
    # (get_observer_tau is left out: on these sightlines the reference itself fails in it -- spectra.py:939 indexes with an
    # empty selection when a sightline's smoothed maxima tie differently before and after the NaN-free comparison)
    out["rand_eq_width"] = np.array(rs.equivalent_width("H", 1, 1215))
    # absorber statistics (thresholds scaled to these small spectra: total columns are 3e13 - 3e14 cm^-2)
    out["rand_eq_width_hist"] = np.concatenate(rs.eq_width_hist("H", 1, 1215, dv=0.1))
    out["rand_line_density_eq_w"] = rs.line_density_eq_w(thresh=0.2, elem="H", ion=1, line=1215)
    out["rand_cddf_line"] = np.concatenate(rs.column_density_function("H", 1, dlogN=0.25, minN=12.5, maxN=15.))
    out["rand_cddf_pixels"] = np.concatenate(rs.column_density_function("H", 1, dlogN=0.25, minN=10., maxN=14., line=False, close=12.))
    out["rand_cddf_dz"] = np.concatenate(rs.column_density_function("H", 1, dlogN=0.25, minN=12.5, maxN=15., dX=False))
    out["rand_omega_abs"] = rs.omega_abs(thresh=5e13, upthresh=1e40)
    out["rand_omega_abs_all"] = rs.omega_abs(thresh=0, upthresh=1e40, elem="C", ion=4)
    out["rand_omega_abs_cddf"] = rs.omega_abs_cddf(thresh=1e13, upthresh=1e15)
    out["rand_line_density"] = rs.line_density(thresh=5e13, upthresh=2e14)
    out["rand_metallicity"] = rs.get_metallicity()
    out["rand_metallicity_w20"] = rs.get_metallicity(width=20.)
    out["rand_ion_metallicity_C4"] = rs.get_ion_metallicity("C", 4)
    out["rand_density_H1"] = rs.get_density("H", 1)
    flux = np.exp(-np.array(out["rand_tau_1215"]))
    snr, cerr = np.linspace(5., 50., rs.NumLos), np.linspace(0.02, 0.3, rs.NumLos)
    out["rand_noise_snr"], out["rand_cont_err"] = snr, cerr
    noisy, noise = rs.add_noise(snr, np.array(flux))
    out["rand_noisy_flux"], out["rand_noise"] = noisy, noise
    out["rand_noisy_single"] = rs.add_noise(snr, np.array(flux[3]), spec_num=3)[0]
    cflux, delta = rs.add_cont_error(cerr, np.array(flux))
    out["rand_cont_flux"], out["rand_cont_delta"] = cflux, delta
    one, d1 = rs.add_cont_error(cerr, np.array(flux[5]), spec_num=5)
    out["rand_cont_single"], out["rand_cont_single_delta"] = one, d1
    out["rand_mean_flux"] = rs.get_mean_flux()
    out["rand_flux_pdf"] = np.array(rs.get_flux_pdf(nbins=10)[1])
    out["rand_flux_pdf_rescaled"] = np.array(rs.get_flux_pdf(nbins=10, mean_flux_desired=0.6)[1])
    kf, pk = rs.get_flux_power_1D()
    out["rand_flux_power_k"], out["rand_flux_power"] = kf, pk
    out["rand_flux_power_rescaled"] = np.array(rs.get_flux_power_1D(mean_flux_desired=0.6)[1])
    # damped-absorber masking: a threshold low enough that the thickest sightlines are edited (on a copy of tau:
    # the filter works in place and the cached array is what the calls above returned)
    tau = np.array(out["rand_tau_1215"])
    thresh = 0.35 * tau.max()
    out["rand_tau_thresh"] = thresh
    out["rand_tau_filtered"] = rs._filter_tau(np.array(tau), tau_thresh=thresh)
    rs.tau[("H", 1, 1215)] = np.array(tau)
    out["rand_mean_flux_thresh"] = rs.get_mean_flux(tau_thresh=thresh)

    # ---- observer tau: needs a spectrograph resolution (res_corr with sigma = 0 yields NaN maxima in the reference)
    ro = mods["randspectra"].RandSpectra(0, hostcases.snapshot(12, 1), numlos=8, thresh=0., res=1.5, spec_res=8., **common)
    out["obs_cofm"] = ro.cofm
    out["obs_tau_H1"] = np.array(ro.get_observer_tau("H", 1))
    out["obs_tau_Si2"] = np.array(ro.get_observer_tau("Si", 2))
    out["obs_tau_C4_number3"] = np.array(ro.get_observer_tau("C", 4, number=3))

    # ---- threshold selection: sightlines below the column density threshold are replaced until ndla are found
    rd = mods["randspectra"].RandSpectra(0, hostcases.snapshot(12, 1), numlos=12, ndla=7, thresh=1.2e14, res=2.0, **common)
    out["dla_cofm"], out["dla_axis"], out["dla_discarded"], out["dla_numlos"] = rd.cofm, rd.axis, rd.discarded, rd.NumLos
    out["dla_colden_H1"] = np.array(rd.get_col_density("H", 1))
    out["dla_tau_1215"] = np.array(rd.get_tau("H", 1, 1215))

    # ---- no self-shielding correction / no damping wings
    rs2 = mods["randspectra"].RandSpectra(0, hostcases.snapshot(12, 1), numlos=12, thresh=0., res=2.0, sf_neutral=False,
                                          turn_off_selfshield=True, **common)
    out["noss_cofm"], out["noss_axis"] = rs2.cofm, rs2.axis
    out["noss_tau_1215"] = np.array(rs2.get_tau("H", 1, 1215))

    # ---- GriddedSpectra on all three axes, top-hat kernel on an Arepo-like snapshot
    gs = mods["griddedspectra"].GriddedSpectra(0, hostcases.snapshot(10, 1, arepo=True), nspec=4, res=2.5, axis=-1, **common)
    out["grid_cofm"], out["grid_axis"], out["grid_nbins"], out["grid_kernel"] = gs.cofm, gs.axis, gs.nbins, gs.kernel_int
    out["grid_tau_1215"] = np.array(gs.get_tau("H", 1, 1215))
    out["grid_colden_H1"] = np.array(gs.get_col_density("H", 1))
    gs1 = mods["griddedspectra"].GriddedSpectra(0, hostcases.snapshot(10, 1), nspec=5, res=2.5, axis=2, **common)
    out["grid1_cofm"], out["grid1_axis"] = gs1.cofm, gs1.axis
    out["grid1_tau_1215"] = np.array(gs1.get_tau("H", 1, 1215))
    out["grid1_proj_pos"] = gs1.get_spectra_proj_pos()

    # ---- plain Spectra with explicit sightlines and the quintic kernel
    cofm = np.array([[100., 200., 300.], [1500., 40., 900.], [700., 700., 700.]])
    axis = np.array([1, 2, 3])
    sp = mods["spectra"].Spectra(0, hostcases.snapshot(12, 1), cofm, axis, res=1.0, kernel="quintic", reload_file=True, **common)
    out["plain_nbins"], out["plain_kernel"] = sp.nbins, sp.kernel_int
    out["plain_tau_1215"] = np.array(sp.get_tau("H", 1, 1215))
    out["plain_colden_He"] = np.array(sp.get_col_density("He", -1))
    out["float64_weight_arrays_rounded"] = len(CONVERTED)
    np.savez_compressed(os.path.join(HERE, "host_reference.npz"), **out)
    print("wrote host_reference.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
