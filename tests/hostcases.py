"""Helpers for the host-class tests: an oracle-backed stand-in for the native boundary module (the
product never imports the oracle; the CPU tests of the host logic inject this checker), and small
synthetic snapshots."""
import numpy as np

from fake_spectra_b200 import synthetic as syn


class OracleBackend:
    """Same two entry points as fake_spectra_b200._spectra_priv, computed by the CPU oracle."""

    def __init__(self, oracle):
        self.oracle = oracle
        self.calls = []

    def _Particle_Interpolate(self, compute_tau, nbins, kernel, box, velfac, atime, lambda_cm, gamma, fosc, amumass, tautail,
                              pos, vel, dens, temp, h, axis, cofm):
        for a in (pos, vel, dens, temp, h):
            if a.dtype != np.float32:
                raise TypeError("One of the data arrays does not have 32-bit float type")
        self.calls.append(("tau" if compute_tau else "colden", pos.shape[0], cofm.shape[0]))
        if compute_tau:
            return self.oracle.compute_tau(nbins, kernel, box, velfac, atime, lambda_cm, gamma, fosc, amumass, tautail, pos, vel,
                                           dens, temp, h, axis, cofm)
        return self.oracle.compute_colden(nbins, kernel, box, velfac, atime, lambda_cm, gamma, fosc, amumass, tautail, pos,
                                          dens, h, axis, cofm)

    def _near_lines(self, box, pos, hh, axis, cofm):
        return self.oracle.near_lines(box, pos, hh, axis, cofm)

    def _count_pairs(self, box, pos, hh, axis, cofm):
        off, _, _ = self.oracle.near_particles(cofm, axis, box, pos, hh)
        return np.diff(off).astype(np.int32)


def snapshot(nside=10, nsegments=1, seed=4, arepo=False):
    return syn.SyntheticSnapshot(nside, seed=seed, nsegments=nsegments, arepo=arepo)
