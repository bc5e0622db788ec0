"""Cloudy ionisation tables: the host side of the metal-ion lookup.

The reference's ``cloudy_tables.convert_cloudy.CloudyTable`` (convert_cloudy.py:123-200) holds
log10(ion fraction) on a regular grid ``[redshift, log10 nH, log10 T, species, ion]`` (the 121 MB of
Cloudy output and its ``cloudy_table.npz`` cache are data of the reference package and are not
shipped here), interpolates it linearly to the snapshot's redshift and looks particles up with
``scipy.ndimage.map_coordinates`` (cubic B-spline, mode "nearest").

Here the table for one (element, ion) is filtered once on the host (scipy's own spline prefilter, on
the table padded by 12 cells of edge values exactly like map_coordinates pads it internally) and kept in HBM as
``fsb_ion_table``; the per-particle lookup runs inside ``fsb_prepare_particles`` (csrc/fsb_prep.cu).
:meth:`CloudyTable.ion` is the reference's host formula, kept for the host-prepared route of
``Spectra._read_particle_data`` and as the checker of the device lookup in the tests.
"""
import os

import numpy as np

NIONS = 17            # convert_cloudy.py:11
RHO_FACTOR = 0.774132  # gas density -> Cloudy hden (convert_cloudy.py:176-183)
SPLINE_PAD = 12       # what scipy.ndimage pads with before filtering in mode "nearest"
SPECIES = ("H", "He", "C", "N", "O", "Ne", "Mg", "Si", "Fe")


def default_directory():
    """Where the tables are looked for when no directory is given: $FAKE_SPECTRA_CLOUDY_DIR."""
    return os.environ.get("FAKE_SPECTRA_CLOUDY_DIR")


class CloudyTable:
    """Ion fractions at one redshift.

    ``CloudyTable(redshift, directory)`` reads ``directory/cloudy_table.npz`` (key ``table``, the
    reference's cache file, convert_cloudy.py:134-139,202-207) and an optional ``redshifts`` key
    (default 0, 1, 2 ... like the reference when the ``zz*`` directories are absent);
    ``CloudyTable(redshift, table=..., reds=..., dens=..., temp=...)`` takes arrays."""

    species = SPECIES
    solar = {"H": 1, "He": 0.1, "C": 3.55e-4, "N": 9.33e-5, "O": 7.41e-4, "Ne": 1.17e-4, "Mg": 3.8e-5, "Si": 3.55e-5,
             "Fe": 3.24e-5}  # convert_cloudy.py:133

    def __init__(self, redshift, directory=None, table=None, reds=None, dens=None, temp=None):
        self.dens = np.arange(-7, 4, 0.2) if dens is None else np.asarray(dens, dtype=np.float64)   # convert_cloudy.py:128
        self.temp = np.arange(3, 8.6, 0.05) if temp is None else np.asarray(temp, dtype=np.float64)  # :129
        if table is None:
            directory = directory if directory is not None else default_directory()
            if directory is None:
                raise IOError("no Cloudy table: pass cdir= (a directory holding the reference's cloudy_table.npz), set "
                              "FAKE_SPECTRA_CLOUDY_DIR, or give Spectra a cloudy_table object")
            self.directory = directory
            with np.load(os.path.join(directory, "cloudy_table.npz")) as f:
                table = f["table"]
                if reds is None and "redshifts" in f.files:
                    reds = f["redshifts"]
        self.table = np.asarray(table, dtype=np.float64)
        if self.table.ndim != 5 or self.table.shape[1:3] != (self.dens.size, self.temp.size):
            raise ValueError("table must be [redshift, %d densities, %d temperatures, species, ion]"
                             % (self.dens.size, self.temp.size))
        nred = self.table.shape[0]
        self.reds = np.arange(0, nred) if reds is None else np.asarray(reds, dtype=np.float64)[:nred]
        if self.reds.size != nred:
            raise ValueError("one redshift per table slice is needed")
        if 4.0 < redshift < 4.1:  # convert_cloudy.py:150-151
            redshift = 4.0
        self.redshift = redshift
        self.red_table = self._at_redshift(redshift)
        self._device = {}

    def _at_redshift(self, redshift):
        """Linear interpolation along the redshift axis (scipy interp1d with its default bounds check)."""
        reds = self.reds
        if reds.size == 1:
            if redshift != reds[0]:
                raise ValueError("redshift %g outside the table (%g)" % (redshift, reds[0]))
            return self.table[0]
        order = np.argsort(reds)
        reds, table = reds[order], self.table[order]
        if redshift < reds[0] or redshift > reds[-1]:
            raise ValueError("redshift %g outside the table range [%g, %g]" % (redshift, reds[0], reds[-1]))
        hi = int(np.clip(np.searchsorted(reds, redshift), 1, reds.size - 1))
        lo = hi - 1
        # the two-weight form scipy's interp1d evaluates (exact at the tabulated redshifts)
        span = reds[hi] - reds[lo]
        return ((redshift - reds[lo]) / span) * table[hi] + ((reds[hi] - redshift) / span) * table[lo]

    def get_temp_bounds(self):
        return (10 ** np.min(self.temp), 10 ** np.max(self.temp))

    def get_dens_bounds(self):
        return (10 ** np.min(self.dens), 10 ** np.max(self.dens))

    def get_red_bounds(self):
        return (np.min(self.reds), np.max(self.reds))

    def get_solar(self, species):
        return self.solar[species]

    def _slice(self, species, ion):
        return self.red_table[:, :, self.species.index(species), ion - 1]

    def ion(self, species, ion, rho, temp):
        """Host lookup with the reference's formula (convert_cloudy.py:167-200); ``rho`` is scaled in place like
        there.  Raises ValueError more than 0.2 dex outside the grid."""
        from scipy.ndimage import map_coordinates
        rho *= RHO_FACTOR
        for values, grid, what in ((rho, self.dens, "Density"), (temp, self.temp, "Temperature")):
            if np.log10(np.max(values)) > np.max(grid) + 0.2:
                raise ValueError("%s %s larger than allowed" % (what, np.max(values)))
            if np.log10(np.min(values)) < np.min(grid) - 0.2:
                raise ValueError("%s %s smaller than allowed" % (what, np.min(values)))
        crho = (np.log10(rho) - self.dens[0]) * (np.size(self.dens) - 1) / (self.dens[-1] - self.dens[0])
        ctemp = (np.log10(temp) - self.temp[0]) * (np.size(self.temp) - 1) / (self.temp[-1] - self.temp[0])
        ions = map_coordinates(self._slice(species, ion), np.vstack((crho, ctemp)), mode="nearest")
        return 10 ** ions

    # ---- device side ----------------------------------------------------------------------------------------
    def spline_coefficients(self, species, ion):
        """B-spline coefficients of one (element, ion) table, padded: what map_coordinates builds internally."""
        from scipy.ndimage import spline_filter
        padded = np.pad(np.ascontiguousarray(self._slice(species, ion)), SPLINE_PAD, mode="edge")
        return np.ascontiguousarray(spline_filter(padded, order=3, mode="nearest", output=np.float64))

    def device_table(self, species, ion, device):
        """(``_lib.IonTable`` describing the table in HBM, the tensor that owns the memory)."""
        import torch
        from . import _lib
        key = (species, ion, str(device))
        if key not in self._device:
            coef = torch.from_numpy(self.spline_coefficients(species, ion)).to(device)
            tb = _lib.IonTable()
            tb.coef = coef.data_ptr()
            tb.nd, tb.nt, tb.pad = self.dens.size, self.temp.size, SPLINE_PAD
            tb.dens0, tb.dens_span = self.dens[0], self.dens[-1] - self.dens[0]
            tb.temp0, tb.temp_span = self.temp[0], self.temp[-1] - self.temp[0]
            (tb.dens_lo, tb.dens_hi), (tb.temp_lo, tb.temp_hi) = self.get_dens_bounds(), self.get_temp_bounds()
            tb.rho_factor = RHO_FACTOR
            self._device[key] = (tb, coef)
        return self._device[key]
