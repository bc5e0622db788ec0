"""Absorber statistics of finished spectra: the consumers of tau and column density that the reference keeps on its
Spectra class (spectra.py:772-799, 833-860, 1063-1206): equivalent widths, column density distribution function, the
mass density and line density of strong absorbers, metallicities.  Post-processing on host arrays (NumLos x nbins
results); none of it touches the particle data or the device.  Mixed into :class:`fake_spectra_b200.spectra.Spectra`.

All of them reduce to three ingredients, which are the helpers below: the column density of every sightline (or of
every group of pixels), the path length of a sightline (absorption distance or redshift interval), and the number of
sightlines that were drawn (kept + discarded by a threshold selection)."""
import numpy as np

_trapezoid = getattr(np, "trapezoid", None) or np.trapz


class AbsorberStatistics:
    """Needs from the host class: get_tau, get_col_density, get_density, units, lines, box, red, hubble, OmegaM, dvbin,
    NumLos, discarded, solar, solarz, cofm, axis."""

    # ---- ingredients ---------------------------------------------------------------------------------------------
    def _sightlines_drawn(self):
        """Sightlines that were tried, including those a selection threshold discarded."""
        return self.NumLos + 1. * self.discarded

    def _path_length(self, dX=True):
        """Absorption distance (dX) or redshift interval of one sightline across the box."""
        if dX:
            return self.units.absorption_distance(self.box, self.red)
        return self.units.redshift_distance(self.box, self.red, self.OmegaM)

    def _sightline_columns(self, elem, ion):
        """Total column density of every sightline, ions / cm^2."""
        return np.sum(self.get_col_density(elem, ion), axis=1)

    # ---- equivalent widths -----------------------------------------------------------------------------------------
    def equivalent_width(self, elem, ion, line):
        """Rest-frame equivalent width of a line on every sightline, in Angstrom: the integral of 1 - exp(-tau) over
        wavelength, one pixel being lambda dv / c wide (spectra.py:833-845)."""
        tau = self.get_tau(elem, ion, line)
        pixel_angstrom = self.dvbin / (self.units.light / 1e5) * line
        return _trapezoid(-np.expm1(-tau), dx=pixel_angstrom, axis=1)

    def eq_width_hist(self, elem, ion, line, dv=0.05):
        """Normalised histogram of log10 equivalent width: (bin centres, density) (spectra.py:847-860)."""
        widths = self.equivalent_width(elem, ion, line)
        logw = np.log10(widths[widths > 0])
        edges = np.arange(logw.min(), logw.max(), dv)
        return (edges[1:] + edges[:-1]) / 2., np.histogram(logw, edges, density=True)[0]

    def line_density_eq_w(self, thresh=0.4, elem="H", ion=1, line=1216):
        """dN/dX of sightlines whose equivalent width exceeds ``thresh`` Angstrom (spectra.py:1186-1199)."""
        widths = self.equivalent_width(elem, ion, line)
        return np.count_nonzero(widths > thresh) / (np.size(widths) + 1. * self.discarded) / self._path_length()

    # ---- column density statistics ---------------------------------------------------------------------------------
    def column_density_function(self, elem="H", ion=1, dlogN=0.2, minN=13, maxN=23., line=True, close=50., dX=True):
        """f(N) = d n / dN dX (or dz): number of absorbers per sightline, column density interval and path length
        (spectra.py:1095-1143).  ``line``: one absorber per sightline (its total column); otherwise the pixels are
        merged in groups of ``close`` km/s and every group counts.  Returns (bin centres, f(N))."""
        edges = 10 ** np.arange(minN, maxN, dlogN)
        centres, widths = (edges[1:] + edges[:-1]) / 2., np.diff(edges)
        if line:
            columns = self._sightline_columns(elem, ion)
        else:
            pixels = self.get_col_density(elem, ion)
            group = max(int(np.round(close / self.dvbin)), 1)
            ngroups = pixels.shape[1] // group  # a remainder of pixels at the end of the sightline is left out
            columns = pixels[:, :ngroups * group].reshape(pixels.shape[0], ngroups, group).sum(axis=2)
        counts = np.histogram(columns, edges)[0]
        return centres, counts / (widths * self._path_length(dX) * self._sightlines_drawn())

    def _rho_abs(self, thresh=10 ** 20.3, upthresh=None, elem="H", ion=1):
        """Comoving mass density (g/cm^3) of the ion in sightlines with thresh < N < upthresh (all of them when no
        threshold is set): mean column x ion mass / (1+z)^2 / box length (spectra.py:1145-1162)."""
        columns = self._sightline_columns(elem, ion)
        if thresh > 0 or upthresh is not None:
            # (the reference compares with upthresh even when it is None, which NumPy refuses: None means no upper bound)
            selected = columns > thresh
            if upthresh is not None:
                selected &= columns < upthresh
            mean_column = columns[selected].sum() / columns.size
        else:
            mean_column = columns.mean()
        mean_column *= columns.size / (columns.size + 1. * self.discarded)
        surface_density = self.lines.get_mass(elem) * self.units.protonmass * mean_column / (1 + self.red) ** 2
        return surface_density / (self.box * self.units.UnitLength_in_cm / self.hubble)

    def rho_DLA(self, thresh=10 ** 20.3):
        """Mass density of neutral hydrogen in damped absorbers, 1e8 Msun / Mpc^3 comoving (spectra.py:1164-1172)."""
        unit = 0.01 * self.units.UnitMass_in_g / self.units.UnitLength_in_cm ** 3
        return self._rho_abs(thresh) / unit

    def omega_abs(self, thresh=10 ** 20.3, upthresh=1e40, elem="H", ion=1):
        """The same density in units of the critical density (spectra.py:1174-1183)."""
        return self._rho_abs(thresh, upthresh, elem=elem, ion=ion) / self.units.rho_crit(self.hubble)

    def omega_abs_cddf(self, thresh=10 ** 20.3, upthresh=1e40, elem="H", ion=1):
        """Omega of the absorbers from the first moment of the column density function (spectra.py:1185-1200)."""
        centres, cddf = self.column_density_function(elem, ion, 0.2, minN=np.log10(thresh), maxN=np.log10(upthresh))
        h0 = self.units.h100 * self.hubble
        prefactor = self.lines.get_mass(elem) * self.units.protonmass / self.units.light * h0 / self.units.rho_crit(self.hubble)
        return prefactor * _trapezoid(cddf * centres, centres)

    def line_density(self, thresh=10 ** 20.3, upthresh=10 ** 40, elem="H", ion=1):
        """dN/dX of sightlines with thresh < N < upthresh (spectra.py:1202-1211)."""
        columns = self._sightline_columns(elem, ion)
        fraction = np.count_nonzero((columns > thresh) & (columns < upthresh)) / columns.size
        fraction *= columns.size / (columns.size + 1. * self.discarded)
        return fraction / self._path_length()

    # ---- metallicities ---------------------------------------------------------------------------------------------
    def get_metallicity(self, width=0.):
        """Metal to hydrogen density ratio of every sightline in solar units; with ``width`` > 0 only the pixels within
        +- width km/s of the strongest hydrogen peak count (spectra.py:772-787)."""
        from . import spec_utils
        metals, hydrogen = self.get_density("Z", -1), self.get_density("H", -1)
        if width > 0:
            roll, hydrogen = spec_utils.get_rolled_spectra(hydrogen)
            metals = np.array([np.roll(row, shift) for row, shift in zip(metals, roll)])
            mid = hydrogen.shape[1] // 2
            half = min(int(width / self.dvbin), mid)
            metals, hydrogen = metals[:, mid - half:mid + half], hydrogen[:, mid - half:mid + half]
        return metals.sum(axis=1) / hydrogen.sum(axis=1) / self.solarz

    def get_ion_metallicity(self, species, ion):
        """Ion to neutral hydrogen density ratio of every sightline over the element's solar abundance (spectra.py:793-799)."""
        return self.get_density(species, ion).sum(axis=1) / self.get_density("H", 1).sum(axis=1) / self.solar[species]

    # ---- observational effects (spectra.py:374-432) ---------------------------------------------------------------
    def add_noise(self, snr, flux, spec_num=-1):
        """Adds Gaussian noise of standard deviation 1 / snr to flux spectra, in place.  Reproducible per spectrum: NumPy's
        global generator is seeded with the spectrum's number (``spec_num`` for a single 1-D spectrum, the row number
        otherwise).  Returns (flux, the noise of all rows concatenated -- empty for a single spectrum)."""
        if np.ndim(flux) == 1:
            np.random.seed(spec_num)
            flux += np.random.normal(0, 1. / snr[spec_num], self.nbins)
            return flux, np.array([])
        drawn = []
        for row in range(np.shape(flux)[0]):
            np.random.seed(row)
            drawn.append(np.random.normal(0, 1. / snr[row], self.nbins))
            flux[row] += drawn[-1]
        return flux, np.concatenate(drawn) if drawn else np.array([])

    def add_cont_error(self, CE, flux, spec_num=-1, u_delta=0.6, l_delta=-0.6):
        """Continuum placement error (eq. 2 of arXiv:2112.03930): every spectrum is divided by 1 + delta, delta drawn from
        a Gaussian of width CE truncated to [l_delta, u_delta]; seeded with twice the spectrum's number so that it differs
        from the noise of add_noise.  In place; returns (flux, delta)."""
        def draw(number, width):
            np.random.seed(2 * number)
            while True:
                delta = np.random.normal(0, width)
                if l_delta <= delta <= u_delta:
                    return delta
        if np.ndim(flux) == 1:
            delta = draw(spec_num, CE[spec_num])
            flux /= (1.0 + delta)
            return flux, delta
        deltas = np.array([draw(row, CE[row]) for row in range(np.shape(flux)[0])])
        flux /= (1.0 + deltas)[:, None]
        return flux, deltas

    # ---- geometry --------------------------------------------------------------------------------------------------
    def get_spectra_proj_pos(self, cofm=None):
        """The two coordinates of every sightline perpendicular to the common axis (spectra.py:1213-1228)."""
        if np.any(self.axis != self.axis[0]):
            raise ValueError("Not all spectra are along the same axis")
        cofm = self.cofm if cofm is None else cofm
        keep = [k for k in range(3) if k != self.axis[0] - 1]
        return cofm[:, keep]
