"""Per-particle gas properties the host classes feed to the native path: hydrogen number density,
temperature and neutral fraction.  Host-side O(N) numpy that prepares the inputs of
``_Particle_Interpolate``; same interface and formulae as the reference's
``gas_properties.GasProperties`` (gas_properties.py:17-200; Rahmati et al. 2013 self-shielding,
Springel & Hernquist 2003 star-formation threshold)."""
import numpy as np

from . import unitsystem

# FG09 UVB tabulated by Rahmati et al. 2013: grey opacity (cm^2) and photoionisation rate (1/s), z = 0..8
_ZZ = np.arange(9.0)
_GRAY_OPAC = np.array([2.59e-18, 2.37e-18, 2.27e-18, 2.15e-18, 2.02e-18, 1.94e-18, 1.82e-18, 1.71e-18, 1.60e-18])
_GAMMA_UVB = np.array([3.99e-14, 3.03e-13, 6e-13, 5.53e-13, 4.31e-13, 3.52e-13, 2.678e-13, 1.81e-13, 9.43e-14])


class GasProperties:
    """redshift, absnap (snapshot object), hubble, fbar, units, sf_neutral: as the reference."""

    def __init__(self, redshift, absnap, hubble=0.71, fbar=0.17, units=None, sf_neutral=True):
        self.units = units if units is not None else unitsystem.UnitSystem()
        self.absnap = absnap
        self.f_bar = fbar
        self.redshift = redshift
        self.sf_neutral = sf_neutral
        self.redshift_coverage = redshift <= _ZZ[-1]
        if self.redshift_coverage:
            # numpy float64 scalars on purpose (the reference holds 0-d float64 arrays from interp1d, gas_properties.py:
            # 51-54): they promote the float32 densities, so the Rahmati formulae below run in double and the neutral
            # fraction is rounded to float32 once, when it is stored
            self.gray_opac = np.float64(np.interp(redshift, _ZZ, _GRAY_OPAC))
            self.gamma_UVB = np.float64(np.interp(redshift, _ZZ, _GAMMA_UVB))
        else:
            print("Warning: no self-shielding at z=", redshift)
        self.gamma = 5. / 3
        self.boltzmann = 1.38066e-16
        self.hubble = hubble
        self.PhysDensThresh = self._get_rho_thresh(hubble)

    # -- Rahmati et al. 2013 ------------------------------------------------------------------------
    def _self_shield_dens(self, temp):
        """Critical self-shielding density, eq. 13 (H atoms / cm^3)."""
        T4 = temp / 1e4
        G12 = self.gamma_UVB / 1e-12
        return 6.73e-3 * (self.gray_opac / 2.49e-18) ** (-2. / 3) * T4 ** 0.17 * G12 ** (2. / 3) * (self.f_bar / 0.17) ** (-1. / 3)

    def _photo_rate(self, nH, temp):
        """Density-dependent photoionisation rate, eq. 14."""
        ratio = nH / self._self_shield_dens(temp)
        return (0.98 * (1 + ratio ** 1.64) ** -2.28 + 0.02 * (1 + ratio) ** -0.84) * self.gamma_UVB

    @staticmethod
    def _recomb_rate(temp):
        """Case-A recombination rate (Hui & Gnedin 1997), cm^3/s."""
        lamb = 315614. / temp
        return 1.269e-13 * lamb ** 1.503 / (1 + (lamb / 0.522) ** 0.47) ** 1.923

    def _neutral_fraction(self, nH, temp):
        """Equilibrium neutral fraction, eq. A8."""
        alpha_A = self._recomb_rate(temp)
        lambda_T = 1.17e-10 * temp ** 0.5 * np.exp(-157809. / temp) / (1 + np.sqrt(temp / 1e5))
        A = alpha_A + lambda_T
        B = 2 * alpha_A + self._photo_rate(nH, temp) / nH + lambda_T
        return (B - np.sqrt(B ** 2 - 4 * A * alpha_A)) / (2 * A)

    # -- interface used by Spectra ------------------------------------------------------------------
    def get_temp(self, part_type, segment):
        """Temperature in K from the internal energy."""
        return self.absnap.get_temp(part_type, segment=segment, units=self.units)

    def _density_conversion(self):
        return np.float32(self.units.UnitDensity_in_cgs * self.hubble ** 2 / self.units.protonmass * (1 + self.redshift) ** 3)

    def get_code_rhoH(self, part_type, segment):
        """Physical H atoms / cm^3 from the code density (h^2 1e10 Msun / kpc^3 comoving)."""
        return self.absnap.get_data(part_type, "Density", segment=segment) * self._density_conversion()

    def get_reproc_HI(self, part_type, segment):
        """Neutral hydrogen fraction: the snapshot's value, replaced above the star-formation
        threshold by the self-shielded equilibrium value at 1e4 K (gas_properties.py:116-146)."""
        nH0 = self.absnap.get_data(part_type, "NeutralHydrogenFraction", segment=segment)
        if not self.sf_neutral:
            return nH0
        density = self.absnap.get_data(part_type, "Density", segment=segment)
        conv = self._density_conversion()
        ind = np.where(density > self.PhysDensThresh / 0.76 / conv)
        if self.redshift_coverage:
            nH0[ind] = self._neutral_fraction(density[ind] * conv, 1e4)
        else:
            nH0[ind] = 1.
        return nH0

    def _get_rho_thresh(self, hubble=0.7, t_0_star=2.27, T_SN=5.73e7, T_c=1000, A_0=573):
        """Star-formation density threshold of the Springel-Hernquist two-phase model, H atoms/cm^3."""
        unit_time = self.units.UnitLength_in_cm / self.units.UnitVelocity_in_cm_per_s
        hy_mass = 0.76
        t_0_star = t_0_star * unit_time / hubble
        beta = 0.264089
        kb_mp = self.boltzmann / self.units.protonmass / (self.gamma - 1)
        u_c = (1 + 3 * hy_mass) / 4 * kb_mp * T_c                 # neutral gas
        ionised = (8 - 5 * (1 - hy_mass)) / 4                     # full ionisation
        u_SN = ionised * kb_mp * T_SN
        rhoinf = 277.476 * self.units.UnitDensity_in_cgs * self.hubble ** 2
        tcool = 4.64419e-10 * unit_time / self.hubble
        u_h = u_SN / A_0
        u_4 = ionised * kb_mp * 1e4
        coolrate = u_h / tcool / rhoinf
        x = (u_h - u_4) / (u_h - u_c)
        physdens = x / (1 - x) ** 2 * (beta * u_SN - (1 - beta) * u_c) / (t_0_star * coolrate)
        return physdens / self.units.protonmass * hy_mass
