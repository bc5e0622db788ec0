"""Units and constants read by the host classes.

The attribute and method names are the ones the reference's spectra driver uses on its
``unitsystem.UnitSystem`` (unitsystem.py:5-62), so code written against either keeps working; the values
of the constants are the reference's own (including its rounded speed of light, unitsystem.py:20), because
they enter results that the parity tests compare.
"""
import math

import numpy as np

# cgs constants, as used by the reference
_CONSTANTS = {
    "light": 2.99e10,             # cm/s
    "protonmass": 1.67262178e-24, # g
    "boltzmann": 1.38066e-16,     # erg/K
    "gravcgs": 6.674e-8,          # cm^3/g/s^2
    "h100": 3.2407789e-18,        # 100 km/s/Mpc in 1/s
    "gamma": 5. / 3,              # adiabatic index of the gas
}


class UnitSystem:
    """Gadget internal units (default: kpc/h, 1e10 Msun/h, km/s) with the derived cgs conversion factors."""

    def __init__(self, UnitMass_in_g=1.98892e43, UnitLength_in_cm=3.085678e21, UnitVelocity_in_cm_per_s=1e5):
        for name, value in _CONSTANTS.items():
            setattr(self, name, value)
        self.UnitMass_in_g, self.UnitLength_in_cm = UnitMass_in_g, UnitLength_in_cm
        self.UnitVelocity_in_cm_per_s = UnitVelocity_in_cm_per_s

    @property
    def UnitDensity_in_cgs(self):
        return self.UnitMass_in_g / self.UnitLength_in_cm ** 3

    @property
    def UnitInternalEnergy_in_cgs(self):
        return self.UnitVelocity_in_cm_per_s ** 2

    def _per_c_times_length(self, rate, speclen):
        """rate [1/s] x comoving length [internal units] / c, multiplied in the reference's order (unitsystem.py:42,51) so
        that the values agree to the last bit."""
        return rate / self.light * speclen * self.UnitLength_in_cm

    def hubble(self, z, omegam0):
        """H(z) in h/s for a flat universe with matter density omegam0."""
        return self.h100 * np.sqrt(omegam0 * (1 + z) ** 3 + (1 - omegam0))

    def absorption_distance(self, speclen, red):
        """Absorption distance X of one sightline of comoving length speclen: (1+z)^2 H0 dL / c."""
        return self._per_c_times_length(self.h100, speclen) * (1 + red) ** 2

    def redshift_distance(self, speclen, red, omegam0):
        """Redshift interval spanned by the comoving length speclen: H(z) dL / c."""
        return self._per_c_times_length(self.hubble(red, omegam0), speclen)

    def rho_crit(self, hubble):
        """Critical density today in g/cm^3 for H0 = 100 hubble km/s/Mpc."""
        return 3 * (self.h100 * hubble) ** 2 / (8 * math.pi * self.gravcgs)
