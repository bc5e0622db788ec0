"""Unit constants used by the host classes (mirrors the attributes of the reference's
``unitsystem.UnitSystem``, unitsystem.py:5-62, which the spectra driver reads)."""
import math


class UnitSystem:
    """Gadget internal units (kpc/h, 1e10 Msun/h, km/s) and cgs constants."""

    def __init__(self, UnitMass_in_g=1.98892e43, UnitLength_in_cm=3.085678e21, UnitVelocity_in_cm_per_s=1e5):
        self.UnitMass_in_g = UnitMass_in_g
        self.UnitLength_in_cm = UnitLength_in_cm
        self.UnitVelocity_in_cm_per_s = UnitVelocity_in_cm_per_s
        self.UnitDensity_in_cgs = UnitMass_in_g / UnitLength_in_cm ** 3
        self.UnitInternalEnergy_in_cgs = UnitVelocity_in_cm_per_s ** 2
        self.light = 2.99e10            # cm/s (value used by the reference, unitsystem.py:20)
        self.protonmass = 1.67262178e-24
        self.boltzmann = 1.38066e-16
        self.gravcgs = 6.674e-8
        self.h100 = 3.2407789e-18       # 100 km/s/Mpc in 1/s
        self.gamma = 5. / 3

    def absorption_distance(self, speclen, red):
        """X(z) per sightline for a comoving length ``speclen`` in kpc/h (unitsystem.py:32-44)."""
        return self.h100 / self.light * speclen * self.UnitLength_in_cm * (1 + red) ** 2

    def hubble(self, z, omegam0):
        return self.h100 * math.sqrt(omegam0 * (1 + z) ** 3 + (1 - omegam0))

    def redshift_distance(self, speclen, red, omegam0):
        return self.hubble(red, omegam0) / self.light * speclen * self.UnitLength_in_cm

    def rho_crit(self, hubble):
        """Critical density at z=0 in g/cm^3 (unitsystem.py:57-62)."""
        h100 = self.h100 * hubble
        return 3 * h100 ** 2 / (8 * math.pi * self.gravcgs)
