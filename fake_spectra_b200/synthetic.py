"""Synthetic Gadget-like gas snapshots (no HDF5 in this image; benchmarks and tests run in memory).

The generator is the one defined in SURVEY.md App. F: a log-normal density field on uniformly
random positions with a power-law temperature-density relation, at z = 3.  Two views are offered:

* :func:`boundary_arrays` — the float32 arrays exactly as they cross the native boundary
  (``_Particle_Interpolate`` arguments, reference py_module.cpp:115): ``pos, vel, dens, temp, h``.
* :class:`SyntheticSnapshot` — an in-memory object duck-typing the reference's
  ``AbstractSnapshot`` interface (abstractsnapshot.py:29-154,253-301) so that the host classes in
  :mod:`fake_spectra_b200.spectra` can be driven exactly like the reference's.
"""
import numpy as np

MEAN_SPACING = 156.25          # kpc/h, SURVEY App. F
KPC_IN_CM = 3.085678e21        # reference unitsystem.py:7
MPC_IN_CM = 3.085678e24        # reference spectra.py:216


class Cosmology:
    """z = 3 flat LCDM used by every synthetic config."""

    def __init__(self, atime=0.25, hubble=0.7, omega_m=0.3, omega_l=0.7, omega_b=0.045):
        self.atime = atime
        self.hubble = hubble
        self.omega_m = omega_m
        self.omega_l = omega_l
        self.omega_b = omega_b

    @property
    def rscale(self):
        """cm per (comoving kpc/h), float32 like reference spectra.py:210."""
        return np.float32(KPC_IN_CM * self.atime / self.hubble)

    @property
    def Hz(self):
        return 100.0 * self.hubble * np.sqrt(self.omega_m / self.atime ** 3 + self.omega_l)

    @property
    def velfac(self):
        """km/s per (comoving kpc/h), reference spectra.py:216."""
        return self.rscale * self.Hz / MPC_IN_CM


def _fields(nside, seed, box=None):
    """Primitive per-particle fields: positions, overdensity, temperature, 3 velocity components.
    ``box`` overrides the default side (mean spacing x nside): a shard of a larger box."""
    npart = int(nside) ** 3
    rng = np.random.default_rng(seed)
    if box is None:
        box = MEAN_SPACING * nside
    pos = (rng.random((npart, 3), dtype=np.float64) * box).astype(np.float32)
    # positions must stay inside [0, box): float32 rounding can land exactly on box
    np.minimum(pos, np.nextafter(np.float32(box), np.float32(0)), out=pos)
    sigma = 1.0
    g = rng.standard_normal(npart)
    delta = np.exp(sigma * g - 0.5 * sigma * sigma)
    temp = 1e4 * delta ** 0.6
    vel = (100.0 * rng.standard_normal((npart, 3))).astype(np.float32)
    return box, pos, delta, temp, vel


def boundary_arrays(nside, seed=42, kernel=1, metal_scale=1.0, box=None):
    """Arrays as they cross the native boundary.

    Returns a dict with ``box`` (kpc/h), ``pos`` f32[N,3], ``vel`` f32[N,3] (physical km/s, already
    multiplied by sqrt(a)), ``dens`` f32[N] (ion number density x rscale, reference
    spectra.py:593-615), ``temp`` f32[N] (K), ``h`` f32[N] (kernel support radius, kpc/h).
    ``kernel`` 0 mimics Arepo (h = Volume^(1/3), abstractsnapshot.py:268), otherwise SPH.
    """
    box, pos, delta, temp, vel = _fields(nside, seed, box)
    hh = MEAN_SPACING * delta ** (-1.0 / 3.0)
    if kernel in (0, 2):
        hh = hh  # cell "radius" = Volume^(1/3) = spacing * delta^(-1/3): same scaling
    dens = 4.4e10 * delta ** 2 * (temp / 1e4) ** (-0.7) * metal_scale
    return {
        "box": float(box),
        "pos": pos,
        "vel": vel,
        "dens": dens.astype(np.float32),
        "temp": temp.astype(np.float32),
        "h": hh.astype(np.float32),
    }


def random_sightlines(box, numlos, seed=23, axis=1):
    """RandSpectra-style sightlines (reference randspectra.py:23-24,32-37).  ``axis`` may be an int
    or "cycle" for 1,2,3,1,2,3,... (config 3 of BASELINE.json)."""
    np.random.seed(seed)
    cofm = box * np.random.random_sample((numlos, 3))
    if axis == "cycle":
        ax = (np.arange(numlos) % 3 + 1).astype(np.int32)
    else:
        ax = np.full(numlos, axis, dtype=np.int32)
    return cofm.astype(np.float64), ax


def grid_sightlines(box, nspec, axis=1):
    """GriddedSpectra-style sightlines (reference griddedspectra.py:59-88)."""
    from .griddedspectra import grid_axes_and_cofm
    ax, cofm = grid_axes_and_cofm(box, nspec, axis)
    return cofm.astype(np.float64), ax.astype(np.int32)


class _Header(dict):
    pass


class SyntheticSnapshot:
    """In-memory snapshot with the AbstractSnapshot duck-type (reference abstractsnapshot.py).

    Fields are stored under their Gadget-HDF5 names.  The gas fields are constructed backwards from
    the App. F primitives so that the reference host pipeline (get_code_rhoH x get_reproc_HI x
    rscale, spectra.py:576-615) reproduces an ion density ~ 4.4e10 delta^2 T4^-0.7 for H I.
    """

    def __init__(self, nside, seed=42, nsegments=1, arepo=False, cosmo=None, with_metals=True):
        self.cosmo = cosmo if cosmo is not None else Cosmology()
        c = self.cosmo
        box, pos, delta, temp, vel = _fields(nside, seed)
        npart = pos.shape[0]
        self.nside = nside
        self.arepo = arepo
        self._nseg = int(nsegments)
        self.header = _Header(BoxSize=float(box), Time=c.atime, HubbleParam=c.hubble, Omega0=c.omega_m,
                              OmegaLambda=c.omega_l, OmegaBaryon=c.omega_b,
                              NumPart_Total=np.array([npart, npart, 0, 0, 0, 0]),
                              TotNumPart=npart, UnitLength_in_cm=KPC_IN_CM, UnitMass_in_g=1.98892e43,
                              UnitVelocity_in_cm_per_s=1e5)
        from .unitsystem import UnitSystem
        units = UnitSystem()
        # Density in code units such that the physical hydrogen number density is delta * nH_mean(z).
        protonmass = units.protonmass
        rho_crit = 3 * (units.h100 * c.hubble) ** 2 / (8 * np.pi * units.gravcgs)
        nH_mean = 0.76 * c.omega_b * rho_crit / protonmass / c.atime ** 3        # physical cm^-3
        conv = units.UnitDensity_in_cgs * c.hubble ** 2 / protonmass / c.atime ** 3  # gas_properties.py:108
        density = (delta * nH_mean / conv).astype(np.float32)
        # neutral fraction: target n_HI*rscale = 4.4e10 delta^2 T4^-0.7 with n_HI = 0.76 * nH * x_HI
        target = 4.4e10 * delta ** 2 * (temp / 1e4) ** (-0.7)
        nH = density.astype(np.float64) * conv
        xHI = np.clip(target / (nH * float(c.rscale) * 0.76), 0.0, 1.0)
        # internal energy from T with fully ionised primordial gas (abstractsnapshot.py:121-154)
        nelec = np.full(npart, 1.158, dtype=np.float32)
        hy = 0.76
        mu_fac = 4.0 / (hy * (3 + 4 * nelec.astype(np.float64)) + 1)
        ienergy = temp / ((units.gamma - 1) * protonmass / units.boltzmann * mu_fac) / units.UnitInternalEnergy_in_cgs
        hsml = MEAN_SPACING * delta ** (-1.0 / 3.0)
        self.fields = {
            "Coordinates": pos,
            "Velocities": (vel / np.sqrt(c.atime)).astype(np.float32),  # Gadget comoving convention
            "Density": density,
            "InternalEnergy": ienergy.astype(np.float32),
            "ElectronAbundance": nelec,
            "NeutralHydrogenAbundance": xHI.astype(np.float32),
        }
        if arepo:
            self.fields["Volume"] = (hsml ** 3).astype(np.float32)
            self.fields["Masses"] = (density * hsml ** 3).astype(np.float32)
        else:
            self.fields["SmoothingLength"] = (2.0 * hsml).astype(np.float32)  # reference halves it (:275)
        if with_metals:
            # GFM_Metals columns follow Spectra.species order H He C N O Ne Mg Si Fe (spectra.py:236)
            zrel = 0.1 * delta ** 0.5  # metallicity relative to solar, rising with density
            solar_massfrac = np.array([0.76, 0.24, 2.4e-3, 7e-4, 5.7e-3, 1.2e-3, 7e-4, 6.7e-4, 1.3e-3])
            metals = np.empty((npart, 9), dtype=np.float32)
            metals[:, 0] = 0.76
            metals[:, 1] = 0.24
            for k in range(2, 9):
                metals[:, k] = (zrel * solar_massfrac[k]).astype(np.float32)
            self.fields["GFM_Metals"] = metals
            self.fields["GFM_Metallicity"] = (zrel * 0.0134).astype(np.float32)
        self._alias = {"Position": "Coordinates", "Velocity": "Velocities", "Mass": "Masses",
                       "NeutralHydrogenFraction": "NeutralHydrogenAbundance", "Metallicity": "GFM_Metallicity"}
        self._units = units
        self._npart = npart

    # -- AbstractSnapshot interface -------------------------------------------------------------
    def get_header_attr(self, attr):
        return self.header[attr]

    def get_kernel(self):
        """0 for Arepo-like (Volume present), 1 for Gadget SPH (abstractsnapshot.py:284-301)."""
        return 0 if self.arepo else 1

    def get_n_segments(self, part_type=0):
        return self._nseg

    def _segment_slice(self, segment):
        if segment is None or segment < 0:
            return slice(0, self._npart)
        edges = np.linspace(0, self._npart, self._nseg + 1).astype(np.int64)
        return slice(int(edges[segment]), int(edges[segment + 1]))

    def get_blocklen(self, part_type, blockname, segment):
        sl = self._segment_slice(segment)
        return sl.stop - sl.start

    def get_data(self, part_type, blockname, segment):
        if part_type != 0:
            raise KeyError("SyntheticSnapshot only carries gas (PartType0)")
        name = self._alias.get(blockname, blockname)
        if name not in self.fields:
            raise KeyError(blockname)
        return np.array(self.fields[name][self._segment_slice(segment)])

    def get_npart(self):
        return self.header["NumPart_Total"]

    def get_omega_baryon(self):
        return self.header["OmegaBaryon"]

    def get_smooth_length(self, part_type, segment):
        """Kernel support radius (abstractsnapshot.py:253-282)."""
        if "Volume" in self.fields:
            return np.power(self.get_data(part_type, "Volume", segment=segment), 1. / 3)
        return self.get_data(part_type, "SmoothingLength", segment=segment) / 2

    def get_units(self):
        return self._units

    def get_peculiar_velocity(self, part_type, segment):
        """Gadget velocities times sqrt(a) (abstractsnapshot.py:114-119)."""
        vel = self.get_data(part_type, "Velocities", segment=segment)
        vel *= np.sqrt(self.header["Time"])
        return vel

    def get_temp(self, part_type, segment, hy_mass=0.76, units=None):
        """Temperature in K from internal energy (abstractsnapshot.py:121-154)."""
        if units is None:
            units = self._units
        ienergy = self.get_data(part_type, "InternalEnergy", segment=segment) * units.UnitInternalEnergy_in_cgs
        nelec = self.get_data(part_type, "ElectronAbundance", segment=segment)
        muienergy = 4 / (hy_mass * (3 + 4 * nelec) + 1) * ienergy
        return (units.gamma - 1) * units.protonmass / units.boltzmann * muienergy


def write_bigfile(snapshot, path, num=0, nfile=2, peculiar=False):
    """Export a :class:`SyntheticSnapshot` in MP-Gadget's BigFile layout under ``path/PART_<num>`` (block names and
    velocity convention of MP-Gadget: Velocity = a v_pec unless ``peculiar``): an input for
    :class:`fake_spectra_b200.abstractsnapshot.BigFileSnapshot`.  Returns ``path``."""
    import os

    from .abstractsnapshot import HDF_TO_BIGFILE, write_bigfile_block
    root = os.path.join(path, "PART_" + str(num).rjust(3, "0"))
    h = snapshot.header
    atime = float(h["Time"])
    attrs = {"BoxSize": float(h["BoxSize"]), "Time": atime, "HubbleParam": float(h["HubbleParam"]),
             "Omega0": float(h["Omega0"]), "OmegaLambda": float(h["OmegaLambda"]), "OmegaBaryon": float(h["OmegaBaryon"]),
             "TotNumPart": np.asarray(h["NumPart_Total"], dtype=np.uint64),
             "UnitLength_in_cm": float(h["UnitLength_in_cm"]), "UnitMass_in_g": float(h["UnitMass_in_g"]),
             "UnitVelocity_in_cm_per_s": float(h["UnitVelocity_in_cm_per_s"]),
             "UsePeculiarVelocity": np.int32(1 if peculiar else 0), "DensityKernel": np.int32(1), "CodeVersion": "synthetic"}
    write_bigfile_block(os.path.join(root, "Header"), None, attrs=attrs)
    for name, data in snapshot.fields.items():
        if name in ("Volume",):
            continue  # MP-Gadget output is SPH
        out = data
        if name == "Velocities":  # stored Gadget-style as v_pec / sqrt(a)
            vpec = data.astype(np.float64) * np.sqrt(atime)
            out = (vpec if peculiar else vpec * atime).astype(np.float32)
        write_bigfile_block(os.path.join(root, "0", HDF_TO_BIGFILE.get(name, name)), out, nfile=nfile)
    return path
