"""Host driver of the sightline interpolation: the reference's ``Spectra`` API
(spectra.py:51-1043: constructor, get_tau, get_col_density, get_density, get_temp, get_velocity,
get_dens_weighted_density, compute_spectra, particles_near_lines, _do_interpolation_work) on top of
the sm_100a kernels of libfsb200.so.

Two routes lead to the same kernels:

* drop-in route — ``_do_interpolation_work`` hands host arrays to
  ``_spectra_priv._Particle_Interpolate`` exactly like the reference does (spectra.py:673);
* resident route (default on a GPU) — per snapshot segment and ion the filtered particle arrays and
  the candidate index stay in HBM (:class:`_SegmentEngine`), so that further lines of the ion, the
  column density and the weighted fields reuse them; the reference rebuilds its index twice per call
  (spectra.py:560-563 + part_int.cpp:22,55).

With ``torch.distributed`` initialised (one process per GPU) the work is sharded by sightline or by
particle (:mod:`fake_spectra_b200.sharding`).  There is no CPU fallback: without the CUDA library
the native calls raise.
"""
import os.path as path

import numpy as np

from . import _spectra_priv
from . import abstractsnapshot as absn
from .absorbers import AbsorberStatistics
from . import cloudy
from . import gas_properties
from . import line_data
from . import sharding
from . import unitsystem

_KERNELS = {"voronoi": 2, "tophat": 0, "quintic": 3, "cubic": 1, "sph": 1}
_PRECISION = {"fp64": 0, "fp32": 1}
_VOIGT = {"fast": 0, "exact": 1}


class _SegmentEngine:
    """Device-resident particles of one (segment, element, ion) plus their candidate index."""

    def __init__(self, spec, pos, vel, elem_den, temp, hh):
        import torch
        from . import native
        self.torch, self.native = torch, native
        dev = torch.device("cuda", torch.cuda.current_device())
        self.dev = dev
        # host arrays (the reference's _read_particle_data) are uploaded; device tensors (_device_particle_data) are taken as is
        up = lambda a: a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)  # noqa: E731
        self.pos, self.vel, self.dens, self.temp, self.h = up(pos), up(vel), up(elem_den), up(temp), up(hh)
        self.cofm = torch.from_numpy(np.ascontiguousarray(spec._my_cofm)).to(dev)
        self.axis = torch.from_numpy(np.ascontiguousarray(spec._my_axis)).to(dev)
        # one candidate index, or one per sightline block when the pairs exceed what an index holds (2^31)
        self.index = native.BlockedIndex(spec.box, self.cofm, self.axis, self.pos, self.h)
        self.cells = None

    def tau(self, params_list, out=None, push=None):
        """float64 CUDA tensor [nlines, nlos_local, nbins]; accumulated into ``out`` when given.  ``push``: see
        native.PeerRows (rows finished by this call are also stored into every rank's full array)."""
        if params_list[0].kernel == 2:
            res = self.torch.stack([self.native.particle_interpolate(1, p, self.pos, self.vel, self.dens, self.temp, self.h,
                                                                     self.axis, self.cofm) for p in params_list])
            if out is not None:
                out += res
                return out
            return res
        return self.index.compute_tau(list(params_list), self.pos, self.vel, self.dens, self.temp, self.h, out=out, push=push)

    def release(self):
        """Free the candidate index now (the tensors go back to torch's allocator with the object)."""
        if self.index is not None:
            self.index.free()
            self.index = None

    def colden(self, params, weights=None):
        """Column density of the ion density (weights None) or of K weight columns
        [K, npart] float32 CUDA in one geometry pass: float64 [K, nlos_local, nbins]."""
        dens = self.dens[None, :] if weights is None else weights
        dens = dens.contiguous()
        if params.kernel == 2:
            return self.torch.stack([self.native.particle_interpolate(0, params, self.pos, None, d.contiguous(), None, self.h,
                                                                      self.axis, self.cofm) for d in dens])
        return self.index.compute_colden(params, self.pos, dens, self.h)


class Spectra(AbsorberStatistics):
    """Interpolates particle densities along sightlines and computes their absorption.

    Positional and keyword arguments are those of the reference (spectra.py:85-87); ``base`` may be
    an in-memory snapshot object (see :mod:`abstractsnapshot`).  Extensions, keyword only:

    precision  "fp64" (parity mode, <= 1e-10 of the reference) or "fp32" (flux within 1e-5)
    voigt      "fast" (this library's profile evaluation) or "exact" (restated Faddeeva::w)
    shard      None, "sightlines" or "particles": partition over the ranks of torch.distributed
    group      process group (default: the world group)
    resident   keep particles + candidate index in HBM between calls (default: on when CUDA is there)
    backend    object providing _Particle_Interpolate / _near_lines (default: the CUDA boundary
               module; the CPU tests of the host logic pass a checker here)
    device_prep  resident route: prepare the particle arrays (selection, temperatures, neutral fractions, species
               densities) on the device from the raw snapshot fields instead of with host numpy (H I and ion == -1)
    seg_pairs  candidate pairs per work item of the kernels (fsb_params.seg_pairs).  None = one work row per
               sightline when sharded (every row is then bit-identical whatever the number of GPUs), automatic
               otherwise (few sightlines are cut into segments with private rows, summed in list order: faster on one
               GPU, same values to rounding); 1 << 30 forces one work row per sightline.
    """

    def __init__(self, num, base, cofm, axis, MPI=None, nbins=None, res=1., cdir=None, savefile="spectra.hdf5",
                 savedir=None, reload_file=False, spec_res=0, load_halo=False, units=None, sf_neutral=True,
                 turn_off_selfshield=False, quiet=False, load_snapshot=True, gasprop=None, gasprop_args=None,
                 kernel=None, use_external_Hz=None, precision="fp64", voigt="fast", shard=None, group=None,
                 resident=None, backend=None, seg_pairs=None, device_prep=True):
        _ = (load_halo, load_snapshot)
        self.num = num
        self.base = base
        self.MPI = MPI
        if MPI is not None:
            self.comm = MPI.COMM_WORLD
            self.rank = self.comm.Get_rank()
            self.size = self.comm.Get_size()
        else:
            self.comm = None
            self.rank = 0
            self.size = 1
        self.units = units if units is not None else unitsystem.UnitSystem()
        # result caches, keyed like the reference's (spectra.py:109-117)
        self.tau_obs, self.tau, self.sfr, self.vel_widths, self.absorber_width = {}, {}, {}, {}, {}
        self.colden, self.velocity, self.temp, self.dens_weight_dens = {}, {}, {}, {}
        self.part_ind = {}
        self.cofm_final = False
        self.num_important = {}
        self.discarded = 0
        self.npart = 0
        self.turn_off_selfshield = turn_off_selfshield
        self.spec_res = spec_res
        self.cdir = cdir
        self.minwidth = 500.
        self.tautail = 1e-7  # spectra.py:135
        self.seg_pairs = seg_pairs
        self.device_prep = bool(device_prep)
        self.precision = _PRECISION[precision]
        self.voigt = _VOIGT[voigt]
        self._backend = backend if backend is not None else _spectra_priv
        self._sharder = sharding.Sharder(shard, group) if shard is not None else sharding.Sharder("sightlines", group)
        if shard is None:
            self._sharder.rank, self._sharder.size = 0, 1  # no partition unless asked for
        self._engines = {}
        try:
            self.snapshot_set = absn.AbstractSnapshotFactory(num, base, comm=self.comm)
            if kernel is None:
                self.kernel_int = self.snapshot_set.get_kernel()
            elif kernel in _KERNELS:
                self.kernel_int = _KERNELS[kernel]
            else:
                raise ValueError("Unrecognised kernel %s" % (kernel,))
        except IOError:
            pass
        if savedir is None and isinstance(base, str):
            savedir = path.join(base, "snapdir_" + str(num).rjust(3, '0'))
            if not path.exists(savedir):
                savedir = path.join(base, "SPECTRA_" + str(num).rjust(3, '0'))
        self.savefile = path.join(savedir, savefile) if savedir is not None else savefile

        if reload_file:
            if not quiet:
                print("Reloading from snapshot (will save to: ", self.savefile, " )", flush=True)
            if cofm is None or axis is None:
                raise RuntimeError("None was passed for cofm or axis. If you are trying to load from a savefile, "
                                   "use reload_file=False.")
            if np.shape(cofm) == (3,):
                cofm = np.array([cofm, ])
            self.cofm = np.asarray(cofm).astype(np.float64)
            if np.shape(axis) == ():
                axis = np.array([axis])
            self.axis = np.asarray(axis).astype(np.int32)
            try:
                self.npart = self.snapshot_set.get_npart()
            except AttributeError as ae:
                raise IOError("Unable to load snapshot ", num, base) from ae
            self.box = self.snapshot_set.get_header_attr("BoxSize")
            self.atime = self.snapshot_set.get_header_attr("Time")
            self.red = 1 / self.atime - 1.
            self.hubble = self.snapshot_set.get_header_attr("HubbleParam")
            self.OmegaM = self.snapshot_set.get_header_attr("Omega0")
            self.OmegaLambda = self.snapshot_set.get_header_attr("OmegaLambda")
            self.omegab = self.snapshot_set.get_omega_baryon()
            try:
                self.units = self.snapshot_set.get_units()
            except KeyError:
                if not quiet:
                    print('No units found. Using kpc/kms/10^10Msun by default')
            self.Hz = use_external_Hz if use_external_Hz else None
        else:
            if not quiet:
                print("Reading pre-computed spectra (from file", self.savefile, " )", flush=True)
            self.load_savefile(self.savefile)

        # conversion factors from internal units (spectra.py:210-232)
        self.rscale = np.float32((self.units.UnitLength_in_cm * self.atime) / self.hubble)
        if self.Hz is None:
            self.Hz = 100.0 * self.hubble * np.sqrt(self.OmegaM / self.atime ** 3 + self.OmegaLambda)
        self.velfac = self.rscale * self.Hz / 3.085678e24
        self.vmax = self.box * self.velfac
        self.NumLos = np.size(self.axis)
        if reload_file:
            if res is None:
                if nbins is not None:
                    self.nbins = nbins
                    res = self.vmax / (1. * nbins)
                else:
                    raise ValueError('pixel width (res) not provided')
            if nbins is None:
                self.nbins = int(self.vmax / res)
            else:
                self.nbins = int(nbins)
            self.dvbin = self.vmax / (1. * self.nbins)
        else:
            self.dvbin = self.vmax / (1. * self.nbins)
            if res is not None:
                assert np.isclose(self.dvbin, res, rtol=1e-2), 'pixel width error'
            if use_external_Hz:
                assert np.isclose(self.Hz, use_external_Hz, rtol=1e-4), 'Hz error'
        self.species = ['H', 'He', 'C', 'N', 'O', 'Ne', 'Mg', 'Si', 'Fe', 'Z']
        self.solar = {"H": 1, "He": 0.0851, "C": 2.69e-4, "N": 6.76e-5, "O": 4.9e-4, "Ne": 8.51e-5, "Mg": 3.98e-5,
                      "Si": 3.24e-5, "Fe": 3.16e-5}
        self.solarz = 0.0134 / 0.7381
        self.lines = line_data.LineData()
        if gasprop is None:
            gasprop = gas_properties.GasProperties
        try:
            gprop_args = {"redshift": self.red, "absnap": self.snapshot_set, "hubble": self.hubble, "units": self.units,
                          "sf_neutral": sf_neutral}
            if gasprop_args is not None:
                gprop_args.update(gasprop_args)
            self.gasprop = gasprop(**gprop_args)
        except AttributeError:
            pass
        if resident is None:
            resident = backend is None and self._cuda_available()
        self.resident = bool(resident)
        self._set_my_sightlines()
        if not quiet:
            print(self.NumLos, " sightlines. resolution: ", self.dvbin, " z=", self.red)

    # ---- partition -----------------------------------------------------------------------------------
    @staticmethod
    def _cuda_available():
        try:
            import torch
            return torch.cuda.is_available()
        except ImportError:
            return False

    def _set_my_sightlines(self):
        sl = self._sharder.my_sightlines(self.NumLos)
        self._my_slice = sl
        self._my_cofm = np.ascontiguousarray(self.cofm[sl])
        self._my_axis = np.ascontiguousarray(self.axis[sl])
        self._engines = {}
        self.part_ind = {}

    def balance_sightlines(self, weights=None, segment=0):
        """Re-partition the sightlines over the ranks (sightline-sharded mode) into contiguous blocks of equal
        WORK instead of equal count: ``weights`` per sightline, by default the number of candidate particles of
        snapshot segment ``segment`` (one cheap count pass on the device; every rank holds the same particles in
        this mode, so every rank derives the same blocks without a collective).  Returns the block edges.
        Drops cached device state; results already computed stay valid (they are stored for all sightlines)."""
        if self._sharder.mode != "sightlines" or self._sharder.size == 1:
            return self._sharder.set_sightlines(self.NumLos)
        if weights is None:
            pos = self.snapshot_set.get_data(0, "Position", segment=segment).astype(np.float32)
            hh = self.snapshot_set.get_smooth_length(0, segment=segment).astype(np.float32)
            weights = self._backend._count_pairs(self.box, pos, hh, self.axis, self.cofm)
        edges = self._sharder.set_sightlines(self.NumLos, weights=np.asarray(weights))
        self._set_my_sightlines()
        return edges

    def set_sightlines(self, cofm, axis):
        """Replace the sightlines (drops cached particle lists, device state and results)."""
        self.cofm = np.asarray(cofm).astype(np.float64)
        self.axis = np.asarray(axis).astype(np.int32)
        self.NumLos = np.size(self.axis)
        for cache in (self.tau, self.colden, self.velocity, self.temp, self.dens_weight_dens, self.tau_obs):
            cache.clear()
        self._set_my_sightlines()

    # ---- savefile (reference spectra.py:266-372,434-499) -----------------------------------------------
    def load_savefile(self, savefile=None):
        """The reference keeps results in an HDF5 file; h5py is not part of this environment and the
        file format is outside the interpolation path."""
        from . import savefile as sf
        sf.load(self, savefile)

    def save_file(self):
        """Writes every result this object holds, including arrays of a reopened savefile that were never touched
        (they are loaded first, like the reference's _load_all_multihash, spectra.py:275-282: the old file becomes
        the .backup and must not be the only copy of anything).  Only one process writes: rank 0 of the sharding group,
        of MPI, and of torch.distributed when a process group exists without sharding."""
        from . import savefile as sf
        for cache, name in ((self.tau_obs, "tau_obs"), (self.tau, "tau"), (self.colden, "colden"), (self.velocity, "velocity"),
                            (self.temp, "temperature"), (self.dens_weight_dens, "density_weight_density")):
            for key in list(cache.keys()):
                self._really_load_array(key, cache, name)
        global_rank = 0
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                global_rank = dist.get_rank()
        except ImportError:
            pass
        if self._sharder.rank == 0 and self.rank == 0 and global_rank == 0:
            sf.save(self, self.savefile)

    def save_cextract(self, path, elem="H", ion=1, line=1215):
        """Optical depths and column densities of one line in the binary layout of the reference's C extractor
        (cextract/main.cpp:247-257; savefile.write_cextract)."""
        from . import savefile as sf
        sf.write_cextract(path, self.red, self.box, self.get_tau(elem, ion, line), self.get_col_density(elem, ion))

    def _really_load_array(self, key, array, array_name):
        """Lazy loading of saved arrays: a one-element placeholder means 'on disc'."""
        if np.size(array[key]) > 1:
            return
        from . import savefile as sf
        array[key] = sf.load_array(self.savefile, array_name, key)

    # ---- particle data (spectra.py:550-617,675-711) ----------------------------------------------------
    def particles_near_lines(self, pos, hh, axis=None, cofm=None):
        """Index list of the particles whose kernel reaches at least one sightline."""
        if axis is None:
            axis = self._my_axis
        if cofm is None:
            cofm = self._my_cofm
        if np.size(axis) == 0:
            return np.zeros(0, dtype=np.int32)
        assert np.min(axis) > 0
        assert np.max(axis) < 4
        return self._backend._near_lines(self.box, pos, hh, np.ascontiguousarray(axis, dtype=np.int32),
                                         np.ascontiguousarray(cofm, dtype=np.float64))

    def get_mass_frac(self, elem, fn, ind):
        """Mass fraction (float32, never negative) of ``elem`` for the particles ``ind`` of segment ``fn``: the snapshot's
        metal table when it has one, primordial hydrogen / helium otherwise; "Z" is the total metallicity
        (spectra.py:686-710)."""
        snap = self.snapshot_set
        if elem == "Z":
            column = snap.get_data(0, "Metallicity", segment=fn)
        else:
            which = self.species.index(elem)
            try:
                column = snap.get_data(0, "GFM_Metals", segment=fn)[:, which]
            except KeyError:
                primordial = (0.76, 0.24)  # no metal table: hydrogen and helium only (IndexError for anything else)
                column = np.full(snap.get_blocklen(0, "Density", segment=fn), primordial[which], dtype=np.float32)
        return np.maximum(np.asarray(column, dtype=np.float32)[ind], np.float32(0))

    def _filter_particles(self, elem_den, pos, velocity, den):
        _ = (pos, velocity, den)
        return np.where(elem_den > 0)

    def _cloudy(self):
        """The Cloudy table at this redshift (spectra.py:641-645): ``self.cloudy_table`` when the caller set one, else
        read from ``cdir`` / $FAKE_SPECTRA_CLOUDY_DIR on first use; None when there is neither."""
        table = self.__dict__.get("cloudy_table")
        if table is None and (self.cdir is not None or cloudy.default_directory() is not None):
            table = self.cloudy_table = cloudy.CloudyTable(self.red, self.cdir)
        return table

    def _get_elem_den(self, elem, ion, den, temp, ind, ind2):
        """Ionisation fraction of a metal ion for host-prepared particles (spectra.py:637-664): the Cloudy table with
        densities and temperatures clipped to its bounds; without a table a snapshot (or subclass) may provide
        ``ion_fraction(elem, ion, nH, temp)``."""
        _ = (ind, ind2)
        table = self._cloudy()
        if table is None:
            fn = getattr(self.snapshot_set, "ion_fraction", None)
            if fn is None:
                raise IOError("ion fractions for %s %d need a Cloudy table (cdir=, $FAKE_SPECTRA_CLOUDY_DIR or a "
                              "cloudy_table attribute) or a snapshot with ion_fraction(elem, ion, nH, temp)" % (elem, ion))
            return np.float32(fn(elem, ion, den, temp))
        clipped = []
        for values, (lo, hi) in ((den, table.get_dens_bounds()), (temp, table.get_temp_bounds())):
            values = np.array(values)  # the lookup scales its density argument in place
            values[values > hi] = hi
            values[values < lo] = lo
            clipped.append(values)
        return np.float32(table.ion(elem, ion, clipped[0], clipped[1]))

    def _near_particles(self, fn, pos, hh):
        """Indices of the particles of segment ``fn`` whose kernel reaches one of this rank's sightlines (kept per
        segment once the sightline set is final; this rank's slice of them in particle-sharded mode)."""
        if self.cofm_final and fn in self.part_ind:
            ind = self.part_ind[fn]
        else:
            ind = self.particles_near_lines(pos, hh)
            if self.cofm_final:
                self.part_ind[fn] = ind
        if self._sharder.mode == "particles" and self._sharder.size > 1:
            ind = ind[self._sharder.my_particles(np.size(ind))]
        return ind

    def _read_particle_data(self, fn, elem, ion, get_tau):
        """Host-prepared inputs of the interpolation for segment ``fn`` (the device route is _device_particle_data):
        (pos, vel, elem_den, temp, hh, amumass), all float32, for the particles near this rank's sightlines; six times
        False when nothing is left.  Values as the reference forms them (spectra.py:550-617): species density =
        (n_H rscale) x mass fraction x neutral or ion fraction / atomic mass; velocities and temperatures are only read
        when something needs them (one-element placeholders otherwise)."""
        nothing = (False,) * 6
        snap, gas = self.snapshot_set, self.gasprop
        f32 = lambda a: np.asarray(a).astype(np.float32)  # noqa: E731
        pos, hh = f32(snap.get_data(0, "Position", segment=fn)), f32(snap.get_smooth_length(0, segment=fn))
        ind = self._near_particles(fn, pos, hh)
        if np.size(ind) == 0:
            return nothing
        pos, hh = pos[ind, :], hh[ind]
        placeholder = np.zeros(1, dtype=np.float32)
        vel = f32(snap.get_peculiar_velocity(0, segment=fn))[ind, :] if get_tau else placeholder
        den = f32(gas.get_code_rhoH(0, segment=fn))[ind]
        metal_ion = ion != -1 and not (elem == "H" and ion == 1)
        temp = placeholder
        if get_tau or (ion != -1 and elem != "H"):
            temp = f32(gas.get_temp(0, segment=fn))[ind]
            temp[temp <= 0] = 1
        amumass = 1 if elem == "Z" else self.lines.get_mass(elem)
        elem_den = (den * self.rscale) * self.get_mass_frac(elem, fn, ind)
        if elem == "H" and ion == 1:
            elem_den *= f32(gas.get_reproc_HI(0, segment=fn)[ind])
        elif metal_ion:
            keep = self._filter_particles(elem_den, pos, vel, den)  # particles with mass in the element
            if np.size(keep) == 0:
                return nothing
            pos, hh, temp = pos[keep], hh[keep], temp[keep]
            if get_tau:
                vel = vel[keep]
            elem_den = elem_den[keep] * self._get_elem_den(elem, ion, den[keep], temp, ind, keep)
        elem_den /= amumass
        return (pos, vel, np.ascontiguousarray(elem_den, dtype=np.float32), temp, hh, amumass)

    # ---- particle data prepared on the device (row f3) ------------------------------------------------------
    _RAW_FIELDS = ("Position", "Velocities", "Density", "InternalEnergy", "ElectronAbundance", "NeutralHydrogenFraction")

    def _device_prep_ok(self, elem, ion):
        """The device route covers H I (snapshot neutral fraction + Rahmati et al. 2013 above the star-formation
        threshold), all ionisation states of an element (ion == -1) and metal ions looked up in a Cloudy table
        (cloudy.CloudyTable).  Needs the stock GasProperties (a user-supplied gasprop class, an overridden
        _get_elem_den / _filter_particles or a snapshot-provided ion_fraction keep the host route) and the CUDA boundary."""
        if not (self.resident and self.device_prep and self._backend is _spectra_priv and self._cuda_available()):
            return False
        if type(self.gasprop) is not gas_properties.GasProperties or elem == "Z":
            return False
        if (elem == "H" and ion == 1) or ion == -1:
            return True
        cls = type(self)
        if cls._get_elem_den is not Spectra._get_elem_den or cls._filter_particles is not Spectra._filter_particles:
            return False
        return isinstance(self._cloudy(), cloudy.CloudyTable)

    def _raw_fields(self, fn):
        """Raw float32 fields of a snapshot segment in HBM (one segment is kept: the next species of the same
        segment reuses the upload)."""
        import torch
        cache = self.__dict__.setdefault("_raw_cache", {})
        if fn in cache:
            return cache[fn]
        cache.clear()
        snap = self.snapshot_set
        dev = torch.device("cuda", torch.cuda.current_device())
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)  # noqa: E731
        raw = {}
        for name in self._RAW_FIELDS:
            try:
                raw[name] = up(snap.get_data(0, name, segment=fn))
            except KeyError:
                raw[name] = None
        from . import native
        try:  # the three cases of get_smooth_length (abstractsnapshot.py:253-282)
            raw["hh"] = native.smoothing_lengths(up(snap.get_data(0, "Volume", segment=fn)), mode=1)
        except KeyError:
            try:
                raw["hh"] = native.smoothing_lengths(up(snap.get_data(0, "SmoothingLength", segment=fn)), mode=0)
            except KeyError:
                raw["hh"] = native.smoothing_lengths(up(snap.get_data(0, "Masses", segment=fn)), raw["Density"], mode=2)
        try:
            raw["metals"] = up(snap.get_data(0, "GFM_Metals", segment=fn))
        except KeyError:
            raw["metals"] = None
        cache[fn] = raw
        return raw

    def _device_particle_data(self, fn, elem, ion):
        """_read_particle_data on the device: (pos, vel, elem_den, temp, hh, amumass) as CUDA tensors for the particles
        of segment ``fn`` near this rank's sightlines (six times False when there are none).  Same selection
        (fsb_near_lines), same formulae, float32 like the host route (values agree to float32 rounding)."""
        import torch
        from . import _lib, native
        none = (False, False, False, False, False, False)
        raw = self._raw_fields(fn)
        if raw["Position"] is None or raw["Position"].shape[0] == 0 or np.size(self._my_axis) == 0:
            return none
        dev = raw["Position"].device
        cofm = torch.from_numpy(np.ascontiguousarray(self._my_cofm)).to(dev)
        axis = torch.from_numpy(np.ascontiguousarray(self._my_axis)).to(dev)
        ind = native.near_lines(self.box, raw["Position"], raw["hh"], axis, cofm)
        if self._sharder.mode == "particles" and self._sharder.size > 1:
            ind = ind[self._sharder.my_particles(int(ind.shape[0]))].contiguous()
        if ind.shape[0] == 0:
            return none
        gp, units = self.gasprop, self.units
        amumass = self.lines.get_mass(elem)
        cfg = _lib.Prep()
        # Gadget HDF5: velocities times sqrt(a); MP-Gadget BigFile: divided by a unless already peculiar
        if hasattr(self.snapshot_set, "velocity_divisor"):
            cfg.velocity_factor, cfg.velocity_divides = self.snapshot_set.velocity_divisor(), 1
        else:
            cfg.velocity_factor, cfg.velocity_divides = np.sqrt(self.snapshot_set.get_header_attr("Time")), 0
        cfg.dens_conv = gp._density_conversion()
        cfg.rscale = self.rscale
        cfg.unit_ienergy = units.UnitInternalEnergy_in_cgs
        cfg.temp_factor = (units.gamma - 1) * units.protonmass / units.boltzmann
        # numpy scalars (unit systems built from file headers) promote the float32 fields to double, Python floats do not
        cfg.temp_double = 1 if isinstance(units.UnitInternalEnergy_in_cgs, np.floating) else 0
        cfg.hy_mass = 0.76
        cfg.nelec_const = 1.0
        nelem = self.species.index(elem)
        cfg.mass_frac_const = float(np.array([0.76, 0.24], dtype=np.float32)[nelem]) if (raw["metals"] is None and nelem < 2) else 0.0
        cfg.amumass = np.float32(amumass)
        cfg.dens_thresh_code = np.float32(gp.PhysDensThresh / 0.76 / gp._density_conversion())
        cfg.neutral_hydrogen = 1 if (elem == "H" and ion == 1) else 0
        cfg.sf_neutral = 1 if gp.sf_neutral else 0
        cfg.redshift_coverage = 1 if gp.redshift_coverage else 0
        if gp.redshift_coverage:
            cfg.gray_opac, cfg.gamma_uvb = gp.gray_opac, gp.gamma_UVB
        cfg.f_bar = gp.f_bar
        if raw["metals"] is None and nelem >= 2:
            raise KeyError("GFM_Metals")  # like the reference: no metal table, no metal mass fraction (spectra.py:700-706)
        mass_frac = raw["metals"][:, nelem] if raw["metals"] is not None else None
        ion_table = None
        if not cfg.neutral_hydrogen and ion != -1:
            # metal ion: drop the particles without mass in the element, then look the ion fraction up (spectra.py:598-610)
            ind = native.select_particles(cfg, ind, raw["Density"], mass_frac)
            if ind.shape[0] == 0:
                return none
            ion_table, _owner = self._cloudy().device_table(elem, ion, dev)
        for name in ("InternalEnergy", "ElectronAbundance") + (("NeutralHydrogenFraction",) if cfg.neutral_hydrogen else ()):
            if raw[name] is None:
                raise KeyError(name)  # as the reference's get_temp / get_reproc_HI would
        pos, vel, elem_den, temp, hh = native.prepare_particles(
            cfg, ind, raw["Position"], raw["Velocities"], raw["Density"], raw["InternalEnergy"], raw["ElectronAbundance"],
            raw["NeutralHydrogenFraction"], raw["hh"], mass_frac, ion_table=ion_table)
        return (pos, vel, elem_den, temp, hh, amumass)

    def find_all_particles(self):
        """Positions and smoothing lengths of all particles near sightlines."""
        pp = np.empty([0, 3])
        hhh = np.array([])
        for i in range(self.snapshot_set.get_n_segments()):
            (pos, _, _, _, hh, amumass) = self._read_particle_data(i, "H", -1, False)
            if amumass is not False:
                pp = np.concatenate([pp, pos])
                hhh = np.concatenate([hhh, hh])
        return pp, hhh

    # ---- the native boundary ---------------------------------------------------------------------------
    def _line(self, elem, ion, ll):
        if ion == -1:
            for ii in range(8):
                try:
                    return self.lines[(elem, ii)][ll]
                except KeyError:
                    continue
            raise KeyError((elem, ion, ll))
        return self.lines[(elem, ion)][ll]

    def _params(self, line, amumass, unsegmented=False):
        from . import _lib
        gamma_X = 0 if self.turn_off_selfshield else line.gamma_X
        # sharded runs keep one work row per sightline: rows are then bit-identical whatever the partition
        seg = (1 << 30) if (unsegmented or self._sharder.size > 1) else 0
        if self.seg_pairs is not None:
            seg = int(self.seg_pairs)
        return _lib.make_params(self.nbins, self.kernel_int, self.box, self.velfac, self.atime, line.lambda_X * 1e-8, gamma_X,
                                line.fosc_X, amumass, self.tautail, precision=self.precision, voigt=self.voigt, seg_pairs=seg)

    def _push_rows(self):
        """True when the ranks exchange result rows through peer-mapped arrays: sightline-sharded, several ranks,
        NCCL process group (one process per GPU of one node), every rank with at least one engine decides alike."""
        if self._sharder.mode != "sightlines" or self._sharder.size == 1 or self._backend is not _spectra_priv:
            return False
        import torch.distributed as dist
        return dist.is_initialized() and dist.get_backend(self._sharder.group) == "nccl"

    def _peer_rows(self, nl):
        """The full [nl, NumLos, nbins] array of this rank, mapped into every other rank (cached per shape)."""
        from . import native
        key = (nl, self.NumLos, self.nbins)
        cache = self.__dict__.setdefault("_peer_cache", {})
        if key not in cache:
            for old in cache.values():
                old.close()
            cache.clear()
            cache[key] = native.PeerRows(nl, self.NumLos, self.nbins, group=self._sharder.group)
        return cache[key]

    def _do_interpolation_work(self, pos, vel, elem_den, temp, hh, amumass, line, get_tau):
        """Run the interpolation on pre-determined host arrays (spectra.py:666-673): the drop-in
        boundary call, on this rank's sightlines."""
        gamma_X = 0 if self.turn_off_selfshield else line.gamma_X
        kw = {}
        if self._backend is _spectra_priv:
            kw = dict(precision=self.precision, voigt=self.voigt)
        if np.size(self._my_axis) == 0:
            return np.zeros([0, self.nbins])
        return self._backend._Particle_Interpolate(get_tau * 1, self.nbins, self.kernel_int, self.box, self.velfac, self.atime,
                                                   line.lambda_X * 1e-8, gamma_X, line.fosc_X, amumass, self.tautail, pos,
                                                   vel, elem_den, temp, hh, self._my_axis, self._my_cofm, **kw)

    def _engine(self, fn, elem, ion):
        """Device-resident particles + candidate index of a segment (None when it has no particles near this rank's
        sightlines).  Cached per (segment, element, ion) for the current sightline set (set_sightlines /
        _set_my_sightlines drop the cache), so that further lines of the ion, its column density and its weighted
        fields reuse the upload and the index.  Least-recently-used engines are released beyond ``max_engines``
        (device memory of multi-segment snapshots stays bounded)."""
        key = (fn, elem, ion)
        if key in self._engines:
            eng = self._engines.pop(key)
            self._engines[key] = eng  # most recently used last
            return eng
        if self._device_prep_ok(elem, ion):
            (pos, vel, elem_den, temp, hh, amumass) = self._device_particle_data(fn, elem, ion)
        else:
            (pos, vel, elem_den, temp, hh, amumass) = self._read_particle_data(fn, elem, ion, True)
        eng = None
        if amumass is not False and np.size(self._my_axis) > 0:
            eng = _SegmentEngine(self, pos, vel, elem_den, temp, hh)
            eng.amumass = amumass
        self._engines[key] = eng
        limit = getattr(self, "max_engines", 4)
        while len(self._engines) > limit:
            old_key = next(iter(self._engines))
            old = self._engines.pop(old_key)
            if old is not None:
                old.release()
        return eng

    def _segments(self):
        """Segments this call loops over (the Voronoi kernel needs all particles at once:
        spectra.py:811-815)."""
        if self.kernel_int == 2:
            # "all segments" is a negative segment number in the reference's snapshot API (abstractsnapshot.py:214,364)
            return [-1] if self.snapshot_set.get_n_segments(part_type=0) > 1 else [0]
        return list(range(self.snapshot_set.get_n_segments(part_type=0)))

    def _interpolate_single_file(self, nsegment, elem, ion, ll, get_tau, load_all_data_first=False):
        """Read arrays and interpolate one segment through the drop-in boundary (spectra.py:501-548)."""
        seg = -1 if load_all_data_first else nsegment
        (pos, vel, elem_den, temp, hh, amumass) = self._read_particle_data(seg, elem, ion, get_tau)
        if amumass is False:
            return np.zeros([np.shape(self._my_cofm)[0], self.nbins], dtype=np.float32)
        line = self._line(elem, ion, ll) if get_tau else self.lines[("H", 1)][1215]
        return self._do_interpolation_work(pos, vel, elem_den, temp, hh, amumass, line, get_tau)

    def _combine(self, local):
        """Recombine the ranks' pieces [K, rows, nbins] (torch tensor or numpy; rows = this rank's sightline block, or
        all sightlines in particle-sharded mode) into the full float64 numpy array [K, NumLos, nbins]."""
        import torch
        was_numpy = isinstance(local, np.ndarray)
        t = torch.from_numpy(np.ascontiguousarray(local, dtype=np.float64)) if was_numpy else local
        if self._sharder.size > 1:
            if was_numpy and self._cuda_available() and self._backend is _spectra_priv:
                t = t.cuda()
            t = self._sharder.combine(t, self.NumLos, dim=1)
        return t.cpu().numpy() if t.is_cuda else t.numpy()

    def compute_spectra(self, elem, ion, ll, get_tau):
        """tau (get_tau) or column density of one species on every sightline: loop over the
        snapshot's segments, accumulate, recombine the ranks (spectra.py:801-831)."""
        return self.compute_spectra_lines(elem, ion, [ll], get_tau)[0]

    def compute_spectra_lines(self, elem, ion, lls, get_tau):
        """Several lines of one ion in one pass over the particles (they share the candidate index
        and every per-particle quantity but the line constants): [len(lls), NumLos, nbins]."""
        nlocal = np.shape(self._my_cofm)[0]
        nl = len(lls) if get_tau else 1
        if self.resident:
            import torch
            acc = torch.zeros((nl, nlocal, self.nbins), dtype=torch.float64, device="cuda")
            # sightline-sharded optical depths on several GPUs: the kernel of the LAST segment stores every finished row
            # (all segments summed) into every rank's full array over NVLink; no gather afterwards (native.PeerRows)
            peer = self._peer_rows(nl) if (get_tau and self.kernel_int != 2 and self._push_rows()) else None
            segs = self._segments()
            pushed = False
            for n, seg in enumerate(segs):
                eng = self._engine(seg, elem, ion)  # used at once: a later _engine call may release it (LRU bound)
                if eng is None:
                    continue
                if get_tau:
                    push = peer.push_spec(0, self._my_slice.start) if (peer is not None and n == len(segs) - 1) else None
                    eng.tau([self._params(self._line(elem, ion, ll), eng.amumass, unsegmented=peer is not None) for ll in lls],
                            out=acc, push=push)
                    pushed = pushed or push is not None
                else:
                    acc += eng.colden(self._params(self.lines[("H", 1)][1215], eng.amumass))
            if peer is not None:
                if not pushed:  # the last segment has no particle near this rank's sightlines: plain copies to every array
                    for v in peer.views:
                        v[:, self._my_slice.start:self._my_slice.start + nlocal] = acc
                torch.cuda.synchronize()
                peer.barrier()
                return peer.full.cpu().numpy()
            local = acc
        else:
            arepo = self.kernel_int == 2
            out = []
            for ll in (lls if get_tau else [0]):
                result = np.array(self._interpolate_single_file(0, elem, ion, ll, get_tau, load_all_data_first=arepo),
                                  dtype=np.float64)
                nseg = 1 if arepo else self.snapshot_set.get_n_segments(part_type=0)
                for nn in range(1, nseg):
                    result += self._interpolate_single_file(nn, elem, ion, ll, get_tau)
                out.append(result)
            local = np.stack(out, axis=0)
        result = self._combine(local)  # [nl, NumLos, nbins]
        if self.MPI is not None:
            # the reference's MPI mode: ranks hold different particles, float32 sum (spectra.py:828-830)
            result = np.ascontiguousarray(result, np.float32)
            self.comm.Allreduce(self.MPI.IN_PLACE, result, op=self.MPI.SUM)
        return result

    # ---- public getters (spectra.py:862-893, 945-1043) -------------------------------------------------
    def get_col_density(self, elem, ion, force_recompute=False):
        """Column density in each pixel, [metal] ions cm^-2."""
        try:
            if force_recompute:
                raise KeyError
            self._really_load_array((elem, ion), self.colden, "colden")
            return self.colden[(elem, ion)]
        except KeyError:
            colden = self.compute_spectra(elem, ion, 0, False)
            self.colden[(elem, ion)] = colden
            return colden

    def get_density(self, elem, ion, force_recompute=False):
        """Density in each pixel, [metal] ions cm^-3."""
        colden = self.get_col_density(elem, ion, force_recompute)
        phys = self.dvbin / self.velfac * self.rscale
        return colden / phys

    def get_tau(self, elem, ion, line, number=-1, force_recompute=False):
        """Optical depth in each pixel for one line (``line`` = int(lambda in Angstrom))."""
        try:
            if force_recompute:
                raise KeyError
            self._really_load_array((elem, ion, line), self.tau, "tau")
            tau = self.tau[(elem, ion, line)]
        except KeyError:
            tau = self.compute_spectra(elem, ion, line, True)
            self.tau[(elem, ion, line)] = tau
        if number >= 0:
            tau = tau[number, :]
        return tau

    def get_tau_lines(self, elem, ion, lines, force_recompute=False):
        """Optical depths of several lines of one ion from one pass (Lya + Lyb ...); fills the same
        cache as get_tau.  Returns {line: tau}."""
        todo = [ll for ll in lines if force_recompute or (elem, ion, ll) not in self.tau]
        if todo:
            taus = self.compute_spectra_lines(elem, ion, todo, True)
            for ll, t in zip(todo, taus):
                self.tau[(elem, ion, ll)] = t
        return {ll: self.get_tau(elem, ion, ll) for ll in lines}

    def _weighted_single_file(self, fn, elem, ion, make_weights):
        """Column-density pass(es) with reweighted densities for one segment: [K, nlocal, nbins].
        make_weights(pos, vel, elem_den, temp) -> list of float32 weight arrays."""
        nlocal = np.shape(self._my_cofm)[0]
        line = self.lines[("H", 1)][1215]
        if self.resident:
            import torch
            eng = self._engine(fn, elem, ion)
            if eng is None:
                return None
            w = make_weights(eng.pos, eng.vel, eng.dens, eng.temp, torch, fn)
            return eng.colden(self._params(line, eng.amumass), torch.stack(w))
        (pos, vel, elem_den, temp, hh, amumass) = self._read_particle_data(fn, elem, ion, True)
        if amumass is False:
            return None
        w = make_weights(pos, vel, elem_den, temp, np, fn)
        return np.stack([np.asarray(self._do_interpolation_work(pos, vel, np.ascontiguousarray(wi, dtype=np.float32), temp, hh,
                                                                amumass, line, False), dtype=np.float64) for wi in w]
                        ).reshape(len(w), nlocal, self.nbins)

    def _get_mass_weight_quantity(self, make_weights, nweights, elem, ion):
        """Sum the weighted passes over segments, recombine the ranks, divide by the density
        (spectra.py:958-981).  Returns [K, NumLos, nbins]."""
        nlocal = np.shape(self._my_cofm)[0]
        acc = None
        for seg in self._segments():
            r = self._weighted_single_file(seg, elem, ion, make_weights)
            if r is None:
                continue
            acc = r if acc is None else acc + r
        if acc is None:
            acc = np.zeros((nweights, nlocal, self.nbins))
        result = self._combine(acc)
        den = np.array(self.get_density(elem, ion))
        den[np.where(den == 0.)] = 1
        return result / den[None, :, :]

    def get_velocity(self, elem, ion):
        """Column-density weighted velocity in each pixel, [NumLos, nbins, 3] km/s; the three
        components share one geometry pass (the reference makes three calls, spectra.py:945-956)."""
        try:
            self._really_load_array((elem, ion), self.velocity, "velocity")
            return self.velocity[(elem, ion)]
        except KeyError:
            phys = np.float32(self.dvbin / self.velfac * self.rscale)
            sqa = np.float32(np.sqrt(self.atime))

            def weights(pos, vel, elem_den, temp, xp, fn):
                return [elem_den * (vel[:, ax] * sqa) / phys for ax in (0, 1, 2)]

            vv = self._get_mass_weight_quantity(weights, 3, elem, ion)
            velocity = np.ascontiguousarray(np.transpose(vv, (1, 2, 0)).astype(np.float32))
            self.velocity[(elem, ion)] = velocity
            return velocity

    def get_temp(self, elem, ion):
        """Density weighted temperature in each pixel."""
        try:
            self._really_load_array((elem, ion), self.temp, "temperature")
            return self.temp[(elem, ion)]
        except KeyError:
            phys = np.float32(self.dvbin / self.velfac * self.rscale)

            def weights(pos, vel, elem_den, temp, xp, fn):
                return [elem_den * temp / phys]

            temp = self._get_mass_weight_quantity(weights, 1, elem, ion)[0]
            self.temp[(elem, ion)] = temp
            return temp

    def get_dens_weighted_density(self, elem, ion):
        """(Ion) density weighted (species) density in each pixel (spectra.py:1015-1043)."""
        try:
            self._really_load_array((elem, ion), self.dens_weight_dens, "density_weight_density")
            return self.dens_weight_dens[(elem, ion)]
        except KeyError:
            phys = np.float32(self.dvbin / self.velfac * self.rscale)
            spec = self

            def weights(pos, vel, elem_den, temp, xp, fn):
                # density of all ionisation states of the element for the same particles
                (_, _, species, _, _, amumass) = spec._read_particle_data(fn, elem, -1, True)
                if amumass is False or species.shape[0] != elem_den.shape[0]:
                    raise ValueError("get_dens_weighted_density needs the ion and species particle selections to "
                                     "coincide (hydrogen, or ions without a zero-density filter)")
                if xp is not np:
                    species = xp.from_numpy(species).to(elem_den.device)
                return [(elem_den / phys) * (species / spec.rscale)]

            dwd = self._get_mass_weight_quantity(weights, 1, elem, ion)[0]
            self.dens_weight_dens[(elem, ion)] = dwd
            return dwd

    def get_observer_tau(self, elem, ion, number=-1, force_recompute=False):
        """Optical depth of the line of an ion whose maximum is closest to unity without being
        saturated (spectra.py:895-943); all lines come from one pass over the particles."""
        try:
            if force_recompute:
                raise KeyError
            self._really_load_array((elem, ion), self.tau_obs, "tau_obs")
            ntau = self.tau_obs[(elem, ion)]
        except KeyError:
            keys = list(self.lines[(elem, ion)].keys())
            tau = self.compute_spectra_lines(elem, ion, keys, True)
            from .spec_utils import res_corr
            maxtaus = np.max(res_corr(tau, self.dvbin, self.spec_res), axis=-1)
            ntau = np.empty([self.NumLos, self.nbins])
            for ii in range(self.NumLos):
                ind = np.where(np.logical_and(maxtaus[:, ii] < 3, maxtaus[:, ii] > 0.1))
                if np.size(ind) > 0:
                    line = np.where(maxtaus[:, ii] == np.max(maxtaus[ind, ii]))
                else:
                    ind2 = np.where(maxtaus[:, ii] > 0.1)
                    if np.size(ind2) > 0:
                        line = np.where(maxtaus[:, ii] == np.min(maxtaus[ind2, ii]))
                    else:
                        line = np.where(maxtaus[:, ii] == np.max(maxtaus[:, ii]))
                ntau[ii, :] = tau[line[0][0], ii, :]
            self.tau_obs[(elem, ion)] = ntau
        if number >= 0:
            ntau = ntau[number, :]
        return ntau

    # -- flux statistics (spectra.py:1254-1301): the reductions run on the device, fluxstatistics.py ------
    def _filter_single_tau_complex(self, tt, taueff, tau_thresh=1e6, thresh2=0.25):
        """One spectrum: damped regions set to ``taueff`` (spectra.py:1219-1252)."""
        from . import fluxstatistics as fstat
        return fstat.mask_damped_region(tt, taueff, tau_thresh=tau_thresh, thresh2=thresh2)

    def _filter_tau(self, tau, tau_thresh=None):
        """Sightlines with an optically thick absorber (maximum above ``tau_thresh``) have it masked
        (spectra.py:1254-1270); alters ``tau`` like the reference."""
        from . import fluxstatistics as fstat
        return fstat.filter_tau(tau, tau_thresh)

    def get_mean_flux(self, elem="H", ion=1, line=1215, tau_thresh=None):
        """Mean flux <exp(-tau)> along the sightlines (spectra.py:1272-1276)."""
        from . import fluxstatistics as fstat
        tau = self._filter_tau(self.get_tau(elem, ion, line), tau_thresh=tau_thresh)
        sf, _, used = fstat.flux_sums(tau)
        return sf / used

    def get_flux_pdf(self, elem="H", ion=1, line=1215, nbins=20, mean_flux_desired=None, tau_thresh=None):
        """Flux PDF: (bin centres, normalised histogram of exp(-tau)) (spectra.py:1278-1282)."""
        from . import fluxstatistics as fstat
        tau = self._filter_tau(self.get_tau(elem, ion, line), tau_thresh=tau_thresh)
        return fstat.flux_pdf(tau, nbins=nbins, mean_flux_desired=mean_flux_desired)

    def get_flux_power_1D(self, elem="H", ion=1, line=1215, mean_flux_desired=None, window=False, tau_thresh=None):
        """1-D power spectrum of delta_F = exp(-tau)/<F> - 1 averaged over the sightlines, without the k = 0
        mode: (k [s/km], P_F [km/s]) (spectra.py:1284-1301)."""
        from . import fluxstatistics as fstat
        tau = self._filter_tau(self.get_tau(elem, ion, line), tau_thresh=tau_thresh)
        if mean_flux_desired is not None and window is True and self.spec_res > 0:
            raise ValueError("Cannot sensibly rescale mean flux with gaussian smoothing")
        kf, avg_flux_power = fstat.flux_power(tau, self.vmax, spec_res=self.spec_res, mean_flux_desired=mean_flux_desired,
                                              window=window)
        return kf[1:], avg_flux_power[1:]

    def get_cofm(self, num=None):
        """Find a bunch more sightlines: should be overridden by child classes"""
        raise NotImplementedError

    def filter_DLA(self, col_den, thresh=10 ** 20.3):
        """Indices of sightlines whose summed column density exceeds ``thresh`` (or lies in a range)."""
        cdsum = np.sum(col_den, axis=1)
        if np.size(thresh) > 1:
            return np.where(np.logical_and(cdsum > thresh[0], cdsum < thresh[1]))
        return np.where(cdsum > thresh)

    def replace_not_DLA(self, ndla, thresh=10 ** 20.3, elem="H", ion=1):
        """Keep drawing sightlines until ``ndla`` of them exceed the column-density threshold
        (spectra.py:713-755)."""
        found = 0
        wanted = ndla
        cofm_DLA = np.empty_like(self.cofm)[:ndla, :]
        col_den_DLA = np.empty((ndla, self.nbins))
        self.set_sightlines(self.cofm, self.axis)
        while True:
            col_den = self.compute_spectra(elem, ion, 1215, False)
            ind = self.filter_DLA(col_den, thresh)
            top = np.min([wanted, found + np.size(ind)])
            cofm_DLA[found:top] = self.cofm[ind][:top - found, :]
            col_den_DLA[found:top] = col_den[ind][:top - found, :]
            found += np.size(ind)
            self.discarded += self.NumLos - np.size(ind)
            print("Discarded: ", self.discarded)
            if found >= wanted:
                break
            self.set_sightlines(self.get_cofm(), self.axis)
        self.discarded = int(self.discarded * 1. * wanted / max(found, 1))  # spectra.py:748
        self.set_sightlines(cofm_DLA, np.ones(ndla) if np.size(self.axis) < ndla else self.axis[:ndla])
        self.colden[(elem, ion)] = col_den_DLA
        self.cofm_final = True
