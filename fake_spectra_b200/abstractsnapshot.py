"""Snapshot access for the host classes (the readers either side of the hot path).

Same interface as the reference's ``abstractsnapshot`` (abstractsnapshot.py:16-400): a factory that returns an
object with ``get_header_attr, get_data, get_n_segments, get_blocklen, get_smooth_length, get_peculiar_velocity,
get_temp, get_kernel, get_npart, get_omega_baryon, get_units``.  Three sources:

* any in-memory object with that duck-type passed as ``base`` (e.g. :class:`fake_spectra_b200.synthetic.SyntheticSnapshot`);
* :class:`BigFileSnapshot` -- MP-Gadget's BigFile layout read with numpy alone (a block is a directory holding a text
  ``header`` -- DTYPE / NMEMB / NFILE and one ``name: items : checksum : 0`` line per data file -- raw little-endian data
  files ``000000, 000001, ...`` and a text ``attr-v2`` with one ``name dtype nmemb HEXBYTES #HUMANE [...]`` line per
  attribute).  The ``bigfile`` package is absent from this environment, so this reader is pinned only by the writer in
  ``write_bigfile_block`` / the round-trip tests, not against files written by MP-Gadget: **parity unpinned**;
* :class:`HDF5Snapshot` -- Gadget/Arepo HDF5 through h5py when h5py is importable (it is not in this image: IOError).

Field names are accepted in either convention and translated (Coordinates/Position, Velocities/Velocity, ...).
Only reading lives here; turning the fields into the interpolation's inputs happens on the device
(``Spectra._device_particle_data`` -> ``fsb_prepare_particles``).
"""
import glob
import os

import numpy as np

from . import unitsystem

_REQUIRED = ("get_header_attr", "get_data", "get_n_segments", "get_smooth_length", "get_peculiar_velocity",
             "get_temp", "get_kernel", "get_npart", "get_omega_baryon", "get_units")

# the names that differ between the two on-disk conventions (abstractsnapshot.py:33-39)
HDF_TO_BIGFILE = {"Coordinates": "Position", "Velocities": "Velocity", "Masses": "Mass",
                  "NeutralHydrogenAbundance": "NeutralHydrogenFraction", "GFM_Metallicity": "Metallicity"}
BIGFILE_TO_HDF = {v: k for k, v in HDF_TO_BIGFILE.items()}


def is_snapshot(obj):
    return all(hasattr(obj, name) for name in _REQUIRED)


def AbstractSnapshotFactory(num, base, comm=None):
    """Same call as the reference's factory (abstractsnapshot.py:16-27); ``base`` may also be a snapshot object."""
    if is_snapshot(base):
        return base
    try:
        return HDF5Snapshot(num, base, comm)
    except IOError as hdf_error:
        try:
            return BigFileSnapshot(num, base, comm)
        except IOError as big_error:
            raise IOError("Not a bigfile or HDF5 snapshot: %s (%s; %s)" % (base, hdf_error, big_error)) from None


def _rank_size(comm):
    if comm is None:
        return 0, 1
    return comm.Get_rank(), comm.Get_size()


class AbstractSnapshot:
    """What both file formats share: derived quantities in terms of ``get_data`` / ``get_header_attr``."""

    def get_header_attr(self, attr):
        raise NotImplementedError

    def get_data(self, part_type, blockname, segment):
        raise NotImplementedError

    def get_n_segments(self, part_type=0):
        raise NotImplementedError

    def get_blocklen(self, part_type, blockname, segment):
        raise NotImplementedError

    def get_kernel(self):
        return 1

    def get_npart(self):
        return self.get_header_attr("TotNumPart")

    def get_omega_baryon(self):
        return self.get_header_attr("OmegaBaryon")

    def get_units(self):
        """Unit system from the header, the Gadget defaults when the header has none (abstractsnapshot.py:100-112)."""
        try:
            return unitsystem.UnitSystem(UnitLength_in_cm=self.get_header_attr("UnitLength_in_cm"),
                                         UnitMass_in_g=self.get_header_attr("UnitMass_in_g"),
                                         UnitVelocity_in_cm_per_s=self.get_header_attr("UnitVelocity_in_cm_per_s"))
        except KeyError:
            print("Warning: Using default (kpc,10^10Msun, km/s) units.")
            return unitsystem.UnitSystem()

    def get_smooth_length(self, part_type, segment):
        """Kernel support radius: half the Gadget smoothing length (abstractsnapshot.py:91-98)."""
        return self.get_data(part_type, "SmoothingLength", segment=segment) / 2

    def get_peculiar_velocity(self, part_type, segment):
        """Gadget's comoving velocities times sqrt(a) (abstractsnapshot.py:114-119)."""
        vel = self.get_data(part_type, "Velocities", segment=segment)
        vel *= np.sqrt(self.get_header_attr("Time"))
        return vel

    def get_temp(self, part_type, segment, hy_mass=0.76, units=None):
        """T = (gamma - 1) mu m_p u / k_B with mu = 4 / (X (3 + 4 n_e) + 1) (abstractsnapshot.py:121-154)."""
        if units is None:
            units = self.get_units()
        ienergy = self.get_data(part_type, "InternalEnergy", segment=segment) * units.UnitInternalEnergy_in_cgs
        nelec = self.get_data(part_type, "ElectronAbundance", segment=segment)
        muienergy = 4 / (hy_mass * (3 + 4 * nelec) + 1) * ienergy
        return (units.gamma - 1) * units.protonmass / units.boltzmann * muienergy


# ---- BigFile ---------------------------------------------------------------------------------------------------
def _parse_block_header(path):
    """(dtype, nmemb, [(file name, items)]) of a BigFile block directory."""
    try:
        with open(os.path.join(path, "header")) as f:
            lines = [ln.strip() for ln in f if ln.strip()]
    except OSError as exc:
        raise KeyError("Not found: " + path) from exc
    fields, files = {}, []
    for ln in lines:
        parts = [p.strip() for p in ln.split(":")]
        if parts[0] in ("DTYPE", "NMEMB", "NFILE"):
            fields[parts[0]] = parts[1]
        elif len(parts) >= 2:
            files.append((parts[0], int(parts[1])))
    if not {"DTYPE", "NMEMB", "NFILE"} <= set(fields):
        raise IOError("malformed BigFile header: " + path)
    if len(files) != int(fields["NFILE"]):
        raise IOError("BigFile header lists %d of %s files: %s" % (len(files), fields["NFILE"], path))
    return np.dtype(fields["DTYPE"]), int(fields["NMEMB"]), files


def _parse_attrs(path):
    """Attributes of a block: name -> numpy array (or str for character data), from the text file ``attr-v2``."""
    attrs = {}
    name = os.path.join(path, "attr-v2")
    if not os.path.exists(name):
        if os.path.exists(os.path.join(path, "attr")) and os.path.getsize(os.path.join(path, "attr")) > 0:
            raise IOError("binary BigFile attributes (format v1) are not supported: " + path)
        return attrs
    with open(name) as f:
        for ln in f:
            tok = ln.split(None, 4)
            if len(tok) < 4:
                continue
            key, dtype, nmemb, raw = tok[0], np.dtype(tok[1]), int(tok[2]), bytes.fromhex(tok[3])
            if dtype.kind == "S":
                attrs[key] = raw[:nmemb].decode("ascii", "replace").rstrip("\x00")
            else:
                attrs[key] = np.frombuffer(raw, dtype=dtype, count=nmemb).copy()
    return attrs


class BigFileBlock:
    """One column: ``block[start:end]`` reads the items from the data files they live in."""

    def __init__(self, path):
        self.path = path
        self.dtype, self.nmemb, self.files = _parse_block_header(path)
        self.size = sum(n for _, n in self.files)
        self.attrs = _parse_attrs(path)

    def __getitem__(self, key):
        if not isinstance(key, slice) or key.step not in (None, 1):
            raise TypeError("BigFile blocks are read by contiguous slices")
        start, end, _ = key.indices(self.size)
        shape = (max(end - start, 0),) + ((self.nmemb,) if self.nmemb > 1 else ())
        out = np.empty(shape, dtype=self.dtype.newbyteorder("="))
        first = 0  # index of the first item of the current file
        for fname, nitems in self.files:
            lo, hi = max(start, first), min(end, first + nitems)
            if lo < hi:
                part = np.fromfile(os.path.join(self.path, fname), dtype=self.dtype, count=(hi - lo) * self.nmemb,
                                   offset=(lo - first) * self.nmemb * self.dtype.itemsize)
                if part.size != (hi - lo) * self.nmemb:
                    raise IOError("short read from " + os.path.join(self.path, fname))
                out[lo - start:hi - start] = part.reshape((hi - lo,) + shape[1:])
            first += nitems
        return out


def write_bigfile_block(path, data, nfile=1, attrs=None):
    """Write ``data`` (items along axis 0) as a BigFile block in ``nfile`` data files, with optional attributes: the
    layout BigFileBlock reads.  Used for test fixtures and by ``synthetic`` to export snapshots."""
    os.makedirs(path, exist_ok=True)
    data = np.ascontiguousarray(data) if data is not None else np.zeros(0, dtype="<i8")
    nmemb = int(np.prod(data.shape[1:])) if data.ndim > 1 else 1
    dtype = data.dtype.newbyteorder("<")
    nfile = 0 if data.shape[0] == 0 else max(1, int(nfile))
    edges = np.linspace(0, data.shape[0], nfile + 1).astype(np.int64)
    with open(os.path.join(path, "header"), "w") as f:
        f.write("DTYPE: %s\nNMEMB: %d\nNFILE: %d\n" % (dtype.str, nmemb, nfile))
        for i in range(nfile):
            chunk = data[edges[i]:edges[i + 1]].astype(dtype, copy=False)
            chunk.tofile(os.path.join(path, "%06X" % i))
            checksum = int(np.frombuffer(chunk.tobytes(), dtype=np.uint8).sum(dtype=np.uint64) & 0xffffffff)
            f.write("%06X: %d : %d : 0\n" % (i, chunk.shape[0], checksum))
    with open(os.path.join(path, "attr-v2"), "w") as f:
        for key, value in (attrs or {}).items():
            if isinstance(value, str):
                raw = value.encode("ascii")
                f.write("%s |S1 %d %s #HUMANE [ %s ]\n" % (key, len(raw), raw.hex().upper(), value))
            else:
                arr = np.atleast_1d(np.asarray(value))
                arr = arr.astype(arr.dtype.newbyteorder("<"))
                f.write("%s %s %d %s #HUMANE [ %s ]\n" % (key, arr.dtype.str, arr.size, arr.tobytes().hex().upper(),
                                                        " ".join(str(v) for v in arr)))


class BigFileSnapshot(AbstractSnapshot):
    """MP-Gadget snapshot ``base/PART_<num>`` (or ``base`` itself): blocks ``<type>/<name>`` and a ``Header`` block
    that carries the attributes.  Particles are divided evenly over the ranks of ``comm`` and each rank's share into
    segments of about ``chunk_size`` particles, like the reference (abstractsnapshot.py:337-375)."""

    def __init__(self, num, base, comm=None):
        self.comm = comm
        self.rank, self.size = _rank_size(comm)
        self.parts_rank = None
        root = os.path.join(str(base), "PART_" + str(num).rjust(3, "0"))
        self.root = root if os.path.exists(root) else str(base)
        if not os.path.exists(os.path.join(self.root, "Header", "header")):
            raise IOError("No BigFile snapshot at " + root)
        self._header = BigFileBlock(os.path.join(self.root, "Header"))
        self._blocks = {}

    def _block(self, part_type, blockname):
        blockname = HDF_TO_BIGFILE.get(blockname, blockname)
        key = "%d/%s" % (part_type, blockname)
        if key not in self._blocks:
            self._blocks[key] = BigFileBlock(os.path.join(self.root, str(part_type), blockname))  # KeyError when absent
        return self._blocks[key]

    def get_header_attr(self, attr):
        value = self._header.attrs[attr]
        if isinstance(value, np.ndarray) and value.size == 1:
            return value[0]
        return value

    def get_n_segments(self, part_type=0, chunk_size=256. ** 3):
        npart = int(self.get_npart()[part_type])
        self.parts_rank = np.full(self.size, npart // self.size, dtype=np.int64)
        self.parts_rank[:npart % self.size] += 1
        return int(max(1, self.parts_rank[self.rank] / chunk_size))

    def _segment_to_partlist(self, part_type, segment):
        """First and one-past-last particle of a segment of this rank (None = to the end: segment < 0 is everything)."""
        if segment is None or segment < 0:
            return (0, None)
        nseg = self.get_n_segments(part_type)
        one = int(self.parts_rank[self.rank] / nseg)
        first = int(self.parts_rank[:self.rank].sum())
        return (first + one * segment, first + one * (segment + 1))

    def get_data(self, part_type, blockname, segment):
        start, end = self._segment_to_partlist(part_type, segment)
        return self._block(part_type, blockname)[start:end]

    def get_blocklen(self, part_type, blockname, segment):
        start, end = self._segment_to_partlist(part_type, segment)
        return (end if end is not None else self._block(part_type, blockname).size) - start

    def get_kernel(self):
        """MP-Gadget's DensityKernel: 1 cubic, 2 quintic (this library's id 3); cubic when absent (abstractsnapshot.py:377-397)."""
        try:
            kernel = int(self.get_header_attr("DensityKernel"))
        except KeyError:
            return 1
        kernel = 3 if kernel == 2 else kernel
        if kernel not in (1, 3):
            raise ValueError("unsupported DensityKernel %d" % kernel)
        return kernel

    def velocity_divisor(self):
        """Stored velocity / this = peculiar velocity (the device route's form of get_peculiar_velocity): a, or 1 when
        the header says UsePeculiarVelocity (abstractsnapshot.py:398-405)."""
        return 1.0 if self.get_header_attr("UsePeculiarVelocity") else self.get_header_attr("Time")

    def get_peculiar_velocity(self, part_type, segment):
        vel = self.get_data(part_type, "Velocity", segment=segment)
        if not self.get_header_attr("UsePeculiarVelocity"):
            vel /= self.get_header_attr("Time")
        return vel


# ---- HDF5 (through h5py, when it exists) ---------------------------------------------------------------------------
class HDF5Snapshot(AbstractSnapshot):
    """Gadget / Arepo HDF5 snapshot ``base/snapdir_<num>/snap_<num>.*.hdf5`` (abstractsnapshot.py:156-301); one segment
    per file; with ``comm`` the files are dealt out to the ranks."""

    def __init__(self, num, base, comm=None):
        try:
            import h5py
        except ImportError as exc:
            raise IOError("h5py is not available in this environment: HDF5 snapshots cannot be read") from exc
        self._h5py = h5py
        self.comm = comm
        rank, size = _rank_size(comm)
        snap = str(num).rjust(3, "0")
        where = str(base)
        if comm is None and os.path.exists(os.path.join(where, "snapdir_" + snap)):
            where = os.path.join(where, "snapdir_" + snap)
        names = sorted(glob.glob(os.path.join(where, "snap_" + snap + "*hdf5"))) or \
            sorted(glob.glob(os.path.join(where, "snapshot_" + snap + "*hdf5")))
        if not names:
            raise IOError("No files found")
        if comm is not None:
            per = len(names) // size
            mine = names[rank * per:(rank + 1) * per]
            if 1 <= rank <= len(names) - per * size:
                mine.append(names[per * size + rank - 1])
            names = mine
        self._files = [n for n in names if h5py.is_hdf5(n)][::-1]
        if not self._files:
            raise IOError("No HDF5 files found")
        self._handle_num = 0
        self._f_handle = h5py.File(self._files[0], "r")

    def __del__(self):
        try:
            self._f_handle.close()
        except Exception:  # noqa: BLE001 (interpreter shutdown)
            pass

    def _open(self, segment):
        if self._handle_num != segment:
            self._f_handle.close()
            self._f_handle = self._h5py.File(self._files[segment], "r")
            self._handle_num = segment
        return self._f_handle

    def get_header_attr(self, attr):
        value = self._f_handle["Header"].attrs[attr]
        if isinstance(value, np.ndarray) and value.size == 1:
            return value.reshape(-1)[0]
        return value

    def get_n_segments(self, part_type=None):
        return len(self._files)

    def get_data(self, part_type, blockname, segment):
        blockname = BIGFILE_TO_HDF.get(blockname, blockname)
        group = "PartType" + str(part_type)
        if segment is None or segment < 0:
            parts = []
            for name in self._files:
                with self._h5py.File(name, "r") as f:
                    parts.append(np.array(f[group][blockname]))
            return np.concatenate(parts)
        return np.array(self._open(segment)[group][blockname])

    def get_blocklen(self, part_type, blockname, segment):
        blockname = BIGFILE_TO_HDF.get(blockname, blockname)
        return self._open(segment)["PartType" + str(part_type)][blockname].len()

    def get_npart(self):
        return self.get_header_attr("NumPart_Total") + 2 ** 32 * self.get_header_attr("NumPart_Total_HighWord")

    def get_omega_baryon(self):
        mass_dm = self.get_header_attr("MassTable")[1] * self.get_header_attr("NumPart_ThisFile")[1]
        mass_bar = np.sum(self._f_handle["PartType0"]["Masses"])
        return mass_bar / (mass_bar + mass_dm) * self.get_header_attr("Omega0")

    def get_smooth_length(self, part_type, segment):
        """Volume^(1/3) for Arepo, SmoothingLength / 2 for Gadget, (Masses / Density)^(1/3) for recent Arepo
        (abstractsnapshot.py:253-282)."""
        try:
            return np.power(self.get_data(part_type, "Volume", segment=segment), 1. / 3)
        except KeyError:
            pass
        try:
            return self.get_data(part_type, "SmoothingLength", segment=segment) / 2
        except KeyError:
            volume = self.get_data(part_type, "Masses", segment=segment) / self.get_data(part_type, "Density", segment=segment)
            return np.power(volume, 1. / 3)

    def get_kernel(self):
        """0 (top hat) for Arepo, 1 (cubic spline) for Gadget (abstractsnapshot.py:284-301)."""
        keys = self._f_handle["PartType0"].keys()
        if "Volume" in keys:
            return 0
        return 1 if "SmoothingLength" in keys else 0
