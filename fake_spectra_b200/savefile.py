"""Saved spectra (the reference's ``spectra.hdf5`` layout, spectra.py:266-372,434-499):

    Header.attrs{redshift,nbins,hubble,box,omegam,omegab,omegal,discarded,npart,Hz}
    spectra/{cofm,axis}   tau/<elem>/<ion>/<line>   colden/<elem>/<ion>   tau_obs/<elem>/<ion>
    velocity/<elem>/<ion>   temperature/<elem>/<ion>   density_weight_density/<elem>/<ion>
    num_important/<elem>/<ion>

Written with h5py when it is importable (byte-compatible with the reference's files); otherwise the
same tree is stored as a NumPy ``.npz`` archive with '/'-joined dataset names (this image has no
h5py).  Outside the interpolation hot path.
"""
import os
import shutil

import numpy as np

_GROUPS = (("tau_obs", "tau_obs"), ("tau", "tau"), ("colden", "colden"), ("velocity", "velocity"),
           ("temperature", "temp"), ("num_important", "num_important"), ("density_weight_density", "dens_weight_dens"))
_HEADER = (("redshift", "red"), ("nbins", "nbins"), ("hubble", "hubble"), ("box", "box"), ("omegam", "OmegaM"),
           ("omegab", "omegab"), ("omegal", "OmegaLambda"), ("discarded", "discarded"), ("npart", "npart"), ("Hz", "Hz"))


def _have_h5py():
    try:
        import h5py  # noqa: F401
        return True
    except ImportError:
        return False


def _flat(spec):
    """{dataset path: array} of everything a Spectra object saves."""
    out = {"spectra/cofm": spec.cofm, "spectra/axis": spec.axis}
    for grp, attr in _GROUPS:
        for key, value in getattr(spec, attr).items():
            if np.size(value) <= 1 and grp != "num_important":
                continue  # lazy placeholder that was never loaded
            out[grp + "/" + "/".join(str(k) for k in key)] = np.asarray(value)
    return out


def save(spec, savefile):
    """Write (with a .backup of any previous file, like the reference)."""
    d = os.path.dirname(savefile)
    if d and not os.path.exists(d):
        os.makedirs(d)
    if os.path.exists(savefile):
        shutil.move(savefile, savefile + ".backup")
    header = {name: getattr(spec, attr) for name, attr in _HEADER}
    data = _flat(spec)
    if _have_h5py() and not savefile.endswith(".npz"):
        import h5py
        with h5py.File(savefile, "w") as f:
            grp = f.create_group("Header")
            for k, v in header.items():
                grp.attrs[k] = v
            for g, _ in _GROUPS:
                f.require_group(g)
            for name, arr in data.items():
                f.create_dataset(name, data=arr)
        return savefile
    arrays = {"Header/" + k: np.asarray(v) for k, v in header.items()}
    arrays.update(data)
    with open(savefile, "wb") as fh:
        np.savez(fh, **arrays)
    return savefile


def _open_npz(savefile):
    try:
        return np.load(savefile, allow_pickle=False)
    except (IOError, OSError, ValueError) as io:
        raise IOError("Could not read saved data from: " + str(savefile) +
                      ". If the file does not exist, try using reload_file=True") from io


def load(spec, savefile):
    """Header, sightlines and lazy placeholders for every saved array (spectra.py:434-499)."""
    if _have_h5py() and not str(savefile).endswith(".npz"):
        import h5py
        try:
            f = h5py.File(savefile, "r")
        except IOError as io:
            raise IOError("Could not read saved data from: " + str(savefile) +
                          ". If the file does not exist, try using reload_file=True") from io
        with f:
            header = dict(f["Header"].attrs)
            names = []
            f.visititems(lambda n, o: names.append(n) if isinstance(o, h5py.Dataset) else None)
            cofm, axis = np.array(f["spectra/cofm"]), np.array(f["spectra/axis"])
            numimp = {n: np.array(f[n]) for n in names if n.startswith("num_important/")}
    else:
        z = _open_npz(savefile)
        header = {k[len("Header/"):]: z[k][()] for k in z.files if k.startswith("Header/")}
        names = [k for k in z.files if not k.startswith("Header/")]
        cofm, axis = z["spectra/cofm"], z["spectra/axis"]
        numimp = {n: z[n] for n in names if n.startswith("num_important/")}
    spec.red = header["redshift"]
    spec.atime = 1. / (1 + spec.red)
    spec.OmegaM = header["omegam"]
    spec.nbins = int(header["nbins"])
    spec.omegab = header["omegab"]
    spec.OmegaLambda = header["omegal"]
    spec.hubble = header["hubble"]
    spec.npart = np.array(header["npart"])
    spec.box = header["box"]
    spec.discarded = header["discarded"]
    spec.Hz = header.get("Hz", None)
    spec.cofm, spec.axis = np.array(cofm), np.array(axis)
    lookup = dict(_GROUPS)
    for name in names:
        parts = name.split("/")
        if parts[0] not in lookup or parts[0] == "num_important":
            continue
        key = (parts[1], int(parts[2])) if len(parts) == 3 else (parts[1], int(parts[2]), int(float(parts[3])))
        getattr(spec, lookup[parts[0]])[key] = np.array([0])  # placeholder: loaded on first use
    for name, arr in numimp.items():
        parts = name.split("/")
        spec.num_important[(parts[1], int(parts[2]))] = arr


def load_array(savefile, array_name, key):
    """One saved array (``array_name`` is the group name on disc)."""
    name = array_name + "/" + "/".join(str(k) for k in key)
    if _have_h5py() and not str(savefile).endswith(".npz"):
        import h5py
        with h5py.File(savefile, "r") as f:
            return np.array(f[name])
    z = _open_npz(savefile)
    if name not in z.files:
        raise KeyError(name)
    return z[name]


# ---- the binary spectra file of the reference's C extractor (cextract/main.cpp:247-257, read back by
# cextract/statistic.c:116-157): a 128-byte header -- redshift f64, box (kpc/h) f64, nbins i32, NumLos i32 and 26 int32
# of padding -- then the H I optical depths and the H I column densities, float64 [NumLos][nbins] each, native byte order.
CEXTRACT_HEADER_BYTES = 128
_CEXTRACT_HEADER = np.dtype([("redshift", "<f8"), ("box", "<f8"), ("nbins", "<i4"), ("numlos", "<i4"), ("pad", "<i4", (26,))])
assert _CEXTRACT_HEADER.itemsize == CEXTRACT_HEADER_BYTES


def write_cextract(path, redshift, box, tau, colden):
    """Write ``tau`` and ``colden`` (float64 [NumLos, nbins]) in the C extractor's ``*_spectra.dat`` layout."""
    tau = np.ascontiguousarray(tau, dtype="<f8")
    colden = np.ascontiguousarray(colden, dtype="<f8")
    if tau.ndim != 2 or tau.shape != colden.shape:
        raise ValueError("tau and colden must both have shape (NumLos, nbins)")
    head = np.zeros(1, dtype=_CEXTRACT_HEADER)
    head["redshift"], head["box"], head["nbins"], head["numlos"] = redshift, box, tau.shape[1], tau.shape[0]
    with open(path, "wb") as f:
        head.tofile(f)
        tau.tofile(f)
        colden.tofile(f)


def read_cextract(path):
    """(redshift, box, tau, colden) from a ``*_spectra.dat`` file; the column densities are None for files that stop
    after the optical depths."""
    with open(path, "rb") as f:
        head = np.fromfile(f, dtype=_CEXTRACT_HEADER, count=1)
        if head.size != 1:
            raise IOError("%s is shorter than the %d-byte header" % (path, CEXTRACT_HEADER_BYTES))
        nbins, numlos = int(head["nbins"][0]), int(head["numlos"][0])
        if nbins <= 0 or numlos < 0:
            raise IOError("%s: bad header (nbins %d, NumLos %d)" % (path, nbins, numlos))
        tau = np.fromfile(f, dtype="<f8", count=nbins * numlos)
        if tau.size != nbins * numlos:
            raise IOError("%s: optical depths truncated" % path)
        colden = np.fromfile(f, dtype="<f8", count=nbins * numlos)
    colden = colden.reshape(numlos, nbins) if colden.size == nbins * numlos else None
    return float(head["redshift"][0]), float(head["box"][0]), tau.reshape(numlos, nbins), colden
