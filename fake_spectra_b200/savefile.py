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
