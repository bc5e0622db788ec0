"""Drop-in for the reference's native module ``fake_spectra._spectra_priv`` (py_module.cpp:358-375)
on the hot path: ``_Particle_Interpolate`` and ``_near_lines`` with the same positional arguments,
dtype/shape checks and exceptions, implemented by the C-ABI host entry points of libfsb200.so
(host buffers in, host buffer out; the copies and the sm_100a kernels are inside the call).
"""
import ctypes as C

import numpy as np

from . import _lib

# module-level defaults that the host classes may override (Spectra(precision=..., voigt=...))
DEFAULT_PRECISION = _lib.PRECISION_FP64
DEFAULT_VOIGT = _lib.VOIGT_FAST


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _result_buffer(shape):
    """The array a call returns when the caller gave none.  Page-locked, from torch's caching host allocator: the
    device-to-host copy of a multi-GB result then runs at the copy engine's speed while the kernel is still computing,
    instead of the driver's staged path into fresh pageable pages (2.3x on the whole call at 2x512^3), and a dropped
    result's block is reused by the next call.  Plain numpy memory when torch is not importable."""
    try:
        import torch
        n = int(np.prod(shape))
        if n > 0 and torch.cuda.is_available():
            return torch.empty(n, dtype=torch.float64, pin_memory=True).numpy().reshape(shape)
    except ImportError:
        pass
    return np.empty(shape, dtype=np.float64)


def _is_f32(a):
    return isinstance(a, np.ndarray) and a.dtype == np.float32


def _Particle_Interpolate(compute_tau, nbins, kernel, box, velfac, atime, lambda_cm, gamma, fosc, amumass, tautail,
                          pos, vel, dens, temp, h, axis, cofm, precision=None, voigt=None, out=None,
                          extra_lines=(), extra_weights=(), seg_pairs=0, extra_ions=()):
    """Optical depth (compute_tau != 0) or column density on every sightline.

    Arguments, order and units as py_module.cpp:115; returns a new float64 array [NumLos, nbins].
    Raises TypeError / ValueError on the same conditions as py_module.cpp:122-151.

    Extensions (keyword only, absent from the reference): ``out`` = preallocated (e.g. pinned)
    float64 result buffer; ``extra_lines`` = [(lambda_cm, gamma, fosc), ...] further lines of the
    same ion computed from the same upload and candidate index, result [1+len, NumLos, nbins];
    ``extra_weights`` (column density only) = further float32 density-like arrays interpolated in
    the same geometry pass, result [1+len, NumLos, nbins]; ``seg_pairs`` = candidate pairs per work item
    (0 = automatic; see fsb_params.seg_pairs); ``extra_ions`` (optical depths only) = [(dens, amumass, [(lambda_cm,
    gamma, fosc), ...]), ...]: further ions of the same particles (their species densities, masses and lines) computed
    from the same upload and the same candidate index; their lines follow the first ion's in the result."""
    for a in (pos, vel, dens, temp, h):
        if not _is_f32(a):
            raise TypeError("One of the data arrays does not have 32-bit float type")
    if not (isinstance(cofm, np.ndarray) and cofm.dtype == np.float64):
        raise TypeError("Sightline positions must have 64-bit float type")
    if not (isinstance(axis, np.ndarray) and axis.dtype == np.int32):
        raise TypeError("Axis must be a 32-bit integer")
    numlos = cofm.shape[0]
    npart = pos.shape[0]
    if npart != dens.shape[-1] or npart != h.shape[0]:
        raise ValueError(" Dens, pos and h must have the same length")
    if cofm.ndim != 2 or numlos != axis.shape[0] or cofm.shape[1] != 3:
        raise ValueError("cofm must have dimensions (np.size(axis),3) ")
    if compute_tau and (vel.shape[0] != npart or temp.shape[0] != npart):
        raise ValueError(" Vel and temp must have the same length as pos when computing tau")
    if extra_weights:
        if compute_tau or extra_lines:
            raise ValueError("extra_weights only apply to column density")
        for w in extra_weights:
            if not _is_f32(w) or w.shape[0] != npart:
                raise TypeError("weight columns must be float32 arrays as long as pos")
        dens = np.stack([dens] + list(extra_weights))
    pos, dens, h = (np.ascontiguousarray(a) for a in (pos, dens, h))
    cofm, axis = np.ascontiguousarray(cofm), np.ascontiguousarray(axis)
    p = _lib.make_params(nbins, kernel, box, velfac, atime, lambda_cm, gamma, fosc, amumass, tautail,
                         precision=DEFAULT_PRECISION if precision is None else precision,
                         voigt=DEFAULT_VOIGT if voigt is None else voigt, seg_pairs=seg_pairs)
    prec = DEFAULT_PRECISION if precision is None else precision
    vgt = DEFAULT_VOIGT if voigt is None else voigt
    plist = [p] + [_lib.make_params(nbins, kernel, box, velfac, atime, lam, gam, fo, amumass, tautail, precision=prec,
                                    voigt=vgt, seg_pairs=seg_pairs) for (lam, gam, fo) in extra_lines]
    line_ion = [0] * len(plist)
    if extra_ions:
        if not compute_tau or extra_weights:
            raise ValueError("extra_ions only apply to optical depths")
        columns = [dens]
        for k, (idens, iamu, ilines) in enumerate(extra_ions):
            if not _is_f32(idens) or idens.shape != (npart,):
                raise TypeError("ion densities must be float32 arrays as long as pos")
            columns.append(idens)
            for (lam, gam, fo) in ilines:
                plist.append(_lib.make_params(nbins, kernel, box, velfac, atime, lam, gam, fo, iamu, tautail, precision=prec,
                                              voigt=vgt, seg_pairs=seg_pairs))
                line_ion.append(k + 1)
        columns = [np.ascontiguousarray(c) for c in columns]  # used where they lie: no stacked copy
    ncols = len(plist) if compute_tau or not extra_weights else 1 + len(extra_weights)
    shape = (numlos, int(nbins)) if ncols == 1 else (ncols, numlos, int(nbins))
    if out is None:
        out = _result_buffer(shape)
    elif out.dtype != np.float64 or out.size != int(np.prod(shape)) or not out.flags.c_contiguous:
        raise ValueError("out must be a C-contiguous float64 array of %s elements" % (shape,))
    parr = (_lib.Params * len(plist))(*plist)
    lib = _lib.load()
    if compute_tau:
        vel, temp = np.ascontiguousarray(vel), np.ascontiguousarray(temp)
        pvel, ptemp = _ptr(vel), _ptr(temp)
    else:
        pvel = ptemp = None
    if extra_ions:
        ions = np.asarray(line_ion, dtype=np.int32)
        cols = (C.c_void_p * len(columns))(*[c.ctypes.data for c in columns])
        rc = lib.fsb_particle_interpolate_ions_host(parr, ncols, _ptr(ions), len(columns), _ptr(pos), pvel, cols,
                                                    ptemp, _ptr(h), npart, _ptr(axis), _ptr(cofm), numlos, _ptr(out))
    else:
        rc = lib.fsb_particle_interpolate_multi_host(1 if compute_tau else 0, parr, ncols, _ptr(pos), pvel, _ptr(dens),
                                                     ptemp, _ptr(h), npart, _ptr(axis), _ptr(cofm), numlos, _ptr(out))
    _lib.check(rc, "_Particle_Interpolate")
    return out.reshape(shape)


def _near_lines(box, pos, hh, axis, cofm):
    """Sorted int32 indices of the particles whose kernel reaches at least one sightline
    (py_module.cpp:25-99, same checks)."""
    if not isinstance(cofm, np.ndarray) or cofm.ndim < 2 or not isinstance(axis, np.ndarray) or axis.ndim < 1:
        raise ValueError("cofm must have dimensions (np.size(axis),3) ")
    if cofm.shape[0] != axis.shape[0] or cofm.shape[1] != 3:
        raise ValueError("cofm must have dimensions (np.size(axis),3) ")
    if cofm.dtype != np.float64 or axis.dtype != np.int32:
        raise ValueError("cofm must have 64-bit float type and axis must be a 32-bit integer")
    if not _is_f32(pos) or not _is_f32(hh):
        raise TypeError("pos and h must have 32-bit float type")
    pos, hh = np.ascontiguousarray(pos), np.ascontiguousarray(hh)
    cofm, axis = np.ascontiguousarray(cofm), np.ascontiguousarray(axis)
    npart = pos.shape[0]
    out = np.empty(max(npart, 1), dtype=np.int32)
    count = C.c_int64(0)
    rc = _lib.load().fsb_near_lines_host(float(box), _ptr(pos), _ptr(hh), npart, _ptr(axis), _ptr(cofm), cofm.shape[0],
                                         _ptr(out), C.byref(count))
    _lib.check(rc, "_near_lines")
    return out[:count.value].copy()


def _rescale_mean_flux(tau, mean_flux_desired, nbins, tol, thresh):
    """Scale factor that brings the mean flux of ``tau`` (float64, first ``nbins`` elements) to
    ``mean_flux_desired`` (Py_mean_flux, py_module.cpp:263-282: same arguments, same TypeError)."""
    if not (isinstance(tau, np.ndarray) and tau.dtype == np.float64):
        raise TypeError("Optical depth must have 64-bit float type")
    from . import fluxstatistics
    return fluxstatistics.mean_flux(np.ravel(np.ascontiguousarray(tau))[:int(nbins)], mean_flux_desired, tol, thresh)


def _count_pairs(box, pos, hh, axis, cofm):
    """Candidate particles per sightline (int32 [NumLos]): the list sizes, without the lists.  Not part of the
    reference's module; used to balance sightline blocks across GPUs (sharding.balanced_blocks)."""
    if cofm.dtype != np.float64 or axis.dtype != np.int32:
        raise ValueError("cofm must have 64-bit float type and axis must be a 32-bit integer")
    if not _is_f32(pos) or not _is_f32(hh):
        raise TypeError("pos and h must have 32-bit float type")
    pos, hh = np.ascontiguousarray(pos), np.ascontiguousarray(hh)
    cofm, axis = np.ascontiguousarray(cofm), np.ascontiguousarray(axis)
    counts = np.zeros(cofm.shape[0], dtype=np.int32)
    rc = _lib.load().fsb_count_pairs_host(float(box), _ptr(pos), _ptr(hh), pos.shape[0], _ptr(axis), _ptr(cofm),
                                          cofm.shape[0], _ptr(counts))
    _lib.check(rc, "_count_pairs")
    return counts
