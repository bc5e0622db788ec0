"""TEST INFRASTRUCTURE — ctypes front-ends for the two CPU checkers.

* :class:`Oracle`    — the C restatement (oracle/fs_oracle.c -> oracle/_build/libfsoracle.so)
* :class:`Reference` — the unmodified reference (oracle/_ref/libfsref.so), when it was built.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this package.  ``fake_spectra_b200`` (the product) must never import it.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")

_LINE_ARGS = [C.c_double] * 7 + [C.c_int, C.c_double]  # lambda gamma fosc amumass velfac box atime kernel tautail


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class _Base:
    """Shared call conventions; subclasses bind the symbol names of their library."""

    prefix = None

    def __init__(self, path):
        self.path = path
        self.lib = C.CDLL(path)
        p = self.prefix
        lib = self.lib
        compute_args = [C.c_int] + [C.c_double] * 7 + [_f64p, _i32p, C.c_int, C.c_int, C.c_double, _f64p]
        self._tau = getattr(lib, p + "compute_tau")
        self._tau.argtypes = compute_args + [_f32p, _f32p, _f32p, _f32p, _f32p, C.c_longlong]
        self._colden = getattr(lib, p + "compute_colden")
        self._colden.argtypes = compute_args + [_f32p, _f32p, _f32p, C.c_longlong]
        self._addc = getattr(lib, p + "add_colden_particle")
        self._addc.argtypes = _LINE_ARGS + [_f64p, C.c_int, C.c_double, C.c_float, C.c_float, C.c_float]
        self._addc.restype = None
        self._addt = getattr(lib, p + "add_tau_particle")
        self._addt.argtypes = _LINE_ARGS + [_f64p, C.c_int, C.c_double, C.c_float, C.c_float, C.c_float,
                                            C.c_float, C.c_float]
        self._addt.restype = None
        self._kfrac = getattr(lib, p + "kern_frac")
        self._kfrac.argtypes = [C.c_int] + [C.c_double] * 5
        self._kfrac.restype = C.c_double

    # -- drivers -------------------------------------------------------------------------
    def compute_tau(self, nbins, kernel, box, velfac, atime, lambda_cm, gamma, fosc, amumass, tautail,
                    pos, vel, dens, temp, h, axis, cofm, out=None):
        """Same argument order as the reference's _Particle_Interpolate (py_module.cpp:115)."""
        cofm, axis = _f64(cofm), _i32(axis)
        nlos = cofm.shape[0]
        if out is None:
            out = np.zeros((nlos, nbins), dtype=np.float64)
        pos = _f32(pos)
        self.last_seconds = self._call_tau(nbins, lambda_cm, gamma, fosc, amumass, box, velfac, atime, cofm, axis,
                                           nlos, kernel, tautail, out, pos, _f32(vel), _f32(dens), _f32(temp),
                                           _f32(h), pos.shape[0])
        return out

    def compute_colden(self, nbins, kernel, box, velfac, atime, lambda_cm, gamma, fosc, amumass, tautail,
                       pos, dens, h, axis, cofm, out=None):
        cofm, axis = _f64(cofm), _i32(axis)
        nlos = cofm.shape[0]
        if out is None:
            out = np.zeros((nlos, nbins), dtype=np.float64)
        pos = _f32(pos)
        self.last_seconds = self._call_colden(nbins, lambda_cm, gamma, fosc, amumass, box, velfac, atime, cofm, axis,
                                              nlos, kernel, tautail, out, pos, _f32(dens), _f32(h), pos.shape[0])
        return out

    # -- single-particle accumulators ----------------------------------------------------
    def add_colden_particle(self, line, row, dr2, dens, ppos, smooth):
        """line = (lambda_cm, gamma, fosc, amumass, velfac, box, atime, kernel, tautail)."""
        self._addc(*line, row, row.shape[0], dr2, dens, ppos, smooth)

    def add_tau_particle(self, line, row, dr2, dens, ppos, pvel, temp, smooth):
        self._addt(*line, row, row.shape[0], dr2, dens, ppos, pvel, temp, smooth)

    def kern_frac(self, kernel, zlow, zhigh, smooth, dr2, zrange):
        return self._kfrac(kernel, zlow, zhigh, smooth, dr2, zrange)


class Oracle(_Base):
    """The C restatement."""

    prefix = "fso_"

    def __init__(self):
        super().__init__(_build.build_oracle())
        lib = self.lib
        self._tau.restype = C.c_int
        self._colden.restype = C.c_int
        lib.fso_faddeeva_re.argtypes = [C.c_double, C.c_double]
        lib.fso_faddeeva_re.restype = C.c_double
        lib.fso_faddeeva_re_many.argtypes = [_f64p, _f64p, _f64p, C.c_longlong]
        lib.fso_erfcx.argtypes = [C.c_double]
        lib.fso_erfcx.restype = C.c_double
        for name in ("fso_cubic_kernel", "fso_quintic_kernel"):
            getattr(lib, name).argtypes = [C.c_double]
            getattr(lib, name).restype = C.c_double
        lib.fso_tau_kern_outer_pub.argtypes = [C.c_double] * 4 + [C.c_int, C.c_double, C.c_double]
        lib.fso_tau_kern_outer_pub.restype = C.c_double
        lib.fso_near_particles.argtypes = [_f64p, _i32p, C.c_int, C.c_double, _f32p, _f32p, C.c_longlong, _i64p,
                                           C.c_void_p, C.c_void_p]
        lib.fso_near_particles.restype = C.c_longlong
        lib.fso_near_lines.argtypes = [C.c_double, _f32p, _f32p, C.c_longlong, _i32p, _f64p, C.c_int, C.c_void_p]
        lib.fso_near_lines.restype = C.c_longlong
        lib.fso_assign_cells.argtypes = [_f64p, _i32p, C.c_double, C.c_int, _i32p, C.c_int, _f32p, _f32p]
        lib.fso_assign_cells.restype = C.c_int
        lib.fso_omp_max_threads.restype = C.c_int
        lib.fso_mean_flux_scale.argtypes = [_f64p, C.c_double, C.c_longlong, C.c_double, C.c_double, C.POINTER(C.c_int)]
        lib.fso_mean_flux_scale.restype = C.c_double

    def _call_tau(self, *a):
        import time
        t0 = time.perf_counter()
        err = self._tau(*a)
        if err:
            raise RuntimeError("oracle: Voronoi cell invariant violated (reference would exit(1))")
        return time.perf_counter() - t0

    def _call_colden(self, *a):
        import time
        t0 = time.perf_counter()
        err = self._colden(*a)
        if err:
            raise RuntimeError("oracle: Voronoi cell invariant violated (reference would exit(1))")
        return time.perf_counter() - t0

    def threads(self):
        return self.lib.fso_omp_max_threads()

    def set_threads(self, n):
        self.lib.fso_omp_set_threads(int(n))

    def profile(self, uu, aa):
        uu = _f64(np.atleast_1d(uu))
        aa = _f64(np.broadcast_to(np.atleast_1d(aa), uu.shape))
        out = np.empty_like(uu)
        self.lib.fso_faddeeva_re_many(uu, aa, out, uu.size)
        return out

    def erfcx(self, y):
        return self.lib.fso_erfcx(y)

    def cubic_kernel(self, q):
        return self.lib.fso_cubic_kernel(q)

    def quintic_kernel(self, q):
        return self.lib.fso_quintic_kernel(q)

    def tau_kern_outer(self, btherm, vdr2, vsmooth, aa, kernel, vlow, vhigh):
        return self.lib.fso_tau_kern_outer_pub(btherm, vdr2, vsmooth, aa, kernel, vlow, vhigh)

    def mean_flux_scale(self, tau, mean_flux_desired, tol=1e-5, thresh=1e30, return_iterations=False):
        """get_mean_flux_scale, py_module.cpp:235-262 (0 for an empty array: fluxstatistics.py:39-40)."""
        tau = _f64(np.ravel(tau))
        if tau.size == 0:
            return (0.0, 0) if return_iterations else 0.0
        it = C.c_int(0)
        s = float(self.lib.fso_mean_flux_scale(tau, float(mean_flux_desired), tau.size, float(tol), float(thresh), C.byref(it)))
        return (s, it.value) if return_iterations else s

    def near_particles(self, cofm, axis, box, pos, h):
        """-> (offsets int64[nlos+1], particle int32[npairs], dr2 float64[npairs])."""
        cofm, axis, pos, h = _f64(cofm), _i32(axis), _f32(pos), _f32(h)
        nlos = cofm.shape[0]
        counts = np.zeros(max(nlos, 1), dtype=np.int64)
        total = self.lib.fso_near_particles(cofm, axis, nlos, box, pos, h, pos.shape[0], counts, None, None)
        part = np.zeros(max(total, 1), dtype=np.int32)
        dr2 = np.zeros(max(total, 1), dtype=np.float64)
        self.lib.fso_near_particles(cofm, axis, nlos, box, pos, h, pos.shape[0], counts,
                                    part.ctypes.data_as(C.c_void_p), dr2.ctypes.data_as(C.c_void_p))
        offsets = np.zeros(nlos + 1, dtype=np.int64)
        np.cumsum(counts[:nlos], out=offsets[1:])
        return offsets, part[:total], dr2[:total]

    def near_lines(self, box, pos, h, axis, cofm):
        cofm, axis, pos, h = _f64(cofm), _i32(axis), _f32(pos), _f32(h)
        out = np.zeros(max(pos.shape[0], 1), dtype=np.int32)
        n = self.lib.fso_near_lines(box, pos, h, pos.shape[0], axis, cofm, cofm.shape[0],
                                    out.ctypes.data_as(C.c_void_p))
        return out[:n].copy()

    def assign_cells(self, cofm, axis, box, line, pos, h):
        cofm, axis, pos, h = _f64(cofm), _i32(axis), _f32(pos), _f32(h)
        offsets, part, _ = self.near_particles(cofm, axis, box, pos, h)
        cand = np.ascontiguousarray(part[offsets[line]:offsets[line + 1]])
        arr = np.zeros(2 * max(cand.size, 1), dtype=np.float32)
        err = self.lib.fso_assign_cells(cofm, axis, box, line, cand if cand.size else np.zeros(1, np.int32),
                                        cand.size, pos, arr)
        return err, arr[:2 * cand.size]


class Reference(_Base):
    """The unmodified reference C++ (oracle/_ref/libfsref.so)."""

    prefix = "ref_"

    @staticmethod
    def available():
        p = _build.ref_lib_path()
        return os.path.exists(p) or _build.have_reference_sources()

    def __init__(self):
        path = _build.build_ref()
        if path is None or not os.path.exists(path):
            raise FileNotFoundError("oracle/_ref/libfsref.so is absent and /root/reference is not mounted")
        super().__init__(path)
        lib = self.lib
        self._tau.restype = C.c_double
        self._colden.restype = C.c_double
        lib.ref_profile.argtypes = [C.c_double, C.c_double]
        lib.ref_profile.restype = C.c_double
        lib.ref_profile_many.argtypes = [_f64p, _f64p, _f64p, C.c_longlong]
        lib.ref_faddeeva_w.argtypes = [C.c_double, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        for name in ("ref_sph_cubic_kernel", "ref_sph_quintic_kernel"):
            getattr(lib, name).argtypes = [C.c_double]
            getattr(lib, name).restype = C.c_double
        lib.ref_tau_kern_outer.argtypes = [C.c_double] * 4 + [C.c_int, C.c_double, C.c_double]
        lib.ref_tau_kern_outer.restype = C.c_double
        lib.ref_near_particles.argtypes = [_f64p, _i32p, C.c_int, C.c_double, _f32p, _f32p, C.c_longlong, _i64p,
                                           C.c_void_p, C.c_void_p, C.c_longlong]
        lib.ref_near_particles.restype = C.c_double
        lib.ref_get_near_lines.argtypes = [_f64p, _i32p, C.c_int, C.c_double, _f32p, C.c_float, _i32p, _f64p, C.c_int]
        lib.ref_get_near_lines.restype = C.c_int
        lib.ref_near_lines.argtypes = [C.c_double, _f32p, _f32p, C.c_longlong, _i32p, _f64p, C.c_int, C.c_void_p,
                                       C.c_longlong]
        lib.ref_near_lines.restype = C.c_longlong
        lib.ref_assign_cells.argtypes = [_f64p, _i32p, C.c_int, C.c_double, C.c_int, _f32p, _f32p, C.c_longlong,
                                         _f32p, C.c_int]
        lib.ref_assign_cells.restype = C.c_int
        lib.ref_omp_max_threads.restype = C.c_int

    def _call_tau(self, *a):
        return self._tau(*a)

    def _call_colden(self, *a):
        return self._colden(*a)

    def threads(self):
        return self.lib.ref_omp_max_threads()

    def set_threads(self, n):
        self.lib.ref_omp_set_threads(int(n))

    def profile(self, uu, aa):
        uu = _f64(np.atleast_1d(uu))
        aa = _f64(np.broadcast_to(np.atleast_1d(aa), uu.shape))
        out = np.empty_like(uu)
        self.lib.ref_profile_many(uu, aa, out, uu.size)
        return out

    def faddeeva_w(self, x, y):
        re, im = C.c_double(), C.c_double()
        self.lib.ref_faddeeva_w(x, y, C.byref(re), C.byref(im))
        return complex(re.value, im.value)

    def cubic_kernel(self, q):
        return self.lib.ref_sph_cubic_kernel(q)

    def quintic_kernel(self, q):
        return self.lib.ref_sph_quintic_kernel(q)

    def tau_kern_outer(self, btherm, vdr2, vsmooth, aa, kernel, vlow, vhigh):
        return self.lib.ref_tau_kern_outer(btherm, vdr2, vsmooth, aa, kernel, vlow, vhigh)

    def near_particles(self, cofm, axis, box, pos, h):
        cofm, axis, pos, h = _f64(cofm), _i32(axis), _f32(pos), _f32(h)
        nlos = cofm.shape[0]
        counts = np.zeros(max(nlos, 1), dtype=np.int64)
        self.lib.ref_near_particles(cofm, axis, nlos, box, pos, h, pos.shape[0], counts, None, None, 0)
        total = int(counts[:nlos].sum())
        part = np.zeros(max(total, 1), dtype=np.int32)
        dr2 = np.zeros(max(total, 1), dtype=np.float64)
        self.last_seconds = self.lib.ref_near_particles(cofm, axis, nlos, box, pos, h, pos.shape[0], counts,
                                                        part.ctypes.data_as(C.c_void_p),
                                                        dr2.ctypes.data_as(C.c_void_p), total)
        offsets = np.zeros(nlos + 1, dtype=np.int64)
        np.cumsum(counts[:nlos], out=offsets[1:])
        return offsets, part[:total], dr2[:total]

    def get_near_lines(self, cofm, axis, box, pos3, hh):
        cofm, axis = _f64(cofm), _i32(axis)
        nlos = cofm.shape[0]
        lines = np.zeros(max(nlos, 1), dtype=np.int32)
        dr2 = np.zeros(max(nlos, 1), dtype=np.float64)
        n = self.lib.ref_get_near_lines(cofm, axis, nlos, box, _f32(pos3), hh, lines, dr2, nlos)
        return lines[:n].copy(), dr2[:n].copy()

    def near_lines(self, box, pos, h, axis, cofm):
        cofm, axis, pos, h = _f64(cofm), _i32(axis), _f32(pos), _f32(h)
        out = np.zeros(max(pos.shape[0], 1), dtype=np.int32)
        n = self.lib.ref_near_lines(box, pos, h, pos.shape[0], axis, cofm, cofm.shape[0],
                                    out.ctypes.data_as(C.c_void_p), out.shape[0])
        return out[:n].copy()

    def assign_cells(self, cofm, axis, box, line, pos, h):
        cofm, axis, pos, h = _f64(cofm), _i32(axis), _f32(pos), _f32(h)
        cap = 2 * pos.shape[0] + 2
        arr = np.zeros(cap, dtype=np.float32)
        n = self.lib.ref_assign_cells(cofm, axis, cofm.shape[0], box, line, pos, h, pos.shape[0], arr, cap)
        return 0, arr[:2 * n].copy()
