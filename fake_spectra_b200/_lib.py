"""ctypes binding of libfsb200.so (the C ABI declared in include/fsb200.h).

There is no fallback: if the library has not been built, cannot be loaded, or a call fails, an
exception is raised.  Nothing in this package imports the CPU oracle.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FSB200_LIB") or os.path.join(HERE, "libfsb200.so")  # FSB200_LIB: tuning variants

FSB_OK = 0
FSB_EINVAL, FSB_ECUDA, FSB_ENOMEM, FSB_EVORONOI, FSB_ENODEV = -1, -2, -3, -4, -5
KERNEL_TOPHAT, KERNEL_CUBIC, KERNEL_VORONOI, KERNEL_QUINTIC = 0, 1, 2, 3
PRECISION_FP64, PRECISION_FP32 = 0, 1
VOIGT_FAST, VOIGT_EXACT = 0, 1


class Params(C.Structure):
    """struct fsb_params."""
    _fields_ = [("nbins", C.c_int32), ("kernel", C.c_int32), ("box", C.c_double), ("velfac", C.c_double),
                ("atime", C.c_double), ("lambda_cm", C.c_double), ("gamma", C.c_double), ("fosc", C.c_double),
                ("amumass", C.c_double), ("tautail", C.c_double), ("precision", C.c_int32), ("voigt", C.c_int32),
                ("seg_pairs", C.c_int32), ("reserved", C.c_int32)]


class Prep(C.Structure):
    """struct fsb_prep."""
    _fields_ = [(n, C.c_float) for n in ("dens_conv", "rscale", "hy_mass",
                                         "nelec_const", "mass_frac_const", "amumass", "dens_thresh_code")] + \
               [(n, C.c_int32) for n in ("velocity_divides", "neutral_hydrogen", "sf_neutral", "redshift_coverage")] + \
               [(n, C.c_double) for n in ("gray_opac", "gamma_uvb", "f_bar", "unit_ienergy", "temp_factor")] + \
               [("temp_double", C.c_int32), ("reserved", C.c_int32), ("velocity_factor", C.c_double)]


class IonTable(C.Structure):
    """struct fsb_ion_table."""
    _fields_ = [("coef", C.c_void_p)] + [(n, C.c_int32) for n in ("nd", "nt", "pad", "reserved")] + \
               [(n, C.c_double) for n in ("dens0", "dens_span", "temp0", "temp_span")] + \
               [(n, C.c_float) for n in ("dens_lo", "dens_hi", "temp_lo", "temp_hi", "rho_factor", "reserved2")]


MAX_PEERS = 16


class Push(C.Structure):
    """struct fsb_push."""
    _fields_ = [("npeers", C.c_int32), ("reserved", C.c_int32), ("line_stride", C.c_int64), ("dest", C.c_void_p * MAX_PEERS)]


class FsbError(RuntimeError):
    def __init__(self, code, what, detail):
        super().__init__("libfsb200: %s (%d): %s" % (what, code, detail))
        self.code = code


_P = C.c_void_p
# name -> (restype, argtypes); every symbol declared in include/fsb200.h
SIGNATURES = {
    "fsb_abi_version": (C.c_int, []),
    "fsb_strerror": (C.c_char_p, [C.c_int]),
    "fsb_last_error": (C.c_char_p, []),
    "fsb_kernel_launches": (C.c_uint64, []),
    "fsb_device_info": (C.c_int, [C.POINTER(C.c_int32)] * 4),
    "fsb_index_build": (C.c_int, [C.c_double, _P, _P, C.c_int32, _P, _P, C.c_int64, _P, C.POINTER(_P)]),
    "fsb_index_build_counted": (C.c_int, [C.c_double, _P, _P, C.c_int32, _P, _P, C.c_int64, _P, _P, C.POINTER(_P)]),
    "fsb_index_free": (C.c_int, [_P, _P]),
    "fsb_index_sizes": (C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "fsb_index_export": (C.c_int, [_P, _P, _P, _P, _P]),
    "fsb_compute_tau": (C.c_int, [_P, C.POINTER(Params), _P, _P, _P, _P, _P, _P, _P, _P]),
    "fsb_compute_tau_multi": (C.c_int, [_P, C.POINTER(Params), C.c_int32, _P, _P, _P, _P, _P, _P, _P, _P]),
    "fsb_compute_tau_multi_push": (C.c_int, [_P, C.POINTER(Params), C.c_int32, _P, _P, _P, _P, _P, _P, C.POINTER(Push), _P]),
    "fsb_compute_tau_multi_range": (C.c_int, [_P, C.POINTER(Params), C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P, _P, _P, _P]),
    "fsb_peer_alloc": (C.c_int, [C.c_int64, C.POINTER(_P), _P]),
    "fsb_peer_free": (C.c_int, [_P]),
    "fsb_peer_open": (C.c_int, [_P, C.POINTER(_P)]),
    "fsb_peer_close": (C.c_int, [_P]),
    "fsb_compute_colden": (C.c_int, [_P, C.POINTER(Params), _P, _P, C.c_int32, _P, _P, _P, _P]),
    "fsb_particle_interpolate": (C.c_int, [C.c_int32, C.POINTER(Params), _P, _P, _P, _P, _P, C.c_int64, _P, _P,
                                           C.c_int32, _P, _P]),
    "fsb_particle_interpolate_host": (C.c_int, [C.c_int32, C.POINTER(Params), _P, _P, _P, _P, _P, C.c_int64, _P, _P,
                                                C.c_int32, _P]),
    "fsb_particle_interpolate_multi_host": (C.c_int, [C.c_int32, C.POINTER(Params), C.c_int32, _P, _P, _P, _P, _P,
                                                      C.c_int64, _P, _P, C.c_int32, _P]),
    "fsb_particle_interpolate_ions_host": (C.c_int, [C.POINTER(Params), C.c_int32, _P, C.c_int32, _P, _P, _P, _P, _P,
                                                     C.c_int64, _P, _P, C.c_int32, _P]),
    "fsb_near_lines": (C.c_int, [C.c_double, _P, _P, C.c_int64, _P, _P, C.c_int32, _P, C.POINTER(C.c_int64), _P]),
    "fsb_near_lines_host": (C.c_int, [C.c_double, _P, _P, C.c_int64, _P, _P, C.c_int32, _P, C.POINTER(C.c_int64)]),
    "fsb_count_pairs": (C.c_int, [C.c_double, _P, _P, C.c_int64, _P, _P, C.c_int32, _P, _P]),
    "fsb_count_pairs_host": (C.c_int, [C.c_double, _P, _P, C.c_int64, _P, _P, C.c_int32, _P]),
    "fsb_assign_cells": (C.c_int, [_P, C.c_double, _P, _P, _P, _P, _P]),
    "fsb_measure_fma_peak": (C.c_int, [C.c_int32, C.POINTER(C.c_double), _P]),
    "fsb_voigt_profile": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int32, _P]),
    "fsb_prepare_particles": (C.c_int, [C.POINTER(Prep), _P, C.c_int64, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int64,
                                        C.POINTER(IonTable), _P, _P, _P, _P, _P, _P]),
    "fsb_smoothing_lengths": (C.c_int, [_P, _P, C.c_int64, C.c_int32, _P, _P]),
    "fsb_prepare_select": (C.c_int, [C.POINTER(Prep), _P, C.c_int64, _P, _P, C.c_int64, _P, C.POINTER(C.c_int64), _P]),
    "fsb_rescale_mean_flux": (C.c_int, [_P, C.c_int64, C.c_double, C.c_double, C.c_double, C.POINTER(C.c_double),
                                        C.POINTER(C.c_int32), _P]),
    "fsb_flux_power": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_double, _P, _P, _P]),
    "fsb_row_max": (C.c_int, [_P, C.c_int64, C.c_int64, _P, _P]),
    "fsb_flux_sums": (C.c_int, [_P, C.c_int64, C.c_double, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                C.POINTER(C.c_int64), _P]),
    "fsb_flux_pdf": (C.c_int, [_P, C.c_int64, C.c_double, C.c_int32, _P, _P]),
    "fsb_delta_flux": (C.c_int, [_P, C.c_int64, C.c_double, C.c_double, _P, _P]),
    "fsb_power_accumulate": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_double, _P, _P]),
}

_lib = None


def load():
    """Load libfsb200.so and bind every declared symbol; raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s has not been built: run `python -m fake_spectra_b200.build` "
                          "(or __graft_entry__.build()); there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.fsb_abi_version() != 1:
        raise ImportError("libfsb200.so ABI version %d, expected 1" % lib.fsb_abi_version())
    _lib = lib
    return lib


def check(rc, what):
    if rc != FSB_OK:
        lib = load()
        raise FsbError(rc, what + ": " + lib.fsb_strerror(rc).decode(), lib.fsb_last_error().decode())


def make_params(nbins, kernel, box, velfac, atime, lambda_cm, gamma, fosc, amumass, tautail,
                precision=PRECISION_FP64, voigt=VOIGT_FAST, seg_pairs=0):
    return Params(int(nbins), int(kernel), float(box), float(velfac), float(atime), float(lambda_cm), float(gamma),
                  float(fosc), float(amumass), float(tautail), int(precision), int(voigt), int(seg_pairs), 0)
