"""CPU-side checks of the drop-in boundary: the C-ABI library loads here (no GPU), exports every
symbol declared in include/fsb200.h, and the Python mirror of the reference's native module
raises the reference's exceptions (py_module.cpp:122-151) before touching the device."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "fsb200.h")).read()
    return sorted(set(re.findall(r"FSB_API[^;(]*?\b(fsb_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for must in ("fsb_particle_interpolate", "fsb_particle_interpolate_host", "fsb_near_lines", "fsb_index_build",
                 "fsb_compute_tau", "fsb_compute_colden", "fsb_assign_cells"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from fake_spectra_b200 import _lib
    lib = _lib.load()
    syms = declared_symbols()
    assert set(syms) == set(_lib.SIGNATURES), "ctypes table and header disagree"
    for s in syms:
        assert getattr(lib, s) is not None
    assert lib.fsb_abi_version() == 1
    assert lib.fsb_strerror(_lib.FSB_EINVAL) == b"invalid argument"


def test_params_struct_layout():
    """sizeof(fsb_params) as laid out by the C compiler: 2*i32 + 8*f64 + 4*i32 = 88 bytes."""
    import ctypes as C
    from fake_spectra_b200 import _lib
    assert C.sizeof(_lib.Params) == 88
    assert _lib.Params.box.offset == 8 and _lib.Params.tautail.offset == 64 and _lib.Params.precision.offset == 72


def _args(n=10, nlos=3):
    pos = np.zeros((n, 3), np.float32)
    one = np.zeros(n, np.float32)
    return dict(pos=pos, vel=pos.copy(), dens=one, temp=one.copy(), h=one.copy(), axis=np.ones(nlos, np.int32),
                cofm=np.zeros((nlos, 3), np.float64))


def _call(priv, **over):
    a = _args()
    a.update(over)
    return priv._Particle_Interpolate(1, 100, 1, 1000., 0.1, 0.25, 1215e-8, 6e8, 0.4, 1.0, 1e-7, a["pos"], a["vel"],
                                      a["dens"], a["temp"], a["h"], a["axis"], a["cofm"])


def test_boundary_type_errors():
    from fake_spectra_b200 import _spectra_priv as priv
    with pytest.raises(TypeError):   # py_module.cpp:122-125
        _call(priv, pos=np.zeros((10, 3), np.float64))
    with pytest.raises(TypeError):   # :126-129
        _call(priv, cofm=np.zeros((3, 3), np.float32))
    with pytest.raises(TypeError):   # :130-133
        _call(priv, axis=np.ones(3, np.int64))
    with pytest.raises(ValueError):  # :141-145
        _call(priv, dens=np.zeros(9, np.float32))
    with pytest.raises(ValueError):  # :147-151
        _call(priv, axis=np.ones(4, np.int32))


def test_near_lines_type_errors():
    from fake_spectra_b200 import _spectra_priv as priv
    a = _args()
    with pytest.raises(TypeError):   # py_module.cpp:51-54
        priv._near_lines(1000., a["pos"].astype(np.float64), a["h"], a["axis"], a["cofm"])
    with pytest.raises(ValueError):  # :47-50
        priv._near_lines(1000., a["pos"], a["h"], a["axis"].astype(np.int64), a["cofm"])
    with pytest.raises(ValueError):  # :43-46
        priv._near_lines(1000., a["pos"], a["h"], np.ones(4, np.int32), a["cofm"])


def test_rescale_and_count_reject_wrong_dtypes_before_the_device():
    """_rescale_mean_flux raises the reference's TypeError (py_module.cpp:274-277) and _count_pairs the
    _near_lines errors, both before any device call."""
    from fake_spectra_b200 import _spectra_priv as priv
    with pytest.raises(TypeError):
        priv._rescale_mean_flux(np.zeros(4, np.float32), 0.5, 4, 1e-5, 1e30)
    with pytest.raises(TypeError):
        priv._count_pairs(1.0, np.zeros((2, 3)), np.zeros(2, np.float32), np.ones(1, np.int32), np.zeros((1, 3)))
    with pytest.raises(ValueError):
        priv._count_pairs(1.0, np.zeros((2, 3), np.float32), np.zeros(2, np.float32), np.ones(1, np.int64), np.zeros((1, 3)))
