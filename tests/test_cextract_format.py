"""The 128-byte-header binary of the reference's C extractor (cextract/main.cpp:247-257, read by statistic.c:116-157),
byte for byte against the layout those fwrite calls produce on a little-endian machine."""
import struct

import numpy as np
import pytest

from fake_spectra_b200 import savefile as sf


def test_cextract_bytes_and_roundtrip(tmp_path):
    rng = np.random.default_rng(1)
    tau, col = rng.random((5, 7)), rng.random((5, 7)) * 1e14
    path = str(tmp_path / "out_spectra.dat")
    sf.write_cextract(path, 2.75, 20000.0, tau, col)
    raw = open(path, "rb").read()
    # fwrite(&redshift, 8) fwrite(&box100, 8) fwrite(&NBINS, 4) fwrite(&NumLos, 4) fwrite(pad, 4, 26) then two double arrays
    want = struct.pack("<ddii", 2.75, 20000.0, 7, 5) + b"\0" * (4 * 26) + tau.astype("<f8").tobytes() + col.astype("<f8").tobytes()
    assert len(raw) == 128 + 2 * 5 * 7 * 8 and raw == want
    z, box, t2, c2 = sf.read_cextract(path)
    assert (z, box) == (2.75, 20000.0) and np.array_equal(t2, tau) and np.array_equal(c2, col)
    open(path, "wb").write(raw[:128 + 5 * 7 * 8])  # a file that stops after the optical depths
    assert sf.read_cextract(path)[3] is None
    open(path, "wb").write(raw[:100])
    with pytest.raises(IOError):
        sf.read_cextract(path)
    with pytest.raises(ValueError):
        sf.write_cextract(path, 1.0, 1.0, tau, col[:3])
