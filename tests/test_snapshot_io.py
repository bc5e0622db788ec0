"""Snapshot readers: the BigFile layout read with numpy (abstractsnapshot.BigFileSnapshot) against the
arrays it was written from, the reference's segment / rank bookkeeping (abstractsnapshot.py:337-375), and the host classes
driven from a snapshot on disc.  The ``bigfile`` package is not available here, so the layout itself is pinned only by
this repository's writer (parity unpinned, see the module docstring)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(__file__))
import cases  # noqa: E402
import hostcases  # noqa: E402
from fake_spectra_b200 import abstractsnapshot as absn  # noqa: E402
from fake_spectra_b200 import randspectra, synthetic  # noqa: E402


class FakeComm:
    def __init__(self, rank, size):
        self.rank, self.size = rank, size

    def Get_rank(self):
        return self.rank

    def Get_size(self):
        return self.size


def test_block_roundtrip_across_data_files(tmp_path):
    rng = np.random.default_rng(0)
    pos = rng.random((1001, 3)).astype(np.float32)
    ids = np.arange(1001, dtype=np.uint64)
    absn.write_bigfile_block(str(tmp_path / "0" / "Position"), pos, nfile=3, attrs={"note": "hello", "x": np.float64(2.5)})
    absn.write_bigfile_block(str(tmp_path / "0" / "ID"), ids, nfile=4)
    blk = absn.BigFileBlock(str(tmp_path / "0" / "Position"))
    assert blk.size == 1001 and blk.nmemb == 3 and len(blk.files) == 3 and blk.dtype == np.dtype("<f4")
    assert blk.attrs["note"] == "hello" and blk.attrs["x"][0] == 2.5
    for sl in (slice(None), slice(0, 1), slice(300, 700), slice(333, 334), slice(990, None), slice(500, 500)):
        assert np.array_equal(blk[sl], pos[sl])
    assert np.array_equal(absn.BigFileBlock(str(tmp_path / "0" / "ID"))[250:760], ids[250:760])
    with pytest.raises(KeyError):
        absn.BigFileBlock(str(tmp_path / "0" / "Nothing"))
    with pytest.raises(TypeError):
        blk[::2]


@pytest.fixture(scope="module")
def on_disc(tmp_path_factory):
    snap = hostcases.snapshot(12, 1)
    path = synthetic.write_bigfile(snap, str(tmp_path_factory.mktemp("bf")), num=3, nfile=3)
    return snap, path


def test_bigfile_snapshot_interface(on_disc):
    snap, path = on_disc
    bf = absn.AbstractSnapshotFactory(3, path)  # no HDF5 there: falls through to BigFile like the reference's factory
    assert isinstance(bf, absn.BigFileSnapshot) and absn.is_snapshot(bf)
    assert bf.get_header_attr("BoxSize") == snap.get_header_attr("BoxSize")
    assert bf.get_header_attr("Time") == snap.get_header_attr("Time")
    assert int(bf.get_npart()[0]) == 12 ** 3 and bf.get_kernel() == 1
    assert bf.get_omega_baryon() == snap.get_omega_baryon()
    assert bf.get_n_segments(0) == 1 and bf.get_blocklen(0, "Density", 0) == 12 ** 3
    for hdf, big in (("Coordinates", "Position"), ("NeutralHydrogenAbundance", "NeutralHydrogenFraction")):
        want = snap.get_data(0, hdf, segment=0)
        assert np.array_equal(bf.get_data(0, hdf, segment=0), want) and np.array_equal(bf.get_data(0, big, segment=-1), want)
    assert np.array_equal(bf.get_smooth_length(0, 0), snap.get_smooth_length(0, 0))
    # header values read from a file are numpy float64: numpy then forms the temperature in double (like the reference on
    # real snapshots), while the in-memory snapshot's Python-float units keep float32
    t_disc, t_mem = bf.get_temp(0, 0), snap.get_temp(0, 0)
    assert t_disc.dtype == np.float64 and t_mem.dtype == np.float32 and np.allclose(t_disc, t_mem, rtol=3e-7)
    # MP-Gadget convention: stored a v_pec, divided by a on reading (float32 rounding of the two conversions)
    assert np.allclose(bf.get_peculiar_velocity(0, 0), snap.get_peculiar_velocity(0, 0), rtol=3e-7, atol=1e-4)
    with pytest.raises(KeyError):
        bf.get_data(0, "Volume", segment=0)
    with pytest.raises(KeyError):
        bf.get_header_attr("NoSuchAttribute")
    with pytest.raises(IOError):
        absn.AbstractSnapshotFactory(4, path + "_missing")


def test_bigfile_ranks_and_segments(on_disc):
    """Particles are dealt to the ranks evenly (the first ranks take the remainder) and a rank's share is cut into
    segments (abstractsnapshot.py:337-375); together the ranks read every particle they own exactly once."""
    snap, path = on_disc
    want = snap.get_data(0, "Density", segment=0)
    got = []
    for rank in range(5):
        bf = absn.BigFileSnapshot(3, path, FakeComm(rank, 5))
        nseg = bf.get_n_segments(0, chunk_size=100.)
        assert nseg == int(bf.parts_rank[rank] / 100.) and bf.parts_rank.sum() == 12 ** 3
        bf.get_n_segments = lambda part_type=0, chunk_size=100., _f=bf.get_n_segments: _f(part_type, chunk_size)
        for seg in range(nseg):
            got.append(bf.get_data(0, "Density", segment=seg))
            assert bf.get_blocklen(0, "Density", seg) == got[-1].shape[0]
    got = np.concatenate(got)
    # the reference drops the remainder of each rank's share when it does not divide into the segments
    assert got.shape[0] <= want.shape[0] and np.all(np.isin(got, want))
    assert got.shape[0] >= want.shape[0] - 5 * 3


def test_spectra_from_a_bigfile_snapshot(on_disc, oracle):
    """The host classes on a snapshot read from disc == on the same arrays held in memory (host-prepared route, the
    CPU checker as the backend); only the velocity differs by float32 rounding of MP-Gadget's convention."""
    snap, path = on_disc
    kw = dict(numlos=16, thresh=0., res=2., quiet=True, backend=hostcases.OracleBackend(oracle))
    mem = randspectra.RandSpectra(3, snap, **kw)
    disc = randspectra.RandSpectra(3, path, **kw)
    assert isinstance(disc.snapshot_set, absn.BigFileSnapshot)
    assert disc.box == mem.box and disc.nbins == mem.nbins and np.array_equal(disc.cofm, mem.cofm)
    rel, same = cases.rel_err(disc.get_col_density("H", 1), mem.get_col_density("H", 1))
    assert same and rel == 0.0
    a, b = disc.get_tau("H", 1, 1215), mem.get_tau("H", 1, 1215)
    big = b > 1e-6 * b.max()
    assert np.max(np.abs(a[big] - b[big]) / b[big]) < 1e-4


def test_hdf5_snapshot_reader_through_a_stand_in(tmp_path, monkeypatch, oracle):
    """HDF5Snapshot (Gadget / Arepo files, one segment per file, abstractsnapshot.py:156-301) driven through the
    dictionary-backed stand-in for h5py (tests/fake_h5py.py; h5py itself is absent here): file discovery, name translation,
    all-segments reads, kernel detection, header access, and the host classes on top of it."""
    import fake_h5py
    monkeypatch.setitem(sys.modules, "h5py", fake_h5py)
    snap = hostcases.snapshot(10, 1)
    n = snap.fields["Density"].shape[0]
    edges = [0, n // 3, n]
    snapdir = tmp_path / "snapdir_007"
    snapdir.mkdir()
    for k in range(2):
        with fake_h5py.File(str(snapdir / ("snap_007.%d.hdf5" % k)), "w") as f:
            head = f.create_group("Header")
            for key, value in snap.header.items():
                head.attrs[key] = np.asarray(value) if isinstance(value, np.ndarray) else value
            head.attrs["NumPart_Total_HighWord"] = np.zeros(6, dtype=np.int64)
            nk = edges[k + 1] - edges[k]
            head.attrs["NumPart_ThisFile"] = np.array([nk, nk, 0, 0, 0, 0])
            head.attrs["MassTable"] = np.array([0., 5.0, 0., 0., 0., 0.])      # dark matter particle mass
            for name, data in snap.fields.items():
                f.create_dataset("PartType0/" + name, data=data[edges[k]:edges[k + 1]])
            f.create_dataset("PartType0/Masses", data=np.ones(nk, dtype=np.float32))
    hs = absn.AbstractSnapshotFactory(7, str(tmp_path))
    assert isinstance(hs, absn.HDF5Snapshot) and hs.get_n_segments() == 2 and hs.get_kernel() == 1
    assert hs.get_header_attr("BoxSize") == snap.get_header_attr("BoxSize") and int(hs.get_npart()[0]) == n
    assert np.isclose(hs.get_omega_baryon(), 1. / 6. * snap.get_header_attr("Omega0"))  # gas 1 : dark matter 5 per particle
    whole = hs.get_data(0, "Position", segment=-1)                       # BigFile name, every file (reverse-sorted like the reference)
    assert whole.shape == (n, 3) and np.array_equal(np.sort(whole, axis=0), np.sort(snap.fields["Coordinates"], axis=0))
    lens = [hs.get_blocklen(0, "Density", s) for s in range(2)]
    assert sorted(lens) == sorted(np.diff(edges).tolist())
    assert np.array_equal(hs.get_smooth_length(0, 0), hs.get_data(0, "SmoothingLength", segment=0) / 2)
    with pytest.raises(KeyError):
        hs.get_data(0, "Volume", segment=0)
    kw = dict(numlos=10, thresh=0., res=2., quiet=True, backend=hostcases.OracleBackend(oracle))
    disc, mem = randspectra.RandSpectra(7, str(tmp_path), **kw), randspectra.RandSpectra(7, snap, **kw)
    assert isinstance(disc.snapshot_set, absn.HDF5Snapshot) and np.array_equal(disc.cofm, mem.cofm)
    a, b = disc.get_tau("H", 1, 1215), mem.get_tau("H", 1, 1215)  # two segments in another particle order: summation order only
    assert np.allclose(a, b, rtol=1e-12, atol=0) and np.array_equal(a == 0, b == 0)
    rel, same = cases.rel_err(disc.get_col_density("H", 1), mem.get_col_density("H", 1))
    assert same and rel < 1e-12
