"""Host classes (Spectra / RandSpectra / GriddedSpectra) on CPU: the reference's API contract
 with the native boundary replaced by an oracle-backed checker, including the
world_size-2 gloo runs of the sightline- and particle-sharded modes."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(__file__))
import hostcases  # noqa: E402
from fake_spectra_b200 import griddedspectra, randspectra, spectra  # noqa: E402


@pytest.fixture()
def backend(oracle):
    return hostcases.OracleBackend(oracle)


def make_rand(backend, nside=10, numlos=14, nsegments=1, **kw):
    snap = hostcases.snapshot(nside, nsegments)
    return randspectra.RandSpectra(0, snap, numlos=numlos, thresh=0., res=2.0, quiet=True, backend=backend, **kw)


def test_constructor_contract(backend):
    rs = make_rand(backend)
    snap = rs.snapshot_set
    assert rs.NumLos == 14 and rs.cofm.shape == (14, 3) and rs.cofm.dtype == np.float64
    assert rs.axis.dtype == np.int32 and np.all(rs.axis == 1)
    # randspectra.py:23-24,32-37: np.random.seed(seed); box * random_sample
    np.random.seed(23)
    assert np.array_equal(rs.cofm, snap.get_header_attr("BoxSize") * np.random.random_sample((14, 3)))
    assert rs.rscale.dtype == np.float32
    assert rs.nbins == int(rs.vmax / 2.0) and np.isclose(rs.dvbin, rs.vmax / rs.nbins)
    assert rs.tautail == 1e-7 and rs.kernel_int == 1
    assert np.isclose(rs.velfac, rs.rscale * rs.Hz / 3.085678e24)
    with pytest.raises(ValueError):
        spectra.Spectra(0, snap, rs.cofm, rs.axis, reload_file=True, quiet=True, kernel="bogus", backend=backend)
    one = spectra.Spectra(0, snap, np.array([1., 2., 3.]), np.array(2), reload_file=True, quiet=True, backend=backend)
    assert one.cofm.shape == (1, 3) and one.axis.shape == (1,) and one.NumLos == 1


def test_get_tau_matches_a_direct_boundary_call(backend, oracle):
    """get_tau == the boundary call on the arrays the reference's _read_particle_data would build
    (spectra.py:550-617, restated here in numpy)."""
    rs = make_rand(backend)
    snap, gp = rs.snapshot_set, rs.gasprop
    pos = snap.get_data(0, "Position", segment=0).astype(np.float32)
    hh = snap.get_smooth_length(0, segment=0).astype(np.float32)
    ind = oracle.near_lines(rs.box, pos, hh, rs.axis, rs.cofm)
    vel = snap.get_peculiar_velocity(0, segment=0).astype(np.float32)[ind]
    den = gp.get_code_rhoH(0, segment=0).astype(np.float32)[ind]
    temp = gp.get_temp(0, segment=0).astype(np.float32)[ind]
    temp[temp <= 0] = 1
    mass_frac = snap.get_data(0, "GFM_Metals", segment=0).astype(np.float32)[:, 0][ind]
    elem_den = (den * rs.rscale) * mass_frac
    elem_den *= gp.get_reproc_HI(0, segment=0)[ind].astype(np.float32)
    elem_den /= rs.lines.get_mass("H")
    line = rs.lines[("H", 1)][1215]
    want = oracle.compute_tau(rs.nbins, 1, rs.box, rs.velfac, rs.atime, line.lambda_X * 1e-8, line.gamma_X, line.fosc_X,
                              rs.lines.get_mass("H"), 1e-7, pos[ind], vel, elem_den.astype(np.float32), temp, hh[ind],
                              rs.axis, rs.cofm)
    tau = rs.get_tau("H", 1, 1215)
    assert tau.shape == (rs.NumLos, rs.nbins) and tau.dtype == np.float64
    assert np.array_equal(tau, want)
    assert rs.get_tau("H", 1, 1215) is tau                      # cached (spectra.py:883-890)
    assert np.array_equal(rs.get_tau("H", 1, 1215, number=3), tau[3])
    assert 0.01 < np.mean(tau) < 50


def test_colden_density_and_weighted_fields(backend):
    rs = make_rand(backend)
    colden = rs.get_col_density("H", 1)
    assert colden.shape == (rs.NumLos, rs.nbins) and np.all(colden >= 0) and colden.max() > 0
    phys = rs.dvbin / rs.velfac * rs.rscale
    assert np.allclose(rs.get_density("H", 1), colden / phys)
    temp = rs.get_temp("H", 1)
    sel = colden > 0
    assert temp.shape == colden.shape
    assert np.all(temp[sel] > 1e2) and np.all(temp[sel] < 1e7)  # a density-weighted mean of particle temperatures
    vel = rs.get_velocity("H", 1)
    assert vel.shape == (rs.NumLos, rs.nbins, 3) and vel.dtype == np.float32
    assert np.all(np.abs(vel) < 2000)
    dwd = rs.get_dens_weighted_density("H", 1)
    assert dwd.shape == colden.shape and np.all(dwd[sel] > 0)
    # turn_off_selfshield: Gamma = 0 at the boundary (spectra.py:669-672)
    g0 = make_rand(backend, turn_off_selfshield=True)
    assert not np.array_equal(g0.get_tau("H", 1, 1215), rs.get_tau("H", 1, 1215))


def test_segments_accumulate(backend):
    """A snapshot split in three segments gives the single-segment result (spectra.py:818-823)."""
    a = make_rand(backend, nsegments=1).get_tau("H", 1, 1215)
    b = make_rand(backend, nsegments=3).get_tau("H", 1, 1215)
    assert np.allclose(a, b, rtol=1e-12, atol=0)
    assert np.array_equal(a == 0, b == 0)


def test_fused_lines_fill_the_cache(backend):
    rs = make_rand(backend)
    both = rs.get_tau_lines("H", 1, [1215, 1025])
    assert set(both) == {1215, 1025}
    assert np.array_equal(both[1215], make_rand(backend).get_tau("H", 1, 1215))
    assert rs.get_tau("H", 1, 1025) is both[1025]
    assert both[1025].max() < both[1215].max()


def test_gridded_spectra_layout(backend):
    snap = hostcases.snapshot(8)
    gs = griddedspectra.GriddedSpectra(0, snap, nspec=4, res=5.0, axis=-1, quiet=True, backend=backend)
    box = snap.get_header_attr("BoxSize")
    assert gs.NumLos == 48 and list(gs.axis[:16]) == [1] * 16 and list(gs.axis[16:32]) == [2] * 16
    # reference griddedspectra.py:68-81: [0,nn,mm], [nn,0,mm], [nn,mm,0] times box/nspec
    assert np.array_equal(gs.cofm[5], box / 4 * np.array([0, 1, 1]))
    assert np.array_equal(gs.cofm[16 + 6], box / 4 * np.array([1, 0, 2]))
    assert np.array_equal(gs.cofm[32 + 7], box / 4 * np.array([1, 3, 0]))
    tau = gs.get_tau("H", 1, 1215)
    assert tau.shape == (48, gs.nbins) and tau.max() > 0
    with pytest.raises(ValueError):
        griddedspectra.GriddedSpectra(0, snap, nspec=4, res=None, nbins=None, quiet=True, backend=backend)
    g2 = griddedspectra.GriddedSpectra(0, snap, nspec=2, res=None, nbins=64, quiet=True, backend=backend)
    assert g2.nbins == 64 and np.isclose(g2.dvbin, g2.vmax / 64)


def test_savefile_roundtrip(backend, tmp_path):
    rs = make_rand(backend, savefile="spectra.npz", savedir=str(tmp_path))
    tau = rs.get_tau("H", 1, 1215)
    col = rs.get_col_density("H", 1)
    rs.save_file()
    snap = rs.snapshot_set
    back = spectra.Spectra(0, snap, None, None, savefile="spectra.npz", savedir=str(tmp_path), res=None, quiet=True,
                           backend=backend)
    assert back.nbins == rs.nbins and np.array_equal(back.cofm, rs.cofm) and np.isclose(back.velfac, rs.velfac)
    assert np.array_equal(back.get_tau("H", 1, 1215), tau)
    assert np.array_equal(back.get_col_density("H", 1), col)
    with pytest.raises(IOError):
        spectra.Spectra(0, snap, None, None, savefile="missing.npz", savedir=str(tmp_path), quiet=True, backend=backend)


def test_savefile_keeps_untouched_arrays(backend, tmp_path):
    """Reopen a savefile, touch one array only, save again: the untouched arrays must survive (the first file becomes
    the .backup, a second save overwrites that)."""
    rs = make_rand(backend, savefile="spectra.npz", savedir=str(tmp_path))
    tau, col = rs.get_tau("H", 1, 1215), rs.get_col_density("H", 1)
    rs.save_file()
    for _ in range(2):
        back = spectra.Spectra(0, rs.snapshot_set, None, None, savefile="spectra.npz", savedir=str(tmp_path), res=None,
                               quiet=True, backend=backend)
        assert np.array_equal(back.get_col_density("H", 1), col)  # tau stays a lazy placeholder
        back.save_file()
    last = spectra.Spectra(0, rs.snapshot_set, None, None, savefile="spectra.npz", savedir=str(tmp_path), res=None, quiet=True,
                           backend=backend)
    assert np.array_equal(last.get_tau("H", 1, 1215), tau) and np.array_equal(last.get_col_density("H", 1), col)


def test_savefile_hdf5_branch(backend, tmp_path, monkeypatch):
    """The h5py branch of the savefile (the reference's spectra.hdf5 tree: Header attributes, spectra/cofm, tau/<elem>/<ion>/
    <line>, colden/<elem>/<ion>, ...; spectra.py:266-372,434-499) driven through a dictionary-backed stand-in for h5py
    (tests/fake_h5py.py; h5py itself is not installed here): same round trip and lazy loading as the .npz tree."""
    import fake_h5py
    monkeypatch.setitem(sys.modules, "h5py", fake_h5py)
    rs = make_rand(backend, savefile="spectra.hdf5", savedir=str(tmp_path))
    tau, taub = rs.get_tau("H", 1, 1215), rs.get_tau("H", 1, 1025)
    col, temp = rs.get_col_density("H", 1), rs.get_temp("H", 1)
    rs.save_file()
    with fake_h5py.File(os.path.join(str(tmp_path), "spectra.hdf5"), "r") as f:
        names = []
        f.visititems(lambda n, o: names.append(n) if isinstance(o, fake_h5py.Dataset) else None)
        assert {"spectra/cofm", "spectra/axis", "tau/H/1/1215", "tau/H/1/1025", "colden/H/1", "temperature/H/1"} <= set(names)
        assert set(f["Header"].attrs) == {"redshift", "nbins", "hubble", "box", "omegam", "omegab", "omegal", "discarded", "npart", "Hz"}
        assert set(f.keys()) >= {"Header", "spectra", "tau_obs", "tau", "colden", "velocity", "temperature", "num_important",
                                 "density_weight_density"}
    back = spectra.Spectra(0, rs.snapshot_set, None, None, savefile="spectra.hdf5", savedir=str(tmp_path), res=None, quiet=True,
                           backend=backend)
    assert back.nbins == rs.nbins and np.array_equal(back.cofm, rs.cofm) and np.array_equal(back.axis, rs.axis)
    assert np.size(back.tau[("H", 1, 1215)]) == 1  # a lazy placeholder until it is asked for
    assert np.array_equal(back.get_tau("H", 1, 1215), tau) and np.array_equal(back.get_tau("H", 1, 1025), taub)
    assert np.array_equal(back.get_col_density("H", 1), col) and np.array_equal(back.get_temp("H", 1), temp)
    back.save_file()  # the first file becomes the backup; every array, loaded or not, is written again
    assert os.path.exists(os.path.join(str(tmp_path), "spectra.hdf5.backup"))
    again = spectra.Spectra(0, rs.snapshot_set, None, None, savefile="spectra.hdf5", savedir=str(tmp_path), res=None, quiet=True,
                            backend=backend)
    assert np.array_equal(again.get_tau("H", 1, 1025), taub)
    with pytest.raises(IOError):
        spectra.Spectra(0, rs.snapshot_set, None, None, savefile="missing.hdf5", savedir=str(tmp_path), quiet=True, backend=backend)


def test_unitsystem_hubble_takes_arrays():
    from fake_spectra_b200 import unitsystem
    u = unitsystem.UnitSystem()
    z = np.array([0., 1., 3.])
    h = u.hubble(z, 0.3) if u.hubble.__code__.co_argcount >= 3 else None
    if h is not None:
        assert h.shape == (3,) and np.all(np.diff(h) > 0)


def test_balanced_blocks():
    from fake_spectra_b200 import sharding
    w = np.array([1, 1, 1, 1, 10, 1, 1, 1, 1, 10])
    e = sharding.balanced_blocks(w, 2)
    assert e[0] == 0 and e[-1] == 10 and abs(w[:e[1]].sum() - w[e[1]:].sum()) <= 10
    assert list(sharding.even_blocks(10, 4)) == [0, 2, 5, 7, 10]
    assert list(sharding.balanced_blocks(np.zeros(6), 3)) == [0, 2, 4, 6]
    e = sharding.balanced_blocks(np.ones(7), 8)
    assert e[0] == 0 and e[-1] == 7 and np.all(np.diff(e) >= 0)


def _worker(rank, world, port, mode, out_dir):
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(__file__))
    import hostcases as hc
    from oracle import Oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = Oracle()
    orc.set_threads(2)
    be = hc.OracleBackend(orc)
    snap = hc.snapshot(10, 2)
    balanced = mode == "sightlines-balanced"
    rs = randspectra.RandSpectra(0, snap, numlos=13, thresh=0., res=2.0, quiet=True, backend=be,
                                 shard="sightlines" if balanced else mode)
    edges = rs.balance_sightlines() if balanced else None
    tau = rs.get_tau("H", 1, 1215)
    vel = rs.get_velocity("H", 1)
    np.savez(os.path.join(out_dir, "r%d.npz" % rank), tau=tau, vel=vel, lines=[c[2] for c in be.calls], parts=[c[1] for c in be.calls],
             edges=np.zeros(0) if edges is None else edges)
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["sightlines", "particles", "sightlines-balanced"])
def test_two_rank_sharding_gloo(backend, tmp_path, mode):
    """world_size 2 over gloo: both ranks end with the full arrays, equal to the unsharded run
    (bitwise for sightline sharding; to summation order for particle sharding)."""
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000) + {"sightlines": 0, "particles": 7, "sightlines-balanced": 13}[mode]
    mp.spawn(_worker, args=(2, port, mode, str(tmp_path)), nprocs=2, join=True)
    ref = make_rand(backend, numlos=13, nsegments=2)
    tau, vel = ref.get_tau("H", 1, 1215), ref.get_velocity("H", 1)
    outs = [np.load(os.path.join(str(tmp_path), "r%d.npz" % r)) for r in range(2)]
    for o in outs:
        if mode == "sightlines":
            assert np.array_equal(o["tau"], tau) and np.array_equal(o["vel"], vel)
            assert set(o["lines"]) <= {6, 7}          # each rank only interpolated its block of the 13 sightlines
        elif mode == "sightlines-balanced":
            # blocks of equal candidate-pair count (counted on segment 0), same edges on both ranks, same results
            assert np.array_equal(o["tau"], tau) and np.array_equal(o["vel"], vel)
            e = o["edges"]
            assert e[0] == 0 and e[-1] == 13 and 0 < e[1] < 13 and np.array_equal(e, outs[0]["edges"])
        else:
            assert np.allclose(o["tau"], tau, rtol=1e-12, atol=0) and np.array_equal(o["tau"] == 0, tau == 0)
            assert np.allclose(o["vel"], vel, rtol=1e-5, atol=1e-4)
            assert set(o["lines"]) == {13}            # every rank did all sightlines for its particles
    assert np.array_equal(outs[0]["tau"], outs[1]["tau"])


def _block_sum_worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    from fake_spectra_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = sharding.Sharder("particles")
    rng = np.random.default_rng(100 + rank)
    partial = torch.from_numpy(rng.random((1, 37, 11)))     # this rank's particles' contribution to every sightline
    want = partial.clone()
    dist.all_reduce(want)                                     # the unblocked sum (Sharder.combine in this mode)
    blocks = sh.reduce_blocks(37, 5)
    works = [sh.sum_block_async(partial, b0, b1) for b0, b1 in blocks]  # issued block by block, as blocks complete
    for w in works:
        w.wait()
    flat = torch.from_numpy(rng.random((37, 11)))             # a 2-D array (one line): rows of a block are contiguous too
    want2 = flat.clone()
    dist.all_reduce(want2)
    for w in [sh.sum_block_async(flat, b0, b1) for b0, b1 in blocks]:
        w.wait()
    bad = None
    try:
        sh.sum_block_async(torch.zeros((2, 37, 11), dtype=torch.float64), 3, 9)
    except ValueError as exc:
        bad = str(exc)
    np.savez(os.path.join(out_dir, "b%d.npz" % rank), got=partial.numpy(), want=want.numpy(), got2=flat.numpy(), want2=want2.numpy(),
             blocks=np.array(blocks), refused=bad is not None)
    dist.destroy_process_group()


def test_blocked_particle_sum_gloo(tmp_path):
    """Particle-sharded mode sums the partial optical depths in sightline blocks, each all-reduce started as its block is
    finished (sharding.Sharder.reduce_blocks / sum_block_async): world size 2 over gloo, block sums ==
    the whole-array sum bit for bit (the same two addends per element), blocks tile the sightlines, a block of a
    multi-line array is refused (its rows are not contiguous)."""
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000) + 31
    mp.spawn(_block_sum_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    outs = [np.load(os.path.join(str(tmp_path), "b%d.npz" % r)) for r in range(2)]
    for o in outs:
        assert np.array_equal(o["got"], o["want"]) and np.array_equal(o["got2"], o["want2"]) and bool(o["refused"])
        b = o["blocks"]
        assert b[0, 0] == 0 and b[-1, 1] == 37 and np.array_equal(b[1:, 0], b[:-1, 1]) and len(b) == 5
    assert np.array_equal(outs[0]["got"], outs[1]["got"])


def test_res_corr_is_the_reference_filter():
    """spec_utils.res_corr == scipy.ndimage.gaussian_filter1d(mode='wrap') (what the reference calls, spec_utils.py:24),
    also for sigma around one pixel where a continuous Gaussian would differ by tens of per cent."""
    ndimage = pytest.importorskip("scipy.ndimage")
    from fake_spectra_b200.spec_utils import res_corr
    f = np.random.default_rng(0).random((4, 3, 160))
    for fwhm, dv in ((8, 1.0), (8, 5.0), (2, 1.0), (30, 1.0)):
        sig = (fwhm / dv) / (2 * np.sqrt(2 * np.log(2)))
        assert np.allclose(res_corr(f, dv, fwhm), ndimage.gaussian_filter1d(f, sig, axis=-1, mode="wrap"), rtol=0, atol=1e-14)
