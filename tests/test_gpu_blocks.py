"""Sightline batching (native.BlockedIndex): a sightline set whose candidate pairs exceed what one index holds (2^31) is
cut into contiguous blocks from a count pass; rows must be bit-identical to the single-index result (every sightline is
an independent work item, part_int.cpp:25-49)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(__file__))
import cases  # noqa: E402


def test_sightline_blocks_partition():
    from fake_spectra_b200.native import sightline_blocks
    counts = [5, 5, 5, 20, 1, 1, 1, 9, 0, 0]
    blocks = sightline_blocks(counts, max_pairs=10)
    assert blocks == [(0, 2), (2, 3), (3, 4), (4, 7), (7, 10)]
    assert sightline_blocks(counts, max_pairs=1000) == [(0, 10)]
    assert sightline_blocks([], max_pairs=10) == [(0, 0)]
    for b0, b1 in blocks:
        assert sum(counts[b0:b1]) <= 10 or b1 - b0 == 1


@pytest.mark.gpu
def test_blocked_index_matches_single_index():
    import torch
    from fake_spectra_b200 import _lib, native
    d = cases.random_case(nside=16, nlos=150, axis="cycle", seed=3)
    p = cases.params(d)
    # one work item per sightline (seg_pairs = 2^30): rows then do not depend on how many sightlines share an index
    prm = [_lib.make_params(**p, seg_pairs=1 << 30),
           _lib.make_params(**dict(p, lambda_cm=1025.7223e-8, fosc=0.07912, gamma=1.897e8), seg_pairs=1 << 30)]
    t = {k: torch.from_numpy(np.ascontiguousarray(d[k])).cuda() for k in ("pos", "vel", "dens", "temp", "h", "cofm", "axis")}
    one = native.CandidateIndex(d["box"], t["cofm"], t["axis"], t["pos"], t["h"])
    many = native.BlockedIndex(d["box"], t["cofm"], t["axis"], t["pos"], t["h"], max_pairs=max(one.npairs // 7, one.max_list))
    assert len(many.blocks) >= 5 and many.npairs == one.npairs and many._single is None
    assert many.blocks[0][0] == 0 and many.blocks[-1][1] == 150
    assert all(a[1] == b[0] for a, b in zip(many.blocks, many.blocks[1:]))
    args = (t["pos"], t["vel"], t["dens"], t["temp"], t["h"])
    assert torch.equal(many.compute_tau(prm, *args), one.compute_tau(prm, *args))
    assert torch.equal(many.compute_tau(prm[0], *args), one.compute_tau(prm[0], *args))
    acc = torch.ones((1, 150, p["nbins"]), dtype=torch.float64, device="cuda")
    many.compute_tau(prm[:1], *args, out=acc)
    assert torch.equal(acc, one.compute_tau(prm[:1], *args, out=torch.ones_like(acc)))  # accumulates into what is there
    w2 = torch.stack([t["dens"], t["dens"] * t["temp"]])
    assert torch.equal(many.compute_colden(prm[0], t["pos"], w2, t["h"]), one.compute_colden(prm[0], t["pos"], w2, t["h"]))
    assert torch.equal(many.compute_colden(prm[0], t["pos"], t["dens"], t["h"]), one.compute_colden(prm[0], t["pos"], t["dens"], t["h"]))
    single = native.BlockedIndex(d["box"], t["cofm"], t["axis"], t["pos"], t["h"])
    assert single._single is not None and torch.equal(single.compute_tau(prm, *args), one.compute_tau(prm, *args))
    with pytest.raises(ValueError):
        many.compute_tau(prm, *args, lines=(0, 10))
