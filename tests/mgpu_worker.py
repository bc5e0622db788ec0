"""Worker of tests/test_gpu_multi.py: one process per GPU (torchrun), NCCL.  Checks the product's own sharded path
(Spectra(shard=...)) on real GPUs against an unsharded run on rank 0:

* sightline sharding with pair-balanced blocks: optical depths of two fused lines (rows pushed into every rank's
  array from inside the tau kernel), column density and weighted fields (NCCL gather): BIT-IDENTICAL to one GPU;
* particle sharding (FP64 NCCL sum): within 1e-12 of one GPU, identical zero pattern.
Prints one JSON line on rank 0 and exits non-zero on any mismatch."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fake_spectra_b200 import spectra, synthetic as syn  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    import datetime
    dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=120))
    snap = syn.SyntheticSnapshot(24, seed=4, nsegments=2)
    box = snap.get_header_attr("BoxSize")
    cofm, axis = syn.random_sightlines(box, 301, seed=23, axis="cycle")
    kw = dict(reload_file=True, res=1.0, quiet=True)
    report = {"world": world}

    sh = spectra.Spectra(0, snap, cofm, axis, shard="sightlines", **kw)
    edges = sh.balance_sightlines()
    taus = sh.get_tau_lines("H", 1, [1215, 1025])
    col = sh.get_col_density("H", 1)
    temp = sh.get_temp("H", 1)
    ps = spectra.Spectra(0, snap, cofm, axis, shard="particles", **kw)
    tau_p = ps.get_tau("H", 1, 1215)
    ok = True
    if rank == 0:
        # one GPU, one work row per sightline like the sharded runs (the default on one GPU cuts few sightlines into
        # segments with private rows: same values to rounding, not bit for bit)
        one = spectra.Spectra(0, snap, cofm, axis, shard=None, seg_pairs=1 << 30, **kw)
        ref = one.get_tau_lines("H", 1, [1215, 1025])
        report["edges"] = [int(e) for e in edges]
        for ll in (1215, 1025):
            same = bool(np.array_equal(taus[ll], ref[ll]))
            report["tau_%d_bitwise" % ll] = same
            ok &= same
        report["mean_tau_1215"] = float(ref[1215].mean())
        ok &= 0.01 < report["mean_tau_1215"] < 50
        same = bool(np.array_equal(col, one.get_col_density("H", 1)))
        report["colden_bitwise"] = same
        ok &= same
        same = bool(np.array_equal(temp, one.get_temp("H", 1)))
        report["temp_bitwise"] = same
        ok &= same
        r1 = ref[1215]
        m = r1 != 0
        rel = float(np.max(np.abs(tau_p - r1)[m] / r1[m]))
        report["particles_max_rel"] = rel
        report["particles_zero_pattern"] = bool(np.array_equal(tau_p == 0, r1 == 0))
        ok &= rel <= 1e-12 and report["particles_zero_pattern"]
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    # every rank must hold the same full arrays
    mine = torch.from_numpy(np.ascontiguousarray(taus[1215])).cuda()
    ref0 = mine.clone()
    dist.broadcast(ref0, src=0)
    same_everywhere = torch.tensor([1 if torch.equal(mine, ref0) else 0], device="cuda")
    dist.all_reduce(same_everywhere, op=dist.ReduceOp.MIN)
    if rank == 0:
        report["all_ranks_hold_the_same_rows"] = bool(same_everywhere.item())
        report["ok"] = bool(flag.item()) and bool(same_everywhere.item())
        print(json.dumps(report))
    rc = 0 if (flag.item() and same_everywhere.item()) else 1
    for sp in (sh, ps):
        for peer in sp.__dict__.get("_peer_cache", {}).values():
            peer.close()
    dist.destroy_process_group()
    sys.exit(rc)


if __name__ == "__main__":
    main()
