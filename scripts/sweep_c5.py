"""BASELINE.json configs[4]: throughput sweep over the number of sightlines and the pixel width on the C3 snapshot
(2 x 512^3), H I Lya tau, cubic-spline kernel, random sightlines cycling over the three axes.  One JSON line per point.

    python scripts/sweep_c5.py [--nside 512] [--numlos 1000 10000 100000 1000000] [--res 1 2 5 10] [--cpu-lines 256]
    python -m torch.distributed.run --nproc-per-node N ... scripts/sweep_c5.py ...     (N GPUs of one node)

A point = candidate-index build + tau of every sightline, inputs resident in HBM, best of 2 after a warm-up (CUDA events,
maximum over the ranks).  With N ranks the sightlines are cut into contiguous blocks of equal candidate-pair count
(sharding.balanced_blocks from one count pass), particles replicated, no data-path collective.  Sightline sets whose pairs
exceed what one index holds (2^31: 10^6 sightlines) are batched by native.BlockedIndex.  Points whose output would not fit
(--max-out-gb per GPU) are skipped and say so.  The CPU column is the unmodified reference (oracle/_ref, OpenMP on all
host threads) on --cpu-lines sightlines of the same set against the full particle set, candidate search included, at
every pixel width; its spectra/s do not depend on the number of sightlines beyond the index build it repeats.
Pixels at least btherm/2 wide take the sub-sampling rule of singleabs.h:110-125.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from fake_spectra_b200 import _lib, native, sharding  # noqa: E402
from fake_spectra_b200 import synthetic as syn  # noqa: E402


def cpu_column(d, cofm, ax, res_list, nlines):
    """spectra/s of the reference C++ per pixel width (rank 0 only)."""
    from oracle import Reference
    if not Reference.available():
        return {}
    ref = Reference()
    # all host threads, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1 to every rank)
    ref.set_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    out = {}
    for res in res_list:
        p = cases.params(d, res=res)
        t0 = time.perf_counter()
        ref.compute_tau(**p, pos=d["pos"], vel=d["vel"], dens=d["dens"], temp=d["temp"], h=d["h"], axis=ax[:nlines], cofm=cofm[:nlines])
        out[res] = nlines / (time.perf_counter() - t0)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nside", type=int, default=512)
    ap.add_argument("--numlos", type=int, nargs="+", default=[1000, 10000, 100000, 1000000])
    ap.add_argument("--res", type=float, nargs="+", default=[1.0, 2.0, 5.0, 10.0])
    ap.add_argument("--max-out-gb", type=float, default=100.0)
    ap.add_argument("--cpu-lines", type=int, default=256)
    a = ap.parse_args()
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl")

    def over_ranks(value, op):
        if world == 1:
            return value
        v = torch.tensor([value], dtype=torch.float64, device="cuda")
        dist.all_reduce(v, op=op)
        return float(v.item())

    d = syn.boundary_arrays(a.nside)
    t = {k: torch.from_numpy(np.ascontiguousarray(d[k])).cuda() for k in ("pos", "vel", "dens", "temp", "h")}
    cofm_all, ax_all = syn.random_sightlines(d["box"], max(a.numlos), axis="cycle")
    cpu = cpu_column(d, cofm_all, ax_all, a.res, a.cpu_lines) if rank == 0 and a.cpu_lines > 0 else {}
    for nlos in a.numlos:
        cofm, ax = cofm_all[:nlos], ax_all[:nlos]
        tc, ta = torch.from_numpy(cofm).cuda(), torch.from_numpy(ax).cuda()
        b0, b1 = 0, nlos
        if world > 1:
            counts = native.count_pairs(d["box"], t["pos"], t["h"], ta, tc).cpu().numpy()
            edges = sharding.balanced_blocks(counts, world)
            b0, b1 = int(edges[rank]), int(edges[rank + 1])
        mc, ma = tc[b0:b1].contiguous(), ta[b0:b1].contiguous()
        for res in a.res:
            p = cases.params(d, res=res)
            gb = (b1 - b0) * p["nbins"] * 8 / 1e9
            if over_ranks(gb, dist.ReduceOp.MAX if dist else None) > a.max_out_gb:
                if rank == 0:
                    print(json.dumps(dict(numlos=nlos, res_kms=res, n_gpus=world, skipped="output %.1f GB per GPU" % gb)), flush=True)
                continue
            prm = _lib.make_params(**p, seg_pairs=(1 << 30) if world > 1 else 0)
            out = torch.zeros((max(b1 - b0, 1), p["nbins"]), dtype=torch.float64, device="cuda")
            best, npairs, nblocks = None, 0, 1
            for rep in range(3):
                out.zero_()
                torch.cuda.synchronize()
                if dist:
                    dist.barrier()
                e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                e0.record()
                idx = native.BlockedIndex(d["box"], mc, ma, t["pos"], t["h"]) if b1 > b0 else None
                e1.record()
                if idx is not None:
                    idx.compute_tau(prm, t["pos"], t["vel"], t["dens"], t["temp"], t["h"], out=out[:b1 - b0])
                e2.record()
                torch.cuda.synchronize()
                if idx is not None:
                    npairs, nblocks = idx.npairs, len(idx.blocks)
                    idx.free()
                tot = over_ranks(e0.elapsed_time(e2) * 1e-3, dist.ReduceOp.MAX if dist else None)
                if rep and (best is None or tot < best[0]):
                    best = (tot, e0.elapsed_time(e1) * 1e-3, e1.elapsed_time(e2) * 1e-3)
            pairs_all = over_ranks(float(npairs), dist.ReduceOp.SUM if dist else None)
            mean_tau = over_ranks(float(out[:b1 - b0].sum().item()), dist.ReduceOp.SUM if dist else None) / (nlos * p["nbins"])
            if rank == 0:
                line = dict(nside=a.nside, numlos=nlos, res_kms=res, nbins=p["nbins"], n_gpus=world, pairs=int(pairs_all), s_total=best[0],
                            s_index_rank0=best[1], s_tau_rank0=best[2], spectra_per_s=nlos / best[0], pairs_per_s=pairs_all / best[0],
                            index_blocks_rank0=nblocks, mean_tau=mean_tau, lib=os.path.basename(_lib.LIB_PATH))
                if res in cpu:
                    line.update(cpu_reference_spectra_per_s=cpu[res], cpu_threads=os.cpu_count(), cpu_lines=a.cpu_lines,
                                speedup=nlos / best[0] / cpu[res])
                print(json.dumps(line), flush=True)
            del out
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
