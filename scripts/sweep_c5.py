"""BASELINE.json configs[4]: throughput sweep over the number of sightlines and the pixel width, H I Lya tau,
cubic-spline kernel, random sightlines cycling over the three axes.  One JSON line per point.

    python scripts/sweep_c5.py [--nside 256] [--numlos 1000 10000 100000 1000000] [--res 1 2 5 10]

A point = candidate-index build + tau of every sightline, inputs resident in HBM, best of 2 after a warm-up
(CUDA events).  Pixels at least btherm/2 wide take the sub-sampling rule of singleabs.h:110-125 (several inner
quadratures per pixel), which the kernel serves by its generic per-pixel route.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from fake_spectra_b200 import _lib, native  # noqa: E402
from fake_spectra_b200 import synthetic as syn  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nside", type=int, default=256)
    ap.add_argument("--numlos", type=int, nargs="+", default=[1000, 10000, 100000, 1000000])
    ap.add_argument("--res", type=float, nargs="+", default=[1.0, 2.0, 5.0, 10.0])
    ap.add_argument("--max-out-gb", type=float, default=60.0)
    a = ap.parse_args()
    torch.cuda.set_device(0)
    d = syn.boundary_arrays(a.nside)
    t = {k: torch.from_numpy(np.ascontiguousarray(d[k])).cuda() for k in ("pos", "vel", "dens", "temp", "h")}
    for nlos in a.numlos:
        cofm, ax = syn.random_sightlines(d["box"], nlos, axis="cycle")
        tc, ta = torch.from_numpy(cofm).cuda(), torch.from_numpy(ax).cuda()
        for res in a.res:
            p = cases.params(d, res=res)
            gb = nlos * p["nbins"] * 8 / 1e9
            if gb > a.max_out_gb:
                print(json.dumps(dict(numlos=nlos, res=res, skipped="output %.1f GB" % gb)), flush=True)
                continue
            prm = _lib.make_params(**p)
            out = torch.zeros((nlos, p["nbins"]), dtype=torch.float64, device="cuda")
            best = None
            for rep in range(3):
                out.zero_()
                torch.cuda.synchronize()
                e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                e0.record()
                idx = native.CandidateIndex(d["box"], tc, ta, t["pos"], t["h"])
                e1.record()
                idx.compute_tau(prm, t["pos"], t["vel"], t["dens"], t["temp"], t["h"], out=out)
                e2.record()
                if rep == 0:
                    ctr = torch.zeros(10, dtype=torch.int64, device="cuda")
                    idx.compute_tau(prm, t["pos"], t["vel"], t["dens"], t["temp"], t["h"], out=out, counters=ctr)
                    c = ctr.cpu().numpy()
                torch.cuda.synchronize()
                npairs = idx.npairs
                idx.free()
                tot = e0.elapsed_time(e2) * 1e-3
                if rep and (best is None or tot < best[0]):
                    best = (tot, e0.elapsed_time(e1) * 1e-3, e1.elapsed_time(e2) * 1e-3)
            print(json.dumps(dict(nside=a.nside, numlos=nlos, res_kms=res, nbins=p["nbins"], pairs=npairs, s_total=best[0],
                                  s_index=best[1], s_tau=best[2], spectra_per_s=nlos / best[0], pairs_per_s=npairs / best[0],
                                  mean_tau=float(out.mean().item()) / 2 if False else float(out.mean().item()),
                                  pixels=int(c[1]), voigt_evals=int(c[2]), steps_by_route=[int(v) for v in c[4:9]],
                                  lib=os.path.basename(_lib.LIB_PATH))), flush=True)
            del out


if __name__ == "__main__":
    main()
