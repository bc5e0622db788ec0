"""Quick on-box probe: FMA peaks, and a timing breakdown of config-1/2-like workloads through the
device API (index build / tau / colden) with the in-kernel counters.  Not the bench."""
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
from fake_spectra_b200 import _lib, native  # noqa: E402
from fake_spectra_b200 import synthetic as syn  # noqa: E402


def timed(fn, reps=3):
    best = 1e30
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best, out


def run(nside, nlos, axis, voigt, reps=2):
    d = syn.boundary_arrays(nside)
    cofm, ax = syn.random_sightlines(d["box"], nlos, axis=axis)
    d["cofm"], d["axis"] = cofm, ax
    t = {k: torch.from_numpy(np.ascontiguousarray(d[k])).cuda() for k in ("pos", "vel", "dens", "temp", "h", "cofm", "axis")}
    p = cases.params(d)
    prm = _lib.make_params(**p, voigt=voigt)
    t_idx, idx = timed(lambda: native.CandidateIndex(d["box"], t["cofm"], t["axis"], t["pos"], t["h"]), reps)
    ctr = torch.zeros(10, dtype=torch.int64, device="cuda")
    out = torch.zeros((nlos, p["nbins"]), dtype=torch.float64, device="cuda")
    t_tau, _ = timed(lambda: idx.compute_tau(prm, t["pos"], t["vel"], t["dens"], t["temp"], t["h"], out=out), reps)
    idx.compute_tau(prm, t["pos"], t["vel"], t["dens"], t["temp"], t["h"], out=out, counters=ctr)
    t_col, _ = timed(lambda: idx.compute_colden(prm, t["pos"], t["dens"], t["h"]), reps)
    prm_b = _lib.make_params(**cases.params(d, line="HI1025"), voigt=voigt)
    out2 = torch.zeros((2, nlos, p["nbins"]), dtype=torch.float64, device="cuda")
    t_tau2, _ = timed(lambda: idx.compute_tau([prm, prm_b], t["pos"], t["vel"], t["dens"], t["temp"], t["h"], out=out2), reps)
    prm32 = _lib.make_params(**p, voigt=voigt, precision=_lib.PRECISION_FP32)
    prm32_b = _lib.make_params(**cases.params(d, line="HI1025"), voigt=voigt, precision=_lib.PRECISION_FP32)
    ref64 = out.clone()
    t_tau32, _ = timed(lambda: idx.compute_tau(prm32, t["pos"], t["vel"], t["dens"], t["temp"], t["h"], out=out.zero_()), reps)
    flux_err = float((torch.exp(-out) - torch.exp(-ref64 / max(reps, 1) if False else -ref64)).abs().max().item()) if False else None
    t_tau32_2, _ = timed(lambda: idx.compute_tau([prm32, prm32_b], t["pos"], t["vel"], t["dens"], t["temp"], t["h"], out=out2), reps)
    c = ctr.cpu().numpy()
    res = dict(nside=nside, nlos=nlos, nbins=p["nbins"], voigt=voigt, npairs=idx.npairs, max_list=idx.max_list,
               t_index=t_idx, t_tau=t_tau, t_tau_lya_lyb=t_tau2, t_tau_fp32=t_tau32, t_tau_fp32_lya_lyb=t_tau32_2, t_colden=t_col, lib=os.path.basename(_lib.LIB_PATH), pairs_per_s=idx.npairs / t_tau, spectra_per_s=nlos / t_tau,
               n_voigt=int(c[2]), voigt_per_s=float(c[2]) / t_tau, pixels=int(c[1]), lane_eff=float(c[1]) / max(float(c[3]), 1),
               routes=[int(v) for v in c[4:9]], check=float(out.mean().item()))
    print(json.dumps(res), flush=True)
    return res


if __name__ == "__main__":
    torch.cuda.set_device(0)
    print(json.dumps(dict(device=torch.cuda.get_device_name(0), info=native.device_info(),
                          fp64_tflops=native.measure_fma_peak(True), fp32_tflops=native.measure_fma_peak(False))), flush=True)
    which = sys.argv[1:] or ["c1"]
    if "c1" in which:
        for v in (1, 0):
            run(64, 1000, 1, v)
    if "c2s" in which:   # config-2-like geometry at reduced size
        run(128, 16384, 1, 0)
    if "c2" in which:
        run(256, 65536, 1, 0, reps=1)
