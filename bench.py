#!/usr/bin/env python
"""Benchmark of the sightline-interpolation hot path (BASELINE.json metric: spectra/s and
sightline-particle pairs/s for Ly-alpha-forest tau).

  python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle/_ref)

A "step" is one pass of the hot path over the whole workload: candidate-index build + optical
depth of every fused line for every sightline.  Default workload = BASELINE.json configs[1]:
GriddedSpectra 256x256 grid (65 536 x-axis sightlines) on a synthetic 2x256^3 snapshot (16.7 M gas
particles, SURVEY App. F generator), cubic-spline SPH kernel, H I Ly-alpha + Ly-beta, 1 km/s pixels.

One process per GPU.  Multi-GPU = sightline sharding with the particle set replicated in each
GPU's HBM and NO data-path collective; weak scaling: every rank processes its
own full 256x256 grid (rank r's grid is shifted by r/N of the grid spacing), so the whole-job
value is the sum over ranks of sightlines / max-over-ranks time.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LINES = {  # lambda (cm), Gamma (1/s), f_osc, amu: SURVEY App. F
    "HI1215": (1215.6701e-8, 6.265e8, 0.4164, 1.00794),
    "HI1025": (1025.7223e-8, 1.897e8, 0.07912, 1.00794),
}
WORKLOADS = {
    # name: (nside, sightline spec, lines, kernel, pixel km/s)
    "c2_grid256_lya_lyb": dict(nside=256, nspec=256, lines=("HI1215", "HI1025"), kernel=1, res=1.0),
    "c1_rand1000_lya": dict(nside=64, numlos=1000, lines=("HI1215",), kernel=1, res=1.0),
    "mini_grid64_lya_lyb": dict(nside=64, nspec=64, lines=("HI1215", "HI1025"), kernel=1, res=1.0),
    # BASELINE.json configs[3] in miniature: Arepo-like top-hat kernel, particles sharded over the ranks
    # (nside^3 cells PER RANK of one common box), every rank computes all sightlines for its cells and the
    # FP64 tau arrays are summed with one NCCL all-reduce per step (the reference's MPI mode, spectra.py:825-831)
    "c4_tophat_pshard": dict(nside=256, numlos=16384, lines=("HI1215",), kernel=0, res=1.0, shard="particles"),
    # BASELINE.json configs[3] at full size when run on 8 GPUs: 8 x 512^3 = 1024^3 cells in a 160 000 kpc/h box
    "c4_tophat_pshard_1024": dict(nside=512, numlos=16384, lines=("HI1215",), kernel=0, res=1.0, shard="particles"),
    "mini_tophat_pshard": dict(nside=64, numlos=2048, lines=("HI1215",), kernel=0, res=1.0, shard="particles"),
}
# Algorithmic FP64 work of THIS library's profile evaluation, per Voigt evaluation
# (one quadrature node of one pixel of one line), FMA = 2 flop.  NEAR route with Gaussian, first line of an
# ion: 14 DFMA (node x 1, table index 2, table Horner 3, A 3, Pe 3, accumulate 2) + 6 DMUL/DADD (index 1,
# x^2 1, Gaussian recurrence 2, kernel weights 2) = 34 flop; every further fused line adds A, Pe and the two
# accumulates = 8 DFMA = 16 flop (node positions, table value and Gaussian are shared).  The t^3..t^7 part of
# the table polynomial runs in FP32 and is NOT counted.  The FAR / no-Gaussian routes are counted at the same
# figures through N only (no extra credit).  280 = the reference algorithm's figure, reported
# separately as reference_equivalent_tflops; it is not the roofline numerator.
FLOP_PER_VOIGT = 34.0
FLOP_PER_VOIGT_FUSED = 16.0
FLOP_PER_VOIGT_REFERENCE = 280.0
TAUTAIL = 1e-7          # reference spectra.py:135


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2_grid256_lya_lyb", choices=sorted(WORKLOADS))
    ap.add_argument("--voigt", default="fast", choices=["fast", "exact"])
    ap.add_argument("--precision", default="fp64", choices=["fp64", "fp32"],
                    help="fp32 = the optional fast path (node sums in single precision, flux within 1e-5 of the reference)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def build_workload(name, rank, world):
    from fake_spectra_b200 import synthetic as syn
    w = dict(WORKLOADS[name])
    w.setdefault("shard", "sightlines")
    if w["shard"] == "particles":
        # rank r holds nside^3 cells (seed 42 + r) of a box sized for world * nside^3 cells; same sightlines everywhere
        box = syn.MEAN_SPACING * w["nside"] * world ** (1.0 / 3.0)
        d = syn.boundary_arrays(w["nside"], seed=42 + rank, kernel=w["kernel"], box=box)
        cos = syn.Cosmology()
        cofm, axis = syn.random_sightlines(box, w["numlos"], seed=23, axis=1)
    else:
        d = syn.boundary_arrays(w["nside"], seed=42, kernel=w["kernel"])
        cos = syn.Cosmology()
        box = d["box"]
    if w["shard"] == "particles":
        pass
    elif "nspec" in w:
        cofm, axis = syn.grid_sightlines(box, w["nspec"], axis=1)
        if world > 1:  # weak scaling: a distinct, shifted grid per rank
            shift = (box / w["nspec"]) * rank / world
            cofm = cofm.copy()
            cofm[:, 1:] += shift
    else:
        cofm, axis = syn.random_sightlines(box, w["numlos"], seed=23 + rank, axis=1)
    velfac = float(cos.velfac)
    nbins = int(box * velfac / w["res"])
    w.update(d)
    w.update(cofm=np.ascontiguousarray(cofm), axis=np.ascontiguousarray(axis), velfac=velfac, atime=cos.atime,
             nbins=nbins, nlos=cofm.shape[0], npart=d["pos"].shape[0], name=name)
    return w


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        power = [float(r[3]) for r in self.rows if len(r) >= 9 and r[3].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "power_w_max": max(power) if power else None}


def make_params(w, line, voigt, precision="fp64"):
    from fake_spectra_b200 import _lib
    lam, gam, fosc, amu = LINES[line]
    return _lib.make_params(w["nbins"], w["kernel"], w["box"], w["velfac"], w["atime"], lam, gam, fosc, amu, TAUTAIL,
                            voigt=_lib.VOIGT_EXACT if voigt == "exact" else _lib.VOIGT_FAST,
                            precision=_lib.PRECISION_FP32 if precision == "fp32" else _lib.PRECISION_FP64)


def run_b200(args):
    import torch
    import torch.distributed as dist
    from fake_spectra_b200 import _lib, native, _spectra_priv

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def gather_ranks(x):
        if world == 1:
            return [x]
        t = torch.zeros(world, dtype=torch.float64, device=dev)
        t[rank] = x
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(v) for v in t.cpu()]

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    w = build_workload(args.workload, rank, world)
    nlines = len(w["lines"])
    params = [make_params(w, ln, args.voigt, args.precision) for ln in w["lines"]]
    fp32 = args.precision == "fp32"
    # FP32 fast path, NEAR route with Gaussian, per Voigt evaluation: node x 1 FFMA, table index 2 FFMA + 1 FADD,
    # degree-3 table Horner 3 FFMA, x^2 1 FMUL, exp(-x^2) 1 FMUL + 1 MUFU (counted 2), A 2 FFMA, Pe 2 FFMA,
    # combine 1 FMUL + 2 FFMA = 12 FFMA + 6 = 30 flop; each further fused line A, Pe, combine = 6 FFMA + 1 = 13 flop
    flop_first, flop_fused = (30.0, 13.0) if fp32 else (FLOP_PER_VOIGT, FLOP_PER_VOIGT_FUSED)
    names = ("pos", "vel", "dens", "temp", "h", "cofm", "axis")
    t = {k: torch.from_numpy(w[k]).to(dev) for k in names}
    out = torch.zeros((nlines, w["nlos"], w["nbins"]), dtype=torch.float64, device=dev)
    fp64_peak = native.measure_fma_peak(not fp32)  # the FMA peak of the precision the node sums run in

    tau_ms = []
    pshard = w["shard"] == "particles"

    idx_ms = []

    def step(counters=None, time_tau=False):
        out.zero_()
        if time_tau:
            i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            i0.record()
        idx = native.CandidateIndex(w["box"], t["cofm"], t["axis"], t["pos"], t["h"])
        if time_tau:
            i1.record()
            idx_ms.append((i0, i1))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        idx.compute_tau(params, t["pos"], t["vel"], t["dens"], t["temp"], t["h"], out=out, counters=counters)
        if time_tau:
            e1.record()
            tau_ms.append((e0, e1))
        if pshard and world > 1:
            dist.all_reduce(out, op=dist.ReduceOp.SUM)  # FP64 [lines, nlos, nbins] over NCCL / NVLink
        npairs = idx.npairs
        idx.free()
        return npairs

    # untimed counter passes: deterministic work counts of one step, per line
    npairs = 0
    n_voigt_line = []
    routes = np.zeros(5, dtype=np.int64)
    for prm in params:
        ctr = torch.zeros(10, dtype=torch.int64, device=dev)
        idx = native.CandidateIndex(w["box"], t["cofm"], t["axis"], t["pos"], t["h"])
        idx.compute_tau(prm, t["pos"], t["vel"], t["dens"], t["temp"], t["h"], out=out[:1], counters=ctr)
        torch.cuda.synchronize()
        npairs = idx.npairs
        idx.free()
        c = ctr.cpu().numpy()
        n_voigt_line.append(int(c[2]))
        routes += c[4:9]
    n_voigt_step = int(sum(n_voigt_line))
    algo_flop_step = flop_first * max(n_voigt_line) + flop_fused * (n_voigt_step - max(n_voigt_line))
    step()
    for _ in range(max(args.warmup - 1, 0)):
        step()
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    launches0 = _lib.load().fsb_kernel_launches()
    barrier()
    torch.cuda.synchronize()
    if rank == 0:
        sampler.start()
    # working set per step: particles (36 B each), candidate index (12 B per pair) and the output.  When it exceeds
    # the 126 MB L2 nothing needs flushing; a small workload gets the L2 overwritten between timed steps (untimed)
    work_bytes = w["npart"] * 36 + 12 * npairs + nlines * w["nlos"] * w["nbins"] * 8
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if work_bytes < (256 << 20) else None
    step_events = []
    for _ in range(args.steps):
        if flush is not None:
            flush.fill_(1)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        step(time_tau=True)
        s1.record()
        step_events.append((s0, s1))
    torch.cuda.synchronize()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = _lib.load().fsb_kernel_launches() - launches0
    elapsed = max_over_ranks(sum(a.elapsed_time(b) for a, b in step_events) * 1e-3)
    # sightline sharding: every rank has its own sightlines; particle sharding: all ranks share them
    total_lines = float(w["nlos"]) if pshard else sum_over_ranks(float(w["nlos"]))
    total_pairs = sum_over_ranks(float(npairs))
    ms_per_step = elapsed / args.steps * 1e3
    value = total_lines * args.steps / elapsed
    pairs_per_s = total_pairs * nlines * args.steps / elapsed
    # one k_tau launch per group of two fused lines (here: one launch for Lya+Lyb)
    n_tau_launches = (nlines + 1) // 2
    tau_launch_s = float(np.mean([a.elapsed_time(b) for a, b in tau_ms])) * 1e-3 / n_tau_launches
    tau_launch_s_ranks = gather_ranks(tau_launch_s)
    tau_launch_s = max(tau_launch_s_ranks)  # the slowest rank bounds the step
    achieved = algo_flop_step / n_tau_launches / tau_launch_s / 1e12
    sanity = float(out[0].mean().item())
    # candidate-index build (K1): HBM-bound by design; algorithmic bytes = 16 B per particle read (pos + h, one
    # axis group here) + 12 B per pair written (int32 particle + f64 dr2), SURVEY 8(d)
    index_s = float(np.mean([a.elapsed_time(b) for a, b in idx_ms])) * 1e-3
    index_bytes = 16.0 * w["npart"] * len(set(int(a) for a in np.unique(w["axis"]))) + 12.0 * npairs
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        hbm_src = "measured (MEASURED_PEAKS.json)"
    except Exception:
        hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"

    # ---- column density (K3) on the same index and particles: one weight column, then three in one pass ----
    colden = None
    if rank == 0 and not pshard:
        idx = native.CandidateIndex(w["box"], t["cofm"], t["axis"], t["pos"], t["h"])
        cout = out[0]
        ctr = torch.zeros(10, dtype=torch.int64, device=dev)
        idx.compute_colden(params[0], t["pos"], t["dens"], t["h"], out=cout.zero_(), counters=ctr)
        torch.cuda.synchronize()
        cpix = int(ctr.cpu()[1])
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(3):
            idx.compute_colden(params[0], t["pos"], t["dens"], t["h"], out=cout)
        c1.record()
        torch.cuda.synchronize()
        col_s = c0.elapsed_time(c1) * 1e-3 / 3
        colden = {"kernel": "k_colden", "ms": col_s * 1e3, "pairs_per_s": idx.npairs / col_s, "pixels": cpix,
                  "kernel_integrals_per_s": cpix / col_s,
                  "note": "one weight column, accumulate into a resident [nlos, nbins] array; each pixel integral is a "
                          "9-node trapezoid of the SPH kernel (absorption.cpp:53-74)"}
        idx.free()
        out.zero_()
        step()  # restore tau in `out` for the statistics below

    # ---- row f2: the consumers of tau while it is still resident (HBM-bound streaming reductions) ----
    flux_stats = None
    if rank == 0:
        from fake_spectra_b200 import fluxstatistics as fstat
        flat = out.view(-1)
        fstat.flux_sums(flat)  # warm-up
        torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        npass = 5
        f0.record()
        for _ in range(npass):
            sums = fstat.flux_sums(flat)
        f1.record()
        torch.cuda.synchronize()
        pass_s = f0.elapsed_time(f1) * 1e-3 / npass
        t0 = time.perf_counter()
        scale = fstat.mean_flux(out[0], 0.7)
        torch.cuda.synchronize()
        newton_s = time.perf_counter() - t0
        flux_stats = {"bound": "hbm", "kernel": "k_flux_sums", "pixels": int(flat.numel()), "ms_per_pass": pass_s * 1e3,
                      "algorithmic_bytes": 8.0 * flat.numel(), "achieved": 8.0 * flat.numel() / pass_s / 1e9, "peak": hbm_peak,
                      "unit": "GB/s", "frac": 8.0 * flat.numel() / pass_s / 1e9 / hbm_peak,
                      "mean_flux": sums[0] / max(sums[2], 1),
                      "rescale_to_0.7": {"scale": scale, "ms": newton_s * 1e3, "pixels": int(out[0].numel())}}

    # ---- end to end through the reference-facing boundary: host buffers in, host buffer out ----
    e2e = None
    if not args.no_e2e:
        pin = {k: torch.from_numpy(w[k]).pin_memory() for k in names}
        hout = torch.empty((nlines, w["nlos"], w["nbins"]), dtype=torch.float64).pin_memory()
        lam, gam, fosc, amu = LINES[w["lines"][0]]
        extra = [LINES[ln][:3] for ln in w["lines"][1:]]
        vg = _lib.VOIGT_EXACT if args.voigt == "exact" else _lib.VOIGT_FAST
        prec = _lib.PRECISION_FP32 if fp32 else _lib.PRECISION_FP64

        def e2e_step_host():
            return _spectra_priv._Particle_Interpolate(
                1, w["nbins"], w["kernel"], w["box"], w["velfac"], w["atime"], lam, gam, fosc, amu, TAUTAIL,
                pin["pos"].numpy(), pin["vel"].numpy(), pin["dens"].numpy(), pin["temp"].numpy(), pin["h"].numpy(),
                pin["axis"].numpy(), pin["cofm"].numpy(), voigt=vg, precision=prec, out=hout.numpy(), extra_lines=extra)

        def e2e_step_pshard():
            # particle-sharded: upload this rank's particles, interpolate all sightlines, sum over ranks
            # on the device (NCCL), read the result back
            dv = {k: pin[k].to(dev, non_blocking=True) for k in names}
            o = torch.zeros((nlines, w["nlos"], w["nbins"]), dtype=torch.float64, device=dev)
            idx = native.CandidateIndex(w["box"], dv["cofm"], dv["axis"], dv["pos"], dv["h"])
            idx.compute_tau(params, dv["pos"], dv["vel"], dv["dens"], dv["temp"], dv["h"], out=o)
            idx.free()
            if world > 1:
                dist.all_reduce(o, op=dist.ReduceOp.SUM)
            hout.copy_(o, non_blocking=True)
            torch.cuda.synchronize()
            return hout.numpy()

        e2e_step = e2e_step_pshard if pshard else e2e_step_host

        e2e_step()  # warm-up (allocations, page faults of the pinned result)
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res = e2e_step()
        torch.cuda.synchronize()
        e2e_elapsed = max_over_ranks(time.perf_counter() - t0)
        barrier()
        h2d = sum(int(pin[k].numel() * pin[k].element_size()) for k in names)
        d2h = int(hout.numel() * hout.element_size())
        e2e = {"value": total_lines * args.steps / e2e_elapsed, "unit": "spectra/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": e2e_elapsed / args.steps * 1e3,
               "call": ("pinned host -> device, native.CandidateIndex + compute_tau, NCCL all-reduce, device -> pinned host" if pshard else
                        "_spectra_priv._Particle_Interpolate(host buffers, extra_lines) -> fsb_particle_interpolate_multi_host"),
               "mean_tau_check": float(np.mean(res[0][: min(64, w["nlos"])]))}
        del pin, hout

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_sample(w, steps=1)

    if rank == 0:
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "tau_traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get(args.workload)
            except Exception:
                traffic = None
        line = {
            "metric": "spectra_per_s", "value": value, "unit": "spectra/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if fp32 else "f64", "data": "synthetic",
            "config": {"workload": args.workload, "particles": int(w["npart"]), "sightlines_per_gpu": int(w["nlos"]),
                       "precision": args.precision,
                       "pixels": int(w["nbins"]), "pixel_kms": w["res"], "lines": list(w["lines"]), "sph_kernel": {0: "tophat", 1: "cubic", 2: "voronoi", 3: "quintic"}[w["kernel"]],
                       "voigt": args.voigt,
                       "parallelism": ("particle-sharded x%d (%d cells per rank), NCCL all-reduce of the FP64 tau array each step"
                                       % (world, w["npart"])) if pshard else "sightline-sharded x%d, particles replicated" % world,
                       "l2": ("inputs and outputs larger than L2 (particles %.2f GB, tau %.2f GB per GPU)" if flush is None else
                              "working set fits the L2 (particles %.3f GB, tau %.3f GB): a 256 MB buffer is overwritten "
                              "between timed steps, outside the timed intervals") % (
                           w["npart"] * 36 / 1e9, nlines * w["nlos"] * w["nbins"] * 8 / 1e9),
                       "step": "index build + tau of all lines for every sightline, inputs resident in HBM"},
            "pairs_per_s": pairs_per_s, "pairs_per_step": total_pairs, "voigt_evals_per_step": n_voigt_step * world,
            "voigt_evals_per_s": n_voigt_step * world * args.steps / elapsed,
            "roofline": {"bound": "fp32" if fp32 else "fp64", "kernel": "k_tau", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved / fp64_peak, "traffic": traffic,
                         "note": ("algorithmic FP32 flop" if fp32 else "algorithmic FP64 flop") + " of this library's profile evaluation (DESIGN.md 5): %.0f per Voigt "
                                 "evaluation of the first line + %.0f per evaluation of each fused line = %.3e flop per launch / "
                                 "mean k_tau launch time %.4f s (CUDA events, timed region); peak = DFMA rate measured on this "
                                 "device by fsb_measure_fma_peak in the same precision (MEASURED_PEAKS.json has no FP64/FP32 FMA entry)" % (
                                     flop_first, flop_fused, algo_flop_step / n_tau_launches, tau_launch_s),
                         "reference_equivalent_tflops": FLOP_PER_VOIGT_REFERENCE * n_voigt_step / n_tau_launches / tau_launch_s / 1e12,
                         "march_steps_by_route": dict(zip(["near_gauss", "near", "far", "straddle", "slow"], [int(v) for v in routes])),
                         "tau_share_of_step": tau_launch_s * n_tau_launches / (elapsed / args.steps),
                         "k_tau_ms_per_rank": [round(v * 1e3, 3) for v in tau_launch_s_ranks],
                         "co_bound": "shared-memory wavefronts (l1tex data pipe ~50 % of peak) and issue slots (~57 %) "
                                     "run as hot as the FP64 pipe (~45 %): profiles/README.md"},
            "index_build": {"bound": "hbm", "limited_by": "latency and atomics: two particle passes with data-dependent cell walks, "
                            "then a per-list sort (DESIGN.md 6); the HBM figure is there for scale", "ms": index_s * 1e3, "algorithmic_bytes": index_bytes,
                            "achieved": index_bytes / index_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                            "frac": index_bytes / index_s / 1e9 / hbm_peak, "peak_source": hbm_src,
                            "share_of_step": index_s / (elapsed / args.steps)},
            "colden": colden, "flux_stats": flux_stats,
            "cpu_baseline": cpu, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "check_mean_tau": sanity,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def cpu_sample(w, steps=1, nsample=None):
    """The reference's own CPU implementation (oracle/_ref when built, else the C restatement) on a
    bounded sample of the workload: a regular subsample of the sightlines against the particles
    that reach them (prefiltered, untimed, as the reference host does: spectra.py:556-568)."""
    from oracle import Oracle, Reference
    try:
        impl, kind = Reference(), "reference"
    except (FileNotFoundError, OSError):
        impl, kind = Oracle(), "port"
    # all host threads, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1 to every rank)
    impl.set_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    cores = impl.threads()
    if nsample is None:
        # ~0.56 core-seconds per sightline at 256^3 (two lines); aim at 10-30 s of CPU work per step
        nsample = int(min(w["nlos"], max(32, 24 * cores * 256 // w["nside"])))
    sel = np.unique(np.linspace(0, w["nlos"] - 1, nsample).astype(np.int64))
    cofm = np.ascontiguousarray(w["cofm"][sel])
    axis = np.ascontiguousarray(w["axis"][sel])
    near = Oracle().near_lines(w["box"], w["pos"], w["h"], axis, cofm) if kind == "port" else \
        impl.near_lines(w["box"], w["pos"], w["h"], axis, cofm)
    sub = {k: np.ascontiguousarray(w[k][near]) for k in ("pos", "vel", "dens", "temp", "h")}
    best = None
    for _ in range(steps):
        t_step = 0.0
        for ln in w["lines"]:
            lam, gam, fosc, amu = LINES[ln]
            t0 = time.perf_counter()
            impl.compute_tau(w["nbins"], w["kernel"], w["box"], w["velfac"], w["atime"], lam, gam, fosc, amu, TAUTAIL,
                             sub["pos"], sub["vel"], sub["dens"], sub["temp"], sub["h"], axis, cofm)
            t_step += time.perf_counter() - t0
        best = t_step if best is None else min(best, t_step)
    return {"value": len(sel) / best, "unit": "spectra/s", "cores": cores, "kind": kind, "seconds_per_step": best,
            "sample": "%d of %d sightlines (regular subsample), %d prefiltered particles, %d line(s) each; prefilter untimed"
                      % (len(sel), w["nlos"], len(near), len(w["lines"]))}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = build_workload(args.workload, 0, 1)
    for _ in range(args.warmup):
        cpu_sample(w, steps=1, nsample=8)  # warm caches / thread pool on a tiny sample
    t0 = time.perf_counter()
    runs = [cpu_sample(w, steps=1) for _ in range(args.steps)]
    wall = time.perf_counter() - t0
    sec = float(np.mean([r["seconds_per_step"] for r in runs]))
    nsel = runs[0]["value"] * runs[0]["seconds_per_step"]
    value = nsel / sec
    cpu = dict(runs[0])
    cpu["value"] = value
    line = {"impl": "reference", "metric": "spectra_per_s", "value": value, "unit": "spectra/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "particles": int(w["npart"]), "pixels": int(w["nbins"]),
                       "lines": list(w["lines"]), "sph_kernel": {0: "tophat", 1: "cubic", 2: "voronoi", 3: "quintic"}[w["kernel"]],
                       "note": "reference C++ (OpenMP, all host threads) on a bounded sightline sample; wall %.1f s" % wall},
            "cpu_baseline": cpu, "e2e": {"value": value, "unit": "spectra/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
