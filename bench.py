#!/usr/bin/env python
"""Benchmark of the sightline-interpolation hot path (BASELINE.json metric: spectra/s and
sightline-particle pairs/s for Ly-alpha-forest tau).

  python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle/_ref)

A "step" is one pass of the hot path over the whole workload: candidate-index build + optical depth of
every line of every ion for every sightline.  Default workload = BASELINE.json configs[2], the largest
configuration that fits one GPU: synthetic 2x512^3 snapshot (134 M gas particles, SURVEY App. F generator),
100 000 random sightlines cycling through the three axes, H I 1215 + 1025 (fused), C IV 1548, Mg II 2796,
cubic-spline SPH kernel, 1 km/s pixels.

One process per GPU.  N > 1 = STRONG scaling of that one fixed workload through the product's own sharding
path (fake_spectra_b200.sharding.Sharder, the class Spectra(shard="sightlines") uses): one count pass
(fsb_count_pairs) balances contiguous sightline blocks by candidate pairs, the particle set is replicated in
every GPU's HBM, every rank interpolates its block, and the result rows are gathered so that every rank holds
the full array (what the reference's MPI Allreduce leaves behind, spectra.py:825-831).  Rank 0 then recomputes
a 128-sightline subsample on its own and checks the gathered rows bit for bit (`parity_check`).
Particle-sharded workloads (c4_*) keep the sightlines common and sum the FP64 partial arrays over NCCL.
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

# ion: (amu, density scale applied to the App. F boundary density: "x1e-4 for metal lines")
IONS = {"HI": (1.00794, 1.0), "CIV": (12.011, 1e-4), "MgII": (24.305, 1e-4)}
LINES = {  # ion, lambda (cm), Gamma (1/s), f_osc: SURVEY App. F (reference atom.dat)
    "HI1215": ("HI", 1215.6701e-8, 6.265e8, 0.4164),
    "HI1025": ("HI", 1025.7223e-8, 1.897e8, 0.07912),
    "CIV1548": ("CIV", 1548.2049e-8, 2.642e8, 0.1899),
    "MgII2796": ("MgII", 2796.3542699e-8, 2.68e8, 0.6155),
}
ALL4 = ("HI1215", "HI1025", "CIV1548", "MgII2796")
WORKLOADS = {
    # BASELINE.json configs[2]
    "c3_rand100k_3axes_4lines": dict(nside=512, numlos=100000, axis="cycle", lines=ALL4, kernel=1, res=1.0),
    # BASELINE.json configs[1]
    "c2_grid256_lya_lyb": dict(nside=256, nspec=256, lines=("HI1215", "HI1025"), kernel=1, res=1.0),
    # BASELINE.json configs[0]
    "c1_rand1000_lya": dict(nside=64, numlos=1000, axis=1, lines=("HI1215",), kernel=1, res=1.0),
    "mini_grid64_lya_lyb": dict(nside=64, nspec=64, lines=("HI1215", "HI1025"), kernel=1, res=1.0),
    "mini3_rand6k_3axes_4lines": dict(nside=64, numlos=6144, axis="cycle", lines=ALL4, kernel=1, res=1.0),
    # BASELINE.json configs[3] in miniature: Arepo-like top-hat kernel, particles sharded over the ranks
    # (nside^3 cells PER RANK of one common box), every rank computes all sightlines for its cells and the
    # FP64 tau arrays are summed over NCCL (the reference's MPI mode, spectra.py:825-831)
    "c4_tophat_pshard": dict(nside=256, numlos=16384, axis=1, lines=("HI1215",), kernel=0, res=1.0, shard="particles"),
    # BASELINE.json configs[3] at full size when run on 8 GPUs: 8 x 512^3 = 1024^3 cells in a 160 000 kpc/h box
    "c4_tophat_pshard_1024": dict(nside=512, numlos=16384, axis=1, lines=("HI1215",), kernel=0, res=1.0, shard="particles"),
    "mini_tophat_pshard": dict(nside=64, numlos=2048, axis=1, lines=("HI1215",), kernel=0, res=1.0, shard="particles"),
}
DEFAULT_WORKLOAD = "c3_rand100k_3axes_4lines"
# Algorithmic FP64 work per Voigt evaluation (one quadrature node of one pixel of one line), DESIGN.md section 5,
# FMA = 2 flop: first line of an ion 14 DFMA + 6 DMUL/DADD = 34 flop, every further fused line 8 DFMA = 16 flop.
# This is the unit of account of rounds 1 and 2 (the full-degree table expansion); builds that evaluate lower-degree
# coefficient polynomials where the damping parameter allows do less than this per evaluation and are still
# accounted at these figures, so that fractions are comparable across iterations (DESIGN.md says so too).
# 280 = the reference algorithm's figure, reported separately; it is not the roofline numerator.
FLOP_PER_VOIGT = 34.0
FLOP_PER_VOIGT_FUSED = 16.0
FLOP_PER_VOIGT_REFERENCE = 280.0
TAUTAIL = 1e-7          # reference spectra.py:135
KERNEL_NAMES = {0: "tophat", 1: "cubic", 2: "voronoi", 3: "quintic"}
UNSEGMENTED = 1 << 30   # fsb_params.seg_pairs: one work item per sightline (results independent of the partition)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--voigt", default="fast", choices=["fast", "exact"])
    ap.add_argument("--precision", default="fp64", choices=["fp64", "fp32"],
                    help="fp32 = the optional fast path (node sums in single precision, flux within 1e-5 of the reference)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the colden / flux statistics / weak-replica legs")
    ap.add_argument("--ref-lines", type=int, default=0, help="sightlines of the reference sample (0 = choose)")
    ap.add_argument("--reduce-blocks", type=int, default=4,
                    help="particle-sharded workloads: sightline blocks whose NCCL sums overlap the next block's kernel (1 = one sum at the end)")
    ap.add_argument("--gather", default="push", choices=["push", "nccl"],
                    help="N > 1, sightline-sharded: how the rows reach every rank: stored into the peers' arrays from inside "
                         "the tau kernel (push), or gathered afterwards with NCCL (sharding.Sharder.combine)")
    return ap.parse_args()


def ion_groups(lines):
    """[(ion, [line names])] in first-appearance order: lines of one ion share a pass."""
    groups = []
    for ln in lines:
        ion = LINES[ln][0]
        for g in groups:
            if g[0] == ion:
                g[1].append(ln)
                break
        else:
            groups.append((ion, [ln]))
    return groups


def build_workload(name, rank, world):
    """Host arrays of the whole workload (every rank builds the same ones from the same seeds, except for
    particle-sharded workloads where rank r holds its own shard of the cells)."""
    from fake_spectra_b200 import synthetic as syn
    w = dict(WORKLOADS[name])
    w.setdefault("shard", "sightlines")
    cos = syn.Cosmology()
    if w["shard"] == "particles":
        # rank r holds nside^3 cells (seed 42 + r) of a box sized for world * nside^3 cells; same sightlines everywhere
        box = syn.MEAN_SPACING * w["nside"] * world ** (1.0 / 3.0)
        d = syn.boundary_arrays(w["nside"], seed=42 + rank, kernel=w["kernel"], box=box)
    else:
        d = syn.boundary_arrays(w["nside"], seed=42, kernel=w["kernel"])
        box = d["box"]
    if "nspec" in w:
        cofm, axis = syn.grid_sightlines(box, w["nspec"], axis=1)
    else:
        cofm, axis = syn.random_sightlines(box, w["numlos"], seed=23, axis=w["axis"])
    velfac = float(cos.velfac)
    nbins = int(box * velfac / w["res"])
    w.update(d)
    w.update(cofm=np.ascontiguousarray(cofm), axis=np.ascontiguousarray(axis), velfac=velfac, atime=cos.atime,
             nbins=nbins, nlos=cofm.shape[0], npart=d["pos"].shape[0], name=name, groups=ion_groups(w["lines"]))
    return w


def ion_density(w, ion):
    """float32 ion density at the boundary: App. F density x the ion's abundance scale."""
    scale = IONS[ion][1]
    return w["dens"] if scale == 1.0 else (w["dens"] * np.float32(scale)).astype(np.float32)


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


def make_params(w, line, voigt, precision="fp64", seg_pairs=0):
    from fake_spectra_b200 import _lib
    ion, lam, gam, fosc = LINES[line]
    return _lib.make_params(w["nbins"], w["kernel"], w["box"], w["velfac"], w["atime"], lam, gam, fosc, IONS[ion][0], TAUTAIL,
                            voigt=_lib.VOIGT_EXACT if voigt == "exact" else _lib.VOIGT_FAST,
                            precision=_lib.PRECISION_FP32 if precision == "fp32" else _lib.PRECISION_FP64,
                            seg_pairs=seg_pairs)


def hbm_peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def run_b200(args):
    import torch
    import torch.distributed as dist
    from fake_spectra_b200 import _lib, native, _spectra_priv, sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))

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

    t_setup = time.perf_counter()
    w = build_workload(args.workload, rank, world)
    groups = w["groups"]
    nlines = len(w["lines"])
    pshard = w["shard"] == "particles"
    fp32 = args.precision == "fp32"
    # sightline-sharded runs pin one work item per sightline: every row is then computed by the same sequence of
    # operations whatever the partition, so N-GPU rows equal 1-GPU rows bit for bit (the parity check below)
    seg = UNSEGMENTED if (world > 1 and not pshard) else 0
    params = {ion: [make_params(w, ln, args.voigt, args.precision, seg) for ln in lns] for ion, lns in groups}
    # FP32 fast path, NEAR route with Gaussian, per Voigt evaluation: 12 FFMA + 6 = 30 flop; fused line 6 FFMA + 1 = 13
    flop_first, flop_fused = (30.0, 13.0) if fp32 else (FLOP_PER_VOIGT, FLOP_PER_VOIGT_FUSED)
    names = ("pos", "vel", "temp", "h", "cofm", "axis")
    t = {k: torch.from_numpy(w[k]).to(dev) for k in names}
    dens = {ion: torch.from_numpy(ion_density(w, ion)).to(dev) for ion, _ in groups}
    fma_peak = native.measure_fma_peak(not fp32)  # the FMA peak of the precision the node sums run in
    L, nbins = w["nlos"], w["nbins"]
    sharder = sharding.Sharder("particles" if pshard else "sightlines")
    setup_s = time.perf_counter() - t_setup

    state = {}
    ev = {"count": [], "index": [], "tau": [], "gather": [], "ion": []}

    def timed(key, fn, on):
        if not on:
            return fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r = fn()
        b.record()
        ev[key].append((a, b))
        return r

    def my_block(on):
        """This rank's sightline block: contiguous, balanced by candidate pairs.  Every rank counts the pairs of ALL
        sightlines against its slice of the (replicated) particle set, the int32 counts are summed over NCCL, and every
        rank derives the same edges from the same totals; the counts of its own block then size its candidate lists
        (fsb_index_build_counted), so no rank makes a counting pass over the whole particle set."""
        if pshard or world == 1:
            return slice(0, L), None

        def count():
            pe = sharding.even_blocks(w["npart"], world)
            ps = slice(int(pe[rank]), int(pe[rank + 1]))
            c = native.count_pairs(w["box"], t["pos"][ps], t["h"][ps], t["axis"], t["cofm"])
            dist.all_reduce(c, op=dist.ReduceOp.SUM)
            return c
        counts = timed("count", count, on)
        sharder.set_sightlines(L, weights=counts.cpu().numpy())
        sl = sharder.my_sightlines(L)
        return sl, counts[sl].contiguous()

    def step(time_parts=False, counters=None):
        sl, my_counts = my_block(time_parts)
        nloc = sl.stop - sl.start
        cofm, axis = t["cofm"][sl].contiguous(), t["axis"][sl].contiguous()
        out = state.get("out")
        if out is None or out.shape[1] != nloc:
            out = state["out"] = torch.empty((nlines, nloc, nbins), dtype=torch.float64, device=dev)
        out.zero_()
        idx = timed("index", lambda: native.CandidateIndex(w["box"], cofm, axis, t["pos"], t["h"], counts=my_counts), time_parts)

        use_push = world > 1 and not pshard and args.gather == "push" and counters is None
        if use_push and "peer" not in state:
            state["peer"] = native.PeerRows(nlines, L, nbins)  # full array on every rank, mapped into all the others

        def all_tau():
            r0 = 0
            marks = [torch.cuda.Event(enable_timing=True) for _ in range(len(groups) + 1)] if time_parts else None
            for gi, (ion, lns) in enumerate(groups):
                if marks:
                    marks[gi].record()
                idx.compute_tau(params[ion], t["pos"], t["vel"], dens[ion], t["temp"], t["h"], out=out[r0:r0 + len(lns)],
                                counters=None if counters is None else counters[gi],
                                push=state["peer"].push_spec(r0, sl.start) if use_push else None)
                r0 += len(lns)
            if marks:
                marks[-1].record()
                ev["ion"].append(marks)
        def all_tau_blocks():
            # particle-sharded: sightline blocks; the FP64 sum of a finished block over the ranks (NCCL, its own stream)
            # runs while the next block is computed
            ion, lns = groups[0]
            works = []
            for (b0, b1) in sharder.reduce_blocks(L, args.reduce_blocks):
                idx.compute_tau(params[ion], t["pos"], t["vel"], dens[ion], t["temp"], t["h"], out=out, lines=(b0, b1))
                works.append(sharder.sum_block_async(out, b0, b1))
            for wk in works:
                if wk is not None:
                    wk.wait()
        blocked = pshard and world > 1 and counters is None and len(groups) == 1 and len(groups[0][1]) == 1 and args.reduce_blocks > 1
        timed("tau", all_tau_blocks if blocked else all_tau, time_parts)
        state["npairs"], state["sl"] = idx.npairs, sl
        idx.free()
        if blocked:
            state["full"] = out
            return out
        if use_push:
            # the rows are already in every rank's array: wait until every rank's kernels have finished
            timed("gather", state["peer"].barrier, time_parts)
            state["full"] = state["peer"].full
        elif world > 1:
            # sightlines: gather the row blocks into the full array on every rank; particles: FP64 sum over NCCL
            state["full"] = timed("gather", lambda: sharder.combine(out, L, dim=1), time_parts)
        else:
            state["full"] = out
        return state["full"]

    # untimed counter pass: deterministic work counts of one step, per ion group
    ctrs = [torch.zeros(10, dtype=torch.int64, device=dev) for _ in groups]
    step(counters=ctrs)
    torch.cuda.synchronize()
    npairs = state["npairs"]
    cvals = [c.cpu().numpy() for c in ctrs]
    routes = sum(c[4:9] for c in cvals)
    # per group: evaluations of all its lines together (c[2]); the first line's share is not separable from a fused
    # launch, so count it from a single-line launch of the group's first line
    n_first, n_all = [], []
    for gi, (ion, lns) in enumerate(groups):
        n_all.append(int(cvals[gi][2]))
        if len(lns) == 1:
            n_first.append(int(cvals[gi][2]))
        else:
            sl = state["sl"]
            idx = native.CandidateIndex(w["box"], t["cofm"][sl].contiguous(), t["axis"][sl].contiguous(), t["pos"], t["h"])
            c1 = torch.zeros(10, dtype=torch.int64, device=dev)
            idx.compute_tau(params[ion][0], t["pos"], t["vel"], dens[ion], t["temp"], t["h"], out=state["out"][:1], counters=c1)
            torch.cuda.synchronize()
            idx.free()
            n_first.append(int(c1.cpu()[2]))
    n_voigt_step = int(sum(n_all))
    flop_per_group = [flop_first * a + flop_fused * (b - a) for a, b in zip(n_first, n_all)]
    algo_flop_step = sum(flop_per_group)
    route_names = ["near_gauss", "near", "far", "straddle", "slow_or_subsampled"]
    for _ in range(max(args.warmup, 0)):
        step()
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    launches0 = _lib.load().fsb_kernel_launches()
    barrier()
    torch.cuda.synchronize()
    if rank == 0:
        sampler.start()
    # working set per step: particles (36 B each + 4 B per further ion), candidate index (16 B per pair) and the output.
    # When it exceeds the 126 MB L2 nothing needs flushing; a small workload gets the L2 overwritten between timed steps
    work_bytes = w["npart"] * (36 + 4 * (len(groups) - 1)) + 16 * npairs + nlines * (state["sl"].stop - state["sl"].start) * nbins * 8
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if work_bytes < (256 << 20) else None
    step_events = []
    for _ in range(args.steps):
        if flush is not None:
            flush.fill_(1)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        step(time_parts=True)
        s1.record()
        step_events.append((s0, s1))
    torch.cuda.synchronize()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = _lib.load().fsb_kernel_launches() - launches0
    elapsed = max_over_ranks(sum(a.elapsed_time(b) for a, b in step_events) * 1e-3)

    def mean_ms(key):
        return float(np.mean([a.elapsed_time(b) for a, b in ev[key]])) if ev[key] else 0.0

    total_pairs = npairs if pshard and world == 1 else sum_over_ranks(float(npairs))
    ms_per_step = elapsed / args.steps * 1e3
    value = L * args.steps / elapsed
    pairs_per_s = total_pairs * nlines * args.steps / elapsed
    # k_tau launches per step: one per group of two fused lines of an ion
    n_tau_launches = sum((len(lns) + 1) // 2 for _, lns in groups)
    tau_ms_ranks = gather_ranks(mean_ms("tau"))
    tau_s = max(tau_ms_ranks) * 1e-3  # the slowest rank bounds the step
    flop_ranks = gather_ranks(float(algo_flop_step))
    achieved = max(flop_ranks) / tau_s / 1e12  # per GPU: the rank with the most work
    index_ms_ranks, count_ms_ranks, gather_ms_ranks = gather_ranks(mean_ms("index")), gather_ranks(mean_ms("count")), gather_ranks(mean_ms("gather"))
    # every cross-rank figure of the JSON line is gathered HERE, by all ranks (rank 0 alone prints)
    voigt_all = sum(gather_ranks(float(n_voigt_step)))
    pairs_ranks = gather_ranks(float(npairs))
    lines_ranks = gather_ranks(float(state["sl"].stop - state["sl"].start))
    full = state["full"]
    sanity = float(full[0].mean().item())
    hbm, hbm_src = hbm_peak()
    naxes = len(set(int(a) for a in np.unique(w["axis"])))
    index_s = max(index_ms_ranks) * 1e-3
    index_bytes = 16.0 * w["npart"] * naxes + 16.0 * npairs  # 16 B per particle and axis group read, 16 B per pair written

    # ---- in-run parity: a subsample of sightlines recomputed on one GPU -----------------------------------------
    parity = None
    nsub = min(128, L)
    sel = np.unique(np.linspace(0, L - 1, nsub).astype(np.int64))
    tsel = torch.from_numpy(sel).to(dev)
    if pshard and world > 1:
        # every rank contributes its particles near the subsample; rank 0 interpolates them all on one GPU
        csub, asub = t["cofm"][tsel].contiguous(), t["axis"][tsel].contiguous()
        near = native.near_lines(w["box"], t["pos"], t["h"], asub, csub).long()
        mine = {k: t[k][near].cpu().numpy() for k in ("pos", "vel", "temp", "h")}
        mine["dens"] = dens[groups[0][0]][near].cpu().numpy()
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        if rank == 0:
            cat = {k: torch.from_numpy(np.concatenate([p[k] for p in parts])).to(dev) for k in mine}
            idx = native.CandidateIndex(w["box"], csub, asub, cat["pos"], cat["h"])
            ref = idx.compute_tau(params[groups[0][0]], cat["pos"], cat["vel"], cat["dens"], cat["temp"], cat["h"])
            idx.free()
            got = full[:, tsel]
            m = ref != 0
            rel = float(((got - ref).abs()[m] / ref[m].abs()).max().item()) if bool(m.any()) else 0.0
            parity = {"rows": int(len(sel)), "mode": "particles", "bitwise": bool(torch.equal(got, ref)), "max_rel": rel,
                      "tolerance": 1e-12, "ok": rel <= 1e-12 and bool(torch.equal(got == 0, ref == 0)),
                      "how": "rank 0 gathers the particles near a regular subsample of sightlines from every rank and interpolates "
                             "them on one GPU; compared with the NCCL-summed rows"}
    elif rank == 0:
        csub, asub = t["cofm"][tsel].contiguous(), t["axis"][tsel].contiguous()
        idx = native.CandidateIndex(w["box"], csub, asub, t["pos"], t["h"])
        ref = torch.empty((nlines, len(sel), nbins), dtype=torch.float64, device=dev).zero_()
        r0 = 0
        for ion, lns in groups:
            pl = [make_params(w, ln, args.voigt, args.precision, UNSEGMENTED) for ln in lns]
            idx.compute_tau(pl, t["pos"], t["vel"], dens[ion], t["temp"], t["h"], out=ref[r0:r0 + len(lns)])
            r0 += len(lns)
        idx.free()
        got = full[:, tsel]
        bitwise = bool(torch.equal(got, ref))
        m = ref != 0
        rel = float(((got - ref).abs()[m] / ref[m].abs()).max().item()) if bool(m.any()) else 0.0
        unsegmented_main = world > 1 or L >= 4096
        parity = {"rows": int(len(sel)), "mode": "sightlines", "bitwise": bitwise, "max_rel": rel,
                  "tolerance": 0.0 if unsegmented_main else 1e-12, "ok": bitwise if unsegmented_main else rel <= 1e-12,
                  "how": "rank 0 recomputes a regular subsample of sightlines against the full particle set on one GPU, "
                         "one work item per sightline; compared with the rows of the %s" % (
                             "gathered full array" if world > 1 else "full run")}
    del tsel

    extras = {}
    if rank == 0 and not pshard and not args.no_extras:
        extras.update(extra_legs(args, w, t, dens, params, native, dev, hbm))

    # ---- weak-scaling replicas (secondary): every rank runs the whole workload on its own -------------------------
    weak = None
    if world > 1 and not pshard and not args.no_extras:
        o2 = torch.empty((nlines, L, nbins), dtype=torch.float64, device=dev)

        def replica():
            o2.zero_()
            idx = native.CandidateIndex(w["box"], t["cofm"], t["axis"], t["pos"], t["h"])
            r0 = 0
            for ion, lns in groups:
                idx.compute_tau(params[ion], t["pos"], t["vel"], dens[ion], t["temp"], t["h"], out=o2[r0:r0 + len(lns)])
                r0 += len(lns)
            idx.free()
        replica()
        barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        nrep = 2
        for _ in range(nrep):
            replica()
        b.record()
        torch.cuda.synchronize()
        rep_s = max_over_ranks(a.elapsed_time(b) * 1e-3)
        weak = {"scaling": "weak", "what": "every rank interpolates the whole workload on its own (N independent replicas)",
                "value": world * L * nrep / rep_s, "unit": "spectra/s", "ms_per_step": rep_s / nrep * 1e3}
        del o2

    # ---- end to end through the reference-facing boundary: host buffers in, host buffers out ----------------------
    e2e = None
    if not args.no_e2e:
        e2e = e2e_leg(args, w, state, groups, params, dens, t, world, rank, dev, dist, barrier, max_over_ranks, pshard, sharder)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_sample(w, nsample=args.ref_lines or None)

    if rank == 0:
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "tau_traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get(args.workload)
            except Exception:
                traffic = None
        step_ms = elapsed / args.steps * 1e3
        line = {
            "metric": "spectra_per_s", "value": value, "unit": "spectra/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak" if pshard else "strong",
            "vs_baseline": None, "dtype": "f32" if fp32 else "f64", "data": "synthetic",
            "config": {"workload": args.workload, "particles": int(w["npart"]), "sightlines": int(L),
                       "precision": args.precision, "pixels": int(nbins), "pixel_kms": w["res"], "lines": list(w["lines"]),
                       "ion_passes": [[ion, lns] for ion, lns in groups], "sph_kernel": KERNEL_NAMES[w["kernel"]], "voigt": args.voigt,
                       "parallelism": ("particle-sharded x%d (%d cells per rank, box grows with N), FP64 NCCL sum of the tau array each step, "
                                       "in %d sightline blocks overlapped with the kernel of the next block" % (world, w["npart"], args.reduce_blocks)) if pshard else
                                      ("sightline-sharded x%d through sharding.Sharder: pair-balanced contiguous blocks of ONE fixed "
                                       "sightline set, particles replicated, rows delivered to every rank (%s)" % (
                                           world, "stored into the peers' arrays over NVLink from inside the tau kernel" if args.gather == "push"
                                           else "NCCL gather after the kernels")),
                       "l2": ("inputs and outputs larger than L2 (particles %.2f GB, tau %.2f GB on this rank)" if flush is None else
                              "working set fits the L2 (particles %.3f GB, tau %.3f GB): a 256 MB buffer is overwritten "
                              "between timed steps, outside the timed intervals") % (
                           w["npart"] * 36 / 1e9, nlines * (state["sl"].stop - state["sl"].start) * nbins * 8 / 1e9),
                       "step": "(N>1: pair count + block edges,) index build, tau of all lines of all ions for every sightline"
                               "(, N>1: gather of the rows), inputs resident in HBM",
                       "setup_s": round(setup_s, 1)},
            "pairs_per_s": pairs_per_s, "pairs_per_step": total_pairs, "voigt_evals_per_step": voigt_all,
            "voigt_evals_per_s": voigt_all * args.steps / elapsed,
            "roofline": {"bound": "fp32" if fp32 else "fp64", "kernel": "k_tau", "achieved": achieved, "peak": fma_peak, "unit": "TFLOP/s",
                         "frac": achieved / fma_peak, "traffic": traffic,
                         "note": ("algorithmic FP32 flop" if fp32 else "algorithmic FP64 flop") + " of the profile evaluation (DESIGN.md 5): %.0f per Voigt "
                                 "evaluation of the first line of an ion + %.0f per evaluation of each fused line = %.3e flop in the %d k_tau launches "
                                 "of a step on the busiest rank / their summed time %.4f s (CUDA events on the launching stream, timed region); "
                                 "peak = FMA rate measured on this device by fsb_measure_fma_peak in the same precision "
                                 "(MEASURED_PEAKS.json has no FP64/FP32 FMA entry)" % (flop_first, flop_fused, max(flop_ranks), n_tau_launches, tau_s),
                         "reference_equivalent_tflops": FLOP_PER_VOIGT_REFERENCE * n_voigt_step / tau_s / 1e12,
                         "march_steps_by_route": dict(zip(route_names, [int(v) for v in routes])),
                         "tau_share_of_step": tau_s * 1e3 / step_ms,
                         "k_tau_ms_per_rank": [round(v, 3) for v in tau_ms_ranks],
                         "k_tau_ms_per_ion_pass": {groups[gi][0]: round(float(np.mean([m[gi].elapsed_time(m[gi + 1]) for m in ev["ion"]])), 3)
                                                   for gi in range(len(groups))} if ev["ion"] else None,
                         # the same fraction per launch (this rank's flop and time): hydrogen (two fused lines, long damping
                         # wings) against the metal ions (one weak line each: a few march steps per particle)
                         "frac_per_ion_pass": {groups[gi][0]: round(flop_per_group[gi] / (float(np.mean([m[gi].elapsed_time(m[gi + 1]) for m in ev["ion"]])) * 1e-3)
                                                                     / 1e12 / fma_peak, 4) for gi in range(len(groups))} if ev["ion"] else None,
                         "march_steps_per_ion_pass": {groups[gi][0]: dict(zip(route_names, [int(v) for v in cvals[gi][4:9]]))
                                                      for gi in range(len(groups))},
                         "pairs_per_ion_pass": int(npairs),
                         "frac_note": "the step's fraction is the time-weighted mix of its launches: the hydrogen launch (two fused lines, "
                                      "7 march steps per candidate pair; the whole of config C2) runs at frac_per_ion_pass['HI'], the "
                                      "single weak metal lines finish a pair in about one march step and are bound by per-particle fixed "
                                      "cost and instruction fetch (DESIGN.md section 5)"},
            "index_build": {"bound": "hbm", "ms": index_s * 1e3, "ms_per_rank": [round(v, 3) for v in index_ms_ranks],
                            "algorithmic_bytes": index_bytes, "achieved": index_bytes / index_s / 1e9, "peak": hbm, "unit": "GB/s",
                            "frac": index_bytes / index_s / 1e9 / hbm, "peak_source": hbm_src, "share_of_step": index_s * 1e3 / step_ms,
                            "note": "16 B per particle per axis group read + 16 B per pair written (particle, dr2, traversal order)"},
            "multi_gpu": None if world == 1 else {"count_pairs_ms_per_rank": [round(v, 3) for v in count_ms_ranks],
                                                  "gather_ms_per_rank": [round(v, 3) for v in gather_ms_ranks],
                                                  "pairs_per_rank": [int(v) for v in pairs_ranks],
                                                  "sightlines_per_rank": [int(v) for v in lines_ranks],
                                                  "weak_replicas": weak},
            "parity_check": parity,
            "cpu_baseline": cpu, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "check_mean_tau": sanity,
        }
        line.update(extras)
        print(json.dumps(line))
    if "peer" in state:
        state["full"] = None
        state["peer"].close()
    if world > 1:
        dist.destroy_process_group()


def extra_legs(args, w, t, dens, params, native, dev, hbm):
    """Column density (K3) and the flux statistics (row f2) on a bounded block of the workload's sightlines."""
    import torch
    from fake_spectra_b200 import fluxstatistics as fstat
    nb = min(w["nlos"], 32768)
    cofm, axis = t["cofm"][:nb].contiguous(), t["axis"][:nb].contiguous()
    ion, lns = w["groups"][0]
    prm = params[ion][0]
    idx = native.CandidateIndex(w["box"], cofm, axis, t["pos"], t["h"])
    cout = torch.zeros((nb, w["nbins"]), dtype=torch.float64, device=dev)
    ctr = torch.zeros(10, dtype=torch.int64, device=dev)
    idx.compute_colden(prm, t["pos"], dens[ion], t["h"], out=cout, counters=ctr)
    torch.cuda.synchronize()
    cpix = int(ctr.cpu()[1])
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for _ in range(3):
        idx.compute_colden(prm, t["pos"], dens[ion], t["h"], out=cout)
    c1.record()
    torch.cuda.synchronize()
    col_s = c0.elapsed_time(c1) * 1e-3 / 3
    # algorithmic traffic of one column: 32 B gathered per pair (index entry 12 B, position 4, h 4, density 4, dr2 8) +
    # the output array written once; algorithmic FP64 work: 9 kernel evaluations per pixel integral, ~12 flop each
    col_bytes = 32.0 * idx.npairs + 8.0 * nb * w["nbins"]
    colden = {"kernel": "k_colden", "sightlines": nb, "ms": col_s * 1e3, "pairs_per_s": idx.npairs / col_s, "pixels": cpix,
              "kernel_integrals_per_s": cpix / col_s,
              "roofline": {"bound": "hbm", "algorithmic_bytes": col_bytes, "achieved": col_bytes / col_s / 1e9, "peak": hbm,
                           "unit": "GB/s", "frac": col_bytes / col_s / 1e9 / hbm,
                           "fp64_tflops": cpix * 9 * 12.0 / col_s / 1e12,
                           "note": "one weight column; 32 B gathered per pair + the [nlos, nbins] array written once; the pass is "
                                   "bound by FP64 issue (9-node trapezoid of the SPH kernel per pixel, absorption.cpp:53-74), "
                                   "fp64_tflops counts 9 x 12 flop per pixel integral"}}
    tau = idx.compute_tau(params[ion], t["pos"], t["vel"], dens[ion], t["temp"], t["h"])
    idx.free()
    flat = tau.view(-1)
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
    scale = fstat.mean_flux(tau[0], 0.7)
    torch.cuda.synchronize()
    newton_s = time.perf_counter() - t0
    fstat.flux_power(tau[0][:256], w["vmax"] if "vmax" in w else 1.0)  # warm-up (tables, attributes)
    torch.cuda.synchronize()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    fstat.flux_power(tau[0], w["vmax"] if "vmax" in w else 1.0)
    p1.record()
    torch.cuda.synchronize()
    power_s = p0.elapsed_time(p1) * 1e-3
    npx = int(tau[0].shape[1])
    n2f = max(d for d in range(1, int(npx ** 0.5) + 1) if npx % d == 0)
    n1f = npx // n2f
    # real FMAs of the two-level transform per sightline: stage A 2 x n2 x (n1/2 + 1) x n1, stage B 4 x (n/2 + 1) x n2
    power_fma = float(tau[0].shape[0]) * (2.0 * n2f * (n1f // 2 + 1) * n1f + 4.0 * (npx // 2 + 1) * n2f)
    flux_power = {"kernel": "k_flux_power", "sightlines": int(tau[0].shape[0]), "pixels": npx, "factors": [n1f, n2f], "ms": power_s * 1e3,
                  "bound": "fp64", "achieved": 2.0 * power_fma / power_s / 1e12, "unit": "TFLOP/s",
                  "note": "own two-level Fourier sum in shared memory (no FFT library), flux contrast formed on the fly; includes the "
                          "mean-flux reduction and the host round trip of the call"}
    flux_stats = {"bound": "hbm", "kernel": "k_flux_sums", "pixels": int(flat.numel()), "ms_per_pass": pass_s * 1e3,
                  "algorithmic_bytes": 8.0 * flat.numel(), "achieved": 8.0 * flat.numel() / pass_s / 1e9, "peak": hbm,
                  "unit": "GB/s", "frac": 8.0 * flat.numel() / pass_s / 1e9 / hbm,
                  "mean_flux": sums[0] / max(sums[2], 1),
                  "rescale_to_0.7": {"scale": scale, "ms": newton_s * 1e3, "pixels": int(tau[0].numel())}}
    return {"colden": colden, "flux_stats": flux_stats, "flux_power": flux_power}


def e2e_leg(args, w, state, groups, params, dens, t, world, rank, dev, dist, barrier, max_over_ranks, pshard, sharder):
    """The same step through the reference-facing boundary: one `_Particle_Interpolate` call per ion with HOST
    buffers (the reference's Python makes one call per line; the extra_lines extension fuses the lines of an ion),
    host->device copies of every input and the device->host copy of the result inside the timed region.  N > 1:
    every rank makes the calls for its block of sightlines (its rows land in its own host buffer).
    Measured twice: with page-locked buffers, and with plain pageable numpy arrays and out=None (the literal
    drop-in signature)."""
    import torch
    from fake_spectra_b200 import _lib, _spectra_priv, native
    fp32 = args.precision == "fp32"
    vg = _lib.VOIGT_EXACT if args.voigt == "exact" else _lib.VOIGT_FAST
    prec = _lib.PRECISION_FP32 if fp32 else _lib.PRECISION_FP64
    L, nbins = w["nlos"], w["nbins"]
    sl = state["sl"]
    nloc = sl.stop - sl.start
    host = {k: w[k] for k in ("pos", "vel", "temp", "h")}
    host["cofm"], host["axis"] = np.ascontiguousarray(w["cofm"][sl]), np.ascontiguousarray(w["axis"][sl])
    hdens = {ion: ion_density(w, ion) for ion, _ in groups}
    maxl = max(len(lns) for _, lns in groups)
    seg = UNSEGMENTED if (world > 1 and not pshard) else 0

    def calls(arr, dn, outbuf, one_call=False):
        res = []
        if one_call and len(groups) > 1:
            # every ion from one upload and one candidate index (extra_ions -> fsb_particle_interpolate_ions_host)
            ion0, lns0 = groups[0]
            _, lam, gam, fosc = LINES[lns0[0]]
            others = [(dn[ion], IONS[ion][0], [LINES[ln][1:] for ln in lns]) for ion, lns in groups[1:]]
            nl_all = sum(len(lns) for _, lns in groups)
            o = None if outbuf is None else outbuf[:nl_all * nloc * nbins]
            r = _spectra_priv._Particle_Interpolate(
                1, nbins, w["kernel"], w["box"], w["velfac"], w["atime"], lam, gam, fosc, IONS[ion0][0], TAUTAIL,
                arr["pos"], arr["vel"], dn[ion0], arr["temp"], arr["h"], arr["axis"], arr["cofm"], voigt=vg, precision=prec,
                out=o, extra_lines=[LINES[ln][1:] for ln in lns0[1:]], seg_pairs=seg, extra_ions=others)
            r = r.reshape(nl_all, nloc, nbins)
            first = 0
            for _, lns in groups:
                res.append(float(np.mean(r[first][: min(64, nloc)])))
                first += len(lns)
            return res
        for ion, lns in groups:
            _, lam, gam, fosc = LINES[lns[0]]
            extra = [LINES[ln][1:] for ln in lns[1:]]
            o = None if outbuf is None else outbuf[:len(lns) * nloc * nbins]
            r = _spectra_priv._Particle_Interpolate(
                1, nbins, w["kernel"], w["box"], w["velfac"], w["atime"], lam, gam, fosc, IONS[ion][0], TAUTAIL,
                arr["pos"], arr["vel"], dn[ion], arr["temp"], arr["h"], arr["axis"], arr["cofm"], voigt=vg, precision=prec,
                out=o, extra_lines=extra, seg_pairs=seg)
            res.append(float(np.mean(r.reshape(-1, nbins)[: min(64, nloc)])))
        return res

    def pshard_step(pin, pdens, hout):
        # particle-sharded: upload this rank's particles, interpolate all sightlines, sum over ranks (NCCL), read back
        dv = {k: pin[k].to(dev, non_blocking=True) for k in pin}
        dd = pdens.to(dev, non_blocking=True)
        o = torch.zeros((len(w["lines"]), L, nbins), dtype=torch.float64, device=dev)
        idx = native.CandidateIndex(w["box"], dv["cofm"], dv["axis"], dv["pos"], dv["h"])
        idx.compute_tau(params[groups[0][0]], dv["pos"], dv["vel"], dd, dv["temp"], dv["h"], out=o)
        idx.free()
        if world > 1:
            o = sharder.combine(o, L, dim=1)
        hout.copy_(o.view(-1), non_blocking=True)
        torch.cuda.synchronize()
        return [float(hout[: 64 * nbins].mean())]

    esteps = max(1, min(args.steps, 5))  # bounded: a C3 step moves 14 GB in and 29 GB out through the host

    def run(fn):
        fn()  # warm-up (allocations, first touch of the buffers)
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(esteps):
            chk = fn()
        torch.cuda.synchronize()
        el = max_over_ranks(time.perf_counter() - t0)
        barrier()
        return el, chk

    pin = {k: torch.from_numpy(host[k]).pin_memory() for k in host}
    pdens = {ion: torch.from_numpy(hdens[ion]).pin_memory() for ion in hdens}
    one_call = (not pshard) and len(groups) > 1
    hout = torch.empty((len(w["lines"]) if one_call else maxl) * nloc * nbins if not pshard else len(w["lines"]) * L * nbins,
                       dtype=torch.float64).pin_memory()
    if pshard:
        el, chk = run(lambda: pshard_step(pin, pdens[groups[0][0]], hout))
        ncalls = 1
    else:
        pa = {k: pin[k].numpy() for k in pin}
        pd = {ion: pdens[ion].numpy() for ion in pdens}
        ho = hout.numpy()
        el, chk = run(lambda: calls(pa, pd, ho, one_call))
        ncalls = 1 if one_call else len(groups)
        if one_call:  # the same through one call per ion (particles uploaded and the index built once per ion)
            el_ion, chk_ion = run(lambda: calls(pa, pd, ho))
    h2d = ncalls * sum(int(host[k].nbytes) for k in host) + (0 if pshard else 0)
    h2d += sum(int(hdens[ion].nbytes) for ion, _ in groups) if not pshard else int(hdens[groups[0][0]].nbytes)
    d2h = len(w["lines"]) * (L if pshard else nloc) * nbins * 8
    e2e = {"value": L * esteps / el, "unit": "spectra/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "ms_per_step": el / esteps * 1e3, "steps": esteps, "buffers": "page-locked",
           "call": ("pinned host -> device, native.CandidateIndex + compute_tau, NCCL sum, device -> pinned host" if pshard else
                    ("1 x _spectra_priv._Particle_Interpolate(host buffers, extra_lines, extra_ions) -> fsb_particle_interpolate_ions_host: "
                     "every line of every ion from one upload and one candidate index%s" if one_call else
                     "%d x _spectra_priv._Particle_Interpolate(host buffers, extra_lines) -> fsb_particle_interpolate_multi_host, "
                     "one call per ion%%s" % ncalls) % ("" if world == 1 else ", this rank's block of sightlines")),
           "mean_tau_check": chk}
    if not pshard and one_call:
        e2e["one_call_per_ion"] = {"value": L * esteps / el_ion, "unit": "spectra/s", "ms_per_step": el_ion / esteps * 1e3,
                                   "h2d_bytes_per_step": int(len(groups) * sum(int(host[k].nbytes) for k in host)
                                                             + sum(int(hdens[ion].nbytes) for ion, _ in groups)),
                                   "buffers": "page-locked", "mean_tau_check": chk_ion}
    del pin, pdens, hout
    if not pshard:
        el2, chk2 = run(lambda: calls(host, hdens, None))
        e2e["pageable"] = {"value": L * esteps / el2, "unit": "spectra/s", "ms_per_step": el2 / esteps * 1e3,
                           "buffers": "pageable numpy arrays in, out=None (np.empty): the literal drop-in signature",
                           "mean_tau_check": chk2}
    return e2e


def core_seconds_per_sightline(w):
    """Rough CPU cost of one sightline (all lines) in core-seconds, to size the reference samples: ~0.4 core-s per
    H I line at 256^3 (measured), proportional to the list length (~ nside); metal lines are narrower (~0.6 of H I)."""
    per = 0.0
    for ion, lns in w["groups"]:
        per += (1.0 if ion == "HI" else 0.6) * len(lns)
    return 0.4 * per * w["nside"] / 256.0


def cpu_sample(w, nsample=None, shift=0.0):
    """The reference's own CPU implementation (oracle/_ref when built, else the C restatement) on a bounded
    sample of the workload: a regular subsample of the sightlines (offset by `shift` of the sampling stride) against
    the FULL particle set, every line of every ion, one compute_tau call per line as the reference's
    Python makes them.  Each call builds the candidate index inside (part_int.cpp:22); that search alone
    (IndexTable::get_near_particles) is timed once, separately, on the same sample."""
    from oracle import Oracle, Reference
    try:
        impl, kind = Reference(), "reference"
    except (FileNotFoundError, OSError):
        impl, kind = Oracle(), "port"
    # all host threads, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1 to every rank)
    impl.set_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    cores = impl.threads()
    nlines = len(w["lines"])
    if nsample is None:  # about 20 s of CPU work
        nsample = int(min(w["nlos"], max(32, 20.0 * cores / core_seconds_per_sightline(w))))
    nsample = int(min(nsample, w["nlos"]))
    stride = w["nlos"] / float(nsample)
    sel = np.unique(np.minimum(((np.arange(nsample) + (shift % 1.0)) * stride).astype(np.int64), w["nlos"] - 1))
    cofm = np.ascontiguousarray(w["cofm"][sel])
    axis = np.ascontiguousarray(w["axis"][sel])
    t0 = time.perf_counter()
    _, part, _ = impl.near_particles(cofm, axis, w["box"], w["pos"], w["h"])
    t_index = impl.last_seconds if kind == "reference" else (time.perf_counter() - t0) / 2  # (the port runs the search twice)
    t_step = 0.0
    for ion, lns in w["groups"]:
        dn = ion_density(w, ion)
        for ln in lns:
            _, lam, gam, fosc = LINES[ln]
            t0 = time.perf_counter()
            impl.compute_tau(w["nbins"], w["kernel"], w["box"], w["velfac"], w["atime"], lam, gam, fosc, IONS[ion][0], TAUTAIL,
                             w["pos"], w["vel"], dn, w["temp"], w["h"], axis, cofm)
            t_step += time.perf_counter() - t0
    return {"value": len(sel) / t_step, "unit": "spectra/s", "cores": cores, "kind": kind, "seconds_per_step": t_step,
            "index_seconds": t_index, "index_share": min(1.0, t_index * nlines / t_step),
            "value_excluding_index": len(sel) / max(t_step - t_index * nlines, 1e-9),
            "pairs_in_sample": int(len(part)), "sightlines_in_sample": int(len(sel)),
            "sample": "%d of %d sightlines (regular subsample) against all %d particles, %d line(s) each, one compute_tau call per "
                      "line; every call rebuilds the candidate index over the full particle set (part_int.cpp:22): that search, timed "
                      "once on its own, is index_seconds (its share of the step is index_share; the full workload amortises it over "
                      "%.0fx more sightlines, value_excluding_index brackets that)" % (len(sel), w["nlos"], w["npart"], nlines, w["nlos"] / float(len(sel)))}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = build_workload(args.workload, 0, 1)
    for _ in range(min(args.warmup, 2)):
        cpu_sample(w, nsample=4)  # warm caches / thread pool on a tiny sample
    # every step takes a different regular subsample; over the K steps they cover >= 1000 distinct sightlines when
    # the workload has them, within a few minutes of wall time
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    per_step = args.ref_lines or int(min(max(32, -(-1000 // max(args.steps, 1))), max(32, 60.0 * cores / core_seconds_per_sightline(w))))
    t0 = time.perf_counter()
    runs = [cpu_sample(w, nsample=per_step, shift=k / float(max(args.steps, 1))) for k in range(args.steps)]
    wall = time.perf_counter() - t0
    sec = float(np.sum([r["seconds_per_step"] for r in runs]))
    nsel = int(np.sum([r["sightlines_in_sample"] for r in runs]))
    value = nsel / sec
    cpu = dict(runs[0])
    cpu.update(value=value, seconds_per_step=sec / args.steps, index_seconds=float(np.mean([r["index_seconds"] for r in runs])),
               index_share=float(np.mean([r["index_share"] for r in runs])),
               value_excluding_index=nsel / max(sec - sum(r["index_seconds"] for r in runs) * len(w["lines"]), 1e-9),
               sightlines_all_steps=nsel)
    line = {"impl": "reference", "metric": "spectra_per_s", "value": value, "unit": "spectra/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec / args.steps * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "particles": int(w["npart"]), "sightlines": int(w["nlos"]), "pixels": int(w["nbins"]),
                       "lines": list(w["lines"]), "sph_kernel": KERNEL_NAMES[w["kernel"]],
                       "note": "reference C++ (OpenMP, all %d host threads of this box); every step interpolates a different regular "
                               "subsample of %d sightlines against the full particle set (%d distinct sightlines over the %d steps); "
                               "wall %.1f s" % (cpu["cores"], per_step, nsel, args.steps, wall)},
            "cpu_baseline": cpu, "e2e": {"value": value, "unit": "spectra/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
