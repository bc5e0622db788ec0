#!/bin/bash
# One GPU-box visit (round 2 of the build; output tags r3<letter>).
# Usage (here): gpurun --timeout 1500 -- 'bash scripts/r3_round.sh <tag> <steps...>'
TAG=${1:-r3a}; shift
STEPS=${@:-"test bench"}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
free -g > $OUT/host.txt; nproc >> $OUT/host.txt
B="python bench.py"
for S in $STEPS; do
case $S in
test) echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" >> $OUT/pytest_gpu.log; tail -8 $OUT/pytest_gpu.log;;
testfast) echo "== pytest -m gpu (parity only)"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_host.py -m gpu -x -q > $OUT/pytest_gpu_fast.log 2>&1; echo "rc=$?" >> $OUT/pytest_gpu_fast.log; tail -5 $OUT/pytest_gpu_fast.log;;
smoke) echo "== smoke"; timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; tail -3 $OUT/smoke.log;;
bench) echo "== bench default (c3)"; timeout 1500 $B --steps 3 --warmup 2 > $OUT/bench_c3.json 2> $OUT/bench_c3.err; tail -c 1500 $OUT/bench_c3.json; tail -3 $OUT/bench_c3.err;;
benchq) echo "== bench c3 quick"; timeout 900 $B --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > $OUT/bench_c3q.json 2> $OUT/bench_c3q.err; tail -c 1200 $OUT/bench_c3q.json; tail -3 $OUT/bench_c3q.err;;
c2) echo "== bench c2"; timeout 900 $B --workload c2_grid256_lya_lyb --steps 3 --warmup 3 > $OUT/bench_c2.json 2> $OUT/bench_c2.err; tail -c 1200 $OUT/bench_c2.json; tail -3 $OUT/bench_c2.err;;
c2q) echo "== bench c2 quick"; timeout 600 $B --workload c2_grid256_lya_lyb --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-extras > $OUT/bench_c2q.json 2> $OUT/bench_c2q.err; tail -c 900 $OUT/bench_c2q.json; tail -3 $OUT/bench_c2q.err;;
c2x) echo "== bench c2 with the extra legs (colden, flux statistics, flux power)"; timeout 600 $B --workload c2_grid256_lya_lyb --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/bench_c2x.json 2> $OUT/bench_c2x.err; python -c "
import json,sys
d=json.loads(open('$OUT/bench_c2x.json').read().strip().splitlines()[-1])
print(json.dumps({k:d[k] for k in ('colden','flux_stats','flux_power') if k in d})[:1500])"; tail -3 $OUT/bench_c2x.err;;
refbench) echo "== bench --impl reference"; timeout 1200 $B --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; tail -c 800 $OUT/bench_reference.json;;
variants) echo "== variants (c2 quick per library)"; for L in fake_spectra_b200/libfsb200*.so; do echo "-- $L"; FSB200_LIB=$PWD/$L timeout 600 $B --workload ${VW:-c2_grid256_lya_lyb} --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-extras 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; print(json.dumps({'lib':'$L','ms':d['ms_per_step'],'tau_ms':r['k_tau_ms_per_rank'],'frac':r['frac'],'index_ms':d['index_build']['ms'],'parity':d['parity_check'],'mean_tau':d['check_mean_tau']}))" | tee -a $OUT/variants.jsonl; done;;
launches) echo "== ncu launch list (c2)"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_c2.csv \
    $B --workload c2_grid256_lya_lyb --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > $OUT/launches_c2.log 2>&1; tail -3 $OUT/launches_c2.log | cut -c1-300;;
launches3) echo "== ncu launch list (c3)"
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_c3.csv \
    $B --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > $OUT/launches_c3.log 2>&1; tail -3 $OUT/launches_c3.log | cut -c1-300;;
ncu) echo "== ncu full k_tau (c2, the timed launch)"
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_tau -s 3 -c 1 -o $OUT/prof_tau_c2 \
    $B --workload c2_grid256_lya_lyb --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > $OUT/prof_tau_c2.log 2>&1; tail -2 $OUT/prof_tau_c2.log | cut -c1-300;;
ncu3) echo "== ncu full k_tau, the three launches (HI, CIV, MgII) of one C3 step"
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_tau -s 4 -c 3 -o $OUT/prof_tau_c3 \
    $B --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > $OUT/prof_tau_c3.log 2>&1; tail -2 $OUT/prof_tau_c3.log | cut -c1-300;;
ncumetal) echo "== ncu full k_tau NL=1 (metal lines, mini3 workload)"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tau -s 8 -c 2 -o $OUT/prof_tau_metal \
    $B --workload mini3_rand6k_3axes_4lines --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > $OUT/prof_tau_metal.log 2>&1; tail -2 $OUT/prof_tau_metal.log | cut -c1-300;;
ncuidx) echo "== ncu full index + colden kernels (c2)"
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_pairs|k_sort|k_colden|k_bin|k_fill|k_cand' -s 0 -c 8 -o $OUT/prof_idx_c2 \
    $B --workload c2_grid256_lya_lyb --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > $OUT/prof_idx_c2.log 2>&1; tail -2 $OUT/prof_idx_c2.log | cut -c1-300;;
ncucol) echo "== ncu full k_colden (c2, first launch of the extras leg)"
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_colden -s 0 -c 1 -o $OUT/prof_colden_c2 \
    $B --workload c2_grid256_lya_lyb --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > $OUT/prof_colden_c2.log 2>&1; tail -2 $OUT/prof_colden_c2.log | cut -c1-300;;
multi) NG=${NG:-2}; WL=${WL:-c2_grid256_lya_lyb}; echo "== bench $WL on $NG GPUs (strong scaling)"
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --workload $WL --steps 3 --warmup 2 > $OUT/bench_${WL}_${NG}gpu.json 2> $OUT/bench_${WL}_${NG}gpu.err; tail -c 2500 $OUT/bench_${WL}_${NG}gpu.json; tail -5 $OUT/bench_${WL}_${NG}gpu.err;;
testmulti) echo "== pytest multi-GPU"; timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s > $OUT/pytest_gpu_multi.log 2>&1; echo "rc=$?" >> $OUT/pytest_gpu_multi.log; tail -12 $OUT/pytest_gpu_multi.log | cut -c1-600;;
sanitizer) echo "== compute-sanitizer"
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_prep.py tests/test_gpu_stats.py tests/test_gpu_blocks.py -m gpu -x -q -k "golden or tiny or empty or voronoi or several_ions or prep_matches or plateau or own_transform or damped or blocked or zero_density" > $OUT/sanitizer_memcheck.log 2>&1; echo "rc=$?" >> $OUT/sanitizer_memcheck.log; tail -4 $OUT/sanitizer_memcheck.log
  timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stats.py -m gpu -x -q -k "candidate_lists_golden or tiny or own_transform or count_pairs" > $OUT/sanitizer_racecheck.log 2>&1; echo "rc=$?" >> $OUT/sanitizer_racecheck.log; tail -4 $OUT/sanitizer_racecheck.log;;
esac
done
ls -la $OUT
