#!/bin/bash
# One GPU-box visit: parity tests, smoke, probe (optionally over tuning variants), bench, ncu launch
# list and a full capture of k_tau.
# Usage (here): gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh <tag> [steps...]'
TAG=${1:-r01}; shift
STEPS=${@:-"test smoke probe bench launches ncu"}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
for S in $STEPS; do
case $S in
test) echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" >> $OUT/pytest_gpu.log; tail -15 $OUT/pytest_gpu.log;;
smoke) echo "== smoke"; timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; tail -3 $OUT/smoke.log;;
probe) echo "== probe"; timeout 600 python scripts/gpu_probe.py c1 c2s > $OUT/probe.log 2>&1; tail -6 $OUT/probe.log;;
variants) echo "== variants"; for L in fake_spectra_b200/libfsb200*.so; do echo "-- $L"; FSB200_LIB=$PWD/$L timeout 600 python scripts/gpu_probe.py ${PROBE:-c1 c2s} 2>&1 | tail -3 | tee -a $OUT/variants.log; done;;
refbench) echo "== bench --impl reference"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; tail -1 $OUT/bench_reference.json | cut -c1-400;;
bench) echo "== bench default"; timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; tail -2 $OUT/bench.json; tail -3 $OUT/bench.err;;
launches) echo "== ncu launch list (mini workload)"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_mini.csv \
    python bench.py --workload mini_grid64_lya_lyb --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/launches_mini.log 2>&1;;
ncu) echo "== ncu full k_tau (mini workload)"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tau -s 2 -c 1 -o $OUT/prof_tau \
    python bench.py --workload mini_grid64_lya_lyb --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/prof_tau.log 2>&1; tail -2 $OUT/prof_tau.log | cut -c1-300;;
esac
done
ls -la $OUT
