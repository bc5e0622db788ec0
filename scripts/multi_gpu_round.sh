#!/bin/bash
# bench.py under torchrun on N GPUs of one box: the default sightline-sharded workload and the particle-sharded
# top-hat workload with the NCCL all-reduce.  Usage: gpurun --gpus N -- 'bash scripts/multi_gpu_round.sh N tag'
N=${1:-2}; TAG=${2:-r01}
OUT=gpurun_out/$TAG; mkdir -p $OUT
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 3 --warmup 3 "${@:2}"; }
run 29511 > $OUT/bench_c2_${N}gpu.json 2> $OUT/bench_c2_${N}gpu.err; tail -1 $OUT/bench_c2_${N}gpu.json | cut -c1-200
run 29512 --workload c4_tophat_pshard > $OUT/bench_c4_pshard_${N}gpu.json 2> $OUT/bench_c4_pshard_${N}gpu.err; tail -1 $OUT/bench_c4_pshard_${N}gpu.json | cut -c1-200
for f in $OUT/*.err; do tail -n 2 "$f"; done | tail -n 6
