# BASELINE configs[3] at full size: 1024^3 top-hat cells sharded over 8 GPUs, NCCL all-reduce of the FP64 tau array
mkdir -p gpurun_out/$1
free -g | head -2
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 3 --warmup 3 --workload c4_tophat_pshard_1024 > gpurun_out/$1/bench_c4_1024_8gpu.json 2> gpurun_out/$1/bench_c4_1024_8gpu.err
tail -1 gpurun_out/$1/bench_c4_1024_8gpu.json | cut -c1-1500; tail -n 3 gpurun_out/$1/bench_c4_1024_8gpu.err
