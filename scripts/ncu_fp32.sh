mkdir -p gpurun_out/r02e
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tau -s 2 -c 1 -o gpurun_out/r02e/prof_tau_fp32 \
    python bench.py --workload mini_grid64_lya_lyb --precision fp32 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02e/prof_tau_fp32.log 2>&1; tail -2 gpurun_out/r02e/prof_tau_fp32.log | cut -c1-200
