# One full ncu capture of k_tau on the bench's own workload (C2): DRAM traffic for roofline.traffic and the pipe
# utilisations at full size.  The first two k_tau launches are the untimed counter passes (one line each).
mkdir -p gpurun_out/r02f
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_tau -s 2 -c 1 -o gpurun_out/r02f/prof_tau_c2 \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02f/prof_tau_c2.log 2>&1; tail -2 gpurun_out/r02f/prof_tau_c2.log | cut -c1-200
