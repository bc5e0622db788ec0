mkdir -p gpurun_out/r01u
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r01u/launches_c2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r01u/launches_c2.log 2>&1
tail -2 gpurun_out/r01u/launches_c2.log | cut -c1-300
