# DRAM traffic of one k_tau launch on the bench workload (single-pass metrics only; cheap)
mkdir -p gpurun_out/$1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:k_tau -s 2 -c 1 --csv --log-file gpurun_out/$1/traffic_c2.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/$1/traffic_c2.log 2>&1
grep -o '"dram__[a-z_.]*","[A-Za-z]*","[0-9.,]*"\|"lts__[a-z_.]*","[%A-Za-z]*","[0-9.,]*"\|"gpu__time[a-z_.]*","[A-Za-z]*","[0-9.,]*"' gpurun_out/$1/traffic_c2.csv
