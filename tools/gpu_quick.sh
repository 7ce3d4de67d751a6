#!/bin/bash
mkdir -p gpurun_out
timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_q.log 2>&1; grep -o '"kernel_ms": [0-9.]*' gpurun_out/bench_q.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 40 --csv --log-file gpurun_out/launches_q.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
grep -E "prologue|field_fwd|composite|pack" gpurun_out/launches_q.csv | awk -F'","' '{print $5, $NF}' | cut -c1-120 | head -12
