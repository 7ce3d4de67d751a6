#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 40 --csv --log-file gpurun_out/launches_r01.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo done
