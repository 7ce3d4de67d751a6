#!/bin/bash
# A/B of a prebuilt library variant (B200R_LIB=<path>): tools/gpu_ab2.sh lab4d_b200/libb200render_nowd.so
mkdir -p gpurun_out
for rep in 1 2; do
  timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_A$rep.log 2>&1; echo "A$rep $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_A$rep.log | head -1) $(grep -o '"forward_call": [0-9.]*, "backward_call": [0-9.]*' gpurun_out/bench_A$rep.log)"
  B200R_LIB=$PWD/$1 timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_B$rep.log 2>&1; echo "B$rep $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_B$rep.log | head -1) $(grep -o '"forward_call": [0-9.]*, "backward_call": [0-9.]*' gpurun_out/bench_B$rep.log)"
done
