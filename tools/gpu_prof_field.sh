#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:field_fwd_kernel -s 3 -c 1 -f -o gpurun_out/field_fwd_r01b \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_field.log 2>&1
tail -3 gpurun_out/ncu_field.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep
