#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -s --timeout=300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "composite-bwd|passed|failed|Error|timed out|exit" gpurun_out/pytest_gpu.log | cut -c1-400 | tail -12
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log
tail -2 gpurun_out/bench.log | cut -c1-1500
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-600
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/launches_v3.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:field_fwd_kernel -s 3 -c 1 -f -o gpurun_out/field_fwd_v3 \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_field.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:composite_fwd_kernel -s 6 -c 2 -f -o gpurun_out/composite_v3 \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_comp.log 2>&1
ls -la gpurun_out | head -30
