#!/bin/bash
# round-end measurement: tests, bench (own arm + reference arm), ncu launch list, full captures of the three kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -s --timeout=120 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|exit" gpurun_out/pytest_gpu.log | tail -3
timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_final.log 2>&1; tail -1 gpurun_out/bench_final.log | cut -c1-250
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-250
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 40 --csv --log-file gpurun_out/launches_r01.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:field_fwd_kernel -s 3 -c 1 -f -o gpurun_out/field_fwd_r01 \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_field.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:composite_fwd_kernel -s 3 -c 1 -f -o gpurun_out/composite_r01 \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_comp.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:prologue_kernel -s 3 -c 1 -f -o gpurun_out/prologue_r01 \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_prol.log 2>&1
ls gpurun_out
