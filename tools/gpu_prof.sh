#!/bin/bash
# launch list + full ncu capture of the dominant kernel (one GPU; never wrap a multi-rank command)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout=300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:field_fwd_kernel -s 3 -c 1 -f -o gpurun_out/field_fwd \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_field.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:composite_fwd_kernel -s 6 -c 1 -f -o gpurun_out/composite_fwd \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_comp.log 2>&1
ls -la gpurun_out
