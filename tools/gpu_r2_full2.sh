#!/bin/bash
# round-2 second half: full GPU suite, smoke, default bench (+ eikonal in the step), reference arm, reference-vs-patched timing, launch list
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout=400 -s > gpurun_out/pytest_gpu_r2b.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_r2b.log
grep -E "passed|failed|exit|^FAILED|^ERROR" gpurun_out/pytest_gpu_r2b.log | tail -12 | cut -c1-250
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke_r2b.log 2>&1; tail -2 gpurun_out/smoke_r2b.log | cut -c1-300
timeout 500 python bench.py > gpurun_out/bench_default_b.log 2> gpurun_out/bench_default_b.err; tail -1 gpurun_out/bench_default_b.log | cut -c1-2500; tail -2 gpurun_out/bench_default_b.err | cut -c1-300
timeout 400 python bench.py --with-eikonal --no-cpu-baseline --steps 50 > gpurun_out/bench_eik.log 2> gpurun_out/bench_eik.err; tail -1 gpurun_out/bench_eik.log | cut -c1-900; tail -2 gpurun_out/bench_eik.err | cut -c1-300
timeout 400 python bench.py --precision fp16 --with-eikonal --no-cpu-baseline --steps 50 > gpurun_out/bench_eik_fp16.log 2> gpurun_out/bench_eik_fp16.err; tail -1 gpurun_out/bench_eik_fp16.log | cut -c1-600
timeout 400 python tools/ref_gpu_check.py > gpurun_out/ref_vs_patched.log 2>&1; grep -E "ms ->" gpurun_out/ref_vs_patched.log | cut -c1-220; tail -2 gpurun_out/ref_vs_patched.log | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"field|wgrad|pack|prologue|composite|absmax|scale_kernel|compose|chain|match|loss" -s 80 -c 80 --csv \
  --log-file gpurun_out/r02_launches_step_eik_fp16x3.csv python bench.py --steps 3 --warmup 3 --pass step --with-eikonal --no-cpu-baseline --no-graph > gpurun_out/bench_under_ncu_eik.log 2>&1
grep -c "field_bwd_kernel" gpurun_out/r02_launches_step_eik_fp16x3.csv
