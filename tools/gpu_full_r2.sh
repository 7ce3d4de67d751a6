#!/bin/bash
# full GPU suite + smoke + default bench + reference arm + sanitizer on two backward cases
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=300 > gpurun_out/pytest_gpu_r2.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_r2.log
grep -E "passed|failed|exit|^FAILED|^ERROR" gpurun_out/pytest_gpu_r2.log | tail -8 | cut -c1-250
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke_r2.log 2>&1; tail -2 gpurun_out/smoke_r2.log | cut -c1-300
timeout 400 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; tail -1 gpurun_out/bench_default.log | cut -c1-1500; tail -2 gpurun_out/bench_default.err | cut -c1-200
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r2.log 2>&1; tail -1 gpurun_out/bench_ref_r2.log | cut -c1-600
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_backward.py -m gpu -q -x --timeout=500 -k "fg_bob-4-16-48 or fg_compquad or bg-4" > gpurun_out/sanitize_r2.log 2>&1; echo "sanitizer exit $?" >> gpurun_out/sanitize_r2.log
grep -E "ERROR SUMMARY|passed|failed|exit|Invalid|out of bounds" gpurun_out/sanitize_r2.log | head -8 | cut -c1-200
