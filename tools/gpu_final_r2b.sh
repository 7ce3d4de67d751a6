#!/bin/bash
# round-end check of the second half: full GPU suite, smoke, default bench, reference arm, reference-vs-patched timings (train + eval)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout=400 > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_final.log
grep -E "passed|failed|exit|^FAILED|^ERROR" gpurun_out/pytest_gpu_final.log | tail -8 | cut -c1-250
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke_final.log 2>&1; tail -1 gpurun_out/smoke_final.log | cut -c1-400
timeout 500 python bench.py > gpurun_out/bench_final.log 2> gpurun_out/bench_final.err; tail -1 gpurun_out/bench_final.log | cut -c1-3000; tail -2 gpurun_out/bench_final.err | cut -c1-300
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_final.log 2>&1; tail -1 gpurun_out/bench_ref_final.log | cut -c1-700
timeout 500 python tools/ref_gpu_check.py > gpurun_out/ref_vs_patched.log 2>&1; grep -E "ms ->" gpurun_out/ref_vs_patched.log | cut -c1-230; tail -1 gpurun_out/ref_vs_patched.log | cut -c1-300
