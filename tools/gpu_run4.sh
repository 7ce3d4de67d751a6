#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s --timeout=300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "compquad|two-field|passed|failed|Error|timed out|exit|assert" gpurun_out/pytest_gpu.log | cut -c1-600 | tail -16
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log
tail -2 gpurun_out/bench.log | cut -c1-1200
