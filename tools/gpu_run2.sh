#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -s --timeout=300 > gpurun_out/pytest_gpu.log 2>&1; rc=$?
echo "pytest exit $rc (cluster=2 build)" >> gpurun_out/pytest_gpu.log
grep -E "parity|passed|failed|Error|error|timed out" gpurun_out/pytest_gpu.log | tail -30
if [ $rc -ne 0 ]; then
  echo "=== rebuilding with B200R_CLUSTER=1 ==="
  B200R_CLUSTER=1 python lab4d_b200/build.py --force > gpurun_out/build_c1.log 2>&1
  timeout 600 python -m pytest tests -m gpu -q -s --timeout=300 > gpurun_out/pytest_gpu_c1.log 2>&1
  echo "pytest exit $? (cluster=1 build)" >> gpurun_out/pytest_gpu_c1.log
  grep -E "parity|passed|failed|Error|error|timed out" gpurun_out/pytest_gpu_c1.log | tail -30
fi
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log
tail -3 gpurun_out/bench.log
