#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_quat.py -m gpu -q --timeout=200 > gpurun_out/quat_tests.log 2>&1; grep -E "passed|failed|^E  " gpurun_out/quat_tests.log | head -20 | cut -c1-300
timeout 500 python tools/ref_gpu_check.py > gpurun_out/ref_vs_patched.log 2>&1; grep -E "ms ->" gpurun_out/ref_vs_patched.log | cut -c1-220; tail -2 gpurun_out/ref_vs_patched.log | cut -c1-300
timeout 300 python tools/profile_patched.py --dq > gpurun_out/profile_patched_dq.log 2>&1; grep -vE "Warning|warn|^\s*\"\"\"" gpurun_out/profile_patched_dq.log | tail -26 | cut -c1-200
