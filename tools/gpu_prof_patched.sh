#!/bin/bash
mkdir -p gpurun_out
timeout 500 python tools/profile_patched.py > gpurun_out/profile_patched.log 2>&1; grep -vE "Warning|warn|^\s*\"\"\"" gpurun_out/profile_patched.log | tail -34 | cut -c1-200
timeout 300 python -m pytest tests/test_gpu_eikonal.py -m gpu -q -s --timeout=300 2>&1 | grep -E "eikonal\] .* g rel|passed|failed" | cut -c1-400
