#!/bin/bash
# differentiable point warp: parity tests, the patched-reference test, timings; ncu --set full of the eikonal chains (base-name regex + skip)
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_warp_points.py -m gpu -q -s --timeout=300 > gpurun_out/warp_tests.log 2>&1; grep -E "^\[warp-points\]|passed|failed|^E  " gpurun_out/warp_tests.log | head -20 | cut -c1-1200
timeout 400 python -m pytest tests/test_gpu_reference.py -m gpu -q -s --timeout=300 > gpurun_out/ref_tests2.log 2>&1; grep -E "passed|failed|^E  " gpurun_out/ref_tests2.log | head -8 | cut -c1-400
timeout 500 python tools/ref_gpu_check.py > gpurun_out/ref_vs_patched.log 2>&1; grep -E "ms ->" gpurun_out/ref_vs_patched.log | grep -E "patched|incl" | cut -c1-220; tail -1 gpurun_out/ref_vs_patched.log | cut -c1-300
timeout 300 python tools/profile_patched.py --dq > gpurun_out/profile_patched_dq.log 2>&1; grep -vE "Warning|warn|^\s*\"\"\"" gpurun_out/profile_patched_dq.log | tail -26 | cut -c1-160
SKIP=4 bash tools/gpu_prof_one.sh field_bwd_kernel eik_reverse --pass step --precision fp16x3 --with-eikonal
SKIP=5 bash tools/gpu_prof_one.sh field_bwd_kernel eik_chain_a --pass step --precision fp16x3 --with-eikonal
