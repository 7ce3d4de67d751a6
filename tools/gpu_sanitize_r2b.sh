#!/bin/bash
# compute-sanitizer (memcheck) over the kernels of the second half of round 2
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_eikonal.py tests/test_gpu_warp_points.py tests/test_gpu_match.py tests/test_gpu_losses.py tests/test_gpu_quat.py -m gpu -q -x --timeout=800 \
  -k "fg_bob-4-8-32 or bg-4-24-33 or window or fg_bob-4-16 or fg_compquad or 4-8-16 or fg-6-16 or comp-8-32 or 1000-4-4 or 333-3-4 or quat_transform" > gpurun_out/sanitize_r2b.log 2>&1; echo "sanitizer exit $?" >> gpurun_out/sanitize_r2b.log
grep -E "ERROR SUMMARY|passed|failed|exit|Invalid|out of bounds" gpurun_out/sanitize_r2b.log | head -8 | cut -c1-200
