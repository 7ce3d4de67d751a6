#!/bin/bash
# normals kernel: parity, the patched-reference eval test; eikonal window test; sanitizer over the kernels of this half of the round
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_normals.py -m gpu -q -s --timeout=300 > gpurun_out/normals_tests.log 2>&1; grep -E "\[normals\]|passed|failed|^E  " gpurun_out/normals_tests.log | head -14 | cut -c1-400
timeout 400 python -m pytest tests/test_gpu_reference.py tests/test_gpu_eikonal.py -m gpu -q -s --timeout=300 > gpurun_out/ref_tests3.log 2>&1; grep -E "fg/bob eval|window|passed|failed|^E  " gpurun_out/ref_tests3.log | head -10 | cut -c1-500
bash tools/gpu_sanitize_r2b.sh
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_normals.py -m gpu -q -x --timeout=280 -k "bg or fg_compquad" > gpurun_out/sanitize_normals.log 2>&1; echo "sanitizer exit $?" >> gpurun_out/sanitize_normals.log
grep -E "ERROR SUMMARY|passed|failed|exit" gpurun_out/sanitize_normals.log | head -4 | cut -c1-200
