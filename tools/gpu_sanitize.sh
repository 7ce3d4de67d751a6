#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -s --timeout=300 -k "single_frame or variants" > gpurun_out/pytest_new.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|Error|assert" gpurun_out/pytest_new.log | head -8 | cut -c1-300
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout=800 -k "matches_reference and (fg_compquad or bg_rigid)" > gpurun_out/sanitize.log 2>&1; echo "sanitizer exit $?"
grep -E "ERROR SUMMARY|Invalid|passed|failed|out of bounds|misaligned" gpurun_out/sanitize.log | head -12 | cut -c1-300
