#!/bin/bash
# eikonal kernels: parity tests, the patched-reference test (eikonal in the loss), a regression slice of the backward suite
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_eikonal.py -m gpu -q -s --timeout=300 > gpurun_out/eik_tests.log 2>&1; echo "eik exit $?" >> gpurun_out/eik_tests.log
grep -E "^\[eikonal\]|passed|failed|exit|Error|error|^E " gpurun_out/eik_tests.log | head -40 | cut -c1-600
timeout 600 python -m pytest tests/test_gpu_reference.py -m gpu -q -s --timeout=400 > gpurun_out/eik_ref.log 2>&1; echo "ref exit $?" >> gpurun_out/eik_ref.log
grep -E "^\[reference\]|passed|failed|exit|^E " gpurun_out/eik_ref.log | head -20 | cut -c1-900
timeout 300 python -m pytest tests/test_gpu_match.py -m gpu -q -s --timeout=200 > gpurun_out/match_tests.log 2>&1; grep -E "^\[match\]|passed|failed|^E " gpurun_out/match_tests.log | head -12 | cut -c1-300
timeout 300 python -m pytest tests/test_gpu_losses.py -m gpu -q -s --timeout=200 > gpurun_out/loss_tests.log 2>&1; grep -E "^\[losses\]|passed|failed|^E " gpurun_out/loss_tests.log | head -12 | cut -c1-700
timeout 400 python -m pytest tests/test_gpu_backward.py -m gpu -q --timeout=300 -k "fg_bob-4-16-48 or bg-4 or chain" > gpurun_out/eik_bwd.log 2>&1; tail -2 gpurun_out/eik_bwd.log | cut -c1-300
