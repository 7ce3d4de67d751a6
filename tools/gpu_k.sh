#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -s --timeout=120 -k "$1" > gpurun_out/pytest_k.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|Error|assert|b200r" gpurun_out/pytest_k.log | head -12 | cut -c1-300
