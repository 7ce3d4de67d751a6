#!/bin/bash
# usage: tools/gpu_k.sh <test file or empty> "<-k expr>" [log name]
mkdir -p gpurun_out
LOG=gpurun_out/${3:-pytest_k}.log
timeout 900 python -m pytest ${1:-tests} -m gpu -q -s --timeout=300 ${2:+-k "$2"} > $LOG 2>&1; echo "pytest exit $?" >> $LOG
grep -E "^\[|passed|failed|Error|error|timed out|exit|assert|b200r:" $LOG | cut -c1-330 | tail -${4:-60}
