#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_backward.py tests/test_gpu_contract.py -m gpu -q --timeout=280 -k "compose or two_field or c4 or C4 or comp" 2>&1 | tail -3 | cut -c1-300
timeout 300 python bench.py --config c4 --steps 30 --no-cpu-baseline > gpurun_out/bench_c4.log 2> gpurun_out/bench_c4.err; tail -1 gpurun_out/bench_c4.log | cut -c1-300; tail -2 gpurun_out/bench_c4.err | cut -c1-200
