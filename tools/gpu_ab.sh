#!/bin/bash
# A/B of a build flag: tools/gpu_ab.sh "<ENV=VAL ...>" (rebuilds on the box, benches, restores nothing: the box is scratch)
mkdir -p gpurun_out
timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_A.log 2>&1; grep -o '"kernel_ms": [0-9.]*' gpurun_out/bench_A.log
env $1 python -m lab4d_b200.build --force > gpurun_out/build_B.log 2>&1; tail -1 gpurun_out/build_B.log
timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_B.log 2>&1; grep -o '"kernel_ms": [0-9.]*' gpurun_out/bench_B.log; grep -m3 "b200r:\|Error" gpurun_out/bench_B.log
