#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do
timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/bench_dbg$i.log 2>&1; echo "bench exit $?"
grep -m8 "b200r:" gpurun_out/bench_dbg$i.log
done
