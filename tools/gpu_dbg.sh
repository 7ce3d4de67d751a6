#!/bin/bash
# stress: many launches of the field kernel under three timing regimes (device-resident, e2e with H2D copies, kernel alone)
mkdir -p gpurun_out
for i in 1 2 3; do
timeout 300 python bench.py --steps 400 --warmup 5 --no-cpu-baseline > gpurun_out/bench_dbg$i.log 2>&1; echo "bench exit $?"
grep -m4 "b200r:" gpurun_out/bench_dbg$i.log
grep -o '"kernel_ms": [0-9.]*' gpurun_out/bench_dbg$i.log
done
