#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -s --timeout=120 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|Error|timed out|exit|assert|b200r:" gpurun_out/pytest_gpu.log | cut -c1-400 | tail -8
for i in 1 2; do
timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu-baseline > gpurun_out/bench_s$i.log 2>&1; echo "bench exit $?"
grep -m4 "b200r:" gpurun_out/bench_s$i.log
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_s$i.log") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d["roofline"]["frac"])
PY
done
