#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_backward.py tests/test_gpu_contract.py tests/test_gpu_reference.py -m gpu -q -x --timeout=200 > gpurun_out/q2.log 2>&1; tail -3 gpurun_out/q2.log | cut -c1-300
for spec in "step fp16x3" "step fp16"; do
  set -- $spec
  timeout 300 python bench.py --steps 50 --warmup 5 --pass $1 --precision $2 --no-cpu-baseline > gpurun_out/bench_$1_$2.log 2> gpurun_out/bench_$1_$2.err
  echo "== $1 $2: $(tail -1 gpurun_out/bench_$1_$2.log | python -c 'import sys,json
try:
  d=json.loads(sys.stdin.read()); print("value %.3e e2e %.3e ms %.3f graph %s phases %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d.get("cuda_graph"), {k:(round(v,3) if isinstance(v,float) else "") for k,v in d["phases_ms"].items() if k!="note"}))
except Exception as e: print("FAILED", e)')"; grep -v Warn gpurun_out/bench_$1_$2.err | tail -2 | cut -c1-300
done
