#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_backward.py -m gpu -q -k compose --timeout=100 2>&1 | tail -3 | cut -c1-300
for spec in "c4 step fp16x3" "c4 forward fp16x3" "c3 step bf16" "c2 step fp16x3" "c2 forward fp16x3"; do
  set -- $spec
  timeout 400 python bench.py --config $1 --steps 20 --warmup 4 --pass $2 --precision $3 --no-cpu-baseline > gpurun_out/bench_$1_$2_$3.log 2> gpurun_out/bench_$1_$2_$3.err
  echo "== $1 $2 $3: $(tail -1 gpurun_out/bench_$1_$2_$3.log | python -c 'import sys,json
try:
  d=json.loads(sys.stdin.read()); print("value %.3e e2e %.3e ms %.3f graph %s phases %s frac %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d.get("cuda_graph"), {k:(round(v,3) if isinstance(v,float) else "") for k,v in d["phases_ms"].items() if k!="note"}, d.get("roofline",{}).get("frac")))
except Exception as e: print("FAILED", e)')"; tail -2 gpurun_out/bench_$1_$2_$3.err | cut -c1-300
done
