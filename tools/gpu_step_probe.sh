#!/bin/bash
# correctness after a kernel change + launch list of the product kernels + step benches
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_backward.py -m gpu -q -x --timeout=200 > gpurun_out/bwd_q.log 2>&1; tail -2 gpurun_out/bwd_q.log | cut -c1-200
bash tools/gpu_launch_step.sh ${1:-fp16}
for spec in "step fp16" "step fp16x3"; do
  set -- $spec
  timeout 300 python bench.py --steps 30 --warmup 5 --pass $1 --precision $2 --no-cpu-baseline > gpurun_out/bench_$1_$2.log 2> gpurun_out/bench_$1_$2.err
  echo "== $1 $2: $(tail -1 gpurun_out/bench_$1_$2.log | python -c 'import sys,json
try:
  d=json.loads(sys.stdin.read()); print("value %.3e e2e %.3e ms %.3f graph %s phases %s frac %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d.get("cuda_graph"), {k:(round(v,3) if isinstance(v,float) else "") for k,v in d["phases_ms"].items() if k!="note"}, d.get("roofline",{}).get("frac")))
except Exception as e: print("FAILED", e)')"; tail -3 gpurun_out/bench_$1_$2.err | cut -c1-300
done
