#!/bin/bash
# N-GPU runs of the step benches (N = $1)
N=${1:-2}
mkdir -p gpurun_out
for spec in "c2 step fp16x3" "c4 step fp16x3" "c2 step fp16"; do
  set -- $spec
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --config $1 --steps 30 --warmup 5 --pass $2 --precision $3 > gpurun_out/bench_n${N}_$1_$2_$3.log 2> gpurun_out/bench_n${N}_$1_$2_$3.err
  echo "== N=$N $1 $2 $3: $(grep '^{' gpurun_out/bench_n${N}_$1_$2_$3.log | tail -1 | python -c 'import sys,json
try:
  d=json.loads(sys.stdin.read()); print("value %.3e e2e %.3e ms %.3f graph %s n_gpus %s grad_bytes %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d.get("cuda_graph"), d["n_gpus"], d.get("grad_buffer_bytes")))
except Exception as e: print("FAILED", e)')"; grep -v Warning gpurun_out/bench_n${N}_$1_$2_$3.err | tail -3 | cut -c1-300
done
