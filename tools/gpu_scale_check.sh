#!/bin/bash
# the driver's scaling launch at N GPUs (N = $1): default bench (C2 step, weak) and the strong-scaling C4 shape
N=${1:-4}
mkdir -p gpurun_out
for spec in "c2" "c4"; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --config $spec --steps 30 --warmup 5 > gpurun_out/scale_n${N}_$spec.log 2> gpurun_out/scale_n${N}_$spec.err
  echo "== N=$N $spec: $(grep '^{' gpurun_out/scale_n${N}_$spec.log | tail -1 | python -c 'import sys,json
try:
  d=json.loads(sys.stdin.read()); print("value %.3e e2e %.3e ms %.3f graph %s n_gpus %s scaling %s grad_bytes %s clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d.get("cuda_graph"), d["n_gpus"], d["scaling"], d.get("grad_buffer_bytes"), d.get("clocks")))
except Exception as e: print("FAILED", e)')"; grep -v Warning gpurun_out/scale_n${N}_$spec.err | tail -3 | cut -c1-300
done
