#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -s --timeout=120 -k "variants or two_field" > gpurun_out/pytest_var.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|parity\] oracle (w|skel18_)" gpurun_out/pytest_var.log | cut -c1-400
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/bench_${N}gpu.log 2>&1; echo "bench$N exit $?"
tail -1 gpurun_out/bench_${N}gpu.log | cut -c1-300
