#!/bin/bash
mkdir -p gpurun_out
echo "== cluster=2 (prebuilt)"; timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms'])"
B200R_CLUSTER=1 python lab4d_b200/build.py --force > gpurun_out/build_c1.log 2>&1
echo "== cluster=1"; timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms'])"
timeout 300 python -m pytest tests -m gpu -q --timeout=300 2>&1 | tail -2
