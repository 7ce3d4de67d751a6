#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv
for i in 1 2 3; do timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['clocks'])"; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"prologue_kernel|field_fwd_kernel|composite_fwd" -s 8 -c 8 --csv --log-file gpurun_out/q.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/q.csv')) if len(r)>5]
h=rows[0]
for r in rows[1:]: print(r[h.index("Kernel Name")][:40], r[h.index("Metric Value")])
PY
