#!/bin/bash
# sweep of the weight-gradient work-split model constants: backward call time (graph replay) per setting
mkdir -p gpurun_out
for cfg in "2.5 0.36 50" "3.0 0.36 50" "3.5 0.36 50" "2.5 0.30 50" "2.5 0.36 35" "3.0 0.36 90" "4.0 0.40 60" "2.0 0.36 50"; do
  set -- $cfg
  B200R_WG_LAT=$1 B200R_WG_BW=$2 B200R_WG_FLUSH=$3 timeout 200 python bench.py --steps 30 --warmup 4 --pass step --precision fp16 --no-cpu-baseline > gpurun_out/wg_sweep.log 2>/dev/null
  echo "lat $1 bw $2 flush $3: $(tail -1 gpurun_out/wg_sweep.log | python -c 'import sys,json
d=json.loads(sys.stdin.read()); print("bwd %.3f ms step %.3f" % (d["phases_ms"]["backward_call"], d["ms_per_step"]))')"
done
