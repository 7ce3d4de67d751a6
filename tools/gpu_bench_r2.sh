#!/bin/bash
# step / forward benches in the three precisions + ncu launch list of the training step
mkdir -p gpurun_out
for spec in "step fp16x3" "step fp16" "forward fp16x3" "forward fp16"; do
  set -- $spec
  timeout 300 python bench.py --steps 30 --warmup 5 --pass $1 --precision $2 --no-cpu-baseline > gpurun_out/bench_$1_$2.log 2>&1
  echo "== $1 $2: $(tail -1 gpurun_out/bench_$1_$2.log | python -c 'import sys,json
try:
  d=json.loads(sys.stdin.read()); print("value %.3e e2e %.3e ms %.3f phases %s frac %s launches %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], {k:(round(v,3) if isinstance(v,float) else "") for k,v in d["phases_ms"].items() if k!="note"}, d.get("roofline",{}).get("frac"), d["gpu_launches"]))
except Exception as e: print("FAILED", e)')"
done
tail -5 gpurun_out/bench_step_fp16x3.log | cut -c1-600
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 120 --csv --log-file gpurun_out/launches_step.csv \
  python bench.py --steps 3 --warmup 3 --pass step --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(open("gpurun_out/launches_step.csv")))
hi=[i for i,r in enumerate(rows) if "Kernel Name" in r][0]
h=rows[hi]; ik=h.index("Kernel Name"); iv=h.index("Metric Value"); iu=h.index("Metric Unit")
agg=collections.OrderedDict()
for r in rows[hi+1:]:
    if len(r)<=iv: continue
    name=r[ik].split("(")[0][:70]; v=float(r[iv].replace(",","")); u=r[iu]
    v*= {"ns":1e-3,"us":1,"ms":1e3,"usecond":1,"nsecond":1e-3,"msecond":1e3}.get(u,1)
    agg.setdefault(name,[]).append(v)
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1]))[:22]:
    print(f"{sum(v):10.1f} us total  {len(v):4d} x {sum(v)/len(v):9.1f} us  {k}")
PY
