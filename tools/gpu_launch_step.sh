#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"field|wgrad|pack|prologue|composite|absmax|scale_kernel|compose" -s 60 -c 60 --csv --log-file gpurun_out/launches_step_k.csv \
  python bench.py --steps 3 --warmup 3 --pass step ${1:+--precision $1} --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(open("gpurun_out/launches_step_k.csv")))
hi=[i for i,r in enumerate(rows) if "Kernel Name" in r][0]
h=rows[hi]; ik=h.index("Kernel Name"); iv=h.index("Metric Value"); iu=h.index("Metric Unit")
agg=collections.OrderedDict()
for r in rows[hi+1:]:
    if len(r)<=iv: continue
    name=r[ik].split("(")[0][:90]; v=float(r[iv].replace(",","")); u=r[iu]
    v*= {"ns":1e-3,"us":1,"ms":1e3}.get(u,1)
    agg.setdefault(name,[]).append(v)
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1]))[:16]:
    print(f"{sum(v):10.1f} us total  {len(v):4d} x {sum(v)/len(v):9.1f} us  {k}")
PY
