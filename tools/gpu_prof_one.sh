#!/bin/bash
# one ncu --set full capture: tools/gpu_prof_one.sh <kernel regex> <tag> <bench args...>
mkdir -p gpurun_out
K=$1; TAG=$2; shift 2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s ${SKIP:-4} -c 1 -f -o gpurun_out/r02_$TAG \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph "$@" > gpurun_out/ncu_$TAG.log 2>&1
ncu -i gpurun_out/r02_$TAG.ncu-rep --page raw --csv > gpurun_out/r02_${TAG}_raw.csv 2>/dev/null
# top-sampled SASS instructions (the .ncu-rep itself is 25-35 MB: it stays on the box unless KEEP_REP=1)
ncu -i gpurun_out/r02_$TAG.ncu-rep --page source --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
h=rows[1]; isrc,isamp=h.index('Source'),h.index('# Samples')
d=[(int(r[isamp]),i,r[isrc].strip()) for i,r in enumerate(rows[2:]) if len(r)>isamp and r[isamp].isdigit()]
tot=sum(x[0] for x in d)
print('total samples',tot,'instructions',len(d))
for c,i,sx in sorted(d,reverse=True)[:25]: print('%7d %5.1f%%  #%5d  %s' % (c,100.0*c/tot,i,sx[:100]))
" > gpurun_out/r02_${TAG}_top_sass.txt
[ "$KEEP_REP" = "1" ] || rm -f gpurun_out/r02_$TAG.ncu-rep
python - "$TAG" <<'PY'
import csv,sys
tag=sys.argv[1]
rows=list(csv.reader(open(f"gpurun_out/r02_{tag}_raw.csv")))
h,u,v=rows[0],rows[1],rows[2]
want=["gpu__time_duration.sum","sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active","dram__bytes_read.sum","dram__bytes_write.sum","smsp__issue_active.avg.pct","sm__inst_executed.sum","smsp__inst_executed.avg.per_cycle_active"]
print("==",tag)
for w in want:
    for i,n in enumerate(h):
        if n==w: print(f"   {w} = {v[i]} {u[i]}")
stall=[(float(v[i].replace(",","")),h[i]) for i in range(len(h)) if "smsp__average_warps_issue_stalled" in h[i] and "per_issue_active" in h[i] and v[i].replace(",","").replace(".","").isdigit()]
for val,n in sorted(stall,reverse=True)[:7]: print(f"   stall {n.split('stalled_')[1].split('_per')[0]} = {val:.2f}")
PY
