#!/bin/bash
# round-2 captures: one ncu --set full of each big kernel of the training step (eager launches), raw pages exported as CSV
mkdir -p gpurun_out
PREC=${1:-fp16}
for spec in "field_bwd_kernel bwd" "wgrad_kernel wgrad" "field_fwd_kernel fwdtrain"; do
  set -- $spec
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$1 -s 4 -c 1 -f -o gpurun_out/r02_$2_$PREC \
    python bench.py --steps 2 --warmup 3 --pass step --precision $PREC --no-cpu-baseline --no-graph > gpurun_out/ncu_$2.log 2>&1
  ncu -i gpurun_out/r02_$2_$PREC.ncu-rep --page raw --csv > gpurun_out/r02_$2_${PREC}_raw.csv 2>/dev/null
  python - "$2" "$PREC" <<'PY'
import csv,sys
tag,prec=sys.argv[1],sys.argv[2]
rows=list(csv.reader(open(f"gpurun_out/r02_{tag}_{prec}_raw.csv")))
h,u,v=rows[0],rows[1],rows[2]
want=["gpu__time_duration.sum","sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active","dram__bytes_read.sum","dram__bytes_write.sum","gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed","smsp__issue_active.avg.pct","sm__inst_executed.sum","launch__registers_per_thread","smsp__average_warp_latency_issue_stalled_long_scoreboard.pct","l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum","lts__t_sectors_op_write.sum","smsp__inst_executed.avg.per_cycle_active"]
print("==",tag,prec)
for w in want:
    for i,n in enumerate(h):
        if n==w: print(f"   {w} = {v[i]} {u[i]}")
stall=[(float(v[i].replace(",","")),h[i]) for i in range(len(h)) if "smsp__average_warps_issue_stalled" in h[i] and "per_issue_active" in h[i] and v[i].replace(",","").replace(".","").isdigit()]
for val,n in sorted(stall,reverse=True)[:6]: print(f"   stall {n.split('stalled_')[1].split('_per')[0]} = {val:.2f}")
PY
done
ls -la gpurun_out/*.ncu-rep | awk '{print $5, $9}'
