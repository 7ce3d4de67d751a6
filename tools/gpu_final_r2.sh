#!/bin/bash
# round-2 final captures: launch lists of one training step (both precisions), ncu --set full of the five kernels
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_backward.py -m gpu -q -x --timeout=200 2>&1 | tail -1 | cut -c1-200
for PREC in fp16x3 fp16; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"field|wgrad|pack|prologue|composite|absmax|scale_kernel|compose|chain" -s 66 -c 66 --csv \
    --log-file gpurun_out/r02_launches_step_$PREC.csv python bench.py --steps 3 --warmup 3 --pass step --precision $PREC --no-cpu-baseline --no-graph > gpurun_out/bench_under_ncu_$PREC.log 2>&1
done
bash tools/gpu_prof_one.sh field_bwd_kernel field_bwd --pass step --precision fp16
bash tools/gpu_prof_one.sh wgrad_kernel wgrad --pass step --precision fp16
bash tools/gpu_prof_one.sh field_fwd_kernel field_fwd_train_fp16 --pass step --precision fp16
bash tools/gpu_prof_one.sh field_fwd_kernel field_fwd_train_fp16x3 --pass step --precision fp16x3
bash tools/gpu_prof_one.sh field_fwd_kernel field_fwd_infer_fp16x3 --pass forward --precision fp16x3
bash tools/gpu_prof_one.sh composite_bwd_kernel composite_bwd --pass step --precision fp16
ls gpurun_out/*.ncu-rep | wc -l
