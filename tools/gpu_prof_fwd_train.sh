#!/bin/bash
mkdir -p gpurun_out
SKIP=2 bash tools/gpu_prof_one.sh field_fwd_kernel fwd_train_x3_final --pass step --precision fp16x3
SKIP=2 bash tools/gpu_prof_one.sh "field_bwd_kernel" dgrad_final --pass step --precision fp16x3
cat gpurun_out/r02_fwd_train_x3_final_top_sass.txt | head -30
cat gpurun_out/r02_dgrad_final_top_sass.txt | head -30
