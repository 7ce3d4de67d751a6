#!/bin/bash
# A/B of the no-watchdog build; C5 bench line (comp_skel-human_dense, 50 instance codes)
bash tools/gpu_ab2.sh lab4d_b200/libb200render_nowd.so
timeout 400 python bench.py --config c5 --steps 30 --no-cpu-baseline > gpurun_out/bench_c5.log 2> gpurun_out/bench_c5.err; tail -1 gpurun_out/bench_c5.log | cut -c1-900; tail -2 gpurun_out/bench_c5.err | cut -c1-300
