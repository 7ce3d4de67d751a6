#!/bin/bash
# quaternion tests; ncu --set full of the eikonal reverse chain, forward chain A and the coalesced composite_bwd
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_quat.py -m gpu -q --timeout=200 > gpurun_out/quat_tests.log 2>&1; grep -E "passed|failed|^E  " gpurun_out/quat_tests.log | head -12 | cut -c1-300
sed -i 's/-s 4 -c 1/-s ${SKIP:-4} -c 1/' tools/gpu_prof_one.sh
SKIP=3 bash tools/gpu_prof_one.sh "field_bwd_kernel.*0,.256,.0,.1" eik_reverse --pass step --precision fp16x3 --with-eikonal
SKIP=4 bash tools/gpu_prof_one.sh "field_bwd_kernel.*0,.256,.0,.1" eik_chain_a --pass step --precision fp16x3 --with-eikonal
SKIP=2 bash tools/gpu_prof_one.sh composite_bwd_kernel composite_bwd --pass step --precision fp16x3
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -1 | cut -c1-400
