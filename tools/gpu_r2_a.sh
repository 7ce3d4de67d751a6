#!/bin/bash
# round-2 call A: MN-major descriptor probe, reference-on-CUDA check, sanity of the committed suite
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 60 tools/mnmajor_probe > gpurun_out/mnmajor.log 2>&1; echo "probe exit $?" >> gpurun_out/mnmajor.log; cat gpurun_out/mnmajor.log
timeout 400 python tools/ref_gpu_check.py 128 16 128 > gpurun_out/ref_gpu.log 2>&1; echo "ref exit $?" >> gpurun_out/ref_gpu.log; tail -4 gpurun_out/ref_gpu.log | cut -c1-300
timeout 500 python -m pytest tests -m gpu -q --timeout=120 -x > gpurun_out/pytest_gpu_a.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_a.log
tail -3 gpurun_out/pytest_gpu_a.log | cut -c1-300
