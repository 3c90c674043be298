#!/bin/bash
# first bring-up on the GPU box: each test file in its own process (a device trap poisons the
# CUDA context), bounded by timeout, all output kept under gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for f in ${WB_TESTS:-test_gpu_kernels test_gpu_mel test_gpu_encoder test_gpu_decoder}; do
  timeout 600 python -m pytest tests/$f.py -m gpu -q -x --timeout 300 > gpurun_out/$f.log 2>&1
  echo "$f exit $?" | tee -a gpurun_out/summary.txt
  tail -40 gpurun_out/$f.log
done
