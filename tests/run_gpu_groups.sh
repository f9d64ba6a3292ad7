#!/bin/bash
# Runs the kernel parity tests group by group, each under its own timeout, so that one hung
# kernel does not hide the results of the others.  Usage (on the GPU box): bash tests/run_gpu_groups.sh
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for grp in "bn_forward_backward or maxpool or colsum" "linear_fwd" "linear_dgrad" "linear_wgrad" "stem" "conv_fprop" "conv_dgrad" "conv_wgrad"; do
  tag=$(echo "$grp" | tr ' ' '_')
  echo "=== group: $grp"
  timeout 300 python -m pytest tests/test_ops_gpu.py -q -m gpu -k "$grp" --tb=line -p no:cacheprovider 2>&1 | tail -25 | tee gpurun_out/ops_$tag.log
  echo "exit: ${PIPESTATUS[0]}"
done
