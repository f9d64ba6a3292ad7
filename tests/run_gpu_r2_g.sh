#!/bin/bash
mkdir -p gpurun_out
python tests/profile_attn.py 2>&1 | tail -1
timeout 300 python -m pytest tests/test_attn_gpu.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -3 | cut -c1-300
CUDA_LAUNCH_BLOCKING=1 timeout 900 python -m pytest tests/test_sam_gpu.py -m gpu -q -x --tb=long -p no:cacheprovider -s -k "step_matches_oracle and e128" > gpurun_out/pytest_sam.log 2>&1
grep -n "Error\|error\|saicv_" gpurun_out/pytest_sam.log | head -30
tail -30 gpurun_out/pytest_sam.log | cut -c1-300
