#!/bin/bash
mkdir -p gpurun_out
python tests/profile_attn.py 2>&1 | tail -1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 6 -c 3 -o gpurun_out/prof_attn_r2 python tests/profile_attn.py --iters 1 > gpurun_out/ncu_attn.log 2>&1; tail -2 gpurun_out/ncu_attn.log
timeout 300 python -m pytest tests/test_attn_gpu.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -5 | cut -c1-300
