#!/bin/bash
mkdir -p gpurun_out
python tests/profile_attn.py 2>&1 | tail -1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 6 -c 3 -o gpurun_out/prof_attn_r2b python tests/profile_attn.py --iters 1 > gpurun_out/ncu_attn.log 2>&1; tail -1 gpurun_out/ncu_attn.log
timeout 900 python -m pytest tests/test_van_gpu.py -m gpu -q --tb=short -p no:cacheprovider -s -k stagewise 2>&1 | tail -8 | cut -c1-700
