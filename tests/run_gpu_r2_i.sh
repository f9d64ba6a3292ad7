#!/bin/bash
# round-2 GPU pass I: SAM tests, BN slab-kernel microbench + ncu, bench with SAM-H sub-record
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt
timeout 300 python -m pytest tests/test_sam_gpu.py -m gpu -x -q > gpurun_out/pytest_sam.log 2>&1; echo "sam rc=$?"
tail -3 gpurun_out/pytest_sam.log
timeout 300 python tests/profile_bn.py > gpurun_out/profile_bn.log 2>&1; echo "bn rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:"bn_|colreduce|add_bf16" -o gpurun_out/prof_bn_r2 -f python tests/profile_bn.py --once > gpurun_out/ncu_bn_r2.log 2>&1; echo "ncu rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sam --dump-ops gpurun_out/ops_r50.csv > gpurun_out/bench_r50.log 2> gpurun_out/bench_r50.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench_r50.err
cut -c1-400 gpurun_out/bench_r50.log
