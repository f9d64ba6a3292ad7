#!/bin/bash
# round-2 GPU pass K: DETR tests, SAM with the GEMM table gradients, attention regression, bench with SAM
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_detr_gpu.py -m gpu -q -s > gpurun_out/pytest_detr.log 2>&1; echo "detr rc=$?"
grep -n "passed\|failed\|Error\|error\|assert\|detr " gpurun_out/pytest_detr.log | head -40
timeout 600 python -m pytest tests/test_sam_gpu.py tests/test_attn_gpu.py tests/test_vit_gpu.py -m gpu -q > gpurun_out/pytest_sam.log 2>&1; echo "sam/attn rc=$?"
tail -3 gpurun_out/pytest_sam.log
timeout 120 python tests/profile_attn.py > gpurun_out/profile_attn.log 2>&1; tail -2 gpurun_out/profile_attn.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sam --dump-ops gpurun_out/ops_r50.csv > gpurun_out/bench_r50.log 2> gpurun_out/bench_r50.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench_r50.err
cut -c1-300 gpurun_out/bench_r50.log
