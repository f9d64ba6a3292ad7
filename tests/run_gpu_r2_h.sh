#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sam_gpu.py -m gpu -q --tb=short -p no:cacheprovider -s > gpurun_out/pytest_sam.log 2>&1
grep -n "rel L2\|passed\|failed\|Error\|worst" gpurun_out/pytest_sam.log | cut -c1-600 | head -40
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dump-ops gpurun_out/ops_r50.csv > gpurun_out/bench_r50.log 2>&1; tail -1 gpurun_out/bench_r50.log | cut -c1-400
