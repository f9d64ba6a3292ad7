#!/bin/bash
# Round-2 GPU pass B: new attention kernels first, then the whole GPU suite and the bench.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_attn_gpu.py tests/test_vit_ops_gpu.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_attn.log 2>&1; echo "attn pytest exit $?"; tail -30 gpurun_out/pytest_attn.log | cut -c1-400
timeout 1800 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -s --deselect tests/test_attn_gpu.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_gpu.log | tail -30 | cut -c1-400
timeout 500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dump-ops gpurun_out/ops_r50.csv > gpurun_out/bench_r50.log 2>&1; tail -1 gpurun_out/bench_r50.log | cut -c1-300
