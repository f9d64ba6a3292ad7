#!/bin/bash
# round-2 GPU pass P: whole GPU suite after the kernel changes (SAMLoss, maxpool, stem im2col, delta, folds), epilogue study, bench with SAM + DETR
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/pytest_gpu.log | cut -c1-400
timeout 300 python tests/profile_epilogue.py > gpurun_out/profile_epilogue.log 2>&1; cat gpurun_out/profile_epilogue.log
timeout 300 python tests/profile_conv.py > gpurun_out/profile_conv.log 2>&1; cat gpurun_out/profile_conv.log
timeout 120 python tests/profile_attn.py > gpurun_out/profile_attn.log 2>&1; tail -1 gpurun_out/profile_attn.log
timeout 1200 python bench.py --steps 10 --warmup 3 --sam --detr --dump-ops gpurun_out/ops_r50.csv > gpurun_out/bench_r50.log 2> gpurun_out/bench_r50.err; echo "bench rc=$?"
tail -c 800 gpurun_out/bench_r50.err
cut -c1-300 gpurun_out/bench_r50.log
