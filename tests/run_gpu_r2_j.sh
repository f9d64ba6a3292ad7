#!/bin/bash
# round-2 GPU pass J: regression tests after the 128-bit vector-access fix, BN microbench, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python tests/profile_bn.py > gpurun_out/profile_bn2.log 2>&1; echo "bn rc=$?"
grep "rows    802816 C   256" gpurun_out/profile_bn2.log
timeout 120 python tests/profile_attn.py > gpurun_out/profile_attn.log 2>&1; tail -3 gpurun_out/profile_attn.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sam --dump-ops gpurun_out/ops_r50.csv > gpurun_out/bench_r50.log 2> gpurun_out/bench_r50.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench_r50.err
cut -c1-300 gpurun_out/bench_r50.log
