#!/bin/bash
# round-2 GPU pass R: what the driver runs at round end - GPU suite, smoke, reference arm, default bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 2 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/bench_ref.log
timeout 900 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_default.err; cut -c1-200 gpurun_out/bench_default.log
