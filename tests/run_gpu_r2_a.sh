#!/bin/bash
# Round-2 GPU pass A: full GPU test suite, smoke, bench (both models, op tables), unmodified-reference torch baselines.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; lscpu | grep -E "Model name|Socket|Core|Thread" >> gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -25 gpurun_out/pytest_gpu.log | cut -c1-600
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 500 python bench.py --steps 10 --warmup 3 --dump-ops gpurun_out/ops_r50.csv > gpurun_out/bench_r50.log 2>&1; tail -1 gpurun_out/bench_r50.log | cut -c1-1500
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-600
rm -f gpurun_out/torch_gpu_baseline.jsonl
for v in bf16 as_shipped tuned; do
  timeout 300 python baseline/torch_gpu_baseline.py --model resnet50 --variant $v --steps 15 --warmup 5 --out gpurun_out/torch_gpu_baseline.jsonl 2>&1 | tail -1 | cut -c1-400
done
for v in bf16 as_shipped; do
  timeout 300 python baseline/torch_gpu_baseline.py --model vit_base_patch16 --variant $v --steps 10 --warmup 4 --out gpurun_out/torch_gpu_baseline.jsonl 2>&1 | tail -1 | cut -c1-400
done
