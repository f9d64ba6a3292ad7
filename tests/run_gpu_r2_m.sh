#!/bin/bash
# round-2 GPU pass M (8 GPUs): bench at N=8 (graph-captured default, then eager), torch DDP baseline at N=8
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_n8.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8.log 2> gpurun_out/bench_n8.err; echo "bench n8 rc=$?"
tail -c 400 gpurun_out/bench_n8.err; cut -c1-260 gpurun_out/bench_n8.log
true
for v in bf16 as_shipped; do
  timeout 400 $TR --master-port 29523 baseline/torch_gpu_baseline.py --model resnet50 --variant $v --steps 15 --warmup 5 --out gpurun_out/torch_gpu_baseline_n8.jsonl > gpurun_out/torch_n8_$v.log 2>&1; echo "torch $v rc=$?"
  tail -1 gpurun_out/torch_n8_$v.log | cut -c1-300
done
timeout 400 $TR --master-port 29524 baseline/torch_gpu_baseline.py --model vit_base_patch16 --variant bf16 --steps 10 --warmup 4 --out gpurun_out/torch_gpu_baseline_n8.jsonl > gpurun_out/torch_n8_vit.log 2>&1; echo "torch vit rc=$?"
tail -1 gpurun_out/torch_n8_vit.log | cut -c1-300
