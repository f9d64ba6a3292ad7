#!/bin/bash
# round-2 GPU pass O (4 GPUs): scaling check of the captured data-parallel step + torch DDP baseline at N=4
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29531 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/bench_n4.log 2> gpurun_out/bench_n4.err; echo "bench n4 rc=$?"
tail -c 300 gpurun_out/bench_n4.err; cut -c1-260 gpurun_out/bench_n4.log
timeout 300 $TR --master-port 29533 baseline/torch_gpu_baseline.py --model resnet50 --variant bf16 --steps 15 --warmup 5 --out gpurun_out/torch_gpu_baseline_n4.jsonl > gpurun_out/torch_n4_bf16.log 2>&1; echo "torch bf16 rc=$?"
tail -1 gpurun_out/torch_n4_bf16.log | cut -c1-300
