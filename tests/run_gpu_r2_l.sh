#!/bin/bash
# round-2 GPU pass L (2 GPUs): DETR tests, DDP tests, bench at N=2 (eager and graph-captured), torch DDP baseline at N=2
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_n2.txt
timeout 600 python -m pytest tests/test_detr_gpu.py tests/test_ddp_gpu.py -m gpu -q -s > gpurun_out/pytest_detr.log 2>&1; echo "detr+ddp rc=$?"
grep -n "passed\|failed\|Error\|detr \|skipped" gpurun_out/pytest_detr.log | head -20
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"
tail -c 400 gpurun_out/bench_n2.err; cut -c1-260 gpurun_out/bench_n2.log
timeout 600 $TR --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --graph-ddp --no-second-model > gpurun_out/bench_n2_graph.log 2> gpurun_out/bench_n2_graph.err; echo "bench n2 graph rc=$?"
tail -c 600 gpurun_out/bench_n2_graph.err; cut -c1-260 gpurun_out/bench_n2_graph.log
for v in bf16 as_shipped; do
  timeout 600 $TR --master-port 29513 baseline/torch_gpu_baseline.py --model resnet50 --variant $v --steps 15 --warmup 5 --out gpurun_out/torch_gpu_baseline_n2.jsonl > gpurun_out/torch_n2_$v.log 2>&1; echo "torch $v rc=$?"
  tail -1 gpurun_out/torch_n2_$v.log | cut -c1-300
done
timeout 600 $TR --master-port 29514 baseline/torch_gpu_baseline.py --model vit_base_patch16 --variant bf16 --steps 10 --warmup 4 --out gpurun_out/torch_gpu_baseline_n2.jsonl > gpurun_out/torch_n2_vit.log 2>&1; echo "torch vit rc=$?"
tail -1 gpurun_out/torch_n2_vit.log | cut -c1-300
