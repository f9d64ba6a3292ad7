#!/bin/bash
# round-2 GPU pass Y (2 GPUs): NCCL data-parallel check + bench at N = 2 with the fused optimizer / captured step
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_ddp_gpu.py tests/test_determinism_gpu.py -m gpu -q > gpurun_out/pytest_y_ddp.log 2>&1; echo "ddp pytest rc=$?"
tail -5 gpurun_out/pytest_y_ddp.log | cut -c1-400
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_y_n2.log 2> gpurun_out/bench_y_n2.err; echo "bench N=2 rc=$?"
tail -8 gpurun_out/bench_y_n2.err | cut -c1-300
python - <<'PY'
import json
for l in open('gpurun_out/bench_y_n2.log'):
    if l.startswith('{'):
        d = json.loads(l)
        print('N=2 resnet50 value', round(d['value']), 'ms', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value']), d.get('ddp_check'), '|', d.get('graph_ddp_check'), '|', d['value_mode'])
        v = d.get('vit_base_patch16')
        if v: print('N=2 vit', round(v['value']), round(v['ms_per_step'], 2), 'e2e', round(v['e2e']['value']), v.get('graph_ddp_check'))
PY
