#!/bin/bash
# round-2 GPU pass AB (2 GPUs): does limiting the CTAs NCCL may occupy help the persistent one-CTA-per-SM GEMM kernels? (default: pass Y)
mkdir -p gpurun_out
for c in 4 8; do
NCCL_MAX_CTAS=$c timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2954$c bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_ab_c$c.log 2> gpurun_out/bench_ab_c$c.err; echo "bench NCCL_MAX_CTAS=$c rc=$?"
python - <<PY
import json
for l in open('gpurun_out/bench_ab_c$c.log'):
    if l.startswith('{'):
        d = json.loads(l)
        v = d.get('vit_base_patch16') or {}
        print('NCCL_MAX_CTAS=$c: resnet50', round(d['value']), round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value']), '| vit', round(v.get('value', 0)), round(v.get('ms_per_step', 0), 2), 'e2e', round((v.get('e2e') or {}).get('value', 0)))
PY
done
