#!/bin/bash
# round-2 GPU pass S: device-side dropout seeds (DETR / ViT tests), host-link probe, e2e with 1 / 2 / 4 copy streams
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_detr_gpu.py tests/test_vit_gpu.py tests/test_prefetcher_gpu.py tests/test_attn_gpu.py -m gpu -q > gpurun_out/pytest_s.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_s.log | cut -c1-300
timeout 120 python tests/profile_h2d.py > gpurun_out/profile_h2d.log 2>&1; cat gpurun_out/profile_h2d.log
for k in 1 2 4; do
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-second-model --prefetch-streams $k > gpurun_out/bench_ps$k.log 2> gpurun_out/bench_ps$k.err
  python - <<PY
import json
for l in open('gpurun_out/bench_ps$k.log'):
    if l.startswith('{'):
        d = json.loads(l); print('copy streams $k: value', round(d['value']), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'], 2), 'ms')
PY
done
