#!/bin/bash
# round-2 GPU pass Q: remaining GPU tests after the first failure, prefetcher test, e2e with the slot prefetcher
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/pytest_gpu.log | cut -c1-400
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r50_q.log 2> gpurun_out/bench_r50_q.err; echo "bench rc=$?"
tail -c 300 gpurun_out/bench_r50_q.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_r50_q.log'):
    if l.startswith('{'):
        d = json.loads(l)
        print('R50', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
        v = d['vit_base_patch16']; print('ViT', v['value'], v['ms_per_step'], 'e2e', v['e2e']['value'], v['e2e']['ms_per_step'])
PY
