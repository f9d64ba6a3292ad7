#!/bin/bash
# round-2 GPU pass X: MAE pre-training on the runtime (token gather kernels, step vs oracle, fused AdamW) + bench --mae
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mae_gpu.py -m gpu -q -x -s > gpurun_out/pytest_x_mae.log 2>&1; echo "mae pytest rc=$?"
grep -E "^mae |passed|failed|Error" gpurun_out/pytest_x_mae.log | head -20
tail -30 gpurun_out/pytest_x_mae.log | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-second-model --mae > gpurun_out/bench_x.log 2> gpurun_out/bench_x.err; echo "bench rc=$?"
tail -5 gpurun_out/bench_x.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_x.log'):
    if l.startswith('{'):
        d = json.loads(l)
        print('resnet50 value', round(d['value']), 'ms', round(d['ms_per_step'], 2))
        v = d.get('vit_base_mae')
        if v: print('mae', round(v['value'], 1), round(v['ms_per_step'], 2), 'eager', round(v.get('eager_ms_per_step'), 2), 'e2e', round(v['e2e']['value'], 1), 'roofline', v['roofline'].get('frac'), 'launches', v['gpu_launches'], v['value_mode'])
PY
