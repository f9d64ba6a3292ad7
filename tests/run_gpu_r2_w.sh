#!/bin/bash
# round-2 GPU pass W: shortcut add fused into the first conv's dgrad (aux by TMA), smem-split policy from the sweep, uint8 input edge; full suite + bench (+ SAM / DETR sub-records)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_w_all.log 2>&1; echo "full pytest rc=$?"
tail -6 gpurun_out/pytest_w_all.log | cut -c1-400
grep -E "^(FAILED|ERROR)" gpurun_out/pytest_w_all.log | head -20
timeout 900 python bench.py --steps 20 --warmup 3 --sam --detr --dump-ops gpurun_out/ops_w.csv > gpurun_out/bench_w.log 2> gpurun_out/bench_w.err; echo "bench rc=$?"
tail -5 gpurun_out/bench_w.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_w.log'):
    if l.startswith('{'):
        d = json.loads(l)
        u8 = d.get('e2e_uint8_input') or {}
        print('resnet50 value', round(d['value']), 'ms', round(d['ms_per_step'], 2), 'eager', round(d.get('eager_ms_per_step'), 2), 'e2e', round(d['e2e']['value']), 'e2e_u8', round(u8.get('value', 0)), 'roofline', d['roofline'].get('frac'), 'launches', d['gpu_launches'])
        print('cpu_baseline', d.get('cpu_baseline'))
        for k in ('vit_base_patch16', 'sam_h_encoder', 'resnet50_detr'):
            v = d.get(k)
            if v: print(k, round(v['value'], 1), round(v['ms_per_step'], 2), 'eager', round(v.get('eager_ms_per_step'), 2), 'e2e', round(v['e2e']['value'], 1), 'u8', round((v.get('e2e_uint8_input') or {}).get('value', 0)), 'roofline', v['roofline'].get('frac'), 'launches', v['gpu_launches'])
PY
