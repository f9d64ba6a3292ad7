#!/bin/bash
# round-2 GPU pass U: fused multi-tensor optimizer tests; ncu source-level capture of the epilogue-bound conv launches; bench with the fused optimizer
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_optim_gpu.py -m gpu -q -x > gpurun_out/pytest_u_optim.log 2>&1; echo "optim pytest rc=$?"
tail -15 gpurun_out/pytest_u_optim.log | cut -c1-300
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_sm100 --launch-skip 3 --launch-count 3 -f -o gpurun_out/gemm_epi python tests/profile_one_gemm.py > gpurun_out/ncu_u.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/ncu_u.log
ls -la gpurun_out/*.ncu-rep
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_u.log 2> gpurun_out/bench_u.err; echo "bench rc=$?"
tail -5 gpurun_out/bench_u.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_u.log'):
    if l.startswith('{'):
        d = json.loads(l)
        print('resnet50 value', round(d['value']), 'ms', round(d['ms_per_step'], 2), 'eager', d.get('eager_ms_per_step'), 'e2e', round(d['e2e']['value']), 'roofline', d['roofline'].get('frac'), 'launches', d['gpu_launches'])
        v = d.get('vit_base_patch16')
        if v: print('vit', round(v['value']), round(v['ms_per_step'], 2), 'eager', v.get('eager_ms_per_step'), 'e2e', round(v['e2e']['value']), 'roofline', v['roofline'].get('frac'), 'launches', v['gpu_launches'])
PY
