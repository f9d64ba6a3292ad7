#!/bin/bash
# round-2 GPU pass V: compile-time epilogue variants + fast GELU terms: parity, sweep; optimizer tests; ncu source capture of attention; full suite; bench
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_ops_gpu.py tests/test_vit_ops_gpu.py tests/test_optim_gpu.py -m gpu -q -x > gpurun_out/pytest_v_ops.log 2>&1; rc=$?
echo "ops pytest rc=$rc"; tail -4 gpurun_out/pytest_v_ops.log | cut -c1-400
if [ $rc -ne 0 ]; then grep -m3 -B5 -A25 "Error" gpurun_out/pytest_v_ops.log | head -80; fi
timeout 600 python tests/profile_gemm_tune.py all > gpurun_out/gemm_tune_v.log 2>&1; echo "tune rc=$?"
timeout 200 python tests/profile_epilogue.py > gpurun_out/profile_epilogue_v.log 2>&1; cat gpurun_out/profile_epilogue_v.log
timeout 200 python tests/profile_attn.py --iters 10 > gpurun_out/profile_attn_v.log 2>&1; cat gpurun_out/profile_attn_v.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:attn_ --launch-skip 6 --launch-count 3 -f -o gpurun_out/attn_src python tests/profile_attn.py --iters 1 > gpurun_out/ncu_v.log 2>&1; echo "ncu rc=$?"
if [ $rc -eq 0 ]; then
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_v_all.log 2>&1; echo "full pytest rc=$?"
tail -6 gpurun_out/pytest_v_all.log | cut -c1-400
fi
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --dump-ops gpurun_out/ops_v.csv > gpurun_out/bench_v.log 2> gpurun_out/bench_v.err; echo "bench rc=$?"
tail -5 gpurun_out/bench_v.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_v.log'):
    if l.startswith('{'):
        d = json.loads(l)
        print('resnet50 value', round(d['value']), 'ms', round(d['ms_per_step'], 2), 'eager', d.get('eager_ms_per_step'), 'e2e', round(d['e2e']['value']), 'roofline', d['roofline'].get('frac'), 'launches', d['gpu_launches'])
        v = d.get('vit_base_patch16')
        if v: print('vit', round(v['value']), round(v['ms_per_step'], 2), 'eager', v.get('eager_ms_per_step'), 'e2e', round(v['e2e']['value']), 'roofline', v['roofline'].get('frac'), 'launches', v['gpu_launches'])
PY
