#!/bin/bash
# round-2 GPU pass T: epilogue with 1-3 staging slices per half + aux operand by TMA: parity, knob sweep, full suite, bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "aux_operand_by_tma or linear or conv" > gpurun_out/pytest_t_ops.log 2>&1; rc=$?
echo "ops pytest rc=$rc"; tail -5 gpurun_out/pytest_t_ops.log | cut -c1-400
if [ $rc -ne 0 ]; then grep -m5 -B2 -A12 "Error\|assert" gpurun_out/pytest_t_ops.log | head -60; exit 1; fi
timeout 600 python tests/profile_gemm_tune.py all > gpurun_out/gemm_tune.log 2>&1; echo "tune rc=$?"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_t_all.log 2>&1; echo "full pytest rc=$?"
tail -6 gpurun_out/pytest_t_all.log | cut -c1-400
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_t.log 2> gpurun_out/bench_t.err; echo "bench rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/bench_t.log'):
    if l.startswith('{'):
        d = json.loads(l)
        print('resnet50 value', round(d['value']), 'ms', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value']), 'roofline', d['roofline'].get('frac'))
        for k in ('vit_base_patch16', 'second_model'):
            if k in d: print(k, {kk: d[k].get(kk) for kk in ('value', 'ms_per_step')})
PY
