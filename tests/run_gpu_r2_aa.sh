#!/bin/bash
# round-2 GPU pass AA: BatchNorm statistics by four extra warps (VAR_STATS_BF16): parity, conv sweep, convnet suites, bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "statistics or conv_fprop or bn_" > gpurun_out/pytest_aa_ops.log 2>&1; rc=$?
echo "ops pytest rc=$rc"; tail -3 gpurun_out/pytest_aa_ops.log | cut -c1-300
if [ $rc -ne 0 ]; then grep -m3 -B5 -A25 "Error\|assert" gpurun_out/pytest_aa_ops.log | head -60; exit 1; fi
timeout 600 python tests/profile_gemm_tune.py conv 2>&1 | grep -E "fprop" > gpurun_out/gemm_tune_aa.log; grep "stats" gpurun_out/gemm_tune_aa.log | grep default
SAICV_GEMM_INLINE_STATS=1 timeout 600 python tests/profile_gemm_tune.py conv 2>&1 | grep -E "fprop \+ stats" | grep default | sed 's/default/inline /' 
timeout 900 python -m pytest tests/test_resnet_gpu.py tests/test_darknet_gpu.py tests/test_determinism_gpu.py tests/test_detr_gpu.py tests/test_van_gpu.py -m gpu -q > gpurun_out/pytest_aa_nets.log 2>&1; echo "nets pytest rc=$?"
tail -3 gpurun_out/pytest_aa_nets.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-second-model > gpurun_out/bench_aa.log 2> gpurun_out/bench_aa.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_aa.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_aa.log'):
    if l.startswith('{'):
        d = json.loads(l)
        print('resnet50 value', round(d['value']), 'ms', round(d['ms_per_step'], 2), 'eager', round(d.get('eager_ms_per_step'), 2), 'e2e', round(d['e2e']['value']), 'roofline', d['roofline'].get('frac'))
PY
