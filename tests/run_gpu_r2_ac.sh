#!/bin/bash
# round-2 GPU pass AC: last changes (optimizer shadow following, batched num_batches_tracked, KD test) + quick bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_distill_gpu.py tests/test_optim_gpu.py tests/test_resnet_gpu.py tests/test_darknet_gpu.py tests/test_determinism_gpu.py tests/test_reference_cuda_parity_gpu.py -m gpu -q > gpurun_out/pytest_ac.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_ac.log | cut -c1-300; grep -E "^(FAILED|ERROR)" gpurun_out/pytest_ac.log | head
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-second-model > gpurun_out/bench_ac.log 2> gpurun_out/bench_ac.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_ac.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_ac.log'):
    if l.startswith('{'):
        d = json.loads(l)
        print('resnet50 value', round(d['value']), 'ms', round(d['ms_per_step'], 2), 'eager', round(d.get('eager_ms_per_step'), 2), 'e2e', round(d['e2e']['value']), 'u8', round((d.get('e2e_uint8_input') or {}).get('value', 0)), 'launches', d['gpu_launches'])
PY
