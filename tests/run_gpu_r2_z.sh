#!/bin/bash
# round-2 GPU pass Z (final evidence): smoke(), ncu launch lists with DRAM bytes of one step (both models), full-set captures of the
# top kernels, the full GPU suite and the default bench line
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_z.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_z.log
for m in resnet50 vit_base_patch16; do
  timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv \
      --log-file gpurun_out/launches_$m.csv python tests/profile_step.py --model $m > gpurun_out/ncu_launch_$m.log 2>&1
  echo "$m launches: $(grep -c gpu__time_duration gpurun_out/launches_$m.csv)"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_sm100 --launch-skip 3 --launch-count 3 -f -o gpurun_out/gemm_epi_final python tests/profile_one_gemm.py > gpurun_out/ncu_z_gemm.log 2>&1; echo "ncu gemm rc=$?"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_z_all.log 2>&1; echo "full pytest rc=$?"
tail -4 gpurun_out/pytest_z_all.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/bench_z.log 2> gpurun_out/bench_z.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_z_ref.log 2> gpurun_out/bench_z_ref.err; echo "ref arm rc=$?"; cut -c1-400 gpurun_out/bench_z_ref.log
python - <<'PY'
import json
for l in open('gpurun_out/bench_z.log'):
    if l.startswith('{'):
        d = json.loads(l)
        u8 = d.get('e2e_uint8_input') or {}
        print('resnet50 value', round(d['value']), 'ms', round(d['ms_per_step'], 2), 'eager', round(d.get('eager_ms_per_step'), 2), 'e2e', round(d['e2e']['value']), 'e2e_u8', round(u8.get('value', 0)), 'roofline', d['roofline'].get('frac'), 'traffic', d['roofline'].get('traffic'), 'launches', d['gpu_launches'], 'clocks', d['clocks'])
        print('cpu_baseline', (d.get('cpu_baseline') or {}).get('value'))
        v = d.get('vit_base_patch16')
        if v: print('vit', round(v['value'], 1), round(v['ms_per_step'], 2), 'e2e', round(v['e2e']['value'], 1), 'u8', round((v.get('e2e_uint8_input') or {}).get('value', 0)), 'roofline', v['roofline'].get('frac'))
PY
