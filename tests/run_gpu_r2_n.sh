#!/bin/bash
# round-2 GPU pass N: SAM (GEMM rel-pos) + ViT dropout/checkpoint tests, smoke, ncu launch lists with DRAM traffic, attention capture, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sam_gpu.py tests/test_vit_gpu.py -m gpu -q > gpurun_out/pytest_sam.log 2>&1; echo "sam/vit rc=$?"
tail -4 gpurun_out/pytest_sam.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
for m in resnet50 vit_base_patch16; do
  timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv \
     --log-file gpurun_out/launches_$m.csv python tests/profile_step.py --model $m > gpurun_out/ncu_launch_$m.log 2>&1; echo "ncu $m rc=$?"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_(fwd|bwd)_sm100" -c 3 -o gpurun_out/prof_attn_r2c -f python tests/profile_attn.py > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 --sam --dump-ops gpurun_out/ops_r50.csv > gpurun_out/bench_r50.log 2> gpurun_out/bench_r50.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench_r50.err
cut -c1-300 gpurun_out/bench_r50.log
