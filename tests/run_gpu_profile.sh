#!/bin/bash
# ncu evidence for profiles/: per-launch device time of one step (both models) and full captures
# of the tcgen05 GEMM engine on representative launches.  Run on the GPU box (1 GPU).
mkdir -p gpurun_out
for m in resnet50 vit_base_patch16; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
      --log-file gpurun_out/launches_$m.csv python tests/profile_step.py --model $m > gpurun_out/ncu_launch_$m.log 2>&1
  echo "$m launches: $(grep -c gpu__time_duration gpurun_out/launches_$m.csv)"
done
# full sets: first GEMM launches of the ResNet-50 forward (stem GEMM, layer1 1x1 / 3x3 convs)
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_sm100 -c 8 \
    -o gpurun_out/prof_gemm_r50 python tests/profile_step.py --model resnet50 --fwd-only > gpurun_out/ncu_full_r50.log 2>&1
# ViT: qkv / proj / fc1 / fc2 forward of the first block (+ patch embedding)
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_sm100 -c 5 \
    -o gpurun_out/prof_gemm_vit python tests/profile_step.py --model vit_base_patch16 --fwd-only > gpurun_out/ncu_full_vit.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attn_ -c 2 \
    -o gpurun_out/prof_attn_vit python tests/profile_step.py --model vit_base_patch16 > gpurun_out/ncu_full_attn.log 2>&1
ls -la gpurun_out/*.ncu-rep
