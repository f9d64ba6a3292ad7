#!/bin/bash
mkdir -p gpurun_out
python tests/profile_attn.py 2>&1 | tail -1
timeout 600 python -m pytest tests/test_attn_gpu.py tests/test_vit_ops_gpu.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -6 | cut -c1-400
timeout 900 python -m pytest tests/test_sam_gpu.py -m gpu -q --tb=short -p no:cacheprovider -s 2>&1 | tail -40 | cut -c1-900
timeout 900 python -m pytest tests/test_van_gpu.py -m gpu -q --tb=short -p no:cacheprovider -s -k stagewise 2>&1 | tail -6 | cut -c1-700
