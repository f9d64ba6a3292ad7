#!/bin/bash
# round-2 GPU pass AD: optimizer resume (load_state_dict mid-run) + smoke on the freshly rebuilt library
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_optim_gpu.py -m gpu -q > gpurun_out/pytest_ad.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_ad.log | cut -c1-300; grep -E "^(FAILED|ERROR)|Error" gpurun_out/pytest_ad.log | head -5
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
