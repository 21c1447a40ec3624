#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 150 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "direct_halo" > gpurun_out/r2y_pytest.log 2>&1; echo "rc=$?"
tail -n 6 gpurun_out/r2y_pytest.log
