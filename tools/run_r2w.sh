#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 70 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 60 -k "sparse_times_sparse or device_transpose or reference_product" > gpurun_out/r2w_pytest.log 2>&1; echo "rc=$?"
tail -n 3 gpurun_out/r2w_pytest.log
