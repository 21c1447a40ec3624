#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 60 python tools/debug_direct.py > gpurun_out/r2g_debug.log 2>&1; echo "rc=$?" >> gpurun_out/r2g_debug.log
tail -n 12 gpurun_out/r2g_debug.log
timeout 200 python -m pytest tests/test_gpu_round2.py -m gpu -x -q --timeout 60 -k "direct_halo" > gpurun_out/r2g_pytest.log 2>&1; echo "tests rc=$?"
tail -n 4 gpurun_out/r2g_pytest.log
