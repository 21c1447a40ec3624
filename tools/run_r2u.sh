#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 32 python -m pytest tests/test_gpu_round2.py -m gpu -x -q --timeout 30 -k "flat_kernel_very_long" 2>&1 | tail -n 2
