#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
timeout 45 python -m pytest tests/test_gpu_round2.py -m gpu -x -q --timeout 40 -k "graph_replay or plan_import_route and int32 and 2" 2>&1 | tail -n 2
