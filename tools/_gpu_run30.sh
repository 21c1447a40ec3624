timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for cb in 1 2 3 4; do echo "== HPCLA_COL_BLOCKS=$cb"; HPCLA_COL_BLOCKS=$cb timeout 300 python tools/tune_spmv.py --workload powerlaw --reps 20 2>&1 | grep -v Warn | tail -1; done | tee gpurun_out/r30_tune_powerlaw_colblocks.log
echo "== default" | tee -a gpurun_out/r30_tune_powerlaw_colblocks.log
timeout 300 python tools/tune_spmv.py --workload powerlaw --reps 20 --cusparse 2>&1 | grep -v "Warn\|M = torch" | tail -2 | tee -a gpurun_out/r30_tune_powerlaw_colblocks.log
