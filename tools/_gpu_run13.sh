timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for w in poisson256 poisson256-i64 stencil27 stencil27-f64 laplace2d; do timeout 300 python tools/tune_spmv.py --workload $w --no-direct 2>&1 | grep -v Warn; done | tee gpurun_out/r13_tune_direct.log
