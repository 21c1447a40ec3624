timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for w in poisson256 poisson256-i64 stencil27 stencil27-f64 laplace2d; do timeout 300 python tools/tune_spmv.py --workload $w 2>&1 | grep -v Warn; done | tee gpurun_out/r16_tune_defaults.log
timeout 200 python bench.py --steps 200 --warmup 10 > gpurun_out/r16_bench_n1.json 2> gpurun_out/r16_bench_n1.err; echo "bench rc=$?"; cat gpurun_out/r16_bench_n1.json
