python -m pytest tests -m gpu -x -q > gpurun_out/r4_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r4_pytest_gpu.log
python bench.py --steps 100 --warmup 5 > gpurun_out/r4_bench_n1.json 2> gpurun_out/r4_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r4_bench_n1.err
cat gpurun_out/r4_bench_n1.json
for chunk in 1024 2048 8192 16384; do HPCLA_STAGE_CHUNK_KB=$chunk python bench.py --steps 20 --warmup 3 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('chunk_kb', $chunk, d['e2e'])"; done
D=$PWD/linearalgebrampi.jl_b200
for lib in libhpcla_b200_pred.so libhpcla_b200_pred8.so; do
echo "== $lib"
HPCLA_LIB=$D/$lib python tools/tune_spmv.py --workload poisson256 2>&1 | grep -v Warn
HPCLA_LIB=$D/$lib python tools/tune_spmv.py --workload stencil27 --sweep 2:3456,4:3456,4:1728,1:3456 2>&1 | grep -v Warn
HPCLA_LIB=$D/$lib python tools/tune_spmv.py --workload stencil27-f64 --sweep 2:3456,1:3456,1:6912,4:1728 2>&1 | grep -v Warn
done
echo "== default"
python tools/tune_spmv.py --workload stencil27-f64 --sweep 1:6912,2:6912 2>&1 | grep -v Warn
