python -m pytest tests -m gpu -x -q > gpurun_out/r3_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r3_pytest_gpu.log
P=$PWD/linearalgebrampi.jl_b200/libhpcla_b200_pred.so
python tools/tune_spmv.py --workload stencil27 --sweep 2:3456,2:3072,2:2048,1:3456,4:3456,2:2304,1:2048 > gpurun_out/r3_tune_stencil27.log 2>&1
HPCLA_LIB=$P python tools/tune_spmv.py --workload stencil27 --sweep 2:3456,2:2460,1:3456,1:2460,4:3456 >> gpurun_out/r3_tune_stencil27.log 2>&1
python tools/tune_spmv.py --workload stencil27-f64 --lanes 1,2,4 --sweep 1:3456,2:3456,1:2304 --cusparse >> gpurun_out/r3_tune_stencil27.log 2>&1
python tools/tune_spmv.py --workload powerlaw --cusparse --reps 20 > gpurun_out/r3_tune_powerlaw.log 2>&1
HPCLA_X_PERSIST=100 python tools/tune_spmv.py --workload powerlaw --reps 20 >> gpurun_out/r3_tune_powerlaw.log 2>&1
HPCLA_X_PERSIST=60 python tools/tune_spmv.py --workload powerlaw --reps 20 >> gpurun_out/r3_tune_powerlaw.log 2>&1
grep -v Warn gpurun_out/r3_tune_*.log | grep -v "M = torch"
