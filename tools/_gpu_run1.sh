set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2_pytest_gpu.log
python tools/tune_spmv.py --workload poisson256 --kinds general --lanes 1,2 --windows 1536,1784,2048 --cusparse > gpurun_out/r2_tune_poisson.log 2>&1
HPCLA_LIB=$PWD/linearalgebrampi.jl_b200/libhpcla_b200_pred.so python tools/tune_spmv.py --workload poisson256 >> gpurun_out/r2_tune_poisson.log 2>&1
python tools/tune_spmv.py --workload stencil27 --kinds general --lanes 1,2,4,8,16 --sweep 4:1296,4:2160,4:2592,8:864,8:1296,2:2592 --cusparse > gpurun_out/r2_tune_stencil27.log 2>&1
HPCLA_LIB=$PWD/linearalgebrampi.jl_b200/libhpcla_b200_pred.so python tools/tune_spmv.py --workload stencil27 --lanes 2,4,8 >> gpurun_out/r2_tune_stencil27.log 2>&1
python tools/tune_spmv.py --workload poisson256-i64 --cusparse > gpurun_out/r2_tune_others.log 2>&1
python tools/tune_spmv.py --workload laplace2d --cusparse >> gpurun_out/r2_tune_others.log 2>&1
python tools/tune_spmv.py --workload powerlaw --kinds rowwalk --cusparse --reps 20 >> gpurun_out/r2_tune_others.log 2>&1
cat gpurun_out/r2_tune_*.log
