timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/tune_spmv.py --workload poisson256 --prefetch 296,592,1184,2368,4736 2>&1 | grep -v Warn | tee gpurun_out/r9_tune_poisson.log
timeout 300 python tools/tune_spmv.py --workload poisson256-i64 --prefetch 592,1184,2368 2>&1 | grep -v Warn | tee -a gpurun_out/r9_tune_poisson.log
timeout 300 python tools/tune_spmv.py --workload stencil27 --prefetch 222,444,888,1776 2>&1 | grep -v Warn | tee gpurun_out/r9_tune_stencil27.log
timeout 300 python tools/tune_spmv.py --workload stencil27-f64 --prefetch 370,740,1480 2>&1 | grep -v Warn | tee -a gpurun_out/r9_tune_stencil27.log
timeout 300 python tools/tune_spmv.py --workload laplace2d --prefetch 592,1184 2>&1 | grep -v Warn | tee -a gpurun_out/r9_tune_poisson.log
