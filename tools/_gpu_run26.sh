D=$PWD/linearalgebrampi.jl_b200
HPCLA_LIB=$D/libhpcla_b200_old059.so timeout 200 python tools/_ab_spmm.py 2>&1 | grep -v Warn
timeout 200 python tools/_ab_spmm.py 2>&1 | grep -v Warn
HPCLA_NO_RUNS=1 timeout 200 python tools/_ab_spmm.py 2>&1 | grep -v Warn
