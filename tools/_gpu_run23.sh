timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
D=$PWD/linearalgebrampi.jl_b200
for lib in libhpcla_b200_old059.so libhpcla_b200.so; do HPCLA_LIB=$D/$lib timeout 200 python tools/_ab_spmm.py 2>&1 | grep -v Warn; done
