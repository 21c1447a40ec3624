D=$PWD/linearalgebrampi.jl_b200
for lib in libhpcla_b200_old059.so libhpcla_b200.so libhpcla_b200_m6e4.so libhpcla_b200_m5e4.so libhpcla_b200_m4e4.so libhpcla_b200_m5e2.so; do HPCLA_LIB=$D/$lib timeout 200 python tools/_ab_spmm.py 2>&1 | grep -v Warn; done
