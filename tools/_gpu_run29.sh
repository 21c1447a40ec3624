D=$PWD/linearalgebrampi.jl_b200
for lib in libhpcla_b200.so libhpcla_b200_f6.so libhpcla_b200_f5.so; do echo "== $lib"; for w in poisson256 stencil27-f64 laplace2d; do HPCLA_LIB=$D/$lib timeout 200 python tools/tune_spmv.py --workload $w 2>&1 | grep -v Warn | tail -1; done; done
