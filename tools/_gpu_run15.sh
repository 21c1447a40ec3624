D=$PWD/linearalgebrampi.jl_b200
for lib in libhpcla_b200.so libhpcla_b200_drop1.so libhpcla_b200_onebar.so; do
echo "== $lib"
for w in poisson256 stencil27 stencil27-f64; do HPCLA_LIB=$D/$lib timeout 300 python tools/tune_spmv.py --workload $w 2>&1 | grep -v Warn; done
done | tee gpurun_out/r15_tune_direct_variants.log
