D=$PWD/linearalgebrampi.jl_b200
for lib in libhpcla_b200_old059.so libhpcla_b200.so; do
HPCLA_LIB=$D/$lib ncu --set full --clock-control none -k regex:spmm_rowwalk -s 1 -c 2 -o gpurun_out/r24_spmm_${lib%.so} python tools/_ab_spmm.py > gpurun_out/r24_ncu_${lib%.so}.log 2>&1; echo "$lib rc=$?"
done
