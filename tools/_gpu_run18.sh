D=$PWD/linearalgebrampi.jl_b200
for lib in libhpcla_b200.so libhpcla_b200_g4.so libhpcla_b200_c8.so libhpcla_b200_c8g1.so; do
echo "== $lib"
HPCLA_LIB=$D/$lib timeout 300 python tools/tune_spmv.py --workload powerlaw --reps 20 2>&1 | grep -v Warn
done | tee gpurun_out/r18_tune_powerlaw_variants.log
echo "== default lib, HPCLA_L2_FETCH=32" | tee -a gpurun_out/r18_tune_powerlaw_variants.log
HPCLA_L2_FETCH=32 timeout 300 python tools/tune_spmv.py --workload powerlaw --reps 20 2>&1 | grep -v Warn | tee -a gpurun_out/r18_tune_powerlaw_variants.log
echo "== default lib, HPCLA_L2_FETCH=128" | tee -a gpurun_out/r18_tune_powerlaw_variants.log
HPCLA_L2_FETCH=128 timeout 300 python tools/tune_spmv.py --workload powerlaw --reps 20 2>&1 | grep -v Warn | tee -a gpurun_out/r18_tune_powerlaw_variants.log
