B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1j_launches_poisson.csv $B > gpurun_out/ncu1.log 2>&1
echo "launch list rc=$?"
$B > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:spmv_rowwalk -s 3 -c 2 -o gpurun_out/r1j_prof_poisson_rowwalk $B > gpurun_out/ncu2.log 2>&1
echo "full poisson rc=$?"
B="python bench.py --workload stencil27-192 --steps 5 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:spmv_rowwalk -s 3 -c 2 -o gpurun_out/r1j_prof_stencil27_rowwalk $B > gpurun_out/ncu3.log 2>&1
echo "full stencil27 rc=$?"
python bench.py --steps 200 --warmup 10 > gpurun_out/r1j_bench_n1.json 2> gpurun_out/r1j_bench_n1.err; echo "bench rc=$?"
