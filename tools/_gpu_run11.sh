timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for w in poisson256-spmm4 poisson256-spmm8 poisson256-spmm16; do
timeout 200 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --workload $w > gpurun_out/r11_bench_$w.json 2> gpurun_out/r11_bench_$w.err; echo "bench $w rc=$?"; tail -2 gpurun_out/r11_bench_$w.err
python -c "
import json; d=json.loads(open('gpurun_out/r11_bench_$w.json').read().strip().splitlines()[-1]); print(d['config']['workload'], d['ms_per_step'], d['value'], d['achieved_gbs'], d['gpu_launches'])"
done
timeout 300 python tools/tune_spmv.py --workload stencil27-f64 --lanes 2,4 2>&1 | grep -v Warn
