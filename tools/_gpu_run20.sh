timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload stencil27-192-T > gpurun_out/r20_bench_s27T.json 2> gpurun_out/r20_bench_s27T.err; echo "bench T rc=$?"; tail -3 gpurun_out/r20_bench_s27T.err
HPCLA_TRANSPOSE=host timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload stencil27-192-T > gpurun_out/r20_bench_s27T_host.json 2> gpurun_out/r20_bench_s27T_host.err; echo "bench T host rc=$?"
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload stencil27-192 > gpurun_out/r20_bench_s27.json 2> gpurun_out/r20_bench_s27.err
for f in r20_bench_s27T r20_bench_s27T_host r20_bench_s27; do python -c "
import json; d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1]); print('$f', d['config']['setup_s'], d['ms_per_step'], d['value'])"; done
