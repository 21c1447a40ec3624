timeout 120 tools/hbm_read_probe 2>&1 | tee gpurun_out/r8_hbm_read_probe.txt
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r8_bench_n1.json 2> gpurun_out/r8_bench_n1.err; echo "bench1 rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r8_bench_n1.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['achieved_gbs'], d['e2e'])"
