timeout 400 python -m pytest tests/test_gpu_nccl.py -x -q > gpurun_out/r35_pytest_nccl.log 2>&1; echo "pytest nccl rc=$?"
tail -8 gpurun_out/r35_pytest_nccl.log | cut -c1-400
TR="timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/r35_bench_n2.json 2> gpurun_out/r35_bench_n2.err; echo "bench2 rc=$?"
HPCLA_NCCL_MAX_CTAS=2 $TR bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/r35_bench_n2_ctas2.json 2> gpurun_out/r35_bench_n2_ctas2.err; echo "bench2 ctas2 rc=$?"
HPCLA_NCCL_MAX_CTAS=8 $TR bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/r35_bench_n2_ctas8.json 2> gpurun_out/r35_bench_n2_ctas8.err; echo "bench2 ctas8 rc=$?"
$TR bench.py --gpus 2 --steps 50 --warmup 3 --workload cg-512 > gpurun_out/r35_bench_n2_cg.json 2> gpurun_out/r35_bench_n2_cg.err; echo "bench2 cg rc=$?"
$TR tools/bench_next_rows.py 2>&1 | grep "^{" > gpurun_out/r35_next_rows_n2.jsonl; echo "next rows rc=$?"; cat gpurun_out/r35_next_rows_n2.jsonl
for f in gpurun_out/r35_bench_n2*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","n_gpus","ms_per_step","achieved_gbs","gpu_launches")}, d["roofline"]["frac"], d["e2e"] and (d["e2e"]["value"], d["e2e"]["ms_per_step"]), d["config"]["workload"])
except Exception as e: print("ERR", e)
PY
done
