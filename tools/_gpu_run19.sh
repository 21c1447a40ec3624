timeout 400 python -m pytest tests/test_gpu_nccl.py -x -q > gpurun_out/r19_pytest_nccl.log 2>&1; echo "pytest nccl rc=$?"
tail -12 gpurun_out/r19_pytest_nccl.log | cut -c1-300
TR="timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/r19_bench_n2.json 2> gpurun_out/r19_bench_n2.err; echo "bench2 rc=$?"; tail -3 gpurun_out/r19_bench_n2.err
$TR bench.py --gpus 2 --steps 50 --warmup 5 --workload stencil27-192 > gpurun_out/r19_bench_n2_s27.json 2> gpurun_out/r19_bench_n2_s27.err; echo "bench2 s27 rc=$?"
$TR bench.py --gpus 2 --steps 30 --warmup 3 --workload poisson256-spmm8 > gpurun_out/r19_bench_n2_spmm8.json 2> gpurun_out/r19_bench_n2_spmm8.err; echo "bench2 spmm rc=$?"; tail -3 gpurun_out/r19_bench_n2_spmm8.err
timeout 120 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r19_bench_reference.json 2> gpurun_out/r19_bench_reference.err; echo "reference rc=$?"; cat gpurun_out/r19_bench_reference.json | cut -c1-600
for f in gpurun_out/r19_bench_n2*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","n_gpus","ms_per_step","achieved_gbs","gpu_launches")}, d["roofline"]["frac"], d["e2e"] and (d["e2e"]["value"], d["e2e"]["ms_per_step"]), d["config"]["workload"], d["clocks"])
except Exception as e: print("ERR", e)
PY
done
