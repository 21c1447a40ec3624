timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "synthetic_stencils or staged or sparse_times_dense or device_transpose or random_ragged" > gpurun_out/r31_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|Invalid|passed|failed|error" gpurun_out/r31_memcheck.log | head -20
tail -5 gpurun_out/r31_memcheck.log
