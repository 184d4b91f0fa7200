#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 70 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_last.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_last.log; grep -E "^FAILED|^ERROR" gpurun_out/pytest_gpu_last.log | head -5
timeout 40 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_last.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_last.log | cut -c1-200
