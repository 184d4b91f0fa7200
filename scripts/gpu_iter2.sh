#!/bin/bash
# Development iteration: GPU tests, kernel timings, the rows bench.py does not exercise.
set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
ALL_VARIANTS=0 timeout 300 python scripts/bench_ops.py > gpurun_out/bench_ops.log 2>&1; cat gpurun_out/bench_ops.log
timeout 600 python scripts/bench_rows.py > gpurun_out/bench_rows.log 2>&1; echo "rows rc=$?"; cat gpurun_out/bench_rows.log
du -sh gpurun_out
