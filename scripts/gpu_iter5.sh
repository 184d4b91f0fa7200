#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log; grep -E "^FAILED|^ERROR" gpurun_out/pytest_gpu.log | head
ALL_VARIANTS=0 timeout 300 python scripts/bench_ops.py > gpurun_out/bench_ops.log 2>&1; cat gpurun_out/bench_ops.log
echo "== bwd variant 20"; GRAFP_MR_BWD_VARIANT=20 ALL_VARIANTS=0 timeout 300 python scripts/bench_ops.py 2>&1 | grep "N=" | cut -c100-
timeout 600 python scripts/bench_rows.py fp > gpurun_out/bench_rows_fp.log 2>&1; cat gpurun_out/bench_rows_fp.log | tail -8
du -sh gpurun_out
