#!/bin/bash
# quick K3 check: aggregation parity tests + isolated timings + ncu of the backward kernel
set +e
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "mr_aggregate or aggregate or smoke or encoder" > gpurun_out/pytest_k3.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_k3.log
ALL_VARIANTS=0 timeout 300 python scripts/bench_ops.py > gpurun_out/bench_ops_k3.log 2>&1; cat gpurun_out/bench_ops_k3.log
bash scripts/gpu_ncu_one.sh mr_bwd_slice 4
