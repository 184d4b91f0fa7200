#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== default (vote-gated selection for K <= 8)" > gpurun_out/bench_ops_ab.log
ALL_VARIANTS=0 timeout 200 python scripts/bench_ops.py 2>&1 | grep "N=" >> gpurun_out/bench_ops_ab.log
echo "== GRAFP_KNN_EPI=queue" >> gpurun_out/bench_ops_ab.log
GRAFP_KNN_EPI=queue ALL_VARIANTS=0 timeout 200 python scripts/bench_ops.py 2>&1 | grep "N=" | cut -c1-60 >> gpurun_out/bench_ops_ab.log
cat gpurun_out/bench_ops_ab.log
