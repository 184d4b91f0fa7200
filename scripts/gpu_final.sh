#!/bin/bash
# End-of-round evidence: GPU tests, isolated kernel timings, bench lines (ours + reference arm), ncu launch list of the
# bench command, ncu --set full of the hot kernels.  Everything lands in gpurun_out/.
set +e
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
ALL_VARIANTS=0 timeout 300 python scripts/bench_ops.py > gpurun_out/bench_ops.log 2>&1; cat gpurun_out/bench_ops.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'knn_stream|knn_self|knn_normalize|mr_aggregate|bn_' -c 40 \
  -o gpurun_out/prof_ops -f python scripts/ncu_ops.py 512 1 > gpurun_out/ncu_ops.log 2>&1; echo "ncu rc=$?"
python scripts/ncu_summary.py gpurun_out/prof_ops.ncu-rep > gpurun_out/ncu_summary.txt 2>&1; cat gpurun_out/ncu_summary.txt
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 5200 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1; echo "launchlist rc=$?"
