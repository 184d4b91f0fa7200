#!/bin/bash
# One GPU session: correctness of the tensor-core k-NN kernels, the GPU test suite, isolated kernel
# timings, the benchmark line and the ncu captures.  Everything lands in gpurun_out/.
set +e
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
cp MEASURED_PEAKS.json gpurun_out/ 2>/dev/null
timeout 300 python scripts/tc_check.py > gpurun_out/tc_check.log 2>&1; echo "tc_check rc=$?"
tail -12 gpurun_out/tc_check.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python scripts/bench_ops.py > gpurun_out/bench_ops.log 2>&1; echo "bench_ops rc=$?"
cat gpurun_out/bench_ops.log
GRAFP_KNN_NO_NH4=1 ALL_VARIANTS=0 timeout 300 python scripts/bench_ops.py > gpurun_out/bench_ops_nh2.log 2>&1; grep "N= 1024" gpurun_out/bench_ops_nh2.log
timeout 120 python scripts/measure_peaks.py > gpurun_out/peaks_self.json 2>&1; cat gpurun_out/peaks_self.json
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
tail -1 gpurun_out/bench.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'knn_stream|knn_self|knn_normalize|mr_aggregate|mr_bwd' -c 20 \
  -o gpurun_out/prof_ops -f python scripts/ncu_ops.py 512 1 > gpurun_out/ncu_ops.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1; echo "launchlist rc=$?"
python scripts/ncu_summary.py gpurun_out/prof_ops.ncu-rep > gpurun_out/ncu_summary.txt 2>&1; tail -40 gpurun_out/ncu_summary.txt
