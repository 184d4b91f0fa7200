#!/bin/bash
# End-of-round evidence: GPU tests, isolated kernel timings, bench lines (ours + reference arm), ncu --set full of the
# hot kernels (summarised on the box; the .ncu-rep is removed because gpurun_out/ only travels back under 64 MiB),
# ncu launch list of the bench command.  Everything lands in gpurun_out/.
set +e
TAG=${1:-r01f}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
ALL_VARIANTS=0 timeout 300 python scripts/bench_ops.py > gpurun_out/bench_ops.log 2>&1; cat gpurun_out/bench_ops.log
timeout 600 python scripts/bench_rows.py > gpurun_out/bench_rows.log 2>&1; cat gpurun_out/bench_rows.log
timeout 300 python scripts/bench_bn.py > gpurun_out/bench_bn.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-300
timeout 900 ncu --set full --clock-control none -k regex:'knn_stream|knn_self|knn_normalize|mr_aggregate|bn_' -c 44 \
  -o gpurun_out/prof_ops -f python scripts/ncu_ops.py 512 1 > gpurun_out/ncu_ops.log 2>&1; echo "ncu rc=$?"
python scripts/ncu_summary.py gpurun_out/prof_ops.ncu-rep > gpurun_out/ncu_summary.txt 2>&1; cat gpurun_out/ncu_summary.txt
python scripts/ncu_stalls.py gpurun_out/prof_ops.ncu-rep > gpurun_out/ncu_stalls.txt 2>&1
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 5200 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1; echo "launchlist rc=$?"
python scripts/summarize_profiles.py $TAG > gpurun_out/summarize.log 2>&1; cp profiles/${TAG}_* gpurun_out/ 2>/dev/null
rm -f gpurun_out/prof_ops.ncu-rep
du -sh gpurun_out
