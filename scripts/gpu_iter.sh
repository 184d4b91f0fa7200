#!/bin/bash
# One development iteration on the GPU box: k-NN kernel cross-check, GPU tests, A/B timings of the kernel
# variants, the bench line and a small ncu capture (the .ncu-rep is summarised on the box and removed: gpurun_out/
# only travels back when it is under 64 MiB).
set +e
mkdir -p gpurun_out
timeout 300 python scripts/tc_check.py > gpurun_out/tc_check.log 2>&1; echo "tc_check rc=$?"; grep -v "agree=1.000000" gpurun_out/tc_check.log | tail -25
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
ALL_VARIANTS=0 timeout 300 python scripts/bench_ops.py > gpurun_out/bench_ops.log 2>&1; cat gpurun_out/bench_ops.log
echo "== vote epilogue"; GRAFP_KNN_EPI=vote ALL_VARIANTS=0 timeout 300 python scripts/bench_ops.py 2>&1 | grep "N=" | cut -c1-60
echo "== BN128 tiles"; GRAFP_KNN_BN128=1 ALL_VARIANTS=0 timeout 300 python scripts/bench_ops.py 2>&1 | grep "N=" | cut -c1-60
for v in 17 18 2 8; do echo "== bwd variant $v"; GRAFP_MR_BWD_VARIANT=$v ALL_VARIANTS=0 timeout 300 python scripts/bench_ops.py 2>&1 | grep "N=" | cut -c100-; done
timeout 300 python scripts/bench_bn.py > gpurun_out/bench_bn.log 2>&1; cat gpurun_out/bench_bn.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench.log
timeout 600 ncu --set full --clock-control none -k regex:'knn_stream|knn_self|mr_aggregate_bwd|finalize' -c 24 \
  -o gpurun_out/prof_ops -f python scripts/ncu_ops.py 512 1 > gpurun_out/ncu_ops.log 2>&1; echo "ncu rc=$?"
python scripts/ncu_summary.py gpurun_out/prof_ops.ncu-rep > gpurun_out/ncu_summary.txt 2>&1; cat gpurun_out/ncu_summary.txt
python scripts/ncu_stalls.py gpurun_out/prof_ops.ncu-rep > gpurun_out/ncu_stalls.txt 2>&1; head -60 gpurun_out/ncu_stalls.txt
rm -f gpurun_out/prof_ops.ncu-rep
du -sh gpurun_out
