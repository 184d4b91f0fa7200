#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "knn" > gpurun_out/pytest_knn.log 2>&1; echo "pytest knn rc=$?"; tail -4 gpurun_out/pytest_knn.log; grep -E "^FAILED|Error" gpurun_out/pytest_knn.log | head -20
timeout 600 python scripts/bench_rows.py stress > gpurun_out/bench_rows_stress.log 2>&1; echo "rows rc=$?"; cat gpurun_out/bench_rows_stress.log
timeout 600 ncu --set full --clock-control none -k regex:'knn_stream' -c 3 -o gpurun_out/prof_k16 -f python scripts/bench_rows.py stress1 > gpurun_out/ncu_k16.log 2>&1; echo "ncu rc=$?"
python scripts/ncu_summary.py gpurun_out/prof_k16.ncu-rep; python scripts/ncu_stalls.py gpurun_out/prof_k16.ncu-rep
rm -f gpurun_out/prof_k16.ncu-rep
du -sh gpurun_out
