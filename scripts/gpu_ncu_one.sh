#!/bin/bash
# ncu --set full of one kernel family from scripts/ncu_ops.py: bash scripts/gpu_ncu_one.sh <regex> [count]
set +e
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$1" -c ${2:-4} \
  -o gpurun_out/prof_one -f python scripts/ncu_ops.py 512 1 > gpurun_out/ncu_one.log 2>&1; echo "ncu rc=$?"
python scripts/ncu_summary.py gpurun_out/prof_one.ncu-rep
