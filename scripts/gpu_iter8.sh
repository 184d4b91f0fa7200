#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "knn or encoder or simclr or smoke" > gpurun_out/pytest_knn.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_knn.log; grep -E "^FAILED|^ERROR" gpurun_out/pytest_knn.log | head
for mode in queue vote; do
  if [ $mode = vote ]; then export GRAFP_KNN_EPI=vote; fi
  timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$mode.log 2>&1
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$mode.log').read().strip().splitlines()[-1])
print('$mode', round(d['ms_per_step'],2), 'ms/step; knn_fwd', round(d['kernels']['knn_fwd']['ms_total']/d['steps'],3), 'ms/step')
PY
done
unset GRAFP_KNN_EPI
ALL_VARIANTS=0 timeout 300 python scripts/bench_ops.py 2>&1 | grep "N=" | cut -c1-60
