#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "gather or edge or maxk or golden or bf16" > gpurun_out/pytest_edge.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_edge.log; grep -E "^FAILED|^ERROR" gpurun_out/pytest_edge.log | head
timeout 300 python scripts/bench_rows.py edge > gpurun_out/bench_rows_edge.log 2>&1; cat gpurun_out/bench_rows_edge.log
echo "== one-pass K3e"; GRAFP_EDGE_BWD_ROW=1 timeout 300 python scripts/bench_rows.py edge 2>&1 | grep "N=" | cut -c1-110
echo "== old forms"; GRAFP_EDGE_ROW_FORM=0 GRAFP_MAXK_ROW_FORM=0 timeout 300 python scripts/bench_rows.py edge 2>&1 | grep "N=" | cut -c1-160
