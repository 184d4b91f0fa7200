#!/bin/bash
set +e
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8.log 2>&1; echo "bench n8 rc=$?"; tail -1 gpurun_out/bench_n8.log
