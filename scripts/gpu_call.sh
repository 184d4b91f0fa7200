#!/bin/bash
# One GPU lease: tests, bench lines, per-kernel timings.  Usage: gpurun -- bash scripts/gpu_call.sh <tag> [steps...]
tag=${1:-r02}
shift
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $out/${tag}_gpu.txt 2>&1
for step in "$@"; do
  case $step in
    tests) timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x > $out/${tag}_pytest_gpu.log 2>&1; tail -5 $out/${tag}_pytest_gpu.log ;;
    tests_all) timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > $out/${tag}_pytest_gpu.log 2>&1; tail -15 $out/${tag}_pytest_gpu.log ;;
    bench) timeout 900 python bench.py --steps 10 --warmup 3 > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err; tail -c 600 $out/${tag}_bench_n1.err; cut -c1-400 $out/${tag}_bench_n1.json ;;
    bench_quick) timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager > $out/${tag}_bench_n1_quick.json 2> $out/${tag}_bench_n1_quick.err; tail -c 600 $out/${tag}_bench_n1_quick.err; cut -c1-400 $out/${tag}_bench_n1_quick.json ;;
    bench_g0) GRAFP_CONV_GEMM=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager > $out/${tag}_bench_n1_gemm0.json 2> $out/${tag}_bench_n1_gemm0.err; tail -c 300 $out/${tag}_bench_n1_gemm0.err; cut -c1-300 $out/${tag}_bench_n1_gemm0.json ;;
    bench_bf16_g0) GRAFP_CONV_GEMM=0 timeout 600 python bench.py --dtype bf16 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager > $out/${tag}_bench_n1_bf16_gemm0.json 2> $out/${tag}_bench_n1_bf16_gemm0.err; tail -c 300 $out/${tag}_bench_n1_bf16_gemm0.err; cut -c1-300 $out/${tag}_bench_n1_bf16_gemm0.json ;;
    bench_bf16) timeout 600 python bench.py --dtype bf16 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager > $out/${tag}_bench_n1_bf16.json 2> $out/${tag}_bench_n1_bf16.err; tail -c 600 $out/${tag}_bench_n1_bf16.err; cut -c1-400 $out/${tag}_bench_n1_bf16.json ;;
    bench_ref) timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $out/${tag}_bench_reference_arm.json 2> $out/${tag}_bench_reference_arm.err; cut -c1-300 $out/${tag}_bench_reference_arm.json ;;
    kernels) timeout 900 python scripts/bench_kernels.py k1 k23 k5 gemm > $out/${tag}_bench_kernels.log 2>&1; tail -60 $out/${tag}_bench_kernels.log ;;
    k1) timeout 600 python scripts/bench_kernels.py k1 > $out/${tag}_bench_k1.log 2>&1; cat $out/${tag}_bench_k1.log ;;
    k23) timeout 600 python scripts/bench_kernels.py k23 > $out/${tag}_bench_k23.log 2>&1; cat $out/${tag}_bench_k23.log ;;
    copies) timeout 600 python scripts/profile_copies.py 512 > $out/${tag}_copies.log 2>&1; cut -c1-250 $out/${tag}_copies.log | head -60 ;;
    k3_test) timeout 900 python -m pytest tests -m gpu -x -q -k "mr_aggregate or aggregation or bf16_hot" > $out/${tag}_pytest_k3.log 2>&1; tail -8 $out/${tag}_pytest_k3.log ;;
    gemm) timeout 600 python scripts/bench_kernels.py gemm > $out/${tag}_bench_gemm.log 2>&1; cat $out/${tag}_bench_gemm.log ;;
    gemm_test) timeout 900 python -m pytest tests -m gpu -x -q -k "conv1x1 or gemm or conv_batch_norm or downsample" > $out/${tag}_pytest_gemm.log 2>&1; tail -25 $out/${tag}_pytest_gemm.log ;;
    k5) timeout 600 python scripts/bench_kernels.py k5 > $out/${tag}_bench_k5.log 2>&1; cat $out/${tag}_bench_k5.log ;;
    prof_bf16) timeout 600 python scripts/profile_step.py 512 --bf16 > $out/${tag}_profile_step_bf16.log 2>&1; head -50 $out/${tag}_profile_step_bf16.log | cut -c1-200 ;;
    prof) timeout 600 python scripts/profile_step.py 512 > $out/${tag}_profile_step.log 2>&1; head -50 $out/${tag}_profile_step.log | cut -c1-200 ;;
    ncu_ops) timeout 900 ncu --set full --clock-control none --import-source on -k regex:"knn_|mr_aggregate|bn_" -c 60 -f -o $out/${tag}_ncu_ops python scripts/ncu_ops.py 512 1 > $out/${tag}_ncu_ops.log 2>&1; tail -3 $out/${tag}_ncu_ops.log
             python scripts/ncu_summary.py $out/${tag}_ncu_ops.ncu-rep > $out/${tag}_ncu_ops_summary.txt 2>&1; python scripts/ncu_stalls.py $out/${tag}_ncu_ops.ncu-rep > $out/${tag}_ncu_ops_stalls.txt 2>&1; cat $out/${tag}_ncu_ops_summary.txt ;;
    bench_rev) GRAFP_BN_REVERSE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager > $out/${tag}_bench_n1_bnrev.json 2> $out/${tag}_bench_n1_bnrev.err; tail -c 300 $out/${tag}_bench_n1_bnrev.err; cut -c1-300 $out/${tag}_bench_n1_bnrev.json ;;
    bench_l2) GRAFP_BN_L2_KEEP_MB=48 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager > $out/${tag}_bench_n1_l2keep48.json 2> $out/${tag}_bench_n1_l2keep48.err; tail -c 300 $out/${tag}_bench_n1_l2keep48.err; cut -c1-300 $out/${tag}_bench_n1_l2keep48.json
              GRAFP_BN_L2_KEEP_MB=80 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager > $out/${tag}_bench_n1_l2keep80.json 2> $out/${tag}_bench_n1_l2keep80.err; cut -c1-300 $out/${tag}_bench_n1_l2keep80.json ;;
    bench_nograph) timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager --graph off > $out/${tag}_bench_n1_nograph.json 2> $out/${tag}_bench_n1_nograph.err; cut -c1-300 $out/${tag}_bench_n1_nograph.json ;;
    launches) timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $out/launches.csv python bench.py --steps 2 --warmup 1 --graph off --cudnn-benchmark off --no-cpu-baseline --no-gpu-eager > $out/${tag}_bench_under_ncu.log 2>&1; wc -l $out/launches.csv ;;
    ncu_full) timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"knn_|mr_aggregate|bn_|ntxent|peak_extract|conv1x1|taps_" -c 130 -f -o /tmp/prof_ops python scripts/ncu_ops.py 512 1 > $out/${tag}_ncu_ops.log 2>&1; tail -2 $out/${tag}_ncu_ops.log
             python scripts/ncu_summary.py /tmp/prof_ops.ncu-rep > $out/${tag}_ncu_hot_kernels_summary.txt 2>&1
             python scripts/ncu_stalls.py /tmp/prof_ops.ncu-rep > $out/${tag}_ncu_stalls.txt 2>&1
             python scripts/dram_traffic.py /tmp/prof_ops.ncu-rep $out/${tag}_dram_traffic.json > /dev/null 2>&1
             cp /tmp/prof_ops.ncu-rep $out/prof_ops.ncu-rep; python scripts/summarize_profiles.py ${tag} ncu_only $out > /dev/null 2>&1; rm -f $out/prof_ops.ncu-rep
             cat $out/${tag}_ncu_hot_kernels_summary.txt ;;
    bench_n2) NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL NCCL_DEBUG_FILE=$out/${tag}_nccl_n2_%p.log timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $out/${tag}_bench_n2.json 2> $out/${tag}_bench_n2.err; for f in $out/${tag}_nccl_n2_*.log; do head -c 20000 $f > $f.head; rm -f $f; done; grep -c "Grad strides" $out/${tag}_bench_n2.err; tail -c 400 $out/${tag}_bench_n2.err; cut -c1-300 $out/${tag}_bench_n2.json ;;
    ncu_bf16) timeout 900 ncu --set full --clock-control none -k regex:"mr_aggregate|bn_|conv1x1|knn_" -c 70 -f -o /tmp/prof_bf16 python scripts/ncu_ops.py 512 1 bf16 > $out/${tag}_ncu_ops_bf16.log 2>&1; tail -2 $out/${tag}_ncu_ops_bf16.log
             python scripts/ncu_summary.py /tmp/prof_bf16.ncu-rep > $out/${tag}_ncu_bf16_summary.txt 2>&1; python scripts/ncu_stalls.py /tmp/prof_bf16.ncu-rep > $out/${tag}_ncu_bf16_stalls.txt 2>&1; cat $out/${tag}_ncu_bf16_summary.txt; cat $out/${tag}_ncu_bf16_stalls.txt | cut -c1-220 ;;
    bench_n2_graph) timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --graph on > $out/${tag}_bench_n2_graph.json 2> $out/${tag}_bench_n2_graph.err; tail -c 1500 $out/${tag}_bench_n2_graph.err; cut -c1-300 $out/${tag}_bench_n2_graph.json ;;
    bench_graph) timeout 600 python bench.py --steps 10 --warmup 3 --graph on --no-cpu-baseline --no-gpu-eager > $out/${tag}_bench_n1_graph.json 2> $out/${tag}_bench_n1_graph.err; tail -c 300 $out/${tag}_bench_n1_graph.err; cut -c1-300 $out/${tag}_bench_n1_graph.json ;;
    bench_n4_graph) timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 4 --steps 8 --warmup 3 > $out/${tag}_bench_n4_graph.json 2> $out/${tag}_bench_n4_graph.err; tail -c 400 $out/${tag}_bench_n4_graph.err; cut -c1-300 $out/${tag}_bench_n4_graph.json ;;
    bench_n8_graph) timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --steps 8 --warmup 3 --graph on > $out/${tag}_bench_n8_graph.json 2> $out/${tag}_bench_n8_graph.err; tail -c 600 $out/${tag}_bench_n8_graph.err; cut -c1-300 $out/${tag}_bench_n8_graph.json ;;
    bench_n8) timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 > $out/${tag}_bench_n8.json 2> $out/${tag}_bench_n8.err; grep -c "Grad strides" $out/${tag}_bench_n8.err; tail -c 400 $out/${tag}_bench_n8.err; cut -c1-300 $out/${tag}_bench_n8.json ;;
    smoke) timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ;;
    *) echo "unknown step $step" ;;
  esac
done
