#!/bin/bash
# ncu evidence for K1 + K3 (run under gpurun, one GPU): launch lists + full captures.
set -x
mkdir -p gpurun_out
K1="python bench.py --quick --no-extras --steps 3 --warmup 3"
$K1 > gpurun_out/plain_k1.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_k1.csv $K1 > gpurun_out/ncu_k1_list.log 2>&1
$K1 > gpurun_out/plain_k1b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:composite -s 6 -c 2 -f -o gpurun_out/prof_k1 $K1 > gpurun_out/ncu_k1_full.log 2>&1
K3="python scripts/mlp_once.py"
$K3 > gpurun_out/plain_k3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_k3.csv $K3 > gpurun_out/ncu_k3_list.log 2>&1
$K3 > gpurun_out/plain_k3b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"fused_mlp|wgrad|posenc_bf16" -s 6 -c 14 -f -o gpurun_out/prof_k3 $K3 > gpurun_out/ncu_k3_full.log 2>&1
ls -la gpurun_out | head -40
