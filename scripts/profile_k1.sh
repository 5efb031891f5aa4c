#!/bin/bash
# ncu evidence for K1 (run under gpurun, one GPU): launch list + full capture of fwd and bwd.
set -x
mkdir -p gpurun_out
CMD="python bench.py --quick --steps 3 --warmup 3"
$CMD > gpurun_out/plain_k1.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_k1.csv $CMD > gpurun_out/ncu_k1_list.log 2>&1
$CMD > gpurun_out/plain_k1b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:composite -s 6 -c 2 -f -o gpurun_out/prof_k1 $CMD > gpurun_out/ncu_k1_full.log 2>&1
ls -la gpurun_out
