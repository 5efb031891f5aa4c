#!/bin/bash
# Round-2 ncu evidence, second half (run under gpurun, one GPU), with the merged backward kernel as the default route:
# (1) every launch of the headline step with its device time and DRAM bytes; (2) full captures of the merged backward
# kernel and of the forward chain of a training step, from one eager step (scripts/step_once.py).
set -x
mkdir -p gpurun_out
STEP="python bench.py --quick --no-extras --no-cpu-baseline --steps 2 --warmup 3"
$STEP > gpurun_out/r2g_plain_step.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 330 --csv \
    --log-file gpurun_out/r2g_launches_step.csv $STEP > gpurun_out/r2g_ncu_step.log 2>&1
ONE="python scripts/step_once.py 2"
$ONE > gpurun_out/r2g_plain_once.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"backward_fused|fused_mlp" -s 3 -c 3 -f \
    -o gpurun_out/r2g_prof_step $ONE > gpurun_out/r2g_ncu_full.log 2>&1
ls -la gpurun_out | grep r2g_
