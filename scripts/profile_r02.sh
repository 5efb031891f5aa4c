#!/bin/bash
# Round-2 ncu evidence (run under gpurun, one GPU): (1) every launch of the headline step with its device time and DRAM
# bytes (the step's traffic, roofline.traffic); (2) full captures of the chain kernel's three instantiations
# (incl. the inference one) and the weight-gradient kernel.  ncu times are cold-cache and serialised: compare shares.
set -x
mkdir -p gpurun_out
STEP="python bench.py --quick --no-extras --no-cpu-baseline --steps 2 --warmup 3"
$STEP > gpurun_out/r2_plain_step.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 330 --csv \
    --log-file gpurun_out/r2_launches_step.csv $STEP > gpurun_out/r2_ncu_step.log 2>&1
K3="python scripts/mlp_once.py"
$K3 > gpurun_out/r2_plain_k3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"fused_mlp|wgrad" -s 2 -c 12 -f -o gpurun_out/r2_prof_k3 $K3 \
    > gpurun_out/r2_ncu_k3_full.log 2>&1
ls -la gpurun_out | grep r2_ | head -40
