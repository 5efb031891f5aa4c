#!/bin/bash
# Same-box A/B of two builds of the library (NFS_B200_LIB): chain kernels (ab_chain.py) and the cfg 3 step (ab_backward.py).
# usage: scripts/dev/ab_libs.sh base "" [rounds]     ("" = the production libnfs_b200.so)
D=$(cd "$(dirname "$0")/../.." && pwd)
L=$D/nerf-few-shot-limitations_b200/nfs_b200
R=${3:-2}
for r in $(seq $R); do
  for v in "$1" "$2"; do
    lib=$L/libnfs_b200${v:+_$v}.so
    echo "== round $r  ${v:-production}"
    NFS_B200_LIB=$lib python $D/scripts/dev/ab_chain.py 2>&1 | tail -1
    NFS_B200_LIB=$lib python $D/scripts/dev/ab_backward.py - 2>&1 | grep "round 1"
  done
done
