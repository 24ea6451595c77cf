#!/bin/bash
# ncu --set full capture (one launch each) of several builds on one shape:  gpu_ncu_ab.sh "B H N d causal" name1 name2 ...
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
SHAPE=$1; shift
for v in "$@"; do
  timeout 120 python scripts/one_launch.py variants/libfa_v_$v.so $SHAPE > gpurun_out/plain_$v.log 2>&1 || { echo "plain run of $v failed"; tail -3 gpurun_out/plain_$v.log; continue; }
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:fwdSm100Kernel -s 1 -c 1 -f -o gpurun_out/ncu_$v python scripts/one_launch.py variants/libfa_v_$v.so $SHAPE > gpurun_out/ncu_$v.log 2>&1
  echo "$v ncu rc=$?"
done
