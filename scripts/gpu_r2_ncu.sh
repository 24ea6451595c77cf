#!/bin/bash
# round 2: ncu launch list + one full capture of the headline kernel (each only after the same command exited 0 without ncu)
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --no-extras"
echo "== ncu launches"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"; tail -4 gpurun_out/r2_launches.csv | cut -c1-300
echo "== ncu full"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fwdSm100 -s 3 -c 2 -f -o gpurun_out/r2_prof $CMD > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"; tail -3 gpurun_out/ncu_full.log
