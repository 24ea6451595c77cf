#!/bin/bash
# tests + sweep + ncu launch list + ncu full capture of the main kernel (each ncu only after a clean plain run)
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest.log
echo "== sweep"; timeout 600 python scripts/sweep.py > gpurun_out/sweep.log 2>&1; echo "sweep rc=$?"; cat gpurun_out/sweep.log | tail -20
if [ "${NCU:-1}" = "1" ]; then
echo "== ncu launches"
python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"; tail -3 gpurun_out/launches.csv | cut -c1-300
echo "== ncu full"
python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fwdSm100Kernel -s 3 -c 2 -f -o gpurun_out/prof python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"; tail -3 gpurun_out/ncu_full.log
fi
