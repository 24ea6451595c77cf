#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_r2_8.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_r2_8.log
echo "== sweep"; timeout 900 python scripts/sweep.py > gpurun_out/r2_sweep.jsonl 2>gpurun_out/sweep.err; echo "rc=$?"; cut -c1-330 gpurun_out/r2_sweep.jsonl
echo "== flush probe"; timeout 300 python scripts/flush_probe.py 2>&1 | grep shape | tee gpurun_out/r2_flush_probe.jsonl
echo "== cfg2"; timeout 300 python scripts/cfg2_probe.py 2>&1 | grep "N1024_d64" | tee gpurun_out/r2_cfg2_probe.jsonl
