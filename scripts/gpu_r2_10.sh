#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_r2_10.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_r2_10.log
echo "== pytest forced 8,0,1"; FA_FORCE_VARIANT=8,0,1 timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_parity_large_gpu.py -m gpu -q --timeout 600 -k "not config4 and not thousands and not long_sequence" 2>&1 | tail -2
echo "== cfg2"; timeout 300 python scripts/cfg2_probe.py 2>&1 | grep "shape" | tee gpurun_out/r2_cfg2_probe.jsonl
