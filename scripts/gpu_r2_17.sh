#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
echo "== pytest forced pair 8,0,1,2"; FA_FORCE_VARIANT=8,0,1,2 timeout 1200 python -m pytest tests/test_parity_gpu.py tests/test_parity_large_gpu.py -m gpu -q --timeout 900 2>&1 | tail -4
echo "== pytest forced pair direct epilogue 8,0,0,2"; FA_FORCE_VARIANT=8,0,0,2 timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_parity_large_gpu.py -m gpu -q --timeout 600 -k "not config4 and not thousands" 2>&1 | tail -3
echo "== tile sweep"; FA_TS_MS=200 FA_TS_ROUNDS=4 timeout 1300 python scripts/tile_sweep.py > gpurun_out/tile_sweep_r2d.jsonl 2> gpurun_out/tile_sweep_r2d.err; grep -c winner gpurun_out/tile_sweep_r2d.jsonl
