#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -k "not config4" > gpurun_out/pytest_r2_11.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_r2_11.log
echo "== trace"; timeout 120 python scripts/trace_cta.py 4 12 1024 64 0 fp16 2>&1 | head -14
echo "== cfg2"; timeout 300 python scripts/cfg2_probe.py 2>&1 | grep "shape" | grep -v "half_items\": 2" | grep "\"softmax_warps\": 8, \"emu\": 0, \"staged_epilogue\": [01]" | tee gpurun_out/r2_cfg2_probe.jsonl
echo "== cycles"; FA_B200_ALLOW_OLDER_LIB=1 FA_AB_SHAPES=0,2,4 FA_CYC_REPS=3 timeout 900 python scripts/cycles.py r1 shipped@8,0,0 2>&1 | grep -v "pass\": 0" | tee gpurun_out/cyc_r2_11.log
