#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
echo "== pytest quick"; timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q --timeout 600 -x 2>&1 | tail -2
echo "== cycles"; FA_AB_SHAPES=0,2,4,5,3 FA_CYC_REPS=3 timeout 900 python scripts/cycles.py noqpf@8,0,0 shipped@8,0,0 2>&1 | grep -v "pass\": 0" | tee gpurun_out/cyc_r2_7.log
echo "== sustained"; FA_AB_SHAPES=2,4,5 FA_SUS_ROUNDS=3 timeout 900 python scripts/ab_sustained.py noqpf shipped 2>&1 | tee gpurun_out/sus_r2_7.log
