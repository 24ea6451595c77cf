#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
echo "== pytest forced 8,0,1"; FA_FORCE_VARIANT=8,0,1 timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_parity_large_gpu.py -m gpu -q --timeout 600 -k "not config4 and not thousands" 2>&1 | tail -2
echo "== trace ST"; FA_FORCE_VARIANT=8,0,1 timeout 120 python scripts/trace_cta.py 64 32 1024 128 1 bf16 2>&1 | grep -n "epilogue\|last PV\|slice" | head -14
echo "== cycles"; FA_AB_SHAPES=0,1,2,4,3,5 FA_CYC_REPS=3 timeout 900 python scripts/cycles.py shipped@8,0,0 shipped@8,0,1 2>&1 | grep -v "pass\": 0" | cut -c1-200 | tee gpurun_out/cyc_r2_13.log
echo "== sustained"; FA_AB_SHAPES=0,2,4 FA_SUS_ROUNDS=3 timeout 900 python scripts/ab_sustained.py shipped@8,0,0 shipped@8,0,1 2>&1 | cut -c1-200 | tee gpurun_out/sus_r2_13.log
