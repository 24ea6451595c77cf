#!/bin/bash
# round 2, GPU call 4: the staged TMA-store epilogue variant (8,0,1): parity, cycles / sustained A/B against (8,0,0), tile sweep
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
for fv in 8,0,1; do
  echo "== pytest forced variant $fv"; FA_FORCE_VARIANT=$fv timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_parity_large_gpu.py -m gpu -q --timeout 600 -k "not config4 and not thousands" > gpurun_out/pytest_v${fv//,/_}.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_v${fv//,/_}.log
done
echo "== cycles"; FA_AB_SHAPES=0,1,2,4,3,5 FA_CYC_REPS=3 timeout 600 python scripts/cycles.py shipped@8,0,0 shipped@8,0,1 2>&1 | grep -v "pass\": 0" | tee gpurun_out/cyc_r2_4.log
echo "== sustained"; FA_AB_SHAPES=0,2,4 FA_SUS_ROUNDS=3 timeout 600 python scripts/ab_sustained.py shipped@8,0,0 shipped@8,0,1 2>&1 | tee gpurun_out/sus_r2_4.log
echo "== tile sweep"; timeout 900 python scripts/tile_sweep.py > gpurun_out/r2_tile_sweep.jsonl 2> gpurun_out/tile_sweep.err; echo "rc=$?"; grep winner gpurun_out/r2_tile_sweep.jsonl
echo "== cfg2 probe"; timeout 600 python scripts/cfg2_probe.py 2>&1 | grep -v "half_items\": 0" | tee gpurun_out/cfg2_probe.log
