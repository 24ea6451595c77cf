#!/bin/bash
# round 2, GPU call 2: parity of every compiled kernel variant, softmax-stage probe, tile sweep, regression A/B against the
# round-1 library, half-item tail schedule on the GPT-2 shape, 13-shape sweep, bench line
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
echo "== pytest (tile table)"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_table.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest_table.log
for fv in 8,0 16,0 16,1; do
  echo "== pytest forced variant $fv"; FA_FORCE_VARIANT=$fv timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_parity_large_gpu.py -m gpu -q --timeout 600 -k "not config4 and not long_sequence and not thousands" > gpurun_out/pytest_v${fv/,/_}.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_v${fv/,/_}.log
done
echo "== tile sweep"; timeout 900 python scripts/tile_sweep.py > gpurun_out/r2_tile_sweep.jsonl 2> gpurun_out/tile_sweep.err; echo "rc=$?"; grep winner gpurun_out/r2_tile_sweep.jsonl
echo "== cycles: round-1 library vs now"; FA_AB_SHAPES=0,1,2,3 FA_CYC_REPS=3 timeout 600 python scripts/cycles.py r1 shipped@8,0 2>&1 | grep -v "pass\": 0" | tee gpurun_out/cyc_r2_2.log
echo "== sustained: round-1 library vs now"; FA_AB_SHAPES=0 FA_SUS_ROUNDS=3 timeout 600 python scripts/ab_sustained.py r1 shipped@8,0 2>&1 | tee gpurun_out/sus_r2_2.log
echo "== cfg2 probe"; timeout 600 python scripts/cfg2_probe.py 2>&1 | tee gpurun_out/cfg2_probe.log
echo "== sweep"; timeout 600 python scripts/sweep.py > gpurun_out/r2_sweep.jsonl 2>&1; echo "rc=$?"; cat gpurun_out/r2_sweep.jsonl
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_2.json 2> gpurun_out/bench_r2_2.err; echo "rc=$?"; cut -c1-600 gpurun_out/bench_r2_2.json
echo "== fa_main props"; timeout 120 ./flash-attention-cuda-c_b200/fa_main --props --N 2048 --B 2 --H 8 2>&1 | tail -14
