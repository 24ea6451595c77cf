#!/bin/bash
# round 2, GPU call 12: final build — whole suite, tile sweep, sweep, cfg2 probe, bench
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_r2_12.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_r2_12.log
for fv in 8,0,0 8,0,1 16,1,0; do
  echo "== pytest forced variant $fv"; FA_FORCE_VARIANT=$fv timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_parity_large_gpu.py -m gpu -q --timeout 600 -k "not config4 and not thousands and not long_sequence" 2>&1 | tail -1
done
echo "== tile sweep"; timeout 900 python scripts/tile_sweep.py > gpurun_out/r2_tile_sweep.jsonl 2> gpurun_out/tile_sweep.err; echo "rc=$?"; grep winner gpurun_out/r2_tile_sweep.jsonl | cut -c1-150
echo "== sweep"; timeout 900 python scripts/sweep.py > gpurun_out/r2_sweep.jsonl 2>gpurun_out/sweep.err; echo "rc=$?"; cut -c1-260 gpurun_out/r2_sweep.jsonl
echo "== cfg2"; timeout 300 python scripts/cfg2_probe.py 2>&1 | grep "shape" | grep "half_items\": 1" | tee gpurun_out/r2_cfg2_probe.jsonl | cut -c1-220
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_12.json 2> gpurun_out/bench_r2_12.err; echo "rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_12.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print('roofline',d['roofline']['frac']); print('sustained',d['sustained']['value']); print('e2e',d['e2e']['value']); print('small',d['small_shapes']['cfg2_fp16_B4_H12_N1024_d64'])
PY
