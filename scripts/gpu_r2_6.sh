#!/bin/bash
# round 2, GPU call 6: whole suite on the table-driven build, 13-shape sweep, bench line (1 GPU)
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_r2_6.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_r2_6.log
echo "== sweep"; timeout 600 python scripts/sweep.py > gpurun_out/r2_sweep.jsonl 2>&1; echo "rc=$?"; cat gpurun_out/r2_sweep.jsonl | cut -c1-200
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_6.json 2> gpurun_out/bench_r2_6.err; echo "rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_6.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print('roofline',d['roofline']['frac']); print('sustained',d['sustained']['value'],d['sustained']['clocks']); print('e2e',d['e2e']['value'],d['e2e']['pcie']); print('small',d['small_shapes'])
PY
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r2_6.json 2> gpurun_out/bench_ref_r2_6.err; echo "rc=$?"; cut -c1-1500 gpurun_out/bench_ref_r2_6.json
