#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
echo "== pytest (table build)"; timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -3
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench_r2_16.json 2> gpurun_out/bench_r2_16.err; tail -c 3000 gpurun_out/bench_r2_16.json
