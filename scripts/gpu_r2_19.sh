#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
echo "== pytest (table build)"; timeout 1200 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -3
echo "== pytest forced pair, direct epilogue (carry / ring steps by heads where the group is even)"; FA_FORCE_VARIANT=8,0,0,2 timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_parity_large_gpu.py -m gpu -q --timeout 600 -k "not config4 and not thousands" 2>&1 | tail -3
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench_r2_19.json 2> gpurun_out/bench_r2_19.err; python - <<PY
import json
d = json.loads(open("gpurun_out/bench_r2_19.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "sustained", d["sustained"]["value"], "e2e", d["e2e"].get("value"), d["roofline"]["kernel"])
PY
