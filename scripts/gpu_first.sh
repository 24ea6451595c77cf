#!/bin/bash
# First GPU contact: stage probes -> compat entry point -> parity tests -> quick timing.  Everything under a
# timeout so a protocol bug cannot hang the box; logs land in gpurun_out/.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
PK=flash-attention-cuda-c_b200
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
echo "== probe"; timeout 120 $PK/tests/probe_umma ${PROBE_SCAN:-0} > gpurun_out/probe.log 2>&1; echo "probe rc=$?"; tail -5 gpurun_out/probe.log
echo "== compat"; timeout 120 $PK/tests/compat_main > gpurun_out/compat.log 2>&1; echo "compat rc=$?"; tail -8 gpurun_out/compat.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest.log
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err
