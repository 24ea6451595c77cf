#!/bin/bash
# round 2, GPU call 1: parity of the 16-softmax-warp layout (shipped build) and of the 8-warp layout (variant w8) on the
# whole suite, then cycle-exact and sustained A/B of the layouts and of the FMA-pipe exp2 share
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,power.limit,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== pytest shipped (w16)"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_w16.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_w16.log
echo "== pytest w8"; FA_B200_LIB=$PWD/variants/libfa_v_w8.so timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_w8.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_w8.log
echo "== cycles"; FA_AB_SHAPES=0,1,2,3,5 FA_CYC_REPS=3 timeout 600 python scripts/cycles.py w8 w16 w16e1 w16e2 w16e3 2>&1 | grep -v "pass\": 0" | tee gpurun_out/cyc_r2_1.log
echo "== sustained"; FA_AB_SHAPES=0,3 FA_SUS_ROUNDS=3 timeout 600 python scripts/ab_sustained.py w8 w16 w16e1 w16e2 2>&1 | tee gpurun_out/sus_r2_1.log
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_1.json 2> gpurun_out/bench_r2_1.err; echo "rc=$?"; cut -c1-1500 gpurun_out/bench_r2_1.json
