#!/bin/bash
# round 2, GPU call 5: where did 2 % go since round 1?  elimination builds + small experiments, all with the (8,0,0) variant
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
export FA_B200_ALLOW_OLDER_LIB=1
echo "== cycles"; FA_AB_SHAPES=0,1,2 FA_CYC_REPS=3 timeout 900 python scripts/cycles.py r1 shipped@8,0,0 xa@8,0,0 xc@8,0,0 xac@8,0,0 xnh@8,0,0 xw100@8,0,0 xw1000@8,0,0 2>&1 | grep -v "pass\": 0" | tee gpurun_out/cyc_r2_5.log
echo "== sustained"; FA_AB_SHAPES=0 FA_SUS_ROUNDS=3 timeout 900 python scripts/ab_sustained.py r1 shipped@8,0,0 xac@8,0,0 xnh@8,0,0 xw100@8,0,0 xw1000@8,0,0 2>&1 | tee gpurun_out/sus_r2_5.log
