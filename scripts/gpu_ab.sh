#!/bin/bash
# A/B a set of library builds (variants/libfa_v_*.so) after the parity tests of the shipped build:  gpu_ab.sh [nshapes] name1 name2 ...
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
NS=${1:-2}; shift
echo "== pytest"; timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest.log
for rep in 1 2; do for v in "$@"; do timeout 200 python scripts/time_lib.py variants/libfa_v_$v.so $v $NS 2>&1 | tail -1; done; done | tee gpurun_out/ab.log
if [ -f flash-attention-cuda-c_b200/libfa_b200_prof.so ]; then
  echo "== phase profile"; timeout 200 python scripts/phase_profile.py 8 32 8192 128 0 2>&1 | tail -1 | tee gpurun_out/phase.log
fi
