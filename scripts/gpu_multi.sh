#!/bin/bash
# 2+ GPU checks: NCCL ring-KV parity, sharded bench lines, C++ host driver, reference kernel beside ours
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
G=${G:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29517"
echo "== ref kernel"; timeout 120 python scripts/ref_kernel_bench.py > gpurun_out/ref_kernel.json 2> gpurun_out/ref_kernel.err; echo "rc=$?"; cat gpurun_out/ref_kernel.json; tail -3 gpurun_out/ref_kernel.err
echo "== dropin"; timeout 60 oracle/_ref/ref_test_dropin > gpurun_out/dropin.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/dropin.log
echo "== ring check"; timeout 600 $TR scripts/ring_check.py > gpurun_out/ring_check.log 2>&1; echo "rc=$?"; grep -E "OK|FAIL|PASSED|Error" gpurun_out/ring_check.log | tail -12
echo "== fa_main"; timeout 300 flash-attention-cuda-c_b200/fa_main --props --gpus $G > gpurun_out/fa_main.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/fa_main.log
for W in cfg3 cfg5 cfg4; do
  echo "== bench $W x$G"; timeout 900 $TR bench.py --gpus $G --steps 5 --warmup 3 --workload $W > gpurun_out/bench_${W}_g$G.json 2> gpurun_out/bench_${W}_g$G.err; echo "rc=$?"; cut -c1-700 gpurun_out/bench_${W}_g$G.json; tail -2 gpurun_out/bench_${W}_g$G.err
done
echo "== bench cfg5 x1"; timeout 600 python bench.py --steps 3 --warmup 3 --workload cfg5 --no-cpu > gpurun_out/bench_cfg5_g1.json 2> gpurun_out/bench_cfg5_g1.err; echo "rc=$?"; cut -c1-500 gpurun_out/bench_cfg5_g1.json
