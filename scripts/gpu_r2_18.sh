#!/bin/bash
# final multi-GPU pass of round 2: ring-KV parity over NCCL + the driver's bench line (cfg3 sharded + ring sub-record)
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
G=${G:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29517"
echo "== ring check"; timeout 600 $TR scripts/ring_check.py > gpurun_out/ring_check_g$G.log 2>&1; echo "rc=$?"; grep -E "OK|FAIL|PASSED|Error" gpurun_out/ring_check_g$G.log | tail -12
echo "== bench x$G"; timeout 900 $TR bench.py --gpus $G --steps 10 --warmup 3 > gpurun_out/bench_final_g$G.json 2> gpurun_out/bench_final_g$G.err; echo "rc=$?"; tail -2 gpurun_out/bench_final_g$G.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_final_g$G.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d["e2e"].get("value"))
print("ring", json.dumps(d.get("ring"))[:1500])
PY
