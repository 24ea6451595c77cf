"""flush_probe.py — why do single launches after an L2 flush time differently from launches held back to back?  Times one
shape three ways: (a) each launch right after a 512 MB flush write (what sweep.py does), (b) flush, then 2 ms of idle, then
the launch, (c) 20 launches back to back.  GPU box only."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "flash-attention-cuda-c_b200")]
import torch, fa_b200
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for (B, H, N, d, causal) in ((8, 32, 8192, 64, True), (8, 32, 8192, 64, False), (8, 32, 8192, 128, True)):
    q, k, v = (torch.randn(B, H, N, d, device="cuda").to(torch.bfloat16) for _ in range(3))
    o = torch.empty_like(q)
    F = 4.0 * B * H * N * N * d * (0.5 if causal else 1.0)
    for _ in range(3): fa_b200.attention_forward(q, k, v, causal=causal, out=o)
    torch.cuda.synchronize()
    res = {}
    for mode in ("flush", "flush+idle", "noflush_single", "back_to_back"):
        ms = []
        for _ in range(8):
            if mode.startswith("flush"): flush.zero_()
            if mode == "flush+idle": torch.cuda.synchronize(); time.sleep(0.002)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            n = 20 if mode == "back_to_back" else 1
            for _ in range(n): fa_b200.attention_forward(q, k, v, causal=causal, out=o)
            b.record(); torch.cuda.synchronize()
            ms.append(a.elapsed_time(b) / n)
        ms.sort(); res[mode] = round(F / ms[len(ms) // 2] / 1e9, 1)
    print(json.dumps({"shape": f"N{N}_d{d}_{'c' if causal else 'nc'}", "tflops": res}), flush=True)
