"""sweep.py — time fa_fwd on a list of shapes (CUDA events, L2 flushed for small working sets). GPU box only."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "flash-attention-cuda-c_b200")]
import torch, fa_b200

SHAPES = [  # B, Hq, Hkv, N, d, causal, dtype
    (8, 32, 32, 8192, 128, True, "bf16"), (8, 32, 32, 8192, 128, False, "bf16"),
    (2, 32, 32, 16384, 128, True, "bf16"), (1, 32, 32, 32768, 128, True, "bf16"), (1, 16, 16, 65536, 128, True, "bf16"),
    (16, 32, 32, 4096, 128, True, "bf16"), (32, 32, 32, 2048, 128, True, "bf16"), (64, 32, 32, 1024, 128, True, "bf16"),
    (4, 64, 8, 8192, 128, True, "bf16"), (8, 32, 32, 8192, 128, True, "fp16"),
    (8, 32, 32, 8192, 64, True, "bf16"), (8, 32, 32, 8192, 64, False, "bf16"), (4, 12, 12, 1024, 64, False, "fp16"),
]
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for (B, Hq, Hkv, N, d, causal, dt) in SHAPES:
    t = {"bf16": torch.bfloat16, "fp16": torch.float16}[dt]
    q = torch.randn(B, Hq, N, d, device="cuda").to(t); k = torch.randn(B, Hkv, N, d, device="cuda").to(t); v = torch.randn(B, Hkv, N, d, device="cuda").to(t)
    o = torch.empty_like(q)
    for _ in range(3): fa_b200.attention_forward(q, k, v, causal=causal, out=o)
    torch.cuda.synchronize()
    ms = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fa_b200.attention_forward(q, k, v, causal=causal, out=o); b.record(); torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    F = 4.0 * B * Hq * N * N * d * (0.5 if causal else 1.0)
    m = sorted(ms)[len(ms) // 2]
    by = (2 * B * Hq * N * d + 2 * B * Hkv * N * d) * 2
    print(json.dumps({"B": B, "Hq": Hq, "Hkv": Hkv, "N": N, "d": d, "causal": causal, "dtype": dt, "ms_median": round(m, 4),
                      "ms_min": round(min(ms), 4), "tflops": round(F / m / 1e9, 1), "gbs": round(by / m / 1e6, 1)}), flush=True)
    del q, k, v, o
