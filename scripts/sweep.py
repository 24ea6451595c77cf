"""sweep.py — time fa_fwd on a list of shapes, two ways per shape.  GPU box only.
  steady    launches held back to back for ~60 ms under one CUDA-event pair after a warm-up: the power-capped steady state
            (what bench.py's `sustained` leg and scripts/tile_sweep.py report); comparable across shapes
  isolated  one launch right after a 512 MB L2 flush, median of 10: the cold-cache burst figure.  NOT comparable across
            shapes within one run: it depends on how warm the GPU is when the shape's turn comes (the first shapes of a run read
            5-10 % high; round 1's "GQA is 10 % slower than MHA" and "d = 64 causal is 27 % below non-causal" were this effect —
            scripts/gqa_probe.py and scripts/flush_probe.py show neither in a same-process A/B)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "flash-attention-cuda-c_b200")]
import torch, fa_b200

SHAPES = [  # B, Hq, Hkv, N, d, causal, dtype
    (8, 32, 32, 8192, 128, True, "bf16"), (8, 32, 32, 8192, 128, False, "bf16"),
    (2, 32, 32, 16384, 128, True, "bf16"), (1, 32, 32, 32768, 128, True, "bf16"), (1, 16, 16, 65536, 128, True, "bf16"),
    (16, 32, 32, 4096, 128, True, "bf16"), (32, 32, 32, 2048, 128, True, "bf16"), (64, 32, 32, 1024, 128, True, "bf16"),
    (4, 64, 8, 8192, 128, True, "bf16"), (4, 64, 64, 8192, 128, True, "bf16"), (8, 32, 32, 8192, 128, True, "fp16"),
    (8, 32, 32, 8192, 64, True, "bf16"), (8, 32, 32, 8192, 64, False, "bf16"), (4, 12, 12, 1024, 64, False, "fp16"),
    # GQA 32/8 (Llama-3-8B's real head layout): CTA pairs cut by four heads of a kv group
    (8, 32, 8, 8192, 128, True, "bf16"), (16, 32, 8, 4096, 128, True, "bf16"), (32, 32, 8, 2048, 128, True, "bf16"), (64, 32, 8, 1024, 128, True, "bf16"),
]
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for (B, Hq, Hkv, N, d, causal, dt) in SHAPES:
    t = {"bf16": torch.bfloat16, "fp16": torch.float16}[dt]
    q = torch.randn(B, Hq, N, d, device="cuda").to(t); k = torch.randn(B, Hkv, N, d, device="cuda").to(t); v = torch.randn(B, Hkv, N, d, device="cuda").to(t)
    o = torch.empty_like(q)
    call = lambda: fa_b200.attention_forward(q, k, v, causal=causal, out=o)
    for _ in range(3): call()
    torch.cuda.synchronize()
    ms = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); call(); b.record(); torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    iso = sorted(ms)[len(ms) // 2]
    n = max(10, int(60.0 / iso))
    for _ in range(max(5, n // 4)): call()
    steady = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n): call()
        b.record(); torch.cuda.synchronize()
        steady.append(a.elapsed_time(b) / n)
    st = sum(steady) / len(steady)
    F = 4.0 * B * Hq * N * N * d * (0.5 if causal else 1.0)
    by = (2 * B * Hq * N * d + 2 * B * Hkv * N * d) * 2
    tile = fa_b200.choose_kernel(B, Hq, Hkv, N, N, d, fa_b200.FA_DTYPE_BF16, causal)
    print(json.dumps({"B": B, "Hq": Hq, "Hkv": Hkv, "N": N, "d": d, "causal": causal, "dtype": dt,
                      "steady_ms": round(st, 4), "steady_tflops": round(F / st / 1e9, 1), "steady_gbs": round(by / st / 1e6, 1),
                      "isolated_ms_median": round(iso, 4), "isolated_tflops": round(F / iso / 1e9, 1),
                      "variant": [tile["softmax_warps"], tile["emu_pairs_per_8"], tile["staged_epilogue"], tile["cta_group"]],
                      "heads_per_item": tile["heads_per_item"], "work_items": tile["work_items"]}), flush=True)
    del q, k, v, o
