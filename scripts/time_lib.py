"""time_lib.py <lib.so> [label] — time a given build of the library on the headline shapes (CUDA events, median of 10)
and check one result against the production build for sanity."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "flash-attention-cuda-c_b200")]
import torch, fa_b200
fa_b200.LIB_PATH = os.path.abspath(sys.argv[1]); label = sys.argv[2] if len(sys.argv) > 2 else os.path.basename(sys.argv[1])
SHAPES = [(8, 32, 32, 8192, 128, True, "bf16"), (8, 32, 32, 8192, 128, False, "bf16"), (32, 32, 32, 2048, 128, True, "bf16"),
          (8, 32, 32, 8192, 64, True, "bf16"), (8, 32, 32, 8192, 128, True, "fp16")]
if len(sys.argv) > 3: SHAPES = SHAPES[:int(sys.argv[3])]
REPS = int(os.environ.get("FA_TIME_REPS", "10"))
res = {}
for (B, Hq, Hkv, N, d, causal, dt) in SHAPES:
    t = {"bf16": torch.bfloat16, "fp16": torch.float16}[dt]
    g = torch.Generator(device="cuda").manual_seed(0)
    q = torch.randn(B, Hq, N, d, device="cuda", generator=g).to(t); k = torch.randn(B, Hkv, N, d, device="cuda", generator=g).to(t); v = torch.randn(B, Hkv, N, d, device="cuda", generator=g).to(t)
    o = torch.empty_like(q)
    for _ in range(3): fa_b200.attention_forward(q, k, v, causal=causal, out=o)
    torch.cuda.synchronize(); ms = []
    for _ in range(REPS):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fa_b200.attention_forward(q, k, v, causal=causal, out=o); b.record(); torch.cuda.synchronize(); ms.append(a.elapsed_time(b))
    F = 4.0 * B * Hq * N * N * d * (0.5 if causal else 1.0); m = sorted(ms)[len(ms) // 2]
    # sanity: one head against torch SDPA in fp32
    ref = torch.nn.functional.scaled_dot_product_attention(q[:1, :1, :1024].float(), k[:1, :1, :1024].float(), v[:1, :1, :1024].float(), is_causal=causal)
    chk = fa_b200.attention_forward(q[:1, :1, :1024].contiguous(), k[:1, :1, :1024].contiguous(), v[:1, :1, :1024].contiguous(), causal=causal)
    res[f"N{N}_d{d}_{'c' if causal else 'nc'}_{dt}"] = {"ms": round(m, 4), "min": round(min(ms), 4), "tflops": round(F / m / 1e9, 1), "err": round((chk.float() - ref).abs().max().item(), 5)}
    del q, k, v, o
print(json.dumps({"lib": label, **res}), flush=True)
