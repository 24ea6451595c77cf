"""ab.py name1 name2 ... — A/B several builds (variants/libfa_v_<name>.so) in ONE process, interleaved round-robin so that
clock / power drift hits all of them alike.  CUDA events, FA_TIME_REPS launches per (build, shape, round), median and min.
Env: FA_AB_SHAPES = comma list of indices into SHAPES (default "0,1"), FA_AB_ROUNDS (default 3)."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "flash-attention-cuda-c_b200")]
import torch, fa_b200
SHAPES = [(8, 32, 32, 8192, 128, True, "bf16"), (8, 32, 32, 8192, 128, False, "bf16"), (32, 32, 32, 2048, 128, True, "bf16"),
          (8, 32, 32, 8192, 64, True, "bf16"), (64, 32, 32, 1024, 128, True, "bf16"), (4, 12, 12, 1024, 64, False, "fp16")]
names = sys.argv[1:]
sel = [int(x) for x in os.environ.get("FA_AB_SHAPES", "0,1").split(",")]
rounds = int(os.environ.get("FA_AB_ROUNDS", "3")); reps = int(os.environ.get("FA_TIME_REPS", "10"))
libs = {}
for n in names:
    fa_b200._lib = None
    fa_b200.LIB_PATH = os.path.join(ROOT, "variants", f"libfa_v_{n}.so") if n != "shipped" else os.path.join(ROOT, "flash-attention-cuda-c_b200", "libfa_b200.so")
    libs[n] = fa_b200.lib()
def use(n): fa_b200._lib = libs[n]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for si in sel:
    B, Hq, Hkv, N, d, causal, dt = SHAPES[si]
    t = {"bf16": torch.bfloat16, "fp16": torch.float16}[dt]
    g = torch.Generator(device="cuda").manual_seed(0)
    q = torch.randn(B, Hq, N, d, device="cuda", generator=g).to(t); k = torch.randn(B, Hkv, N, d, device="cuda", generator=g).to(t); v = torch.randn(B, Hkv, N, d, device="cuda", generator=g).to(t)
    o = torch.empty_like(q)
    small = (2 * q.numel() + 2 * k.numel()) * 2 < (256 << 20)
    ref = torch.nn.functional.scaled_dot_product_attention(q[:1, :2, :1024].float(), k[:1, :2 if Hkv == Hq else 1, :1024].float().repeat_interleave(1 if Hkv == Hq else 2, 1), v[:1, :2 if Hkv == Hq else 1, :1024].float().repeat_interleave(1 if Hkv == Hq else 2, 1), is_causal=causal)
    F = 4.0 * B * Hq * N * N * d * (0.5 if causal else 1.0)
    res = {n: [] for n in names}; err = {}
    for n in names:
        use(n)
        for _ in range(3): fa_b200.attention_forward(q, k, v, causal=causal, out=o)
        chk = fa_b200.attention_forward(q[:1, :2, :1024].contiguous(), k[:1, :2 if Hkv == Hq else 1, :1024].contiguous(), v[:1, :2 if Hkv == Hq else 1, :1024].contiguous(), causal=causal)
        err[n] = round((chk.float() - ref).abs().max().item(), 5)
    torch.cuda.synchronize()
    for r in range(rounds):
        for n in names[r % len(names):] + names[:r % len(names)]:      # rotate the order: no build always runs first (coolest)
            use(n)
            for _ in range(reps):
                if small: flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fa_b200.attention_forward(q, k, v, causal=causal, out=o); b.record(); torch.cuda.synchronize(); res[n].append(a.elapsed_time(b))
    for n in names:
        ms = sorted(res[n]); m = ms[len(ms) // 2]
        print(json.dumps({"shape": f"N{N}_d{d}_{'c' if causal else 'nc'}_{dt}_B{B}", "lib": n, "ms_med": round(m, 4), "ms_min": round(ms[0], 4), "tflops_med": round(F / m / 1e9, 1), "tflops_best": round(F / ms[0] / 1e9, 1), "err": err[n]}), flush=True)
    del q, k, v, o
