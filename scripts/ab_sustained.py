"""ab_sustained.py name1 name2 ... — wall-clock A/B in the power-capped steady state (what bench.py measures): per round and
build, FA_SUS_WARM untimed launches to settle clocks, then FA_SUS_N launches back to back under ONE CUDA-event pair; builds
are interleaved round-robin (rotating start) so drift hits all alike.  Reports per-build mean / min / max of the per-round
ms per launch.  Env: FA_AB_SHAPES (default "0"), FA_SUS_ROUNDS (6), FA_SUS_N (120), FA_SUS_WARM (40)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "flash-attention-cuda-c_b200")]
import torch, fa_b200
SHAPES = [(8, 32, 32, 8192, 128, True, "bf16"), (8, 32, 32, 8192, 128, False, "bf16"), (32, 32, 32, 2048, 128, True, "bf16"),
          (8, 32, 32, 8192, 64, True, "bf16"), (64, 32, 32, 1024, 128, True, "bf16"), (4, 12, 12, 1024, 64, False, "fp16"),
          (16, 32, 32, 4096, 128, True, "bf16"), (4, 32, 32, 16384, 128, True, "bf16"), (16, 32, 32, 4096, 128, False, "bf16"), (128, 32, 32, 512, 128, True, "bf16"),
          # GQA 32/8 (10..14): the pair kernel can cut its pairs by heads there
          (64, 32, 8, 1024, 128, True, "bf16"), (32, 32, 8, 2048, 128, True, "bf16"), (16, 32, 8, 4096, 128, True, "bf16"), (8, 32, 8, 8192, 128, True, "bf16"),
          (1, 64, 8, 32768, 128, True, "bf16"),
          # small launches (15..19): 96 / 128 / 256 pair items by rows; 128 / 64 by heads
          (4, 12, 12, 1024, 128, False, "bf16"), (2, 16, 16, 2048, 128, False, "bf16"), (4, 16, 16, 2048, 128, False, "bf16"),
          (2, 16, 4, 2048, 128, True, "bf16"), (1, 16, 4, 2048, 128, False, "bf16")]
names = sys.argv[1:]
sel = [int(x) for x in os.environ.get("FA_AB_SHAPES", "0").split(",")]
rounds = int(os.environ.get("FA_SUS_ROUNDS", "6")); n = int(os.environ.get("FA_SUS_N", "120")); warm = int(os.environ.get("FA_SUS_WARM", "40"))
libs = {}
force = {}
for nm in names:     # "<lib>@sw,emu" forces a compiled kernel variant of that library
    base, _, fv = nm.partition("@")
    fa_b200._lib = None
    fa_b200.LIB_PATH = os.path.join(ROOT, "variants", f"libfa_v_{base}.so") if base != "shipped" else os.path.join(ROOT, "flash-attention-cuda-c_b200", "libfa_b200.so")
    libs[nm] = fa_b200.lib()
    force[nm] = tuple(int(x) for x in (fv + ",0,0").split(",")[:4]) if fv else None
for si in sel:
    B, Hq, Hkv, N, d, causal, dt = SHAPES[si]
    t = {"bf16": torch.bfloat16, "fp16": torch.float16}[dt]
    g = torch.Generator(device="cuda").manual_seed(0)
    q = torch.randn(B, Hq, N, d, device="cuda", generator=g).to(t); k = torch.randn(B, Hkv, N, d, device="cuda", generator=g).to(t); v = torch.randn(B, Hkv, N, d, device="cuda", generator=g).to(t)
    o = torch.empty_like(q)
    F = 4.0 * B * Hq * N * N * d * (0.5 if causal else 1.0)
    res = {nm: [] for nm in names}
    for r in range(rounds):
        for nm in names[r % len(names):] + names[:r % len(names)]:
            fa_b200._lib = libs[nm]
            try:
                fv4 = force[nm] or (0, 0, 0, 0)
                libs[nm].fa_debug_force_variant(*fv4[:3])
                if hasattr(libs[nm], "fa_debug_force_cta_group"): libs[nm].fa_debug_force_cta_group(fv4[3])
            except (AttributeError, TypeError): pass
            for _ in range(warm): fa_b200.attention_forward(q, k, v, causal=causal, out=o)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(n): fa_b200.attention_forward(q, k, v, causal=causal, out=o)
            b.record(); torch.cuda.synchronize()
            res[nm].append(a.elapsed_time(b) / n)
    for nm in names:
        ms = res[nm]; m = sum(ms) / len(ms)
        print(json.dumps({"shape": f"N{N}_d{d}_{'c' if causal else 'nc'}_{dt}_B{B}", "lib": nm, "ms_mean": round(m, 4), "ms_min": round(min(ms), 4), "ms_max": round(max(ms), 4),
                          "tflops_mean": round(F / m / 1e9, 1), "rounds": [round(x, 3) for x in ms]}), flush=True)
    del q, k, v, o
