"""tile_sweep.py — the measurement behind the tile table (FlashAttention.cu: kTileTable; fa_tile_table in the C ABI).
For every (head dim, causal, sequence-length bucket) every compiled kernel variant is timed on the same inputs in one
process — variants interleaved round-robin, each round FA_TS_MS milliseconds of back-to-back launches under one CUDA-event
pair after a warm-up (the power-capped steady state bench.py's `sustained` leg reports) — and the fastest is named.
Writes one JSON object per (shape, variant) plus one "winner" object per shape.  GPU box only.
Env: FA_TS_ROUNDS (3), FA_TS_MS (60), FA_TS_DTYPE (bf16)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "flash-attention-cuda-c_b200")]
import torch, fa_b200

VARIANTS = [(8, 0, 0, 1), (8, 0, 1, 1), (16, 1, 0, 1), (8, 0, 1, 2)]      # (softmax warps, exp2 pairs of 8 on the FMA pipe, staged TMA-store epilogue, CTAs per MMA: 2 = the CTA-pair kernel, d = 128 only)
rounds = int(os.environ.get("FA_TS_ROUNDS", "3")); budget_ms = float(os.environ.get("FA_TS_MS", "60"))
dt = {"bf16": torch.bfloat16, "fp16": torch.float16}[os.environ.get("FA_TS_DTYPE", "bf16")]
for d in (128, 64):
    for causal in (False, True):
        for N in (512, 1024, 2048, 4096, 8192, 16384, 32768):
            BH = max(8, min(4096, 256 * 8192 // N))
            B, H = max(1, BH // 32), min(32, BH)
            g = torch.Generator(device="cuda").manual_seed(N + d)
            q, k, v = (torch.randn(B, H, N, d, device="cuda", generator=g).to(dt) for _ in range(3))
            o = torch.empty_like(q)
            F = 4.0 * B * H * N * N * d * (0.5 if causal else 1.0)
            variants = [vv for vv in VARIANTS if vv[3] == 1 or d == 128]
            res = {vv: [] for vv in variants}
            n_launch = {}
            for vv in variants:      # calibrate the launch count per round
                fa_b200.force_variant(*vv)
                for _ in range(3): fa_b200.attention_forward(q, k, v, causal=causal, out=o)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fa_b200.attention_forward(q, k, v, causal=causal, out=o); b.record(); torch.cuda.synchronize()
                n_launch[vv] = max(5, int(budget_ms / max(a.elapsed_time(b), 1e-3)))
            for r in range(rounds):
                for vv in variants[r % len(variants):] + variants[:r % len(variants)]:
                    fa_b200.force_variant(*vv)
                    for _ in range(max(3, n_launch[vv] // 4)): fa_b200.attention_forward(q, k, v, causal=causal, out=o)
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    for _ in range(n_launch[vv]): fa_b200.attention_forward(q, k, v, causal=causal, out=o)
                    b.record(); torch.cuda.synchronize()
                    res[vv].append(a.elapsed_time(b) / n_launch[vv])
            best = None
            for vv in variants:
                ms = sum(res[vv]) / len(res[vv])
                rec = {"d": d, "causal": causal, "N": N, "B": B, "H": H, "dtype": str(dt).split(".")[-1], "softmax_warps": vv[0], "emu": vv[1], "staged_epilogue": vv[2], "cta_group": vv[3],
                       "ms_mean": round(ms, 5), "ms_min": round(min(res[vv]), 5), "tflops": round(F / ms / 1e9, 1), "launches_per_round": n_launch[vv]}
                print(json.dumps(rec), flush=True)
                if best is None or ms < best[1]: best = (vv, ms)
            print(json.dumps({"winner": True, "d": d, "causal": causal, "N": N, "softmax_warps": best[0][0], "emu": best[0][1], "staged_epilogue": best[0][2], "cta_group": best[0][3],
                              "tflops": round(F / best[1] / 1e9, 1)}), flush=True)
            del q, k, v, o
fa_b200.force_variant(0, 0, 0, 0)
