"""cfg2_probe.py — BASELINE configs[1] (GPT-2 shape: B=4 H=12 N=1024 d=64 non-causal fp16), the launch- and tail-bound case:
eager and CUDA-graph-replay time per launch for every kernel variant, with the half-item tail schedule on and off, and the
slowest CTA's cycle count.  GPU box only."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "flash-attention-cuda-c_b200")]
import torch, fa_b200
L = fa_b200.lib()
L.fa_debug_set_profile_buffer.argtypes = [ctypes.c_void_p]
shapes = [(4, 12, 1024, 64, False, torch.float16), (4, 12, 1024, 64, True, torch.float16), (2, 16, 2048, 128, True, torch.bfloat16), (1, 24, 1024, 128, False, torch.bfloat16)]
prof = torch.zeros(32, dtype=torch.int64, device="cuda")


def timeit(fn, iters, reps=5):
    for _ in range(20): fn()
    torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters): fn()
        b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / iters)
    return best


for (B, H, N, d, causal, dt) in shapes:
    g = torch.Generator(device="cuda").manual_seed(0)
    q, k, v = (torch.randn(B, H, N, d, device="cuda", generator=g).to(dt) for _ in range(3))
    o = torch.empty_like(q)
    F = 4.0 * B * H * N * N * d * (0.5 if causal else 1.0)
    for (sw, emu, epi) in ((8, 0, 0), (8, 0, 1), (16, 1, 0)):
        for half in (0, 2, 1):      # no half items / half items on slot 0 alone / split-KV half items
            fa_b200.force_variant(sw, emu, epi); L.fa_debug_half_items(half)
            fn = lambda: fa_b200.attention_forward(q, k, v, causal=causal, out=o)
            eager = timeit(fn, 200)
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                for _ in range(20): fn()
            graph = timeit(gr.replay, 10) / 20
            L.fa_debug_set_profile_buffer(prof.data_ptr()); cyc = []
            for _ in range(5):
                prof.zero_(); fn(); torch.cuda.synchronize(); cyc.append(int(prof[30].item()))
            L.fa_debug_set_profile_buffer(None)
            print(json.dumps({"shape": f"B{B}_H{H}_N{N}_d{d}_{'c' if causal else 'nc'}", "softmax_warps": sw, "emu": emu, "staged_epilogue": epi, "half_items": half,
                              "eager_us": round(eager * 1e3, 2), "graph_us": round(graph * 1e3, 2), "tflops_graph": round(F / graph / 1e9, 1),
                              "max_cta_cycles": min(cyc)}), flush=True)
fa_b200.force_variant(0, 0, 0); L.fa_debug_half_items(1)
