"""gqa_probe.py name... — GQA against MHA at equal FLOPs (cycles of the slowest CTA and back-to-back time), for the listed builds."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "flash-attention-cuda-c_b200")]
import torch, fa_b200
names = sys.argv[1:]
libs = {}
for n in names:
    fa_b200._lib = None
    fa_b200.LIB_PATH = os.path.join(ROOT, "variants", f"libfa_v_{n}.so") if n != "shipped" else os.path.join(ROOT, "flash-attention-cuda-c_b200", "libfa_b200.so")
    libs[n] = fa_b200.lib(); libs[n].fa_debug_set_profile_buffer.argtypes = [ctypes.c_void_p]
prof = torch.zeros(32, dtype=torch.int64, device="cuda")
for (B, Hq, Hkv, N, causal) in ((4, 64, 64, 8192, True), (4, 64, 8, 8192, True), (4, 64, 8, 8192, False), (4, 64, 64, 8192, False), (1, 64, 8, 32768, True), (1, 64, 64, 32768, True)):
    g = torch.Generator(device="cuda").manual_seed(0)
    q = torch.randn(B, Hq, N, 128, device="cuda", generator=g).to(torch.bfloat16)
    k, v = (torch.randn(B, Hkv, N, 128, device="cuda", generator=g).to(torch.bfloat16) for _ in range(2))
    o = torch.empty_like(q)
    F = 4.0 * B * Hq * N * N * 128 * (0.5 if causal else 1.0)
    ref = None
    for n in names:
        fa_b200._lib = libs[n]
        for _ in range(3): fa_b200.attention_forward(q, k, v, causal=causal, out=o)
        if ref is None: ref = o.clone()
        same = bool(torch.equal(o, ref))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20): fa_b200.attention_forward(q, k, v, causal=causal, out=o)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 20
        libs[n].fa_debug_set_profile_buffer(prof.data_ptr()); cyc = []
        for _ in range(3):
            prof.zero_(); fa_b200.attention_forward(q, k, v, causal=causal, out=o); torch.cuda.synchronize(); cyc.append(int(prof[30].item()))
        libs[n].fa_debug_set_profile_buffer(None)
        print(json.dumps({"shape": f"B{B}_Hq{Hq}_Hkv{Hkv}_N{N}_{'c' if causal else 'nc'}", "lib": n, "ms": round(ms, 4), "tflops": round(F / ms / 1e9, 1),
                          "max_cta_cycles": min(cyc), "same_bits_as_first": same}), flush=True)
    del q, k, v, o
