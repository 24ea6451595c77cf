"""pair_soak.py [n_shapes] [seed] — random shapes through the CTA-pair kernel and the 1-CTA kernel (half items off, so both do the
same per-row arithmetic): outputs and LSE must be BIT-identical.  Also repeats one shape many times on two streams at once.
GPU box only."""
import ctypes, os, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "flash-attention-cuda-c_b200")]
import torch, fa_b200
n_shapes = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
L = fa_b200.lib()
L.fa_debug_half_items.argtypes = [ctypes.c_int]
L.fa_debug_half_items(0)
bad = 0
for i in range(n_shapes):
    g = rng.choice([1, 1, 2, 4, 8])
    Hkv = rng.choice([1, 2, 3, 4]); Hq = Hkv * g
    B = rng.choice([1, 2, 3, 5])
    Nq = rng.choice([1, 17, 128, 129, 255, 256, 257, 511, 512, 513, 700, 1000, 1024, 1500, 2048, 3000, rng.randrange(1, 4097)])
    Nk = Nq if rng.random() < 0.5 else rng.choice([1, 64, 127, 128, 129, 300, 512, 1000, 2048, 4096, rng.randrange(1, 6000)])
    causal = rng.random() < 0.6
    dt = rng.choice([torch.bfloat16, torch.float16])
    stg = rng.choice([0, 1])
    gen = torch.Generator(device="cuda").manual_seed(i)
    q = torch.randn(B, Hq, Nq, 128, device="cuda", generator=gen).to(dt)
    k = torch.randn(B, Hkv, Nk, 128, device="cuda", generator=gen).to(dt)
    v = torch.randn(B, Hkv, Nk, 128, device="cuda", generator=gen).to(dt)
    fa_b200.force_variant(8, 0, stg, 1)
    o1, l1 = fa_b200.attention_forward(q, k, v, causal=causal, return_lse=True)
    fa_b200.force_variant(8, 0, stg, rng.choice([2, 2, 3, 4]))      # pairs cut by as many heads as the group allows / by rows / by at most two heads
    o2, l2 = fa_b200.attention_forward(q, k, v, causal=causal, return_lse=True)
    torch.cuda.synchronize()
    same = torch.equal(o1.view(torch.int16), o2.view(torch.int16)) and torch.equal(l1.view(torch.int32), l2.view(torch.int32))
    if not same:
        bad += 1
        print(f"MISMATCH shape {i}: B{B} Hq{Hq} Hkv{Hkv} Nq{Nq} Nk{Nk} causal={causal} {dt} staged={stg}: max|dO|={(o1.float() - o2.float()).abs().max().item():.3e}", flush=True)
print(f"{n_shapes} random shapes, {bad} mismatches")
# two streams, the pair kernel back to back on both, results checked at the end
q = torch.randn(2, 8, 2048, 128, device="cuda").bfloat16(); k = torch.randn(2, 8, 2048, 128, device="cuda").bfloat16(); v = torch.randn(2, 8, 2048, 128, device="cuda").bfloat16()
fa_b200.force_variant(8, 0, 1, 1); ref = fa_b200.attention_forward(q, k, v, causal=True); torch.cuda.synchronize()
fa_b200.force_variant(8, 0, 1, 2)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
outs = [torch.empty_like(q) for _ in range(2)]
for it in range(400):
    for s, o in ((s1, outs[0]), (s2, outs[1])):
        with torch.cuda.stream(s):
            fa_b200.attention_forward(q, k, v, causal=True, out=o)
torch.cuda.synchronize()
ok = all(torch.equal(o.view(torch.int16), ref.view(torch.int16)) for o in outs)
print("two streams x 400 pair launches:", "OK" if ok else "MISMATCH")
fa_b200.force_variant(0, 0, 0, 0); L.fa_debug_half_items(1)
sys.exit(0 if bad == 0 and ok else 1)
