"""pair_check.py — CTA-pair kernel (cta_group::2) against the 1-CTA kernel on the same inputs: max |difference| of O and LSE
per shape.  GPU box only.  Usage: pair_check.py [staged=1]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "flash-attention-cuda-c_b200")]
import torch, fa_b200
stg = int(sys.argv[1]) if len(sys.argv) > 1 else 1
shapes = [(1, 1, 1, 512, 512, False), (1, 1, 1, 512, 512, True), (1, 2, 2, 1024, 1024, True), (2, 4, 2, 1000, 1000, True),
          (1, 2, 1, 700, 1300, False), (1, 2, 2, 300, 2000, True), (2, 8, 8, 2048, 2048, True), (8, 32, 32, 8192, 8192, True)]
for B, Hq, Hkv, Nq, Nk, causal in shapes:
    g = torch.Generator(device="cuda").manual_seed(Nq + Nk)
    q = torch.randn(B, Hq, Nq, 128, device="cuda", generator=g).bfloat16()
    k = torch.randn(B, Hkv, Nk, 128, device="cuda", generator=g).bfloat16()
    v = torch.randn(B, Hkv, Nk, 128, device="cuda", generator=g).bfloat16()
    fa_b200.force_variant(8, 0, stg, 1)
    o1, l1 = fa_b200.attention_forward(q, k, v, causal=causal, return_lse=True)
    torch.cuda.synchronize()
    fa_b200.force_variant(8, 0, stg, 2)
    o2, l2 = fa_b200.attention_forward(q, k, v, causal=causal, return_lse=True)
    torch.cuda.synchronize()
    do = (o1.float() - o2.float()).abs().max().item()
    fin = torch.isfinite(l1) & torch.isfinite(l2)
    dl = (l1[fin] - l2[fin]).abs().max().item() if fin.any() else 0.0
    print(f"B{B} Hq{Hq} Hkv{Hkv} Nq{Nq} Nk{Nk} causal={causal}: max|dO|={do:.3e} max|dLSE|={dl:.3e} nan={bool(torch.isnan(o2.float()).any())}", flush=True)
fa_b200.force_variant(0, 0, 0, 0)
