"""make_golden.py — generate golden vectors by running the REFERENCE's own check.py (imported from
/root/reference, which only exists in the build container).  Run once there; the .npz outputs are committed
so the oracle and the CUDA path can be checked against the real reference on the GPU box too.

    python tests/golden/make_golden.py

Each fixture holds Q, K, V in check.py's (batch, seq_len, d_model) layout, num_heads, and the reference's
`output` (and `attn` for the small ones).  check.py has no causal option, so causal cases have no reference
golden; they are pinned by the rule restated from tests/main.cu:81 and cross-checked against torch SDPA.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/check.py"


def load_reference():
    spec = importlib.util.spec_from_file_location("reference_check", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)   # its demo sits under `if __name__ == "__main__"`, so import is side-effect free
    return mod


def main():
    if not os.path.exists(REF):
        sys.exit("reference not mounted; golden vectors are generated in the build container only")
    ref = load_reference()
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    cases = {
        # the reference's own demo / known-answer test: all ones in -> attn 0.25, output 1 (check.py:28-43)
        "kat_ones_b1_n4_h2_dm8": dict(ones=True, B=1, N=4, H=2, dm=8),
        # BASELINE.json configs[0]: single-head fp32 B=1 H=1 N=256 d=64
        "cfg1_b1_n256_h1_d64": dict(ones=False, B=1, N=256, H=1, dm=64),
        # multi-head, multi-batch, ragged length (not a tile multiple)
        "mh_b2_n200_h4_d32": dict(ones=False, B=2, N=200, H=4, dm=128),
        # head dim 128, several heads
        "mh_b1_n384_h2_d128": dict(ones=False, B=1, N=384, H=2, dm=256),
    }
    for name, c in cases.items():
        if c["ones"]:
            Q = torch.ones(c["B"], c["N"], c["dm"]); K = Q.clone(); V = Q.clone()
        else:
            # seeds 0, 1, 2 for Q, K, V
            Q = torch.randn(c["B"], c["N"], c["dm"], generator=torch.Generator().manual_seed(0))
            K = torch.randn(c["B"], c["N"], c["dm"], generator=torch.Generator().manual_seed(1))
            V = torch.randn(c["B"], c["N"], c["dm"], generator=torch.Generator().manual_seed(2))
        out, attn = ref.multi_head_attention(Q, K, V, c["H"])
        payload = dict(Q=Q.numpy(), K=K.numpy(), V=V.numpy(), num_heads=np.int32(c["H"]), output=out.numpy())
        if c["N"] <= 8:
            payload["attn"] = attn.numpy()
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **payload)
        print(f"{name}: out mean {out.mean().item():+.6f}  -> {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
