"""CPU tests (-m "not gpu"): pin the oracle against the reference's golden vectors and known-answer test.

Golden vectors come from the reference's own check.py (tests/golden/make_golden.py).  The causal rule has no
reference golden (check.py has no mask), so it is checked against an independent implementation
(torch.nn.functional.scaled_dot_product_attention) and by properties.
"""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import oracle


def _to_bhnd(x, heads):
    b, n, dm = x.shape
    return np.ascontiguousarray(x.reshape(b, n, heads, dm // heads).transpose(0, 2, 1, 3))


def _golden(golden_dir):
    files = sorted(glob.glob(os.path.join(golden_dir, "*.npz")))
    assert files, "golden fixtures missing"
    return files


def test_golden_files_present(golden_dir):
    names = {os.path.basename(f) for f in _golden(golden_dir)}
    assert "kat_ones_b1_n4_h2_dm8.npz" in names and "cfg1_b1_n256_h1_d64.npz" in names


def test_c_oracle_matches_reference_golden(golden_dir):
    for f in _golden(golden_dir):
        g = np.load(f)
        h = int(g["num_heads"])
        o = oracle.attention_fwd(_to_bhnd(g["Q"], h), _to_bhnd(g["K"], h), _to_bhnd(g["V"], h))
        ref = _to_bhnd(g["output"], h)
        # reference output is fp32 torch; oracle accumulates in double: agreement to fp32 round-off
        np.testing.assert_allclose(o, ref, rtol=2e-5, atol=2e-6, err_msg=f)


def test_torch_port_matches_reference_golden(golden_dir):
    for f in _golden(golden_dir):
        g = np.load(f)
        out, attn = oracle.multi_head_attention(torch.from_numpy(g["Q"]), torch.from_numpy(g["K"]),
                                                torch.from_numpy(g["V"]), int(g["num_heads"]))
        np.testing.assert_allclose(out.numpy(), g["output"], rtol=1e-5, atol=1e-6, err_msg=f)
        if "attn" in g:
            np.testing.assert_allclose(attn.numpy(), g["attn"], rtol=1e-6, atol=1e-7)


def test_known_answer_all_ones():
    # the reference's only KAT: Q = K = V = 1 -> O = 1 (tests/main.cu:33-35 with B=H=1, N=16, D=16; check.py:36-38)
    x = np.ones((1, 1, 16, 16), np.float32)
    for causal in (False, True):
        o = oracle.attention_fwd(x, x, x, causal=causal)
        np.testing.assert_allclose(o, 1.0, rtol=0, atol=1e-7)
    g = np.ones((1, 4, 8), np.float32)
    out, attn = oracle.multi_head_attention(torch.from_numpy(g), torch.from_numpy(g), torch.from_numpy(g), 2)
    assert torch.allclose(attn, torch.full_like(attn, 0.25)) and torch.allclose(out, torch.ones_like(out))


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("shape", [(1, 2, 2, 37, 37, 16), (2, 4, 2, 64, 64, 32), (1, 8, 1, 50, 90, 64), (1, 2, 2, 1, 33, 16)])
def test_c_oracle_matches_torch_sdpa(shape, causal):
    B, Hq, Hkv, Nq, Nk, d = shape
    g = torch.Generator().manual_seed(3)
    q = torch.randn(B, Hq, Nq, d, generator=g)
    k = torch.randn(B, Hkv, Nk, d, generator=g)
    v = torch.randn(B, Hkv, Nk, d, generator=g)
    o, lse = oracle.attention_fwd(q.numpy(), k.numpy(), v.numpy(), causal=causal, return_lse=True)
    rep = Hq // Hkv
    kk, vv = k.repeat_interleave(rep, 1), v.repeat_interleave(rep, 1)
    mask = None
    if causal:   # bottom-right aligned: key j visible to query i iff j <= i + (Nk - Nq)
        i = torch.arange(Nq)[:, None]
        j = torch.arange(Nk)[None, :]
        mask = j <= i + (Nk - Nq)
    ref = torch.nn.functional.scaled_dot_product_attention(q.double(), kk.double(), vv.double(), attn_mask=mask)
    np.testing.assert_allclose(o, ref.float().numpy(), rtol=1e-5, atol=1e-6)
    s = (q.double() @ kk.double().transpose(-1, -2)) / d ** 0.5
    if mask is not None:
        s = s.masked_fill(~mask, float("-inf"))
    np.testing.assert_allclose(lse, torch.logsumexp(s, -1).float().numpy(), rtol=1e-5, atol=1e-5)


def test_oracle_causal_rows_without_keys_are_zero():
    # Nq > Nk with bottom-right alignment: the first Nq-Nk queries see no key -> zero output, lse = -inf
    q = np.random.default_rng(0).standard_normal((1, 1, 6, 8), dtype=np.float32)
    k = np.random.default_rng(1).standard_normal((1, 1, 4, 8), dtype=np.float32)
    o, lse = oracle.attention_fwd(q, k, k, causal=True, return_lse=True)
    assert np.all(o[0, 0, :2] == 0) and np.all(np.isneginf(lse[0, 0, :2])) and np.all(np.isfinite(lse[0, 0, 2:]))


def test_oracle_properties_linearity_and_rowsum():
    rng = np.random.default_rng(5)
    q, k = (rng.standard_normal((1, 2, 40, 16), dtype=np.float32) for _ in range(2))
    v1, v2 = (rng.standard_normal((1, 2, 40, 16), dtype=np.float32) for _ in range(2))
    o12 = oracle.attention_fwd(q, k, v1 + v2, causal=True)
    np.testing.assert_allclose(o12, oracle.attention_fwd(q, k, v1, causal=True) + oracle.attention_fwd(q, k, v2, causal=True),
                               rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(oracle.attention_fwd(q, k, np.ones_like(v1), causal=True), 1.0, atol=1e-6)


def test_oracle_rejects_bad_sizes():
    x = np.ones((1, 3, 4, 8), np.float32)
    y = np.ones((1, 2, 4, 8), np.float32)
    with pytest.raises(ValueError):
        oracle.attention_fwd(x, y, y)   # Hq % Hkv != 0
