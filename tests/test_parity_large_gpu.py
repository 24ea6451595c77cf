"""GPU parity tests (-m gpu), part 2: the regimes the first suite did not reach.

  * the multi-item regime of the persistent kernel: more work items than SMs, so every CTA hands over between items
    (o_free hand-back, barrier phase wrap, virtual score steps across items, scheduler mailbox) — swept over head dim,
    causal, dtype, output alignment (256-bit / 128-bit epilogue stores), Nq = Nk and ragged Nq != Nk, GQA, and the
    ring-KV carry mode;
  * BASELINE.json sizes: sampled query-row blocks of N = 32K / 64K / 128K causal problems against the oracle (rows
    [r0, r0+256) of a causal problem are an Nq=256, Nk=r0+256 bottom-right-aligned problem), config 4 at its full
    single-GPU size (Q is 4.3 G elements: offsets beyond 2^31 and 2^32), d = 64 at N = 8K;
  * ring-KV through sharding.ring_attention's default CUDA operators at world size 1, against the oracle;
  * launch-path state: thousands of launches across two streams, a graph replay beside eager launches, two host
    threads inside fa_fwd_host.

The oracle (oracle/attention_oracle.c) is the ground truth everywhere; where a case has hundreds of (batch, head)
slices the oracle checks a sample of them and a plain fp32 torch matmul-softmax-matmul on the GPU (TF32 off) checks
every element, so that a wrong hand-over anywhere in the queue cannot hide.  Tolerance: BASELINE.json's 2e-2 max-abs for
bf16 / fp16.
"""
import os
import threading

import numpy as np
import pytest
import torch

import fa_b200
import sharding
from oracle import oracle

pytestmark = pytest.mark.gpu

TOL16 = 2e-2


@pytest.fixture(scope="module", autouse=True)
def _lib():
    assert torch.cuda.is_available(), "GPU tests need a B200"
    assert os.path.exists(fa_b200.LIB_PATH), "libfa_b200.so missing: the CUDA path must be built, there is no fallback"
    fa_b200.lib()
    torch.backends.cuda.matmul.allow_tf32 = False
    before = fa_b200.launch_count()
    yield
    assert fa_b200.launch_count() > before


def _rand(shape, dtype, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(shape, device="cuda", generator=g).to(dtype)


def _torch_ref(q, k, v, causal, scale=None):
    """fp32 matmul-softmax-matmul on the GPU (secondary, every element); GQA by head repetition; bottom-right causal."""
    B, Hq, Nq, d = q.shape
    Hkv, Nk = k.shape[1], k.shape[2]
    kf = k.float().repeat_interleave(Hq // Hkv, dim=1)
    vf = v.float().repeat_interleave(Hq // Hkv, dim=1)
    s = (q.float() @ kf.transpose(-1, -2)) * (scale if scale else d ** -0.5)
    if causal:
        i = torch.arange(Nq, device=q.device)[:, None]
        j = torch.arange(Nk, device=q.device)[None, :]
        s = s.masked_fill(j > i + (Nk - Nq), float("-inf"))
    m = s.amax(dim=-1, keepdim=True)
    m = torch.where(torch.isfinite(m), m, torch.zeros_like(m))
    p = torch.exp(s - m)
    l = p.sum(dim=-1, keepdim=True)
    o = (p @ vf) / torch.where(l > 0, l, torch.ones_like(l))
    lse = torch.where(l > 0, m + torch.log(l), torch.full_like(l, float("-inf"))).squeeze(-1)
    return o, lse


def _oracle_slices(q, k, v, o, lse, causal, pairs):
    """Oracle check of the given (batch, head) slices."""
    g = q.shape[1] // k.shape[1]
    for (b, h) in pairs:
        hk = h // g
        o_ref, l_ref = oracle.attention_fwd(q[b:b + 1, h:h + 1].float().cpu().numpy(), k[b:b + 1, hk:hk + 1].float().cpu().numpy(),
                                            v[b:b + 1, hk:hk + 1].float().cpu().numpy(), causal=causal, return_lse=True)
        err = np.abs(o[b:b + 1, h:h + 1].float().cpu().numpy() - o_ref).max()
        assert err <= TOL16, f"(b={b}, h={h}): max abs error {err:.3e} vs oracle"
        if lse is not None:
            got = lse[b:b + 1, h:h + 1].cpu().numpy()
            fin = np.isfinite(l_ref)
            assert (np.isfinite(got) == fin).all()
            np.testing.assert_allclose(got[fin], l_ref[fin], rtol=0, atol=2e-3)


# ---- the multi-item regime ------------------------------------------------------------------------------------
@pytest.mark.parametrize("nq,nk", [(1000, 1000), (1000, 1200)])
@pytest.mark.parametrize("o_off", [0, 8])          # elements: 0 -> 256-bit epilogue stores, 8 (16 bytes) -> 128-bit stores
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("d", [128, 64])
def test_many_work_items(d, causal, dtype, o_off, nq, nk):
    # 7 x 15 heads x 4 query blocks = 420 work items on 148 CTAs: every CTA runs 2-3 items back to back
    B, Hq, Hkv = 7, 15, (15 if d == 128 else 5)
    q, k, v = _rand((B, Hq, nq, d), dtype, 1), _rand((B, Hkv, nk, d), dtype, 2), _rand((B, Hkv, nk, d), dtype, 3)
    big = torch.full((B, Hq, nq, d + 16), 512.0, dtype=dtype, device="cuda")
    out = big[..., o_off:o_off + d]
    assert out.data_ptr() % 32 == (16 if o_off else 0)
    _, lse = fa_b200.attention_forward(q, k, v, causal=causal, return_lse=True, out=out)
    torch.cuda.synchronize()
    assert (big[..., :o_off] == 512.0).all() and (big[..., o_off + d:] == 512.0).all()
    o_t, lse_t = _torch_ref(q, k, v, causal)
    assert torch.isfinite(out.float()).all()
    assert (out.float() - o_t).abs().max().item() <= TOL16
    fin = torch.isfinite(lse_t)
    assert (torch.isfinite(lse) == fin).all() and (lse[fin] - lse_t[fin]).abs().max().item() <= 2e-3
    _oracle_slices(q, k, v, out, lse, causal, [(0, 0), (3, 7), (B - 1, Hq - 1)])


@pytest.mark.parametrize("d", [128, 64])
def test_many_work_items_carry_mode(d):
    # fa_fwd_carry with more items than CTAs: three key ranges folded into the fp32 running pair, the last one causal
    B, Hq, Hkv, nq, nk = 7, 15, 5, 1000, 1200
    q, k, v = _rand((B, Hq, nq, d), torch.bfloat16, 4), _rand((B, Hkv, nk, d), torch.bfloat16, 5), _rand((B, Hkv, nk, d), torch.bfloat16, 6)
    acc_o = torch.zeros(B, Hq, nq, d, device="cuda")
    acc_l = torch.full((B, Hq, nq), float("-inf"), device="cuda")
    bounds = [0, 384, 640, nk]
    for i, (s, e) in enumerate(zip(bounds[:-1], bounds[1:])):
        fa_b200.attention_forward_carry(q, k[:, :, s:e], v[:, :, s:e], acc_o, acc_l, causal=(i == len(bounds) - 2))
    torch.cuda.synchronize()
    # keys < 640 fully visible; keys 640.. causal, bottom-right aligned inside their block: key j visible iff j - 640 <= i + (560 - 1000)
    o1, l1 = _torch_ref(q, k[:, :, :640], v[:, :, :640], False)
    o2, l2 = _torch_ref(q, k[:, :, 640:], v[:, :, 640:], True)
    l_ref = torch.logaddexp(l1, l2)
    w1 = torch.exp(l1 - l_ref).nan_to_num(0.0)[..., None]
    w2 = torch.exp(l2 - l_ref).nan_to_num(0.0)[..., None]
    o_ref = o1 * w1 + o2 * w2
    assert (acc_o - o_ref).abs().max().item() <= TOL16
    assert (acc_l - l_ref).abs().max().item() <= 2e-3
    # and one slice against the oracle, combined the same way
    b, h, hk = B - 1, Hq - 1, (Hq - 1) // (Hq // Hkv)
    qq, kk, vv = (t.float().cpu().numpy() for t in (q[b:b + 1, h:h + 1], k[b:b + 1, hk:hk + 1], v[b:b + 1, hk:hk + 1]))
    a1, m1 = oracle.attention_fwd(qq, kk[:, :, :640], vv[:, :, :640], return_lse=True)
    a2, m2 = oracle.attention_fwd(qq, kk[:, :, 640:], vv[:, :, 640:], causal=True, return_lse=True)
    mm = np.logaddexp(m1, m2)
    with np.errstate(invalid="ignore"):
        ref = a1 * np.nan_to_num(np.exp(m1 - mm))[..., None] + a2 * np.nan_to_num(np.exp(m2 - mm))[..., None]
    assert np.abs(acc_o[b:b + 1, h:h + 1].cpu().numpy() - ref).max() <= TOL16


# ---- the CTA-pair kernel (clusters of 2, tcgen05 cta_group::2) and its 1-CTA counterpart, each forced --------------------
@pytest.mark.parametrize("cta_group", [1, 2, 3, 4])   # 1-CTA kernel; pairs cut by four heads (4 query heads per kv group); by rows; by two heads
@pytest.mark.parametrize("staged", [0, 1])
@pytest.mark.parametrize("causal,nq,nk", [(False, 1000, 1000), (True, 1000, 1000), (True, 700, 1300), (False, 513, 384), (True, 2100, 2100)])
def test_cta_pair_and_single_kernels_forced(cta_group, staged, causal, nq, nk):
    # The launcher picks the pair kernel for large d = 128 launches (MHA: non-causal, causal Nk >= 8K; GQA: always); here every
    # kernel runs every shape: 512-row pair items with a ragged last item (rows past Nq in the peer CTA only: nq = 513), fully
    # masked leader half tiles on every causal diagonal, two heads of a kv group in one MMA, Nq != Nk, more items than pairs, LSE.
    B, Hq, Hkv, d = 6, 12, 3, 128
    dtype = torch.float16 if (nq + staged) % 2 else torch.bfloat16
    q, k, v = _rand((B, Hq, nq, d), dtype, 31), _rand((B, Hkv, nk, d), dtype, 32), _rand((B, Hkv, nk, d), dtype, 33)
    try:
        fa_b200.force_variant(8, 0, staged, cta_group)
        out, lse = fa_b200.attention_forward(q, k, v, causal=causal, return_lse=True)
        torch.cuda.synchronize()
    finally:
        fa_b200.force_variant(0, 0, 0, 0)
    o_t, lse_t = _torch_ref(q, k, v, causal)
    assert torch.isfinite(out.float()).all()
    assert (out.float() - o_t).abs().max().item() <= TOL16
    fin = torch.isfinite(lse_t)
    assert (torch.isfinite(lse) == fin).all() and (lse[fin] - lse_t[fin]).abs().max().item() <= 2e-3
    _oracle_slices(q, k, v, out, lse, causal, [(0, 0), (B - 1, Hq - 1)])


def test_cta_pair_kernel_bit_identical_to_single_kernel_on_random_shapes():
    # Both kernels do the same per-row arithmetic in the same order (half items off: a split-KV tail sums in another order), so
    # O and LSE must agree to the bit on any shape — 120 random ones: B, GQA group, ragged Nq, Nq != Nk, causal, dtype, epilogue.
    import ctypes
    import random
    rng = random.Random(7)
    L = fa_b200.lib()
    L.fa_debug_half_items.argtypes = [ctypes.c_int]
    try:
        L.fa_debug_half_items(0)
        for i in range(120):
            g = rng.choice([1, 1, 2, 3, 4, 8])      # odd groups: pairs cut by rows; even ones: by heads
            Hkv = rng.choice([1, 2, 3]); Hq = Hkv * g
            B = rng.choice([1, 2, 3])
            Nq = rng.choice([1, 17, 128, 129, 255, 256, 257, 511, 512, 513, 700, 1000, 1024, 1500, 2048, rng.randrange(1, 3000)])
            Nk = Nq if rng.random() < 0.5 else rng.choice([1, 64, 127, 128, 129, 300, 512, 1000, 2048, rng.randrange(1, 4000)])
            causal = rng.random() < 0.6
            dt = rng.choice([torch.bfloat16, torch.float16])
            stg = rng.choice([0, 1])
            q, k, v = _rand((B, Hq, Nq, 128), dt, 3 * i), _rand((B, Hkv, Nk, 128), dt, 3 * i + 1), _rand((B, Hkv, Nk, 128), dt, 3 * i + 2)
            fa_b200.force_variant(8, 0, stg, 1)
            o1, l1 = fa_b200.attention_forward(q, k, v, causal=causal, return_lse=True)
            fa_b200.force_variant(8, 0, stg, rng.choice([2, 2, 3, 4]))      # pairs cut by as many heads as the group allows / by rows / by <= 2 heads
            o2, l2 = fa_b200.attention_forward(q, k, v, causal=causal, return_lse=True)
            torch.cuda.synchronize()
            what = f"shape {i}: B{B} Hq{Hq} Hkv{Hkv} Nq{Nq} Nk{Nk} causal={causal} {dt} staged={stg}"
            assert torch.equal(o1.view(torch.int16), o2.view(torch.int16)), what
            assert torch.equal(l1.view(torch.int32), l2.view(torch.int32)), what
    finally:
        fa_b200.force_variant(0, 0, 0, 0)
        L.fa_debug_half_items(1)


def test_cta_pair_kernel_carry_window():
    # ring-step form on the pair kernel: two key ranges folded into a row window of a larger fp32 accumulator
    B, Hq, Hkv, nq, nk, d = 3, 8, 8, 1024, 1536, 128
    q, k, v = _rand((B, Hq, nq, d), torch.bfloat16, 41), _rand((B, Hkv, nk, d), torch.bfloat16, 42), _rand((B, Hkv, nk, d), torch.bfloat16, 43)
    rows, off = nq + 512, 256
    acc_o = torch.zeros(B, Hq, rows, d, device="cuda")
    acc_l = torch.full((B, Hq, rows), float("-inf"), device="cuda")
    try:
        fa_b200.force_variant(8, 0, 0, 2)
        fa_b200.attention_forward_carry(q, k[:, :, :512], v[:, :, :512], acc_o, acc_l, causal=False, row_offset=off)
        fa_b200.attention_forward_carry(q, k[:, :, 512:], v[:, :, 512:], acc_o, acc_l, causal=True, row_offset=off)
        torch.cuda.synchronize()
    finally:
        fa_b200.force_variant(0, 0, 0, 0)
    o1, l1 = _torch_ref(q, k[:, :, :512], v[:, :, :512], False)
    o2, l2 = _torch_ref(q, k[:, :, 512:], v[:, :, 512:], True)
    l_ref = torch.logaddexp(l1, l2)
    o_ref = o1 * torch.exp(l1 - l_ref).nan_to_num(0.0)[..., None] + o2 * torch.exp(l2 - l_ref).nan_to_num(0.0)[..., None]
    assert (acc_o[:, :, off:off + nq] - o_ref).abs().max().item() <= TOL16
    assert (acc_l[:, :, off:off + nq] - l_ref).abs().max().item() <= 2e-3
    assert (acc_o[:, :, :off] == 0).all() and (acc_o[:, :, off + nq:] == 0).all()      # rows outside the window untouched
    b, h = B - 1, Hq - 1
    qq, kk, vv = (t.float().cpu().numpy() for t in (q[b:b + 1, h:h + 1], k[b:b + 1, h:h + 1], v[b:b + 1, h:h + 1]))
    a1, m1 = oracle.attention_fwd(qq, kk[:, :, :512], vv[:, :, :512], return_lse=True)
    a2, m2 = oracle.attention_fwd(qq, kk[:, :, 512:], vv[:, :, 512:], causal=True, return_lse=True)
    mm = np.logaddexp(m1, m2)
    with np.errstate(invalid="ignore"):
        ref = a1 * np.nan_to_num(np.exp(m1 - mm))[..., None] + a2 * np.nan_to_num(np.exp(m2 - mm))[..., None]
    assert np.abs(acc_o[b:b + 1, h:h + 1, off:off + nq].cpu().numpy() - ref).max() <= TOL16


@pytest.mark.parametrize("nk", [128, 384, 1024])           # 1 key tile (slot 1 gets none), odd and even tile counts
@pytest.mark.parametrize("d,dtype", [(128, torch.bfloat16), (64, torch.float16)])
def test_half_item_tail_split_kv(d, dtype, nk):
    # a launch whose last wave is short: 148 query blocks as 256-row items + the rest as 128-row half items; non-causal with
    # Nk a multiple of 128, so the half items run split-KV on both query-tile slots (slot t takes key tiles t, t+2, ...,
    # merged in the epilogue).  176 blocks: 28 x 2 half items; the ragged Nq = 200 leaves the second half of every block with
    # 72 valid rows.
    B, H, nq = 8, 22, 200
    q, k, v = _rand((B, H, nq, d), dtype, 90), _rand((B, H, nk, d), dtype, 91), _rand((B, H, nk, d), dtype, 92)
    out, lse = fa_b200.attention_forward(q, k, v, causal=False, return_lse=True)
    torch.cuda.synchronize()
    o_t, lse_t = _torch_ref(q, k, v, False)
    assert (out.float() - o_t).abs().max().item() <= TOL16
    assert (lse - lse_t).abs().max().item() <= 2e-3
    _oracle_slices(q, k, v, out, lse, False, [(0, 0), (B - 1, H - 1), (B - 1, H - 2)])      # the last blocks are the half items
    # the same launch with the tail on slot 0 alone and with no half items at all gives the same answer to rounding
    L = fa_b200.lib()
    try:
        for mode in (2, 0):
            L.fa_debug_half_items(mode)
            alt = fa_b200.attention_forward(q, k, v, causal=False)
            assert (alt.float() - out.float()).abs().max().item() <= 1e-2
    finally:
        L.fa_debug_half_items(1)


def test_many_work_items_rescale_path():
    # the lazy O rescale (row max growing by more than 2^8 along the keys) in the multi-item regime
    B, H, nq, nk, d = 5, 16, 768, 1536, 128
    q, k, v = _rand((B, H, nq, d), torch.bfloat16, 7), _rand((B, H, nk, d), torch.bfloat16, 8), _rand((B, H, nk, d), torch.bfloat16, 9)
    ramp = torch.linspace(0.05, 4.0, nk, device="cuda")[None, None, :, None]
    k = (k.float() * ramp).to(torch.bfloat16)
    q = (q.float() * 3).to(torch.bfloat16)
    out, lse = fa_b200.attention_forward(q, k, v, causal=False, return_lse=True)
    o_t, lse_t = _torch_ref(q, k, v, False)
    assert (out.float() - o_t).abs().max().item() <= TOL16
    _oracle_slices(q, k, v, out, lse, False, [(0, 0), (B - 1, H - 1)])


# ---- long sequences: sampled query-row blocks against the oracle ----------------------------------------------------
@pytest.mark.parametrize("n,d,heads", [(32768, 128, 2), (65536, 128, 2), (131072, 128, 1), (8192, 64, 4)])
def test_long_sequence_sampled_rows(n, d, heads):
    q, k, v = (_rand((1, heads, n, d), torch.bfloat16, 20 + i) for i in range(3))
    out, lse = fa_b200.attention_forward(q, k, v, causal=True, return_lse=True)
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all()
    h = heads - 1
    for r0 in (0, n // 2 - 100, n - 256):     # first block, an unaligned block in the middle, the last (heaviest) block
        nk = r0 + 256
        o_ref, l_ref = oracle.attention_fwd(q[:, h:h + 1, r0:r0 + 256].float().cpu().numpy(), k[:, h:h + 1, :nk].float().cpu().numpy(),
                                            v[:, h:h + 1, :nk].float().cpu().numpy(), causal=True, return_lse=True)
        err = np.abs(out[:, h:h + 1, r0:r0 + 256].float().cpu().numpy() - o_ref).max()
        assert err <= TOL16, f"N={n} rows {r0}..{r0 + 255}: max abs error {err:.3e}"
        np.testing.assert_allclose(lse[:, h:h + 1, r0:r0 + 256].cpu().numpy(), l_ref, rtol=0, atol=2e-3)


def test_config4_full_size_gqa_32k():
    # BASELINE.json configs[3] on one GPU: B=16, Hq=64, Hkv=8, N=32768, d=128 causal bf16 (Q/O 8.6 GB each, K/V 1 GB each):
    # element offsets of the last batches exceed 2^31 and 2^32 (the reference's flat int row*D_HEAD addressing,
    # kernels/loaders.cuh:57, is the class of bug this would catch)
    B, Hq, Hkv, N, d = 16, 64, 8, 32768, 128
    free, _ = torch.cuda.mem_get_info()
    if free < 40 << 30:
        pytest.skip("needs 40 GB of free device memory")
    q = torch.empty(B, Hq, N, d, dtype=torch.bfloat16, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(40)
    for b in range(B):
        q[b].normal_(generator=g)
    k, v = _rand((B, Hkv, N, d), torch.bfloat16, 41), _rand((B, Hkv, N, d), torch.bfloat16, 42)
    assert q.numel() >= 2 ** 32      # the last (batch, head) slices start beyond element 2^32 - 2^22
    out = torch.empty_like(q)
    # (a) softmax weights sum to one everywhere: V = 1 -> O = 1
    fa_b200.attention_forward(q, k, torch.ones_like(v), causal=True, out=out)
    for b in range(B):
        assert (out[b].float() - 1).abs().max().item() <= 1e-2, f"batch {b}"
    # (b) the real V: row 0 of every head sees only key 0 -> O[.., 0, :] == V[.., 0, :] bit for bit
    _, lse = fa_b200.attention_forward(q, k, v, causal=True, return_lse=True, out=out)
    torch.cuda.synchronize()
    assert torch.equal(out[:, :, 0], v[:, :, 0].repeat_interleave(Hq // Hkv, dim=1))
    assert torch.isfinite(lse).all()
    # (c) sampled row blocks against the oracle, including the very last (batch 15, head 63: offsets > 2^32)
    for (b, h, r0) in ((B - 1, Hq - 1, N - 256), (B - 1, Hq - 1, 5000), (8, 1, N // 2 - 128), (0, 0, 0)):
        hk, nk = h // (Hq // Hkv), r0 + 256
        o_ref, l_ref = oracle.attention_fwd(q[b:b + 1, h:h + 1, r0:r0 + 256].float().cpu().numpy(), k[b:b + 1, hk:hk + 1, :nk].float().cpu().numpy(),
                                            v[b:b + 1, hk:hk + 1, :nk].float().cpu().numpy(), causal=True, return_lse=True)
        err = np.abs(out[b:b + 1, h:h + 1, r0:r0 + 256].float().cpu().numpy() - o_ref).max()
        assert err <= TOL16, f"(b={b}, h={h}, rows {r0}..): max abs error {err:.3e}"
        np.testing.assert_allclose(lse[b:b + 1, h:h + 1, r0:r0 + 256].cpu().numpy(), l_ref, rtol=0, atol=2e-3)


# ---- ring-KV at world size 1 (the default CUDA operators), against the oracle ------------------------------------
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("hq,hkv,n,d", [(4, 2, 1024, 128), (6, 6, 640, 64), (16, 4, 2560, 128)])
def test_ring_attention_world1_vs_oracle(causal, hq, hkv, n, d):
    # one rank owns the whole sequence: a single fa_fwd_carry_window launch into the fp32 running pair + fa_cast_out
    q, k, v = _rand((2, hq, n, d), torch.bfloat16, 50), _rand((2, hkv, n, d), torch.bfloat16, 51), _rand((2, hkv, n, d), torch.bfloat16, 52)
    before = fa_b200.launch_count()
    out, lse = sharding.ring_attention(q, k, v, causal=causal, return_lse=True)
    torch.cuda.synchronize()
    assert fa_b200.launch_count() - before == 2      # one carry launch + one cast, both through libfa_b200.so
    o_ref, l_ref = oracle.attention_fwd(q.float().cpu().numpy(), k.float().cpu().numpy(), v.float().cpu().numpy(), causal=causal, return_lse=True)
    assert np.abs(out.float().cpu().numpy() - o_ref).max() <= TOL16
    np.testing.assert_allclose(lse.cpu().numpy(), l_ref, rtol=0, atol=2e-3)


# ---- launch-path state ------------------------------------------------------------------------------------------
def test_thousands_of_launches_across_two_streams():
    # the work-item counter of the persistent kernel is per stream and self-resetting: 2 x 1,300 launches queued on two
    # streams without a synchronise in between (more than any fixed pool of per-launch counters) all give the same bits
    q, k, v = (_rand((2, 3, 700, 128), torch.bfloat16, 60 + i) for i in range(3))     # 18 items
    ref = fa_b200.attention_forward(q, k, v, causal=True)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    outs = {s1: [torch.empty_like(ref) for _ in range(4)], s2: [torch.empty_like(ref) for _ in range(4)]}
    for i in range(1300):
        for st in (s1, s2):
            with torch.cuda.stream(st):
                fa_b200.attention_forward(q, k, v, causal=True, out=outs[st][i % 4])
    torch.cuda.synchronize()
    assert all(torch.equal(o, ref) for st in (s1, s2) for o in outs[st])


def test_graph_replay_beside_eager_launches():
    # a launch recorded into a CUDA graph owns its work-item counter: replays on another stream run beside eager launches
    # on the stream it was captured from
    q, k, v = (_rand((4, 8, 1024, 128), torch.bfloat16, 70 + i) for i in range(3))    # 128 items
    ref = fa_b200.attention_forward(q, k, v, causal=True)
    o_graph, o_eager = torch.empty_like(ref), torch.empty_like(ref)
    cap = torch.cuda.Stream()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(cap):
        fa_b200.attention_forward(q, k, v, causal=True, out=o_eager)      # this stream's eager counter exists before the capture
        cap.synchronize()
        with torch.cuda.graph(g, stream=cap):
            fa_b200.attention_forward(q, k, v, causal=True, out=o_graph)
    other = torch.cuda.Stream()
    for _ in range(20):
        with torch.cuda.stream(other):
            g.replay()
        with torch.cuda.stream(cap):
            fa_b200.attention_forward(q, k, v, causal=True, out=o_eager)
    torch.cuda.synchronize()
    assert torch.equal(o_graph, ref) and torch.equal(o_eager, ref)


def test_fa_fwd_host_from_two_threads():
    # fa_fwd_host keeps its staging buffers per device behind a lock: two host threads calling it at once both get
    # the device path's bits
    cases = []
    for i in range(2):
        g = torch.Generator().manual_seed(80 + i)
        q = torch.randn(3, 8, 600 + 100 * i, 128, generator=g).to(torch.bfloat16)
        k = torch.randn(3, 2, 600 + 100 * i, 128, generator=g).to(torch.bfloat16)
        v = torch.randn(3, 2, 600 + 100 * i, 128, generator=g).to(torch.bfloat16)
        cases.append((q.pin_memory(), k.pin_memory(), v.pin_memory(), torch.empty_like(q).pin_memory()))
    errs = []

    def work(c):
        try:
            torch.cuda.set_device(0)
            for _ in range(5):
                fa_b200.attention_forward_host(c[0], c[1], c[2], c[3], causal=True)
        except Exception as ex:   # surfaced below
            errs.append(ex)

    th = [threading.Thread(target=work, args=(c,)) for c in cases]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    for (q, k, v, o) in cases:
        assert torch.equal(o, fa_b200.attention_forward(q.cuda(), k.cuda(), v.cuda(), causal=True).cpu())
