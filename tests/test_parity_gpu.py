"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI (fa_b200.py -> libfa_b200.so), against the
CPU oracle on identical seeded inputs, against the reference's golden vectors, and — at BASELINE.json's full
sizes — through size-independent properties.

Tolerances (BASELINE.json north_star): max abs error <= 2e-2 for bf16 / fp16, <= 1e-4 relative for fp32,
always against an fp32(+) softmax on the same (already rounded) inputs.
"""
import glob
import os
import subprocess

import numpy as np
import pytest
import torch

import fa_b200
from oracle import oracle

pytestmark = pytest.mark.gpu

TOL16 = 2e-2     # max abs, bf16 / fp16
RTOL32 = 1e-4    # relative, fp32
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "flash-attention-cuda-c_b200")


@pytest.fixture(scope="module", autouse=True)
def _lib():
    assert torch.cuda.is_available(), "GPU tests need a B200"
    assert os.path.exists(fa_b200.LIB_PATH), "libfa_b200.so missing: the CUDA path must be built, there is no fallback"
    fa_b200.lib()
    before = fa_b200.launch_count()
    yield
    assert fa_b200.launch_count() > before, "no kernel of libfa_b200.so was launched by these tests"


def _inputs(B, Hq, Hkv, Nq, Nk, d, dtype, seed=0):
    # seeds 0/1/2 for Q/K/V, N(0,1) in fp32 then rounded to the I/O dtype
    gen = [torch.Generator().manual_seed(seed + i) for i in range(3)]
    q = torch.randn(B, Hq, Nq, d, generator=gen[0]).to(dtype)
    k = torch.randn(B, Hkv, Nk, d, generator=gen[1]).to(dtype)
    v = torch.randn(B, Hkv, Nk, d, generator=gen[2]).to(dtype)
    return q, k, v


def _check(q, k, v, causal, scale=None, lse=True):
    o_ref, lse_ref = oracle.attention_fwd(q.float().numpy(), k.float().numpy(), v.float().numpy(), causal=causal,
                                          scale=scale, return_lse=True)
    out = fa_b200.attention_forward(q.cuda(), k.cuda(), v.cuda(), causal=causal, scale=scale, return_lse=lse)
    torch.cuda.synchronize()
    o, l = (out if lse else (out, None))
    o = o.float().cpu().numpy()
    assert np.isfinite(o).all()
    if q.dtype == torch.float32:
        err = np.abs(o - o_ref).max() / max(np.abs(o_ref).max(), 1e-30)
        assert err <= RTOL32, f"fp32 relative error {err:.3e}"
        np.testing.assert_allclose(o, o_ref, rtol=RTOL32, atol=2e-5)
    else:
        err = np.abs(o - o_ref).max()
        assert err <= TOL16, f"max abs error {err:.3e} > {TOL16}"
    if l is not None:
        l = l.cpu().numpy()
        fin = np.isfinite(lse_ref)
        assert (np.isfinite(l) == fin).all()
        np.testing.assert_allclose(l[fin], lse_ref[fin], rtol=0, atol=2e-3 if q.dtype != torch.float32 else 1e-4)
    return err


# ---- 16-bit tensor-core path ---------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("d", [128, 64])
def test_16bit_basic(dtype, causal, d):
    _check(*_inputs(2, 3, 3, 512, 512, d, dtype), causal=causal)


@pytest.mark.parametrize("n", [1, 17, 127, 128, 129, 200, 255, 256, 257, 333, 640, 1000])
@pytest.mark.parametrize("causal", [False, True])
def test_16bit_ragged_lengths(n, causal):
    # tile-boundary and ragged shapes: TMA zero fill + masking of the tail
    _check(*_inputs(1, 2, 2, n, n, 128, torch.bfloat16, seed=n), causal=causal)


@pytest.mark.parametrize("nq,nk", [(128, 512), (100, 1000), (512, 128), (300, 200), (1, 777)])
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("d", [64, 128])
def test_16bit_nq_ne_nk(nq, nk, causal, d):
    # bottom-right aligned causal mask; with Nq > Nk the first rows see no key -> zeros, lse = -inf.  Both head dims: the
    # MMA issuers are split by query tile at d = 64 and by type at d = 128, and a query tile without any visible key tile
    # (virtual steps on the shared score buffer) takes a different path in each
    _check(*_inputs(1, 2, 2, nq, nk, d, torch.bfloat16, seed=nq + nk), causal=causal)


@pytest.mark.parametrize("hq,hkv", [(8, 1), (8, 2), (6, 3), (64, 8)])
def test_16bit_gqa(hq, hkv):
    _check(*_inputs(2, hq, hkv, 384, 384, 128, torch.bfloat16, seed=hq), causal=True)


@pytest.mark.parametrize("seed", range(24))
def test_16bit_random_shapes(seed):
    # randomised sweep over batch, heads (GQA ratios), ragged Nq != Nk, head dim, dtype, causal, scale
    rng = np.random.default_rng(1000 + seed)
    hkv = int(rng.choice([1, 2, 3]))
    hq = hkv * int(rng.choice([1, 2, 4]))
    nq, nk = int(rng.integers(1, 700)), int(rng.integers(1, 900))
    d = int(rng.choice([64, 128]))
    dtype = [torch.bfloat16, torch.float16][int(rng.integers(0, 2))]
    causal = bool(rng.integers(0, 2))
    scale = None if rng.integers(0, 2) else float(rng.uniform(0.05, 0.3))
    _check(*_inputs(int(rng.integers(1, 4)), hq, hkv, nq, nk, d, dtype, seed=seed), causal=causal, scale=scale)


def test_concurrent_streams_do_not_share_work_counters():
    # two persistent launches in flight on different streams must not steal each other's work items
    q, k, v = _inputs(2, 8, 8, 1024, 1024, 128, torch.bfloat16, seed=21)
    qc, kc, vc = q.cuda(), k.cuda(), v.cuda()
    ref = fa_b200.attention_forward(qc, kc, vc, causal=True)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    outs = []
    for _ in range(6):
        for st in (s1, s2):
            with torch.cuda.stream(st):
                outs.append(fa_b200.attention_forward(qc, kc, vc, causal=True))
    torch.cuda.synchronize()
    assert all(torch.equal(o, ref) for o in outs)


def test_cuda_graph_capture_and_replay():
    # launch-bound shapes belong in a CUDA graph: fa_fwd is capturable (counter memset + one kernel per call, tensor maps
    # baked in as kernel parameters) and a replay on new data in the same buffers gives the eager result bit for bit
    q, k, v = (t.cuda() for t in _inputs(2, 4, 2, 512, 512, 128, torch.bfloat16, seed=31))
    o = torch.empty_like(q)
    o2 = torch.empty_like(q)
    fa_b200.attention_forward(q, k, v, causal=True, out=o)          # warm-up outside the capture (smem attribute, counter pool)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fa_b200.attention_forward(q, k, v, causal=True, out=o)
        fa_b200.attention_forward(q, k, v, causal=False, out=o2)
    for seed in (32, 33):
        qn, kn, vn = _inputs(2, 4, 2, 512, 512, 128, torch.bfloat16, seed=seed)
        q.copy_(qn); k.copy_(kn); v.copy_(vn)
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(o, fa_b200.attention_forward(q, k, v, causal=True))
        assert torch.equal(o2, fa_b200.attention_forward(q, k, v, causal=False))
        ref = oracle.attention_fwd(qn.float().numpy(), kn.float().numpy(), vn.float().numpy(), causal=True)
        assert np.abs(o.float().cpu().numpy() - ref).max() <= TOL16


def test_16bit_long_rows_trigger_rescale():
    # growing score magnitude along the key axis forces the lazy O rescale path (max grows by > 2^8 repeatedly)
    q, k, v = _inputs(1, 2, 2, 256, 2048, 128, torch.bfloat16, seed=11)
    ramp = torch.linspace(0.05, 4.0, 2048)[None, None, :, None]
    k = (k.float() * ramp).to(torch.bfloat16)
    q = (q.float() * 3).to(torch.bfloat16)
    _check(q, k, v, causal=False)
    _check(q, k, v, causal=False, scale=0.5)


def test_16bit_all_ones_known_answer():
    # the reference's KAT generalised: all ones in -> all ones out (tests/main.cu:33-35)
    x = torch.ones(1, 2, 300, 128, dtype=torch.bfloat16)
    for causal in (False, True):
        o = fa_b200.attention_forward(x.cuda(), x.cuda(), x.cuda(), causal=causal).float().cpu()
        assert torch.allclose(o, torch.ones_like(o), atol=1e-2)


@pytest.mark.parametrize("dtype,d", [(torch.bfloat16, 128), (torch.float16, 64), (torch.float32, 64)])
@pytest.mark.parametrize("nq,nk", [(333, 333), (1, 129), (257, 100)])
def test_no_write_outside_the_output_rows(dtype, d, nq, nk):
    # compute-sanitizer is not available on this pool, so the bounds check is our own: O is a view of rows
    # [3, 3+Nq) inside a sentinel-filled buffer (ragged Nq: the last 128-row tile hangs over the end), LSE sits between
    # two sentinel bands; nothing outside the view may change and everything inside must be overwritten.
    B, H = 2, 3
    q, k, v = (t.cuda() for t in _inputs(B, H, H, nq, nk, d, dtype, seed=nq + d))
    sent = 512.0
    big = torch.full((B, H, nq + 8, d), sent, dtype=dtype, device="cuda")
    out = big[:, :, 3:3 + nq, :]
    lse_big = torch.full((B * H * nq + 64,), sent, dtype=torch.float32, device="cuda")
    strides = (fa_b200.ctypes.c_longlong * 12)(*(list(q.stride()[:3]) + list(k.stride()[:3]) + list(v.stride()[:3]) + list(out.stride()[:3])))
    for causal in (0, 1):
        big.fill_(sent)
        lse_big.fill_(sent)
        rc = fa_b200.lib().fa_fwd_strided(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), lse_big.data_ptr() + 32 * 4,
                                          B, H, H, nq, nk, d, fa_b200._dtype_code(q), 0.0, causal, strides, fa_b200._stream_ptr(q))
        assert rc == 0
        torch.cuda.synchronize()
        assert (big[:, :, :3] == sent).all() and (big[:, :, 3 + nq:] == sent).all(), "write outside the output rows"
        assert (lse_big[:32] == sent).all() and (lse_big[32 + B * H * nq:] == sent).all(), "write outside the LSE range"
        o_ref, lse_ref = oracle.attention_fwd(q.float().cpu().numpy(), k.float().cpu().numpy(), v.float().cpu().numpy(),
                                              causal=bool(causal), return_lse=True)
        tol = TOL16 if dtype != torch.float32 else RTOL32 * max(np.abs(o_ref).max(), 1.0)
        assert np.abs(out.float().cpu().numpy() - o_ref).max() <= tol
        got = lse_big[32:32 + B * H * nq].view(B, H, nq).cpu().numpy()
        fin = np.isfinite(lse_ref)
        assert (np.isfinite(got) == fin).all()      # every LSE slot was written (no sentinel left)
        np.testing.assert_allclose(got[fin], lse_ref[fin], rtol=0, atol=2e-3)


@pytest.mark.parametrize("col_off", [0, 8])
def test_output_view_alignment_paths(col_off):
    # the epilogue uses 256-bit stores when O is 32-byte aligned and 128-bit stores otherwise: an output view that starts
    # 16 bytes into a wider, sentinel-filled buffer takes the second path; the columns beside the view must stay untouched
    B, H, n, d = 1, 3, 300, 128
    q, k, v = (t.cuda() for t in _inputs(B, H, H, n, n, d, torch.bfloat16, seed=77))
    big = torch.full((B, H, n, d + 16), 512.0, dtype=torch.bfloat16, device="cuda")
    out = big[..., col_off:col_off + d]
    assert out.data_ptr() % 32 == (16 if col_off else 0)
    fa_b200.attention_forward(q, k, v, causal=True, out=out)
    torch.cuda.synchronize()
    assert (big[..., :col_off] == 512.0).all() and (big[..., col_off + d:] == 512.0).all()
    o_ref = oracle.attention_fwd(q.float().cpu().numpy(), k.float().cpu().numpy(), v.float().cpu().numpy(), causal=True)
    assert np.abs(out.float().cpu().numpy() - o_ref).max() <= TOL16


@pytest.mark.parametrize("causal", [False, True])
def test_nan_in_one_query_row_stays_in_that_row(causal):
    # NaN / Inf guard: a poisoned query row yields a non-finite output row (as in check.py's softmax) and must neither hang
    # the barrier protocol nor leak into any other row of the tile
    q, k, v = _inputs(1, 2, 2, 300, 300, 128, torch.bfloat16, seed=5)
    q[0, 1, 130, 7] = float("nan")
    q[0, 0, 5, :] = float("inf")
    o = fa_b200.attention_forward(q.cuda(), k.cuda(), v.cuda(), causal=causal).float().cpu().numpy()
    bad = np.zeros((1, 2, 300), dtype=bool)
    bad[0, 1, 130] = bad[0, 0, 5] = True
    assert not np.isfinite(o[bad]).any()
    qc = q.clone()
    qc[0, 1, 130, 7] = 0.0
    qc[0, 0, 5, :] = 0.0
    o_ref = oracle.attention_fwd(qc.float().numpy(), k.float().numpy(), v.float().numpy(), causal=causal)
    assert np.abs(o[~bad] - o_ref[~bad]).max() <= TOL16


def test_config2_gpt2_shape_fp16():
    # BASELINE.json configs[1]: B=4 H=12 N=1024 d=64 non-causal fp16
    _check(*_inputs(4, 12, 12, 1024, 1024, 64, torch.float16), causal=False)


def test_check_py_layout_strided(golden_dir):
    # check.py's (batch, seq, d_model) layout through fa_fwd_strided, against the reference's golden outputs:
    # head dim 128 on the bf16 tensor-core path, head dim 32 on the fp32 path
    g = np.load(os.path.join(golden_dir, "mh_b1_n384_h2_d128.npz"))
    h = int(g["num_heads"])
    Q, K, V = (torch.from_numpy(g[n]).to(torch.bfloat16) for n in ("Q", "K", "V"))
    out = fa_b200.multi_head_attention(Q.cuda(), K.cuda(), V.cuda(), h).float().cpu()
    ref, _ = oracle.multi_head_attention(Q.float(), K.float(), V.float(), h)
    assert (out - ref).abs().max().item() <= TOL16
    # and against the reference's own fp32 output on unrounded inputs (bf16 input rounding included)
    assert np.abs(out.numpy() - g["output"]).max() <= 5e-2
    g = np.load(os.path.join(golden_dir, "mh_b2_n200_h4_d32.npz"))
    Q, K, V = (torch.from_numpy(g[n]).cuda() for n in ("Q", "K", "V"))
    out = fa_b200.multi_head_attention(Q, K, V, int(g["num_heads"])).cpu().numpy()
    assert np.abs(out - g["output"]).max() / np.abs(g["output"]).max() <= RTOL32


# ---- fp32 path ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("shape", [(1, 1, 256, 64), (2, 3, 100, 32), (1, 2, 77, 128), (1, 1, 16, 16), (1, 4, 513, 80)])
def test_fp32_path(shape, causal):
    B, H, N, d = shape
    _check(*_inputs(B, H, H, N, N, d, torch.float32, seed=N), causal=causal)


def test_fp32_gqa_and_nq_ne_nk():
    _check(*_inputs(1, 4, 2, 70, 130, 64, torch.float32), causal=True)
    _check(*_inputs(1, 4, 1, 130, 70, 64, torch.float32), causal=True)


def test_config1_reference_golden_fp32(golden_dir):
    # BASELINE.json configs[0]: B=1 H=1 N=256 d=64 fp32, output of the reference's check.py
    g = np.load(os.path.join(golden_dir, "cfg1_b1_n256_h1_d64.npz"))
    Q, K, V = (torch.from_numpy(g[n]).cuda() for n in ("Q", "K", "V"))
    out = fa_b200.multi_head_attention(Q, K, V, int(g["num_heads"])).cpu().numpy()
    err = np.abs(out - g["output"]).max() / np.abs(g["output"]).max()
    assert err <= RTOL32, err
    x = torch.ones(1, 1, 16, 16).cuda()
    o = torch.empty_like(x)
    fa_b200.two_loader_mha_flash_attention(x, x, x, o, 1, 1, 16, 0.25, False)   # the KAT through the kernel's argument list
    assert torch.allclose(o.cpu(), torch.ones(1, 1, 16, 16), atol=1e-6)


# ---- error behaviour on the device ----------------------------------------------------------------------
def test_errors_are_return_codes():
    x = torch.zeros(1, 1, 64, 96, dtype=torch.bfloat16).cuda()
    with pytest.raises(fa_b200.FaError, match="supports d in"):
        fa_b200.attention_forward(x, x, x)
    y = torch.zeros(1, 1, 64, 24, dtype=torch.float32).cuda()
    with pytest.raises(fa_b200.FaError):
        fa_b200.attention_forward(y, y, y)


# ---- compat entry point + stage probes (native binaries) ------------------------------------------------
def _run(path):
    assert os.path.exists(path), f"{path} not built"
    r = subprocess.run([path], capture_output=True, text=True, timeout=120)
    print(r.stdout[-3000:], r.stderr[-2000:])
    return r


def test_compat_template_under_reference_launch_contract():
    r = _run(os.path.join(PKG, "tests", "compat_main"))
    assert r.returncode == 0 and "COMPAT PASSED" in r.stdout


def test_stage_probes_tma_umma_tmem():
    r = _run(os.path.join(PKG, "tests", "probe_umma"))
    assert r.returncode == 0 and "PROBE PASSED" in r.stdout


def test_stage_probe_softmax_step():
    # one online-softmax step on a known score tile, in the 16-lane TMEM view the kernel uses, against a host restatement
    # (the slot of the reference's empty tests/test_computers.cu)
    r = _run(os.path.join(PKG, "tests", "probe_softmax"))
    assert r.returncode == 0 and "SOFTMAX PROBE PASSED" in r.stdout


def test_reference_own_test_runs_against_our_headers():
    # the reference's tests/main.cu, unmodified, compiled against our kernels/ (oracle/_ref/ref_test_dropin)
    path = os.path.join(ROOT, "oracle", "_ref", "ref_test_dropin")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/ref_test_dropin not built (reference sources absent at build time)")
    r = _run(path)
    assert r.returncode == 0
    assert "Max absolute difference vs CPU reference: 0.000000" in r.stdout


# ---- host-buffer end-to-end call ------------------------------------------------------------------------
def test_fa_fwd_host_matches_device_path():
    q, k, v = _inputs(3, 8, 2, 700, 700, 128, torch.bfloat16, seed=5)
    qp, kp, vp = (t.pin_memory() for t in (q, k, v))
    out = torch.empty_like(q).pin_memory()
    fa_b200.attention_forward_host(qp, kp, vp, out, causal=True)
    dev = fa_b200.attention_forward(q.cuda(), k.cuda(), v.cuda(), causal=True).cpu()
    assert torch.equal(out, dev)


# ---- ring building block ---------------------------------------------------------------------------------
def test_merge_partial_equals_full_attention():
    q, k, v = _inputs(1, 2, 2, 256, 1024, 128, torch.bfloat16, seed=9)
    qc, kc, vc = q.cuda(), k.cuda(), v.cuda()
    acc_o = torch.zeros(1, 2, 256, 128, device="cuda")
    acc_l = torch.full((1, 2, 256), float("-inf"), device="cuda")
    for s in range(0, 1024, 256):
        po, pl = fa_b200.attention_forward(qc, kc[:, :, s:s + 256].contiguous(), vc[:, :, s:s + 256].contiguous(), return_lse=True)
        fa_b200.merge_partial(acc_o, acc_l, po, pl)
    o_ref, lse_ref = oracle.attention_fwd(q.float().numpy(), k.float().numpy(), v.float().numpy(), return_lse=True)
    assert np.abs(acc_o.cpu().numpy() - o_ref).max() <= TOL16
    np.testing.assert_allclose(acc_l.cpu().numpy(), lse_ref, atol=2e-3)
    out16 = fa_b200.cast_out(acc_o, torch.empty(1, 2, 256, 128, dtype=torch.bfloat16, device="cuda"))
    assert np.abs(out16.float().cpu().numpy() - o_ref).max() <= TOL16


def test_carry_mode_equals_full_attention():
    # fa_fwd_carry: the merge fused into the kernel epilogue, over strided key slices, causal last block included
    q, k, v = _inputs(2, 4, 2, 384, 1024, 128, torch.bfloat16, seed=13)
    qc, kc, vc = q.cuda(), k.cuda(), v.cuda()
    for causal in (False, True):
        acc_o = torch.zeros(2, 4, 384, 128, device="cuda")
        acc_l = torch.full((2, 4, 384), float("-inf"), device="cuda")
        bounds = [0, 256, 640, 1024]
        for i, (s, e) in enumerate(zip(bounds[:-1], bounds[1:])):
            last = i == len(bounds) - 2
            # only the last key block (which ends at Nk) carries the bottom-right aligned causal mask
            fa_b200.attention_forward_carry(qc, kc[:, :, s:e], vc[:, :, s:e], acc_o, acc_l, causal=causal and last)
        if causal:   # reference: keys < 640 fully visible, keys 640.. causal with offset (1024-640) - 384 = 0
            kk = k.float().numpy(); vv = v.float().numpy(); qq = q.float().numpy()
            o1, l1 = oracle.attention_fwd(qq, kk[:, :, :640], vv[:, :, :640], return_lse=True)
            o2, l2 = oracle.attention_fwd(qq, kk[:, :, 640:], vv[:, :, 640:], causal=True, return_lse=True)
            lse_ref = np.logaddexp(l1, l2)
            o_ref = o1 * np.exp(l1 - lse_ref)[..., None] + o2 * np.exp(l2 - lse_ref)[..., None]
        else:
            o_ref, lse_ref = oracle.attention_fwd(q.float().numpy(), k.float().numpy(), v.float().numpy(), return_lse=True)
        assert np.abs(acc_o.cpu().numpy() - o_ref).max() <= TOL16
        np.testing.assert_allclose(acc_l.cpu().numpy(), lse_ref, atol=2e-3)


# ---- full-size properties (BASELINE.json configs[2]: B=8 H=32 N=8192 d=128 causal bf16) ------------------
def test_config3_full_size_properties():
    B, H, N, d = 8, 32, 8192, 128
    g = torch.Generator(device="cuda").manual_seed(0)
    q = torch.randn(B, H, N, d, device="cuda", generator=g).to(torch.bfloat16)
    k = torch.randn(B, H, N, d, device="cuda", generator=g).to(torch.bfloat16)
    v1 = torch.randn(B, H, N, d, device="cuda", generator=g).to(torch.bfloat16)
    # (a) softmax weights sum to one: V = 1 -> O = 1
    ones = torch.ones_like(v1)
    o = fa_b200.attention_forward(q, k, ones, causal=True)
    assert (o.float() - 1).abs().max().item() <= 1e-2
    del ones, o
    # (b) row 0 of every head sees only key 0 -> O[.., 0, :] == V[.., 0, :] exactly
    o1, lse = fa_b200.attention_forward(q, k, v1, causal=True, return_lse=True)
    assert torch.equal(o1[:, :, 0], v1[:, :, 0])
    assert torch.isfinite(o1.float()).all() and torch.isfinite(lse).all()
    # (c) linearity in V
    v2 = torch.randn(B, H, N, d, device="cuda", generator=g).to(torch.bfloat16)
    o2 = fa_b200.attention_forward(q, k, v2, causal=True)
    o12 = fa_b200.attention_forward(q, k, (v1.float() + v2.float()).to(torch.bfloat16), causal=True)
    assert (o12.float() - (o1.float() + o2.float())).abs().max().item() <= 6e-2
    # (d) two (batch, head) slices against the oracle at full length
    for (b, h) in ((0, 0), (7, 31)):
        o_ref = oracle.attention_fwd(q[b:b + 1, h:h + 1].float().cpu().numpy(), k[b:b + 1, h:h + 1].float().cpu().numpy(),
                                     v1[b:b + 1, h:h + 1].float().cpu().numpy(), causal=True)
        assert np.abs(o1[b:b + 1, h:h + 1].float().cpu().numpy() - o_ref).max() <= TOL16
    # (e) determinism
    assert torch.equal(o1, fa_b200.attention_forward(q, k, v1, causal=True))
