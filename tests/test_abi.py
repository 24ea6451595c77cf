"""CPU tests (-m "not gpu"): the C-ABI library builds for sm_100a, loads without a GPU, exports every symbol that
include/fa_b200.h declares, and its host-only helpers behave like the reference's helpers.hpp contracts."""
import ctypes
import os
import re

import pytest

import fa_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    fa_b200.build()
    return fa_b200.lib()


def _declared():
    hdr = open(os.path.join(ROOT, "include", "fa_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(fa_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_exported(L):
    names = _declared()
    assert len(names) >= 12
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/fa_b200.h but not exported"
    assert sorted(fa_b200.EXPORTS) == names


def test_library_is_sm100a_native():
    import subprocess
    sass = subprocess.run(["cuobjdump", "-sass", fa_b200.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM", "STTM", "UTCBAR"):
        assert mnemonic in sass, f"{mnemonic} missing from SASS: the tcgen05/TMA path did not compile"
    assert "HMMA." not in sass.replace("UTCHMMA", "")   # no legacy mma.sync path


def test_host_helpers(L):
    # getNumCta (helpers.hpp:33-36) without the divisibility assert: ragged sizes round up
    assert L.fa_num_cta(8192, 256) == 32 and L.fa_num_cta(8193, 256) == 33 and L.fa_num_cta(0, 256) == 0
    # calculateSizeBlockQ / KV (helpers.hpp:8-30): the sm_100a tile table
    assert L.fa_block_q(128, fa_b200.FA_DTYPE_BF16) == 256 and L.fa_block_kv(128, fa_b200.FA_DTYPE_BF16) == 128
    assert L.fa_block_q(64, fa_b200.FA_DTYPE_F32) == 64
    assert b"sm_100a" in L.fa_version()
    # the measured tile table behind them: every (d, causal) has a base row, rows are consistent with the built kernels,
    # and fa_choose_tile answers with the row of the largest n_min that fits
    rows = fa_b200.tile_table()
    assert {(r["d"], r["causal"]) for r in rows if r["n_min"] == 0} == {(128, 0), (128, 1), (64, 0), (64, 1)}
    for r in rows:
        assert r["block_q"] == 256 and r["block_kv"] == 128 and r["softmax_warps"] in (8, 16) and r["cta_group"] in (1, 2)
        assert r["issuer_by_type"] == (1 if r["d"] == 128 else 0) and r["tflops"] > 0
        # compiled variants: (8 warps, direct), (8 warps, staged TMA-store epilogue: 4 ring slots at d = 128), (16 warps + FMA-pipe exp2);
        # the CTA-pair kernel (cta_group 2: d = 128, 8 warps, six half-tile ring slots) exists with either epilogue
        assert (r["softmax_warps"], r["emu_pairs_per_8"], r["staged_epilogue"]) in ((8, 0, 0), (8, 0, 1), (16, 1, 0))
        if r["cta_group"] == 2:
            assert r["d"] == 128 and r["softmax_warps"] == 8 and r["stages"] == 6
        else:
            assert r["stages"] == (8 if r["d"] == 64 else (4 if r["staged_epilogue"] else 5))
    for d in (64, 128):
        for causal in (0, 1):
            for nk in (100, 1024, 4096, 100000):
                got = fa_b200.choose_tile(d, fa_b200.FA_DTYPE_BF16, causal, nk, nk)
                want = max((r for r in rows if r["d"] == d and r["causal"] == causal and nk >= r["n_min"]), key=lambda r: r["n_min"])
                assert got == want
    assert fa_b200.choose_tile(64, fa_b200.FA_DTYPE_F32, 0, 10, 10)["block_q"] == 64
    with pytest.raises(fa_b200.FaError):
        fa_b200.choose_tile(96, fa_b200.FA_DTYPE_BF16, 0, 10, 10)


def test_argument_validation_without_gpu(L):
    # validation happens before any CUDA call, so it can be exercised on a CPU-only box
    assert L.fa_fwd(None, None, None, None, None, 1, 1, 1, 16, 16, 64, 2, 0.0, 0, None) == -1
    assert b"null" in L.fa_last_error()
    buf = ctypes.create_string_buffer(64)
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert L.fa_fwd(p, p, p, p, None, 1, 3, 2, 16, 16, 64, 2, 0.0, 0, None) == -1     # Hq % Hkv != 0
    assert L.fa_fwd(p, p, p, p, None, 1, 1, 1, 0, 16, 64, 2, 0.0, 0, None) == -1      # empty sequence
    assert L.fa_fwd(p, p, p, p, None, 1, 1, 1, 16, 16, 64, 7, 0.0, 0, None) == -1     # unknown dtype
    assert L.fa_merge_partial(None, None, None, None, 0, 0, 2, None) == -1
    # "ask, allocate, call": the answer is always 0 bytes, and the same shapes fa_fwd rejects are rejected here
    assert L.fa_workspace_bytes(8, 32, 32, 8192, 8192, 128, fa_b200.FA_DTYPE_BF16) == 0
    assert L.fa_workspace_bytes(16, 64, 8, 32768, 32768, 128, fa_b200.FA_DTYPE_F16) == 0
    assert L.fa_workspace_bytes(1, 1, 1, 256, 256, 64, fa_b200.FA_DTYPE_F32) == 0
    assert L.fa_workspace_bytes(1, 3, 2, 16, 16, 64, 2) == -1 and L.fa_workspace_bytes(1, 1, 1, 16, 16, 96, 2) < 0
    assert L.fa_launch_count() == 0                                                    # nothing was launched


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(fa_b200, "_lib", None)
    monkeypatch.setattr(fa_b200, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(fa_b200.FaError):
        fa_b200.lib()


def test_no_cpu_fallback_for_cpu_tensors(L):
    import torch
    x = torch.zeros(1, 1, 16, 64, dtype=torch.bfloat16)
    with pytest.raises(fa_b200.FaError):
        fa_b200.attention_forward(x, x, x)


def test_product_does_not_touch_oracle():
    pkg = os.path.join(ROOT, "flash-attention-cuda-c_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp", ".h", ".sh")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                for pat in (r"^\s*(from|import)\s+oracle", r"liboracle", r"oracle/", r"oracle\.", r"attention_oracle"):
                    assert not re.search(pat, txt, flags=re.M), f"{f} uses the oracle ({pat}): the product path must not"


def test_work_item_plan_of_small_launches(L):
    # the tail of a small launch: whole waves of 256-row items, and a remainder that would leave more than half of the SMs idle
    # is queued as 128-row half items (two per remaining block).  Every block is covered exactly once either way.
    ll = ctypes.c_longlong
    L.fa_debug_plan_counts.argtypes = [ll, ctypes.c_int, ctypes.POINTER(ll), ctypes.POINTER(ll)]
    def plan(blocks, ctas=148):
        a, b = ll(), ll()
        assert L.fa_debug_plan_counts(blocks, ctas, ctypes.byref(a), ctypes.byref(b)) == 0
        return a.value, b.value
    assert plan(192) == (148, 148 + 2 * 44)          # BASELINE configs[1]: 4 x 12 heads x 4 query blocks
    assert plan(148) == (148, 148) and plan(296) == (296, 296)      # whole waves: nothing to split
    assert plan(100) == (100, 100)                   # a single partial wave with more than half of the SMs busy stays whole
    assert plan(40) == (0, 80)                       # a handful of blocks: all halves, twice as many SMs busy
    assert plan(148 + 75) == (223, 223)              # remainder above half a wave: a full-item wave is cheaper than two half-item rounds
    assert plan(8192) == (8192, 8192)                # many waves (config 3): left alone
    for blocks in range(1, 700):
        n_full, total = plan(blocks)
        assert 0 <= n_full <= blocks and total == n_full + 2 * (blocks - n_full)
        assert n_full % 148 == 0 or n_full == blocks


def test_fast_division_of_the_item_decode(L):
    # every role decodes every work item: the three divisions by run-time values are multiply + shift with host-made magic
    # numbers; exact for all 0 <= n < 2^31
    import random
    L.fa_debug_fast_div.argtypes = [ctypes.c_uint, ctypes.c_uint]
    rng = random.Random(0)
    ds = list(range(1, 300)) + [2 ** k + e for k in range(1, 31) for e in (-1, 0, 1) if 1 <= 2 ** k + e < 2 ** 31] + [rng.randrange(1, 2 ** 31) for _ in range(300)]
    for d in ds:
        for n in [0, 1, d - 1, d, d + 1, 2 * d - 1, 2 * d, 2 ** 31 - 1] + [rng.randrange(0, 2 ** 31) for _ in range(8)]:
            if 0 <= n < 2 ** 31:
                assert L.fa_debug_fast_div(d, n) == n // d, (d, n)


def test_work_item_decode_covers_every_row_once(L):
    # The kernels' item decode (loaders.cuh: decode_item / decode_pair_item, compiled for the host as well) replayed on the CPU
    # for whole launches: every (batch, head, query row) belongs to exactly one query tile of one item — for the 1-CTA kernels
    # with their half-item / split-KV tail, for CTA pairs cut by rows, by two heads and by four heads — and every tile is given
    # exactly the key tiles its rows can see (pairs cut by rows: those of the whole 256-row MMA tile, identical in both CTAs).
    import numpy as np
    ip = ctypes.c_int
    L.fa_debug_decode_items.argtypes = [ip] * 9 + [ctypes.POINTER(ip), ip]
    cases = [(1, 3, 3, 1000, 1000, 1), (2, 4, 2, 700, 1300, 1), (1, 2, 2, 513, 384, 0), (2, 8, 2, 2100, 2100, 1), (1, 4, 1, 300, 2000, 1),
             (3, 6, 3, 129, 640, 0), (1, 2, 2, 1, 1, 1), (2, 2, 2, 900, 500, 1), (4, 12, 12, 1024, 1024, 0), (2, 16, 2, 777, 777, 1)]
    for mode in (0, 1, 2, 3):
        for (B, Hq, Hkv, Nq, Nk, causal) in cases:
            group = Hq // Hkv
            if (mode == 2 and group % 2) or (mode == 3 and group % 4):
                continue
            for split_half in ((0, 1) if mode == 0 else (0,)):
                cap = 8 * B * Hq * (Nq // 128 + 2)
                buf = (ip * (12 * cap))()
                n = L.fa_debug_decode_items(mode, B, Hq, Hkv, Nq, Nk, causal, 148, split_half, buf, cap)
                assert n > 0, (mode, B, Hq, Hkv, Nq, Nk, causal)
                rec = np.frombuffer(buf, dtype=np.int32, count=12 * n).reshape(n, 12)
                cover = np.zeros((B, Hq, Nq), dtype=np.int32)
                n_all = (Nk + 127) // 128
                off = Nk - Nq
                def visible_tiles(last_row):          # key tiles a tile whose last row is last_row can see
                    if not causal:
                        return n_all
                    return 0 if last_row + off < 0 else min(n_all, (last_row + off) // 128 + 1)
                for b, h, hkv, q0, rows, split, n_kv, n_steps, nt0, nt1, stride, rank in rec.tolist():
                    assert hkv == h // group and 0 <= b < B and 0 <= h < Hq
                    hstep = 1 if mode == 3 else 0      # pairs cut by four heads: slot t is head h + t, same rows
                    assert n_kv == max(nt0, nt1) or split
                    if split and q0 >= Nq:             # the second half of a ragged last block: nothing to do
                        assert n_kv == 0 and nt0 == 0 and nt1 == 0
                        continue
                    if split:                          # split-KV half item: both slots work on rows [q0, q0 + 128), keys dealt out alternately
                        assert mode == 0 and rows == 128 and nt0 + nt1 == n_kv == visible_tiles(q0 + 127) and n_steps == nt0 and nt0 - nt1 in (0, 1)
                        cover[b, h, q0:min(q0 + 128, Nq)] += 1
                        continue
                    assert n_steps == n_kv
                    for t, nt in ((0, nt0), (1, nt1)):
                        r0 = q0 + t * stride
                        ht = h + t * hstep
                        assert ht // group == hkv
                        if mode == 0 and t * 128 >= rows:      # half item on slot 0 alone: slot 1 has no rows
                            assert nt == 0
                            continue
                        if r0 >= Nq:
                            # a tile past the end of the sequence does nothing — unless it is the peer half of a pair tile whose
                            # leader half has rows (then it runs in lock-step on zero-filled Q rows and stores nothing)
                            assert nt == 0 or (mode == 1 and rank == 1 and r0 - 128 < Nq)
                            continue
                        cover[b, ht, r0:min(r0 + 128, Nq)] += 1
                        if mode == 1:                          # the MMA tile is 256 rows: leader rows r0.., peer rows r0 + 128..
                            first = r0 - 128 * rank
                            assert nt == visible_tiles(first + 255)
                        else:
                            assert nt == visible_tiles(r0 + 127)
                assert (cover == 1).all(), (mode, B, Hq, Hkv, Nq, Nk, causal, split_half, np.argwhere(cover != 1)[:4])
                if mode:                                   # the two CTAs of a pair walk the same steps
                    a, c = rec[0::2], rec[1::2]
                    assert (a[:, [0, 2, 4, 6, 7, 8, 9]] == c[:, [0, 2, 4, 6, 7, 8, 9]]).all()
                    assert (c[:, 1] - a[:, 1] == {1: 0, 2: 1, 3: 2}[mode]).all() and (c[:, 3] - a[:, 3] == (128 if mode == 1 else 0)).all()
                if mode == 3:                              # both slots of both CTAs reach the diagonal together
                    assert (rec[:, 8] == rec[:, 9]).all() and (rec[:, 10] == 0).all()


def test_launcher_rules_on_top_of_the_tile_table(L):
    # fa_choose_kernel = the table row + the launcher's rules, host arithmetic only (148 SMs assumed without a GPU)
    bf = fa_b200.FA_DTYPE_BF16
    ck = fa_b200.choose_kernel
    # BASELINE configs[2] (MHA 8K causal) and its non-causal twin: CTA pairs cut by rows, 512-row items, six half-tile ring slots
    k = ck(8, 32, 32, 8192, 8192, 128, bf, True)
    assert (k["cta_group"], k["heads_per_item"], k["stages"], k["staged_epilogue"], k["work_items"]) == (2, 1, 6, 1, 8 * 32 * 16)
    assert ck(8, 32, 32, 8192, 8192, 128, bf, False)["cta_group"] == 2
    # MHA causal below 8K: 1-CTA kernel with the staged epilogue (four ring slots)
    k = ck(16, 32, 32, 4096, 4096, 128, bf, True)
    assert (k["cta_group"], k["stages"], k["staged_epilogue"], k["work_items"]) == (1, 4, 1, 16 * 32 * 16)
    # GQA: pairs cut by as many heads of a kv group as the group size allows, at every length
    k = ck(32, 32, 8, 2048, 2048, 128, bf, True)
    assert (k["cta_group"], k["heads_per_item"], k["work_items"]) == (2, 4, 32 * 8 * 16)             # 128-row items of four heads
    k = ck(32, 12, 6, 2048, 2048, 128, bf, True)
    assert (k["cta_group"], k["heads_per_item"], k["work_items"]) == (2, 2, 32 * 6 * 8)              # 256-row items of two heads
    assert ck(32, 12, 4, 2048, 2048, 128, bf, True)["cta_group"] == 1                                 # groups of 3: the table's causal row
    k = ck(16, 64, 8, 32768, 32768, 128, bf, True)                                                    # BASELINE configs[3]
    assert (k["cta_group"], k["heads_per_item"]) == (2, 4)
    # small launches whose tail the 1-CTA kernel smooths with half items stay there (192 blocks: 148 full items + 88 halves)
    k = ck(4, 12, 12, 1024, 1024, 128, bf, False)
    assert (k["cta_group"], k["work_items"]) == (1, 148 + 2 * 44)
    assert ck(2, 16, 16, 2048, 2048, 128, bf, False)["cta_group"] == 2                                # 256 blocks, no half-item tail: pairs
    # d = 64 has no pair kernel; fp32 has one geometry
    assert ck(8, 32, 8, 8192, 8192, 64, bf, True)["cta_group"] == 1 and ck(8, 32, 8, 8192, 8192, 64, bf, True)["softmax_warps"] == 16
    assert ck(1, 1, 1, 256, 256, 64, fa_b200.FA_DTYPE_F32, False)["work_items"] == 4
    # a reserve of 8 SMs (a communication kernel beside the launch) keeps everything on 1-CTA kernels
    L.fa_set_sm_reserve(8)
    try:
        assert ck(8, 32, 32, 8192, 8192, 128, bf, False)["cta_group"] == 1
    finally:
        L.fa_set_sm_reserve(0)
    with pytest.raises(fa_b200.FaError):
        ck(1, 3, 2, 16, 16, 128, bf, False)

