"""CPU tests (-m "not gpu") of the multi-GPU host logic: unit sharding, the zig-zag ring schedule, and the ring-KV
driver run for real over gloo with world_size 2 (and 4), with oracle-backed stand-ins for the local attention and
merge operators (the CUDA operators are exercised by the GPU tests)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import sharding
from oracle import oracle


def test_shard_units_partition_is_exact_and_contiguous():
    for B, Hkv, world in [(8, 32, 8), (16, 8, 8), (3, 5, 4), (1, 1, 8), (2, 3, 1)]:
        covered = []
        for r in range(world):
            u0, n = sharding.shard_units(B, Hkv, world, r)
            covered += list(range(u0, u0 + n))
        assert covered == list(range(B * Hkv))
        sizes = [sharding.shard_units(B, Hkv, world, r)[1] for r in range(world)]
        assert max(sizes) - min(sizes) <= 1


def _visible(qc, kc):   # chunk-level causal visibility: 0 none, 1 diagonal, 2 full
    return 2 if kc < qc else (1 if kc == qc else 0)


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_ring_schedule_covers_exactly_the_causal_block_pairs(world):
    need = {(qc, kc) for qc in range(2 * world) for kc in range(2 * world) if kc <= qc}
    got = set()
    work = []
    for r in range(world):
        own = dict(zip("ab", sharding.zigzag_chunks(world, r)))
        w = 0
        sched = sharding.ring_schedule(world, r, True)
        assert [e[0] for e in sched] == list(range(world)), "exactly one attention launch per ring step"
        for (s, src, qp, kp, c) in sched:
            assert src == (r - s) % world
            theirs = dict(zip("ab", sharding.zigzag_chunks(world, src)))
            for qn in qp:
                for kn in kp:
                    if c and kn > qn:
                        continue      # inside a causal call the block above the diagonal is masked: the kernel skips its tiles
                    qc, kc = own[qn], theirs[kn]
                    vis = _visible(qc, kc)
                    assert vis > 0, "a fully masked block pair was scheduled"
                    assert (vis == 1) == (c and qc == kc)
                    assert (qc, kc) not in got
                    got.add((qc, kc))
                    w += 1 if vis == 2 else 0.5
        work.append(w)
    assert got == need
    assert max(work) == min(work), f"zig-zag must balance causal work: {work}"


def _cpu_step(q, k, v, causal, acc_o, acc_lse, row_offset):
    """CPU stand-in for fa_fwd_carry_window: oracle attention over this key range, folded into rows
    [row_offset, row_offset + Nq) of the running (O, LSE) pair."""
    o, lse = oracle.attention_fwd(q.numpy(), k.numpy(), v.numpy(), causal=causal, return_lse=True)
    o, lse = torch.from_numpy(o), torch.from_numpy(lse)
    ao, al = acc_o[:, :, row_offset:row_offset + q.shape[2]], acc_lse[:, :, row_offset:row_offset + q.shape[2]]
    new = torch.logaddexp(al, lse)
    wa = torch.exp(al - new).nan_to_num(0.0)
    wb = torch.exp(lse - new).nan_to_num(0.0)
    ao.mul_(wa[..., None]).add_(o * wb[..., None])
    al.copy_(new)


def _worker(rank, world, port, causal, shape, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        B, Hq, Hkv, N, d = shape
        g = torch.Generator().manual_seed(0)
        q = torch.randn(B, Hq, N, d, generator=g)
        k = torch.randn(B, Hkv, N, d, generator=g)
        v = torch.randn(B, Hkv, N, d, generator=g)
        if causal:
            ql, kl, vl = (sharding.zigzag_split(t, world, rank) for t in (q, k, v))
        else:
            ql, kl, vl = (t.chunk(world, dim=2)[rank].contiguous() for t in (q, k, v))
        o, lse = sharding.ring_attention(ql, kl, vl, causal=causal, step_fn=_cpu_step,
                                         finish_fn=lambda acc, like: acc.to(like.dtype), return_lse=True)
        outs = [torch.empty_like(o) for _ in range(world)]
        dist.all_gather(outs, o)
        if rank == 0:
            full = sharding.zigzag_merge(outs, world) if causal else torch.cat(outs, dim=2)
            ref = oracle.attention_fwd(q.numpy(), k.numpy(), v.numpy(), causal=causal)
            ret["err"] = float(np.abs(full.numpy() - ref).max())
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


@pytest.mark.parametrize("world,causal", [(2, True), (2, False), (4, True)])
def test_ring_attention_over_gloo_matches_oracle(world, causal):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), causal, (1, 4, 2, 16 * world * 2, 16), ret), nprocs=world, join=True)
    assert ret["err"] <= 2e-5, ret["err"]
