"""sharding.py — the two multi-GPU modes of the attention forward path.  The reference has no
multi-device code at all; both modes are built on the single-GPU kernel behind the C ABI.

1. (batch x kv-head) sharding — no communication.  `shard_units` gives rank r a contiguous slice of the
   B*Hkv independent units (each unit carries its Hq/Hkv query heads, so K/V are never duplicated).

2. ring-KV — sequence-sharded long context.  Q/O stay resident on their rank; the K/V block of every rank
   travels once around the ring, one hop per step, overlapped with the MMAs of the current step.  Two transports:
     "p2p"   point-to-point send/recv (NCCL over NVLink on GPUs; gloo in the CPU tests), posted before the local
             attention calls.  NCCL moves the data with SM kernels, so the persistent attention kernel leaves a few
             SMs free for it (fa_set_sm_reserve); at 8 ranks the hop bandwidth (~77 GB/s measured) becomes the limit.
     "peer"  the K/V double buffer lives in symmetric memory (torch.distributed._symmetric_memory: every rank maps
             every peer's buffer over NVLink); each rank PULLS its predecessor's block with a plain device copy on a
             side stream (~690 GB/s measured, no SMs beyond a one-CTA barrier kernel per step) and device-side
             barriers order the hops.  Default on GPUs ("auto") when the rendezvous succeeds.  Partial results over disjoint key
   ranges are folded into an fp32 (O, log-sum-exp) carry inside the attention kernel's epilogue (fa_fwd_carry) —
   the reference's running (max, sum) recurrence (reference: kernels/utils.cuh:63-80) applied across ring steps
   instead of across tiles.
   Causal work is balanced with the zig-zag layout: the sequence is cut into 2P chunks and rank r owns chunks
   r and 2P-1-r, so every rank does the same amount of unmasked work at every step and fully masked block pairs
   are never launched.  Every ring step is ONE kernel launch into the rank's single fp32 accumulator
   (fa_fwd_carry_window: a step in which only the late local chunk sees the arriving keys addresses the accumulator
   rows of that chunk), P launches per pass plus one cast.

The ring driver takes the local "attend and fold into the carry" operator as an argument: on GPUs it defaults to
the CUDA path (fa_b200.attention_forward_carry); the CPU tests (gloo, world_size 2) pass a CPU stand-in to check
schedule and plumbing.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple


def shard_units(B: int, Hkv: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous slice [u0, u0+n) of the B*Hkv (batch, kv-head) units owned by `rank`."""
    units = B * Hkv
    u0 = units * rank // world
    u1 = units * (rank + 1) // world
    return u0, u1 - u0


def zigzag_chunks(world: int, rank: int) -> Tuple[int, int]:
    """Indices (of 2*world sequence chunks) owned by `rank` in the causal zig-zag layout."""
    return rank, 2 * world - 1 - rank


def zigzag_split(x, world: int, rank: int, dim: int = 2):
    """Take this rank's two chunks of a full-sequence tensor (test / setup helper)."""
    import torch
    chunks = x.chunk(2 * world, dim=dim)
    a, b = zigzag_chunks(world, rank)
    return torch.cat([chunks[a], chunks[b]], dim=dim).contiguous()


def zigzag_merge(parts, world: int, dim: int = 2):
    """Inverse of zigzag_split over the list of per-rank tensors (test helper)."""
    import torch
    chunks = [None] * (2 * world)
    for r, p in enumerate(parts):
        a, b = p.chunk(2, dim=dim)
        ia, ib = zigzag_chunks(world, r)
        chunks[ia], chunks[ib] = a, b
    return torch.cat(chunks, dim=dim)


def ring_schedule(world: int, rank: int, causal: bool):
    """The local attention call of one rank at every ring step: a list of (step, src_rank, q_part, kv_part, causal_flag)
    with exactly ONE entry per step.  Parts: 'ab' = all local rows, 'a' / 'b' = first / second local chunk.
      step 0            ('ab', 'ab', causal)  the rank's own block: in the zig-zag layout chunk a precedes chunk b in the
                                              global order too, so the local [a ; b] sequence against itself under the
                                              ordinary causal mask is exactly the two diagonal blocks plus (b, a)
      sender < rank     ('ab', 'a',  full)    only the sender's early chunk is visible — to both local chunks
      sender > rank     ('b',  'ab', full)    only the local late chunk sees the sender — both of its chunks
    Every entry is n^2 / 2 score elements (n = local rows), so all ranks do equal work at every step, and block pairs
    that are fully masked never reach the GPU (inside the step-0 call the kernel skips masked tiles)."""
    out = []
    for s in range(world):
        src = (rank - s) % world
        if not causal:
            out.append((s, src, "ab", "ab", False))
        elif s == 0:
            out.append((s, src, "ab", "ab", True))
        elif src < rank:
            out.append((s, src, "ab", "a", False))
        else:
            out.append((s, src, "b", "ab", False))
    return out


def _default_ops():
    import torch
    import fa_b200

    def step(q, k, v, causal, acc_o, acc_lse, row_offset):
        fa_b200.attention_forward_carry(q, k, v, acc_o, acc_lse, causal=causal, row_offset=row_offset)

    def finish(acc, like):
        return fa_b200.cast_out(acc, torch.empty(acc.shape, dtype=like.dtype, device=acc.device))

    return step, finish


class _PeerRing:
    """K/V double buffer in symmetric memory + the copy stream the hops run on.  One per (shape, dtype, group)."""

    def __init__(self, kv_shape, dtype, device, group):
        import math
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.kv_shape = tuple(kv_shape)
        self.dtype = dtype
        self.numel = math.prod(self.kv_shape)
        self.buf = symm.empty((2, self.numel), dtype=dtype, device=device)
        self.hdl = symm.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        self.copy_stream = torch.cuda.Stream(device=device)

    def slot(self, i):
        return self.buf[i].view(self.kv_shape)

    def peer_slot(self, rank, i):
        return self.hdl.get_buffer(rank, self.kv_shape, self.dtype, i * self.numel)


_PEER_RINGS = {}          # key -> _PeerRing, or None when the ranks agreed that this key has no symmetric-memory ring
_PEER_RING_LIMIT = 8      # symmetric buffers kept alive; the least recently created one is dropped beyond this


def _peer_ring(kv_shape, dtype, device, group):
    """Cached symmetric-memory ring for this (shape, dtype, group), or None when it cannot be had.  The decision is
    COLLECTIVE: every rank tries, the outcomes are combined with an all-reduce (MIN), and either all ranks use the ring or
    none does — a rank-local failure (out of memory for a new shape, say) can therefore never leave one rank in the
    point-to-point protocol while the others wait in a symmetric-memory barrier.  A failure is remembered for its own
    key only."""
    import torch
    import torch.distributed as dist
    key = (tuple(kv_shape), dtype, str(device), id(group))
    if key in _PEER_RINGS:
        return _PEER_RINGS[key]
    ring, err = None, None
    try:
        ring = _PeerRing(kv_shape, dtype, device, group)
    except Exception as ex:
        err = ex
    ok = torch.tensor([1 if ring is not None else 0], device=device, dtype=torch.int32)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
    if int(ok.item()) == 0:
        if ring is not None:
            del ring
        import warnings
        warnings.warn(f"symmetric-memory ring unavailable on some rank ({err}); using point-to-point send/recv for this shape")
        ring = None
    while len(_PEER_RINGS) >= _PEER_RING_LIMIT:
        _PEER_RINGS.pop(next(iter(_PEER_RINGS)))
    _PEER_RINGS[key] = ring
    return ring


def ring_attention(q, k, v, causal: bool = True, group=None, step_fn: Optional[Callable] = None,
                   finish_fn: Optional[Callable] = None, return_lse: bool = False, transport: str = "auto"):
    """Sequence-sharded attention.  q [B,Hq,n,d], k/v [B,Hkv,n,d] are this rank's shard of the sequence
    (zig-zag layout when causal: [chunk r ; chunk 2P-1-r], contiguous otherwise).  Returns this rank's shard of O.
    One K/V hop and ONE attention launch per step (step_fn(q, k, v, causal, acc_o, acc_lse, row_offset) folds the step
    into the rank's single fp32 accumulator; q's rows are the accumulator's rows [row_offset, row_offset + len(q)))."""
    import torch
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0   # degenerate ring: one rank owns the whole sequence, no communication
    if step_fn is None or finish_fn is None:
        d_step, d_finish = _default_ops()
        step_fn, finish_fn = step_fn or d_step, finish_fn or d_finish

    B, Hq, n, d = q.shape
    half = n // 2
    if causal and n % 2:
        raise ValueError("causal ring needs an even local length (two zig-zag chunks)")
    ring = None
    if world > 1 and q.is_cuda and transport in ("auto", "peer"):
        ring = _peer_ring((2,) + tuple(k.shape), k.dtype, k.device, group)
        if ring is None and transport == "peer":
            raise RuntimeError("transport='peer' requested but symmetric memory is unavailable")

    def part(t, which):          # rows of the local sequence axis (dim -2); slices stay strided views
        if which == "ab":
            return t
        return t[..., :half, :] if which == "a" else t[..., half:, :]

    acc_o = torch.zeros(B, Hq, n, d, device=q.device, dtype=torch.float32)
    acc_lse = torch.full((B, Hq, n), float("-inf"), device=q.device, dtype=torch.float32)
    sched = ring_schedule(world, rank, causal)

    def compute(s, kv):
        for (_, _, qp, kp, c) in [e for e in sched if e[0] == s]:
            step_fn(part(q, qp), part(kv[0], kp), part(kv[1], kp), c, acc_o, acc_lse, half if qp == "b" else 0)

    if ring is not None:
        # ---- "peer": pull the predecessor's block out of its symmetric-memory slot on a side stream ----
        prev = (rank - 1) % world
        cur = torch.cuda.current_stream(q.device)
        ring.hdl.barrier(channel=0)                   # every rank has finished with the previous call's slots
        ring.slot(0)[0].copy_(k)
        ring.slot(0)[1].copy_(v)
        ring.hdl.barrier(channel=0)                   # slot 0 of every rank holds its own block
        ev_prev = cur.record_event()
        ev_pull = None
        for s in range(world):
            if s + 1 < world:
                ring.copy_stream.wait_event(ev_prev)  # s = 0: slots filled; s > 0: my step s-1 no longer reads slot (s+1)%2
                with torch.cuda.stream(ring.copy_stream):
                    if s > 0:
                        ring.hdl.barrier(channel=1)   # all ranks: hop s-1 landed and step s-1 computed
                    ring.slot((s + 1) & 1).copy_(ring.peer_slot(prev, s & 1))
                    ev_next = ring.copy_stream.record_event()
            if ev_pull is not None:
                cur.wait_event(ev_pull)               # this step's block has landed
            compute(s, ring.slot(s & 1))
            ev_prev = cur.record_event()
            if s + 1 < world:
                ev_pull = ev_next
    else:
        # ---- "p2p": K and V travel as one buffer, a single send/recv pair per hop ----
        kv = torch.stack([k, v]).contiguous()
        nxt = torch.empty_like(kv) if world > 1 else None
        send_to, recv_from = (rank + 1) % world, (rank - 1) % world
        if group is not None:
            send_to, recv_from = dist.get_global_rank(group, send_to), dist.get_global_rank(group, recv_from)
        for s in range(world):
            reqs = []
            if s + 1 < world:   # post the hop for the NEXT step first: it overlaps the attention call below
                ops = [dist.P2POp(dist.isend, kv, send_to, group), dist.P2POp(dist.irecv, nxt, recv_from, group)]
                reqs = dist.batch_isend_irecv(ops)
            compute(s, kv)
            for r in reqs:
                r.wait()
            if s + 1 < world:
                kv, nxt = nxt, kv
    o = finish_fn(acc_o, q)
    return (o, acc_lse) if return_lse else o
