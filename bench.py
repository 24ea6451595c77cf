#!/usr/bin/env python
"""bench.py — attention forward throughput on B200, the metric of BASELINE.json.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg3|cfg2|cfg4|cfg3nc]

A step is one pass of the hot path (fa_fwd through the C ABI) over one batch of synthetic Q/K/V.
Default workload = BASELINE.json configs[2]: B=8 H=32 N=8192 d=128 causal bf16 (the shape the metric is quoted on).
  value      whole-job TFLOP/s with inputs resident in HBM: the K steps are ONE device-timed region (a CUDA-event pair on
             the launch stream around all of them, launch gaps included), max over ranks
  e2e        same metric through fa_fwd_host: pinned HOST buffers, H2D + kernel + D2H inside the timed region
  roofline   tensor-core bound: algorithmic FLOPs (4*B*Hq*Nq*Nk*d, halved when causal) / mean per-launch kernel time (one
             event pair per step) vs the measured
             cuBLAS bf16 peak in MEASURED_PEAKS.json
  cpu_baseline  the torch-CPU restatement of the reference's check.py path timed on this box's host cores on a
             bounded sample of the same workload (rank 0, N=1 only)
  sustained  the same launch held back to back for >= 0.7 s under ONE event pair (the power-capped steady state), with
             the clocks sampled during it; `value` / `roofline` stay the K-step burst the contract asks for
  e2e.pcie   the same host<->device bytes moved by plain pinned copies on the same rank, no kernel: the floor of e2e
N > 1: one process per GPU (torchrun); (batch x head) units are sharded with NO data-path collective — every
rank runs the per-GPU workload on its own units ("weak" scaling); only the timing reduction uses NCCL.  The line then
also carries a `ring` sub-record: BASELINE configs[4] (N=128K causal, B=1 H=8, sequence-sharded ring-KV, strong-scaled
over the N ranks) with both hop transports, its scaling against one GPU running the whole sequence, and its error
against that single-GPU result and against the oracle on sampled rows.
--impl reference times the reference's CPU implementation (check.py path, torch-CPU port in oracle/) instead, and
runs the reference's own CUDA kernel (oracle/_ref, unmodified) on the shapes it can execute, in a subprocess.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

# stdout carries exactly one JSON line: whatever NCCL has to say (NCCL_DEBUG=INFO through its logger, NCCL_DEBUG=VERSION
# through a bare printf) goes to stderr — file descriptor 1 points at stderr until the line is printed (emit()).
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
sys.stdout.flush()
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    os.dup2(_REAL_STDOUT, 1)
    print(json.dumps(line), flush=True)
    os.dup2(2, 1)


ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "flash-attention-cuda-c_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

WORKLOADS = {
    # name: (B, Hq, Hkv, N, d, causal, dtype, description)
    "cfg3": (8, 32, 32, 8192, 128, True, "bf16", "Llama-3-8B attention shape B=8 H=32 N=8192 d=128 causal bf16 (BASELINE configs[2])"),
    "cfg3nc": (8, 32, 32, 8192, 128, False, "bf16", "B=8 H=32 N=8192 d=128 non-causal bf16"),
    "cfg2": (4, 12, 12, 1024, 64, False, "fp16", "GPT-2 shape B=4 H=12 N=1024 d=64 non-causal fp16 (BASELINE configs[1])"),
    "cfg4": (16, 64, 8, 32768, 128, True, "bf16", "GQA Hq=64 Hkv=8 N=32K B=16 d=128 causal bf16 (BASELINE configs[3]); per GPU: B=16/N ranks"),
    "cfg5": (1, 8, 8, 131072, 128, True, "bf16", "long context N=128K d=128 causal bf16, B=1 H=8, sequence-sharded ring-KV (BASELINE configs[4]); per GPU: N/ranks rows, zig-zag"),
}
STRONG = {"cfg4", "cfg5"}   # total work fixed as ranks grow; cfg2/cfg3 replicate the per-GPU workload (weak)


def metric_name(wl):
    B, Hq, Hkv, N, d, causal, dtype, _ = wl
    n = f"{N // 1024}K" if N % 1024 == 0 else str(N)
    gqa = f", GQA {Hq}/{Hkv}" if Hq != Hkv else ""
    return (f"attention fwd TFLOP/s ({dtype}, d={d}, N={n}{' causal' if causal else ''}{gqa}), whole job; "
            "roofline.frac = fraction of measured dense bf16 TC peak")



def config_for(wl, wl_name, world):
    """The `config` object of the JSON line — identical in both arms (--impl ours / reference) for the same flags."""
    B, Hq, Hkv, N, d, causal, dtype, desc = wl
    if wl_name == "cfg4":
        B = max(1, B // world)
    if wl_name == "cfg5":
        N = N // world
    return {"workload": desc, "per_gpu": {"B": B, "Hq": Hq, "Hkv": Hkv, "N": N, "d": d, "causal": causal}, "n_gpus": world}


def flops(B, Hq, Nq, Nk, d, causal):
    f = 4.0 * B * Hq * Nq * Nk * d
    return f / 2 if causal else f


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["bf16_tflops"]), float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), float(p["hbm_gbs"]), "measured"
    except Exception:
        return 1590.0, 1400.0, 6650.0, "fallback"


class ClockSampler:
    """SM clock, power and clock-event reasons of one GPU, sampled in-process through NVML every ~2 ms by a thread, each
    sample time-stamped; `stop(window)` reports the median over the samples that fall INSIDE the timed region (the GPU is busy
    for all of it by construction), plus every sample's median.  The default timed region is ~70 ms: nvidia-smi's own loop
    (-lms) delivers 3-4 samples in that time and its power reading lags by more than the region, so neither its clock nor
    a power threshold can tell "under load" there.  Falls back to an nvidia-smi -lms 20 subprocess (power-threshold
    classification: >= half of the enforced limit) when pynvml is missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,"
         "enforced.power.limit")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc, self.nv, self.stop_flag, self.thread = gpu_index, [], None, None, False, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = self.idx
            if vis:
                ent = [e.strip() for e in vis.split(",") if e.strip()]
                if self.idx < len(ent) and ent[self.idx].isdigit():
                    phys = int(ent[self.idx])
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nv = (pynvml, h)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.limit = pynvml.nvmlDeviceGetEnforcedPowerLimit(h) / 1000.0
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nv = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv, h = self.nv
        names = (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown", "nvmlClocksThrottleReasonHwSlowdown"),
                 ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown", "nvmlClocksThrottleReasonHwThermalSlowdown"),
                 ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown", "nvmlClocksThrottleReasonSwThermalSlowdown"),
                 ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap", "nvmlClocksThrottleReasonSwPowerCap"))
        bits = [(n, getattr(nv, a, None) or getattr(nv, b, 0)) for n, a, b in names]
        while not self.stop_flag:
            try:
                t = time.perf_counter()
                mhz = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                watts = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                self.rows.append((t, mhz, watts, tuple(n for n, bit in bits if mask & bit)))
            except Exception:
                pass
            time.sleep(0.002)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self, window=None):
        """window = (t0, t1) in time.perf_counter() seconds: the timed region."""
        if self.nv is not None:
            self.stop_flag = True
            self.thread.join(timeout=1.0)
            rows = list(self.rows)
            thr = 0.5 * self.limit
            if window is not None:
                load = [r for r in rows if window[0] <= r[0] <= window[1]]
                how = "samples time-stamped inside the timed region"
            else:
                load = [r for r in rows if r[2] >= thr]
                how = "samples drawing >= half of the enforced power limit"
            reasons = sorted({n for r in load for n in r[3]})
            return {"sm_mhz": statistics.median([r[1] for r in load]) if load else None,
                    "sm_mhz_all_samples": statistics.median([r[1] for r in rows]) if rows else None,
                    "sm_mhz_min_under_load": min(r[1] for r in load) if load else None, "sm_max_mhz": self.max_mhz,
                    "power_w_max": max((r[2] for r in rows), default=None), "power_limit_w": self.limit,
                    "samples": len(rows), "samples_under_load": len(load), "under_load_means": how, "source": "NVML, 2 ms poll",
                    "reasons": reasons, "reasons_any_sample": sorted({n for r in rows for n in r[3]})}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["NVML and nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        sm, mx, reasons, power, limit = [], [], set(), [], None
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
                for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                    if r[col].lower().startswith("active"):
                        reasons.add(name)
                limit = float(r[9])
            except Exception:
                continue
        # "under load" = drawing at least half of the enforced power limit (the attention kernel runs AT the limit; an
        # idle B200 draws about a quarter of it), not half of whatever the largest sample happened to be
        thr = 0.5 * limit if limit else (0.5 * max(power) if power else 0.0)
        load = [c for c, w in zip(sm, power) if w >= thr]
        return {"sm_mhz": statistics.median(load) if load else None,
                "sm_mhz_all_samples": statistics.median(sm) if sm else None,
                "sm_mhz_min_under_load": min(load) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "power_limit_w": limit, "load_threshold_w": thr,
                "samples": len(sm), "samples_under_load": len(load), "under_load_means": "samples drawing >= half of the enforced power limit",
                "source": "nvidia-smi -lms 20", "reasons": sorted(reasons)}


def cpu_attention_sample(N, d, causal, heads, torch):
    """The check.py path (torch-CPU port, oracle.multi_head_attention) on `heads` heads of the workload, fp32."""
    from oracle import oracle
    g = torch.Generator().manual_seed(0)
    Q = torch.randn(1, N, heads * d, generator=g)
    K = torch.randn(1, N, heads * d, generator=g)
    V = torch.randn(1, N, heads * d, generator=g)
    t0 = time.perf_counter()
    out, _ = oracle.multi_head_attention(Q, K, V, heads, causal=causal)
    dt = time.perf_counter() - t0
    return dt, float(out.abs().mean())


def run_reference(args, wl):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    B, Hq, Hkv, N, d, causal, dtype, desc = wl
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    heads = 2 if N >= 8192 else min(Hq, 12)
    t1, _ = cpu_attention_sample(N, d, causal, 1, torch)   # calibrate: keep the whole run within a few minutes
    budget = 150.0 / max(1, args.steps + args.warmup)
    heads = max(1, min(heads * 4, int(budget / max(t1, 1e-3)), Hq))
    for _ in range(args.warmup):
        cpu_attention_sample(N, d, causal, heads, torch)
    times = [cpu_attention_sample(N, d, causal, heads, torch)[0] for _ in range(args.steps)]
    ms = 1e3 * sum(times) / len(times)
    val = flops(1, heads, N, N, d, causal) / (ms * 1e-3) / 1e12
    sample = f"{heads} of {B * Hq} (batch, head) slices of the workload per step, fp32, [N,N] scores materialised (check.py path)"
    # the reference's own CUDA kernel beside it, on the shapes it can execute (fp32, B*H = 1, grid = 1): a subprocess that
    # loads oracle/_ref/libref_kernel.so and nothing of this repository's product
    ref_cuda = {"unavailable": "not run"}
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ref_kernel_only.py")], capture_output=True, text=True, timeout=240)
        ref_cuda = json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 and r.stdout.strip() else {"unavailable": (r.stderr or "no output")[-300:]}
    except Exception as ex:
        ref_cuda = {"unavailable": str(ex)[:300]}
    world = int(os.environ.get("WORLD_SIZE", "1"))
    line = {"impl": "reference", "metric": metric_name(wl), "value": val, "unit": "TFLOP/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if args.workload in STRONG else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_for(wl, args.workload, world),
            "cpu_baseline": {"value": val, "unit": "TFLOP/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "reference_cuda_kernel": ref_cuda,
            "gpu_launches": 0}
    emit(line)
    return 0


def run_ours(args, wl, wl_name):
    import torch
    import fa_b200
    B, Hq, Hkv, N, d, causal, dtype, desc = wl
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if wl_name == "cfg4":
        B = max(1, B // world)   # 128 kv groups sharded by batch: strong-scaled config, 16/N batches per GPU
    ring = wl_name == "cfg5"
    N_total = N
    import sharding
    if ring:
        N = N // world           # this rank's rows (two zig-zag chunks of N/(2*world))
        if world > 1:
            fa_b200.set_sm_reserve(4)   # room for the one-CTA barrier kernels (peer transport) / NCCL kernels (p2p)
    tdt = {"bf16": torch.bfloat16, "fp16": torch.float16}[dtype]
    dev = torch.device("cuda", local)
    g = torch.Generator(device=dev).manual_seed(rank)
    q = torch.randn(B, Hq, N, d, device=dev, generator=g).to(tdt)
    k = torch.randn(B, Hkv, N, d, device=dev, generator=g).to(tdt)
    v = torch.randn(B, Hkv, N, d, device=dev, generator=g).to(tdt)
    o = torch.empty_like(q)
    F = flops(B, Hq, N, N, d, causal)
    if ring:
        F = flops(B, Hq, N_total, N_total, d, causal) / world   # this rank's share of the whole-sequence work
    es = 2
    alg_bytes = (2 * B * Hq * N * d + 2 * B * Hkv * N * d) * es
    flush = None
    if alg_bytes < 256 << 20:   # working set near L2 size: flush L2 between timed iterations
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step():
        if ring:
            sharding.ring_attention(q, k, v, causal=causal)
        else:
            fa_b200.attention_forward(q, k, v, causal=causal, out=o)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    launches0 = fa_b200.launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for a, b in evs:
        if flush is not None:
            flush.zero_()
        a.record()
        step()
        b.record()
    ev1.record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = fa_b200.launch_count() - launches0
    kernel_ms = [a.elapsed_time(b) for a, b in evs]
    # the K steps as ONE device-timed region (first launch to last retirement, gaps between launches included); with an L2
    # flush between steps the flush writes are not part of a step, so the per-step event times are summed instead
    total_ms = ev0.elapsed_time(ev1) if flush is None else sum(kernel_ms)
    clocks = sampler.stop((t_wall0, t_wall0 + t_wall)) if rank == 0 else None
    tt = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_per_step = float(tt.item()) / args.steps
    value = F * world / (ms_per_step * 1e-3) / 1e12

    def reduce_max(x):
        """max over ranks of a python float (a failure anywhere is -1 on that rank and poisons the result for all: every rank
        reaches this collective whether its local leg raised or not)."""
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        bad = torch.tensor([1.0 if x < 0 else 0.0], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(bad, op=dist.ReduceOp.MAX)
        return -1.0 if bad.item() > 0 else float(t.item())

    # ---- sustained: the same launch held for >= 0.7 s under one event pair (power-capped steady state) ------
    sustained_rec = None
    if not ring and not args.no_extras:
        n_sus = int(min(20000, max(200, 700.0 / max(ms_per_step, 1e-3))))
        sus_sampler = ClockSampler(local)
        if rank == 0:
            sus_sampler.start()
        for _ in range(10):
            step()
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_sus0 = time.perf_counter()
        s0.record()
        for _ in range(n_sus):
            step()
        s1.record()
        barrier()
        t_sus1 = time.perf_counter()
        sus_ms = reduce_max(s0.elapsed_time(s1)) / n_sus
        sus_clocks = sus_sampler.stop((t_sus0, t_sus1)) if rank == 0 else None
        sustained_rec = {"launches": n_sus, "ms_per_step": sus_ms, "value": F * world / (sus_ms * 1e-3) / 1e12, "unit": "TFLOP/s",
                         "per_gpu": F / (sus_ms * 1e-3) / 1e12, "clocks": sus_clocks,
                         "note": "same launch back to back under one CUDA-event pair, max over ranks; L2 not flushed (inputs > L2 for the default workload)"}

    # ---- end to end: pinned host buffers through fa_fwd_host, copies inside the timed region --------------
    e2e = None
    e2e_steps = max(1, min(args.steps, 3))
    if ring or args.no_extras:
        e2e = {"value": None, "error": "ring-KV keeps Q/K/V sharded and resident on the GPUs; the host-buffer path is measured on cfg3" if ring else "skipped (--no-extras)"}
    else:
        dt, pcie_dt, err = -1.0, -1.0, None
        try:
            hq, hk, hv = (t.cpu().pin_memory() for t in (q, k, v))
            ho = torch.empty_like(hq).pin_memory()
            fa_b200.attention_forward_host(hq, hk, hv, ho, causal=causal)   # warm-up (allocates the staging buffers)
        except Exception as ex:   # report, never hide
            err = str(ex)
        barrier()
        if err is None:
            try:
                t0 = time.perf_counter()
                for _ in range(e2e_steps):
                    fa_b200.attention_forward_host(hq, hk, hv, ho, causal=causal)
                torch.cuda.synchronize()
                dt = (time.perf_counter() - t0) / e2e_steps
            except Exception as ex:
                err = str(ex)
        dt_all = reduce_max(dt)
        result_check = float(ho[0, 0, :4].float().abs().mean()) if err is None else None    # before the copy probe reuses `ho`
        # the same bytes as plain pinned copies, no kernel: one cudaMemcpyAsync per tensor and chunk on three streams, the
        # chunking fa_fwd_host uses (96 MB of units per chunk) — the PCIe floor under e2e on this rank, and (max over
        # ranks) what the ranks of one box get when they copy at the same time
        if err is None:
            try:
                units = B * Hkv
                gq = Hq // Hkv
                q_unit, kv_unit = gq * N * d, N * d
                upc = max(1, min(units, (96 << 20) // ((2 * q_unit + 2 * kv_unit) * es + gq * N * 4)))
                hqf, hkf, hvf, hof = hq.view(units, q_unit), hk.view(units, kv_unit), hv.view(units, kv_unit), ho.view(units, q_unit)
                streams = [torch.cuda.Stream(device=dev) for _ in range(3)]
                stage = [(torch.empty(upc, q_unit, dtype=tdt, device=dev), torch.empty(upc, kv_unit, dtype=tdt, device=dev),
                          torch.empty(upc, kv_unit, dtype=tdt, device=dev), torch.empty(upc, q_unit, dtype=tdt, device=dev)) for _ in range(3)]

                def copies():
                    for ci, u0 in enumerate(range(0, units, upc)):
                        nu = min(upc, units - u0)
                        sq, sk, sv, so = stage[ci % 3]
                        with torch.cuda.stream(streams[ci % 3]):
                            sq[:nu].copy_(hqf[u0:u0 + nu], non_blocking=True)
                            sk[:nu].copy_(hkf[u0:u0 + nu], non_blocking=True)
                            sv[:nu].copy_(hvf[u0:u0 + nu], non_blocking=True)
                            hof[u0:u0 + nu].copy_(so[:nu], non_blocking=True)
                    torch.cuda.synchronize()

                copies()
                barrier()
                t0 = time.perf_counter()
                for _ in range(e2e_steps):
                    copies()
                pcie_dt = (time.perf_counter() - t0) / e2e_steps
            except Exception as ex:
                err = "pcie probe: " + str(ex)
        else:
            barrier()
        pcie_all = reduce_max(pcie_dt)
        if dt_all > 0:
            h2d = (hq.numel() + hk.numel() + hv.numel()) * es
            d2h = ho.numel() * es
            e2e = {"value": F * world / dt_all / 1e12, "unit": "TFLOP/s", "h2d_bytes_per_step": h2d,
                   "d2h_bytes_per_step": d2h, "ms_per_step": dt_all * 1e3, "steps": e2e_steps,
                   "api": "fa_fwd_host (C ABI, pinned host buffers, 3-stream H2D/kernel/D2H pipeline)",
                   "result_check": result_check,
                   "pcie": ({"ms_per_step": pcie_all * 1e3, "h2d_gbs": h2d / pcie_all / 1e9, "d2h_gbs": d2h / pcie_all / 1e9,
                             "frac_of_e2e": pcie_all / dt_all,
                             "note": "same bytes, same chunking, plain pinned cudaMemcpyAsync on 3 streams with no kernel, max over ranks "
                                     "(per-rank rates; the ranks of one box share the host's PCIe / memory bandwidth)"}
                            if pcie_all > 0 else {"error": err})}
        else:
            e2e = {"value": None, "error": err or "failed on another rank"}

    # ---- ring-KV (BASELINE configs[4]) beside the sharded workload whenever there is more than one rank ----
    ring_rec = None
    if world > 1 and not ring and not args.no_extras:
        ring_rec = ring_leg(torch, dist, fa_b200, dev, rank, world)

    small = None
    if rank == 0 and world == 1 and not ring and not args.no_extras:
        small = small_shapes(torch, fa_b200, dev)

    if rank == 0:
        burst, sustained, hbm, how = peaks()
        k_ms = sum(kernel_ms) / len(kernel_ms)          # this rank's mean per-step time from the per-step event pairs
        kernel_name = "fa::fwdFp32Kernel"
        if dtype != "fp32":      # what the launcher's tile table picks for this workload
            tc = fa_b200.choose_kernel(B, Hq, Hkv, N, N, d, {"bf16": fa_b200.FA_DTYPE_BF16, "fp16": fa_b200.FA_DTYPE_F16}[dtype], causal)
            kernel_name = "fa::fwdSm100PairKernel (CTA pairs, cta_group::2)" if tc["cta_group"] == 2 else "fa::fwdSm100Kernel"
            kernel_name += (f" [softmax warps {tc['softmax_warps']}, staged epilogue {tc['staged_epilogue']}, ring slots {tc['stages']}, "
                            f"{tc['work_items']} work items, {tc['heads_per_item']} head(s) per item]")
        per_gpu = F / (k_ms * 1e-3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get(wl_name)
            except Exception:
                traffic = None
        roofline = {"bound": "tensor", "achieved": per_gpu, "peak": burst, "unit": "TFLOP/s", "frac": per_gpu / burst,
                    "peak_source": f"MEASURED_PEAKS.json bf16_tflops ({how}, burst: kernel timed alone)",
                    "frac_of_nominal_2250": per_gpu / 2250.0,
                    "traffic": traffic, "algorithmic_bytes": alg_bytes, "algorithmic_flops": F,
                    "hbm_gbs_achieved": alg_bytes / (k_ms * 1e-3) / 1e9, "hbm_peak_gbs": hbm,
                    "kernel": kernel_name, "kernel_ms": k_ms}
        if sustained_rec is not None:
            sustained_rec["peak_sustained"] = sustained
            sustained_rec["frac_of_sustained_peak"] = sustained_rec["per_gpu"] / sustained
        cpu = None
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            t1, _ = cpu_attention_sample(N, d, causal, 1, torch)
            heads = max(1, min(Hq * B, int(12.0 / max(t1, 1e-3))))
            dt, _ = cpu_attention_sample(N, d, causal, heads, torch)
            cpu = {"value": flops(1, heads, N, N, d, causal) / dt / 1e12, "unit": "TFLOP/s", "cores": torch.get_num_threads(),
                   "kind": "port", "seconds": dt,
                   "sample": f"{heads} of {B * Hq} (batch, head) slices, fp32, torch-CPU port of check.py:4-25 (scores materialised)"}
        line = {"metric": metric_name(wl), "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if wl_name in STRONG else "weak", "vs_baseline": None,
                "dtype": dtype, "data": "synthetic",
                "config": config_for(wl, wl_name, world),
                "method": {"sharding": ("sequence-sharded ring-KV, zig-zag causal layout: one K/V hop and one attention launch per step, overlapped; transport "
                                        + ("symmetric-memory peer pull over NVLink (copy on a side stream), NCCL send/recv as fallback"
                                           if (world > 1 and any(r is not None for r in sharding._PEER_RINGS.values())) else "NCCL send/recv")
                                        if ring else "(batch x head) units per rank, no data-path collective"),
                           "l2": ("working set %.2f GiB > 126 MB L2" % (alg_bytes / 2**30)) if flush is None else "L2 flushed (256 MiB write) between timed iterations",
                           "timing": ("one CUDA-event pair around the K steps on the launch stream (launch gaps included)" if flush is None
                                      else "CUDA events per step on the launch stream, summed (the L2 flush between steps is not timed)") + "; max over ranks"},
                "roofline": roofline, "sustained": sustained_rec, "cpu_baseline": cpu, "e2e": e2e, "ring": ring_rec, "small_shapes": small,
                "gpu_launches": launches, "clocks": clocks,
                "wall_s_timed_region": t_wall, "kernel_ms_min": min(kernel_ms), "kernel_ms_max": max(kernel_ms)}
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def small_shapes(torch, fa_b200, dev):
    """Ours on the shapes the reference's own CUDA kernel can execute (the reference arm's `reference_cuda_kernel` record
    times that kernel on the same shapes) and on BASELINE configs[1] (GPT-2 shape), eager and under CUDA-graph replay."""
    def timeit(fn, iters):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters

    out = {}
    try:
        g = torch.Generator(device=dev).manual_seed(0)
        for name, N in (("cfg1_fp32_N256_d64", 256), ("fp32_N2048_d64", 2048)):
            q, k, v = (torch.randn(1, 1, N, 64, device=dev, generator=g) for _ in range(3))
            o = torch.empty_like(q)
            ms = timeit(lambda: fa_b200.attention_forward(q, k, v, out=o), 50)
            out[name] = {"ms": ms, "gflops": 4.0 * N * N * 64 / ms / 1e6, "kernel": "fa::fwdFp32Kernel"}
        B, H, N, d = 4, 12, 1024, 64
        q, k, v = (torch.randn(B, H, N, d, device=dev, generator=g).to(torch.float16) for _ in range(3))
        o = torch.empty_like(q)
        F, by = 4.0 * B * H * N * N * d, 4 * B * H * N * d * 2
        ms = timeit(lambda: fa_b200.attention_forward(q, k, v, out=o), 200)
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for _ in range(20):
                fa_b200.attention_forward(q, k, v, out=o)
        ms_g = timeit(gr.replay, 20) / 20
        out["cfg2_fp16_B4_H12_N1024_d64"] = {"eager_us": ms * 1e3, "graph_replay_us": ms_g * 1e3, "tflops_eager": F / ms / 1e9,
                                              "tflops_graph": F / ms_g / 1e9, "gbs_eager": by / ms / 1e6, "gbs_graph": by / ms_g / 1e6,
                                              "note": "back-to-back launches, working set 25 MB stays in L2; 60 % of TC peak would be 12.9 us"}
    except Exception as ex:
        out["error"] = str(ex)
    return out


def ring_leg(torch, dist, fa_b200, dev, rank, world):
    """BASELINE configs[4]: N=128K causal bf16 d=128, B=1 H=8, sequence-sharded over the ranks (zig-zag), strong-scaled.
    Every rank builds the same full Q/K/V from one seed, runs the whole sequence on its own GPU (the single-GPU time and the
    result the ring is compared with), then the ring with both hop transports."""
    import sharding
    B, H, N, d = 1, 8, 131072, 128
    rec = {"workload": WORKLOADS["cfg5"][7], "n_gpus": world, "kernel": fa_b200.lib().fa_version().decode()}
    try:
        g = torch.Generator(device=dev).manual_seed(1234)
        q, k, v = (torch.randn(B, H, N, d, device=dev, generator=g).to(torch.bfloat16) for _ in range(3))
        o_full = torch.empty_like(q)
        F = flops(B, H, N, N, d, True)

        def timed(fn, iters):
            for _ in range(2):
                fn()
            dist.barrier(); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(iters):
                fn()
            b.record()
            dist.barrier(); torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b) / iters], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        t1 = timed(lambda: fa_b200.attention_forward(q, k, v, causal=True, out=o_full), 3)
        rec["single_gpu_ms"] = t1
        rec["single_gpu_tflops"] = F / t1 / 1e9
        ql, kl, vl = (sharding.zigzag_split(t, world, rank) for t in (q, k, v))
        want = sharding.zigzag_split(o_full, world, rank).float()
        # SMs the persistent attention kernel leaves free: 4 for the one-CTA barrier kernels of the peer transport, 8 for
        # NCCL's send/recv kernels (profiles/multi/r2_ring_p2p_tune_g8.jsonl: 0-4 SMs 8.7 ms per pass, 8 SMs 6.5 ms, more is worse)
        for transport, reserve in (("peer", 4), ("p2p", 8)):
            fa_b200.set_sm_reserve(reserve)
            try:
                ms = timed(lambda: sharding.ring_attention(ql, kl, vl, causal=True, transport=transport), 5)
                out = sharding.ring_attention(ql, kl, vl, causal=True, transport=transport)
                e = torch.tensor([(out.float() - want).abs().max().item()], device=dev, dtype=torch.float64)
                dist.all_reduce(e, op=dist.ReduceOp.MAX)
                rec[transport] = {"ms_per_pass": ms, "value": F / ms / 1e9, "unit": "TFLOP/s", "scaling_vs_single_gpu": t1 / ms,
                                  "max_abs_err_vs_single_gpu": float(e.item()), "launches_per_pass": world + 1, "sm_reserve": reserve}
                if rank == 0 and "oracle_sample_max_abs_err" not in rec:
                    # rank 0 owns the last chunk of the sequence: its last 128 local rows are the global rows N-128 .. N-1
                    import numpy as np
                    from oracle import oracle
                    o_ref = oracle.attention_fwd(q[:, :1, N - 128:].float().cpu().numpy(), k[:, :1].float().cpu().numpy(),
                                                 v[:, :1].float().cpu().numpy(), causal=True)
                    rec["oracle_sample_max_abs_err"] = float(np.abs(out[:, :1, -128:].float().cpu().numpy() - o_ref).max())
            except Exception as ex:
                rec[transport] = {"error": str(ex)[:300]}
        fa_b200.set_sm_reserve(0)
        rec["transports"] = {"peer": "symmetric-memory pull over NVLink on a copy stream (torch.distributed._symmetric_memory)",
                             "p2p": "NCCL send/recv (batch_isend_irecv) posted before the step's attention launch"}
    except Exception as ex:
        rec["error"] = str(ex)[:300]
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="profiling runs: only warm-up + the timed region (no sustained / e2e / ring / small-shape legs)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, wl)
    return run_ours(args, wl, args.workload)


if __name__ == "__main__":
    sys.exit(main())
