#!/usr/bin/env python
"""Benchmark of the SSNT lattice forward-backward hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload cfg2]

One "step" = one loss+grad pass of the hot path over one batch of synthetic log-probs.  The default
workload is BASELINE.json configs[1] — B=32 U=128 T=800 fp32 — per GPU (weak scaling: every rank owns
B=32 independent utterances; the only value that crosses GPUs is the scalar loss).  Prints ONE JSON line
(rank 0).

* value      lattice cells/s, inputs resident in HBM, CUDA-event timed, max over ranks.
* e2e        same metric through the C-ABI with HOST buffers the way the reference's DEVICE_CPU callers hold
             them — ordinary pageable numpy arrays: H2D of the inputs and D2H of log-likelihoods, loss and
             both gradient tensors inside the timed region (the number with caller-pinned buffers is
             reported beside it).
* roofline   one step's kernels (chunk operators, boundary sweep, chunk interiors, re-run check):
             algorithmic bytes (16 B/cell: read emit+shift, write two gradients) / device time of the step,
             against the measured HBM copy bandwidth.
* cpu_baseline  the CPU oracle's fp32 port (oracle/, the restatement standing in for the Rust reference,
             which has no forward-backward and cannot be built here) on all host cores, bounded sample.
             --impl reference runs only that arm (on the GLOBAL batch of the same N).
* secondary  driver-run numbers of the other BASELINE configs: tone-latent lattice (configs[2]), the
             B=4096 U=256 T=2000 sweep sharded B/N per rank (configs[4], strong scaling), and the
             decoding path (configs[3]: beam steps, whole-loop decode, back-trace, edit distance).

Multi-GPU: one process per GPU.  The loss all-reduce is not a host-issued collective: the kernel that
reduces a call's loss stores it into every rank's slot buffer over NVLink (ssnt_tts_loss_exchange_*), so
it replays with the CUDA graphs that hold the steps; NCCL only carries the timing reduction and a
cross-check of the exchanged loss outside the timed region.

L2 policy: the working set of one step fits the 126 MB L2, so the timed loop rotates through independent
input/output/scratch sets (> 3x L2 in total); every step's inputs come from HBM.
"""
from __future__ import annotations

import argparse
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (B, T, U, scaling) — B per GPU for weak scaling, B in total (sharded B/N) for strong scaling
    "cfg1": (1, 120, 32, "weak"),
    "cfg2": (32, 800, 128, "weak"),       # BASELINE configs[1], the headline
    "cfg3": (32, 800, 128, "weak"),       # tone-latent lattice, K = 4 tone classes (BASELINE configs[2])
    "cfg5": (4096, 2000, 256, "strong"),  # BASELINE configs[4]: B=4096 in total, B/N per rank
    "cfg5s": (512, 2000, 256, "weak"),    # a 1/8 slice of configs[4] per GPU
}
TONE_K = {"cfg3": 4}             # workloads that run the tone-latent lattice, and their class count
BYTES_PER_CELL = 16  # read log_emit+log_shift, write grad_emit+grad_shift (fp32); x K for the tone lattice
L2_BYTES = 126e6


def load_product():
    name = "ssnt_tts_rust_b200"
    if name in sys.modules:
        return sys.modules[name]
    pkg = os.path.join(ROOT, "ssnt-tts-rust_b200")
    spec = importlib.util.spec_from_file_location(name, os.path.join(pkg, "__init__.py"),
                                                  submodule_search_locations=[pkg])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def workload_text(name, world):
    B, T, U, scaling = WORKLOADS[name]
    K = TONE_K.get(name, 0)
    what = f"tone-latent (K={K}) SSNT loss+grad" if K else "batched SSNT loss+grad"
    if scaling == "strong":
        return (f"{name}: {what} fp32 B={B} U={U} T={T} in total, sharded B/N per GPU "
                f"(BASELINE configs[4]), full lengths")
    ref = {"cfg2": " (BASELINE configs[1])", "cfg3": " (BASELINE configs[2])"}.get(name, "")
    return f"{name}: {what} fp32 B={B} U={U} T={T} per GPU{ref}, full lengths"


# ---- synthetic data: counter-based, keyed by the GLOBAL (b, t, u) index -------------------------
def _mix64_np(x):
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def synthetic_numpy(b0, B, T, U, seed=1234):
    """z ~ N(0,1) from a splitmix64 counter; log_emit = log sigmoid(z), log_shift = log sigmoid(-z)."""
    with np.errstate(over="ignore"):
        idx = (np.arange(b0 * T * U, (b0 + B) * T * U, dtype=np.uint64) * np.uint64(2)
               + np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15))
        u1 = (_mix64_np(idx) >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))
        u2 = (_mix64_np(idx + np.uint64(1)) >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))
    z = np.sqrt(-2.0 * np.log(u1 + 1e-300)) * np.cos(2.0 * np.pi * u2)
    le = -np.logaddexp(0.0, -z)
    ls = -np.logaddexp(0.0, z)
    return le.reshape(B, T, U).astype(np.float32), ls.reshape(B, T, U).astype(np.float32)


def synthetic_torch(b0, B, T, U, device, seed=1234):
    """Same distribution generated on the device (cheap hash of the global cell index), a few utterances at a time."""
    import torch
    le = torch.empty(B, T, U, device=device)
    ls = torch.empty(B, T, U, device=device)

    def mix(x):
        x = (x ^ (x >> 30)) * -4658895280553007687      # 0xBF58476D1CE4E5B9 as int64
        x = (x ^ (x >> 27)) * -7723592293110705685      # 0x94D049BB133111EB as int64
        return x ^ (x >> 31)

    off = (seed * 0x9E3779B97F4A7C15) & ((1 << 64) - 1)
    off = off - (1 << 64) if off >= (1 << 63) else off   # two's-complement int64
    step = max(1, int(16e6 // (T * U)))                  # ~16 M cells of temporaries at a time
    for c0 in range(0, B, step):
        nb = min(step, B - c0)
        n = nb * T * U
        idx = torch.arange((b0 + c0) * T * U, (b0 + c0) * T * U + n, device=device, dtype=torch.int64)
        k = idx * 2 + off
        u1 = ((mix(k) >> 11) & ((1 << 53) - 1)).double() / float(1 << 53)
        u2 = ((mix(k + 1) >> 11) & ((1 << 53) - 1)).double() / float(1 << 53)
        z = (torch.sqrt(-2.0 * torch.log(u1 + 1e-300)) * torch.cos(2.0 * torch.pi * u2)).float()
        le[c0:c0 + nb] = torch.nn.functional.logsigmoid(z).reshape(nb, T, U)
        ls[c0:c0 + nb] = torch.nn.functional.logsigmoid(-z).reshape(nb, T, U)
    return le, ls


def synthetic_tone_numpy(b0, B, T, U, K, seed=1234):
    """Tone-latent inputs: the two-way emit/shift split per (cell, class) from the same counter hash
    (the class axis is folded into the token axis), log_tone = log_softmax of N(0,1)^K."""
    le, ls = synthetic_numpy(b0, B, T, U * K, seed)
    z, _ = synthetic_numpy(b0, B, 1, U * K, seed + 1)
    z = z.reshape(B, U, K).astype(np.float64)
    lt = z - np.log(np.exp(z).sum(-1, keepdims=True))
    return le.reshape(B, T, U, K), ls.reshape(B, T, U, K), lt.astype(np.float32)


def synthetic_tone_torch(b0, B, T, U, K, device, seed=1234):
    import torch
    le, ls = synthetic_torch(b0, B, T, U * K, device, seed)
    z, _ = synthetic_torch(b0, B, 1, U * K, device, seed + 1)
    lt = torch.log_softmax(z.reshape(B, U, K).double(), dim=-1).float().contiguous()
    return le.reshape(B, T, U, K), ls.reshape(B, T, U, K), lt


# ---- clocks -------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._gpu = gpu_index
        self._th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self._gpu)], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for n, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._th.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._th.join(timeout=6)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---- CPU arm ----------------------------------------------------------------------------------------------
def _cpu_runner(B, T, U, K):
    import oracle
    oracle.build()
    if K:
        inputs = synthetic_tone_numpy(0, B, T, U, K)
        return (lambda: oracle.tone_latent_forward_backward(*inputs, precision="f32")), oracle.get_threads()
    inputs = synthetic_numpy(0, B, T, U)
    return (lambda: oracle.forward_backward(*inputs, precision="f32")), oracle.get_threads()


def cpu_arm(B, T, U, budget_s=12.0, min_reps=2, max_reps=200, K=0):
    """Times the oracle's fp32 port (multi-threaded over the batch like rayon, persistent pool) on one batch of
    the workload, repeated until ~budget_s of wall time.  Returns (cells/s, cores, reps, median s/step)."""
    run, cores = _cpu_runner(B, T, U, K)
    run()  # touch: thread pool, page faults
    reps, t0 = 0, time.perf_counter()
    times = []
    while reps < max_reps and (reps < min_reps or time.perf_counter() - t0 < budget_s):
        s = time.perf_counter()
        run()
        times.append(time.perf_counter() - s)
        reps += 1
    per = float(np.median(times))
    return B * T * U / per, cores, reps, per


def run_reference(args):
    """The reference arm: the CPU path on the box's host cores, on the GLOBAL batch the N-GPU arm processes per
    step (weak scaling: N x B utterances; strong scaling: B), timed per step like cpu_arm (median)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    if rank != 0:
        return
    B, T, U, scaling = WORKLOADS[args.workload]
    K = TONE_K.get(args.workload, 0)
    Bg = B * world if scaling == "weak" else B
    # bounded sample: at most ~64 utterances' worth of the shape per step on the CPU, scaled (and said so)
    Bs = min(Bg, max(32, int(2.7e7 // (T * U * max(K, 1)))))
    run, cores = _cpu_runner(Bs, T, U, K)
    for _ in range(max(args.warmup, 1)):
        run()
    times = []
    for _ in range(args.steps):
        s = time.perf_counter()
        run()
        times.append(time.perf_counter() - s)
    per = float(np.median(times)) * (Bg / Bs)   # seconds per global step
    cells = Bg * T * U
    val = cells / per
    line = {
        "impl": "reference", "metric": "ssnt_fwd_bwd_lattice_cells_per_sec", "value": val, "unit": "cells/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * per,
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_text(args.workload, world), "global_batch": Bg,
                   "n_batches": Bg // B if scaling == "weak" else 1,
                   "note": "the reference crate has no forward-backward and cannot be built here (no cargo); this arm is "
                           "the oracle's fp32 C++ port of the authored spec, batch-parallel over all host cores like rayon, "
                           "median step time over the GLOBAL batch of this GPU count"
                           + (f" (timed on {Bs} utterances per step and scaled linearly)" if Bs != Bg else "")},
        "cpu_baseline": {"value": val, "unit": "cells/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} passes over {Bs} utterances of U={U} T={T}"
                                   + (f" K={K}" if K else "") + f", median {np.median(times) * 1e3:.1f} ms each"},
        "e2e": {"value": val, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---- GPU arm ----------------------------------------------------------------------------------------------
class Lattice:
    """Rotating input/output/scratch sets of one lattice workload on one GPU, and CUDA graphs over them."""

    def __init__(self, P, dev, B, T, U, K, b_global0, set_stride, min_sets_bytes=3.2 * L2_BYTES, group=1, max_bytes=60e9,
                 logits=False):
        import torch
        self.P, self.dev, self.B, self.T, self.U, self.K = P, dev, B, T, U, K
        kk = K or 1
        self.cells = B * T * U
        self.ws_bytes = (P.tone_latent_forward_backward_workspace_bytes(B, T, U, K) if K
                         else P.forward_backward_logits_workspace_bytes(B, T, U) if logits
                         else P.forward_backward_workspace_bytes(B, T, U))
        self.set_bytes = self.cells * kk * (8 if logits else 16) + self.ws_bytes
        n = max(1, int(np.ceil(min_sets_bytes / self.set_bytes)))
        n = ((n + group - 1) // group) * group            # whole graphs
        if self.set_bytes * n > max_bytes:
            n = max(1, int(max_bytes // self.set_bytes))
        self.nsets = n
        self.loss_all = torch.zeros(n, device=dev)
        self.sets = []
        for s in range(n):
            ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=dev)
            b0 = b_global0 + s * set_stride
            if K:
                inp = synthetic_tone_torch(b0, B, T, U, K, dev)
                out = (torch.empty(B, device=dev), self.loss_all[s:s + 1], torch.empty(B, T, U, K, device=dev),
                       torch.empty(B, T, U, K, device=dev), torch.empty(B, U, K, device=dev))
            elif logits:
                le, ls = synthetic_torch(b0, B, T, U, dev)
                inp = ((le - ls).contiguous(),)      # z = log sigmoid(z) - log sigmoid(-z)
                del le, ls
                out = (torch.empty(B, device=dev), self.loss_all[s:s + 1], torch.empty(B, T, U, device=dev))
            else:
                inp = synthetic_torch(b0, B, T, U, dev)
                out = (torch.empty(B, device=dev), self.loss_all[s:s + 1],
                       torch.empty(B, T, U, device=dev), torch.empty(B, T, U, device=dev))
            self.sets.append((inp, ws, out))
        self.call = P.tone_latent_forward_backward if K else (P.forward_backward_logits if logits else P.forward_backward)
        self.graphs = []

    def run_set(self, i):
        inp, ws, out = self.sets[i % self.nsets]
        return self.call(*inp, workspace=ws, out=out)

    def capture(self, group):
        """One CUDA graph per `group` consecutive sets (one C-ABI call per step, captured)."""
        import torch
        self.group = group
        for j in range(self.nsets // group):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                for i in range(j * group, (j + 1) * group):
                    self.run_set(i)
            self.graphs.append(g)
        for g in self.graphs:   # warm-up replay
            g.replay()
        torch.cuda.synchronize()

    def replay_steps(self, n, first=0):
        """Replays n steps (n is a multiple of the graph length, or there are no graphs)."""
        if not self.graphs:
            for i in range(n):
                self.run_set(first + i)
            return (first + n - 1) % self.nsets
        r = n // self.group
        for j in range(r):
            self.graphs[(first + j) % len(self.graphs)].replay()
        return (((first + r - 1) % len(self.graphs)) + 1) * self.group - 1   # index of the last step's set

    def time_steps(self, n, tail=None, head=None):
        """Device time of n steps in ms (CUDA events on torch's current stream); `tail` runs inside the timed region,
        `head` is enqueued just before it starts."""
        import torch
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if head:
            head()
        ev0.record()
        last = self.replay_steps(n)
        extra = tail() if tail else None
        ev1.record()
        torch.cuda.synchronize()
        return ev0.elapsed_time(ev1), last, extra


def graph_len(steps, cap=10):
    """Largest divisor of `steps` that is <= cap: the timed region is whole replays, whatever --steps is."""
    for g in range(min(cap, steps), 0, -1):
        if steps % g == 0:
            return g
    return 1


def kernel_description(kind, K):
    if K:
        return ("tone_split_kernel (block-float; cluster of 4 CTAs per utterance) + tone_fb_kernel "
                "(log domain, only the utterances the first kernel flagged)")
    return {6: "time-parallel block-float kernels: tp_build_kernel (chunk transfer operators, TMA-fed, one warp per 16-frame chunk) "
               "+ tp_combine_kernel (banded mat-vec sweep over the chunk boundaries, TMA ring, per-group exponents) "
               "+ tp_fill_kernel (chunk interiors and gradients) + fb_log_warp_kernel (re-run of flagged utterances only, loss)",
            8: "warp-serial block-float kernels (large batches): ws_forward_kernel (one warp per utterance, alpha in registers, "
               "rows through a TMA ring, a checkpoint every 8-16 rows) + ws_backward_kernel (chunks from the end: alpha re-run "
               "from the checkpoint, beta and gradients fused) + fb_log_warp_kernel (re-run of flagged utterances only, loss)",
            4: "fb_split_kernel (block-float; cluster of 4 CTAs per utterance: 2 recursion CTAs + 2 gradient CTAs, "
               "rows by ld.global.cg + L2 prefetch, DSMEM flags)",
            2: "fb_bf_kernel (block-float, warp-specialised cluster of 2 CTAs, TMA ring)",
            1: "fb_log_warp_kernel (log domain, cluster of 2 warps, TMA ring)",
            0: "fb_generic_kernel"}.get(kind)


def launches_per_step(kind, K):
    return 2 if K else {6: 4, 8: 3}.get(kind, 1)


def measure_lattice(P, dev, name, B, steps, warmup, world, rank, fb_kernel=-1, use_graph=True, exchange=False, logits=False):
    """Device-timed throughput of one lattice workload on this rank.  Returns a dict (per-rank numbers)."""
    import torch
    import torch.distributed as dist
    _, T, U, scaling = WORKLOADS[name]
    K = TONE_K.get(name, 0)
    kk = K or 1
    g = graph_len(steps) if use_graph else 1
    P.set_fb_kernel(fb_kernel)
    # every rank and every set draws its own utterances from the global counter space
    lat = Lattice(P, dev, B, T, U, K, b_global0=rank * B, set_stride=world * B, group=g, logits=logits)
    for i in range(max(warmup, 3)):
        lat.run_set(i)
    torch.cuda.synchronize()
    kind = P.fb_kernel_used()
    if use_graph:
        if world > 1:
            dist.barrier()
        lat.capture(g)
    red = torch.zeros(1, device=dev)
    tail = (lambda: P.loss_allreduce(red)) if exchange else None
    head = None
    if exchange:
        # device-side rendezvous on top of the host barrier: one more graph of steps, then the kernel that waits for
        # every rank's loss entry — all ranks' streams pass this point within microseconds of each other, so the
        # timed region (a few hundred microseconds at --steps 20) does not start with the host barrier's exit skew
        def head():
            lat.replay_steps(g if use_graph else 1)
            P.loss_allreduce(red)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms, last, _ = lat.time_steps(steps, tail, head)
    local_loss = float(lat.loss_all[last].item())
    out = {"ms": ms, "kind": kind, "cells": lat.cells, "nsets": lat.nsets, "set_bytes": lat.set_bytes, "graph_len": g,
           "ngraphs": len(lat.graphs), "local_loss": local_loss, "K": K, "T": T, "U": U, "B": B,
           "loss_allreduced": float(red.item()) if exchange else local_loss}
    # kernel-only timing for the roofline (same rotation, no exchange read)
    ksteps = max(steps, 2 * g)
    ksteps -= ksteps % g
    kms, _, _ = lat.time_steps(ksteps)
    out["kernel_ms"] = kms / ksteps
    out["lat"] = lat
    return out


def e2e_lattice(P, lat, world, host_group, steps, pinned):
    """The same pass through the C-ABI with HOST buffers (H2D + D2H inside the timed region), wall-clock timed."""
    import torch
    import torch.distributed as dist
    hsets = []
    for s in range(2):
        inp, _, out = lat.sets[s % lat.nsets]
        if pinned:
            hin = [x.cpu().pin_memory().numpy() for x in inp]
            hout = [torch.empty(x.shape).pin_memory().numpy() for x in out]
        else:   # what a DEVICE_CPU caller holds: ordinary pageable allocations
            hin = [np.array(x.cpu().numpy(), copy=True) for x in inp]
            hout = [np.empty(tuple(x.shape), np.float32) for x in out]
        hsets.append((hin, hout))

    def step(i):
        hin, hout = hsets[i % 2]
        lat.call(*hin, out=tuple(hout))
        loss = float(hout[1][0])
        if world > 1:   # the host owns the scalar here: a host-side all-reduce (gloo), 4 bytes
            t = torch.tensor([loss], dtype=torch.float64)
            dist.all_reduce(t, group=host_group)
            loss = float(t.item())
        return loss

    for i in range(2):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        loss = step(i)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    h2d = sum(x.size * 4 for x in hsets[0][0])
    d2h = sum(x.size * 4 for x in hsets[0][1])
    return dt, loss, h2d, d2h


def decoding_secondary(P, dev):
    """BASELINE configs[3]: beam=8, B=64, 150 input tokens, <= 1000 output frames; CUDA-event timed."""
    import torch
    out = {}

    def graph_time(fn, per_graph=20, replays=10):
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            for _ in range(per_graph):
                fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(replays):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / (per_graph * replays)   # us per call

    def event_time(fn, reps=20):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / reps

    B, W, D = 64, 8, 16
    gen = torch.Generator(device=dev).manual_seed(4)
    h = torch.log_softmax(torch.randn(B, W, D, device=dev, generator=gen), dim=-1)
    lp = torch.zeros(B, W, device=dev)
    fin = torch.zeros(B, W, dtype=torch.bool, device=dev)
    tot = torch.zeros(B, W, dtype=torch.int32, device=dev)
    tab = torch.arange(D, dtype=torch.int32, device=dev)
    tt = torch.zeros(B, W, dtype=torch.int32, device=dev)
    uu = torch.zeros(B, W, dtype=torch.int32, device=dev)
    il = torch.full((B,), 150, dtype=torch.int32, device=dev)
    ol = torch.full((B,), 1000, dtype=torch.int32, device=dev)
    us = graph_time(lambda: P.ssnt_tts_v2_beam_search_decode(h, lp, fin, tot, tab, tt, uu, il, ol, W, D, 0, False, True))
    out["v2_beam_step"] = {"shape": f"B={B} W={W} D={D}", "us_per_step": us, "steps_per_s": 1e6 / us,
                           "how": "one C-ABI call per output frame (the reference's ABI), captured in a CUDA graph"}
    hk = torch.log_softmax(torch.randn(B, W, 4, device=dev, generator=gen), dim=-1)
    us = graph_time(lambda: P.tone_latent_beam_search_decode(hk, lp, fin, tt, uu, il, W, 4, 0))
    out["tone_beam_step"] = {"shape": f"B={B} W={W} K=4", "us_per_step": us, "steps_per_s": 1e6 / us}
    if hasattr(P, "ssnt_tts_v2_decode_loop"):
        S = 150   # one step per input token in test mode
        hs = torch.log_softmax(torch.randn(B, S, W, D, device=dev, generator=gen), dim=-1)
        us = event_time(lambda: P.ssnt_tts_v2_decode_loop(hs, tab, il, ol, W, D, 0, False, True, 1000, -1), reps=10)
        out["v2_decode_loop"] = {"shape": f"B={B} W={W} D={D} steps={S} max_u=1000", "us_per_utterance_batch": us,
                                 "us_per_step": us / S, "steps_per_s": S * 1e6 / us,
                                 "how": "ONE launch: every step, back-trace of all beams and upsampling"}
    for L in (150, 1000):
        a = torch.randint(0, 50, (B, L), dtype=torch.int32, device=dev, generator=gen)
        b = torch.randint(0, 50, (B, L), dtype=torch.int32, device=dev, generator=gen)
        al = torch.full((B,), L, dtype=torch.int32, device=dev)
        us = graph_time(lambda: P.levenshtein_edit_distance(a, b, al, al))   # kernel time: an eager call is launch-bound at L=150
        out[f"edit_distance_L{L}"] = {"shape": f"B={B} lengths {L}", "us": us,
                                      "cell_updates_per_s": B * L * L / (us * 1e-6)}
    Tb = 1000
    bb = torch.randint(0, W, (B, Tb, W), dtype=torch.int32, device=dev, generator=gen)
    fb = torch.arange(W, dtype=torch.int32, device=dev).repeat(B, 1)
    out["order_beam_branch"] = {"shape": f"B={B} T={Tb} W={W}", "us": event_time(lambda: P.order_beam_branch(fb, bb, W))}
    return out


def run_b200(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    host_group = None
    P = load_product()
    P.lib()
    if world > 1:
        # the ranks share the host's cores: split them between the ranks' copy threads (host-pointer calls)
        os.environ.setdefault("SSNT_COPY_THREADS", str(max(2, min(12, (os.cpu_count() or 8) // world - 1))))
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        host_group = dist.new_group(backend="gloo")
        P.connect_loss_exchange()     # NVLink slot buffers: the loss all-reduce lives inside the kernels from here on
    name = args.workload
    Bcfg, T, U, scaling = WORKLOADS[name]
    B = Bcfg if scaling == "weak" else P.shard_range(Bcfg, rank, world)[1] - P.shard_range(Bcfg, rank, world)[0]
    K = TONE_K.get(name, 0)
    kk = K or 1
    steps = args.steps

    with ClockSampler(local_rank) as clk:
        m = measure_lattice(P, dev, name, B, steps, args.warmup, world, rank, fb_kernel=args.fb_kernel,
                            use_graph=not args.no_graph, exchange=world > 1)
        lat = m["lat"]
        # keep sampling a little under load if the run was very short: a FIXED number of extra replays (the same
        # on every rank — every call also advances the loss exchange's call counter, which must stay in step)
        if m["ms"] < 300:
            extra = lat.group if lat.graphs else 1
            est_step_s = max(m["cells"] / 5e10, 2e-5)      # from the shape only: identical on every rank
            for _ in range(min(2000, max(1, int(0.4 / (est_step_s * extra))))):
                lat.replay_steps(extra)
            torch.cuda.synchronize()
    ms = m["ms"]
    tt = torch.tensor([ms, m["kernel_ms"]], device=dev, dtype=torch.float64)
    cells_total = torch.tensor([float(m["cells"])], device=dev, dtype=torch.float64)
    nccl_loss = torch.tensor([m["local_loss"]], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(cells_total)
        dist.all_reduce(nccl_loss)      # cross-check of the in-kernel exchange, outside the timed region
    ms, k_ms = float(tt[0].item()), float(tt[1].item())
    cells_total = float(cells_total.item())

    # ---- e2e through the C-ABI with HOST buffers ---------------------------------------------------
    e2e_steps = max(3, min(steps, 20))
    e_dt, e_loss, h2d, d2h = e2e_lattice(P, lat, world, host_group, e2e_steps, pinned=False)
    p_dt, _, _, _ = e2e_lattice(P, lat, world, host_group, e2e_steps, pinned=True)
    te = torch.tensor([e_dt, p_dt], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e_dt, p_dt = float(te[0].item()), float(te[1].item())

    peak, peak_src = measured_peak_gbs()
    achieved = BYTES_PER_CELL * kk * m["cells"] / (k_ms * 1e-3) / 1e9
    g = m["graph_len"]
    line = {
        "metric": "ssnt_fwd_bwd_lattice_cells_per_sec",
        "value": cells_total * steps / (ms * 1e-3),
        "unit": "cells/s",
        "n_gpus": world, "steps": steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / steps,
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": workload_text(name, world),
            "global_batch": int(round(cells_total / (T * U))),
            "parallelism": f"batch-sharded dp{world}; the scalar loss is exchanged by NVLink peer stores from the kernel "
                           f"that reduces it (no host-issued collective)" if world > 1 else "single GPU",
            "launch": (f"{m['ngraphs']} CUDA graph(s) of {g} steps each (one C-ABI call per step, captured), "
                       f"{steps // g} replays in the timed region"
                       + ("; the timed region ends with the kernel that sums the ranks' exchanged losses" if world > 1 else "")
                       if m["ngraphs"] else "one C-ABI call per step from the host"),
            "l2_policy": f"rotating {m['nsets']} independent input/output/scratch sets "
                         f"({m['nsets'] * m['set_bytes'] / 1e6:.0f} MB > 3x 126 MB L2); inputs come from HBM every step",
            "fb_kernel": kernel_description(m["kind"], K),
            "loss_check": m["loss_allreduced"],
            "loss_check_nccl": float(nccl_loss.item()),
        },
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": None, "traffic_source": None, "peak_source": peak_src,
                     "kernel": "all kernels of one step (device time per step, CUDA events around graph replays)",
                     "kernel_ms": k_ms, "algorithmic_bytes_per_launch": BYTES_PER_CELL * kk * m["cells"]},
        "e2e": {"value": cells_total * e2e_steps / e_dt, "unit": "cells/s", "host_memory": "pageable",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "ms_per_step": 1e3 * e_dt / e2e_steps, "loss_check": e_loss,
                "pinned_host_memory": {"value": cells_total * e2e_steps / p_dt, "ms_per_step": 1e3 * p_dt / e2e_steps},
                "note": "host-pointer C-ABI call on ordinary numpy arrays: the library stages them chunk by chunk through its own "
                        "page-locked buffers (host-side copies on its copy threads, streaming stores) overlapped with the DMA "
                        "and the kernels of the neighbouring chunks; caller-pinned buffers are used in place",
                "copy_threads": int(os.environ.get("SSNT_COPY_THREADS", "0")) or min(12, max(1, (os.cpu_count() or 2) // 2)),
                "host_cores": os.cpu_count()},
        "gpu_launches": steps * launches_per_step(m["kind"], K),
        "clocks": clk.summary(),
    }
    traffic_file = os.path.join(ROOT, "profiles", "fb_traffic_bytes.json")
    if os.path.exists(traffic_file):
        try:
            tf = json.load(open(traffic_file))
            line["roofline"]["traffic"] = tf.get(name)
            line["roofline"]["traffic_source"] = tf.get("source", "static: one ncu --set full capture, not measured in this run")
        except Exception:
            pass
    del lat, m["lat"]
    torch.cuda.empty_cache()

    # ---- the other BASELINE configs, driver-run --------------------------------------------------------
    if not args.no_secondary:
        sec = {}
        ssteps = 10
        for sname in ("cfg3", "cfg5"):
            if sname == name:
                continue
            sB, sT, sU, sscal = WORKLOADS[sname]
            sK = TONE_K.get(sname, 0)
            if sscal == "strong":
                lo, hi = P.shard_range(sB, rank, world)
                sBr = hi - lo
            else:
                sBr = sB
            st = 5 if sname == "cfg5" else ssteps
            sm = measure_lattice(P, dev, sname, sBr, st, 3, world, rank, exchange=world > 1)
            v = torch.tensor([sm["ms"], sm["kernel_ms"]], device=dev, dtype=torch.float64)
            c = torch.tensor([float(sm["cells"])], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(v, op=dist.ReduceOp.MAX)
                dist.all_reduce(c)
            sach = BYTES_PER_CELL * (sK or 1) * sm["cells"] / (float(v[1].item()) * 1e-3) / 1e9
            sec[sname] = {
                "workload": workload_text(sname, world), "scaling": sscal, "batch_per_gpu": sBr,
                "value": float(c.item()) * st / (float(v[0].item()) * 1e-3), "unit": "cells/s",
                "ms_per_step": float(v[0].item()) / st, "steps": st,
                "roofline": {"bound": "hbm", "achieved": sach, "peak": peak, "unit": "GB/s", "frac": sach / peak,
                             "kernel_ms": float(v[1].item()),
                             "algorithmic_bytes_per_launch": BYTES_PER_CELL * (sK or 1) * sm["cells"]},
                "fb_kernel": kernel_description(sm["kind"], sK), "loss_check": sm["loss_allreduced"],
                "l2_policy": f"{sm['nsets']} set(s) of {sm['set_bytes'] / 1e6:.0f} MB",
            }
            del sm
            torch.cuda.empty_cache()
        # raw-logit entry (one logit instead of two log-probs; 12 algorithmic bytes per cell) at the headline shape
        lm = measure_lattice(P, dev, "cfg2", WORKLOADS["cfg2"][0], ssteps, 3, world, rank, exchange=world > 1, logits=True)
        v = torch.tensor([lm["ms"], lm["kernel_ms"]], device=dev, dtype=torch.float64)
        c = torch.tensor([float(lm["cells"])], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(v, op=dist.ReduceOp.MAX)
            dist.all_reduce(c)
        lach = 12 * lm["cells"] / (float(v[1].item()) * 1e-3) / 1e9
        sec["cfg2_logits"] = {
            "workload": "ssnt_tts_forward_backward_logits at the cfg2 shape (B=32 U=128 T=800 per GPU): log-sigmoids fused into "
                        "the kernels, gradient chained through them",
            "value": float(c.item()) * ssteps / (float(v[0].item()) * 1e-3), "unit": "cells/s",
            "ms_per_step": float(v[0].item()) / ssteps, "steps": ssteps,
            "roofline": {"bound": "hbm", "achieved": lach, "peak": peak, "unit": "GB/s", "frac": lach / peak,
                         "kernel_ms": float(v[1].item()), "algorithmic_bytes_per_launch": 12 * lm["cells"],
                         "note": "12 B/cell: read one logit, write one gradient (and the upstream log-sigmoid kernels disappear)"},
            "loss_check": lm["loss_allreduced"]}
        del lm
        torch.cuda.empty_cache()
        if rank == 0:
            sec["cfg4_decoding"] = decoding_secondary(P, dev)
        line["secondary"] = sec

    if rank == 0 and world == 1 and not args.no_cpu:
        v, cores, reps, per = cpu_arm(Bcfg, T, U, K=K)
        line["cpu_baseline"] = {"value": v, "unit": "cells/s", "cores": cores, "kind": "port",
                                "sample": f"{reps} passes over one B={Bcfg} U={U} T={T}{' K=%d' % K if K else ''} batch, "
                                          f"median {per * 1e3:.1f} ms each (oracle fp32 port, batch-parallel, persistent pool)"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        P.disconnect_loss_exchange()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-secondary", action="store_true", help="skip the other BASELINE configs")
    ap.add_argument("--no-graph", action="store_true", help="issue every step from Python instead of replaying CUDA graphs")
    ap.add_argument("--fb-kernel", type=int, default=-1,
                    help="-1 auto, 0 generic, 1 log-warp, 2 block-float fused, 4 block-float split-role, 6 time-parallel block-float, "
                         "8 warp-serial block-float (large batches)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
