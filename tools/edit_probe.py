"""Times the batched edit distance (device pointers, CUDA events over graph replays) and checks it against the oracle.
    [SSNT_EDIT_WAVEFRONT=1] python tools/edit_probe.py [B L ...]     (default: 64 150  64 1000  64 1024  512 150)"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import load_product
import oracle

def main():
    a = [int(x) for x in sys.argv[1:]] or [64, 150, 64, 1000, 64, 1024, 512, 150]
    P = load_product(); P.lib()
    dev = torch.device("cuda", 0)
    for B, L in zip(a[0::2], a[1::2]):
        rng = np.random.default_rng(L)
        x = rng.integers(0, 60, (B, L)).astype(np.int32)
        y = x.copy()
        m = rng.random((B, L)) < 0.3
        y[m] = rng.integers(0, 60, m.sum())
        ln = np.full(B, L, np.int32)
        want = oracle.levenshtein_edit_distance(x, y, ln, ln)
        dx, dy, dl = (torch.as_tensor(v, device=dev) for v in (x, y, ln))
        got = P.levenshtein_edit_distance(dx, dy, dl, dl)
        torch.cuda.synchronize()
        ok = bool((got.cpu().numpy() == want).all())
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            for _ in range(10):
                P.levenshtein_edit_distance(dx, dy, dl, dl)
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 100
        print(f"B={B} L={L}: {us:.1f} us, {B * L * L / us / 1e3:.1f} G cell updates/s, exact={ok}", flush=True)

main()
