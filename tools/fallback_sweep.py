"""Sweeps lattice shapes through the auto-dispatched forward-backward and reports, per shape, the kernel family taken,
how many utterances were re-run in the log domain (block-float fallbacks) and a coarse time per call.
    python tools/fallback_sweep.py [B] [tone]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import load_product, synthetic_torch, synthetic_tone_torch

def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    tone = len(sys.argv) > 2 and sys.argv[2] == "tone"
    P = load_product(); P.lib()
    dev = torch.device("cuda", 0)
    bad = 0
    for U in ((32, 64, 128) if tone else (32, 64, 96, 128, 160, 192, 256)):
        for T in (200, 400, 800, 1600, 3200):
            if U > T:
                continue
            reps = 4
            if tone:
                inps = [synthetic_tone_torch(s * B, B, T, U, 4, dev) for s in range(reps)]
                call = P.tone_latent_forward_backward
                count = P.fb_fallback_count   # one counter for both lattices
                used = (lambda: P.tone_kernel_used()) if hasattr(P, "tone_kernel_used") else (lambda: -1)
            else:
                inps = [synthetic_torch(s * B, B, T, U, dev) for s in range(reps)]
                call = P.forward_backward
                count = P.fb_fallback_count
                used = P.fb_kernel_used
            call(*inps[0]); torch.cuda.synchronize()
            f0 = count()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for inp in inps:
                out = call(*inp)
            e1.record(); torch.cuda.synchronize()
            nfb = count() - f0
            us = e0.elapsed_time(e1) * 1e3 / reps
            rows = (out[2] + out[3]).sum(dim=2) if not tone else (out[2] + out[3]).sum(dim=(2, 3))
            ok = bool(torch.allclose(rows, torch.ones_like(rows), atol=3e-4))
            flag = "  <-- re-runs" if nfb else ""
            bad += (not ok)
            print(f"B={B} U={U:4d} T={T:5d} T/U={T / U:5.1f} kind {used()}: {us:9.1f} us/call  re-run {nfb:4d} of {reps * B}  rows sum to 1: {ok}{flag}", flush=True)
    print("row-sum failures:", bad)

main()
