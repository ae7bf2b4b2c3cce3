"""Times the tone-latent lattice kernels at one shape with CUDA events over rotating buffer sets (inputs from HBM every
step, like bench.py) and checks the kinds against each other.   python tools/tone_probe.py B T U K [kinds...]"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import load_product, synthetic_tone_torch

def main():
    args = [int(x) for x in sys.argv[1:]]
    B, T, U, K = (args[:4] if len(args) >= 4 else (32, 800, 128, 4))
    kinds = args[4:] or [1, 2]
    P = load_product(); P.lib()
    dev = torch.device("cuda", 0)
    ws_bytes = P.tone_latent_forward_backward_workspace_bytes(B, T, U, K)
    cells = B * T * U
    set_bytes = cells * 16 * K + ws_bytes
    nsets = max(2, min(12, int(np.ceil(3.2 * 126e6 / set_bytes))))
    sets = []
    for s in range(nsets):
        inp = synthetic_tone_torch(s * B, B, T, U, K, dev)
        out = (torch.empty(B, device=dev), torch.empty(1, device=dev), torch.empty(B, T, U, K, device=dev),
               torch.empty(B, T, U, K, device=dev), torch.empty(B, U, K, device=dev))
        sets.append((inp, torch.empty(ws_bytes, dtype=torch.uint8, device=dev), out))
    ref = None
    for kind in kinds:
        P.set_tone_kernel(kind)
        def run(i):
            inp, ws, out = sets[i % nsets]
            return P.tone_latent_forward_backward(*inp, workspace=ws, out=out)
        fb0 = P.fb_fallback_count()
        for i in range(nsets):
            run(i)
        torch.cuda.synchronize()
        res = [x.clone() for x in sets[0][2]]
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        ev0.record()
        for i in range(reps * nsets):
            run(i)
        ev1.record(); torch.cuda.synchronize()
        us = ev0.elapsed_time(ev1) * 1e3 / (reps * nsets)
        msg = (f"tone kind {kind} used {P.tone_kernel_used()}: {us:.1f} us/step, {cells / us / 1e3:.1f} G cells/s, "
               f"frac {16 * K * cells / us / 1e3 / 6548.8:.3f}, fallbacks {P.fb_fallback_count() - fb0}")
        if ref is None:
            ref = res
        else:
            msg += " | vs first kind: ll %.2e ge %.2e gs %.2e gt %.2e" % (
                float(((res[0] - ref[0]).abs() / ref[0].abs()).max()), float((res[2] - ref[2]).abs().max()),
                float((res[3] - ref[3]).abs().max()), float((res[4] - ref[4]).abs().max()))
        print(msg, flush=True)
    P.set_tone_kernel(-1)

if __name__ == "__main__":
    main()
