"""Times the host-pointer lattice call (C-ABI, H2D + kernels + D2H) on pageable or caller-pinned numpy buffers.
    [SSNT_FB_CHUNKS=n] [SSNT_COPY_THREADS=n] python tools/e2e_probe.py [pageable|pinned] [B T U]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, numpy as np
from bench import load_product, synthetic_torch
P = load_product()
mode = sys.argv[1] if len(sys.argv) > 1 else "pageable"
B, T, U = [int(x) for x in sys.argv[2:5]] if len(sys.argv) >= 5 else (32, 800, 128)
le, ls = synthetic_torch(0, B, T, U, torch.device("cuda"))
if mode == "pinned":
    args = (le.cpu().pin_memory().numpy(), ls.cpu().pin_memory().numpy())
    out = tuple(torch.empty(s).pin_memory().numpy() for s in ((B,), (1,), (B, T, U), (B, T, U)))
else:
    args = (np.array(le.cpu().numpy(), copy=True), np.array(ls.cpu().numpy(), copy=True))
    out = tuple(np.empty(s, np.float32) for s in ((B,), (1,), (B, T, U), (B, T, U)))
for i in range(3): P.forward_backward(*args, out=out)
ts = []
for i in range(20):
    t0 = time.perf_counter()
    P.forward_backward(*args, out=out)
    ts.append(time.perf_counter() - t0)
print(mode, "chunks", os.environ.get("SSNT_FB_CHUNKS"), "threads", os.environ.get("SSNT_COPY_THREADS"),
      "ms/call median %.3f min %.3f" % (np.median(ts) * 1e3, min(ts) * 1e3), "loss", float(out[1][0]), "cores", os.cpu_count(), flush=True)
