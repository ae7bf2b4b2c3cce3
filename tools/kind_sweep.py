"""Times the lattice kernel kinds over batch sizes (profiling aid, needs a GPU)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import load_product, synthetic_torch
P = load_product()
dev = torch.device("cuda")
shapes = [(32, 2000, 256), (256, 2000, 256), (32, 400, 64)]
for (B, T, U) in shapes:
    le, ls = synthetic_torch(0, B, T, U, dev)
    ws = torch.empty(P.forward_backward_workspace_bytes(B, T, U), dtype=torch.uint8, device=dev)
    out = (torch.empty(B, device=dev), torch.empty(1, device=dev), torch.empty(B, T, U, device=dev), torch.empty(B, T, U, device=dev))
    res = []
    for kind in (2, 4):
        P.set_fb_kernel(kind)
        f0 = P.fb_fallback_count()
        for _ in range(3):
            P.forward_backward(le, ls, workspace=ws, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 10
        for _ in range(n):
            P.forward_backward(le, ls, workspace=ws, out=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        res.append(f"kind {kind}: {ms*1e3:8.1f} us  {B*T*U/ms/1e6:7.1f} Gcells/s  loss {float(out[1][0]):.3f} fb {P.fb_fallback_count()-f0}")
    print((B, T, U), " | ".join(res), flush=True)
