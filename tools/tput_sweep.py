import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import load_product, synthetic_torch
P = load_product()
dev = torch.device("cuda")
for (B, T, U) in [(64, 800, 128), (148, 800, 128), (512, 800, 128), (1024, 400, 64)]:
    le, ls = synthetic_torch(0, B, T, U, dev)
    ws = torch.empty(P.forward_backward_workspace_bytes(B, T, U), dtype=torch.uint8, device=dev)
    out = (torch.empty(B, device=dev), torch.empty(1, device=dev), torch.empty(B, T, U, device=dev), torch.empty(B, T, U, device=dev))
    P.set_fb_kernel(2)
    for _ in range(3): P.forward_backward(le, ls, workspace=ws, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): P.forward_backward(le, ls, workspace=ws, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("two_per_sm", os.environ.get("SSNT_BF_TWO_PER_SM", "auto"), (B, T, U), f"{ms*1e3:8.1f} us {B*T*U/ms/1e6:7.1f} Gcells/s loss {float(out[1][0]):.2f}", flush=True)
