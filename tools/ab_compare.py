import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, numpy as np
from bench import load_product, synthetic_torch
P = load_product()
if os.environ.get("ABLIB"):
    P.LIB_PATH = os.path.join(ROOT, os.environ["ABLIB"])
P.lib(); P.set_fb_kernel(int(os.environ.get("KIND", "4")))
B, T, U = (int(x) for x in os.environ.get("SHAPE", "32,800,128").split(","))
dev = torch.device("cuda")
NS = 10
sets = []
for i in range(NS):
    le, ls = synthetic_torch(i * B, B, T, U, dev)
    ws = torch.empty(P.forward_backward_workspace_bytes(B, T, U), dtype=torch.uint8, device=dev)
    out = (torch.empty(B, device=dev), torch.empty(1, device=dev), torch.empty(B, T, U, device=dev), torch.empty(B, T, U, device=dev))
    sets.append((le, ls, ws, out))
for i in range(30):
    le, ls, ws, out = sets[i % NS]; P.forward_backward(le, ls, workspace=ws, out=out)
torch.cuda.synchronize()
N = 300
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(N)]
for i in range(N):
    le, ls, ws, out = sets[i % NS]
    ev[i][0].record(); P.forward_backward(le, ls, workspace=ws, out=out); ev[i][1].record()
torch.cuda.synchronize()
t = np.array([a.elapsed_time(b) for a, b in ev]) * 1e3
print(f"{os.environ.get('ABLIB','current'):28s} kind {P.fb_kernel_used()} median {np.median(t):.2f} us  p10 {np.percentile(t,10):.2f}  p90 {np.percentile(t,90):.2f}  fallbacks {P.fb_fallback_count()}")
