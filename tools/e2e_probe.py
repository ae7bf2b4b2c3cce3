import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, numpy as np
from bench import load_product, synthetic_torch
P = load_product()
B, T, U = 32, 800, 128
le, ls = synthetic_torch(0, B, T, U, torch.device("cuda"))
h = dict(le=le.cpu().pin_memory(), ls=ls.cpu().pin_memory(), ll=torch.empty(B).pin_memory(), loss=torch.empty(1).pin_memory(),
         ge=torch.empty(B, T, U).pin_memory(), gs=torch.empty(B, T, U).pin_memory())
args = (h["le"].numpy(), h["ls"].numpy())
out = (h["ll"].numpy(), h["loss"].numpy(), h["ge"].numpy(), h["gs"].numpy())
for i in range(3): P.forward_backward(*args, out=out)
t0 = time.perf_counter()
for i in range(20): P.forward_backward(*args, out=out)
torch.cuda.synchronize()
print("chunks", os.environ.get("SSNT_FB_CHUNKS"), "ms/call", (time.perf_counter() - t0) / 20 * 1e3, "loss", float(h["loss"][0]))
