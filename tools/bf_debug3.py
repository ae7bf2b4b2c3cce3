import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from bench import load_product, synthetic_torch
P = load_product()
B, T, U = 32, 800, 128
SU = U + 32
P.set_fb_kernel(2)
s, b = 3, 2
le, ls = synthetic_torch(s * B, B, T, U, torch.device("cuda"))
ws = torch.zeros(P.forward_backward_workspace_bytes(B, T, U), dtype=torch.uint8, device="cuda")
ll, loss, ge, gs = P.forward_backward(le, ls, workspace=ws)
torch.cuda.synchronize()
scr = ws[: B * (T + 1) * SU * 4].view(torch.float32).view(B, T + 1, SU)[b].cpu()
vals = scr[:, :U].numpy(); exps = scr[:, U:].contiguous().view(torch.int32).numpy()
np.set_printoptions(linewidth=250, precision=3)
m = (T + 1) // 2
print("m", m)
bad = np.nonzero(~np.isfinite(vals).all(axis=1) | (np.abs(vals) > 1e30).any(axis=1))[0]
print("rows non-finite / huge:", bad[:40])
mx = np.abs(vals).max(axis=1)
for t in list(range(m, m + 3)) + list(range(T - 60, T + 1, 1)):
    lanes = vals[t].reshape(32, 4)
    lm = np.abs(lanes).max(axis=1)
    print(t, "lane log2 max", np.where(lm > 0, np.floor(np.log2(np.maximum(lm, 1e-45))), -999).astype(int)[-10:], "exps", exps[t][-10:])
