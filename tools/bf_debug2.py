"""Finds the utterances fb_bf_kernel flags and shows where their block-float result is wrong (needs a GPU)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from bench import load_product, synthetic_torch
import oracle
P = load_product()
B, T, U = 32, 800, 128
SU = U + 32
P.set_fb_kernel(2)
for s in range(4):
    le, ls = synthetic_torch(s * B, B, T, U, torch.device("cuda"))
    ws = torch.zeros(P.forward_backward_workspace_bytes(B, T, U), dtype=torch.uint8, device="cuda")
    ll, loss, ge, gs = P.forward_backward(le, ls, workspace=ws)
    torch.cuda.synchronize()
    st = ws[B * (T + 1) * SU * 4: B * (T + 1) * SU * 4 + 4 * B].view(torch.int32).cpu().numpy()
    bad = np.nonzero(st)[0]
    print("set", s, "status nonzero:", [(int(b), int(st[b])) for b in bad])
    for b in bad:
        w = oracle.forward_backward(le[b:b+1].cpu().numpy(), ls[b:b+1].cpu().numpy())
        g1 = ge[b].cpu().numpy(); g2 = gs[b].cpu().numpy()
        e1 = np.abs(g1 - w[2][0]).max(axis=1); e2 = np.abs(g2 - w[3][0]).max(axis=1)
        print("  b", b, "ll", float(ll[b]), "oracle", w[0][0], "max err ge", e1.max(), "gs", e2.max())
        rows = np.nonzero((e1 > 1e-4) | (e2 > 1e-4))[0]
        print("  rows with err > 1e-4:", rows[:20], "... count", len(rows))
        if len(rows):
            t = rows[0]
            cols = np.nonzero(np.abs(g1[t] - w[2][0][t]) > 1e-5)[0]
            print("  first bad row", t, "cols", cols[:16], "got", g1[t][cols[:6]], "want", w[2][0][t][cols[:6]])
        rs = (g1 + g2).sum(1)
        print("  row sums min/max", rs.min(), rs.max(), "argmax dev", np.abs(rs - 1).argmax())
