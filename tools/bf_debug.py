"""Bisecting aid for fb_bf_kernel (needs a GPU): fallbacks and errors vs the fp64 oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from bench import load_product, synthetic_torch
P = load_product()
B, T, U = int(os.environ.get("B", "32")), int(os.environ.get("T", "800")), 128
sets = [synthetic_torch(i * B, B, T, U, torch.device("cuda")) for i in range(4)]
P.set_fb_kernel(2)
f0 = P.fb_fallback_count()
worst = 0.0
for it in range(40):
    le, ls = sets[it % 4]
    ll, loss, ge, gs = P.forward_backward(le, ls)
    rows = (ge + gs).sum(2)
    worst = max(worst, float((rows - 1).abs().max()))
torch.cuda.synchronize()
print("DEBUG_SKIP", os.environ.get("SSNT_BF_DEBUG_SKIP"), "fallbacks", P.fb_fallback_count() - f0, "of", 40 * B, "worst row-sum dev", worst)
