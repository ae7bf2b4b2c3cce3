import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from bench import load_product, synthetic_numpy
import oracle
P = load_product()
kind = int(os.environ.get("KIND", "4"))
for (B, T, U) in [(1, 16, 128), (2, 100, 128), (3, 333, 128), (4, 800, 128), (2, 200, 64), (2, 300, 256), (32, 800, 128), (33, 801, 128)]:
    le, ls = synthetic_numpy(0, B, T, U)
    want = oracle.forward_backward(le, ls)
    P.set_fb_kernel(kind)
    f0 = P.fb_fallback_count()
    ll, loss, ge, gs = P.forward_backward(torch.as_tensor(le).cuda(), torch.as_tensor(ls).cuda())
    torch.cuda.synchronize()
    ll = ll.cpu().numpy(); ge = ge.cpu().numpy(); gs = gs.cpu().numpy()
    print((B, T, U), "kind", P.fb_kernel_used(), "fallbacks", P.fb_fallback_count() - f0,
          "ll relerr %.2e" % (np.abs(ll - want[0]).max() / np.abs(want[0]).max()),
          "ge err %.2e gs err %.2e" % (np.abs(ge - want[2]).max(), np.abs(gs - want[3]).max()), flush=True)
