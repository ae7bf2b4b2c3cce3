"""One lattice call at a given shape (for the in-kernel timing / trace builds).   python tools/tp_once.py [B T U kind nsets]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import load_product, synthetic_torch
a = [int(x) for x in sys.argv[1:]]
B, T, U, kind, nsets = (a + [32, 800, 128, 6, 1][len(a):])[:5]
P = load_product(); P.lib()
dev = torch.device("cuda", 0)
P.set_fb_kernel(kind)
for s in range(nsets):
    inp = synthetic_torch(s * B, B, T, U, dev)
    ll, loss, ge, gs = P.forward_backward(*inp)
    torch.cuda.synchronize()
    print("set", s, "loss", float(loss), "kind used", P.fb_kernel_used(), "fallbacks so far", P.fb_fallback_count(), flush=True)
