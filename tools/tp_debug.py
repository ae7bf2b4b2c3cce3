import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from bench import load_product
from lattice_util import make_inputs
B, T, U = [int(x) for x in sys.argv[1:4]]
P = load_product(); P.lib()
le, ls = make_inputs(B, T, U, seed=1)
P.set_fb_kernel(int(sys.argv[4]) if len(sys.argv) > 4 else 6)
d = lambda a: torch.as_tensor(a, device="cuda")
print("calling", flush=True)
out = P.forward_backward(d(le), d(ls))
torch.cuda.synchronize()
print("done; ll", out[0][:4].tolist(), "loss", out[1].tolist(), flush=True)
if os.environ.get("SSNT_TP_DEBUG_STAGES", "4") == "4":
    import oracle; oracle.build()
    want = oracle.forward_backward(le, ls)
    print("ll err", np.abs(out[0].cpu().numpy() - want[0]).max(), "ge err", np.abs(out[2].cpu().numpy() - want[2]).max(),
          "gs err", np.abs(out[3].cpu().numpy() - want[3]).max(), "fallbacks", P.fb_fallback_count())
