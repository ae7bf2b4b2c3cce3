import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from bench import load_product, synthetic_numpy
P = load_product()
B, T, U = (int(x) for x in sys.argv[1:4])
le, ls = synthetic_numpy(0, B, T, U)
P.set_fb_kernel(int(os.environ.get("KIND", "4")))
ll, loss, ge, gs = P.forward_backward(torch.as_tensor(le).cuda(), torch.as_tensor(ls).cuda())
torch.cuda.synchronize()
print((B, T, U), "ok", float(loss[0]))
