import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import load_product
P = load_product()
dev = torch.device("cuda")
B, T, U, K = 32, 800, 128, 4
g = torch.Generator(device="cuda").manual_seed(3)
z = torch.randn(B, T, U, K, device=dev, generator=g)
le, ls = torch.nn.functional.logsigmoid(z), torch.nn.functional.logsigmoid(-z)
lt = torch.log_softmax(torch.randn(B, U, K, device=dev, generator=g), dim=-1)
for _ in range(3):
    r = P.tone_latent_forward_backward(le, ls, lt)
torch.cuda.synchronize()
print("ok", float(r[1][0]))
