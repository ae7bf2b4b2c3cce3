"""Where the warps of tone_split_kernel spend their cycles (profiling aid, needs a GPU)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from ctypes import c_void_p
from bench import load_product
P = load_product()
dev = torch.device("cuda")
B, T, U, K = 32, 800, 128, 4
g = torch.Generator(device="cuda").manual_seed(3)
z = torch.randn(B, T, U, K, device=dev, generator=g)
le, ls = torch.nn.functional.logsigmoid(z), torch.nn.functional.logsigmoid(-z)
lt = torch.log_softmax(torch.randn(B, U, K, device=dev, generator=g), dim=-1)
stats = torch.zeros(4 * B * 8 * 8, dtype=torch.int64, device=dev)
for i in range(4):
    if i == 3: P.lib().ssnt_tts_debug_set_fb_stats(c_void_p(stats.data_ptr()))
    P.tone_latent_forward_backward(le, ls, lt)
torch.cuda.synchronize()
P.lib().ssnt_tts_debug_set_fb_stats(c_void_p(0))
s = stats.view(B, 4, 8, 8).double().cpu().mean(0)
for r in (0, 1):
    print(f"chain CTA {r}: recursion total {s[r,0,0]:.0f} wait_ready {s[r,0,1]:.0f} wait_free {s[r,0,2]:.0f} | prep w1 total {s[r,1,0]:.0f} wait_slot {s[r,1,1]:.0f} | prep w6 total {s[r,6,0]:.0f} wait_slot {s[r,6,1]:.0f}")
for r in (2, 3):
    print(f"grad CTA {r}: warp0 total {s[r,0,0]:.0f} waiting {s[r,0,1]:.0f} first_row_at {s[r,0,2]:.0f} | warp7 total {s[r,7,0]:.0f} waiting {s[r,7,1]:.0f}")
