"""Where the warps of fb_split_kernel spend their cycles (profiling aid, needs a GPU)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import load_product, synthetic_torch
from ctypes import c_void_p
P = load_product()
B, T, U = 32, 800, 128
dev = torch.device("cuda")
NSETS = int(os.environ.get("NSETS", "6"))
sets = [synthetic_torch(i * B, B, T, U, dev) for i in range(NSETS)]
stats = torch.zeros(4 * B * 16 * 16 + 4 * B * 4 + 4096, dtype=torch.int64, device=dev)
P.set_fb_kernel(4)
NIT = 3 * NSETS + 1
for i in range(NIT):
    if i == NIT - 1:
        P.lib().ssnt_tts_debug_set_fb_stats(c_void_p(stats.data_ptr()))
    le, ls = sets[i % NSETS]
    P.forward_backward(le, ls)
torch.cuda.synchronize()
P.lib().ssnt_tts_debug_set_fb_stats(c_void_p(0))
s = stats[:4 * B * 16 * 16].view(B, 4, 16, 16).double().cpu()
for rank in (0, 1):
    m = s[:, rank].mean(0)
    print(f"chain CTA {rank}: recursion total {m[0,0]:.0f} wait_ready {m[0,1]:.0f} first_ready_at {m[0,2]:.0f} rows {m[0,3]:.0f} handoff {m[0,4]:.0f} mid {m[0,5]:.0f} warmup_end_at {m[0,6]:.0f} | prep(w1) total {m[1,0]:.0f} wait_slot {m[1,1]:.0f} first_stage_at {m[1,2]:.0f} [loop start {m[1,3]:.0f} loads issued {m[1,4]:.0f} loaded {m[1,5]:.0f} stored {m[1,6]:.0f}] | copy-out(w4) total {m[4,0]:.0f} wait_state {m[4,1]:.0f} fence {m[4,2]:.0f}")
for rank in (2, 3):
    m = s[:, rank].mean(0)
    for w in (0, 7, 15):
        print(f"gradient CTA {rank} warp {w}: total {m[w,0]:.0f} grad {m[w,3]:.0f} grad_start_at {m[w,5]:.0f} waiting {m[w,6]:.0f}")

tl = stats[4 * B * 16 * 16: 4 * B * 16 * 16 + 4 * B * 4].view(B, 4, 4).double().cpu()
t0 = tl[:, :, 0].min()
for rank in range(4):
    r = tl[:, rank] - t0
    print(f"rank {rank}: entry min/max {r[:,0].min():.0f}/{r[:,0].max():.0f} ns | after first cluster.sync {r[:,1].min():.0f}/{r[:,1].max():.0f} | role done {r[:,2].min():.0f}/{r[:,2].max():.0f} | exit {r[:,3].min():.0f}/{r[:,3].max():.0f}")
