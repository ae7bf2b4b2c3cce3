"""Prints where the warps of fb_bf_kernel spend their cycles (profiling aid, needs a GPU)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import load_product, synthetic_torch
from ctypes import c_void_p

P = load_product()
B, T, U = 32, 800, 128
dev = torch.device("cuda")
NSETS = int(os.environ.get('NSETS', '7'))
sets = [synthetic_torch(i * B, B, T, U, dev) for i in range(NSETS)]
stats = torch.zeros(2 * B * 8 * 16 + 4 * 256 * 4, dtype=torch.int64, device=dev)
P.set_fb_kernel(int(os.environ.get("KIND", "2")))
NIT = 3 * NSETS + 1
for i in range(NIT):
    if i == NIT - 1:
        P.lib().ssnt_tts_debug_set_fb_stats(c_void_p(stats.data_ptr()))
    le, ls = sets[i % NSETS]
    P.forward_backward(le, ls)
torch.cuda.synchronize()
print('fallbacks so far:', P.fb_fallback_count(), 'of', NIT * B, 'utterances')
P.lib().ssnt_tts_debug_set_fb_stats(c_void_p(0))
tl = stats[2 * B * 8 * 16:].view(4, 256, 4).cpu()
s = stats[:2 * B * 8 * 16].view(B, 2, 8, 16).double().cpu()
names = ["chain", "helper1", "helper2", "helper3", "producer", "helper5", "helper6", "helper7"]
for rank in (0, 1):
    print("rank", rank, "(alpha)" if rank == 0 else "(beta)")
    for w in range(8):
        m = s[:, rank, w].mean(0)
        print(f"  {names[w]:9s} total {m[0]:9.0f} cyc | blocked: raw_full {m[1]:8.0f} prep_full {m[2]:8.0f} "
              f"state_full {m[3]:8.0f} slot_free/handoff {m[4]:8.0f} cluster_sync {m[5]:8.0f} | busy {m[0]-m[1]-m[2]-m[3]-m[4]-m[5]:8.0f} | phase1 ends at {m[6]:8.0f} | prep {s[:, rank, w, 7].div(1000000, rounding_mode='floor').mean():8.0f} post {s[:, rank, w, 7].remainder(1000000).mean():8.0f}")

if os.environ.get("TIMELINE"):
    print("stage | producer: wait_start issue | prep(pair of stage, half0): wait_start got_data done | chain: wait_start got_prep rows_done | post: wait_start got_state done")
    for k in list(range(0, 30)) + list(range(44, 70)) + list(range(94, 101)):
        print(f"{k:4d} | {tl[1,k,0]:7d} {tl[1,k,1]:7d} | {tl[2,k,0]:7d} {tl[2,k,1]:7d} {tl[2,k,2]:7d} | {tl[0,k,0]:7d} {tl[0,k,1]:7d} {tl[0,k,2]:7d} | {tl[3,k,0]:7d} {tl[3,k,1]:7d} {tl[3,k,2]:7d}")
