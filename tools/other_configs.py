"""Timings of the non-headline BASELINE configs (3: tone-latent lattice, 4: beam decode + edit distance)
on the GPU with the CPU oracle beside them (reporting aid, needs a GPU)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from bench import load_product
import oracle
P = load_product(); oracle.build()
dev = torch.device("cuda")

def gpu_time(fn, reps=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def graph_time(fn, per_graph=20, replays=10):
    """GPU time per call with the host out of the way: the calls are captured into a CUDA graph (the
    Python wrapper costs ~60 us per call, far more than these small kernels)."""
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(per_graph): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (per_graph * replays)

def cpu_time(fn, reps=3):
    fn(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    return (time.perf_counter() - t0) / reps * 1e3

# ---- config 3: tone-latent lattice B=32 U=128 T=800 K=4
B, T, U, K = 32, 800, 128, 4
g = torch.Generator(device="cuda").manual_seed(3)
z = torch.randn(B, T, U, K, device=dev, generator=g)
le, ls = torch.nn.functional.logsigmoid(z), torch.nn.functional.logsigmoid(-z)
lt = torch.log_softmax(torch.randn(B, U, K, device=dev, generator=g), dim=-1)
ms = gpu_time(lambda: P.tone_latent_forward_backward(le, ls, lt), reps=5)
cells = B * T * U
len_, lsn, ltn = le.cpu().numpy(), ls.cpu().numpy(), lt.cpu().numpy()
cms = cpu_time(lambda: oracle.tone_latent_forward_backward(len_, lsn, ltn, precision="f32"), reps=3)
print(f"cfg3 tone-latent lattice B={B} U={U} T={T} K={K}: GPU {ms*1e3:.0f} us = {cells/ms/1e6:.2f} Gcells/s "
      f"({64*cells/ms/1e6:.0f} GB/s algorithmic, {64*cells/ms/1e6/6548.8*100:.1f}% of HBM roofline) | CPU oracle {cms:.0f} ms ({oracle.get_threads()} threads)")

# ---- config 4: v2 beam step B=64 W=8 D=16, edit distance B=64 len 1000 / 150
B, W, D = 64, 8, 16
rng = np.random.default_rng(4)
h = torch.log_softmax(torch.randn(B, W, D, device=dev), dim=-1)
st = [torch.zeros(B, W, device=dev), torch.zeros(B, W, dtype=torch.bool, device=dev), torch.zeros(B, W, dtype=torch.int32, device=dev)]
tab = torch.arange(D, dtype=torch.int32, device=dev)
tt = torch.zeros(B, W, dtype=torch.int32, device=dev); uu = torch.zeros(B, W, dtype=torch.int32, device=dev)
il = torch.full((B,), 150, dtype=torch.int32, device=dev); ol = torch.full((B,), 1000, dtype=torch.int32, device=dev)
ms = gpu_time(lambda: P.ssnt_tts_v2_beam_search_decode(h, st[0], st[1], st[2], tab, tt, uu, il, ol, W, D, 0, False, True), reps=200)
hn = h.cpu().numpy(); s0 = [x.cpu().numpy() for x in st]
cms = cpu_time(lambda: oracle.ssnt_tts_v2_beam_search_decode(hn, s0[0], s0[1], s0[2], tab.cpu().numpy(), tt.cpu().numpy(), uu.cpu().numpy(), il.cpu().numpy(), ol.cpu().numpy(), W, D, 0, False, True), reps=50)
gms = graph_time(lambda: P.ssnt_tts_v2_beam_search_decode(h, st[0], st[1], st[2], tab, tt, uu, il, ol, W, D, 0, False, True))
print(f"cfg4 v2 beam step B={B} W={W} D={D} (device pointers): GPU {ms*1e3:.1f} us/step through the Python wrapper, {gms*1e3:.1f} us/step as a CUDA graph (pre-fill + kernel) | CPU oracle {cms*1e3:.1f} us/step")
hk = torch.log_softmax(torch.randn(B, W, 4, device=dev), dim=-1)
ms = gpu_time(lambda: P.tone_latent_beam_search_decode(hk, st[0], st[1], tt, uu, il, W, 4, 0), reps=200)
gms = graph_time(lambda: P.tone_latent_beam_search_decode(hk, st[0], st[1], tt, uu, il, W, 4, 0))
print(f"cfg4 tone-latent beam step B={B} W={W} K=4: GPU {ms*1e3:.1f} us/step through the Python wrapper, {gms*1e3:.1f} us/step as a CUDA graph")
for L in (150, 1000):
    a = torch.randint(0, 50, (B, L), dtype=torch.int32, device=dev); b = torch.randint(0, 50, (B, L), dtype=torch.int32, device=dev)
    al = torch.full((B,), L, dtype=torch.int32, device=dev)
    ms = gpu_time(lambda: P.levenshtein_edit_distance(a, b, al, al), reps=50)
    an, bn, aln = a.cpu().numpy(), b.cpu().numpy(), al.cpu().numpy()
    cms = cpu_time(lambda: oracle.levenshtein_edit_distance(an, bn, aln, aln), reps=5)
    print(f"cfg4 edit distance B={B} L={L}: GPU {ms*1e3:.1f} us = {B*L*L/ms/1e6:.2f} G cell updates/s | CPU oracle {cms:.2f} ms = {B*L*L/cms/1e6:.2f} G/s")
T_, Wd = 1000, 8
bb = torch.randint(0, Wd, (B, T_, Wd), dtype=torch.int32, device=dev); fin = torch.arange(Wd, dtype=torch.int32, device=dev).repeat(B, 1)
ms = gpu_time(lambda: P.order_beam_branch(fin, bb, Wd), reps=50)
print(f"cfg4 back-trace of all beams B={B} T={T_} W={Wd}: GPU {ms*1e3:.1f} us")
