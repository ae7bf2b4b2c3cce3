"""Small invocations of every kernel family (a quick self-check on a GPU box; compute-sanitizer is closed on this pool); results are
compared with the oracle so that a silent corruption shows up too."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from bench import load_product
from lattice_util import make_inputs, ragged_lengths
import oracle
P = load_product(); oracle.build()
d = lambda x: torch.as_tensor(np.ascontiguousarray(x)).cuda()
ok = True
for kind, (B, T, U) in [(4, (2, 72, 64)), (5, (1, 40, 64)), (2, (2, 72, 64)), (3, (1, 40, 64)), (2, (2, 50, 20)), (1, (2, 40, 36)), (0, (2, 30, 33)), (4, (1, 70, 128))]:
    le, ls = make_inputs(B, T, U, seed=kind * 10 + T)
    t_len, u_len = ragged_lengths(B, T, U, seed=kind)
    want = oracle.forward_backward(le, ls, t_len, u_len)
    P.set_fb_kernel(kind)
    ll, loss, ge, gs = P.forward_backward(d(le), d(ls), d(t_len), d(u_len))
    torch.cuda.synchronize()
    err = max(float(np.abs(ge.cpu().numpy() - want[2]).max()), float(np.abs(gs.cpu().numpy() - want[3]).max()))
    good = err < 1e-4 and np.allclose(ll.cpu().numpy(), want[0], rtol=1e-5)
    ok &= bool(good)
    print("fb kind", kind, (B, T, U), "used", P.fb_kernel_used(), "max grad err %.2e" % err, "OK" if good else "MISMATCH", flush=True)
P.set_fb_kernel(-1)
le, ls, lt = make_inputs(2, 20, 8, seed=5, K=4)
w = oracle.tone_latent_forward_backward(le, ls, lt)
r = P.tone_latent_forward_backward(d(le), d(ls), d(lt)); torch.cuda.synchronize()
good = np.abs(r[2].cpu().numpy() - w[2]).max() < 1e-4; ok &= bool(good); print("tone fb", "OK" if good else "MISMATCH", flush=True)
rng = np.random.default_rng(0)
a = rng.integers(0, 5, (4, 40)).astype(np.int32); b = rng.integers(0, 5, (4, 40)).astype(np.int32)
al = rng.integers(0, 41, 4).astype(np.int32); bl = rng.integers(0, 41, 4).astype(np.int32)
good = np.array_equal(P.levenshtein_edit_distance(d(a), d(b), d(al), d(bl)).cpu().numpy(), oracle.levenshtein_edit_distance(a, b, al, bl))
ok &= bool(good); print("edit distance", "OK" if good else "MISMATCH", flush=True)
B, W, D = 3, 4, 6
h = np.log(rng.dirichlet(np.ones(D), (B, W))).astype(np.float32)
z = np.zeros((B, W), np.float32); zi = np.zeros((B, W), np.int32); zb = np.zeros((B, W), np.bool_)
tab = np.arange(D, dtype=np.int32); il = np.full(B, 5, np.int32); ol = np.full(B, 12, np.int32)
*o, bad = oracle.ssnt_tts_v2_beam_search_decode(h, z, zb, zi, tab, zi, zi, il, ol, W, D, 0, False, True)
g = P.ssnt_tts_v2_beam_search_decode(d(h), d(z), d(zb), d(zi), d(tab), d(zi), d(zi), d(il), d(ol), W, D, 0, False, True)
good = all(np.array_equal(np.asarray(x.cpu()), y) for x, y in zip(g, o)); ok &= bool(good); print("v2 beam step", "OK" if good else "MISMATCH", flush=True)
bb = rng.integers(0, W, (B, 9, W)).astype(np.int32); fin = np.tile(np.arange(W, dtype=np.int32), (B, 1))
good = np.array_equal(P.order_beam_branch(d(fin), d(bb), W).cpu().numpy(), oracle.order_beam_branch(fin, bb, W)); ok &= bool(good)
print("back-trace", "OK" if good else "MISMATCH", flush=True)
print("ALL OK" if ok else "SOME MISMATCH")
sys.exit(0 if ok else 1)
