"""numpy fp32 emulation of the register-sweep time-parallel lattice kernels (csrc/fb_tp4.cuh, kind 8): groups of
L = 4 frames, one warp per sweep, per-lane power-of-two frames that lag the data by two steps.  Checked against the
fp64 oracle.  Development aid, not a test."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
f32 = np.float32
DEAD = -(1 << 20)
GUARD = 64

def pow2c(x):
    """2^x as float32, exponent clamped to [-127 -> 0, 127]"""
    x = np.asarray(x, np.int64)
    out = np.ldexp(f32(1.0), np.clip(x, -126, 127)).astype(f32)
    return np.where(x < -126, f32(0), out).astype(f32)

def probs(le, ls, T, U, L):
    max_t, max_u = le.shape
    C = (T + L - 1) // L
    e = np.zeros((C * L, max_u), f32); s = np.zeros((C * L, max_u), f32)
    e[:T, :U] = np.exp(le[:T, :U].astype(f32))
    s[:T - 1, :U - 1] = np.exp(ls[:T - 1, :U - 1].astype(f32))
    e[T:, :] = 1.0
    return e, s, C

def build(e, s, g, L):
    U = e.shape[1]
    Q = np.zeros((L + 1, U), f32); Q[0, :] = 1.0
    for l in range(L):
        el, sl = e[g * L + l], s[g * L + l]
        new = (el[None, :] * Q).astype(f32)
        new[1:, 1:] = (new[1:, 1:] + sl[None, :-1] * Q[:-1, :-1]).astype(f32)
        Q = new
    return Q

def lane_exp(y, CPL):
    m = y.reshape(-1, CPL).max(axis=1)
    alive = m > 0
    ex = np.where(alive, np.floor(np.log2(np.maximum(m, 1e-45))).astype(np.int64), 0)
    return alive, ex

def sweep(Qs, U, max_u, CPL, L, direction):
    """returns list of (y, F) boundary vectors in sweep order; frames lag two steps."""
    nl = max_u // CPL
    G = len(Qs)
    y = np.zeros(max_u, f32); y[0 if direction == 0 else U - 1] = 1
    F = np.zeros(nl, np.int64)          # frame of the current vector y_{k-1}
    Fn = np.zeros(nl, np.int64)         # frame of y_k (decided one step ago)
    out = [(y.copy(), F.copy())]
    for k in range(G):
        Q = Qs[k] if direction == 0 else Qs[G - 1 - k]
        # ---- off the chain: frame of y_{k+1} from y_{k-1} ----
        alive, ex = lane_exp(y, CPL)
        A = np.where(alive, F + ex, DEAD)
        if direction == 0:
            A1 = np.concatenate([[DEAD], A[:-1]]); A2 = np.concatenate([[DEAD, DEAD], A[:-2]])
        else:
            A1 = np.concatenate([A[1:], [DEAD]]); A2 = np.concatenate([A[2:], [DEAD, DEAD]])
        # own largest exponent, but at most GUARD below the two upstream lanes' (whatever can arrive within the two
        # steps of lag came from there, and a step grows a value by at most 2^4)
        F2 = np.maximum(A, np.maximum(A1, A2) - GUARD)
        F2 = np.where(F2 > DEAD // 2, F2, Fn)
        F2 = Fn + np.clip(F2 - Fn, -126, 126)
        # ---- the step: y_k = c * P y_{k-1}, incoming scaled by kin ----
        c = pow2c(-(Fn - F))                      # per lane
        if direction == 0:
            Fnb = np.concatenate([[DEAD], F[:-1]])
        else:
            Fnb = np.concatenate([F[1:], [DEAD]])
        kin = pow2c(Fnb - F)
        new = np.zeros(max_u, f32)
        for l in range(nl):
            for r in range(CPL):
                i = l * CPL + r
                acc = f32(0)
                for d in range(L + 1):
                    if direction == 0:
                        j = i - d
                        if j < 0: break
                        q = Q[d, i]
                    else:
                        j = i + d
                        if j >= max_u: break
                        q = Q[d, j]
                    x = y[j]
                    if j // CPL != l:
                        assert abs(j // CPL - l) == 1
                        x = f32(x * kin[l])
                    acc = f32(acc + q * x)
                new[i] = f32(acc * c[l])
        y = new
        F, Fn = Fn, F2
        out.append((y.copy(), F.copy()))
    return out

def run(le, ls, T, U, L=4, CPL=4):
    max_t, max_u = le.shape
    e, s, G = probs(le, ls, T, U, L)
    Qs = [build(e, s, g, L) for g in range(G)]
    nl = max_u // CPL
    A = sweep(Qs, U, max_u, CPL, L, 0)
    Bs = sweep(Qs, U, max_u, CPL, L, 1)
    Bv = Bs[::-1]   # Bv[g] = beta at boundary g
    ya, Fa = A[G]; zf = np.log2(max(float(ya[U - 1]), 1e-300)) + Fa[(U - 1) // CPL]
    yb, Fb = Bv[0]; zb = np.log2(max(float(yb[0]), 1e-300)) + Fb[0]
    ll = zf * np.log(2.0)
    ge = np.zeros((max_t, max_u), f32); gs = np.zeros((max_t, max_u), f32)
    rowsum_dev = 0.0
    for g in range(G):
        (va, ea0), (vb, eb0) = A[g], Bv[g + 1]
        # per-lane renormalisation, then frames for the group: own exponent, but not more than GUARD below the neighbour's
        al, xa = lane_exp(va, CPL); bl, xb = lane_exp(vb, CPL)
        ea = np.where(al, ea0 + xa, DEAD); eb = np.where(bl, eb0 + xb, DEAD)
        va = (va.reshape(nl, CPL) * pow2c(-xa)[:, None]).astype(f32).reshape(-1)
        vb = (vb.reshape(nl, CPL) * pow2c(-xb)[:, None]).astype(f32).reshape(-1)
        fa = np.maximum(ea, np.concatenate([[DEAD], ea[:-1]]) - GUARD)
        fb = np.maximum(eb, np.concatenate([eb[1:], [DEAD]]) - GUARD)
        a = np.zeros((L, max_u), f32)
        a[0] = (va.reshape(nl, CPL) * pow2c(ea - fa)[:, None]).reshape(-1)
        ka = np.repeat(pow2c(np.concatenate([[DEAD], fa[:-1]]) - fa), CPL)
        ka[np.arange(max_u) % CPL != 0] = 1
        for l in range(L - 1):
            el, sl = e[g * L + l], s[g * L + l]
            nxt = el * a[l]
            nxt[1:] += (sl[:-1] * a[l][:-1]) * ka[1:]
            a[l + 1] = nxt.astype(f32)
        kb = np.repeat(pow2c(np.concatenate([fb[1:], [DEAD]]) - fb), CPL)
        kb[np.arange(max_u) % CPL != CPL - 1] = 1
        b = (vb.reshape(nl, CPL) * pow2c(eb - fb)[:, None]).reshape(-1).astype(f32)
        dd = (fa + fb).astype(np.float64) - zf
        sc = np.repeat(np.where(dd > -140, np.exp2(np.clip(dd, -126, 126)), 0).astype(f32), CPL)
        for l in range(L - 1, -1, -1):
            t = g * L + l
            el, sl = e[t], s[t]
            p1 = (el * b).astype(f32)
            bn = np.zeros(max_u, f32); bn[:-1] = b[1:] * kb[:-1]
            p2 = (sl * bn).astype(f32)
            if t < T:
                g1 = (a[l] * sc) * p1; g2 = (a[l] * sc) * p2
                ge[t] = g1; gs[t] = g2
                rowsum_dev = max(rowsum_dev, abs(float(g1.sum() + g2.sum()) - 1.0))
            b = (p1 + p2).astype(f32)
    return ll, ge, gs, zf - zb, rowsum_dev

if __name__ == "__main__":
    import oracle
    from lattice_util import make_inputs, ragged_lengths
    oracle.build()
    cases = [(2, 800, 128, 1), (2, 333, 128, 2), (2, 130, 128, 3), (1, 600, 256, 5)]
    if len(sys.argv) > 1:
        cases = [tuple(int(x) for x in sys.argv[1:5])]
    for (B, T, U, seed) in cases:
        le, ls = make_inputs(B, T, U, seed=seed)
        t_len, u_len = ragged_lengths(B, T, U, seed=seed)
        t_len[0], u_len[0] = T, min(T, U)
        ll64, _, ge64, gs64 = oracle.forward_backward(le, ls, t_len, u_len)
        for b in range(B):
            ll, ge, gs, zd, dev = run(le[b], ls[b], int(t_len[b]), int(u_len[b]), CPL=U // 32)
            rel = np.abs(ge - ge64[b]) / np.maximum(ge64[b], 1e-6)
            print(f"T={t_len[b]} U={u_len[b]} ll {ll:.6f} vs {ll64[b]:.6f} rel {abs(ll-ll64[b])/abs(ll64[b]):.2e} "
                  f"ge err {np.abs(ge-ge64[b]).max():.2e} gs err {np.abs(gs-gs64[b]).max():.2e} ge relerr {rel.max():.2e} "
                  f"zdiff {zd:.2e} rowdev {dev:.2e}", flush=True)
