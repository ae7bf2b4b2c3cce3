"""numpy fp32 emulation of the time-parallel (chunked transfer-operator) lattice algorithm of
csrc/fb_tp.cuh, checked against the fp64 oracle.  Development aid, not a test.

Chunk c covers frames [cL, cL+L).  P_c = M_{cL+L-1} ... M_{cL} (banded, bandwidth L+1) is built as
Q_c(i, d) = P_c(i, i-d); alpha_{c+1} = P_c alpha_c, beta_c = P_c^T beta_{c+1}; interiors are then filled
independently per chunk.  Boundary vectors carry one power-of-two exponent per lane (CPL tokens)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
f32 = np.float32

def probs(le, ls, T, U, L):
    max_t, max_u = le.shape
    C = (T + L - 1) // L
    e = np.zeros((C * L, max_u), f32); s = np.zeros((C * L, max_u), f32)
    e[:T, :U] = np.exp(le[:T, :U].astype(f32))
    s[:T - 1, :U - 1] = np.exp(ls[:T - 1, :U - 1].astype(f32))
    e[T:, :] = 1.0  # identity rows after the last frame
    return e, s, C

def build(e, s, c, L):
    U = e.shape[1]
    Q = np.zeros((L + 1, U), f32)   # Q[d, i] = P(i, i-d)
    Q[0, :] = 1.0
    for l in range(L):
        el, sl = e[c * L + l], s[c * L + l]
        new = np.zeros_like(Q)
        new[:, :] = el[None, :] * Q
        new[1:, 1:] += sl[None, :-1] * Q[:-1, :-1]
        Q = new.astype(f32)
    return Q

def pow2(x):
    return np.ldexp(f32(1.0), x).astype(f32)

def renorm(v, ex, CPL):
    """per-lane renormalisation of mantissas v (frame ex per lane) -> lane max in [1, 2)"""
    lanes = v.reshape(-1, CPL)
    m = lanes.max(axis=1)
    sh = np.where(m > 0, np.floor(np.log2(np.maximum(m, 1e-45))).astype(np.int64), 0)
    out = (lanes * pow2(-sh)[:, None]).astype(f32)
    nex = np.where(m > 0, ex + sh, -(1 << 20))
    return out.reshape(-1), nex

def combine_fwd(Q, v, ex, CPL, L):
    U = v.shape[0]; nl = U // CPL
    out = np.zeros(U, f32); oex = np.zeros(nl, np.int64)
    span = (L + CPL - 1) // CPL
    for l in range(nl):
        src = [k for k in range(max(0, l - span), l + 1)]
        eref = max(ex[k] for k in src)
        for r in range(CPL):
            i = l * CPL + r
            acc = f32(0)
            for d in range(L + 1):
                j = i - d
                if j < 0: break
                dd = ex[j // CPL] - eref
                a = f32(v[j] * pow2(dd)) if dd >= -126 else f32(0)
                acc = f32(acc + Q[d, i] * a)
            out[i] = acc
        oex[l] = eref
    return renorm(out, oex, CPL)

def combine_bwd(Q, v, ex, CPL, L):
    U = v.shape[0]; nl = U // CPL
    out = np.zeros(U, f32); oex = np.zeros(nl, np.int64)
    span = (L + CPL - 1) // CPL
    for l in range(nl):
        src = [k for k in range(l, min(nl, l + span + 1))]
        eref = max(ex[k] for k in src)
        for r in range(CPL):
            j = l * CPL + r
            acc = f32(0)
            for d in range(L + 1):
                i = j + d
                if i >= U: break
                dd = ex[i // CPL] - eref
                b = f32(v[i] * pow2(dd)) if dd >= -126 else f32(0)
                acc = f32(acc + Q[d, i] * b)
            out[j] = acc
        oex[l] = eref
    return renorm(out, oex, CPL)

def run(le, ls, T, U, L=16, CPL=4):
    max_t, max_u = le.shape
    e, s, C = probs(le, ls, T, U, L)
    Qs = [build(e, s, c, L) for c in range(C)]
    nl = max_u // CPL
    av = np.zeros(max_u, f32); av[0] = 1; aex = np.zeros(nl, np.int64); aex[1:] = -(1 << 20)
    A = [(av, aex)]
    for c in range(C):
        A.append(combine_fwd(Qs[c], *A[-1], CPL, L))
    bv = np.zeros(max_u, f32); bv[U - 1] = 1; bex = np.full(nl, -(1 << 20), np.int64); bex[(U - 1) // CPL] = 0
    Bv = [None] * (C + 1); Bv[C] = (bv, bex)
    for c in range(C - 1, -1, -1):
        Bv[c] = combine_bwd(Qs[c], *Bv[c + 1], CPL, L)
    # Z at every boundary, as (mantissa, exponent)
    def dot(a, b):
        (va, ea), (vb, eb) = a, b
        ex = ea + eb
        m = ex.max()
        terms = (va.reshape(-1, CPL) * vb.reshape(-1, CPL)).sum(axis=1).astype(np.float64)
        return np.log2(max((terms * np.exp2((ex - m).astype(np.float64))).sum(), 1e-300)) + m
    Z2 = np.array([dot(A[c], Bv[c]) for c in range(C + 1)])
    ll = Z2[0] * np.log(2.0)
    ge = np.zeros((max_t, max_u), f32); gs = np.zeros((max_t, max_u), f32)
    rowsum_dev = 0.0
    z0 = Z2[0]
    for c in range(C):
        (va, ea), (vb, eb) = A[c], Bv[c + 1]
        # fill frames, fixed over the chunk: F_l = max(ex_l, F_{l-1} - dec) (a lane the front has not reached yet
        # takes its neighbour's frame lowered by dec, so that the mass entering it neither overflows nor flushes)
        dec = 96 // ((L + CPL - 1) // CPL)
        fa = ea.copy()
        for l in range(1, nl): fa[l] = max(ea[l], fa[l - 1] - dec)
        fb = eb.copy()
        for l in range(nl - 2, -1, -1): fb[l] = max(eb[l], fb[l + 1] - dec)
        def p2(dd):
            return pow2(dd) if dd >= -126 else f32(0)
        a = np.zeros((L, max_u), f32)
        for u in range(max_u): a[0, u] = va[u] * p2(ea[u // CPL] - fa[u // CPL])
        ka = np.ones(max_u, f32)   # factor applied to a(u-1) when it enters token u
        for u in range(1, max_u): ka[u] = p2(fa[(u - 1) // CPL] - fa[u // CPL])
        for l in range(L - 1):
            el, sl = e[c * L + l], s[c * L + l]
            nxt = el * a[l]
            nxt[1:] += (sl[:-1] * a[l][:-1]) * ka[1:]
            a[l + 1] = nxt.astype(f32)
        kb = np.ones(max_u, f32)   # factor applied to b(u+1) when it enters token u
        for u in range(max_u - 1): kb[u] = p2(fb[(u + 1) // CPL] - fb[u // CPL])
        b = np.zeros(max_u, f32)
        for u in range(max_u): b[u] = vb[u] * p2(eb[u // CPL] - fb[u // CPL])
        sc = np.zeros(max_u, f32)
        for u in range(max_u):
            dd = fa[u // CPL] + fb[u // CPL] - z0
            sc[u] = f32(np.exp2(min(max(dd, -126.0), 126.0))) if dd > -140 else f32(0)
        for l in range(L - 1, -1, -1):
            t = c * L + l
            el, sl = e[t], s[t]
            p1 = (el * b).astype(f32)
            bn = np.zeros(max_u, f32); bn[:-1] = b[1:] * kb[:-1]
            p2 = (sl * bn).astype(f32)
            if t < T:
                g1 = (a[l] * sc) * p1; g2 = (a[l] * sc) * p2
                ge[t] = g1; gs[t] = g2
                rowsum_dev = max(rowsum_dev, abs(float(g1.sum() + g2.sum()) - 1.0))
            b = (p1 + p2).astype(f32)
    return ll, ge, gs, Z2, rowsum_dev

if __name__ == "__main__":
    import oracle
    from lattice_util import make_inputs, ragged_lengths
    oracle.build()
    for (B, T, U, seed) in [(2, 800, 128, 1), (2, 333, 128, 2), (2, 130, 128, 3), (2, 64, 64, 4)]:
        le, ls = make_inputs(B, T, U, seed=seed)
        t_len, u_len = ragged_lengths(B, T, U, seed=seed)
        t_len[0], u_len[0] = T, min(T, U)
        ll64, _, ge64, gs64 = oracle.forward_backward(le, ls, t_len, u_len)
        for b in range(B):
            ll, ge, gs, Z2, dev = run(le[b], ls[b], int(t_len[b]), int(u_len[b]), CPL=U // 32)
            print(f"T={t_len[b]} U={u_len[b]} ll {ll:.6f} vs {ll64[b]:.6f} rel {abs(ll-ll64[b])/abs(ll64[b]):.2e} "
                  f"ge err {np.abs(ge-ge64[b]).max():.2e} gs err {np.abs(gs-gs64[b]).max():.2e} "
                  f"Zspread {(Z2.max()-Z2.min())*np.log(2):.2e} rowdev {dev:.2e}")
