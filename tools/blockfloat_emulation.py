"""CPU emulation (numpy float32) of the block-floating-point lattice recursion used by
fb_bf_kernel: probabilities instead of log-probabilities, one shared power-of-two exponent per
lane (CPL consecutive tokens), re-normalised every G rows.  Used to validate the numerics and the
range-management rules against the fp64 oracle before running on the GPU.  Not part of the
product or of the test-suite."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import numpy as np

f32 = np.float32
TARGET, SLACK, G = 24, 32, 8
BIG = -100000


def ilogb(x):
    """floor(log2 x) from the exponent field (x > 0), like the kernel's bit extraction."""
    b = np.asarray(x, f32).view(np.int32)
    return ((b >> 23) & 0xFF).astype(np.int64) - 127


def pow2(k):
    k = np.clip(k, -149, 127)
    return np.ldexp(f32(1.0), k.astype(np.int64)).astype(f32)


def scale_pow2(v, k):
    """v * 2^k in two fp32 multiplies (k may exceed the single-factor range)."""
    k = np.clip(k, -290, 290)
    k1 = k // 2
    return (v * pow2(k1)[..., None]).astype(f32) * pow2(k - k1)[..., None]


PIPELINED = True
DEC_AT = 6   # the decision is taken from the state before this row of the stage, applied at the next stage start


def sweep(e, s, U, CPL, direction):
    """Runs one full sweep (all T rows) and returns the list of (v[lanes,CPL], exp[lanes]) per row
    BEFORE the step of that row.  direction=+1: alpha from row 0; -1: beta from row T-1 (state of
    row t is beta(t+1))."""
    T = e.shape[0]
    L = (U + CPL - 1) // CPL
    Upad = L * CPL
    ep = np.zeros((T, Upad), f32); ep[:, :U] = e
    sp = np.zeros((T, Upad), f32); sp[:, :U] = s
    sp[:, U - 1] = 0; sp[T - 1, :] = 0
    v = np.zeros((L, CPL), f32)
    ex = np.zeros(L, np.int64)
    if direction > 0:
        v[0, 0] = 1
    else:
        v[(U - 1) // CPL, (U - 1) % CPL] = 1
    g = np.ones(L, f32)
    if direction > 0: g[0] = 0
    else: g[-1] = 0
    dec = None
    rows = []
    order = range(T) if direction > 0 else range(T - 1, -1, -1)
    for j, t in enumerate(order):
        if j % G == 0 and PIPELINED:
            if dec is not None:
                new = dec
                v = scale_pow2(v, ex - new)
                ex = new
                d = np.zeros(L, np.int64)
                if direction > 0: d[1:] = ex[:-1] - ex[1:]
                else: d[:-1] = ex[1:] - ex[:-1]
                g = pow2(np.clip(d, -126, 126))
                if direction > 0: g[0] = 0
                else: g[-1] = 0
        if j % G == DEC_AT and PIPELINED:
            mloc = v.max(axis=1)
            own = np.where(mloc > 0, ex + ilogb(np.maximum(mloc, f32(1e-45))) - TARGET, BIG)
            edge = v[:, -1] if direction > 0 else v[:, 0]
            amag = np.where(edge > 0, ex + ilogb(np.maximum(edge, f32(1e-45))), BIG)
            nb = np.full(L, BIG, np.int64)
            if direction > 0: nb[1:] = amag[:-1]
            else: nb[:-1] = amag[1:]
            new = np.maximum(own, nb - TARGET - SLACK)
            dec = np.where(new <= BIG // 2, ex, new)
        if j % G == 0 and not PIPELINED:
            # ---- renorm ----
            mloc = v.max(axis=1)
            own = np.where(mloc > 0, ex + ilogb(np.maximum(mloc, f32(1e-45))) - TARGET, BIG)
            edge = v[:, -1] if direction > 0 else v[:, 0]
            amag = np.where(edge > 0, ex + ilogb(np.maximum(edge, f32(1e-45))), BIG)
            nb = np.full(L, BIG, np.int64)
            if direction > 0:
                nb[1:] = amag[:-1]
            else:
                nb[:-1] = amag[1:]
            new = np.maximum(own, nb - TARGET - SLACK)
            new = np.where(new <= BIG // 2, ex, new)
            v = scale_pow2(v, ex - new)
            ex = new
            d = np.zeros(L, np.int64)
            if direction > 0:
                d[1:] = ex[:-1] - ex[1:]
            else:
                d[:-1] = ex[1:] - ex[:-1]
            g = pow2(np.clip(d, -126, 126))
            if direction > 0: g[0] = 0
            else: g[-1] = 0
        rows.append((t, v.copy(), ex.copy()))
        E = ep[t].reshape(L, CPL); S = sp[t].reshape(L, CPL)
        if direction > 0:
            b = (v * S).astype(f32)
            inc = np.zeros(L, f32); inc[1:] = b[:-1, -1]
            nv = np.empty_like(v)
            nv[:, 1:] = (v[:, 1:] * E[:, 1:] + b[:, :-1]).astype(f32)
            nv[:, 0] = (inc * g + (v[:, 0] * E[:, 0]).astype(f32)).astype(f32)
        else:
            # beta(t,u) = e*beta(t+1,u) + s*beta(t+1,u+1)
            inc = np.zeros(L, f32); inc[:-1] = v[1:, 0]
            nbv = np.empty_like(v)
            nbv[:, :-1] = v[:, 1:]
            nbv[:, -1] = (inc * g).astype(f32)
            nv = (E * v + (S * nbv).astype(f32)).astype(f32)
        v = nv
    rows.append((None, v.copy(), ex.copy()))  # final state (alpha(T) unused / beta(0))
    return rows


def forward_backward_bf(le, ls, CPL=4):
    T, U = le.shape
    e = np.exp(le.astype(np.float64)).astype(f32); s = np.exp(ls.astype(np.float64)).astype(f32)
    A = sweep(e, s, U, CPL, +1)     # A[t] = alpha(t)
    Bw = sweep(e, s, U, CPL, -1)    # Bw[j]: state before step of row t=T-1-j is beta(t+1)
    beta = {}
    for (t, v, ex) in Bw[:-1]:
        beta[t + 1] = (v, ex)
    beta[0] = (Bw[-1][1], Bw[-1][2])
    L = A[0][1].shape[0]
    ge = np.zeros((T, U)); gs = np.zeros((T, U))
    m = (T + 1) // 2
    # Z at the meeting row m-1 in (mantissa, exponent) form
    def terms(t):
        va, ea = A[t][1].astype(np.float64), A[t][2]
        vb, eb = beta[t + 1][0].astype(np.float64), beta[t + 1][1]
        ep = np.zeros(L * CPL); ep[:U] = e[t]
        sp = np.zeros(L * CPL); sp[:U] = s[t]; sp[U - 1] = 0
        if t == T - 1: sp[:] = 0
        E = ep.reshape(L, CPL); S = sp.reshape(L, CPL)
        vbn = np.empty_like(vb); vbn[:, :-1] = vb[:, 1:]
        nxt = np.zeros(L); nxt[:-1] = vb[1:, 0] * np.ldexp(1.0, np.clip(eb[1:] - eb[:-1], -1000, 1000))
        vbn[:, -1] = nxt
        return va * E * vb, va * S * vbn, ea + eb
    ge_m, gs_m, Em = terms(m - 1)
    w = (ge_m + gs_m).sum(axis=1)
    M = np.max(np.where(w > 0, Em + np.floor(np.log2(np.maximum(w, 1e-300))), BIG))
    tot = (w * np.ldexp(1.0, np.clip(Em - M, -1000, 1000).astype(np.int64))).sum()
    ll = (np.log2(tot) + M) * np.log(2.0)
    for t in range(T):
        a, b, Et = terms(t)
        F = np.ldexp(1.0, np.clip(Et - M, -1000, 1000).astype(np.int64)) / tot
        ge[t] = (a * F[:, None]).reshape(-1)[:U]
        gs[t] = (b * F[:, None]).reshape(-1)[:U]
    return ll, ge, gs


if __name__ == "__main__":
    import oracle
    from lattice_util import make_inputs
    cases = {}
    le, ls = make_inputs(1, 800, 128, seed=1234); cases["cfg2 random"] = (le[0], ls[0])
    T, U = 400, 64
    cases["uniform"] = (np.full((T, U), np.log(0.5), f32), np.full((T, U), np.log(0.5), f32))
    rng = np.random.default_rng(3)
    cases["peaked random -30"] = ((-30.0 * (rng.random((T, U)) < 0.5) - 1e-3).astype(f32),
                                  (-30.0 * (rng.random((T, U)) < 0.5) - 1e-3).astype(f32))
    cases["always shift"] = (np.full((T, U), np.log(1e-10), f32), np.full((T, U), np.log1p(-1e-10), f32))
    cases["never shift"] = (np.full((T, U), np.log1p(-1e-6), f32), np.full((T, U), np.log(1e-6), f32))
    le, ls = make_inputs(1, 2000, 256, seed=5); cases["cfg5 random CPL8"] = (le[0], ls[0])
    for name, (le, ls) in cases.items():
        cpl = 8 if "CPL8" in name else 4
        ll64, _, ge64, gs64 = oracle.forward_backward(le[None], ls[None], precision="f64")
        with np.errstate(all="ignore"):
            ll, ge, gs = forward_backward_bf(le, ls, cpl)
        print(f"{name:22s} ll {ll:14.6f} oracle {ll64[0]:14.6f} rel {abs(ll-ll64[0])/abs(ll64[0]):.2e} "
              f"grad max abs err {max(np.abs(ge-ge64[0]).max(), np.abs(gs-gs64[0]).max()):.2e} "
              f"rowsum dev {np.abs((ge+gs).sum(1)-1).max():.2e}")
